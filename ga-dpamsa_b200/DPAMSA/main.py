"""DPAMSA entry points with the reference's names (reference DPAMSA/main.py).

``inference`` (main.py:238-310): one greedy episode of the agent over every whole sequence set of a dataset module,
metrics and the report / CSV in the shared format. The Q-network runs on the GPU (DQN.predict ->
gad_qnet_forward_host), the single environment is host-side bookkeeping (DPAMSA/env.py).

``train`` (main.py:95-235): the reference's DQN training loop - epsilon-greedy episodes,
replay memory, one optimisation step per environment step, epsilon decay per episode, early stopping on the
100-episode moving average, scalars to TensorBoard, weights saved in the reference layout - on the torch-autograd
network of DPAMSA/trainer.py. It does not start a TensorBoard server or a browser (main.py:62-93 does).
"""
from __future__ import annotations

import csv
import os

import config
import utils
from DPAMSA.dqn import DQN
from DPAMSA.env import Environment

INFERENCE_MODEL = 'model_3x30'


def output_parameters():
    print("\n---------- DPAMSA inference ----------")
    print(f"Scores: match {config.MATCH_REWARD}, mismatch {config.MISMATCH_PENALTY}, gap {config.GAP_PENALTY}\n")


def _device_episodes(groups, model_path):
    """Greedy full-board episodes of many sequence sets at once: every set whose
    sequences all have the same length is a board whose agent window is the whole board, so the lock-step device
    episodes of the mutation phase (gad_mutate: Environment.reset / step / padding, DQN.predict) align all sets of one
    shape in a single call. groups: {(rows, cols): [(name, seqs), ...]}. Returns {name: aligned rows as code lists}."""
    import torch
    from DPAMSA.dqn import load_qnet
    from DPAMSA.env import nucleotides_map
    from population import DevicePopulation
    out = {}
    for (rows, cols), items in groups.items():
        boards = [[[nucleotides_map[ch] for ch in seq] for seq in seqs] for _, seqs in items]
        qnet = load_qnet(model_path, rows, cols)
        pop = DevicePopulation.from_lists(boards, device=config.DEVICE)
        n = len(boards)
        idx = torch.arange(n, dtype=torch.int32, device=pop.device)
        tile = torch.zeros((n, 2), dtype=torch.int32, device=pop.device)
        max_reward = rows * (rows - 1) / 2 * config.MATCH_REWARD                   # env.py:79
        pop.mutate(qnet, idx, tile, cols * max_reward, width=cols)
        for (name, _), aligned in zip(items, pop.to_lists()):
            out[name] = aligned
    return out


def inference(dataset=None, start=0, end=-1, model_path=INFERENCE_MODEL, truncate_file=True, batched=True):
    """main.py:238-310. ``batched`` (default) aligns all equal-length sets of one shape in one lock-step device run;
    sets with sequences of different lengths, and ``batched=False``, take the reference's one-step-at-a-time loop.
    Both give the same alignments."""
    if dataset is None:
        raise ValueError("no dataset module given")
    output_parameters()
    tag = os.path.splitext(dataset.file_name)[0]
    os.makedirs(config.DPAMSA_REPORTS_PATH, exist_ok=True)
    os.makedirs(config.DPAMSA_INF_CSV_PATH, exist_ok=True)
    report_file_name = os.path.join(config.DPAMSA_REPORTS_PATH, f"{tag}.txt")
    csv_file_name = os.path.join(config.DPAMSA_INF_CSV_PATH, f"{tag}_DPAMSA_results.csv")
    if truncate_file:
        open(report_file_name, 'w').close()
        with open(csv_file_name, 'w', newline='') as fh:
            csv.writer(fh).writerow(utils.CSV_COLUMNS_QTY_FIRST)
    names = [n for n in dataset.datasets[start:end if end != -1 else len(dataset.datasets)] if hasattr(dataset, n)]
    done_on_device = {}
    if batched:
        groups = {}
        for name in names:
            seqs = getattr(dataset, name)
            lengths = {len(seq) for seq in seqs}
            if len(seqs) >= 1 and len(lengths) == 1 and min(lengths) >= 1:
                groups.setdefault((len(seqs), lengths.pop()), []).append((name, seqs))
        done_on_device = _device_episodes(groups, model_path)
    for name in names:
        env = Environment(getattr(dataset, name))
        if name in done_on_device:
            env.aligned = done_on_device[name]
            env.not_aligned = [[] for _ in range(env.row)]
        else:
            agent = DQN(env.action_number, env.row, env.max_len, env.max_len * env.max_reward)
            agent.load(model_path)
            state = env.reset()
            while True:
                _, state, done = env.step(agent.predict(state))
                if done == 0:
                    break
            env.padding()
        m = utils.calculate_metrics(env)
        with open(report_file_name, 'a') as fh:
            fh.write(utils.report_block(name, m, env.get_alignment()))
        with open(csv_file_name, 'a', newline='') as fh:
            csv.writer(fh).writerow([name, m['QTY'], m['AL'], m['SP'], m['EM'], m['CS']])
    print(f"\nInference completed successfully.\nReport saved at: {report_file_name}\nCSV saved at: {csv_file_name}\n")
    return report_file_name, csv_file_name


class _Scalars:
    """TensorBoard scalars under ``<RUNS_PATH>/<dataset file>/<set name>`` (main.py:108-121); when the tensorboard package
    is missing the same scalars go to ``scalars.csv`` in that directory."""

    def __init__(self, log_dir):
        os.makedirs(log_dir, exist_ok=True)
        self.tb = self.csv = None
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.tb = SummaryWriter(log_dir)
        except Exception:
            self.csv = open(os.path.join(log_dir, "scalars.csv"), "w", newline="")
            csv.writer(self.csv).writerow(["tag", "value", "step"])

    def add(self, tag, value, step):
        if self.tb is not None:
            self.tb.add_scalar(tag, value, step)
        else:
            csv.writer(self.csv).writerow([tag, value, step])

    def flush(self):
        (self.tb or self.csv).flush()

    def close(self):
        (self.tb or self.csv).close()


def output_training_parameters():
    """main.py:42-60"""
    print("\n---------- DPAMSA training ----------")
    for label, value in (("Gap penalty", config.GAP_PENALTY), ("Mismatch penalty", config.MISMATCH_PENALTY),
                         ("Match reward", config.MATCH_REWARD), ("Episode", config.MAX_EPISODE),
                         ("Batch size", config.BATCH_SIZE), ("Replay memory size", config.REPLAY_MEMORY_SIZE),
                         ("Alpha", config.ALPHA), ("Epsilon", config.EPSILON), ("Gamma", config.GAMMA),
                         ("Delta", config.DELTA), ("Decrement iteration", config.DECREMENT_ITERATION),
                         ("Update iteration", config.UPDATE_ITERATION), ("Device", config.DEVICE_NAME)):
        print("{}: {}".format(label, value))


def train(dataset=None, start=0, end=-1, model_path='new_model_3x30', early_stopping_patience=200):
    """Train on every sequence set ``dataset.datasets[start:end]`` in turn, saving ``<model_path>.pth`` after each
    (main.py:95-235). Returns the per-set episode logs (list of dicts) - the reference returns nothing."""
    if dataset is None:
        raise ValueError("no dataset module given")
    output_training_parameters()
    log_dir = os.path.join(config.RUNS_PATH, os.path.splitext(dataset.file_name)[0])
    writers = {}
    for name in dataset.datasets:
        w = writers[name] = _Scalars(os.path.join(log_dir, name))
        for tag, v in (("Training/Loss", 0), ("Training/Reward", 0), ("Training/Steps", 0),
                       ("Training/Epsilon", config.EPSILON), ("Metrics/SP", 0), ("Metrics/CS", 0)):
            w.add(tag, v, 0)
        w.flush()
    history = {}
    for name in dataset.datasets[start:end if end != -1 else len(dataset.datasets)]:
        if not hasattr(dataset, name):
            continue
        env = Environment(getattr(dataset, name))
        agent = DQN(env.action_number, env.row, env.max_len, env.max_len * env.max_reward)
        try:
            agent.load(model_path + '.pth')       # as in the reference (main.py:141-144): looks for <model_path>.pth.pth
        except Exception:
            pass
        best_avg_reward = -float('inf')
        no_improve = 0
        recent = []
        log = history[name] = []
        for episode in range(config.MAX_EPISODE):
            state = env.reset()
            episode_reward = episode_loss = steps = 0
            while True:
                action = agent.select(state)
                reward, next_state, done = env.step(action)
                agent.replay_memory.push((state, next_state, action, reward, done))
                loss = agent.update()
                episode_reward += reward
                if loss is not None:
                    episode_loss += loss
                steps += 1
                if done == 0:
                    break
                state = next_state
            agent.update_epsilon()
            sp_score = env.calc_score()
            column_score = env.calc_exact_matched() / len(env.aligned[0])
            w = writers[name]
            for tag, v in (("Training/Loss", episode_loss), ("Training/Reward", episode_reward), ("Training/Steps", steps),
                           ("Training/Epsilon", agent.current_epsilon), ("Metrics/SP", sp_score),
                           ("Metrics/CS", column_score)):
                w.add(tag, v, episode)
            log.append({"episode": episode, "loss": episode_loss, "reward": episode_reward, "steps": steps,
                        "epsilon": agent.current_epsilon, "sp": sp_score, "cs": column_score})
            recent.append(episode_reward)
            if len(recent) > 100:                 # moving average over the last 100 episodes (main.py:207-222)
                recent.pop(0)
                avg = sum(recent) / len(recent)
                if avg > best_avg_reward:
                    best_avg_reward = avg
                    no_improve = 0
                else:
                    no_improve += 1
                if no_improve >= early_stopping_patience:
                    print(f"Early stopping activated. No improvement for {early_stopping_patience} episodes.")
                    break
        writers[name].close()
        agent.save(model_path)
    print("\nTraining completed successfully.")
    return history
