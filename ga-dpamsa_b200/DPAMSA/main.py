"""Plain DPAMSA inference (no GA) with the reference's entry point (reference DPAMSA/main.py:238-310):
one greedy episode of the agent over every whole sequence set of a dataset module, metrics and the
report / CSV in the shared format. The Q-network runs on the GPU (DQN.predict -> gad_qnet_forward_host),
the single environment is host-side bookkeeping (DPAMSA/env.py). Training (main.py:95-235) is outside the
generation-loop scope of this build."""
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


def inference(dataset=None, start=0, end=-1, model_path=INFERENCE_MODEL, truncate_file=True):
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
    for name in dataset.datasets[start:end if end != -1 else len(dataset.datasets)]:
        if not hasattr(dataset, name):
            continue
        env = Environment(getattr(dataset, name))
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


def train(*_args, **_kwargs):
    raise NotImplementedError("DQN training (reference DPAMSA/main.py:95-235) is outside the generation-loop scope")
