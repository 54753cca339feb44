"""Training path against vectors recorded from the UNMODIFIED reference trainer
(oracle/make_golden_train.py -> tests/golden/train_vectors.json): replay memory bookkeeping incl. the overwrite rule,
seeded sampling, epsilon-greedy selection, three optimisation steps with dropout active (loss and every tensor of
the evaluation / target networks), the epsilon schedule. Runs on CPU torch; no CUDA library call."""
import json
import os
import random

import numpy as np
import pytest
import torch

import ga_oracle as O
from conftest import GOLDEN


@pytest.fixture(scope="module")
def vec():
    with open(os.path.join(GOLDEN, "train_vectors.json")) as fh:
        return json.load(fh)


@pytest.fixture()
def agent(vec, monkeypatch):
    import config
    for k, v in vec["knobs"].items():
        monkeypatch.setattr(config, k, v)
    from DPAMSA.dqn import DQN
    torch.set_num_threads(1)
    a = DQN(2 ** vec["rows"] - 1, vec["rows"], vec["cols"], vec["max_value"])
    a.train_device = "cpu"
    a._ensure_training()
    a.eval_net.load_state_dict({k: torch.from_numpy(v) for k, v in
                                O.synth_qnet_weights(vec["rows"], vec["cols"], vec["weight_seed"]).items()})
    a.target_net.load_state_dict({k: torch.from_numpy(v) for k, v in
                                  O.synth_qnet_weights(vec["rows"], vec["cols"], vec["weight_seed"] + 1).items()})
    return a


def digest(t, at):
    a = t.detach().cpu().double().numpy().ravel()
    return float(a.sum()), float(np.abs(a).sum()), a[at].tolist()


def check_digests(state_dict, want):
    assert set(state_dict) == set(want)
    for k, w in want.items():
        s, ab, vals = digest(state_dict[k], w["at"])
        assert abs(ab - w["abs"]) <= 2e-6 * w["abs"] + 1e-9, k
        assert abs(s - w["sum"]) <= 2e-6 * w["abs"] + 1e-9, k
        assert np.allclose(vals, w["values"], rtol=2e-5, atol=1e-7), k


def test_replay_memory_matches_reference(vec, agent):
    tr = vec["transitions"]
    for t, want in zip(tr, vec["memory_trace"]):
        agent.replay_memory.push(tuple(t))
        mem = agent.replay_memory
        assert (mem.ptr, mem.size) == (want["ptr"], want["size"])
        assert [tr.index(list(x)) for x in mem.storage] == want["slots"]
    random.seed(vec["sample_seed"])
    state, nxt, action, reward, done = agent.replay_memory.sample(4)
    assert [int(a) for a in action] == vec["sample_actions"]
    assert [int(r) for r in reward] == vec["sample_rewards"]
    assert len(state) == len(nxt) == len(done) == 4


def test_select_matches_reference(vec, agent):
    tr = vec["transitions"]
    random.seed(vec["select_seed"])
    np.random.seed(vec["select_seed"])
    torch.manual_seed(vec["select_seed"])
    agent.current_epsilon = vec["select_epsilon"]
    got = [int(agent.select(tr[i % len(tr)][0])) for i in range(len(vec["selected"]))]
    assert got == vec["selected"]


def test_update_steps_match_reference(vec, agent):
    for t in vec["transitions"]:
        agent.replay_memory.push(tuple(t))
    random.seed(vec["update_seeds"][0])
    torch.manual_seed(vec["update_seeds"][1])
    losses = [agent.update() for _ in vec["losses"]]
    assert np.allclose(losses, vec["losses"], rtol=2e-5), (losses, vec["losses"])
    check_digests(agent.eval_net.state_dict(), vec["eval_digest"])
    check_digests(agent.target_net.state_dict(), vec["target_digest"])
    assert agent.qnet is None                                   # any frozen device copy is invalidated by an update


def test_update_waits_for_a_full_batch_and_epsilon_schedule(vec, agent):
    for t in vec["transitions"][:vec["knobs"]["BATCH_SIZE"] - 1]:
        agent.replay_memory.push(tuple(t))
    assert agent.update() is None
    got = []
    for _ in vec["epsilon_schedule"]:
        agent.update_epsilon()
        got.append(agent.current_epsilon)
    assert np.allclose(got, vec["epsilon_schedule"], rtol=0, atol=1e-12)


def test_trained_weights_keep_the_reference_file_layout(vec, agent, tmp_path):
    from DPAMSA.dqn import STATE_KEYS
    for t in vec["transitions"]:
        agent.replay_memory.push(tuple(t))
    agent.update()
    agent.save("trained", path=str(tmp_path))
    sd = torch.load(os.path.join(str(tmp_path), "trained.pth"), map_location="cpu")
    assert list(sd) and set(sd) == set(STATE_KEYS)
    L = vec["rows"] * (vec["cols"] + 1)
    assert tuple(sd["encoder.position_enc.pos_table"].shape) == (1, L, 64)
    assert tuple(sd["l1.weight"].shape) == (1028, L * 64) and tuple(sd["l3.weight"].shape) == (3, 512)
    assert all(v.dtype == torch.float32 for v in sd.values())
