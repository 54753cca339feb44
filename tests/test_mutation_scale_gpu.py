"""Mutation parity at the BASELINE shapes: sampled individuals of a full-size mutation
phase compared bit-exactly with the oracle's per-individual episode on the reference's own torch fp32 ops
(O.qnet_forward_torch, which reproduces the reference Net's recorded Q-values bit for bit), the duplicate
elimination of gad_mutate against the run-everything path, and greedy-action parity as a measured rate on
4 096 recorded reference states per agent window (reference GA.py:353-373, DPAMSA/dqn.py:151-154)."""
import os
import random

import numpy as np
import pytest

import ga_oracle as O
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

WEIGHTS = (4, -4, -4)


@pytest.fixture(scope="module")
def gpu(native_lib):
    native_lib.require_device()
    import torch
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def weight_files(gpu):
    import torch
    import config
    out = {}
    os.makedirs(config.DPAMSA_WEIGHTS_PATH, exist_ok=True)
    for rows, cols in ((3, 30), (6, 30)):
        sd = O.synth_qnet_weights(rows, cols, 7)
        name = "syn%dx%d" % (rows, cols)
        torch.save({k: torch.from_numpy(v) for k, v in sd.items()}, os.path.join(config.DPAMSA_WEIGHTS_PATH, name + ".pth"))
        out[(rows, cols)] = (name, sd)
    return out


def synthetic_population(R, C, P, seed, clone_rate=0.25):
    """The GA.generate_population recipe (GA.py:71-94) with a vectorised generator: boards are inputs to both sides."""
    rng = np.random.default_rng(seed)
    base = rng.integers(1, 5, size=(R, C), dtype=np.uint8)
    gaps = max(1, round(C * 0.05))
    W = C + gaps
    stride = -(-W // 16) * 16
    boards = np.full((P, R, stride), 5, np.uint8)
    rowlen = np.full((P, R), C, np.int32)
    boards[:, :, :C] = base[None]
    n_clone = round(P * clone_rate)
    m = P - n_clone
    keys = rng.random((m, R, W - 1))
    slots = np.argsort(keys, axis=2)[:, :, :gaps]
    is_gap = np.zeros((m, R, W), dtype=bool)
    np.put_along_axis(is_gap, slots, True, axis=2)
    src = np.clip(np.cumsum(~is_gap, axis=2) - 1, 0, C - 1)
    vals = np.take_along_axis(np.broadcast_to(base[None], (m, R, C)), src, axis=2)
    vals[is_gap] = 5
    boards[n_clone:, :, :W] = vals
    rowlen[n_clone:] = W
    return boards, rowlen


def device_population(boards, rowlen, rw, cw):
    from population import DevicePopulation
    extra = (rw - 1) * cw
    wide = np.full((boards.shape[0], boards.shape[1], -(-(boards.shape[2] + extra) // 16) * 16), 5, np.uint8)
    wide[:, :, :boards.shape[2]] = boards
    return DevicePopulation.from_numpy(wide, rowlen, capacity=boards.shape[0])


def one_generation(dev, net, mode, rw, cw, P, seed):
    """mutation -> fitness -> selection -> crossover with the fitness re-evaluations of GA.run (GA.py:494-526)."""
    idx = dev.rank_order(mode, round(P * 0.25))
    tile = dev.worst_tiles(idx, rw, cw, mode, *WEIGHTS)
    dev.mutate(net, idx, tile, cw * (rw * (rw - 1) / 2 * 4))
    dev.fitness(*WEIGHTS)
    dev.hof_update(mode)
    n_keep = dev.select(mode, P // 2)
    dev.fitness(*WEIGHTS)
    dev.hof_update(mode)
    rng = random.Random(seed)
    plan = np.array(O.crossover_plan(n_keep, dev.R, P - n_keep, rng), dtype=np.int32).reshape(-1, 3)
    dev.crossover(plan, 0)
    dev.fitness(*WEIGHTS)
    dev.hof_update(mode)


@pytest.mark.parametrize("R,C,P,rw,cw,mode,n_sample,warm_generations", [
    (6, 60, 65536, 6, 30, "mo", 256, 2),          # c3 at full population: 16 384 episodes
    (20, 200, 16384, 3, 30, "cs", 64, 1),         # c4 board shape (population reduced: the CPU side is the limit)
])
def test_sampled_mutation_bit_exact_at_baseline_shapes(gpu, weight_files, R, C, P, rw, cw, mode, n_sample, warm_generations):
    import torch
    from DPAMSA.dqn import load_qnet
    name, sd = weight_files[(rw, cw)]
    net = load_qnet(name, rw, cw)
    boards, rowlen = synthetic_population(R, C, P, R * 11 + C)
    dev = device_population(boards, rowlen, rw, cw)
    dev.fitness(*WEIGHTS)
    dev.hof_update(mode)
    for g in range(warm_generations):              # so that the mutated set is no longer one clone repeated
        one_generation(dev, net, mode, rw, cw, P, 100 + g)
    M = round(P * 0.25)
    idx = dev.rank_order(mode, M).clone()
    tile = dev.worst_tiles(idx, rw, cw, mode, *WEIGHTS)
    pick = np.sort(np.random.default_rng(5).choice(M, n_sample, replace=False))
    ind = idx.cpu().numpy()[pick]
    tiles = tile.cpu().numpy()[pick]
    sel = torch.from_numpy(ind.astype(np.int64)).cuda()
    before_b = dev.cur.boards[sel].cpu().numpy()
    before_r = dev.cur.rowlen[sel].cpu().numpy()
    e0, u0 = dev.episodes, dev.unique_episodes
    dev.mutate(net, idx, tile, cw * (rw * (rw - 1) / 2 * 4))
    after_b = dev.cur.boards[sel].cpu().numpy()
    after_r = dev.cur.rowlen[sel].cpu().numpy()
    assert dev.episodes - e0 == M and 1 <= dev.unique_episodes - u0 <= M

    sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
    forward = lambda states, mv: O.qnet_forward_torch(sd_t, states, mv)
    memo = {}
    flips = 0
    for k in range(n_sample):
        board = [before_b[k, r, :before_r[k, r]].tolist() for r in range(R)]
        fr, fc = int(tiles[k, 0]), int(tiles[k, 1])
        _, t = O.worst_tile(board, mode, rw, cw)
        assert t == (fr, fr + rw, fc, fc + cw)                                  # a15 at this scale
        key = tuple(tuple(row[fc:fc + cw]) for row in board[fr:fr + rw])
        if key not in memo:                                                      # the oracle's episode, once per window
            work = [list(r) for r in board]
            O.mutate_board(work, t, sd, forward=forward)
            memo[key] = [row[fc:len(row) - (len(board[fr + i]) - fc - cw)] for i, row in enumerate(work[fr:fr + rw])]
        for i, row in enumerate(memo[key]):
            board[fr + i][fc:fc + cw] = row
        got = [after_b[k, r, :after_r[k, r]].tolist() for r in range(R)]
        flips += got != board
    print("\n%dx%d P=%d: %d sampled individuals (%d distinct windows) of %d episodes (%d run), %d differ from the oracle"
          % (R, C, P, n_sample, len(memo), M, dev.unique_episodes - u0, flips))
    assert flips == 0


def test_duplicate_elimination_equals_running_every_episode(gpu, weight_files, monkeypatch):
    """gad_mutate with and without GAD_MUTATE_DEDUP on the same population (clones, survivors and children
    of two generations): identical boards, and the counters say what ran."""
    from DPAMSA.dqn import load_qnet
    R, C, P, rw, cw, mode = 6, 60, 8192, 6, 30, "mo"
    name, _ = weight_files[(rw, cw)]
    net = load_qnet(name, rw, cw)
    results = {}
    for dedup in ("1", "0"):
        monkeypatch.setenv("GAD_MUTATE_DEDUP", dedup)
        boards, rowlen = synthetic_population(R, C, P, 77)
        dev = device_population(boards, rowlen, rw, cw)
        dev.fitness(*WEIGHTS)
        dev.hof_update(mode)
        per_gen = []
        for g in range(3):
            e0, u0 = dev.episodes, dev.unique_episodes
            one_generation(dev, net, mode, rw, cw, P, 200 + g)
            per_gen.append((dev.episodes - e0, dev.unique_episodes - u0))
        results[dedup] = (dev.to_numpy(), dev.scores_numpy(), per_gen, dev.hof_numpy())
    (b1, r1), s1, gens1, hof1 = results["1"]
    (b0, r0), s0, gens0, hof0 = results["0"]
    assert np.array_equal(r1, r0)
    width = int(r1.max())
    assert np.array_equal(b1[:, :, :width], b0[:, :, :width])
    for a, b in zip(s1, s0):
        assert np.array_equal(a, b)
    assert np.array_equal(hof1[0], hof0[0]) and hof1[1] == hof0[1]
    assert all(e == round(P * 0.25) for e, _ in gens1 + gens0)
    assert all(u == e for e, u in gens0)                                        # switch off: everything runs
    assert all(u < e for e, u in gens1)
    print("\nunique / episodes per generation:", gens1)


def test_action_parity_rate_on_recorded_reference_states(gpu, weight_files):
    """4 096 states per window from the reference's own greedy trajectories (oracle/make_golden_qnet_big.py):
    the device Q-net must pick the reference's action on every state whose reference top-2 margin exceeds the
    north star's tolerance (1e-3 relative on Q); the flip count over ALL states is reported (0 measured on B200)."""
    from DPAMSA.dqn import load_qnet
    path = os.path.join(GOLDEN, "qnet_states_big.npz")
    z = np.load(path)
    assert int(z["weight_seed"]) == 7
    for (rows, cols), (name, sd) in weight_files.items():
        tag = "%dx%d" % (rows, cols)
        states, act, top1, margin = z["states_" + tag], z["action_" + tag], z["top1_" + tag], z["margin_" + tag]
        assert len(states) >= 4096
        max_value = cols * rows * (rows - 1) / 2 * 4
        net = load_qnet(name, rows, cols)
        q, a = net.forward(states.astype(np.uint8), max_value)
        flips = a != act
        tol = 1e-3 * np.abs(top1) + 1e-5 * max_value
        decided = margin > 2 * tol                          # both Q-values may move by the tolerance
        print("\n%s: %d states, %d flips overall, %d with a margin above the tolerance; smallest reference margin %.3g"
              % (tag, len(states), int(flips.sum()), int((flips & decided).sum()), float(margin.min())))
        assert not (flips & decided).any()
        assert int(flips.sum()) <= int((~decided).sum())    # a flip is only tolerable on a numerically tied state
        head = z["q_head_" + tag]
        assert np.all(np.abs(q[:len(head)] - head) <= 1e-3 * np.abs(head) + 1e-5 * max_value)
        assert np.all(np.abs(q.max(1) - top1) <= 1e-3 * np.abs(top1) + 1e-5 * max_value)


def test_q_values_do_not_depend_on_the_batch(gpu, weight_files):
    """l1 sums its k-slices in one canonical order whether a tile is computed by one CTA (large batches) or split
    over 2 or 4 k-slice groups with a separate fold (small batches, gemm_tc.cu): a state's Q-values are bit-identical
    in every batch size class - what makes duplicate elimination and sharding (both change the batches) exact."""
    from DPAMSA.dqn import QNet
    z = np.load(os.path.join(GOLDEN, "qnet_states_big.npz"))
    for (rows, cols), (name, sd) in weight_files.items():
        states = z["states_%dx%d" % (rows, cols)].astype(np.uint8)
        max_value = cols * rows * (rows - 1) / 2 * 4
        net = QNet(sd, rows, cols)
        q_all, a_all = net.forward(states, max_value)               # 4 096 rows: whole-K units
        for n in (1, 97, 130, 896, 900, 1792, 1800, 2049, 3000, 3700):   # every split class of the cost model, whole-K with half units
            q, a = net.forward(states[:n], max_value)
            assert np.array_equal(q, q_all[:n]), n
            assert np.array_equal(a, a_all[:n]), n
