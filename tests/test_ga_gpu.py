"""The generation loop on the GPU (ranking, selection, hall of fame, crossover, tile search, Q-net,
mutation, GA.run) through the C-ABI and the reference-shaped facade, against the oracle and the
golden fixtures generated from the unmodified reference (oracle/make_golden.py). Integer, index and
board results are bit-exact; Q-values follow the north star's 1e-3 relative tolerance."""
import copy
import fractions
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


@pytest.fixture()
def cfg():
    """config knobs restored after the test."""
    import config
    names = ["AGENT_WINDOW_ROW", "AGENT_WINDOW_COLUMN", "GA_ITERATIONS", "POPULATION_SIZE", "CLONE_RATE", "GAP_RATE",
             "SELECTION_RATE", "MUTATION_RATE", "MATCH_REWARD", "MISMATCH_PENALTY", "GAP_PENALTY", "CROSSOVER_KIND",
             "MUTATION_TILE", "MUTATION_PROBABILITY", "CROSSOVER_PROBABILITY", "SELECTION_SCHEME", "TOURNAMENT_RATE",
             "ELITISM_PROBABILITY"]
    saved = {n: getattr(config, n) for n in names}
    yield config
    for n, v in saved.items():
        setattr(config, n, v)


@pytest.fixture(scope="module")
def weight_files(gpu):
    """Substitute weights (oracle.synth_qnet_weights, the ones the golden traces were made with)
    written in the reference's .pth layout under config.DPAMSA_WEIGHTS_PATH."""
    import torch
    import config
    z = np.load(os.path.join(GOLDEN, "qnet_vectors.npz"))
    seed = int(z["weight_seed"])
    out = {}
    os.makedirs(config.DPAMSA_WEIGHTS_PATH, exist_ok=True)
    for rows, cols in ((3, 30), (6, 30), (2, 5)):
        sd = O.synth_qnet_weights(rows, cols, seed)
        name = "syn%dx%d" % (rows, cols)
        torch.save({k: torch.from_numpy(v) for k, v in sd.items()},
                   os.path.join(config.DPAMSA_WEIGHTS_PATH, name + ".pth"))
        out[(rows, cols)] = (name, sd)
    return out


def tagged_board(i):
    """A 2x4 board of valid codes that encodes the integer i (base 4)."""
    digits = [1 + (i >> (2 * k)) % 4 for k in range(4)]
    return [digits, digits[::-1]]


def untag(board):
    return sum((d - 1) << (2 * k) for k, d in enumerate(board[0][:4]))


def scores_to_arrays(scores, mode):
    n = len(scores)
    sp = np.zeros(n, np.int32)
    em = np.zeros(n, np.int32)
    al = np.ones(n, np.int32)
    for i, t in enumerate(scores):
        cs = None
        if len(t) == 3:
            sp[i], cs = t[1], t[2]
        elif mode == "sp":
            sp[i] = t[1]
        else:
            cs = t[1]
        if cs is not None:
            fr = fractions.Fraction(cs).limit_denominator(1000)
            assert fr.numerator / fr.denominator == cs
            em[i], al[i] = fr.numerator, fr.denominator
    return sp, em, al


def device_population(gpu, n, scores=None, mode="sp"):
    from population import DevicePopulation
    dev = DevicePopulation.from_lists([tagged_board(i) for i in range(n)])
    if scores is not None:
        sp, em, al = scores_to_arrays(scores, mode)
        dev.cur.sp[:n].copy_(gpu.from_numpy(sp))
        dev.cur.em[:n].copy_(gpu.from_numpy(em))
        dev.cur.al[:n].copy_(gpu.from_numpy(al))
    return dev


# ------------------------------------------------------------------------------------------- a9
def test_rank_golden(gpu, primitives):
    import utils
    for rec in primitives["rank"]:
        scores = [tuple(t) for t in rec["scores"]]
        assert utils.get_index_of_the_best_fitted_individuals(scores, rec["n"]) == rec["out"]
        mode = "mo" if len(scores[0]) == 3 else "sp"
        dev = device_population(gpu, len(scores), scores, mode)
        assert dev.rank_order(mode, rec["n"]).cpu().tolist() == rec["out"]


def test_rank_large_random_against_oracle(gpu):
    rng = random.Random(5)
    n = 20000
    scores = [(i, rng.randint(-40, 40) * 4, rng.randint(0, 63) / 63) for i in range(n)]
    for mode, sc in (("sp", [(i, s) for i, s, _ in scores]), ("cs", [(i, c) for i, _, c in scores]), ("mo", scores)):
        dev = device_population(gpu, n, sc, mode)
        assert dev.rank_order(mode, 5000).cpu().tolist() == O.rank_best(sc, 5000)


# ------------------------------------------------------------------------------------------- a10 / a11
def test_selection_and_pareto_golden(gpu, primitives):
    for rec in primitives["selection"]:
        scores = [tuple(t) for t in rec["scores"]]
        dev = device_population(gpu, rec["n"], scores, rec["mode"])
        top_n = int(rec["n"] * rec["rate"])
        n_keep = dev.select(rec["mode"], top_n)
        assert [untag(b) for b in dev.to_lists()] == rec["keep"]
        assert n_keep == len(rec["keep"])
    for rec in primitives["pareto"]:
        scores = [tuple(t) for t in rec["scores"]]
        dev = device_population(gpu, len(scores), scores, "mo")
        keep, _ = dev.select_flags("mo", -1)
        assert np.nonzero(keep.cpu().numpy())[0].tolist() == rec["out"]


@pytest.mark.parametrize("mode", ["sp", "cs", "mo"])
def test_selection_large_random_against_oracle(gpu, mode):
    rng = random.Random(17)
    n = 6000
    full = [(i, rng.randint(-12, 12) * 4, rng.randint(0, 20) / 21) for i in range(n)]
    sc = full if mode == "mo" else [(i, s) for i, s, _ in full] if mode == "sp" else [(i, c) for i, _, c in full]
    for rate in (0.5, 0.01, 0.9):
        dev = device_population(gpu, n, sc, mode)
        keep, count = dev.select_flags(mode, int(n * rate))
        want = O.selection_keep(sc, mode, n, rate, fast=True)
        assert np.nonzero(keep.cpu().numpy())[0].tolist() == sorted(want)
        assert int(count.item()) == len(want)


def test_compaction_keeps_boards_scores_and_order(gpu):
    from population import DevicePopulation
    rng = random.Random(3)
    pop = [[[rng.choice([1, 2, 3, 4]) for _ in range(rng.randint(20, 70))] for _ in range(5)] for _ in range(300)]
    for b in pop:
        w = max(map(len, b))
        for row in b:
            row.extend([5] * (w - len(row)))
    dev = DevicePopulation.from_lists(copy.deepcopy(pop))
    dev.fitness(*WEIGHTS)
    sp, em, al = [a.copy() for a in dev.scores_numpy()]
    scores = [(i, int(sp[i])) for i in range(len(pop))]
    keep = sorted(O.selection_keep(scores, "sp", len(pop), 0.3))
    assert dev.select("sp", int(len(pop) * 0.3)) == len(keep)
    cleaned = copy.deepcopy(pop)
    O.fitness_scores(cleaned, "sp")
    assert dev.to_lists() == [cleaned[i] for i in keep]
    sp2, em2, al2 = dev.scores_numpy()
    assert sp2.tolist() == sp[keep].tolist() and em2.tolist() == em[keep].tolist() and al2.tolist() == al[keep].tolist()


# ------------------------------------------------------------------------------------------- a8
def test_hall_of_fame_golden(gpu, primitives):
    for rec in primitives["hof"]:
        mode = rec["mode"]
        n = len(rec["scores1"])
        dev = device_population(gpu, n, [tuple(t) for t in rec["scores1"]], mode)
        dev.hof_update(mode)
        _, (idx, sp, em, al) = dev.hof_numpy()          # al is a synthetic denominator here: read the raw board
        board = dev.hof_board[:, :4].cpu().numpy()
        assert idx == rec["first"][1][0] and untag(board.tolist()) == rec["first"][0][0][0] - 1
        got = (idx, sp) if mode == "sp" else (idx, sp, em / al)
        assert list(got) == rec["first"][1]
        # second population: boards tagged 100+i in the golden record -> tag i+64 here
        sp2, em2, al2 = scores_to_arrays([tuple(t) for t in rec["scores2"]], mode)
        dev2_boards = [tagged_board(64 + i) for i in range(n)]
        from population import pack_boards
        b, r = pack_boards(dev2_boards, stride=dev.stride)
        dev.cur.boards[:n].copy_(gpu.from_numpy(b))
        dev.cur.sp[:n].copy_(gpu.from_numpy(sp2))
        dev.cur.em[:n].copy_(gpu.from_numpy(em2))
        dev.cur.al[:n].copy_(gpu.from_numpy(al2))
        dev.hof_update(mode)
        _, (idx, sp, em, al) = dev.hof_numpy()
        board = dev.hof_board[:, :4].cpu().numpy()
        want_board, want_score = rec["second"]
        assert idx == want_score[0]
        tag = want_board[0][0]                       # golden: i+1 (first population) or 100+i (second)
        assert untag(board.tolist()) == (tag - 100 + 64 if tag >= 100 else tag - 1)
        got = (idx, sp) if mode == "sp" else (idx, sp, em / al)
        assert list(got) == want_score


@pytest.mark.parametrize("mode", ["sp", "cs", "mo"])
def test_hall_of_fame_large_random_against_oracle(gpu, mode):
    """The one-launch decision (hof_decide_kernel) on grids of many blocks: score ranges over a grid barrier,
    first maximum among many ties, the record re-entering as element n, three updates in a row."""
    rng = random.Random(23)
    for n in (1500, 70001, 200000):
        dev = None
        hof = None
        boards = [tagged_board(i) for i in range(n)]
        for rnd in range(3):
            hi = 6 + 3 * rnd if rnd < 2 else 4                     # third round: the old record usually stays
            full = [(i, rng.randint(-hi, hi) * 4, rng.randint(0, 2 * hi) / 21) for i in range(n)]
            sc = full if mode == "mo" else [(i, s) for i, s, _ in full] if mode == "sp" else [(i, c) for i, _, c in full]
            if dev is None:
                dev = device_population(gpu, n, sc, mode)
            else:
                sp, em, al = scores_to_arrays(sc, mode)
                dev.cur.sp[:n].copy_(gpu.from_numpy(sp))
                dev.cur.em[:n].copy_(gpu.from_numpy(em))
                dev.cur.al[:n].copy_(gpu.from_numpy(al))
            dev.hof_update(mode)
            hof = O.hof_update(hof, boards, sc)
            rec = dev.hof.cpu().tolist()
            assert rec[5:] == [0, 0, 0]                            # barrier counter and ticket are back to zero
            got = (rec[1], rec[2]) if mode == "sp" else (rec[1], rec[3] / rec[4]) if mode == "cs" else \
                (rec[1], rec[2], rec[3] / rec[4])
            assert got == tuple(hof[1]), (n, rnd, got, hof[1])
            assert dev.hof_board[:, :4].cpu().tolist() == [row[:4] for row in hof[0]]


# ------------------------------------------------------------------------------------------- a3 / a12 / a13
def test_population_and_crossover_rng_replay(gpu, primitives, cfg, native_lib):
    """Same CPython `random` stream as the reference: populations, crossover children and the state of
    the generator afterwards are identical."""
    import GA as ga_mod
    for rec in primitives["population"]:
        cfg.POPULATION_SIZE = rec["P"]
        cfg.CLONE_RATE, cfg.GAP_RATE = 0.25, 0.05
        random.seed(rec["seed"])
        g = ga_mod.GA(rec["seqs"], "sp")
        g.generate_population()
        assert g.population == rec["population"]
        n_surv = rec["n_survivors"]
        g.population = rec["population"][:n_surv]
        dev = g._dev
        plan = np.empty((rec["P"] - n_surv, 3), np.int32)
        state, meta = ga_mod._mt_state()
        native_lib.call("gad_mt_crossover_plan_host", state.ctypes.data, n_surv, dev.R, rec["P"] - n_surv, None,
                        plan.ctypes.data)
        ga_mod._mt_restore(state, meta)
        dev.crossover(plan, 0)
        assert g.population == rec["after_crossover"]
        assert random.random() == rec["next_random"]


def test_crossover_plan_matches_python_random(gpu, native_lib):
    import GA as ga_mod
    for seed, (n_par, rows, n_child) in enumerate([(2, 2, 50), (3, 6, 200), (1000, 20, 5000), (32768, 6, 32768)]):
        random.seed(seed)
        want = O.crossover_plan(n_par, rows, n_child)
        after = random.random()
        random.seed(seed)
        plan = np.empty((n_child, 3), np.int32)
        state, meta = ga_mod._mt_state()
        native_lib.call("gad_mt_crossover_plan_host", state.ctypes.data, n_par, rows, n_child, None, plan.ctypes.data)
        ga_mod._mt_restore(state, meta)
        assert plan.tolist() == [list(t) for t in want]
        assert random.random() == after
    with pytest.raises(ValueError):
        state, _ = ga_mod._mt_state()
        native_lib.call("gad_mt_crossover_plan_host", state.ctypes.data, 1, 6, 4, None, plan.ctypes.data)


def test_crossover_children_against_oracle(gpu):
    from population import DevicePopulation
    rng = random.Random(8)
    parents = []
    for _ in range(40):
        w = rng.randint(5, 90)
        parents.append([[rng.choice([1, 2, 3, 4, 5]) for _ in range(w)] for _ in range(7)])
    for kind in (0, 1):
        plan = []
        for _ in range(300):
            a, b = rng.sample(range(40), 2)
            hi = 6 if kind == 0 else min(len(parents[a][0]), len(parents[b][0])) - 1
            plan.append((a, b, rng.randint(1, hi)))
        dev = DevicePopulation.from_lists(copy.deepcopy(parents), capacity=400)
        dev.crossover(np.array(plan, np.int32), kind)
        child = O.horizontal_child if kind == 0 else O.vertical_child
        want = parents + [child(parents[a], parents[b], c) for a, b, c in plan]
        assert dev.to_lists() == want
        # the children obey the layout invariant: a fitness pass gives the oracle's result
        dev.fitness(*WEIGHTS)
        ref = copy.deepcopy(want)
        sc = O.fitness_scores(ref, "mo")
        assert dev.to_lists() == ref
        assert dev.scores_numpy()[0].tolist() == [s for _, s, _ in sc]


# ------------------------------------------------------------------------------------------- a15
def test_worst_tile_golden(gpu, primitives, cfg):
    import utils
    for rec in primitives["worst"]:
        cfg.AGENT_WINDOW_ROW, cfg.AGENT_WINDOW_COLUMN = rec["rw"], rec["cw"]
        score, tile = utils.calculate_worst_fitted_sub_board(rec["board"], rec["mode"])
        assert list(tile) == rec["tile"]
        assert score == rec["score"]
    for rec in primitives["tiles"]:
        board = [[1] * rec["W"] for _ in range(rec["R"])]
        assert [list(t) for t in utils.get_all_different_sub_range(board, rec["rw"], rec["cw"])] == rec["out"]


def test_worst_tiles_population_against_oracle(gpu):
    from population import DevicePopulation
    rng = random.Random(12)
    for R, W, rw, cw in ((6, 63, 3, 30), (6, 63, 6, 30), (20, 210, 3, 30), (7, 59, 3, 30)):
        pop = [[[rng.choice([1, 2, 3, 4, 5, 1, 2]) for _ in range(W)] for _ in range(R)] for _ in range(120)]
        dev = DevicePopulation.from_lists(pop)
        for mode in ("sp", "cs", "mo"):
            for w in (WEIGHTS, (5, -3, -7)):
                got = dev.worst_tiles(None, rw, cw, mode, *w).cpu().numpy()
                for p, board in enumerate(pop):
                    _, t = O.worst_tile(board, mode, rw, cw, O.Weights(*w))
                    assert (got[p, 0], got[p, 1]) == (t[0], t[2])


# ------------------------------------------------------------------------------------------- a16-a18
def test_qnet_golden_vectors(gpu, weight_files):
    """Q-values of the reference Net (eval mode) on recorded states: 1e-3 relative (plus a floor of
    1e-5 * max_value for Q near 0) and identical greedy actions."""
    from DPAMSA.dqn import load_qnet
    z = np.load(os.path.join(GOLDEN, "qnet_vectors.npz"))
    for (rows, cols), (name, sd) in weight_files.items():
        tag = "%dx%d" % (rows, cols)
        states, q_ref = z["states_" + tag], z["q_" + tag]
        max_value = cols * rows * (rows - 1) / 2 * 4
        net = load_qnet(name, rows, cols)
        q, a = net.forward(states.astype(np.uint8), max_value)
        assert np.all(np.abs(q - q_ref) <= 1e-3 * np.abs(q_ref) + 1e-5 * max_value), np.abs(q - q_ref).max()
        assert np.array_equal(a, q_ref.argmax(1))
        q64 = O.qnet_forward(sd, states, max_value)
        assert np.all(np.abs(q - q64) <= 1e-3 * np.abs(q64) + 1e-5 * max_value)


def test_qnet_tensor_core_l1_against_fp32_tiles(gpu, weight_files, monkeypatch):
    """The tcgen05 split-fp16 l1 (hi.hi + lo.hi + hi.lo, fp32 accumulation in TMEM) agrees with the fp32
    SIMT tiles to fp32 rounding on every recorded state, for every batch size class (partial tiles)."""
    from DPAMSA.dqn import QNet
    z = np.load(os.path.join(GOLDEN, "qnet_vectors.npz"))
    for (rows, cols), (name, sd) in weight_files.items():
        states = z["states_%dx%d" % (rows, cols)].astype(np.uint8)
        max_value = cols * rows * (rows - 1) / 2 * 4
        monkeypatch.setenv("GAD_QNET_SIMT", "1")
        q_simt, a_simt = QNet(sd, rows, cols).forward(states, max_value)
        monkeypatch.setenv("GAD_QNET_SIMT", "0")
        net = QNet(sd, rows, cols)
        for n in (1, 3, len(states)):
            q_tc, a_tc = net.forward(states[:n], max_value)
            assert np.all(np.abs(q_tc - q_simt[:n]) <= 2e-5 * np.abs(q_simt[:n]) + 1e-6 * max_value), \
                np.abs(q_tc - q_simt[:n]).max()
            assert np.array_equal(a_tc, a_simt[:n])
        big = np.concatenate([states] * (300 // len(states) + 1))[:300]          # 3 M-tiles, the last one partial
        q_big, _ = net.forward(big, max_value)
        assert np.array_equal(q_big[:len(states)], net.forward(states, max_value)[0])


@pytest.mark.parametrize("variant", ["fa", "mma", "simt"])
def test_qnet_encoder_variants_agree(gpu, weight_files, monkeypatch, variant):
    """The alternative encoders (flash-attention style on mma.sync; table logits + mma.sync; table logits + fp32
    tiles) give the default tcgen05 attention kernel's Q-values to fp32 rounding and the same greedy actions."""
    from DPAMSA.dqn import QNet
    z = np.load(os.path.join(GOLDEN, "qnet_vectors.npz"))
    for (rows, cols), (name, sd) in weight_files.items():
        states = z["states_%dx%d" % (rows, cols)].astype(np.uint8)
        max_value = cols * rows * (rows - 1) / 2 * 4
        net = QNet(sd, rows, cols)
        monkeypatch.delenv("GAD_QNET_ENCODER", raising=False)
        q0, a0 = net.forward(states, max_value)
        monkeypatch.setenv("GAD_QNET_ENCODER", variant)
        q1, a1 = net.forward(states, max_value)
        assert np.all(np.abs(q1 - q0) <= 3e-5 * np.abs(q0) + 2e-6 * max_value), np.abs(q1 - q0).max()
        assert np.array_equal(a1, a0)


@pytest.mark.parametrize("rows,cols", [(4, 30), (5, 30), (4, 20)])
def test_qnet_tc_encoder_other_tilings(gpu, monkeypatch, rows, cols):
    """The tcgen05 attention on windows whose tiling differs from the golden ones: 4x30 is one 128-row tile (two
    threads per row in warps q and q+4 with the MMA issuer in a ninth warp), 5x30 two 96-row tiles with a short
    second tile, 4x20 one 96-row tile. Random ragged states, fully padded rows included."""
    from DPAMSA.dqn import QNet
    sd = O.synth_qnet_weights(rows, cols, 11)
    rng = np.random.default_rng(rows * 100 + cols)
    L = rows * (cols + 1)
    states = np.zeros((300, L), np.uint8)
    for s in states:
        pos = 0
        for _ in range(rows):
            rem = int(rng.integers(0, cols + 1))
            s[pos:pos + rem] = rng.integers(1, 5, rem)
            s[pos + rem] = 5
            pos += rem + 1
    for s in states[200:260]:                                   # padding between tokens: only reachable through predict()
        s[rng.integers(0, L, 12)] = 0                           # with hand-made states; takes the kernel's gather path
    states[260:264] = 0                                         # fully masked states
    max_value = cols * rows * (rows - 1) / 2 * 4
    net = QNet(sd, rows, cols)
    monkeypatch.setenv("GAD_QNET_ENCODER", "fa")
    q0, a0 = net.forward(states, max_value)
    monkeypatch.setenv("GAD_QNET_ENCODER", "tc")
    q1, a1 = net.forward(states, max_value)
    assert np.all(np.abs(q1 - q0) <= 3e-5 * np.abs(q0) + 2e-6 * max_value), np.abs(q1 - q0).max()
    assert np.array_equal(a1, a0)


def test_dqn_and_environment_facade(gpu, weight_files, primitives):
    from DPAMSA.dqn import DQN
    from DPAMSA.env import Environment
    name, sd = weight_files[(2, 5)]
    env = Environment([[1, 2, 3, 4, 1], [2, 2, 3, 1, 1]], convert_data=False)
    agent = DQN(env.action_number, env.row, env.max_len, env.max_len * env.max_reward)
    agent.load(name)
    oenv = O.Env([[1, 2, 3, 4, 1], [2, 2, 3, 1, 1]])
    state = env.reset()
    assert state == oenv.reset()
    while True:
        a = agent.predict(state)
        want = O.greedy_action(O.qnet_forward(sd, [state], env.max_len * env.max_reward, np.float32)[0])
        assert a.shape == (1,) and int(a[0]) == want
        r, state, done = env.step(a)
        r2, s2, d2 = oenv.step(want)
        assert (r, state, done) == (r2, s2, d2)
        if done == 0:
            break
    env.padding()
    oenv.padding()
    assert env.aligned == oenv.aligned
    for rec in primitives["env"]:
        env = Environment(copy.deepcopy(rec["data"]), convert_data=False)
        assert env.max_reward == rec["max_reward"] and env.action_number == rec["action_number"]
        assert env.reset() == rec["steps"][0]["state"]
        for st in rec["steps"][1:]:
            r, s, d = env.step(st["action"])
            assert (r, s, d) == (st["reward"], st["state"], st["done"])
            assert env.aligned == st["aligned"]
        env.padding()
        assert env.aligned == rec["padded"]
    with pytest.raises(FileNotFoundError):
        DQN(7, 3, 30, 360.0).load("no_such_model")


@pytest.mark.parametrize("rows,cols,R,W", [(3, 30, 6, 63), (6, 30, 6, 64), (3, 30, 7, 95), (2, 5, 5, 23)])
def test_mutation_against_oracle(gpu, weight_files, rows, cols, R, W):
    """Batched device episodes + padding + splice == the oracle's per-individual GA.py:335-373."""
    from DPAMSA.dqn import load_qnet
    from population import DevicePopulation
    name, sd = weight_files[(rows, cols)]
    net = load_qnet(name, rows, cols)
    rng = random.Random(rows * 100 + W)
    pop = [[[rng.choice([1, 2, 3, 4, 5, 1, 2, 3, 4]) for _ in range(W)] for _ in range(R)] for _ in range(24)]
    dev = DevicePopulation.from_lists(copy.deepcopy(pop))
    gpu_idx = gpu.tensor([3, 0, 7, 23, 11, 12, 5], dtype=gpu.int32, device="cuda")
    tile = dev.worst_tiles(gpu_idx, rows, cols, "sp", *WEIGHTS)
    max_value = cols * rows * (rows - 1) / 2 * 4
    dev.mutate(net, gpu_idx, tile, max_value)
    want = copy.deepcopy(pop)
    for i in gpu_idx.cpu().tolist():
        _, t = O.worst_tile(want[i], "sp", rows, cols)
        O.mutate_board(want[i], t, sd, qdtype=np.float32)
    assert dev.to_lists() == want
    dev.fitness(*WEIGHTS)
    sc = O.fitness_scores(want, "mo")
    assert dev.to_lists() == want
    assert dev.scores_numpy()[0].tolist() == [s for _, s, _ in sc]


# ------------------------------------------------------------------------------------------- a19: whole runs
def run_facade(tr, cfg, model):
    import GA as ga_mod
    for k, v in tr["knobs"].items():
        setattr(cfg, k, v)
    random.seed(tr["seed"])
    g = ga_mod.GA(tr["seqs"], tr["mode"])
    log = []
    inner = g.calculate_fitness_score

    def logged():
        inner()
        hof = g.hall_of_fame
        log.append({"scores": [list(s) for s in g.population_score], "hof": O.decode(hof[0]),
                    "hof_score": list(hof[1])})
    g.calculate_fitness_score = logged
    out = g.run(model)
    return out, log, g


def test_ga_run_reproduces_reference_traces(gpu, ga_traces, weight_files, cfg):
    """GA.run through the facade == the unmodified reference's recorded per-pass scores, hall of fame
    and final alignment, for every recorded (mode, seed, knobs) - population generation, mutation with
    greedy Q-net actions, selection and crossover included."""
    done = 0
    for tr in ga_traces["traces"]:
        kb = tr["knobs"]
        name, _ = weight_files[(kb["AGENT_WINDOW_ROW"], kb["AGENT_WINDOW_COLUMN"])]
        out, log, _ = run_facade(tr, cfg, name)
        assert len(log) == len(tr["passes"])
        for k, (mine, ref) in enumerate(zip(log, tr["passes"])):
            assert mine["scores"] == ref["scores"], (tr["mode"], tr["seed"], k)
            assert mine["hof"] == ref["hof"] and mine["hof_score"] == ref["hof_score"], (tr["mode"], tr["seed"], k)
        assert out == tr["result"]
        done += 1
    assert done >= 20


def test_ga_run_against_oracle_larger_population(gpu, weight_files, cfg, datasets_json):
    """P=256 on a 6x60 set, all three modes, against the oracle's whole loop (too slow for the golden file)."""
    name, sd = weight_files[(3, 30)]
    seqs = datasets_json["synthetic_dataset_6x60bp"]["test2"]
    for mode in ("sp", "cs", "mo"):
        knobs = {"AGENT_WINDOW_ROW": 3, "AGENT_WINDOW_COLUMN": 30, "GA_ITERATIONS": 2, "POPULATION_SIZE": 256,
                 "CLONE_RATE": 0.25, "GAP_RATE": 0.05, "SELECTION_RATE": 0.5, "MUTATION_RATE": 0.1}
        tr = {"knobs": knobs, "seed": 11, "seqs": seqs, "mode": mode}
        out, log, _ = run_facade(tr, cfg, name)
        random.seed(11)
        olog = []
        oout, _ = O.ga_run(seqs, mode, sd, O.Knobs(**knobs), log=olog, fast_pareto=True)
        assert len(log) == len(olog)
        for mine, ref in zip(log, olog):
            assert mine["scores"] == [list(s) for s in ref[1]]
            assert mine["hof"] == ref[2] and mine["hof_score"] == list(ref[3])
        assert out == oout


PAPER_VARIANTS = [
    {"MUTATION_TILE": "random"},
    {"MUTATION_PROBABILITY": 0.85},
    {"CROSSOVER_PROBABILITY": 0.5},
    {"SELECTION_SCHEME": "tournament", "TOURNAMENT_RATE": 0.25},
    {"ELITISM_PROBABILITY": 0.4},
    {"MUTATION_TILE": "random", "MUTATION_PROBABILITY": 0.85, "CROSSOVER_PROBABILITY": 0.5,
     "SELECTION_SCHEME": "tournament", "TOURNAMENT_RATE": 0.25, "ELITISM_PROBABILITY": 0.4, "SELECTION_RATE": 0.45},
]


@pytest.mark.parametrize("variant", PAPER_VARIANTS, ids=lambda v: "+".join(sorted(v)))
def test_paper_variants_against_oracle(gpu, weight_files, cfg, datasets_json, variant):
    """The GA variants of the paper's supplement (tables A3 / A5: random-tile mutation, mutation / crossover /
    elitism probabilities, tournament selection) - extensions without reference code, defined in config.py and
    restated by the oracle: every pass of GA.run (scores, hall of fame) and the final alignment equal the
    oracle's on the same `random` stream, 'sp' and 'mo'."""
    name, sd = weight_files[(3, 30)]
    seqs = datasets_json["synthetic_dataset_6x60bp"]["test3"]
    for mode in ("sp", "mo"):
        knobs = {"AGENT_WINDOW_ROW": 3, "AGENT_WINDOW_COLUMN": 30, "GA_ITERATIONS": 3, "POPULATION_SIZE": 60,
                 "CLONE_RATE": 0.25, "GAP_RATE": 0.05, "SELECTION_RATE": 0.5, "MUTATION_RATE": 0.5}
        knobs.update(variant)
        tr = {"knobs": knobs, "seed": 29, "seqs": seqs, "mode": mode}
        out, log, _ = run_facade(tr, cfg, name)
        random.seed(29)
        olog = []
        oout, _ = O.ga_run(seqs, mode, sd, O.Knobs(**knobs), log=olog, fast_pareto=True)
        assert len(log) == len(olog), (mode, len(log), len(olog))
        for k, (mine, ref) in enumerate(zip(log, olog)):
            assert mine["scores"] == [list(s) for s in ref[1]], (mode, k)
            assert mine["hof"] == ref[2] and mine["hof_score"] == list(ref[3]), (mode, k)
        assert out == oout
    # both sides drew the same numbers in the same order: the generator ends in the same state
    knobs = {"AGENT_WINDOW_ROW": 3, "AGENT_WINDOW_COLUMN": 30, "GA_ITERATIONS": 1, "POPULATION_SIZE": 24,
             "CLONE_RATE": 0.25, "GAP_RATE": 0.05, "SELECTION_RATE": 0.5, "MUTATION_RATE": 0.5}
    knobs.update(variant)
    run_facade({"knobs": knobs, "seed": 5, "seqs": seqs, "mode": "sp"}, cfg, name)
    after_facade = random.getstate()
    random.seed(5)
    O.ga_run(seqs, "sp", sd, O.Knobs(**knobs), fast_pareto=True)
    assert random.getstate() == after_facade


def test_run_validation_errors(gpu, weight_files, cfg):
    import GA as ga_mod
    name, _ = weight_files[(3, 30)]
    cfg.POPULATION_SIZE = 6
    cfg.AGENT_WINDOW_ROW, cfg.AGENT_WINDOW_COLUMN = 3, 30
    with pytest.raises(ValueError):
        ga_mod.GA(["ACGT" * 10, "ACGA" * 10], "sp").run(name)            # 2 rows < window rows
    with pytest.raises(ValueError):
        ga_mod.GA(["ACGTACGT", "ACGAACGT", "ACGAACGA"], "sp").run(name)   # 8 columns < window columns
    with pytest.raises(KeyError):
        ga_mod.GA(["ACGN" * 10] * 3, "sp").run(name)                      # unknown nucleotide (GA.py:79)


def test_vertical_crossover_run(gpu, weight_files, cfg, datasets_json):
    """Vertical crossover is an extension without reference code (parity unpinned): the run must keep
    the population size and every child must be a column-wise recombination of two survivors."""
    import GA as ga_mod
    name, _ = weight_files[(3, 30)]
    cfg.POPULATION_SIZE, cfg.GA_ITERATIONS, cfg.CROSSOVER_KIND = 40, 1, "vertical"
    random.seed(2)
    g = ga_mod.GA(datasets_json["dataset1_6x30bp"]["test0"], "sp")
    out = g.run(name)
    assert len(g.population) == 40 and len(out) == 6
    assert len({len(r) for r in out}) == 1


def test_vertical_crossover_known_answers(gpu, native_lib):
    """a13: hand-written cases from the written specification (tests/golden/vertical_crossover.json): the device
    kernel, the oracle's restatement and the specification sentence agree; the cut rule of the plan generator
    is randint(1, min(width1, width2) - 1) on the same `random` stream as a python restatement."""
    import json
    import GA as ga_mod
    from population import DevicePopulation
    with open(os.path.join(GOLDEN, "vertical_crossover.json")) as fh:
        gold = json.load(fh)
    for case in gold["cases"]:
        p1, p2, cut, want = case["p1"], case["p2"], case["cut"], case["child"]
        assert [a[:cut] + b[cut:] for a, b in zip(p1, p2)] == want          # the specification sentence itself
        assert O.vertical_child(p1, p2, cut) == want
        dev = DevicePopulation.from_lists([copy.deepcopy(p1), copy.deepcopy(p2)], capacity=4, extra_capacity=8)
        dev.crossover(np.array([[0, 1, cut]], np.int32), 1)
        assert dev.to_lists()[2] == want
        assert dev.to_lists()[:2] == [p1, p2]                                # parents untouched
    for rule in gold["cut_rule"]:
        widths = np.array([rule["w1"], rule["w2"]], np.int32)
        seen = set()
        for seed in range(40):
            random.seed(seed)
            plan = np.empty((8, 3), np.int32)
            state, meta = ga_mod._mt_state()
            native_lib.call("gad_mt_crossover_plan_host", state.ctypes.data, 2, 3, 8, widths.ctypes.data, plan.ctypes.data)
            ga_mod._mt_restore(state, meta)
            after = random.random()
            random.seed(seed)
            want = []
            while len(want) < 8:                                             # GA.py:281-293 with the vertical cut rule
                a, b = random.choice(range(2)), random.choice(range(2))
                if a == b:
                    continue
                want.append([a, b, random.randint(1, min(int(widths[a]), int(widths[b])) - 1)])
            assert plan.tolist() == want and random.random() == after
            seen.update(int(c) for c in plan[:, 2])
        assert seen == set(rule["cuts"])


def test_debug_mode_prints_like_the_reference(gpu, weight_files, cfg, datasets_json, capsys):
    """a22: GA.run(model, debug_mode=True) materialises the device population for print_population /
    print_hall_of_fame (GA.py:375-441) after every phase; what it prints is the population the oracle holds."""
    import GA as ga_mod
    name, sd = weight_files[(3, 30)]
    cfg.POPULATION_SIZE, cfg.GA_ITERATIONS = 5, 1
    seqs = datasets_json["dataset1_6x30bp"]["test0"]
    random.seed(4)
    g = ga_mod.GA(seqs, "mo")
    out = g.run(name, debug_mode=True)
    text = capsys.readouterr().out
    random.seed(4)
    want, score = O.ga_run(seqs, "mo", sd, O.Knobs(POPULATION_SIZE=5, GA_ITERATIONS=1), fast_pareto=True)
    assert out == want
    assert text.count("CURRENT POPULATION (Iteration") == 4          # initial + after mutation, selection, crossover
    assert text.count("Individual 1:") == 4 and "Iteration 1/1..." in text
    assert "HALL OF FAME - BEST INDIVIDUAL SO FAR:" in text and "Population entirely printed." in text
    last = text[text.rindex("HALL OF FAME"):]
    for row in want:
        assert row in last
    assert "SP Score: %s" % (score[1],) in last and "CS Score: %s" % (score[2],) in last
    g2 = ga_mod.GA(seqs, "sp")
    g2.print_hall_of_fame()
    g2.print_population()
    text = capsys.readouterr().out
    assert "Hall of Fame is empty!" in text and "Empty Population!" in text


# ------------------------------------------------------------------------------------------- multi-GPU (section 8(e))
def _run_sharded_check(world):
    import subprocess
    import sys
    from conftest import ROOT
    script = os.path.join(ROOT, "scripts", "sharded_check.py")
    if world == 1:
        cmd = [sys.executable, script]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), script]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "SHARDED-OK world=%d" % world in out.stdout, (out.stdout[-2000:], out.stderr[-4000:])


def test_sharded_loop_single_rank_reproduces_reference_traces(gpu):
    """The sharded code path (logical indices, gathered-score ranking, gad_crossover_from) with one rank."""
    _run_sharded_check(1)


def test_sharded_loop_two_gpus_reproduces_reference_traces(gpu):
    """Two ranks over NCCL: same per-pass scores, hall of fame and result as the reference for every trace."""
    if gpu.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _run_sharded_check(2)


# ------------------------------------------------------------------------------------------- callers (section 8(f) rank 2)
def test_inference_driver_writes_reference_report_format(gpu, weight_files, cfg, datasets_json, tmp_path):
    """mainGA.inference over a dataset module: the report blocks parse with the golden-report parser and
    their SP / EM / AL recompute from the alignment with the oracle (the 1340-record format)."""
    import types
    import mainGA
    name, sd = weight_files[(3, 30)]
    sets = datasets_json["dataset1_6x30bp"]
    mod = types.SimpleNamespace(file_name="unit_6x30bp.py", datasets=list(sets)[:3], **{k: sets[k] for k in list(sets)[:3]})
    cfg.POPULATION_SIZE, cfg.GA_ITERATIONS = 8, 2
    cfg.AGENT_WINDOW_ROW, cfg.AGENT_WINDOW_COLUMN = 3, 30
    for mode in ("sp", "mo"):
        random.seed(1)
        report, csv_path = mainGA.inference(mode, mod, model_path=name)
        blocks = open(report).read().split("File: ")[1:]
        assert len(blocks) == 3
        for blk, key in zip(blocks, mod.datasets):
            lines = blk.strip("\n").split("\n")
            assert lines[0] == key
            head = dict(ln.rsplit(":", 1) for ln in lines[1:6])
            rows = [ln.strip() for ln in lines[lines.index("Alignment:") + 1:] if ln.strip()]
            board = O.encode(rows)
            R, W = len(board), len(board[0])
            assert int(head["Number of Sequences (QTY)"]) == R == 6
            assert int(head["Alignment Length (AL)"]) == W
            assert int(head["Sum of Pairs (SP)"]) == O.sum_of_pairs(board, 0, R, 0, W)
            assert int(head["Exact Matches (EM)"]) == O.exact_columns(board, 0, R, 0, W)
            assert head["Column Score (CS)"].strip() == "%.3f" % O.column_score(board, 0, R, 0, W)
            assert [r.replace("-", "") for r in rows] == [s.replace("-", "") for s in sets[key]]   # residues preserved
        assert len(open(csv_path).read().strip().split("\n")) == 4
    # the oracle's loop gives the same winning alignment for the first set
    random.seed(1)
    kb = O.Knobs(POPULATION_SIZE=8, GA_ITERATIONS=2)
    want, _ = O.ga_run(sets[mod.datasets[0]], "sp", sd, kb)
    random.seed(1)
    import GA as ga_mod
    assert ga_mod.GA(sets[mod.datasets[0]], "sp").run(name) == want


# ------------------------------------------------------------------------------------------- BASELINE sizes: properties
def _synthetic_population(R, C, P, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(1, 5, size=(R, C), dtype=np.uint8)
    gaps = max(1, round(C * 0.05))
    W = C + gaps
    stride = -(-W // 16) * 16
    boards = np.full((P, R, stride), 5, np.uint8)
    rowlen = np.full((P, R), C, np.int32)
    boards[:, :, :C] = base[None]
    n_clone = P // 4
    pos = rng.integers(0, C, size=(P - n_clone, R, gaps))
    for g in range(gaps):                                  # insert one gap per pass (vectorised over boards and rows)
        p = pos[:, :, g][:, :, None]
        idx = np.arange(stride)[None, None, :]
        sub = boards[n_clone:]
        shifted = np.concatenate((sub[:, :, :1], sub[:, :, :-1]), axis=2)
        boards[n_clone:] = np.where(idx < p, sub, np.where(idx == p, 5, shifted))
        rowlen[n_clone:] += 1
    return boards, rowlen


@pytest.mark.parametrize("R,C,P,rw,cw,mode", [(6, 60, 65536, 6, 30, "mo"), (20, 200, 16384, 3, 30, "cs"),
                                              (64, 1000, 1024, 3, 30, "sp")])
def test_generation_properties_at_baseline_shapes(gpu, weight_files, cfg, R, C, P, rw, cw, mode):
    """One whole generation at the BASELINE board shapes (c3 at full population; c4/c5 shapes at a reduced
    population), checked through size-independent properties: every board keeps each row's residues in order
    (mutation and crossover only move gaps / recombine whole rows of boards with the same residues), the
    population size is restored, scores of a sample agree with the oracle, the hall of fame dominates."""
    from population import DevicePopulation
    from DPAMSA.dqn import load_qnet
    name, _ = weight_files[(rw, cw)]
    boards, rowlen = _synthetic_population(R, C, P, R * 7 + C)
    residues = [row[row != 5].tolist() for row in boards[0]]
    extra = (rw - 1) * cw
    wide = np.full((P, R, -(-(boards.shape[2] + extra) // 16) * 16), 5, np.uint8)
    wide[:, :, :boards.shape[2]] = boards
    dev = DevicePopulation.from_numpy(wide, rowlen, capacity=P)
    net = load_qnet(name, rw, cw)
    w = WEIGHTS
    dev.fitness(*w)
    dev.hof_update(mode)
    n_mut = round(P * 0.25)
    idx = dev.rank_order(mode, n_mut)
    tile = dev.worst_tiles(idx, rw, cw, mode, *w)
    assert int(tile.min().item()) >= 0
    dev.mutate(net, idx, tile, cw * (rw * (rw - 1) / 2 * 4))
    dev.fitness(*w)
    dev.hof_update(mode)
    n_keep = dev.select(mode, P // 2)
    assert (n_keep == P // 2) if mode != "mo" else (n_keep >= 1)
    dev.fitness(*w)
    dev.hof_update(mode)
    random.seed(9)
    plan = np.array(O.crossover_plan(n_keep, R, P - n_keep), dtype=np.int32).reshape(-1, 3)
    dev.crossover(plan, 0)
    sp, em, al = [t.clone() for t in dev.fitness(*w)]
    dev.hof_update(mode)
    assert dev.n == P
    b, r = dev.to_numpy()
    rng = np.random.default_rng(1)
    for p in rng.choice(P, 64, replace=False):
        assert len(set(r[p].tolist())) == 1                                   # rectangular after the pass
        for k in range(R):
            row = b[p, k, :r[p, k]]
            assert row[row != 5].tolist() == residues[k]                      # residues preserved, in order
        board = [b[p, k, :r[p, k]].tolist() for k in range(R)]
        assert int(sp[p]) == O.sum_of_pairs(board, 0, R, 0, len(board[0]))
        assert int(em[p]) == O.exact_columns(board, 0, R, 0, len(board[0]))
        assert not any(all(row[c] == 5 for row in board) for c in range(len(board[0])))   # no all-gap column left
    hb, (hidx, hsp, hem, hal) = dev.hof_numpy()
    if mode == "sp":
        assert hsp >= int(sp.max())
    elif mode == "cs":
        assert hem / hal >= float((em.double() / al.double()).max())
    for k in range(R):
        assert hb[k][hb[k] != 5].tolist() == residues[k]


def test_benchmark_callers_thin_host_versions(gpu, weight_files, cfg, datasets_json, tmp_path, monkeypatch):
    """run_benchmarks.py's other entry points: plain DPAMSA inference (greedy episode per whole set, against the
    oracle's episode), an external tool run (a fake tool that echoes the FASTA), the menu and the summary."""
    import types
    import utils
    name, sd = weight_files[(3, 30)]
    sets = datasets_json["synthetic_dataset_3x30bp"]
    keys = list(sets)[:2]
    mod = types.SimpleNamespace(file_name="unit_3x30bp.py", datasets=keys, **{k: sets[k] for k in keys})
    csv_path = utils.run_dpamsa_inference(mod, "unit_3x30bp", name)
    lines = open(csv_path).read().strip().split("\n")
    assert len(lines) == 3
    for key, line in zip(keys, lines[1:]):
        board = O.encode(sets[key])
        O.mutate_board(board, (0, 3, 0, 30), sd, qdtype=np.float64)        # one greedy episode over the whole board
        R, W = len(board), len(board[0])
        got = line.split(",")
        assert got[0] == key and int(got[1]) == R and int(got[2]) == W
        assert int(got[3]) == O.sum_of_pairs(board, 0, R, 0, W) and int(got[4]) == O.exact_columns(board, 0, R, 0, W)
    # external tool plumbing with a stand-in command
    fasta = tmp_path / "set0.fasta"
    fasta.write_text(">a\nACGT-A\n>b\nACGTTA\n>c\nAC-TTA\n")
    monkeypatch.setitem(cfg.TOOLS, "Echo", {"command": lambda src, dst: ["cat", src],
                                            "output_dir": str(tmp_path / "out"), "report_dir": str(tmp_path / "rep")})
    rows = utils.run_tool_and_generate_report("Echo", [str(fasta)], "unit")
    board = O.encode(["ACGT-A", "ACGTTA", "AC-TTA"])
    assert rows == [["set0.fasta", 3, 6, O.sum_of_pairs(board, 0, 3, 0, 6), O.exact_columns(board, 0, 3, 0, 6),
                     O.column_score(board, 0, 3, 0, 6)]]
    assert "File: set0.fasta" in open(tmp_path / "rep" / "unit.txt").read()
    tool_csv = utils.save_inference_csv(rows, "Echo", "unit")
    summary = utils.plot_metrics({"Echo": tool_csv, "DPAMSA": csv_path}, "unit")
    assert set(summary) == {"Echo", "DPAMSA"} and summary["Echo"]["SP"] == rows[0][3]
    answers = iter(["x", "7", "2"])
    monkeypatch.setattr("builtins.input", lambda _prompt: next(answers))
    assert utils.display_menu() == 2


def test_training_loop_end_to_end(gpu, monkeypatch, tmp_path):
    """DPAMSA.main.train (reference main.py:95-235) on a tiny set: episodes run through the device-scored environment,
    the optimiser steps once the replay memory holds a batch, the saved file loads back into the device Q-net, and
    the frozen device copy of the freshly trained weights takes the same greedy decisions as the torch network in
    eval mode."""
    import types
    import torch
    import config
    import DPAMSA.main as dmain
    from DPAMSA.dqn import DQN
    from DPAMSA.env import Environment
    for k, v in {"MAX_EPISODE": 4, "BATCH_SIZE": 8, "REPLAY_MEMORY_SIZE": 32, "UPDATE_ITERATION": 4,
                 "DECREMENT_ITERATION": 2, "ALPHA": 1e-4}.items():
        monkeypatch.setattr(config, k, v)
    monkeypatch.setattr(config, "RUNS_PATH", str(tmp_path / "runs"))
    monkeypatch.setattr(config, "DPAMSA_WEIGHTS_PATH", str(tmp_path))
    seqs = ["ACGTAGCT", "AGTCAGT", "ACGAGCT"]
    mod = types.SimpleNamespace(file_name="unit_train.py", datasets=["set0"], set0=seqs)
    random.seed(3)
    np.random.seed(3)
    torch.manual_seed(3)
    history = dmain.train(mod, model_path="unit_trained")
    log = history["set0"]
    assert len(log) == 4 and all(e["steps"] >= 8 for e in log)
    assert any(e["loss"] > 0 for e in log) and all(np.isfinite(e["loss"]) for e in log)
    assert log[-1]["epsilon"] == pytest.approx(config.EPSILON - 2 * config.DELTA)
    assert os.path.isdir(os.path.join(str(tmp_path), "runs", "unit_train", "set0"))
    # the file goes straight back into the generation loop's loader
    env = Environment(seqs)
    agent = DQN(env.action_number, env.row, env.max_len, env.max_len * env.max_reward)
    agent.load("unit_trained")
    state = env.reset()
    q_dev = agent.q_values(state)
    agent._ensure_training()                                   # torch copy of the same file, dropout off
    agent.eval_net.eval()
    with torch.no_grad():
        q_torch = agent.eval_net(torch.LongTensor(state).unsqueeze(0).to(agent.train_device))[0].cpu().numpy()
    assert np.allclose(q_dev, q_torch, rtol=1e-3, atol=1e-3 * agent.max_value * 1e-2)
    assert int(agent.predict(state)[0]) == int(np.argmax(q_torch))


def test_dpamsa_inference_batched_equals_stepwise(gpu, weight_files, datasets_json):
    """DPAMSA.main.inference: all equal-length sets of a module aligned in one lock-step device run give the very
    report and CSV of the reference's one-step-at-a-time loop; a set with sequences of different lengths takes
    that loop in both modes."""
    import types
    import DPAMSA.main as dmain
    name, _ = weight_files[(3, 30)]
    sets = datasets_json["synthetic_dataset_3x30bp"]
    keys = list(sets)[:6]
    data = {k: sets[k] for k in keys}
    ragged = [sets[keys[0]][0], sets[keys[0]][1][:-2], sets[keys[0]][2][:-1]]     # 30, 28, 29 -> still a 3x30 model
    data["ragged"] = ragged
    mod = types.SimpleNamespace(file_name="unit_batched.py", datasets=keys[:3] + ["ragged"] + keys[3:], **data)
    outs = []
    for batched in (False, True):
        report, csv_path = dmain.inference(mod, model_path=name, batched=batched)
        outs.append((open(report).read(), open(csv_path).read()))
    assert outs[0] == outs[1]
    assert outs[0][0].count("File: ") == 7 and len(outs[0][1].strip().split("\n")) == 8
