"""Pin the oracle (oracle/ga_oracle.py + ga_oracle.c): against the committed golden fixtures made
from the reference (tests/golden, oracle/make_golden.py), and - when /root/reference exists - against
the unmodified reference run in-process. CPU only."""
import copy
import os
import random

import numpy as np
import pytest

import ga_oracle as O
from conftest import GOLDEN


# ------------------------------------------------------------------------------- golden: scoring
def test_report_vectors_python_oracle(report_vectors):
    assert len(report_vectors) == 1340
    for rec in report_vectors[::7]:
        board = O.encode([s.upper() for s in rec["rows"]])
        R, W = len(board), len(board[0])
        assert O.sum_of_pairs(board, 0, R, 0, W) == rec["sp"]
        assert O.exact_columns(board, 0, R, 0, W) == rec["em"]
        assert W == rec["al"]


def test_report_vectors_c_oracle(report_vectors):
    """All 1340 published alignments through the C restatement, as one ragged population per shape."""
    by_rows = {}
    for rec in report_vectors:
        by_rows.setdefault(len(rec["rows"]), []).append(rec)
    for R, recs in by_rows.items():
        pop = [O.encode([s.upper() for s in rec["rows"]]) for rec in recs]
        boards, rowlen = O.to_arrays(pop)
        sp, em, al = O.fitness_population_c(boards, rowlen, nthreads=2)
        for i, rec in enumerate(recs):
            board = pop[i]
            has_allgap = any(all(row[c] == 5 for row in board) for c in range(len(board[0])))
            if has_allgap:       # the fitness pass cleans first; reports are scored uncleaned
                continue
            assert (sp[i], em[i], al[i]) == (rec["sp"], rec["em"], rec["al"]), rec["name"]


def test_c_oracle_equals_python_oracle_on_ragged_boards():
    rng = random.Random(11)
    pop = []
    for _ in range(300):
        W = rng.randint(1, 50)
        b = [[rng.choice([1, 2, 3, 4, 5, 5]) for _ in range(rng.randint(max(0, W - 5), W))] for _ in range(5)]
        if rng.random() < 0.4:
            for row in b:
                if len(row) > 3:
                    row[2] = 5
        pop.append(b)
    boards, rowlen = O.to_arrays(pop)
    sp, em, al = O.fitness_population_c(boards, rowlen)
    ref_pop = copy.deepcopy(pop)
    scores = O.fitness_scores(ref_pop, "mo")
    assert O.to_lists(boards, rowlen) == ref_pop
    for i, (_, s, c) in enumerate(scores):
        assert s == sp[i]
        assert c == (em[i] / al[i] if al[i] else 0)


# ------------------------------------------------------------------------------- golden: primitives
def test_rank_pareto_selection_hof(primitives):
    for rec in primitives["rank"]:
        assert O.rank_best([tuple(t) for t in rec["scores"]], rec["n"]) == rec["out"]
    for rec in primitives["pareto"]:
        sc = [tuple(t) for t in rec["scores"]]
        assert [t[0] for t in O.pareto_front(sc)] == rec["out"]
        assert [t[0] for t in O.pareto_front_fast(sc)] == rec["out"]
    for rec in primitives["selection"]:
        sc = [tuple(t) for t in rec["scores"]]
        for fast in (False, True):
            keep = O.selection_keep(sc, rec["mode"], rec["n"], rec["rate"], fast=fast)
            assert sorted(keep) == rec["keep"]
    for rec in primitives["hof"]:
        n = len(rec["scores1"])
        hof = O.hof_update(None, [[[i + 1]] for i in range(n)], [tuple(t) for t in rec["scores1"]])
        assert [hof[0], list(hof[1])] == rec["first"]
        hof = O.hof_update(hof, [[[100 + i]] for i in range(n)], [tuple(t) for t in rec["scores2"]])
        assert [hof[0], list(hof[1])] == rec["second"]


def test_clean_tiles_worst(primitives):
    for rec in primitives["clean"]:
        b = copy.deepcopy(rec["in"])
        O.pad_rows(b)
        O.clean_unnecessary_gaps(b)
        assert b == rec["out"]
        boards, rowlen = O.to_arrays([rec["in"]])
        O.fitness_population_c(boards, rowlen)
        assert O.to_lists(boards, rowlen)[0] == rec["out"]
    for rec in primitives["tiles"]:
        assert [list(t) for t in O.tile_grid(rec["R"], rec["W"], rec["rw"], rec["cw"])] == rec["out"]
    for rec in primitives["worst"]:
        score, tile = O.worst_tile(rec["board"], rec["mode"], rec["rw"], rec["cw"])
        assert score == rec["score"] and list(tile) == rec["tile"]
        arr, _ = O.to_arrays([rec["board"]])
        tsp, tem = O.tile_scores_c(arr[0], len(rec["board"][0]), rec["rw"], rec["cw"])
        tiles = O.tile_grid(len(rec["board"]), len(rec["board"][0]), rec["rw"], rec["cw"])
        assert [O.sum_of_pairs(rec["board"], *t) for t in tiles] == tsp.tolist()
        assert [O.exact_columns(rec["board"], *t) for t in tiles] == tem.tolist()


def test_population_and_crossover_rng_replay(primitives):
    for rec in primitives["population"]:
        random.seed(rec["seed"])
        pop = O.generate_population(rec["seqs"], rec["P"], 0.25, 0.05)
        assert pop == rec["population"]
        parents = pop[:rec["n_survivors"]]
        plan = O.crossover_plan(len(parents), len(parents[0]), rec["P"] - len(parents))
        children = [O.horizontal_child(parents[a], parents[b], c) for a, b, c in plan]
        assert parents + children == rec["after_crossover"]
        assert random.random() == rec["next_random"]      # the RNG stream is consumed identically


def test_env_episodes(primitives):
    for rec in primitives["env"]:
        env = O.Env(rec["data"])
        assert env.max_reward == rec["max_reward"] and env.action_number == rec["action_number"]
        assert env.reset() == rec["steps"][0]["state"]
        for st in rec["steps"][1:]:
            r, s, d = env.step(st["action"])
            assert (r, s, d) == (st["reward"], st["state"], st["done"])
            assert env.aligned == st["aligned"]
        env.padding()
        assert env.aligned == rec["padded"]


def test_env_known_answer_trace():
    """SURVEY.md 8(a) a16 probe trace."""
    env = O.Env([[1, 2, 5], [3, 4, 1]])
    assert env.reset() == [1, 2, 5, 5, 3, 4, 1, 5]
    _, s, _ = env.step(0b10)
    assert env.aligned == [[1], [5]] and s == [2, 5, 5, 3, 4, 1, 5, 0]
    env.step(0b10)
    env.step(0b10)
    r, _, d = env.step(0b00)
    assert (r, d) == (-4.0, 0)
    env.padding()
    assert env.aligned == [[1, 2, 5, 5, 5, 5], [5, 5, 5, 3, 4, 1]]


# ------------------------------------------------------------------------------- golden: Q-net and full runs
def test_qnet_golden_vectors():
    z = np.load(os.path.join(GOLDEN, "qnet_vectors.npz"))
    seed = int(z["weight_seed"])
    for tag, rows, cols in (("3x30", 3, 30), ("6x30", 6, 30), ("2x5", 2, 5)):
        sd = O.synth_qnet_weights(rows, cols, seed)
        states, q_ref = z["states_" + tag], z["q_" + tag]
        max_value = cols * rows * (rows - 1) / 2 * 4
        q = O.qnet_forward(sd, states, max_value)
        # tolerance of the north star: 1e-3 relative (plus an absolute floor for Q near 0)
        assert np.all(np.abs(q - q_ref) <= 1e-3 * np.abs(q_ref) + 1e-3 * max_value * 1e-2)
        assert np.array_equal(q.argmax(1), q_ref.argmax(1))


def test_ga_traces_reproduced_by_oracle(ga_traces):
    """The oracle's whole generation loop reproduces the reference's per-pass scores and hall of
    fame for every recorded (mode, seed, knobs) - including greedy Q-net mutation."""
    seed = ga_traces["weight_seed"]
    weights = {}
    done = 0
    for tr in ga_traces["traces"]:
        kb = tr["knobs"]
        if kb["POPULATION_SIZE"] > 64:
            continue                        # the P=1024 trace is replayed by the GPU tests only
        key = (kb["AGENT_WINDOW_ROW"], kb["AGENT_WINDOW_COLUMN"])
        if key not in weights:
            weights[key] = O.synth_qnet_weights(key[0], key[1], seed)
        random.seed(tr["seed"])
        log = []
        out, hof_score = O.ga_run(tr["seqs"], tr["mode"], weights[key], O.Knobs(**kb), log=log)
        assert out == tr["result"]
        assert len(log) == len(tr["passes"])
        for mine, ref in zip(log, tr["passes"]):
            assert [list(s) for s in mine[1]] == ref["scores"]
            assert mine[2] == ref["hof"] and list(mine[3]) == ref["hof_score"]
        done += 1
        if done >= 8:
            break
    assert done >= 8


# ------------------------------------------------------------------------------- live reference (container only)
ref_shim = pytest.importorskip("ref_shim")


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present (GPU box)")
def test_live_reference_scoring_and_tiles():
    import subprocess
    import sys
    code = r'''
import sys, random, copy
sys.path.insert(0, %r)
import ref_shim, ga_oracle as O
ref = ref_shim.load()
rng = random.Random(3)
for _ in range(150):
    R, W = rng.randint(2, 8), rng.randint(1, 40)
    b = [[rng.choice([1, 2, 3, 4, 5, 5]) for _ in range(W)] for _ in range(R)]
    fr = rng.randint(0, R - 1); tr = rng.randint(fr + 1, R); fc = rng.randint(0, W - 1); tc = rng.randint(fc, W)
    assert O.sum_of_pairs(b, fr, tr, fc, tc) == ref.utils.get_sum_of_pairs(b, fr, tr, fc, tc)
    assert O.column_score(b, fr, tr, fc, tc) == ref.utils.get_column_score(b, fr, tr, fc, tc)
    b1, b2 = copy.deepcopy(b), copy.deepcopy(b)
    O.clean_unnecessary_gaps(b1); ref.utils.clean_unnecessary_gaps(b2)
    assert b1 == b2
    rw, cw = rng.randint(1, R), rng.randint(1, W)
    assert O.tile_grid(R, W, rw, cw) == ref.utils.get_all_different_sub_range(b, rw, cw)
print("LIVE-OK")
''' % os.path.dirname(O.__file__)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "LIVE-OK" in out.stdout, out.stderr[-2000:]
