"""CPython-compatible MT19937 replay (host entry points of libgadpamsa, no GPU needed): crossover plans and
populations - whole and sharded - must equal what CPython's `random` produces through the oracle's
restatement of GA.py:71-94 / GA.py:281-293, and leave the generator in the same state."""
import random

import numpy as np
import pytest

import ga_oracle as O


def mt_state():
    version, words, gauss = random.getstate()
    return np.array(words, dtype=np.uint32), (version, gauss)


def mt_restore(arr, meta):
    random.setstate((meta[0], tuple(int(x) for x in arr), meta[1]))


def host_population(nat, seqs, P, clone_rate, gap_rate, first=0, step=1):
    codes = O.encode(seqs)
    R = len(codes)
    seqlen = np.array([len(s) for s in codes], np.int32)
    ngaps = np.array([max(1, round(len(s) * gap_rate)) for s in codes], np.int32)
    cstride = int(seqlen.max())
    arr = np.full((R, cstride), 5, np.uint8)
    for r, s in enumerate(codes):
        arr[r, :len(s)] = s
    stride = -(-(cstride + int(ngaps.max())) // 16) * 16
    n_own = len(range(first, P, step))
    boards = np.empty((n_own, R, stride), np.uint8)
    rowlen = np.empty((n_own, R), np.int32)
    state, meta = mt_state()
    if step == 1:
        nat.call("gad_mt_population_host", state.ctypes.data, arr.ctypes.data, seqlen.ctypes.data, cstride,
                 ngaps.ctypes.data, R, P, round(P * clone_rate), boards.ctypes.data, rowlen.ctypes.data, stride)
    else:
        nat.call("gad_mt_population_shard_host", state.ctypes.data, arr.ctypes.data, seqlen.ctypes.data, cstride,
                 ngaps.ctypes.data, R, P, round(P * clone_rate), first, step, boards.ctypes.data, rowlen.ctypes.data,
                 stride)
    mt_restore(state, meta)
    return [[boards[p, r, :rowlen[p, r]].tolist() for r in range(R)] for p in range(n_own)], boards, rowlen


@pytest.mark.parametrize("seed,R,C,P,gap_rate", [(0, 3, 30, 9, 0.05), (1, 6, 60, 64, 0.05), (2, 4, 101, 33, 0.08),
                                                 (3, 20, 200, 40, 0.05), (4, 5, 1000, 6, 0.05)])
def test_population_whole_and_sharded(native_lib, seed, R, C, P, gap_rate):
    rng = random.Random(100 + seed)
    seqs = ["".join(rng.choice("ACGT-") for _ in range(C - (r % 3))) for r in range(R)]
    random.seed(seed)
    want = O.generate_population(seqs, P, 0.25, gap_rate)
    after = random.random()
    random.seed(seed)
    got, boards, rowlen = host_population(native_lib, seqs, P, 0.25, gap_rate)
    assert got == want
    assert random.random() == after
    for p in range(P):                                   # layout invariant: gap code beyond every row
        for r in range(R):
            assert (boards[p, r, rowlen[p, r]:] == 5).all()
    for world in (2, 3, 8):
        for rank in range(world):
            random.seed(seed)
            part, _, _ = host_population(native_lib, seqs, P, 0.25, gap_rate, rank, world)
            assert part == want[rank::world]
            assert random.random() == after              # every rank leaves the generator where the reference does


def test_population_blocks_and_worker_threads(native_lib, monkeypatch):
    """The board builders run on worker threads, one block of individuals behind the (sequential) generator: many
    small blocks, every thread count, rows long enough for the run-copy emitter - same population as CPython's."""
    rng = random.Random(7)
    seqs = ["".join(rng.choice("ACGT") for _ in range(300 - 7 * r)) for r in range(8)]
    random.seed(11)
    want = O.generate_population(seqs, 600, 0.25, 0.05)
    after = random.random()
    for threads, block in (("1", "100"), ("3", "100"), ("8", "64"), ("16", "1000")):
        monkeypatch.setenv("GAD_HOST_THREADS", threads)
        monkeypatch.setenv("GAD_HOST_BLOCK", block)
        random.seed(11)
        got, boards, rowlen = host_population(native_lib, seqs, 600, 0.25, 0.05)
        assert got == want
        assert random.random() == after
        assert all((boards[p, r, rowlen[p, r]:] == 5).all() for p in range(0, 600, 37) for r in range(8))


def test_crossover_plan_host(native_lib):
    for seed, (n_par, rows, n_child) in enumerate([(2, 2, 40), (3, 6, 100), (77, 20, 1000), (4096, 64, 5000)]):
        random.seed(seed)
        want = O.crossover_plan(n_par, rows, n_child)
        after = random.random()
        random.seed(seed)
        plan = np.empty((n_child, 3), np.int32)
        state, meta = mt_state()
        native_lib.call("gad_mt_crossover_plan_host", state.ctypes.data, n_par, rows, n_child, None, plan.ctypes.data)
        mt_restore(state, meta)
        assert plan.tolist() == [list(t) for t in want]
        assert random.random() == after
    state, _ = mt_state()
    with pytest.raises(ValueError):                      # the reference would loop forever (GA.py:285-290)
        native_lib.call("gad_mt_crossover_plan_host", state.ctypes.data, 1, 6, 4, None, plan.ctypes.data)
    with pytest.raises(ValueError):
        native_lib.call("gad_mt_crossover_plan_host", state.ctypes.data, 5, 1, 4, None, plan.ctypes.data)


@pytest.mark.parametrize("seed,n_parents,n_rows,n_children,prob", [(0, 2, 2, 50, 0.5), (1, 75, 6, 75, 0.5), (2, 500, 20, 2000, 0.85),
                                                                  (3, 10, 3, 400, 0.0), (4, 10, 3, 400, 1.0)])
def test_crossover_plan_with_probability_matches_python_random(native_lib, seed, n_parents, n_rows, n_children, prob):
    """gad_mt_crossover_plan_prob_host == the oracle's restatement on CPython's `random` (choice, choice, retry on equal
    parents, randint, then random.random() against the probability), horizontal and vertical, generator state included."""
    random.seed(seed)
    want = O.crossover_plan_variant(n_parents, n_rows, n_children, prob, n_rows)
    after = random.random()
    random.seed(seed)
    plan = np.empty((n_children, 3), np.int32)
    state, meta = mt_state()
    native_lib.call("gad_mt_crossover_plan_prob_host", state.ctypes.data, n_parents, n_rows, n_children, None, float(prob),
                    n_rows, plan.ctypes.data)
    mt_restore(state, meta)
    assert [tuple(int(v) for v in row) for row in plan] == want
    assert random.random() == after
    if prob == 0.0:
        assert (plan[:, 2] == n_rows).all()
    if prob == 1.0:
        assert ((plan[:, 2] >= 1) & (plan[:, 2] < n_rows)).all()
    # vertical: the cut comes from the narrower parent's width
    wrng = random.Random(50 + seed)
    widths = np.array([wrng.randint(2, 90) for _ in range(n_parents)], np.int32)
    random.seed(1000 + seed)
    want_v = []
    while len(want_v) < n_children:
        a = random.choice(range(n_parents))
        b = random.choice(range(n_parents))
        if a == b:
            continue
        cut = random.randint(1, int(min(widths[a], widths[b])) - 1)
        if random.random() >= prob:
            cut = 1 << 20
        want_v.append((a, b, cut))
    after = random.random()
    random.seed(1000 + seed)
    state, meta = mt_state()
    native_lib.call("gad_mt_crossover_plan_prob_host", state.ctypes.data, n_parents, n_rows, n_children, widths.ctypes.data,
                    float(prob), 1 << 20, plan.ctypes.data)
    mt_restore(state, meta)
    assert [tuple(int(v) for v in row) for row in plan] == want_v
    assert random.random() == after


@pytest.mark.parametrize("seed,n,prob,tiles", [(0, 40, 0.85, True), (1, 1000, 0.5, False), (2, 300, 1.0, True), (3, 64, 0.0, True),
                                               (4, 5000, 0.85, True)])
def test_mutation_variant_draws_match_python_random(native_lib, seed, n, prob, tiles):
    """gad_mt_mutation_draws_host == python: per selected individual random.random() (only below probability 1), a skip
    at or above the probability, then random.randrange(number of tiles); generator state included."""
    trng = random.Random(70 + seed)
    ntiles = np.array([trng.randint(1, 40) for _ in range(n)], np.int32)
    random.seed(seed)
    want_keep, want_tile = [], []
    for i in range(n):
        if prob < 1.0 and random.random() >= prob:
            want_keep.append(0)
            want_tile.append(-1)
            continue
        want_keep.append(1)
        want_tile.append(random.randrange(int(ntiles[i])) if tiles else -1)
    after = random.random()
    random.seed(seed)
    keep = np.empty(n, np.uint8)
    tile = np.empty(n, np.int32)
    state, meta = mt_state()
    native_lib.call("gad_mt_mutation_draws_host", state.ctypes.data, n, float(prob), ntiles.ctypes.data if tiles else None,
                    keep.ctypes.data, tile.ctypes.data)
    mt_restore(state, meta)
    assert keep.tolist() == want_keep and tile.tolist() == want_tile
    assert random.random() == after
    if tiles:
        bad = ntiles.copy()
        bad[n // 2] = 0
        state, meta = mt_state()
        if prob > 0.0:
            with pytest.raises(ValueError):
                native_lib.call("gad_mt_mutation_draws_host", state.ctypes.data, n, 1.0, bad.ctypes.data, keep.ctypes.data,
                                tile.ctypes.data)
