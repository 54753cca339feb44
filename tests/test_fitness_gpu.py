"""CUDA fitness pass (gad_fitness / gad_fitness_host / gad_window_score_host through the C-ABI)
against the oracle and the reference's published golden vectors. Bit-exact integer parity."""
import copy
import random

import numpy as np
import pytest

import ga_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(native_lib):
    native_lib.require_device()
    import torch
    torch.cuda.set_device(0)
    return torch


def run_device(pop, weights=(4, -4, -4), stride=None):
    from population import DevicePopulation
    import torch
    dev = DevicePopulation.from_lists(pop, stride=stride)
    sp, em, al = dev.fitness(*weights)
    torch.cuda.synchronize()
    return sp.cpu().numpy(), em.cpu().numpy(), al.cpu().numpy(), dev


def run_oracle(pop, weights=(4, -4, -4), stride=None):
    boards, rowlen = O.to_arrays(pop, stride=stride)
    sp, em, al = O.fitness_population_c(boards, rowlen, O.Weights(*weights))
    return sp, em, al, O.to_lists(boards, rowlen)


def assert_parity(pop, weights=(4, -4, -4), stride=None):
    sp, em, al, dev = run_device(copy.deepcopy(pop), weights, stride)
    osp, oem, oal, oboards = run_oracle(copy.deepcopy(pop), weights, dev.stride)
    assert np.array_equal(sp, osp)
    assert np.array_equal(em, oem)
    assert np.array_equal(al, oal)
    assert dev.to_lists() == oboards
    assert dev.widest() == (int(oal.max()) if len(oal) else 0)
    return sp, em, al


def test_golden_report_vectors(gpu, report_vectors):
    """All 1340 published alignments: device SP / EM / AL equal the recorded values."""
    by_rows = {}
    for rec in report_vectors:
        by_rows.setdefault(len(rec["rows"]), []).append(rec)
    checked = 0
    for R, recs in by_rows.items():
        pop = [O.encode([s.upper() for s in rec["rows"]]) for rec in recs]
        sp, em, al = assert_parity(pop)
        for i, rec in enumerate(recs):
            board = pop[i]
            if any(all(row[c] == 5 for row in board) for c in range(len(board[0]))):
                continue
            assert (sp[i], em[i], al[i]) == (rec["sp"], rec["em"], rec["al"])
            checked += 1
    assert checked > 1300


@pytest.mark.parametrize("R,Wmax,stride", [(2, 9, None), (3, 31, None), (6, 63, 64), (6, 63, 128), (4, 101, None),
                                           (12, 150, None), (20, 210, None), (5, 300, 512), (64, 130, None)])
def test_random_ragged_boards(gpu, R, Wmax, stride):
    rng = random.Random(R * 1000 + Wmax)
    pop = []
    for _ in range(257):
        W = rng.randint(1, Wmax)
        b = [[rng.choice([1, 2, 3, 4, 5, 5]) for _ in range(rng.randint(max(0, W - 7), W))] for _ in range(R)]
        if rng.random() < 0.5:
            for c in rng.sample(range(W), min(W, rng.randint(1, 6))):
                for row in b:
                    if c < len(row):
                        row[c] = 5
        pop.append(b)
    assert_parity(pop, stride=stride)
    assert_parity(pop, weights=(5, -3, -7), stride=stride)      # the three pair classes are separated


def test_edge_cases(gpu):
    pop = [
        [[], [], []],                                       # empty board
        [[5, 5, 5], [5, 5], [5]],                           # only gaps -> cleans to width 0
        [[1], [1], [1]],                                    # single exact column
        [[1, 2, 3, 4], [], [5, 5, 5, 5, 5, 5, 1]],          # very ragged
        [[5, 1, 5, 5, 2, 5], [5, 1, 5, 5, 3, 5], [5, 5, 5, 5, 4, 5]],
        [[1, 2, 3, 4, 1, 2, 3, 4, 1, 2, 3, 4, 1, 2, 3, 4]] * 3,
    ]
    sp, em, al = assert_parity(pop)
    assert al.tolist() == [0, 0, 1, 5, 2, 16]      # board 3: columns 4 and 5 are all-gap after padding
    assert em.tolist()[2] == 1 and em.tolist()[5] == 16
    assert sp.tolist()[5] == 16 * 3 * 4


def test_single_row_and_wide_rows(gpu):
    rng = random.Random(5)
    pop = [[[rng.choice([1, 2, 3, 4, 5]) for _ in range(rng.randint(1, 1100))]] for _ in range(40)]
    assert_parity(pop)
    pop = [[[rng.choice([1, 2, 3, 4, 5]) for _ in range(1050)] for _ in range(64)] for _ in range(6)]
    assert_parity(pop)


def test_clean_boards_are_left_untouched(gpu):
    """A second pass over already-clean boards is idempotent (GA.selection re-evaluates survivors)."""
    import torch
    rng = random.Random(9)
    pop = [[[rng.choice([1, 2, 3, 4]) for _ in range(60)] for _ in range(6)] for _ in range(100)]
    sp, em, al, dev = run_device(pop)
    before = dev.boards.clone()
    sp2, em2, al2 = dev.fitness(4, -4, -4)
    torch.cuda.synchronize()
    assert torch.equal(before, dev.boards)
    assert np.array_equal(sp, sp2.cpu().numpy()) and np.array_equal(al, al2.cpu().numpy())


def test_host_buffer_entry_point(gpu, native_lib):
    from population import pack_boards, unpack_boards
    rng = random.Random(21)
    pop = [[[rng.choice([1, 2, 3, 4, 5, 5]) for _ in range(rng.randint(20, 33))] for _ in range(4)] for _ in range(500)]
    boards, rowlen = pack_boards(copy.deepcopy(pop))
    P, R, stride = boards.shape
    sp = np.zeros(P, np.int32)
    em = np.zeros(P, np.int32)
    al = np.zeros(P, np.int32)
    native_lib.call("gad_fitness_host", boards.ctypes.data, rowlen.ctypes.data, P, R, stride, 4, -4, -4,
                    sp.ctypes.data, em.ctypes.data, al.ctypes.data, 1)
    osp, oem, oal, oboards = run_oracle(pop, stride=stride)
    assert np.array_equal(sp, osp) and np.array_equal(em, oem) and np.array_equal(al, oal)
    assert unpack_boards(boards, rowlen) == oboards


def test_facade_free_functions(gpu):
    """utils.get_sum_of_pairs / get_column_score / clean_unnecessary_gaps keep the reference's
    signatures and return types (utils.py:27-128, 390-428)."""
    import utils
    import config
    rng = random.Random(4)
    for _ in range(40):
        R, W = rng.randint(2, 8), rng.randint(1, 70)
        b = [[rng.choice([1, 2, 3, 4, 5, 5]) for _ in range(W)] for _ in range(R)]
        fr = rng.randint(0, R - 1)
        tr = rng.randint(fr + 1, R)
        fc = rng.randint(0, W - 1)
        tc = rng.randint(fc, W)
        sp = utils.get_sum_of_pairs(b, fr, tr, fc, tc)
        cs = utils.get_column_score(b, fr, tr, fc, tc)
        assert isinstance(sp, int) and sp == O.sum_of_pairs(b, fr, tr, fc, tc)
        assert cs == O.column_score(b, fr, tr, fc, tc)
        b1, b2 = copy.deepcopy(b), copy.deepcopy(b)
        utils.clean_unnecessary_gaps(b1)
        O.clean_unnecessary_gaps(b2)
        assert b1 == b2
    old = (config.MATCH_REWARD, config.MISMATCH_PENALTY, config.GAP_PENALTY)
    try:
        config.MATCH_REWARD, config.MISMATCH_PENALTY, config.GAP_PENALTY = 5, -3, -7   # read live
        b = [[1, 2, 5, 4], [1, 3, 5, 4], [2, 3, 1, 5]]
        assert utils.get_sum_of_pairs(b, 0, 3, 0, 4) == O.sum_of_pairs(b, 0, 3, 0, 4, O.Weights(5, -3, -7))
    finally:
        config.MATCH_REWARD, config.MISMATCH_PENALTY, config.GAP_PENALTY = old
    assert utils.get_nucleotides_seqs([[1, 2, 3, 5], [4, 3, 2, 1]]) == ["ATC-", "GCTA"]
    assert utils.get_column_score([[1, 2, 3, 3]] * 3, 0, 3, 0, 4) == 1.0


def test_full_size_population_properties(gpu):
    """BASELINE config c3 size (P=65,536, 6x60): size-independent properties instead of the oracle -
    (1) row permutation invariance, (2) idempotence of the cleaned boards, (3) a sampled oracle check."""
    import torch
    from population import DevicePopulation
    rng = np.random.default_rng(3)
    P, R, W = 65536, 6, 63
    boards = rng.integers(1, 6, size=(P, R, 64), dtype=np.uint8)
    boards[:, :, W:] = 5
    boards[rng.random((P, 1, 64)).repeat(R, 1) < 0.03] = 5          # some all-gap columns
    rowlen = np.full((P, R), W, np.int32)
    dev = DevicePopulation.from_numpy(boards.copy(), rowlen.copy())
    sp, em, al = [t.clone() for t in dev.fitness(4, -4, -4)]
    perm = DevicePopulation.from_numpy(np.ascontiguousarray(boards[:, ::-1, :]), rowlen.copy())
    sp_p, em_p, al_p = perm.fitness(4, -4, -4)
    torch.cuda.synchronize()
    assert torch.equal(sp, sp_p) and torch.equal(em, em_p) and torch.equal(al, al_p)
    sp2, em2, al2 = dev.fitness(4, -4, -4)
    torch.cuda.synchronize()
    assert torch.equal(sp, sp2) and torch.equal(em, em2) and torch.equal(al, al2)
    idx = rng.choice(P, 2048, replace=False)
    ob, orl = boards[idx].copy(), rowlen[idx].copy()
    osp, oem, oal = O.fitness_population_c(ob, orl)
    assert np.array_equal(sp.cpu().numpy()[idx], osp)
    assert np.array_equal(em.cpu().numpy()[idx], oem)
    assert np.array_equal(al.cpu().numpy()[idx], oal)


def test_randomized_shape_sweep(gpu):
    """Many (rows, width, stride, weights) combinations, including rows > 30 (several nibble-counter epochs),
    rows up to 255, widths past 1000 and boards made mostly of gaps - device == oracle, bit for bit."""
    rng = np.random.default_rng(2025)
    shapes = [(1, 5), (2, 1), (7, 16), (8, 17), (9, 64), (29, 40), (30, 33), (31, 47), (32, 100), (61, 130),
              (64, 1050), (100, 90), (255, 20), (13, 1200), (5, 257)]
    for R, W in shapes:
        P = 40 if R * W < 20000 else 6
        stride = -(-(W + int(rng.integers(0, 40))) // 16) * 16
        boards = np.zeros((P, R, stride), np.uint8)
        rowlen = np.zeros((P, R), np.int32)
        for p in range(P):
            gap_p = rng.choice([0.05, 0.3, 0.9])
            w = int(rng.integers(max(1, W - 20), W + 1))
            for r in range(R):
                n = int(rng.integers(max(0, w - 9), w + 1))
                row = rng.integers(1, 5, size=n).astype(np.uint8)
                row[rng.random(n) < gap_p] = 5
                boards[p, r, :n] = row
                rowlen[p, r] = n
            for c in rng.integers(0, w, size=3):                  # a few all-gap columns
                boards[p, :, c] = np.where(np.arange(R) < R, 5, 5)
        pop = O.to_lists(boards, rowlen)
        for weights in ((4, -4, -4), (7, -2, -11)):
            assert_parity(copy.deepcopy(pop), weights=weights, stride=stride)


def _ragged_population(rng, P, R, W, stride, gap_p=0.2, stale=True):
    """P ragged boards with all-gap columns, laid out with the library's invariant: gap code from each row's
    length up to roundup4(board width); beyond that arbitrary stale bytes (kernels must mask them)."""
    w = rng.integers(max(1, W - 20), W + 1, size=P)
    n = np.minimum(w[:, None], np.maximum(0, w[:, None] - rng.integers(0, 10, size=(P, R))))
    n[np.arange(P), rng.integers(0, R, size=P)] = w                      # at least one row has the full width
    boards = rng.integers(1, 5, size=(P, R, stride), dtype=np.uint8)
    boards[rng.random((P, R, stride)) < gap_p] = 5
    allgap = rng.random((P, 1, stride)) < 0.04
    boards[np.broadcast_to(allgap, boards.shape)] = 5
    col = np.arange(stride)[None, None, :]
    pad_to = ((w + 3) // 4 * 4)[:, None, None]
    boards[(col >= n[:, :, None]) & (col < pad_to)] = 5
    if stale:
        junk = rng.integers(0, 256, size=boards.shape, dtype=np.uint8)
        boards = np.where(col >= pad_to, junk, boards)
    else:
        boards[np.broadcast_to(col >= pad_to, boards.shape)] = 5
    return boards, n.astype(np.int32)


@pytest.mark.parametrize("R,W,stride,hint", [
    (6, 63, 224, None),      # c3 shape: nibble counters, 4 lanes per board, one unit per lane
    (3, 32, 64, None),       # 3x30 data
    (2, 9, 16, None),        # one lane per board
    (12, 150, 208, None),    # 12x150 (supplement table A6), two units per lane: compaction falls back to global memory
    (15, 100, 128, None),    # largest nibble-counter row count
    (16, 100, 128, None),    # smallest byte-counter row count
    (20, 210, 288, None),    # c4 shape
    (40, 250, 256, None),    # widest box (256 columns), several flush epochs
    (6, 120, 224, 63),       # width hint too small: boards wider than the box take the global path inside the launch
    (6, 63, 224, 500),       # hint larger than the stride
])
def test_tma_staged_kernel_against_oracle(gpu, R, W, stride, hint):
    """The TMA-staged fitness kernel (fitness.cu, populations of at least 2 048 boards) on ragged boards with
    all-gap columns and stale bytes beyond the board width: scores, cleaned boards and row lengths == the C oracle.
    Two passes: the second one sees clean boards (nothing to compact) and must change nothing."""
    import torch
    from population import DevicePopulation
    rng = np.random.default_rng(R * 1000 + W)
    P = 4099                                                        # not a multiple of the box: the last box is partial
    boards, rowlen = _ragged_population(rng, P, R, W, stride)
    dev = DevicePopulation.from_numpy(boards.copy(), rowlen.copy())
    if hint is not None:
        dev.width_hint = hint
    for weights in ((4, -4, -4), (7, -2, -11)):
        sp, em, al = [t.cpu().numpy() for t in dev.fitness(*weights)]
        torch.cuda.synchronize()
        ob, orl = boards.copy(), rowlen.copy()
        osp, oem, oal = O.fitness_population_c(ob, orl, O.Weights(*weights))
        assert np.array_equal(sp, osp) and np.array_equal(em, oem) and np.array_equal(al, oal)
        gb, grl = dev.to_numpy()
        assert np.array_equal(grl, orl)
        assert O.to_lists(gb, grl) == O.to_lists(ob, orl)
        assert dev.widest() == int(oal.max())


def test_tma_and_register_kernels_agree_at_baseline_shapes(gpu, monkeypatch):
    """c3 and c4 shaped populations through both fitness kernels (GAD_FITNESS_TMA=0 selects the register-staged
    one): identical scores and boards."""
    import torch
    from population import DevicePopulation
    rng = np.random.default_rng(11)
    for R, W, stride, P in ((6, 63, 224, 65536), (20, 210, 288, 16384)):
        boards, rowlen = _ragged_population(rng, P, R, W, stride, gap_p=0.1)
        out = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("GAD_FITNESS_TMA", flag)
            dev = DevicePopulation.from_numpy(boards.copy(), rowlen.copy())
            sp, em, al = [t.cpu().numpy() for t in dev.fitness(4, -4, -4)]
            torch.cuda.synchronize()
            b, r = dev.to_numpy()
            out[flag] = (sp, em, al, r, O.to_lists(b[:512], r[:512]))
        for x, y in zip(out["1"][:4], out["0"][:4]):
            assert np.array_equal(x, y)
        assert out["1"][4] == out["0"][4]


# ------------------------------------------------------------------ GA.calculate_fitness_score as one launch
def _random_population(rng, P, R, C, ragged, min_flips=0):
    """Boards of R copies of one random sequence with a few random cells (symbols or gaps) changed; a share of them
    ragged and with all-gap columns (the cleaning branch)."""
    base = np.repeat(rng.integers(1, 5, size=(1, C)), R, axis=0)
    pop = []
    for p in range(P):
        b = base.copy()
        flips = rng.integers(0, R * C, size=rng.integers(min_flips, 12))
        b.flat[flips] = rng.integers(1, 6, size=len(flips))
        rows = [list(map(int, r)) for r in b]
        if ragged and p % 5 == 0:
            col = int(rng.integers(0, C))
            for r in rows:
                r.insert(col, 5)                       # an all-gap column: cleaned away by the pass
            rows[int(rng.integers(0, R))].extend([5] * int(rng.integers(1, 4)))
        pop.append(rows)
    return pop


@pytest.mark.parametrize("mode", ["sp", "cs", "mo"])
@pytest.mark.parametrize("P,R,C", [(5000, 6, 60), (2500, 20, 120), (2048, 3, 30), (600, 6, 60)])
def test_fitness_pass_one_launch_equals_two_launches_and_the_oracle(gpu, mode, P, R, C):
    """gad_fitness_pass (the hall-of-fame decision in the tail of the TMA-staged fitness kernel; P = 600 takes the
    two launches) against gad_fitness + gad_hof_update on a copy and against the oracle's GA.update_hall_of_fame,
    three passes in a row: first record, a better population, a worse one (the record keeps its place)."""
    from population import DevicePopulation
    rng = np.random.default_rng(P + R)
    hof = None
    one = two = None
    for rnd in range(3):
        pop = _random_population(rng, P, R, C, ragged=True, min_flips=2 if rnd == 0 else 0)
        if rnd == 2:                                    # everybody worse than the record: rows of gaps and one symbol
            pop = [[[5] * C for _ in range(R)] for _ in range(P)]
            for p, b in enumerate(pop):
                b[p % R][p % C] = 1 + p % 4
        boards, rowlen = O.to_arrays(copy.deepcopy(pop), stride=-(-(C + 8) // 16) * 16)
        if one is None:
            one = DevicePopulation.from_numpy(boards.copy(), rowlen.copy())
            two = DevicePopulation.from_numpy(boards.copy(), rowlen.copy())
        else:
            for dev in (one, two):
                dev.cur.boards[:P].copy_(gpu.from_numpy(boards))
                dev.cur.rowlen[:P].copy_(gpu.from_numpy(rowlen))
        one.fitness_pass(4, -4, -4, mode)
        two.fitness(4, -4, -4)
        two.hof_update(mode)
        gpu.cuda.synchronize()
        osp, oem, oal = O.fitness_population_c(boards, rowlen)          # cleans `boards` in place like the pass
        for a, b2 in ((one.sp, two.sp), (one.em, two.em), (one.al, two.al)):
            assert gpu.equal(a[:P], b2[:P])
        assert np.array_equal(one.sp[:P].cpu().numpy(), osp) and np.array_equal(one.al[:P].cpu().numpy(), oal)
        assert one.to_lists() == two.to_lists() == O.to_lists(boards, rowlen)
        rec1, rec2 = one.hof.cpu().tolist(), two.hof.cpu().tolist()
        assert rec1 == rec2 and rec1[5:] == [0, 0, 0], (rnd, rec1, rec2)
        assert gpu.equal(one.hof_board, two.hof_board)
        cleaned = O.to_lists(boards, rowlen)
        if mode == "sp":
            sc = [(i, int(osp[i])) for i in range(P)]
        elif mode == "cs":
            sc = [(i, int(oem[i]) / int(oal[i]) if oal[i] else 0.0) for i in range(P)]
        else:
            sc = [(i, int(osp[i]), int(oem[i]) / int(oal[i]) if oal[i] else 0.0) for i in range(P)]
        hof = O.hof_update(hof, cleaned, sc)
        got = (rec1[1], rec1[2]) if mode == "sp" else (rec1[1], rec1[3] / rec1[4]) if mode == "cs" else \
            (rec1[1], rec1[2], rec1[3] / rec1[4])
        assert got == tuple(hof[1]), (rnd, got, hof[1])
        board, _ = one.hof_numpy()
        assert board.tolist() == hof[0]
        if rnd == 2:
            assert rec1[1] == P                         # the record won against the whole population
