"""GA.generate_population on the device (csrc/mt_device.cu) against the host replay of CPython's `random`
(gad_mt_population_host, itself pinned to CPython in tests/test_mt_host_cpu.py) and against the oracle's
restatement of reference GA.py:71-94: same boards, same row lengths, same generator state afterwards - for every
board shape of BASELINE.json, ragged sequences, sequences that already contain gaps, sharded ownership, and
generator segments so small that a population spans hundreds of them."""
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


def make_ga(seqs, P, clone_rate=0.25, gap_rate=0.05):
    import config
    import GA as ga_mod
    config.POPULATION_SIZE, config.CLONE_RATE, config.GAP_RATE = P, clone_rate, gap_rate
    return ga_mod.GA(seqs, "sp")


def sequences(rng, rows, cols, ragged=False, with_gaps=False):
    out = []
    for _ in range(rows):
        n = cols - (rng.randint(0, 7) if ragged else 0)
        alphabet = "ACGT-" if with_gaps else "ACGT"
        out.append("".join(rng.choice(alphabet) for _ in range(n)))
    return out


@pytest.fixture()
def knobs():
    import config
    saved = {k: getattr(config, k) for k in ("POPULATION_SIZE", "CLONE_RATE", "GAP_RATE")}
    yield
    for k, v in saved.items():
        setattr(config, k, v)


@pytest.mark.parametrize("rows,cols,P,ragged,with_gaps,own", [
    (6, 30, 5, False, False, (0, 1)),          # c1: the reference's default run
    (3, 30, 1024, False, True, (0, 1)),        # c2, inputs with '-' inside
    (6, 60, 65536, False, False, (0, 1)),      # c3 at full population
    (6, 63, 4099, True, True, (0, 1)),         # ragged rows: the draw pattern has the full period
    (20, 200, 8192, False, False, (0, 1)),     # c4 board shape
    (64, 1000, 96, False, False, (0, 1)),      # c5 board shape: 50 gaps per row (two gap registers per lane)
    (64, 1000, 96, True, False, (1, 8)),       # ... ragged, and only the boards of rank 1 of 8
    (6, 60, 10000, False, False, (3, 4)),      # sharded ownership
    (4, 101, 777, False, False, (0, 1)),       # ENCODE-like 4x101
])
def test_device_population_equals_host_replay(gpu, knobs, monkeypatch, rows, cols, P, ragged, with_gaps, own):
    rng = random.Random(rows * 1000 + cols + P)
    seqs = sequences(rng, rows, cols, ragged, with_gaps)
    for segment in (None, "8192"):                       # default segments, and hundreds of tiny ones
        if segment and P * rows * cols > 3_000_000:
            continue
        if segment:
            monkeypatch.setenv("GAD_POPGEN_SEGMENT", segment)
        g = make_ga(seqs, P)
        random.seed(P + 7)
        dev_boards, dev_rowlen = g._generate_device(*own)
        after_dev = random.random()
        random.seed(P + 7)
        host_boards, host_rowlen = g._generate_host(*own, min_stride=dev_boards.shape[2])
        after_host = random.random()
        assert after_dev == after_host                   # the generator was advanced by exactly the same words
        assert np.array_equal(dev_rowlen.cpu().numpy(), host_rowlen)
        assert np.array_equal(dev_boards.cpu().numpy(), host_boards)
        monkeypatch.delenv("GAD_POPGEN_SEGMENT", raising=False)


def test_generate_population_against_the_oracle(gpu, knobs):
    """Through the facade (GA.generate_population -> device generator) against the oracle's python restatement
    with the same seed; then the stream continues identically (the crossover plan that follows agrees)."""
    rng = random.Random(3)
    seqs = sequences(rng, 6, 30)
    g = make_ga(seqs, 333)
    random.seed(11)
    g.generate_population()
    got = g.population
    tail = [random.random() for _ in range(3)]
    py = random.Random(11)
    want = O.generate_population(seqs, 333, 0.25, 0.05, py)
    assert got == want
    assert tail == [py.random() for _ in range(3)]


def test_host_generator_still_selectable(gpu, knobs, monkeypatch):
    monkeypatch.setenv("GAD_POPGEN", "host")
    rng = random.Random(5)
    seqs = sequences(rng, 6, 60)
    g = make_ga(seqs, 500)
    assert g._generate_device() is None
    random.seed(2)
    g.generate_population()
    py = random.Random(2)
    assert g.population == O.generate_population(seqs, 500, 0.25, 0.05, py)
