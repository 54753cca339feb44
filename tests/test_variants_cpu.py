"""The oracle's restatement of the paper-style GA variants (supplement tables A3 / A5; extensions without reference
code, PARITY UNPINNED): what each knob does, and that the defaults draw nothing from `random` beyond the reference's
own draws (the recorded reference traces are replayed with them in test_oracle_pins.py)."""
import copy
import random

import pytest

import ga_oracle as O


def _scores(n, rng, mode):
    if mode == "sp":
        return [(i, rng.randint(-40, 40)) for i in range(n)]
    return [(i, rng.randint(-40, 40), rng.randint(0, 20) / 20) for i in range(n)]


def test_tournament_keeps_top_n_distinct_and_prefers_the_better_ranked():
    rng = random.Random(1)
    for mode in ("sp", "mo"):
        sc = _scores(150, rng, mode)
        keep = O.tournament_keep(sc, mode, 150, 0.45, 0.25, random.Random(2))
        assert len(keep) == int(150 * 0.45) and keep <= set(range(150))
        order = O.rank_best(sc, 150)
        rank = {p: r for r, p in enumerate(order)}
        # a quarter of the pool per tournament: the survivors are far better ranked than a uniform draw would be
        assert sum(rank[p] for p in keep) / len(keep) < 0.6 * (149 / 2)
    # the whole pool in every tournament = truncation by rank
    sc = _scores(40, rng, "sp")
    assert O.tournament_keep(sc, "sp", 40, 0.5, 1.0, random.Random(3)) == set(O.rank_best(sc, 20))


def test_crossover_probability_zero_copies_parent1_and_one_equals_the_reference_plan():
    plan = O.crossover_plan_variant(10, 6, 50, 0.0, 6, random.Random(4))
    assert all(c == 6 for _, _, c in plan)
    p1 = [[1, 2, 3]] * 6
    p2 = [[4, 4, 4]] * 6
    assert O.horizontal_child(p1, p2, 6) == p1
    a, b = random.Random(5), random.Random(5)
    with_prob = O.crossover_plan_variant(10, 6, 50, 1.0, 6, a)
    plain = O.crossover_plan(10, 6, 50, b)
    assert len(with_prob) == len(plain) == 50
    # probability 1 keeps every cut inside the board (one extra random() per child, so the streams differ)
    assert all(1 <= c <= 5 for _, _, c in with_prob)


def test_mutation_probability_and_random_tile():
    rng = random.Random(6)
    seqs = ["".join(rng.choice("ACGT") for _ in range(60)) for _ in range(6)]
    sd = O.synth_qnet_weights(3, 30, 7)
    pop = O.generate_population(seqs, 12, 0.25, 0.05, random.Random(7))
    sc = O.fitness_scores(pop, "sp")
    before = copy.deepcopy(pop)
    O.mutation(pop, sc, "sp", 6, 3, 30, sd, probability=0.0, rng=random.Random(8))
    assert pop == before                                    # nobody mutates, and nothing but random() was drawn
    r = random.Random(9)
    O.mutation(pop, sc, "sp", 6, 3, 30, sd, tile_choice="random", rng=r)
    changed = [i for i in range(12) if pop[i] != before[i]]
    assert changed and set(changed) <= set(O.rank_best(sc, 6))
    probe = random.Random(9)
    tiles = O.tile_grid(6, 60, 3, 30)
    drawn = [tiles[probe.randrange(len(tiles))] for _ in range(6)]
    assert r.getstate() == probe.getstate() and all(t in tiles for t in drawn)     # one randrange per selected individual


def test_defaults_are_the_reference_loop():
    rng = random.Random(10)
    seqs = ["".join(rng.choice("ACGT") for _ in range(30)) for _ in range(3)]
    sd = O.synth_qnet_weights(3, 30, 7)
    a, b = random.Random(11), random.Random(11)
    kb = O.Knobs(POPULATION_SIZE=8, GA_ITERATIONS=2)
    out_a = O.ga_run(seqs, "mo", sd, kb, rng=a)
    explicit = O.Knobs(POPULATION_SIZE=8, GA_ITERATIONS=2, MUTATION_TILE="worst", MUTATION_PROBABILITY=1.0,
                       CROSSOVER_PROBABILITY=1.0, SELECTION_SCHEME="truncation", ELITISM_PROBABILITY=0.0)
    out_b = O.ga_run(seqs, "mo", sd, explicit, rng=b)
    assert out_a == out_b and a.getstate() == b.getstate()


def test_facade_mutation_variant_host_logic_matches_the_python_loop(native_lib):
    """GA._mutation_variant (draws by the library's MT19937 replay, tile coordinates vectorised) against the plain
    python loop over `random` it replaces, on a stand-in population object (no GPU involved: host logic only)."""
    import types

    import numpy as np
    import torch

    import GA as ga_mod

    rng = random.Random(21)
    n_pop, R, rw, cw = 500, 7, 3, 30
    minlen = [rng.randint(30, 200) for _ in range(n_pop)]
    rowlen = torch.tensor([[m + rng.randint(0, 3) if r else m for r in range(R)] for m in minlen], dtype=torch.int32)
    order = list(range(n_pop))
    rng.shuffle(order)
    order = order[:200]
    dev = types.SimpleNamespace(cur=types.SimpleNamespace(rowlen=rowlen), n=n_pop, R=R, device="cpu")
    me = types.SimpleNamespace(_dev=dev)
    for m_prob, random_tile in ((0.85, True), (1.0, True), (0.5, False)):
        random.seed(33)
        kept, tiles = [], []
        for i in order:                                   # the loop the oracle's `mutation` runs
            if m_prob < 1.0 and random.random() >= m_prob:
                continue
            if random_tile:
                grid = O.tile_grid(R, minlen[i], rw, cw)
                fr, _, fc, _ = grid[random.randrange(len(grid))]
                tiles.append([fr, fc])
            kept.append(i)
        after = random.random()
        random.seed(33)
        d_idx, d_tile = ga_mod.GA._mutation_variant(me, torch.tensor(order, dtype=torch.int32), rw, cw, m_prob, random_tile)
        assert d_idx.dtype == torch.int32 and d_idx.tolist() == kept
        assert (d_tile is None) == (not random_tile)
        if random_tile:
            assert d_tile.dtype == torch.int32 and d_tile.tolist() == tiles
        assert random.random() == after
    random.seed(1)
    assert ga_mod.GA._mutation_variant(me, torch.tensor(order, dtype=torch.int32), rw, cw, 0.0, True) == (None, None)


def test_facade_tournament_selection_host_logic_matches_the_oracle(native_lib, monkeypatch):
    """GA._tournament_selection on a stand-in population whose ranking comes from the oracle: the survivor set and
    the generator state equal oracle.tournament_keep's."""
    import types

    import torch

    import config
    import GA as ga_mod

    rng = random.Random(5)
    for mode in ("sp", "mo"):
        n = 150
        sc = _scores(n, rng, mode)
        got = {}
        dev = types.SimpleNamespace(n=n, device="cpu",
                                    rank_order=lambda m, k, sc=sc: torch.tensor(O.rank_best(sc, k), dtype=torch.int32),
                                    compact_with=lambda keep: got.update(keep=keep.clone()))
        me = types.SimpleNamespace(_dev=dev, mode=mode)
        monkeypatch.setattr(config, "TOURNAMENT_RATE", 0.25)
        top_n = int(n * 0.45)
        random.seed(8)
        ga_mod.GA._tournament_selection(me, top_n)
        after = random.random()
        random.seed(8)
        want = O.tournament_keep(sc, mode, n, 0.45, 0.25)
        assert set(torch.nonzero(got["keep"]).flatten().tolist()) == want
        assert random.random() == after


def test_facade_elitism_host_logic(native_lib, monkeypatch):
    """GA._elitism on a stand-in population: one random.random() per generation (none at probability 0); below the
    probability the hall-of-fame board and width overwrite the worst-ranked individual and the population is scored
    again."""
    import types

    import torch

    import config
    import GA as ga_mod

    n, R, stride = 6, 3, 16
    boards = torch.arange(n * R * stride, dtype=torch.int64).reshape(n, R, stride).to(torch.uint8)
    rowlen = torch.full((n, R), 12, dtype=torch.int32)
    hof_board = torch.full((R, stride), 7, dtype=torch.uint8)
    order = torch.tensor([2, 0, 5, 1, 3, 4], dtype=torch.int32)             # worst-ranked individual: 4
    calls = []
    dev = types.SimpleNamespace(n=n, cur=types.SimpleNamespace(boards=boards.clone(), rowlen=rowlen.clone()),
                                hof=torch.tensor([1, 2, 40, 9, 10, 0, 0, 0], dtype=torch.int32), hof_board=hof_board,
                                rank_order=lambda mode, k: order[:k])
    me = types.SimpleNamespace(_dev=dev, mode="sp", calculate_fitness_score=lambda: calls.append(1))
    monkeypatch.setattr(config, "ELITISM_PROBABILITY", 0.0)
    before = random.getstate()
    ga_mod.GA._elitism(me)
    assert random.getstate() == before and not calls and torch.equal(dev.cur.boards, boards)
    monkeypatch.setattr(config, "ELITISM_PROBABILITY", 0.4)
    hits = 0
    for seed in range(12):
        dev.cur.boards.copy_(boards)
        dev.cur.rowlen.copy_(rowlen)
        calls.clear()
        random.seed(seed)
        expect = random.random() < 0.4
        after = random.random()
        random.seed(seed)
        ga_mod.GA._elitism(me)
        assert random.random() == after
        assert bool(calls) == expect
        if expect:
            hits += 1
            assert torch.equal(dev.cur.boards[4], hof_board) and dev.cur.rowlen[4].tolist() == [10] * R
            keep = [i for i in range(n) if i != 4]
            assert torch.equal(dev.cur.boards[keep], boards[keep])
        else:
            assert torch.equal(dev.cur.boards, boards)
    assert 0 < hits < 12


def test_facade_crossover_plan_with_probability_matches_the_oracle(native_lib):
    """GA._plan_with_probability (the library's replay of CPython's generator) == oracle.crossover_plan_variant on
    `random`, generator state included."""
    import numpy as np

    import GA as ga_mod

    for seed, (n_parents, n_rows, n_children, prob) in enumerate([(30, 6, 30, 0.5), (75, 6, 75, 0.85), (4, 2, 200, 0.1)]):
        random.seed(seed)
        want = O.crossover_plan_variant(n_parents, n_rows, n_children, prob, n_rows)
        after = random.random()
        random.seed(seed)
        plan = ga_mod.GA._plan_with_probability(n_parents, n_rows, n_children, prob, None, 0)
        assert plan.dtype == np.int32 and [tuple(int(v) for v in r) for r in plan] == want
        assert random.random() == after
    with pytest.raises(ValueError):
        ga_mod.GA._plan_with_probability(1, 6, 4, 0.5, None, 0)
