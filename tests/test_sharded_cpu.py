"""Host logic of the sharded generation loop (ga-dpamsa_b200/sharded.py) on CPU: world_size-2 and -3
gloo process groups, with the oracle plugged in as the compute backend. The sharded run must reproduce
the single-process oracle run (and therefore the reference) exactly - per-pass scores, hall of fame and
the final alignment - which pins the logical-index bookkeeping, the padded all-gathers, the survivor
exchange and the dealing of children. No GPU, no CUDA library call."""
import os
import random
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ga_oracle as O
from conftest import PKG, ROOT


class OracleBackend:
    """CPU stand-in for sharded.DeviceBackend built from the oracle's functions (test infrastructure)."""

    def __init__(self, sd, knobs):
        self.device = torch.device("cpu")
        self.sd = sd
        self.kb = knobs
        self.boards = []          # local list of boards (python lists)
        self.R = 0

    def generate_shard(self, sequences, P, rank, world):
        pop = O.generate_population(sequences, P, self.kb.CLONE_RATE, self.kb.GAP_RATE, random)
        b, r = O.to_arrays(pop[rank::world], stride=64)
        b[b == 0] = 5
        return b, r

    def adopt(self, boards, rowlen, capacity):
        self.boards = O.to_lists(boards, rowlen)
        self.R = boards.shape[1]
        self.stride = 512

    def board(self, local, width):
        return torch.tensor(self.boards[local], dtype=torch.uint8)[:, :width].contiguous()

    def survivors(self):
        b = np.full((len(self.boards), self.R, self.stride), 5, np.uint8)
        r = np.zeros((len(self.boards), self.R), np.int32)
        for p, board in enumerate(self.boards):
            for k, row in enumerate(board):
                b[p, k, :len(row)] = row
                r[p, k] = len(row)
        return torch.from_numpy(b), torch.from_numpy(r)

    def validate_window(self):
        pass

    def fitness(self):
        sc = O.fitness_scores(self.boards, "mo")
        sp = torch.tensor([s for _, s, _ in sc], dtype=torch.int32)
        al = torch.tensor([len(b[0]) for b in self.boards], dtype=torch.int32)
        em = torch.tensor([O.exact_columns(b, 0, len(b), 0, len(b[0])) for b in self.boards], dtype=torch.int32)
        return sp, em, al

    @staticmethod
    def _tuples(scores, n, mode):
        sp, em, al = (t.tolist() for t in scores)
        out = []
        for i in range(n):
            cs = em[i] / al[i] if al[i] > 0 else 0
            out.append((i, sp[i], cs) if mode == "mo" else (i, sp[i] if mode == "sp" else cs))
        return out

    def hof_argbest(self, scores, n, mode, hof_score):
        pool = self._tuples(scores, n, mode)
        if hof_score is not None:
            _, sp, em, al = hof_score
            cs = em / al if al > 0 else 0
            pool.append((n, sp, cs) if mode == "mo" else (n, sp if mode == "sp" else cs))
        return O.rank_best(pool, 1)[0]

    def select_flags(self, scores, n, mode, top_n):
        keep = O.selection_keep(self._tuples(scores, n, mode), mode, self.kb.POPULATION_SIZE,
                                self.kb.SELECTION_RATE, fast=True)
        flags = torch.zeros(n, dtype=torch.uint8)
        flags[sorted(keep)] = 1
        dst = (torch.cumsum(flags.to(torch.int32), 0) - flags.to(torch.int32)).to(torch.int32)
        return flags, dst, len(keep)

    def rank_order(self, scores, n, mode, k):
        return torch.tensor(O.rank_best(self._tuples(scores, n, mode), k), dtype=torch.int32)

    def compact(self, local_keep):
        self.boards = [b for b, k in zip(self.boards, local_keep.tolist()) if k]

    def crossover_plan(self, n_parents, n_children):
        return np.array(O.crossover_plan(n_parents, self.R, n_children, random), dtype=np.int32).reshape(-1, 3)

    def crossover_from(self, all_boards, all_rowlen, pos, plan):
        parents = O.to_lists(all_boards.numpy(), all_rowlen.numpy())      # padded buffer: unused slots have rowlen 0
        for a, b, c in plan.tolist():
            self.boards.append(O.horizontal_child(parents[int(pos[a])], parents[int(pos[b])], c))

    def mutate(self, local_idx, model_path, mode, global_width):
        for i in local_idx.tolist():
            _, tile = O.worst_tile(self.boards[i], mode, self.kb.AGENT_WINDOW_ROW, self.kb.AGENT_WINDOW_COLUMN)
            O.mutate_board(self.boards[i], tile, self.sd)


def _worker(rank, world, port, case, queue):
    sys.path[:0] = [PKG, os.path.join(ROOT, "oracle")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["GADPAMSA_NO_MKDIR"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import config
        import sharded
        seqs, mode, seed, knobs = case
        for k, v in knobs.items():
            setattr(config, k, v)
        kb = O.Knobs(**knobs)
        sd = O.synth_qnet_weights(kb.AGENT_WINDOW_ROW, kb.AGENT_WINDOW_COLUMN, 7)
        random.seed(seed)
        g = sharded.ShardedGA(seqs, mode, OracleBackend(sd, kb))
        log = []
        inner = g.calculate_fitness_score

        def logged():
            inner()
            hof = g.hall_of_fame
            log.append((g.population_score, O.decode(hof[0]), tuple(hof[1])))
        g.calculate_fitness_score = logged
        out = g.run("unused")
        counts = g._counts(int(g.gidx.numel()))
        queue.put((rank, out, log, counts, g.comm_bytes))
    finally:
        dist.destroy_process_group()


def run_sharded(world, case, port):
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(results)


@pytest.mark.parametrize("world,mode,P,port", [(2, "sp", 12, 29611), (2, "mo", 10, 29612), (3, "cs", 11, 29613)])
def test_sharded_run_equals_single_process_oracle(datasets_json, world, mode, P, port):
    seqs = datasets_json["dataset1_6x30bp"]["test2"]
    knobs = {"AGENT_WINDOW_ROW": 3, "AGENT_WINDOW_COLUMN": 30, "GA_ITERATIONS": 2, "POPULATION_SIZE": P,
             "CLONE_RATE": 0.25, "GAP_RATE": 0.05, "SELECTION_RATE": 0.5, "MUTATION_RATE": 0.25}
    seed = 5
    random.seed(seed)
    olog = []
    oout, _ = O.ga_run(seqs, mode, O.synth_qnet_weights(3, 30, 7), O.Knobs(**knobs), log=olog, fast_pareto=True)
    results = run_sharded(world, (seqs, mode, seed, knobs), port)
    for rank, out, log, counts, comm in results:
        assert out == oout
        assert len(log) == len(olog)
        for mine, ref in zip(log, olog):
            assert [tuple(t) for t in mine[0]] == [tuple(t) for t in ref[1]]
            assert mine[1] == ref[2] and tuple(mine[2]) == tuple(ref[3])
        assert sum(counts) == P and max(counts) - min(counts) <= 1        # every rank ends with its share
        assert comm > 0


def test_shard_bounds():
    sys.path.insert(0, PKG)
    os.environ["GADPAMSA_NO_MKDIR"] = "1"
    import sharded
    assert sharded.shard_bounds(10, 3) == [0, 4, 7, 10]
    assert sharded.shard_bounds(8, 8) == list(range(9))
    assert sharded.shard_bounds(1048576, 8)[1] == 131072
