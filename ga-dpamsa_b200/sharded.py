"""The generation loop with the population sharded over the GPUs of one box (one process per GPU,
``torch.distributed``; NCCL on GPUs, gloo in the CPU tests of the host logic).

Every individual carries its *logical* index - the position it has in the reference's single list
(GA.py:54) - so all order-dependent decisions (stable sorts, first-maximum ties, survivors "in original
order", crossover parents) are taken on score arrays arranged by logical index and are bit-identical
to a single-GPU run for any world size. Physical placement is free: survivors stay on the rank that
owns them, children are dealt out to refill every rank to its share.

Per fitness pass : one all-gather of (sp, em, al, logical index) - 16 bytes per individual - and, when
                   the hall of fame changes, one broadcast of the winner's board from its owner.
Per generation   : one all-gather of the survivors' boards ("elite" exchange) so that each rank can
                   build its children from any pair of parents.
Fitness, tile search, Q-net episodes, compaction and crossover itself run locally with no collective.

The algorithm is written against a small backend interface so that the exchange logic can be
exercised on CPU (tests/test_sharded_cpu.py plugs the oracle in); ``DeviceBackend`` is the product.
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch
import torch.distributed as dist

import config

nucleotides_map = {'A': 1, 'T': 2, 'C': 3, 'G': 4, 'a': 1, 't': 2, 'c': 3, 'g': 4, '-': 5}


class _Ticker:
    """Per-step wall times of a phase (GAD_SHARDED_TIMING=1): synchronises after every step, so only for diagnosis."""

    def __init__(self, sink):
        import time
        self.sink, self.time = sink, time
        torch.cuda.synchronize()
        self.t = time.perf_counter()

    def __call__(self, name):
        torch.cuda.synchronize()
        now = self.time.perf_counter()
        self.sink[name] = self.sink.get(name, 0.0) + (now - self.t)
        self.t = now


class _NoTicker:
    def __call__(self, name):
        pass


def shard_bounds(total: int, world: int):
    """Contiguous shares of `total` items: the first total % world ranks get one more."""
    base, extra = divmod(total, world)
    lo = [r * base + min(r, extra) for r in range(world + 1)]
    return lo


class ShardedGA:
    """Same phases and results as ``GA`` (GA.py:38-545) over a sharded population."""

    def __init__(self, sequences, mode, backend, group=None):
        self.sequences = sequences
        self.mode = mode
        self.backend = backend
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.population_size = config.POPULATION_SIZE
        self.current_iteration = 0
        self.n_global = 0
        self.gidx = None             # logical index of every local individual (tensor on backend.device)
        self.g_scores = None         # (sp, em, al) int32 tensors arranged by logical index
        self.owner = None            # rank that owns each logical index
        self.hof_board = None        # replicated on every rank: (boards uint8 [R, w] tensor)
        self.hof_score = None        # (logical index, sp, em, al)
        self.comm_bytes = 0          # bytes this rank received through collectives
        self.timing = {} if os.environ.get("GAD_SHARDED_TIMING") == "1" else None      # per-step wall times (diagnosis)
        self._counts_cache = None    # individuals per rank (device backends: kept on the host between phases)
        self._owner_pattern = None
        self._hof_dev = None         # device backends: hall-of-fame record / board stay on the device
        self._hof_board_dev = None
        self._hof_valid = False

    # ------------------------------------------------------------------ collectives
    def _all_gather_padded(self, t: torch.Tensor, counts, compact=True):
        """All-gather of per-rank tensors whose first dimension differs: pad to the largest and gather into
        one buffer. compact=True returns the concatenation of the valid parts in rank order; compact=False
        returns the padded buffer [world * max(counts), ...] (rank r's rows start at r * max(counts)) and
        saves the second copy - used for the survivors' boards."""
        if self.world == 1:
            return t
        mx = max(counts)
        out = torch.empty((self.world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        if t.shape[0] == mx:
            src = t.contiguous()
        else:
            src = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            src[:t.shape[0]] = t
        dist.all_gather_into_tensor(out, src, group=self.group)
        self.comm_bytes += src.numel() * src.element_size() * (self.world - 1)
        if not compact:
            return out
        return torch.cat([out[r * mx:r * mx + c] for r, c in enumerate(counts)])

    def _counts(self, n_local: int):
        if self.world == 1:
            return [n_local]
        t = torch.tensor([n_local], dtype=torch.int64, device=self.backend.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return [int(o.item()) for o in out]

    # ------------------------------------------------------------------ a3
    def generate_population(self):
        """Every rank replays the same `random` stream (GA.py:71-94) and keeps its slice."""
        P = self.population_size
        lo = shard_bounds(P, self.world)
        # interleaved placement (individual i -> rank i % world): the clones - the first CLONE_RATE of the
        # list and usually the best scored, hence the first to be mutated - spread evenly over the ranks.
        # Every rank consumes the whole random stream but materialises only its own boards.
        mine = np.arange(self.rank, P, self.world)
        if not (hasattr(self.backend, "generate_and_adopt") and
                self.backend.generate_and_adopt(self.sequences, P, self.rank, self.world, lo[1] - lo[0])):
            boards, rowlen = self.backend.generate_shard(self.sequences, P, self.rank, self.world)
            self.backend.adopt(boards, rowlen, capacity=lo[1] - lo[0])
        self.gidx = torch.from_numpy(mine).to(self.backend.device, torch.int64)
        self.n_global = P

    # ------------------------------------------------------------------ a4 + a8
    def calculate_fitness_score(self):
        if getattr(self.backend, "device_pass", False) and self.world > 1:
            return self._fitness_pass_device()
        sp, em, al = self.backend.fitness()                       # local, no collective
        counts = self._counts(int(self.gidx.numel()))
        packed = torch.stack((sp.to(torch.int64), em.to(torch.int64), al.to(torch.int64), self.gidx), dim=1)
        allp = self._all_gather_padded(packed, counts)
        n = self.n_global
        g = allp[:, 3]
        dev = self.backend.device
        self.g_scores = tuple(torch.zeros(n, dtype=torch.int32, device=dev).index_copy_(0, g, allp[:, k].to(torch.int32))
                              for k in range(3))
        owner_of_pos = torch.repeat_interleave(torch.arange(self.world, device=dev),
                                               torch.tensor(counts, device=dev))
        self.owner = torch.zeros(n, dtype=torch.int64, device=dev).index_copy_(0, g, owner_of_pos)
        self.update_hall_of_fame()

    # -- the same pass without a single host synchronisation (device backends) ---------------------------------
    def _rank_counts(self):
        """Individuals per rank, cached: they only change in selection and crossover, which know them."""
        if self._counts_cache is None or sum(self._counts_cache) != self.n_global:
            self._counts_cache = self._counts(int(self.gidx.numel()))
        return self._counts_cache

    def _fitness_pass_device(self):
        """calculate_fitness_score + update_hall_of_fame as a stream of kernels and two collectives, nothing read
        back by the host: (1) the local fitness kernel; (2) ONE all-gather of (sp, em, al, logical index) as int32,
        ranks padded to the largest share with entries that point at a spare slot; (3) a scatter into the
        logical-index order; (4) the hall-of-fame decision on every rank (gad_hof_argbest, winner index stays on
        the device); (5) the winner's board by an all-reduce in which only the owner contributes non-zero bytes
        (a board is ~1 kB: cheaper than learning the owner on the host for a broadcast); (6) a device-side select
        between the new and the old record. Decisions are the same kernels as the single-GPU path."""
        be = self.backend
        dev = be.device
        sp, em, al = be.fitness()
        n, n_loc = self.n_global, int(self.gidx.numel())
        blk = be.pass_block(max(int(config.POPULATION_SIZE), int(self.population_size), n)) \
            if hasattr(be, "pass_block") else None
        if blk is not None:
            return self._fitness_pass_fused(blk, sp, em, al, n, n_loc)
        counts = self._rank_counts()
        mx = max(counts)
        packed = torch.empty((mx, 4), dtype=torch.int32, device=dev)
        packed[:n_loc, 0] = sp
        packed[:n_loc, 1] = em
        packed[:n_loc, 2] = al
        packed[:n_loc, 3] = self.gidx
        if n_loc < mx:
            packed[n_loc:] = torch.tensor([0, 0, 1, n], dtype=torch.int32, device=dev)      # padding -> spare slot n
        allp = torch.empty((self.world * mx, 4), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allp, packed, group=self.group)
        self.comm_bytes += packed.numel() * 4 * (self.world - 1)
        g = allp[:, 3].to(torch.int64)
        buf = torch.zeros((4, n + 1), dtype=torch.int32, device=dev)
        buf[0].index_copy_(0, g, allp[:, 0])
        buf[1].index_copy_(0, g, allp[:, 1])
        buf[2].index_copy_(0, g, allp[:, 2])
        if self._owner_pattern is None or self._owner_pattern.numel() != self.world * mx:
            self._owner_pattern = torch.arange(self.world, device=dev, dtype=torch.int32).repeat_interleave(mx)
        buf[3].index_copy_(0, g, self._owner_pattern)
        self.g_scores = (buf[0, :n], buf[1, :n], buf[2, :n])
        self.owner = buf[3, :n].to(torch.int64)
        # hall of fame, device side
        if self._hof_dev is None:
            self._hof_dev = torch.zeros(8, dtype=torch.int32, device=dev)                  # {valid, index, sp, em, al}
            self._hof_board_dev = torch.full((be.R, be.pop.stride), 5, dtype=torch.uint8, device=dev)
        if self._hof_board_dev.shape[1] != be.pop.stride:                                 # the population was re-laid out
            wide = torch.full((be.R, be.pop.stride), 5, dtype=torch.uint8, device=dev)
            wide[:, :self._hof_board_dev.shape[1]] = self._hof_board_dev
            self._hof_board_dev = wide
        w = be.argbest_device(self.g_scores, n, self.mode, self._hof_dev if self._hof_valid else None)   # int32[1]; == n: the record stays
        self._hof_valid = True
        changed = w < n                                                                    # bool[1]
        wl = torch.clamp(w, max=n - 1).to(torch.int64)
        slot = torch.full((n + 1,), -1, dtype=torch.int64, device=dev)
        slot.index_copy_(0, self.gidx, torch.arange(n_loc, device=dev))
        my_slot = slot.index_select(0, wl)                                                 # -1: another rank owns the winner
        mine = (my_slot >= 0) & changed
        cand = be.pop.cur.boards.index_select(0, torch.clamp(my_slot, min=0))[0] * mine.to(torch.uint8)
        dist.all_reduce(cand, group=self.group)                                            # only the owner's bytes are non-zero
        self.comm_bytes += cand.numel()
        self._hof_board_dev = torch.where(changed, cand, self._hof_board_dev)
        new = torch.stack((torch.ones_like(w), w, buf[0].index_select(0, wl), buf[1].index_select(0, wl),
                           buf[2].index_select(0, wl))).flatten()
        old = self._hof_dev[:5].clone()
        old[1] = n                                                                         # GA.py:127-132: the record re-enters as index n
        self._hof_dev[:5] = torch.where(changed, new, old)
        self.hof_board, self.hof_score = None, None                                        # materialised on demand

    def _fitness_pass_fused(self, blk, sp, em, al, n, n_loc):
        """The pass without a collective call: three launches. (1) the local fitness kernel; (2) gad_pass_scatter -
        this rank's (sp, em, al, owner, slot) stored straight into EVERY rank's score table over NVLink (symmetric
        memory), then its arrival flag raised on every rank; (3) gad_pass_decide - waits for all flags, takes
        GA.update_hall_of_fame's decision on the complete table in one kernel (identically on every rank), and the
        rank that owns the winner stores its board into every rank's hall of fame while the others wait for it.
        The tables ARE the logical-index arrays: no packing, no all-gather, no scatter, nothing read by the host.
        Two tables alternate with the pass number (a rank that runs ahead writes the other one and cannot start
        the pass after that before every rank has arrived in this one)."""
        be = self.backend
        if self._hof_dev is None:
            self._hof_dev = torch.zeros(8, dtype=torch.int32, device=be.device)            # {valid, index, sp, em, al}
        table = blk.run_pass(sp, em, al, self.gidx, n_loc, n, self.mode, self._hof_dev, self._hof_valid, self.rank,
                             self.world)                                                   # int32 [5, cap1]
        self._hof_valid = True
        self._hof_board_dev = blk.hof_board
        self.g_scores = (table[0, :n], table[1, :n], table[2, :n])
        self.owner = table[3, :n]
        self.comm_bytes += 20 * n_loc * (self.world - 1)
        self.hof_board, self.hof_score = None, None                                        # materialised on demand

    def _hof_from_device(self):
        rec = self._hof_dev.cpu().tolist()
        self.hof_score = (rec[1], rec[2], rec[3], rec[4])
        self.hof_board = self._hof_board_dev[:, :rec[4]].clone()

    def update_hall_of_fame(self):
        """GA.py:110-136 on the gathered scores; the winner's board is broadcast by its owner."""
        n = self.n_global
        w = self.backend.hof_argbest(self.g_scores, n, self.mode, self.hof_score)
        if w >= n:                                               # the old hall of fame keeps its place
            self.hof_score = (n,) + tuple(self.hof_score[1:])
            return
        owner = int(self.owner[w].item())
        sp, em, al = (int(t[w].item()) for t in self.g_scores)
        if owner == self.rank:
            local = int((self.gidx == w).nonzero()[0].item())
            board = self.backend.board(local, al)
        else:
            board = torch.empty((self.backend.R, al), dtype=torch.uint8, device=self.backend.device)
        if self.world > 1:
            dist.broadcast(board, src=owner, group=self.group)
            self.comm_bytes += board.numel()
        self.hof_board = board
        self.hof_score = (w, sp, em, al)

    @property
    def hall_of_fame(self):
        if self.hof_board is None and self._hof_dev is not None:
            self._hof_from_device()
        if self.hof_board is None:
            return []
        idx, sp, em, al = self.hof_score
        cs = em / al if al > 0 else 0
        score = (idx, sp, cs) if self.mode == 'mo' else (idx, sp if self.mode == 'sp' else cs)
        return ([row.tolist() for row in self.hof_board.cpu().numpy()], score)

    @property
    def population_score(self):
        sp, em, al = (t.cpu().numpy() for t in self.g_scores)
        out = []
        for i in range(self.n_global):
            cs = int(em[i]) / int(al[i]) if al[i] > 0 else 0
            out.append((i, int(sp[i]), cs) if self.mode == 'mo' else (i, int(sp[i]) if self.mode == 'sp' else cs))
        return out

    # ------------------------------------------------------------------ a10
    def selection(self):
        """GA.py:232-260: the survivor set is decided on the global scores; every rank compacts its own."""
        top_n = int(self.population_size * config.SELECTION_RATE)
        keep, dst, count = self.backend.select_flags(self.g_scores, self.n_global, self.mode, top_n)
        local_keep = keep[self.gidx].to(torch.bool)
        self.backend.compact(local_keep)
        self.gidx = dst[self.gidx][local_keep].to(torch.int64)
        self.n_global = count
        # The crossover plan depends only on the survivor count and the `random` stream, not on any score: start
        # drawing it now (a sequential MT19937 walk over every child of the WHOLE population, on every rank) on a
        # worker thread, under the compaction and the fitness pass that follow. horizontal_crossover joins it.
        if hasattr(self.backend, "crossover_plan_start") and getattr(self, "_plan_job", None) is None:
            self._plan_job = self.backend.crossover_plan_start(count, int(config.POPULATION_SIZE) - count)
        self.calculate_fitness_score()

    # ------------------------------------------------------------------ a12
    def horizontal_crossover(self):
        """GA.py:281-301: one shared plan; survivors' boards all-gathered; each rank builds its block of
        children (logical indices S + block) so that it ends up with its share of the population."""
        P = int(config.POPULATION_SIZE)
        S = self.n_global
        n_children = P - S
        if n_children > 0:
            job, self._plan_job = getattr(self, "_plan_job", None), None
            if job is not None:
                plan = self.backend.crossover_plan_finish(job, S, n_children)
            else:
                plan = self.backend.crossover_plan(S, n_children)        # identical on every rank
            counts = self._counts(int(self.gidx.numel()))
            lo = shard_bounds(P, self.world)
            need = [max(0, (lo[r + 1] - lo[r]) - counts[r]) for r in range(self.world)]
            assert sum(need) == n_children, (need, n_children)
            first = sum(need[:self.rank])
            mine = plan[first:first + need[self.rank]]
            if getattr(self.backend, "peer_ready", False):
                # fused exchange + crossover: only the (rank, slot) address of every survivor is exchanged
                # (8 bytes each); the kernel reads the parents' rows from the owners' buffers over NVLink
                mx = max(counts)
                all_gidx = self._all_gather_padded(self.gidx, counts)
                dev = self.backend.device
                where = torch.cat([torch.arange(r * mx, r * mx + c, device=dev) for r, c in enumerate(counts)])
                pos = torch.zeros(S, dtype=torch.int64, device=dev)
                pos.index_copy_(0, all_gidx, where)
                self.backend.crossover_peer(mx, pos, mine)
                self.comm_bytes += int(mine.shape[0]) * self.backend.R * self.backend.pop.stride * (self.world - 1) // self.world
                kids = torch.arange(S + first, S + first + need[self.rank], dtype=torch.int64, device=dev)
                self.gidx = torch.cat((self.gidx, kids))
                self.n_global = P
                self.calculate_fitness_score()
                return
            boards, rowlen = self.backend.survivors()
            all_boards = self._all_gather_padded(boards, counts, compact=False)      # [world * max, R, stride]
            all_rowlen = self._all_gather_padded(rowlen, counts, compact=False)
            all_gidx = self._all_gather_padded(self.gidx, counts)
            mx = max(counts)
            dev = self.backend.device
            where = torch.cat([torch.arange(r * mx, r * mx + c, device=dev) for r, c in enumerate(counts)]) \
                if self.world > 1 else torch.arange(S, device=dev)
            pos = torch.zeros(S, dtype=torch.int64, device=dev)
            pos.index_copy_(0, all_gidx, where)
            self.backend.crossover_from(all_boards, all_rowlen, pos, mine)
            kids = torch.arange(S + first, S + first + need[self.rank], dtype=torch.int64, device=self.backend.device)
            self.gidx = torch.cat((self.gidx, kids))
            self.n_global = P
        self.calculate_fitness_score()

    # ------------------------------------------------------------------ a14
    def mutation(self, model_path):
        n = min(round(config.POPULATION_SIZE * config.MUTATION_RATE), self.n_global)
        # 'sp' / 'cs': the selection keeps exactly top_n individuals whatever the scores: draw the crossover plan
        # (a sequential MT19937 walk over every child of the whole population, on every rank) under the mutation phase
        self._plan_job = None
        if self.mode in ("sp", "cs") and hasattr(self.backend, "crossover_plan_start"):
            top_n = int(self.population_size * config.SELECTION_RATE)
            if 2 <= top_n <= self.n_global:
                self._plan_job = self.backend.crossover_plan_start(top_n, int(config.POPULATION_SIZE) - top_n)
        if n <= 0:
            return
        order = self.backend.rank_order(self.g_scores, self.n_global, self.mode, n)      # logical indices
        chosen = torch.zeros(self.n_global, dtype=torch.bool, device=self.backend.device)
        chosen[order.to(torch.int64)] = True
        local = chosen[self.gidx].nonzero().flatten().to(torch.int32)
        # the stride must grow identically on every rank (the survivor all-gather needs equal strides)
        width = int(self.g_scores[2].max().item())
        if self.world > 1 and getattr(self.backend, "global_mutation", False):
            self._mutation_global(local, chosen, model_path, width)
        else:
            self.backend.mutate(local, model_path, self.mode, width)

    def _mutation_global(self, local, chosen, model_path, width):
        """The mutation phase with the windows of ALL ranks pooled: every rank extracts the windows of its own
        individuals (180 bytes each for a 6x30 agent), one all-gather makes the pool identical everywhere, every rank
        finds the distinct windows (deterministically: the representative is the smallest index), the u-th distinct
        window runs on rank u % world, and one all-gather of the results (aligned rows, heads, step count) lets every
        rank splice its own individuals. Each distinct window runs ONCE in the whole job - duplicates across ranks
        included - and every rank runs the same number of episodes. Results are those of the single-GPU path: an
        episode is a pure function of its window and Q-values do not depend on the batch (gemm_tc.cu)."""
        be, dev, world, rank = self.backend, self.backend.device, self.world, self.rank
        tick = _Ticker(self.timing) if self.timing is not None else _NoTicker()
        # individuals to mutate per rank, from the replicated ranking (no collective): one host read
        counts = torch.bincount(self.owner[chosen].to(torch.int64), minlength=world).cpu().tolist()
        tick("counts")
        m_loc = int(local.numel())
        assert m_loc == counts[rank]
        if sum(counts) == 0:
            return
        windows, tile = be.mutation_windows(local, model_path, self.mode, width)          # uint8 [m_loc, rows*cols]
        tick("tiles+windows")
        pool = self._all_gather_padded(windows, counts)                                    # [M, rows*cols], rank-major
        tick("gather windows")
        M = int(pool.shape[0])
        rep = be.windows_rep(pool)                                                         # int32 [M]
        flags = rep == torch.arange(M, dtype=torch.int32, device=dev)
        uidx = torch.cumsum(flags.to(torch.int64), 0) - 1                                  # ordinal of every distinct window
        n_unique = int(flags.sum().item())
        per_rank = -(-n_unique // world)
        mine = (flags & (uidx % world == rank)).nonzero().flatten().to(torch.int32)        # ascending: j-th = ordinal rank + j*world
        tick("distinct")
        aligned, heads, steps = be.run_episodes(pool, mine, per_rank, reserve=-(-M // world))
        tick("episodes")
        if world > 1:
            a_all = torch.empty((world * per_rank,) + tuple(aligned.shape[1:]), dtype=torch.uint8, device=dev)
            h_all = torch.empty((world * per_rank,) + tuple(heads.shape[1:]), dtype=torch.uint8, device=dev)
            s_all = torch.empty((world * per_rank,), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(a_all, aligned, group=self.group)
            dist.all_gather_into_tensor(h_all, heads, group=self.group)
            dist.all_gather_into_tensor(s_all, steps, group=self.group)
            self.comm_bytes += (aligned.numel() + heads.numel() + 4 * steps.numel()) * (world - 1)
        else:
            a_all, h_all, s_all = aligned, heads, steps
        tick("gather results")
        off = sum(counts[:rank])
        u = uidx[rep[off:off + m_loc].to(torch.int64)]
        slot = ((u % world) * per_rank + u // world).to(torch.int32)
        be.splice_results(local, tile, windows, slot, a_all, h_all, s_all, width)
        tick("splice")

    # ------------------------------------------------------------------ a19
    def run(self, model_path, debug_mode=False):
        self.generate_population()
        self.calculate_fitness_score()
        self.backend.validate_window()
        for _ in range(config.GA_ITERATIONS):
            self.current_iteration += 1
            self.mutation(model_path)
            self.calculate_fitness_score()
            self.selection()
            self.horizontal_crossover()
        best, _ = self.hall_of_fame
        return ["".join("ATCG-"[v - 1] for v in row) for row in best]


class DeviceBackend:
    """The product backend: libgadpamsa kernels on this rank's GPU."""

    def __init__(self, device=None, peer=None):
        """peer=None: read parents from peer memory when torch's symmetric memory is usable on this group
        (GAD_SHARDED_EXCHANGE=allgather forces the NCCL all-gather of the survivors' boards)."""
        import os
        import _native as nat
        nat.require_device()
        self.nat = nat
        self.device = torch.device(device or config.DEVICE)
        torch.cuda.set_device(self.device)
        self.pop = None
        self.R = 0
        self._rank_tmp = None
        self.peer_ready = False
        self._handles = []
        self._blk = self._blk_hdl = self._blk_peer = self._blk_tables = self.hof_board = None      # the pass block
        self._blk_cap1 = self._blk_pitch = self._blk_R = self._blk_seq = 0
        want = peer if peer is not None else os.environ.get("GAD_SHARDED_EXCHANGE", "peer") != "allgather"
        self._symm = None
        if want and dist.is_initialized() and dist.get_world_size() > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self._symm = symm_mem
            except Exception:                        # pragma: no cover - depends on the torch build
                self._symm = None

    def _symm_alloc(self, shape, dtype):
        return self._symm.empty(*shape, dtype=dtype, device=self.device)

    def _rendezvous(self):
        """Map both population buffers of every rank into this process (collective)."""
        self._handles = []
        world = dist.get_world_size()
        for buf in (self.pop.cur, self.pop.alt):
            hb = self._symm.rendezvous(buf.boards, dist.group.WORLD)
            hr = self._symm.rendezvous(buf.rowlen, dist.group.WORLD)
            self._handles += [hb, hr]
            buf.peer_boards = torch.tensor([int(p) for p in hb.buffer_ptrs[:world]], dtype=torch.int64, device=self.device)
            buf.peer_rowlen = torch.tensor([int(p) for p in hr.buffer_ptrs[:world]], dtype=torch.int64, device=self.device)
        self.peer_ready = True

    # -- population ---------------------------------------------------------------------------
    def generate_shard(self, sequences, P, rank, world):
        import GA as ga_mod
        g = ga_mod.GA.__new__(ga_mod.GA)
        g.sequences, g.mode, g.population_size, g._dev = sequences, 'sp', P, None
        g._device = self.device
        # room for two worst-case mutations: re-laying out symmetric buffers later costs a rendezvous (~0.1 s)
        extra = max(32, 2 * (int(config.AGENT_WINDOW_ROW) - 1) * int(config.AGENT_WINDOW_COLUMN))
        width = max(len(s) for s in sequences) + max(1, round(max(len(s) for s in sequences) * config.GAP_RATE))
        return g._generate_host(rank, world, min_stride=width + extra)     # final stride at once: no second host copy

    def generate_and_adopt(self, sequences, P, rank, world, capacity):
        """The shard generated on the device (GA._generate_device): every rank walks the whole stream on its own GPU
        and keeps its own boards. False when the shape is outside the device generator (the host replay follows)."""
        import GA as ga_mod
        from population import DevicePopulation
        g = ga_mod.GA.__new__(ga_mod.GA)
        g.sequences, g.mode, g.population_size, g._dev = sequences, 'sp', P, None
        g._device = self.device
        extra = max(32, 2 * (int(config.AGENT_WINDOW_ROW) - 1) * int(config.AGENT_WINDOW_COLUMN))
        width = max(len(s) for s in sequences) + max(1, round(max(len(s) for s in sequences) * config.GAP_RATE))
        made = g._generate_device(rank, world, min_stride=width + extra)
        if made is None:
            return False
        boards, rowlen = made
        self.pop = DevicePopulation(boards, rowlen, capacity=max(capacity, 1),
                                    allocator=self._symm_alloc if self._symm else None)
        self.pop.width_hint = width
        self.R = self.pop.R
        if self._symm:
            self.pop.ensure_alt()
            self.pop.on_realloc = self._on_realloc
            self._rendezvous()
        return True

    def adopt(self, boards, rowlen, capacity):
        from population import DevicePopulation
        extra = max(16, (int(config.AGENT_WINDOW_ROW) - 1) * int(config.AGENT_WINDOW_COLUMN))
        # from the global width when this rank owns no individual (population smaller than the world): every
        # rank must arrive at the same stride, peers read each other's rows with it
        widest = int(rowlen.max()) if rowlen.size else 0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            w = torch.tensor([widest], dtype=torch.int64, device=self.device if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
            widest = int(w.item())
        need = max(-(-(widest + extra) // 16) * 16, boards.shape[2])
        if boards.shape[2] < need:
            wide = np.full((boards.shape[0], boards.shape[1], need), 5, dtype=np.uint8)
            wide[:, :, :boards.shape[2]] = boards
            boards = wide
        self.pop = DevicePopulation.from_numpy(np.ascontiguousarray(boards), np.ascontiguousarray(rowlen),
                                               device=self.device, capacity=max(capacity, 1),
                                               allocator=self._symm_alloc if self._symm else None)
        self.R = self.pop.R
        if self._symm:
            self.pop.ensure_alt()
            self.pop.on_realloc = self._on_realloc
            self._rendezvous()

    def _on_realloc(self):
        self.pop.ensure_alt()
        self._rendezvous()

    def crossover_peer(self, peer_slots, pos, plan):
        self.pop.crossover_peer(peer_slots, pos, plan)

    def board(self, local, width):
        return self.pop.cur.boards[local, :, :width].contiguous()

    def survivors(self):
        p = self.pop
        return p.cur.boards[:p.n], p.cur.rowlen[:p.n]

    def validate_window(self):
        if config.AGENT_WINDOW_ROW > self.R:
            raise ValueError("AGENT_WINDOW_ROW is larger than the sequence count.")

    # -- phases -------------------------------------------------------------------------------
    def _weights(self):
        return int(config.MATCH_REWARD), int(config.MISMATCH_PENALTY), int(config.GAP_PENALTY)

    def fitness(self):
        return self.pop.fitness(*self._weights())

    def _ranker(self, n):
        from population import Ranker
        if self._rank_tmp is None or self._rank_tmp.capacity < n:
            self._rank_tmp = Ranker(max(n, int(config.POPULATION_SIZE)), self.device)
        return self._rank_tmp

    def hof_argbest(self, scores, n, mode, hof_score):
        return self._ranker(n).argbest(scores, n, mode, hof_score)

    device_pass = True            # ShardedGA._fitness_pass_device: the whole pass without a host synchronisation

    # -- the pass block (ShardedGA._fitness_pass_fused; layout in include/gadpamsa.h, gad_pass_scatter) ----------
    FLAGS_BYTES = 256

    def pass_block(self, cap):
        """This rank's pass block in symmetric memory - int32 flags[64] | int32 tables [2][5][cap + 1] | hall-of-fame
        board uint8 [R][pitch] - mapped on every rank (collective when it has to be allocated or grown: capacity
        and stride are the same on every rank, so all ranks arrive here together). None when symmetric memory is
        not in use (GAD_SHARDED_SCORES=nccl, one rank, no peer access): the NCCL pass follows."""
        if self._symm is None or os.environ.get("GAD_SHARDED_SCORES", "peer") == "nccl":
            return None
        p = self.pop
        cap1 = int(cap) + 1
        if self._blk is None or self._blk_cap1 < cap1 or self._blk_pitch < p.stride or self._blk_R != self.R:
            world = dist.get_world_size()
            cap1 = max(cap1, self._blk_cap1)
            pitch = -(-max(2 * p.stride, self._blk_pitch) // 16) * 16
            table_bytes = -(-(2 * 5 * cap1 * 4) // 256) * 256
            hof_bytes = -(-(self.R * pitch) // 256) * 256
            # staging area: per pass parity and sender a header and four contiguous int32 arrays (scores + logical
            # index) of up to cap_s entries; GAD_SHARDED_SCATTER=direct keeps the stores at the logical index
            staged = os.environ.get("GAD_SHARDED_SCATTER", "staged") != "direct"
            cap_s = -(-cap1 // 4) * 4 if staged else 0
            stage_bytes = 2 * world * (16 + 16 * cap_s) if staged else 0
            t = self._symm.empty(self.FLAGS_BYTES + table_bytes + hof_bytes + stage_bytes, dtype=torch.uint8, device=self.device)
            t.zero_()
            board = t[self.FLAGS_BYTES + table_bytes:self.FLAGS_BYTES + table_bytes + self.R * pitch].view(self.R, pitch)
            board.fill_(5)
            if self._blk is not None and self._blk_R == self.R:                 # keep the record's board
                board[:, :self._blk_pitch].copy_(self.hof_board)
            hdl = self._symm.rendezvous(t, dist.group.WORLD)
            self._blk, self._blk_hdl, self._blk_cap1, self._blk_pitch, self._blk_R = t, hdl, cap1, pitch, self.R
            self._blk_table_off = self.FLAGS_BYTES
            self._blk_hof_off = self.FLAGS_BYTES + table_bytes
            self._blk_stage_off = self.FLAGS_BYTES + table_bytes + hof_bytes if staged else 0
            self._blk_cap_s = cap_s
            self._blk_peer = torch.tensor([int(b) for b in hdl.buffer_ptrs[:world]], dtype=torch.int64, device=self.device)
            self._blk_tables = t[self.FLAGS_BYTES:self.FLAGS_BYTES + 2 * 5 * cap1 * 4].view(torch.int32).view(2, 5, cap1)
            self.hof_board = board
            self._blk_seq = 0                                                   # the device counts passes the same way
            torch.cuda.synchronize(self.device)
            hdl.barrier(channel=0)                                              # every block is zeroed before anyone stores
        return self

    def run_pass(self, sp, em, al, gidx, n_loc, n, mode, hof_rec, hof_valid, rank, world):
        from population import MODE_ID
        nat, p = self.nat, self.pop
        st = nat.current_stream_ptr()
        gidx = gidx.contiguous()
        nat.call("gad_pass_scatter", nat.ptr(self._blk_peer), 0, self._blk_table_off, self._blk_cap1, self._blk_stage_off,
                 self._blk_cap_s, nat.ptr(sp), nat.ptr(em), nat.ptr(al), nat.ptr(gidx), n_loc, rank, world, st)
        rk = self._ranker(n)
        nat.call("gad_pass_decide", nat.ptr(self._blk_peer), 0, self._blk_table_off, self._blk_hof_off, self._blk_cap1,
                 self._blk_stage_off, self._blk_cap_s, self._blk_pitch, n, MODE_ID[mode], int(bool(hof_valid)),
                 nat.ptr(p.cur.boards), self.R, p.stride,
                 nat.ptr(hof_rec), rank, world, nat.ptr(rk.ws), rk.ws.numel(), st)
        self._blk_seq += 1
        return self._blk_tables[self._blk_seq & 1]

    def argbest_device(self, scores, n, mode, hof_dev):
        return self._ranker(n).argbest_device(scores, n, mode, hof_dev)

    def select_flags(self, scores, n, mode, top_n):
        return self._ranker(n).select_flags(scores, n, mode, top_n)

    def rank_order(self, scores, n, mode, k):
        return self._ranker(n).rank_order(scores, n, mode, k)

    def compact(self, local_keep):
        self.pop.compact_with(local_keep)

    def crossover_plan(self, n_parents, n_children):
        import GA as ga_mod
        plan = np.empty((n_children, 3), dtype=np.int32)
        state, meta = ga_mod._mt_state()
        self.nat.call("gad_mt_crossover_plan_host", state.ctypes.data, n_parents, self.R, n_children, None,
                      plan.ctypes.data)
        ga_mod._mt_restore(state, meta)
        return plan

    def crossover_plan_start(self, n_parents, n_children):
        """GA.PlanJob: the plan drawn on a worker thread while the GPU works on the phases before the crossover."""
        import GA as ga_mod
        return ga_mod.PlanJob.start(n_parents, self.R, n_children)

    def crossover_plan_finish(self, job, n_parents, n_children):
        plan = job.finish(n_parents, self.R, n_children)
        return plan if plan is not None else self.crossover_plan(n_parents, n_children)

    def crossover_from(self, all_boards, all_rowlen, pos, plan):
        self.pop.crossover_from(all_boards, all_rowlen, pos, plan)

    # -- mutation with the windows of all ranks pooled (ShardedGA._mutation_global) ---------------------------
    global_mutation = os.environ.get("GAD_SHARDED_MUTATION", "global") != "local"      # A/B: every rank on its own windows

    def mutation_windows(self, local_idx, model_path, mode, global_width):
        from DPAMSA import dqn as _dqn
        rw, cw = int(config.AGENT_WINDOW_ROW), int(config.AGENT_WINDOW_COLUMN)
        self.pop.ensure_stride(global_width + (rw - 1) * cw)
        qnet = _dqn.load_qnet(model_path, rw, cw)
        M = int(local_idx.numel())
        tile = self.pop.worst_tiles(local_idx, rw, cw, mode, *self._weights()) if M else \
            torch.empty((0, 2), dtype=torch.int32, device=self.device)
        if M and int(tile.min().item()) < 0:
            raise ValueError("min() arg is an empty sequence")          # no tile fits (utils.py:291)
        windows = torch.empty((M, rw * cw), dtype=torch.uint8, device=self.device)
        b = self.pop.cur
        self.nat.call("gad_mutate_windows", qnet.handle, self.nat.ptr(b.boards), self.R, self.pop.stride,
                      self.nat.ptr(local_idx), self.nat.ptr(tile), M, self.nat.ptr(windows), self.nat.current_stream_ptr())
        self._mut_qnet = qnet
        self._mut_max_value = cw * (rw * (rw - 1) / 2 * config.MATCH_REWARD)
        return windows, tile

    def windows_rep(self, pool):
        M = int(pool.shape[0])
        slots = 1024
        while slots < 2 * M:
            slots <<= 1
        rep = torch.empty(M, dtype=torch.int32, device=self.device)
        table = torch.empty(slots, dtype=torch.int32, device=self.device)
        tmp = torch.empty(max(M, 1), dtype=torch.int32, device=self.device)
        pool = pool.contiguous()
        self.nat.call("gad_windows_rep", self.nat.ptr(pool), M, int(pool.shape[1]), self.nat.ptr(rep), self.nat.ptr(table),
                      slots, self.nat.ptr(tmp), self.nat.current_stream_ptr())
        return rep

    def run_episodes(self, pool, mine, padded, reserve=0):
        import ctypes
        q = self._mut_qnet
        U = int(mine.numel())
        aligned = torch.zeros((padded, q.rows, q.rows * q.cols), dtype=torch.uint8, device=self.device)
        heads = torch.zeros((padded, q.rows), dtype=torch.uint8, device=self.device)
        steps = torch.full((padded,), -1, dtype=torch.int32, device=self.device)
        fw = ctypes.c_int64(0)
        pool = pool.contiguous()
        self.nat.call("gad_episodes_run", q.handle, self.nat.ptr(pool), self.nat.ptr(mine), U, int(reserve),
                      float(self._mut_max_value),
                      self.nat.ptr(aligned), self.nat.ptr(heads), self.nat.ptr(steps), ctypes.byref(fw),
                      self.nat.current_stream_ptr())
        st = (ctypes.c_int64 * 4)()
        self.nat.call("gad_mutate_stats", q.handle, st)
        p = self.pop
        p.forwards += fw.value
        p.unique_episodes += U
        p.states += int(st[3])
        return aligned, heads, steps

    def splice_results(self, local_idx, tile, windows, slot, aligned, heads, steps, width):
        q = self._mut_qnet
        p = self.pop
        M = int(local_idx.numel())
        p.episodes += M
        p.width_hint = width + q.cols
        if M == 0:
            return
        p._status.zero_()
        b = p.cur
        self.nat.call("gad_splice_results", q.handle, self.nat.ptr(b.boards), self.nat.ptr(b.rowlen), self.R, p.stride,
                      self.nat.ptr(local_idx), self.nat.ptr(tile), M, self.nat.ptr(windows), self.nat.ptr(slot.contiguous()),
                      self.nat.ptr(aligned), self.nat.ptr(heads), self.nat.ptr(steps), self.nat.ptr(p._status),
                      self.nat.current_stream_ptr())
        if int(p._status.item()) & 1:
            raise self.nat.GadCapacityError("a mutated row outgrew the board stride")

    def mutate(self, local_idx, model_path, mode, global_width):
        from DPAMSA import dqn as _dqn
        rw, cw = int(config.AGENT_WINDOW_ROW), int(config.AGENT_WINDOW_COLUMN)
        self.pop.ensure_stride(global_width + (rw - 1) * cw)
        if local_idx.numel() == 0:
            return
        qnet = _dqn.load_qnet(model_path, rw, cw)
        tile = self.pop.worst_tiles(local_idx, rw, cw, mode, *self._weights())
        self.pop.mutate(qnet, local_idx, tile, cw * (rw * (rw - 1) / 2 * config.MATCH_REWARD), width=global_width)
