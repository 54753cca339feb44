"""Device-resident population and the GA phases on it.

Replaces the reference's ``list[P] of list[R] of list[W] of int`` (GA.py:54-56) by
uint8 ``boards[P][R][stride]`` + int32 ``rowlen[P][R]`` in HBM (two buffers: selection compacts the
survivors from one into the other). PyTorch owns the allocations and the stream; all arithmetic is
done by libgadpamsa kernels (``_native``). Layout invariant (include/gadpamsa.h): with
w = max_r rowlen[p][r], the cells [rowlen[p][r], roundup4(w)) of every row hold the gap code 5.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

import _native as nat

GAP = 5
MAX_AL_MASK = (1 << 20) - 1          # include/gadpamsa.h GAD_MAX_AL_MASK
MODE_ID = nat.MODE_ID


def round_up(x: int, m: int) -> int:
    return -(-x // m) * m


def pack_boards(population, stride=None, extra_capacity=0):
    """list-of-boards -> numpy (boards uint8 [P,R,stride], rowlen int32 [P,R]) with the layout
    invariant established (everything beyond a row's own cells is the gap code)."""
    P = len(population)
    R = len(population[0]) if P else 1
    width = max((len(row) for b in population for row in b), default=0)
    need = round_up(max(width + extra_capacity, 1), 16)
    stride = max(stride or 0, need)
    boards = np.full((P, R, stride), GAP, dtype=np.uint8)
    rowlen = np.zeros((P, R), dtype=np.int32)
    for p, b in enumerate(population):
        if len(b) != R:
            raise ValueError("all boards of a population must have the same number of rows")
        for r, row in enumerate(b):
            n = len(row)
            rowlen[p, r] = n
            if n:
                boards[p, r, :n] = row
    return boards, rowlen


def unpack_boards(boards: np.ndarray, rowlen: np.ndarray):
    """numpy (boards, rowlen) -> list-of-boards of python ints (the reference's representation)."""
    return [[boards[p, r, :rowlen[p, r]].tolist() for r in range(boards.shape[1])]
            for p in range(boards.shape[0])]


class _Buffer:
    def __init__(self, cap, R, stride, dev, boards=None, rowlen=None, alloc=None):
        if boards is None:
            if alloc is None:
                boards = torch.full((cap, R, stride), GAP, dtype=torch.uint8, device=dev)
                rowlen = torch.zeros((cap, R), dtype=torch.int32, device=dev)
            else:                                  # e.g. symmetric memory, so that peers can read the boards
                boards = alloc((cap, R, stride), torch.uint8)
                boards.fill_(GAP)
                rowlen = alloc((cap, R), torch.int32)
                rowlen.zero_()
        self.boards = boards
        self.rowlen = rowlen
        self.peer_boards = None                    # device int64[world]: this buffer's address on every rank
        self.peer_rowlen = None
        self.sp = torch.zeros(cap, dtype=torch.int32, device=dev)
        self.em = torch.zeros(cap, dtype=torch.int32, device=dev)
        self.al = torch.zeros(cap, dtype=torch.int32, device=dev)


class DevicePopulation:
    """Up to ``capacity`` boards of R rows resident on one GPU, plus the GA state that goes with them
    (scores, hall of fame, ranking workspace)."""

    def __init__(self, boards: torch.Tensor, rowlen: torch.Tensor, capacity=None, allocator=None):
        assert boards.is_cuda and boards.dtype == torch.uint8 and boards.dim() == 3
        assert rowlen.dtype == torch.int32 and rowlen.shape == boards.shape[:2]
        assert boards.is_contiguous() and rowlen.is_contiguous()
        assert boards.shape[2] % 16 == 0
        self.device = boards.device
        self.n = boards.shape[0]
        self.capacity = max(capacity or 0, self.n, 1)
        self.R = boards.shape[1]
        self.stride = boards.shape[2]
        self._alloc = allocator
        self.on_realloc = None                   # called after the buffers were re-allocated (sharded: re-rendezvous)
        if self.capacity == self.n and allocator is None:      # adopt the caller's tensors as the first buffer
            self.cur = _Buffer(self.capacity, self.R, self.stride, self.device, boards, rowlen)
        else:
            self.cur = self._new_buffer(self.capacity)
            self.cur.boards[:self.n].copy_(boards)
            self.cur.rowlen[:self.n].copy_(rowlen)
        self.alt = None                          # second buffer, allocated by the first selection
        dev = self.device
        self.max_al = torch.zeros(1, dtype=torch.int64, device=dev)     # tagged: low 20 bits = widest cleaned board
        self.hof_board = torch.full((self.R, self.stride), GAP, dtype=torch.uint8, device=dev)
        self.hof = torch.zeros(8, dtype=torch.int32, device=dev)
        self.hof_valid = False
        self._ws = None
        self._keys = None
        self._order = None
        self._keep = None
        self._dst = None
        self._count = torch.zeros(1, dtype=torch.int32, device=dev)
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.width_hint = 0                      # expected board width (lanes per board in the fitness pass)
        self.evals = 0                           # boards scored so far (reference-equivalent accounting)
        self.forwards = 0                        # lock-step Q-net forwards so far
        self.episodes = 0                        # mutation episodes so far (reference-equivalent accounting)
        self.unique_episodes = 0                 # of those, the ones that ran (distinct windows, see gad_mutate_stats)
        self.states = 0                          # Q-net states evaluated so far

    # ------------------------------------------------------------------ construction / readback
    def _new_buffer(self, cap):
        return _Buffer(cap, self.R, self.stride, self.device, alloc=self._alloc)

    def ensure_alt(self):
        if self.alt is None or self.alt.boards.shape != self.cur.boards.shape:
            self.alt = self._new_buffer(self.capacity)

    @classmethod
    def from_lists(cls, population, device="cuda", stride=None, extra_capacity=0, capacity=None):
        nat.require_device()
        boards, rowlen = pack_boards(population, stride, extra_capacity)
        return cls.from_numpy(boards, rowlen, device, capacity)

    @classmethod
    def from_numpy(cls, boards: np.ndarray, rowlen: np.ndarray, device="cuda", capacity=None, allocator=None):
        nat.require_device()
        pop = cls(torch.from_numpy(boards).to(device), torch.from_numpy(rowlen).to(device), capacity, allocator)
        pop.width_hint = int(rowlen.max()) if rowlen.size else 0
        return pop

    def to_numpy(self):
        return self.cur.boards[:self.n].cpu().numpy(), self.cur.rowlen[:self.n].cpu().numpy()

    def to_lists(self):
        return unpack_boards(*self.to_numpy())

    # legacy names used by the first tests / bench
    @property
    def P(self) -> int:
        return self.n

    @property
    def boards(self):
        return self.cur.boards[:self.n]

    @property
    def rowlen(self):
        return self.cur.rowlen[:self.n]

    @property
    def sp(self):
        return self.cur.sp[:self.n]

    @property
    def em(self):
        return self.cur.em[:self.n]

    @property
    def al(self):
        return self.cur.al[:self.n]

    def widest(self) -> int:
        """Width of the widest board after the last fitness pass (one 8-byte read)."""
        return int(self.max_al.item()) & MAX_AL_MASK

    def scores_numpy(self):
        s = torch.stack((self.cur.sp[:self.n], self.cur.em[:self.n], self.cur.al[:self.n])).cpu().numpy()
        return s[0], s[1], s[2]

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return nat.current_stream_ptr()

    def _workspace(self):
        if self._ws is None:
            nbytes = ctypes.c_size_t(0)
            nat.call("gad_workspace_bytes", self.capacity, ctypes.byref(nbytes))
            dev = self.device
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            self._keys = torch.empty(self.capacity + 1, dtype=torch.float64, device=dev)
            self._order = torch.empty(self.capacity + 1, dtype=torch.int32, device=dev)
            self._keep = torch.empty(self.capacity, dtype=torch.uint8, device=dev)
            self._dst = torch.empty(self.capacity, dtype=torch.int32, device=dev)
        return self._ws

    def ensure_stride(self, need: int):
        """Re-lay the population out with a wider row stride (rows grow under mutation)."""
        need = round_up(need, 16)
        if need <= self.stride:
            return
        need = max(need, round_up(self.stride + self.stride // 2, 16))      # grow geometrically: re-laying out a
        old = self.stride                                                    # sharded population costs a rendezvous
        self.stride = need
        for name in ("cur", "alt"):
            buf = getattr(self, name)
            if buf is None:
                continue
            wide = self._new_buffer(buf.boards.shape[0])
            wide.boards[:, :, :old].copy_(buf.boards)
            for f in ("rowlen", "sp", "em", "al"):
                getattr(wide, f).copy_(getattr(buf, f))
            setattr(self, name, wide)
        hof = torch.full((self.R, need), GAP, dtype=torch.uint8, device=self.device)
        hof[:, :old].copy_(self.hof_board)
        self.hof_board = hof
        if self.on_realloc:
            self.on_realloc()

    def ensure_capacity(self, cap: int):
        if cap <= self.capacity:
            return
        for name in ("cur", "alt"):
            buf = getattr(self, name)
            if buf is None:
                continue
            big = self._new_buffer(cap)
            for f in ("boards", "rowlen", "sp", "em", "al"):
                getattr(big, f)[:self.capacity].copy_(getattr(buf, f))
            setattr(self, name, big)
        self.capacity = cap
        self._ws = None
        if self.on_realloc:
            self.on_realloc()

    # ------------------------------------------------------------------ a4-a7 + a8
    def fitness(self, match: int, mismatch: int, gap: int):
        """One fused pad+clean+SP+EM pass over the live boards (GA.py:156-178). Asynchronous on the
        current stream; results land in self.sp / self.em / self.al (int32 device tensors)."""
        b = self.cur
        nat.call("gad_fitness", nat.ptr(b.boards), nat.ptr(b.rowlen), self.n, self.R, self.stride,
                 int(self.width_hint), int(match), int(mismatch), int(gap), nat.ptr(b.sp), nat.ptr(b.em),
                 nat.ptr(b.al), nat.ptr(self.max_al), self._stream())
        self.evals += self.n
        return self.sp, self.em, self.al

    def fitness_pass(self, match: int, mismatch: int, gap: int, mode: str):
        """GA.calculate_fitness_score (GA.py:138-181): the scoring pass and GA.update_hall_of_fame - one launch for
        the shapes the TMA-staged kernel takes (the decision rides in the kernel's tail), else two."""
        ws = self._workspace()
        b = self.cur
        nat.call("gad_fitness_pass", nat.ptr(b.boards), nat.ptr(b.rowlen), self.n, self.R, self.stride,
                 int(self.width_hint), int(match), int(mismatch), int(gap), nat.ptr(b.sp), nat.ptr(b.em),
                 nat.ptr(b.al), nat.ptr(self.max_al), MODE_ID[mode], nat.ptr(self.hof_board), nat.ptr(self.hof),
                 int(self.hof_valid), nat.ptr(ws), ws.numel(), self._stream())
        self.evals += self.n
        self.hof_valid = True
        return self.sp, self.em, self.al

    def hof_update(self, mode: str):
        """GA.update_hall_of_fame (GA.py:110-136) on device."""
        ws = self._workspace()
        b = self.cur
        nat.call("gad_hof_update", nat.ptr(b.boards), nat.ptr(b.sp), nat.ptr(b.em), nat.ptr(b.al), self.n, self.R,
                 self.stride, MODE_ID[mode], nat.ptr(self.hof_board), nat.ptr(self.hof), int(self.hof_valid),
                 nat.ptr(ws), ws.numel(), self._stream())
        self.hof_valid = True

    def hof_numpy(self):
        """(board uint8 [R, width], (index, sp, em, al)) of the hall of fame."""
        rec = self.hof.cpu().numpy()
        width = int(rec[4])
        return self.hof_board[:, :width].cpu().numpy(), (int(rec[1]), int(rec[2]), int(rec[3]), width)

    # ------------------------------------------------------------------ a9
    def rank_order(self, mode: str, n: int):
        """Device tensor with utils.get_index_of_the_best_fitted_individuals(scores, n)."""
        ws = self._workspace()
        b = self.cur
        st = self._stream()
        nat.call("gad_rank_keys", nat.ptr(b.sp), nat.ptr(b.em), nat.ptr(b.al), self.n, MODE_ID[mode], None,
                 nat.ptr(self._keys), nat.ptr(ws), ws.numel(), st)
        nat.call("gad_rank_order", nat.ptr(self._keys), self.n, nat.ptr(self._order), nat.ptr(ws), ws.numel(), st)
        return self._order[:min(n, self.n)]

    # ------------------------------------------------------------------ a10 / a11
    def select_flags(self, mode: str, top_n: int):
        """Survivor flags of GA.selection (top_n = -1: the Pareto front only). Returns (keep, count)."""
        ws = self._workspace()
        b = self.cur
        nat.call("gad_select", nat.ptr(b.sp), nat.ptr(b.em), nat.ptr(b.al), self.n, MODE_ID[mode], int(top_n),
                 nat.ptr(self._keep), nat.ptr(self._dst), nat.ptr(self._count), nat.ptr(ws), ws.numel(),
                 self._stream())
        return self._keep[:self.n], self._count

    def select(self, mode: str, top_n: int) -> int:
        """GA.py:232-259: keep the survivors, in original order. Returns their number (one 4-byte
        device->host read: the crossover plan needs it on the host)."""
        self.select_flags(mode, top_n)
        self.ensure_alt()
        s, d = self.cur, self.alt
        nat.call("gad_compact", nat.ptr(s.boards), nat.ptr(s.rowlen), nat.ptr(s.sp), nat.ptr(s.em), nat.ptr(s.al),
                 nat.ptr(self._keep), nat.ptr(self._dst), self.n, self.R, self.stride, nat.ptr(d.boards),
                 nat.ptr(d.rowlen), nat.ptr(d.sp), nat.ptr(d.em), nat.ptr(d.al), self._stream())
        self.cur, self.alt = d, s
        both = torch.stack((self._count[0].to(torch.int64), self.max_al[0])).cpu()   # one small read: survivor count + widest board
        self.n = int(both[0])
        self.width_hint = int(both[1]) & MAX_AL_MASK
        return self.n

    # ------------------------------------------------------------------ a12 / a13
    def crossover(self, plan: np.ndarray, kind: int = 0):
        """Append len(plan) children built from (parent1, parent2, cut) triples (GA.py:296)."""
        n_children = int(plan.shape[0])
        if n_children == 0:
            return
        assert plan.dtype == np.int32 and plan.shape[1] == 3
        self.ensure_capacity(self.n + n_children)
        d_plan = torch.from_numpy(plan).to(self.device, non_blocking=False)
        b = self.cur
        nat.call("gad_crossover", nat.ptr(b.boards), nat.ptr(b.rowlen), self.n, nat.ptr(d_plan), n_children, self.R,
                 self.stride, int(kind), self._stream())
        self.n += n_children

    # ------------------------------------------------------------------ a14-a18
    def worst_tiles(self, idx, rw, cw, mode, match, mismatch, gap):
        """Device int32 [M, 2] with (from_row, from_col) of every listed individual's worst tile."""
        M = self.n if idx is None else int(idx.numel())
        tile = torch.empty((M, 2), dtype=torch.int32, device=self.device)
        b = self.cur
        nat.call("gad_worst_tiles", nat.ptr(b.boards), nat.ptr(b.rowlen), self.R, self.stride, nat.ptr(idx), M,
                 int(rw), int(cw), MODE_ID[mode], int(match), int(mismatch), int(gap), nat.ptr(tile), self._stream())
        return tile

    def compact_with(self, keep: torch.Tensor):
        """Keep the local boards flagged in `keep` (bool [n]), in order (sharded selection)."""
        self._workspace()
        n = self.n
        k8 = keep.to(torch.uint8).contiguous()
        dst = (torch.cumsum(keep.to(torch.int32), 0, dtype=torch.int32) - keep.to(torch.int32)).contiguous()
        self.ensure_alt()
        s, d = self.cur, self.alt
        nat.call("gad_compact", nat.ptr(s.boards), nat.ptr(s.rowlen), nat.ptr(s.sp), nat.ptr(s.em), nat.ptr(s.al),
                 nat.ptr(k8), nat.ptr(dst), n, self.R, self.stride, nat.ptr(d.boards), nat.ptr(d.rowlen),
                 nat.ptr(d.sp), nat.ptr(d.em), nat.ptr(d.al), self._stream())
        self.cur, self.alt = d, s
        self.n = int(keep.sum().item())

    def crossover_from(self, src_boards, src_rowlen, pos, plan: np.ndarray, kind: int = 0):
        """Append children whose parents are rows of an all-gathered survivor buffer; `plan` holds
        logical parent indices, `pos` maps a logical index to its row in `src_boards`."""
        n_children = int(plan.shape[0])
        if n_children == 0:
            return
        assert src_boards.shape[1] == self.R and src_boards.shape[2] == self.stride, "ranks disagree on the board stride"
        self.ensure_capacity(self.n + n_children)
        d_plan = torch.from_numpy(np.ascontiguousarray(plan)).to(self.device)
        d_plan[:, 0] = pos[d_plan[:, 0].long()].to(torch.int32)
        d_plan[:, 1] = pos[d_plan[:, 1].long()].to(torch.int32)
        d_plan = d_plan.contiguous()
        src_boards, src_rowlen = src_boards.contiguous(), src_rowlen.contiguous()
        b = self.cur
        nat.call("gad_crossover_from", nat.ptr(src_boards), nat.ptr(src_rowlen), nat.ptr(b.boards), nat.ptr(b.rowlen),
                 self.n, nat.ptr(d_plan), n_children, self.R, self.stride, int(kind), self._stream())
        torch.cuda.current_stream().synchronize()          # src buffers are temporaries of the caller
        self.n += n_children

    def crossover_peer(self, peer_slots: int, pos, plan: np.ndarray, kind: int = 0):
        """Append children whose parents are read from the owning ranks' buffers over NVLink (fused exchange +
        crossover, gad_crossover_peer). `plan` holds logical parent indices; pos[logical] = rank * peer_slots + slot."""
        n_children = int(plan.shape[0])
        if n_children == 0:
            return
        assert self.n + n_children <= self.capacity, "symmetric buffers cannot grow inside a collective phase"
        d_plan = torch.from_numpy(np.ascontiguousarray(plan)).to(self.device)
        d_plan[:, 0] = pos[d_plan[:, 0].long()].to(torch.int32)
        d_plan[:, 1] = pos[d_plan[:, 1].long()].to(torch.int32)
        d_plan = d_plan.contiguous()
        b = self.cur
        nat.call("gad_crossover_peer", nat.ptr(b.peer_boards), nat.ptr(b.peer_rowlen), int(peer_slots), nat.ptr(b.boards),
                 nat.ptr(b.rowlen), self.n, nat.ptr(d_plan), n_children, self.R, self.stride, int(kind), self._stream())
        self.n += n_children

    def mutate(self, qnet, idx, tile, max_value: float, width=None):
        """GA.py:335-373 for the listed individuals with their tiles; boards grow in place."""
        M = int(tile.shape[0])
        if M == 0:
            return
        if width is None:
            width = int(self.cur.rowlen[:self.n].max().item())
        self.ensure_stride(width + (qnet.rows - 1) * qnet.cols)
        self.width_hint = width + qnet.cols              # mutated boards are wider until the next pass cleans them
        self._status.zero_()
        fw = ctypes.c_int64(0)
        b = self.cur
        nat.call("gad_mutate", qnet.handle, nat.ptr(b.boards), nat.ptr(b.rowlen), self.R, self.stride, nat.ptr(idx),
                 nat.ptr(tile), M, float(max_value), nat.ptr(self._status), ctypes.byref(fw), self._stream())
        self.forwards += fw.value
        st = (ctypes.c_int64 * 4)()
        nat.call("gad_mutate_stats", qnet.handle, st)
        self.episodes += int(st[0])                 # what GA.py:331-373 would run
        self.unique_episodes += int(st[1])          # what ran here (distinct windows)
        self.states += int(st[3])                   # Net.forward evaluations
        if int(self._status.item()) & 1:
            raise nat.GadCapacityError("a mutated row outgrew the board stride")


class Ranker:
    """Ranking / selection / hall-of-fame decisions on explicit score tensors (used on the all-gathered
    global scores of a sharded population)."""

    def __init__(self, capacity: int, device):
        self.capacity = capacity
        self.device = device
        nbytes = ctypes.c_size_t(0)
        nat.call("gad_workspace_bytes", capacity, ctypes.byref(nbytes))
        self.ws = torch.empty(nbytes.value, dtype=torch.uint8, device=device)
        self.keys = torch.empty(capacity + 1, dtype=torch.float64, device=device)
        self.order = torch.empty(capacity + 1, dtype=torch.int32, device=device)
        self.keep = torch.empty(capacity, dtype=torch.uint8, device=device)
        self.dst = torch.empty(capacity, dtype=torch.int32, device=device)
        self.count = torch.zeros(1, dtype=torch.int32, device=device)
        self.winner = torch.zeros(1, dtype=torch.int32, device=device)
        self.hof = torch.zeros(8, dtype=torch.int32, device=device)

    def rank_order(self, scores, n, mode, k):
        sp, em, al = scores
        st = nat.current_stream_ptr()
        nat.call("gad_rank_keys", nat.ptr(sp), nat.ptr(em), nat.ptr(al), n, MODE_ID[mode], None, nat.ptr(self.keys),
                 nat.ptr(self.ws), self.ws.numel(), st)
        nat.call("gad_rank_order", nat.ptr(self.keys), n, nat.ptr(self.order), nat.ptr(self.ws), self.ws.numel(), st)
        return self.order[:min(k, n)]

    def select_flags(self, scores, n, mode, top_n):
        sp, em, al = scores
        nat.call("gad_select", nat.ptr(sp), nat.ptr(em), nat.ptr(al), n, MODE_ID[mode], int(top_n), nat.ptr(self.keep),
                 nat.ptr(self.dst), nat.ptr(self.count), nat.ptr(self.ws), self.ws.numel(), nat.current_stream_ptr())
        return self.keep[:n], self.dst[:n], int(self.count.item())

    def argbest_device(self, scores, n, mode, hof_dev):
        """Like argbest with the hall-of-fame record given as a device int32[8] (or None) and the winner left on the
        device (int32[1] tensor): nothing is read back."""
        sp, em, al = (t.contiguous() for t in scores)
        nat.call("gad_hof_argbest", nat.ptr(sp), nat.ptr(em), nat.ptr(al), n, MODE_ID[mode],
                 None if hof_dev is None else nat.ptr(hof_dev), nat.ptr(self.winner), nat.ptr(self.ws), self.ws.numel(),
                 nat.current_stream_ptr())
        return self.winner.clone()

    def argbest(self, scores, n, mode, hof_score):
        sp, em, al = scores
        hof_ptr = None
        if hof_score is not None:
            self.hof.copy_(torch.tensor([1, hof_score[0], hof_score[1], hof_score[2], hof_score[3], 0, 0, 0],
                                        dtype=torch.int32))
            hof_ptr = nat.ptr(self.hof)
        nat.call("gad_hof_argbest", nat.ptr(sp), nat.ptr(em), nat.ptr(al), n, MODE_ID[mode], hof_ptr,
                 nat.ptr(self.winner), nat.ptr(self.ws), self.ws.numel(), nat.current_stream_ptr())
        return int(self.winner.item())
