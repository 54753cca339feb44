"""Device-resident population: uint8 boards[P][R][stride] + int32 rowlen[P][R] in HBM.

Replaces the reference's ``list[P] of list[R] of list[W] of int`` (GA.py:54-56). PyTorch owns
the allocations and the stream; all arithmetic is done by libgadpamsa kernels (``_native``).
Layout invariant (see include/gadpamsa.h): with w = max_r rowlen[p][r], the cells
[rowlen[p][r], roundup4(w)) of every row hold the gap code 5.
"""
from __future__ import annotations

import numpy as np
import torch

import _native as nat

GAP = 5


def round_up(x: int, m: int) -> int:
    return -(-x // m) * m


def pack_boards(population, stride=None, extra_capacity=0):
    """list-of-boards -> pinned-able numpy (boards uint8 [P,R,stride], rowlen int32 [P,R]) with the
    layout invariant established (everything beyond a row's own cells is the gap code)."""
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


class DevicePopulation:
    """P boards of R rows resident on one GPU."""

    def __init__(self, boards: torch.Tensor, rowlen: torch.Tensor):
        assert boards.is_cuda and boards.dtype == torch.uint8 and boards.dim() == 3
        assert rowlen.dtype == torch.int32 and rowlen.shape == boards.shape[:2]
        assert boards.is_contiguous() and rowlen.is_contiguous()
        self.boards = boards
        self.rowlen = rowlen
        P = boards.shape[0]
        dev = boards.device
        self.sp = torch.zeros(P, dtype=torch.int32, device=dev)
        self.em = torch.zeros(P, dtype=torch.int32, device=dev)
        self.al = torch.zeros(P, dtype=torch.int32, device=dev)
        self.max_al = torch.zeros(1, dtype=torch.int32, device=dev)

    # ------------------------------------------------------------------ construction / readback
    @classmethod
    def from_lists(cls, population, device="cuda", stride=None, extra_capacity=0):
        nat.require_device()
        boards, rowlen = pack_boards(population, stride, extra_capacity)
        return cls(torch.from_numpy(boards).to(device), torch.from_numpy(rowlen).to(device))

    @classmethod
    def from_numpy(cls, boards: np.ndarray, rowlen: np.ndarray, device="cuda"):
        nat.require_device()
        return cls(torch.from_numpy(boards).to(device), torch.from_numpy(rowlen).to(device))

    def to_lists(self):
        return unpack_boards(self.boards.cpu().numpy(), self.rowlen.cpu().numpy())

    @property
    def P(self) -> int:
        return self.boards.shape[0]

    @property
    def R(self) -> int:
        return self.boards.shape[1]

    @property
    def stride(self) -> int:
        return self.boards.shape[2]

    # ------------------------------------------------------------------ a4-a7
    def fitness(self, match: int, mismatch: int, gap: int):
        """One fused pad+clean+SP+EM pass (GA.py:156-178). Asynchronous on the current stream;
        results land in self.sp / self.em / self.al (int32 device tensors)."""
        nat.require_device()
        self.max_al.zero_()
        nat.call("gad_fitness", nat.ptr(self.boards), nat.ptr(self.rowlen), self.P, self.R, self.stride,
                 int(match), int(mismatch), int(gap), nat.ptr(self.sp), nat.ptr(self.em), nat.ptr(self.al),
                 nat.ptr(self.max_al), nat.current_stream_ptr())
        return self.sp, self.em, self.al
