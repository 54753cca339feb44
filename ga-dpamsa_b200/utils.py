"""GA utilities with the reference's names and positional signatures (reference utils.py:27-469),
computed by libgadpamsa on the GPU. Boards passed to these free functions are the reference's
python lists; they are copied to the device, scored there and the result is returned as the
same python type the reference returns (int SP, float CS, list of indices, ...).

There is no CPU fallback: without the CUDA library or a GPU these functions raise.
The benchmarking half of the reference's utils.py (utils.py:472-975) lives in ``bench_utils``
and is re-exported at the bottom so ``run_benchmarks.py`` finds the same names.
"""
from __future__ import annotations

import ctypes

import numpy as np

import config
import _native as nat

_LETTERS = "ATCG-"          # reference utils.py:458


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _pack_window(chromosome, from_row, to_row, from_column, to_column):
    """The half-open window of a list-of-lists board as a dense uint8 [rows, stride] array."""
    rows, cols = to_row - from_row, to_column - from_column
    stride = max(4, -(-cols // 4) * 4)
    arr = np.full((rows, stride), 5, dtype=np.uint8)
    for i in range(rows):
        row = chromosome[from_row + i]
        if to_column > len(row):
            raise IndexError("list index out of range")          # what the reference raises
        arr[i, :cols] = row[from_column:to_column]
    return arr, stride


def _window_score(chromosome, from_row, to_row, from_column, to_column):
    """(SP, exact-match columns) of a non-empty window, scored on the device."""
    nat.require_device()
    arr, stride = _pack_window(chromosome, from_row, to_row, from_column, to_column)
    sp = ctypes.c_int64(0)
    em = ctypes.c_int32(0)
    nat.call("gad_window_score_host", arr.ctypes.data, arr.shape[0], stride, 0, to_row - from_row, 0,
             to_column - from_column, int(config.MATCH_REWARD), int(config.MISMATCH_PENALTY),
             int(config.GAP_PENALTY), ctypes.byref(sp), ctypes.byref(em))
    return sp.value, em.value


# ----------------------------------------------------------------------------------------------
# a6 / a7: scoring (reference utils.py:27-128)
# ----------------------------------------------------------------------------------------------
def get_sum_of_pairs(chromosome, from_row, to_row, from_column, to_column):
    """Sum-of-pairs of a half-open window (reference utils.py:62-76). Returns a python int."""
    if to_row - from_row < 2 or to_column <= from_column:
        return 0
    return int(_window_score(chromosome, from_row, to_row, from_column, to_column)[0])


def get_column_score(chromosome, from_row, to_row, from_column, to_column):
    """Fraction of columns whose cells all equal the first row's (reference utils.py:119-128);
    python true division of device-counted integers, 0 when the window has no columns."""
    num_columns = to_column - from_column
    if num_columns <= 0:
        return 0
    if to_row <= from_row:
        return num_columns / num_columns          # all() over no rows is True for every column
    em = _window_score(chromosome, from_row, to_row, from_column, to_column)[1]
    return em / num_columns


# ----------------------------------------------------------------------------------------------
# a5: all-gap column removal (reference utils.py:390-428), in place on the caller's lists
# ----------------------------------------------------------------------------------------------
def check_if_there_are_all_gaps(row, from_index):
    """reference utils.py:383-387 (tiny host helper kept for API completeness)."""
    return False if any(v != 5 for v in row[from_index:]) else from_index


def clean_unnecessary_gaps(aligned_sequence):
    """Remove every all-gap column of a rectangular board, in place (reference utils.py:404-428)."""
    if not aligned_sequence or not aligned_sequence[0]:
        return
    nat.require_device()
    from population import pack_boards
    width = len(aligned_sequence[0])
    if any(len(r) != width for r in aligned_sequence):
        raise IndexError("list index out of range")              # reference behaviour on ragged input
    boards, rowlen = pack_boards([aligned_sequence])
    P, R, stride = boards.shape
    out = np.zeros(3, dtype=np.int32)
    nat.call("gad_fitness_host", boards.ctypes.data, rowlen.ctypes.data, 1, R, stride,
             int(config.MATCH_REWARD), int(config.MISMATCH_PENALTY), int(config.GAP_PENALTY),
             out[0:1].ctypes.data, out[1:2].ctypes.data, out[2:3].ctypes.data, 1)
    new_w = int(rowlen[0, 0])
    for r, row in enumerate(aligned_sequence):
        row[:] = boards[0, r, :new_w].tolist()


def get_nucleotides_seqs(chromosome):
    """int codes -> strings (reference utils.py:464-467); pure formatting, stays on the host."""
    return ["".join(_LETTERS[v - 1] for v in sequence) for sequence in chromosome]
