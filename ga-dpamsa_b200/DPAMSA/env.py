"""Alignment environment with the reference's interface (DPAMSA/env.py:61-90,139-174,225-351) for
single boards on the host side of the boundary: callers use it to hold an alignment and ask for its
metrics (mainGA.py:124-134). Scores are computed on the GPU (utils.get_sum_of_pairs /
get_column_score); the batched episodes of GA.mutation never go through this class - they run in
csrc/qnet.cu. The Tk GUI of the reference (Windows only) is not provided; ``render`` is a no-op.
"""
from __future__ import annotations

import copy

import numpy as np

import config
import utils

nucleotides_map = {'A': 1, 'T': 2, 'C': 3, 'G': 4, 'a': 1, 't': 2, 'c': 3, 'g': 4, '-': 5}
nucleotides = ['A', 'T', 'C', 'G', '-']


class Environment:
    def __init__(self, data, nucleotide_size=50, text_size=25, show_nucleotide_name=True, convert_data=True):
        if convert_data:
            self.data = [[nucleotides_map[char] for char in seq] for seq in data]     # KeyError like env.py:67
        else:
            self.data = data
        self.row = len(data)
        self.max_len = max([len(data[i]) for i in range(len(data))])
        self.show_nucleotide_name = show_nucleotide_name
        self.nucleotide_size = nucleotide_size
        self.max_window_width = 1800
        self.text_size = text_size
        self.action_number = 2 ** self.row - 1                                       # env.py:78
        self.max_reward = self.row * (self.row - 1) / 2 * config.MATCH_REWARD        # env.py:79
        self.aligned = [[] for _ in range(self.row)]
        self.not_aligned = copy.deepcopy(self.data)

    # ------------------------------------------------------------------ state machine (env.py:139-259)
    def _state(self):
        state = []
        for rest in self.not_aligned:
            state.extend(rest if len(rest) > 0 else [])
            state.append(5)
        state.extend([0] * (self.row * (self.max_len + 1) - len(state)))
        return state

    def reset(self):
        self.aligned = [[] for _ in range(self.row)]
        self.not_aligned = copy.deepcopy(self.data)
        return self._state()

    def step(self, action):
        action = int(np.asarray(action).reshape(-1)[0])
        for bit in range(self.row):
            if 0 == (action >> bit) & 0x1 and 0 == len(self.not_aligned[bit]):
                return -self.max_reward, self._state(), 0
        total_len = 0
        for bit in range(self.row):
            if 0 == (action >> bit) & 0x1:
                self.aligned[bit].append(self.not_aligned[bit].pop(0))
            else:
                self.aligned[bit].append(5)
            total_len += len(self.not_aligned[bit])
        last = len(self.aligned[0]) - 1                                              # env.py:163-172
        reward = utils.get_sum_of_pairs(self.aligned, 0, self.row, last, last + 1)
        return reward, self._state(), 1 if total_len > 0 else 0

    def padding(self):
        """env.py:344-351"""
        longest = max((len(r) for r in self.not_aligned), default=0)
        for i in range(len(self.not_aligned)):
            rest = self.not_aligned[i]
            self.aligned[i].extend(rest)
            self.aligned[i].extend([5] * (longest - len(rest)))
            rest.clear()

    # ------------------------------------------------------------------ metrics (env.py:261-302)
    def calc_score(self):
        return utils.get_sum_of_pairs(self.aligned, 0, self.row, 0, len(self.aligned[0]))

    def calc_exact_matched(self):
        width = len(self.aligned[0])
        if width == 0:
            return 0
        return utils.count_exact_columns(self.aligned, 0, self.row, 0, width)

    def set_alignment(self, seqs):
        """env.py:304-315"""
        if isinstance(seqs[0][0], int):
            self.aligned = seqs
        else:
            self.aligned = [[nucleotides_map[char] for char in seq] for seq in seqs]
        self.not_aligned = [[] for _ in range(len(self.data))]

    def get_alignment(self):
        """env.py:322-336"""
        out = []
        for seq in self.aligned:
            out.append(''.join(nucleotides[n - 1] for n in seq) if isinstance(seq[0], int) else ''.join(seq))
        return '\n'.join(out)

    def render(self):
        pass
