"""Experience replay buffer of the DQN trainer (reference DPAMSA/replay_memory.py:28-84): same attributes
(``storage``, ``max_size``, ``size``, ``ptr``), same ``push`` / ``sample`` behaviour, including what the reference
does once the buffer is full - it overwrites slot ``ptr - 1`` and then advances ``ptr`` (replay_memory.py:55-57),
so the first overwrite lands on the newest entry, not the oldest. Sampling draws from CPython's ``random`` module
exactly like the reference, so a seeded run picks the same batches.
"""
from __future__ import annotations

import random

import config


class ReplayMemory:
    def __init__(self):
        self.storage = []
        self.max_size = config.REPLAY_MEMORY_SIZE
        self.size = 0
        self.ptr = 0
        self.previous_hash = None

    def push(self, data: tuple):
        """data = (state, next_state, action, reward, done)"""
        if len(self.storage) < self.max_size:
            self.storage.append(data)
            self.ptr += 1
            self.size += 1
            return
        self.storage[self.ptr - 1] = data
        self.ptr = (self.ptr + 1) % self.max_size

    def sample(self, batch_size):
        """Five parallel lists (states, next_states, actions, rewards, done flags) of ``batch_size`` transitions."""
        picked = random.sample(self.storage, batch_size)
        cols = tuple(zip(*picked)) if picked else ((), (), (), (), ())
        return tuple(list(c) for c in cols)
