"""DQN agent facade (reference DPAMSA/dqn.py:82-213): same constructor, ``load`` / ``predict`` /
``save``; the network itself is libgadpamsa's folded-table Q-net (csrc/qnet.cu) on the GPU, in eval
mode. Training (``select`` / ``update`` / replay memory, dqn.py:123-185) is outside the generation
loop and not provided.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

import config
import _native as nat

# state_dict keys in the order gad_qnet_create expects (reference dqn.py:54-63, models.py:158-166,235-239)
STATE_KEYS = (
    "encoder.src_word_emb.weight", "encoder.position_enc.pos_table",
    "encoder.layer_norm.weight", "encoder.layer_norm.bias",
    "encoder.self_attention.w_qs.weight", "encoder.self_attention.w_ks.weight",
    "encoder.self_attention.w_vs.weight", "encoder.self_attention.fc.weight",
    "encoder.self_attention.layer_norm.weight", "encoder.self_attention.layer_norm.bias",
    "l1.weight", "l1.bias", "l2.weight", "l2.bias", "l3.weight", "l3.bias",
)


class QNet:
    """Owner of one ``gad_qnet`` handle built from a reference state_dict."""

    def __init__(self, state_dict, rows: int, cols: int):
        nat.require_device()
        arrays = []
        for k in STATE_KEYS:
            if k not in state_dict:
                raise KeyError("weights file lacks %r (reference state_dict layout, dqn.py:187-199)" % k)
            v = state_dict[k]
            a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            arrays.append(np.ascontiguousarray(a, dtype=np.float32))
        L = rows * (cols + 1)
        A = 2 ** rows - 1
        d_k = arrays[4].shape[0]
        d_v = arrays[6].shape[0]
        h1, h2 = arrays[10].shape[0], arrays[12].shape[0]
        expect = {0: (6, 64), 1: (1, L, 64), 4: (d_k, 64), 5: (d_k, 64), 6: (d_v, 64), 7: (64, d_v),
                  10: (h1, L * 64), 12: (h2, h1), 14: (A, h2), 15: (A,)}
        for i, shp in expect.items():
            if tuple(arrays[i].shape) != shp:      # what load_state_dict(strict) reports in the reference
                raise RuntimeError("size mismatch for %s: checkpoint %s, agent window %dx%d needs %s"
                                   % (STATE_KEYS[i], tuple(arrays[i].shape), rows, cols, shp))
        self.rows, self.cols, self.L, self.A = rows, cols, L, A
        self.state = dict(zip(STATE_KEYS, arrays))
        ptrs = (ctypes.c_void_p * 16)(*[a.ctypes.data for a in arrays])
        handle = ctypes.c_void_p()
        nat.call("gad_qnet_create", ctypes.byref(handle), rows, cols, d_k, d_v, h1, h2, ptrs)
        self.handle = handle

    def forward(self, states, max_value: float):
        """Q-values (float32 [B, A]) and greedy actions (int32 [B]) of uint8 token states [B, L]."""
        s = np.ascontiguousarray(states, dtype=np.uint8).reshape(-1, self.L)
        B = s.shape[0]
        q = np.empty((B, self.A), np.float32)
        a = np.empty(B, np.int32)
        nat.call("gad_qnet_forward_host", self.handle, s.ctypes.data, B, float(max_value), q.ctypes.data, a.ctypes.data)
        return q, a

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                nat.lib().gad_qnet_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


_cache = {}


def load_qnet(filename, rows, cols, path=None):
    """Weights are read once per (file, mtime, window) instead of once per individual (GA.py:354-355)."""
    path = config.DPAMSA_WEIGHTS_PATH if path is None else path
    full = os.path.join(path, "{}.pth".format(filename))
    key = (os.path.abspath(full), os.path.getmtime(full) if os.path.exists(full) else None, rows, cols)
    if key not in _cache:
        sd = torch.load(full, map_location="cpu")           # FileNotFoundError like the reference
        _cache.clear()
        _cache[key] = QNet(sd, rows, cols)
    return _cache[key]


class DQN:
    """reference DPAMSA/dqn.py:105-121."""

    def __init__(self, action_number, seq_num, max_seq_len, max_value):
        self.seq_num = seq_num
        self.max_seq_len = max_seq_len
        self.action_number = action_number
        self.max_value = max_value
        self.current_epsilon = config.EPSILON
        self.update_step_counter = 0
        self.epsilon_step_counter = 0
        self.qnet = None

    def load(self, filename, path=None):
        """dqn.py:201-213"""
        self.qnet = load_qnet(filename, self.seq_num, self.max_seq_len, path)
        if self.qnet.A != self.action_number:
            raise RuntimeError("size mismatch for l3.weight: agent has %d actions, expected %d"
                               % (self.qnet.A, self.action_number))

    def predict(self, state):
        """dqn.py:151-154: greedy action as a numpy array of shape (1,)."""
        if self.qnet is None:
            raise RuntimeError("DQN.predict before DQN.load: this build has no randomly initialised network")
        _, a = self.qnet.forward(np.asarray(state, dtype=np.uint8)[None, :], self.max_value)
        return a.astype(np.int64)

    def q_values(self, state):
        q, _ = self.qnet.forward(np.asarray(state, dtype=np.uint8)[None, :], self.max_value)
        return q[0]

    def save(self, filename, path=None):
        """dqn.py:187-199"""
        path = config.DPAMSA_WEIGHTS_PATH if path is None else path
        torch.save({k: torch.from_numpy(v) for k, v in self.qnet.state.items()},
                   os.path.join(path, "{}.pth".format(filename)))

    def select(self, state):
        raise NotImplementedError("epsilon-greedy training (dqn.py:123-149) is outside the generation loop")

    def update(self):
        raise NotImplementedError("DQN training (dqn.py:156-185) is outside the generation loop")

    def update_epsilon(self):
        raise NotImplementedError("DQN training (dqn.py:123-131) is outside the generation loop")
