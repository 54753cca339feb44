"""DQN agent facade (reference DPAMSA/dqn.py:82-213): same constructor and methods.

``load`` / ``predict`` run the network as libgadpamsa's folded-table Q-net (csrc/qnet.cu, csrc/attn_tc.cu) on the GPU,
in eval mode - that is the generation loop's path. ``select`` / ``update`` / ``update_epsilon`` / ``replay_memory``
(dqn.py:123-185) train a torch-autograd copy of the same network (DPAMSA/trainer.py)
with the reference's epsilon-greedy policy, Adam(lr=ALPHA), MSE loss and target-network schedule; ``save`` writes the
reference's state_dict layout either way, so trained weights go straight back into ``load``.
"""
from __future__ import annotations

import ctypes
import os

import random

import numpy as np
import torch

import config
import _native as nat
from DPAMSA.replay_memory import ReplayMemory
from DPAMSA.trainer import TrainNet

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
    device = torch.cuda.current_device() if torch.cuda.is_available() else -1      # a handle lives on one GPU
    key = (os.path.abspath(full), os.path.getmtime(full) if os.path.exists(full) else None, rows, cols, device)
    if key not in _cache:
        sd = torch.load(full, map_location="cpu")           # FileNotFoundError like the reference
        for old in [k for k in _cache if k[4] == device]:   # one resident agent per device
            del _cache[old]
        _cache[key] = QNet(sd, rows, cols)
    return _cache[key]


class DQN:
    """reference DPAMSA/dqn.py:82-213."""

    def __init__(self, action_number, seq_num, max_seq_len, max_value):
        self.seq_num = seq_num
        self.max_seq_len = max_seq_len
        self.action_number = action_number
        self.max_value = max_value
        self.current_epsilon = config.EPSILON
        self.update_step_counter = 0
        self.epsilon_step_counter = 0
        self.qnet = None                  # device Q-net of the loaded / frozen weights (the generation loop's path)
        self.replay_memory = ReplayMemory()
        # the trainable networks are only built when training is touched (dqn.py:110-121 builds them eagerly)
        self.eval_net = None
        self.target_net = None
        self.optimizer = None
        self.train_device = config.DEVICE if torch.cuda.is_available() else "cpu"

    # ------------------------------------------------------------------ inference: libgadpamsa
    def load(self, filename, path=None):
        """dqn.py:201-213"""
        self.qnet = load_qnet(filename, self.seq_num, self.max_seq_len, path)
        if self.qnet.A != self.action_number:
            raise RuntimeError("size mismatch for l3.weight: agent has %d actions, expected %d"
                               % (self.qnet.A, self.action_number))
        if self.eval_net is not None:
            self.eval_net.load_state_dict({k: torch.from_numpy(v) for k, v in self.qnet.state.items()})

    def _frozen(self):
        if self.qnet is None and self.eval_net is not None:           # weights trained in this process
            self.qnet = QNet(self.eval_net.state_dict(), self.seq_num, self.max_seq_len)
        if self.qnet is None:
            raise RuntimeError("DQN.predict before DQN.load or any training step: there are no weights yet")
        return self.qnet

    def predict(self, state):
        """dqn.py:151-154: greedy action as a numpy array of shape (1,)."""
        _, a = self._frozen().forward(np.asarray(state, dtype=np.uint8)[None, :], self.max_value)
        return a.astype(np.int64)

    def q_values(self, state):
        q, _ = self._frozen().forward(np.asarray(state, dtype=np.uint8)[None, :], self.max_value)
        return q[0]

    def save(self, filename, path=None):
        """dqn.py:187-199: the evaluation network's state_dict as ``<path>/<filename>.pth``."""
        path = config.DPAMSA_WEIGHTS_PATH if path is None else path
        if self.eval_net is not None:
            sd = {k: v.cpu() for k, v in self.eval_net.state_dict().items()}
        elif self.qnet is not None:
            sd = {k: torch.from_numpy(v) for k, v in self.qnet.state.items()}
        else:
            raise RuntimeError("DQN.save: no weights loaded or trained")
        torch.save(sd, os.path.join(path, "{}.pth".format(filename)))
        print("\nModel weights saved as {}.pth in {}\n".format(filename, path))

    # ------------------------------------------------------------------ training: torch autograd (DPAMSA/trainer.py)
    def _ensure_training(self):
        if self.eval_net is not None:
            return
        args = (self.seq_num, self.max_seq_len, self.action_number, self.max_value, self.train_device)
        self.eval_net = TrainNet(*args)
        self.target_net = TrainNet(*args)
        if self.qnet is not None:                                     # continue from the loaded weights
            self.eval_net.load_state_dict({k: torch.from_numpy(v) for k, v in self.qnet.state.items()})
        self.optimizer = torch.optim.Adam(self.eval_net.parameters(), lr=config.ALPHA)

    def update_epsilon(self):
        """dqn.py:123-127"""
        self.epsilon_step_counter += 1
        if self.epsilon_step_counter % config.DECREMENT_ITERATION == 0:
            self.current_epsilon = max(0, self.current_epsilon - config.DELTA)

    def select(self, state):
        """dqn.py:129-149: epsilon-greedy; the greedy branch runs the evaluation network as it is (train mode)."""
        self._ensure_training()
        if random.random() <= self.current_epsilon:
            return np.random.randint(0, self.action_number)
        tokens = torch.LongTensor(state).unsqueeze_(0).to(self.train_device)
        with torch.no_grad():
            values = self.eval_net.forward(tokens)
        return torch.argmax(values, 1).cpu().numpy()[0]

    def update(self):
        """dqn.py:156-185: one optimisation step on a replay batch; returns the loss, or None while the memory
        holds fewer than BATCH_SIZE transitions. The target network is refreshed every UPDATE_ITERATION calls."""
        self._ensure_training()
        self.update_step_counter += 1
        if self.update_step_counter % config.UPDATE_ITERATION == 0:
            self.target_net.load_state_dict(self.eval_net.state_dict())
        if self.replay_memory.size < config.BATCH_SIZE:
            return None
        state, next_state, action, reward, done = self.replay_memory.sample(config.BATCH_SIZE)
        dev = self.train_device
        b_state = torch.LongTensor(state).to(dev)
        b_next = torch.LongTensor(next_state).to(dev)
        b_action = torch.LongTensor(action).to(dev)
        b_reward = torch.FloatTensor(reward).to(dev)
        b_done = torch.FloatTensor(done).to(dev)
        q_eval = self.eval_net(b_state).gather(1, b_action.unsqueeze(-1)).squeeze(1)
        q_next = self.target_net(b_next).max(1)[0].detach()
        q_target = b_reward + b_done * config.GAMMA * q_next         # done == 0 marks the terminal step (env.py:259)
        loss = torch.nn.functional.mse_loss(q_eval, q_target)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        self.qnet = None                                             # frozen copy is stale now
        return loss.item()
