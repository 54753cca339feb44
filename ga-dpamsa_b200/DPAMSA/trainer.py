"""The trainable form of the Q-network (reference DPAMSA/models.py:54-249, dqn.py:26-79) for ``DQN.select`` /
``DQN.update`` / ``DPAMSA.main.train`` - SURVEY.md section 8(f) rank 1: producing the weight files the generation
loop consumes.

Training is not on the generation loop and runs on torch autograd (library code, like the reference's own trainer),
not on libgadpamsa: the folded tables of csrc/qnet.cu are built once per weight file and would have to be rebuilt
after every optimiser step. What this module guarantees instead is the *contract* with the hot path:

* parameters live in a flat dict under the reference's ``state_dict`` keys (``STATE_KEYS`` of DPAMSA/dqn.py), so a file
  written by ``DQN.save`` after training loads into ``gad_qnet_create`` and into the reference unchanged;
* the forward issues the same torch operators in the same order and shapes as the reference modules (embedding +
  sinusoid positions, dropout 0.1, LayerNorm(eps=1e-6), single-head attention with d_k = d_v = 164 and the
  ``-1e9`` padding mask, dropout on the attention weights and on the output projection, residual, LayerNorm,
  l1 / l2 with LeakyReLU, l3 with tanh times ``max_value``), so with the same ``torch.manual_seed`` it consumes the
  dropout stream identically and a seeded update step reproduces the reference's loss and weights
  (tests/test_training_cpu.py against tests/golden/train_vectors.json).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

D_MODEL = 64          # dqn.py:54
D_ATT = 164           # models.py:231 (d_k = d_v)
H1, H2 = 1028, 512    # dqn.py:57-61
P_DROP = 0.1          # models.py:66,153,232

TRAINABLE = (
    "encoder.src_word_emb.weight",
    "encoder.layer_norm.weight", "encoder.layer_norm.bias",
    "encoder.self_attention.w_qs.weight", "encoder.self_attention.w_ks.weight",
    "encoder.self_attention.w_vs.weight", "encoder.self_attention.fc.weight",
    "encoder.self_attention.layer_norm.weight", "encoder.self_attention.layer_norm.bias",
    "l1.weight", "l1.bias", "l2.weight", "l2.bias", "l3.weight", "l3.bias",
)
POS_KEY = "encoder.position_enc.pos_table"


def sinusoid_table(n_position: int, d_hid: int = D_MODEL) -> torch.Tensor:
    """models.py:113-133: angle(pos, j) = pos / 10000^(2*(j//2)/d); sin on even, cos on odd channels; fp64 then fp32."""
    j = np.arange(d_hid)
    angle = np.arange(n_position, dtype=np.float64)[:, None] / np.power(10000.0, 2 * (j // 2) / d_hid)[None, :]
    angle[:, 0::2] = np.sin(angle[:, 0::2])
    angle[:, 1::2] = np.cos(angle[:, 1::2])
    return torch.from_numpy(angle.astype(np.float32))[None]


def _linear_init(out_f: int, in_f: int, bias: bool, gen):
    """torch.nn.Linear's default: kaiming_uniform(a=sqrt(5)) = U(-1/sqrt(in), 1/sqrt(in)) for weight and bias."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound if bias else None
    return w, b


class TrainNet:
    """Parameters + forward of the reference ``Net`` (dqn.py:26-79) as a flat dict of leaf tensors."""

    def __init__(self, seq_num: int, max_seq_len: int, action_number: int, max_value: float, device="cpu", generator=None):
        self.L = seq_num * (max_seq_len + 1)
        self.action_number = action_number
        self.max_value = float(max_value)
        self.device = torch.device(device)
        self.training = True                      # the reference never leaves train mode
        p = {}
        emb = torch.randn(6, D_MODEL, generator=generator)
        emb[0] = 0                                # padding_idx=0 (models.py:235)
        p["encoder.src_word_emb.weight"] = emb
        for ln in ("encoder.layer_norm", "encoder.self_attention.layer_norm"):
            p[ln + ".weight"] = torch.ones(D_MODEL)
            p[ln + ".bias"] = torch.zeros(D_MODEL)
        p["encoder.self_attention.w_qs.weight"], _ = _linear_init(D_ATT, D_MODEL, False, generator)
        p["encoder.self_attention.w_ks.weight"], _ = _linear_init(D_ATT, D_MODEL, False, generator)
        p["encoder.self_attention.w_vs.weight"], _ = _linear_init(D_ATT, D_MODEL, False, generator)
        p["encoder.self_attention.fc.weight"], _ = _linear_init(D_MODEL, D_ATT, False, generator)
        p["l1.weight"], p["l1.bias"] = _linear_init(H1, self.L * D_MODEL, True, generator)
        p["l2.weight"], p["l2.bias"] = _linear_init(H2, H1, True, generator)
        p["l3.weight"], p["l3.bias"] = _linear_init(action_number, H2, True, generator)
        self.params = {k: p[k].to(self.device).requires_grad_(True) for k in TRAINABLE}
        self.pos_table = sinusoid_table(self.L).to(self.device)

    # ---- state_dict contract (dqn.py:187-213) ----
    def parameters(self):
        return [self.params[k] for k in TRAINABLE]

    def state_dict(self):
        out = {k: v.detach().clone() for k, v in self.params.items()}
        out[POS_KEY] = self.pos_table.clone()
        return out

    def load_state_dict(self, sd):
        missing = [k for k in TRAINABLE + (POS_KEY,) if k not in sd]
        if missing:
            raise RuntimeError("Missing key(s) in state_dict: %s" % ", ".join(missing))
        with torch.no_grad():
            for k in TRAINABLE:
                src = torch.as_tensor(sd[k])
                if tuple(src.shape) != tuple(self.params[k].shape):
                    raise RuntimeError("size mismatch for %s: checkpoint %s, model %s"
                                       % (k, tuple(src.shape), tuple(self.params[k].shape)))
                self.params[k].copy_(src.to(self.device, torch.float32))
            self.pos_table = torch.as_tensor(sd[POS_KEY]).to(self.device, torch.float32).clone()

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    # ---- forward (models.py:71-95,169-207,241-249; dqn.py:68-79) ----
    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        p, tr = self.params, self.training
        mask = (tokens != 0).unsqueeze(-2)                                  # [B,1,L]: keys that are not padding
        x = F.embedding(tokens, p["encoder.src_word_emb.weight"], padding_idx=0)
        x = x + self.pos_table[:, :x.size(1)].clone().detach()
        x = F.dropout(x, P_DROP, tr)
        x = F.layer_norm(x, (D_MODEL,), p["encoder.layer_norm.weight"], p["encoder.layer_norm.bias"], 1e-6)
        B, n = x.size(0), x.size(1)
        residual = x
        q = F.linear(x, p["encoder.self_attention.w_qs.weight"]).view(B, n, 1, D_ATT).transpose(1, 2)
        k = F.linear(x, p["encoder.self_attention.w_ks.weight"]).view(B, n, 1, D_ATT).transpose(1, 2)
        v = F.linear(x, p["encoder.self_attention.w_vs.weight"]).view(B, n, 1, D_ATT).transpose(1, 2)
        attn = torch.matmul(q / (D_ATT ** 0.5), k.transpose(2, 3))
        attn = attn.masked_fill(mask.unsqueeze(1) == 0, -1e9)
        attn = F.dropout(F.softmax(attn, dim=-1), P_DROP, tr)
        o = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, n, -1)
        o = F.dropout(F.linear(o, p["encoder.self_attention.fc.weight"]), P_DROP, tr)
        o = o + residual
        o = F.layer_norm(o, (D_MODEL,), p["encoder.self_attention.layer_norm.weight"],
                         p["encoder.self_attention.layer_norm.bias"], 1e-6)
        h = o.view(B, -1)
        h = F.leaky_relu(F.linear(h, p["l1.weight"], p["l1.bias"]))
        h = F.leaky_relu(F.linear(h, p["l2.weight"], p["l2.bias"]))
        h = torch.tanh(F.linear(h, p["l3.weight"], p["l3.bias"]))
        return torch.mul(h, self.max_value)

    __call__ = forward
