"""How many greedy actions flip when the first dense layer (l1: 95 % of the weights, 3x the FLOPs of a plain fp16
product in this build) runs with fewer split terms - the measurement DESIGN.md's "3-term split" rests on.

CPU emulation on the recorded reference states (tests/golden/qnet_states_big.npz: 4 096 states per agent window from
the reference's own greedy trajectories, with the reference's action and top-2 margin): everything up to l1's input and
everything after l1 in float64; l1 itself as the kernel forms it - activations Z and weights W * 2^s split into fp16
hi + lo parts, products accumulated in float32 (BLAS order, not the tensor cores' - representative, not bit-equal) -
with 1 term (hi.hi), 2 terms (hi.hi + lo.hi or hi.hi + hi.lo) and the 3 terms the kernel issues, plus plain fp32 and a
bf16 product for scale. Per variant: max |dQ| / max_value against the float64 forward, flipped greedy actions, and the
share of states whose top-2 margin is below twice that error (what a margin-gated scheme would have to re-run).

    python scripts/l1_precision_flips.py > profiles/r02_l1_precision_flips.json        (CPU only, about a minute)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ga_oracle as O          # noqa: E402  (a measurement script: the oracle is the float64 reference here)


def to_bf16(x):
    u = np.asarray(x, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).view(np.float32)


def split16(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float64)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def pre_l1(sd, states):
    """the float64 forward up to l1's input (ga_oracle.qnet_forward, same operations)"""
    W = {k: np.asarray(v, np.float64) for k, v in sd.items()}
    tok = np.asarray(states, np.int64)
    B, L = tok.shape
    x = W["encoder.src_word_emb.weight"][tok] + W["encoder.position_enc.pos_table"][:, :L]
    x = O._layer_norm(x, W["encoder.layer_norm.weight"], W["encoder.layer_norm.bias"])
    q = x @ W["encoder.self_attention.w_qs.weight"].T
    k = x @ W["encoder.self_attention.w_ks.weight"].T
    v = x @ W["encoder.self_attention.w_vs.weight"].T
    att = (q / np.float64(q.shape[-1] ** 0.5)) @ np.swapaxes(k, 1, 2)
    att = np.where((tok != 0)[:, None, :], att, -1e9)
    att = np.exp(att - att.max(-1, keepdims=True))
    att = att / att.sum(-1, keepdims=True)
    o = (att @ v) @ W["encoder.self_attention.fc.weight"].T + x
    o = O._layer_norm(o, W["encoder.self_attention.layer_norm.weight"], W["encoder.self_attention.layer_norm.bias"])
    return o.reshape(B, -1), W


def post_l1(W, h, max_value):
    h = np.where(h > 0, h, 0.01 * h)
    h = h @ W["l2.weight"].T + W["l2.bias"]
    h = np.where(h > 0, h, 0.01 * h)
    h = h @ W["l3.weight"].T + W["l3.bias"]
    return np.tanh(h) * max_value


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "qnet_states_big.npz"))
    seed = int(z["weight_seed"])
    out = {"source": "tests/golden/qnet_states_big.npz (reference greedy trajectories), substitute weights seed %d" % seed,
           "emulation": "l1 only; float64 before and after; fp16 split operands, float32 accumulation by BLAS", "windows": {}}
    for rows, cols in ((3, 30), (6, 30)):
        tag = "%dx%d" % (rows, cols)
        sd = O.synth_qnet_weights(rows, cols, seed)
        states = z["states_" + tag].astype(np.int64)
        max_value = cols * rows * (rows - 1) / 2 * 4.0
        Z, W = pre_l1(sd, states)
        W1, b1 = W["l1.weight"], W["l1.bias"]
        q_ref = post_l1(W, Z @ W1.T + b1, max_value)
        a_ref = q_ref.argmax(1)
        srt = np.sort(q_ref, 1)
        margin = srt[:, -1] - srt[:, -2]
        s = int(np.floor(np.log2(32768.0 / np.abs(W1).max())))          # the kernel's weight pre-scale 2^s
        zh, zl = split16(Z)
        wh, wl = split16(W1 * 2.0 ** s)
        inv = 2.0 ** -s
        hh = zh @ wh.T
        lh = zl @ wh.T
        hl = zh @ wl.T
        variants = {
            "fp32 (plain float32 product)": Z.astype(np.float32) @ W1.astype(np.float32).T,
            "bf16 x bf16, 1 term": (to_bf16(Z) @ to_bf16(W1).T),
            "fp16 split, 1 term (hi.hi)": hh * inv,
            "fp16 split, 2 terms (hi.hi + lo.hi)": (hh + lh) * inv,
            "fp16 split, 2 terms (hi.hi + hi.lo)": (hh + hl) * inv,
            "fp16 split, 3 terms (hi.hi + lo.hi + hi.lo) - what the kernel issues": (hh + lh + hl) * inv,
        }
        res = {}
        for name, h in variants.items():
            qv = post_l1(W, h.astype(np.float64) + b1, max_value)
            err = float(np.abs(qv - q_ref).max())
            flips = int((qv.argmax(1) != a_ref).sum())
            res[name] = {"max_abs_dq_over_max_value": err / max_value, "max_abs_dq": err, "flipped_actions": flips,
                         "states_with_margin_below_2x_error": int((margin < 2 * err).sum())}
        out["windows"][tag] = {"states": int(states.shape[0]), "max_value": max_value, "weight_prescale_log2": s,
                               "reference_matches_recorded_actions": int((a_ref == z["action_" + tag]).sum()),
                               "smallest_top2_margin": float(margin.min()),
                               "margin_percentiles_1_10_50": [float(np.percentile(margin, p)) for p in (1, 10, 50)],
                               "variants": res}
    print(json.dumps(out, indent=1))


main()
