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
    python scripts/l1_precision_flips.py --episodes 3000 > profiles/r02_l1_precision_flips_100k.json
        also simulates that many greedy episodes per window in lock step (float64 actions, windows cut from a
        population generated like GA.generate_population at the c3 shape) and measures the same on every state they
        visit - 100 000+ states per window, a quarter of an hour on 8 cores
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


def measure(sd, states, max_value, chunk=2048):
    """per variant: max |dQ|, flips, and the margins of all states (for the gate statistics), streamed in chunks"""
    W1 = np.asarray(sd["l1.weight"], np.float64)
    s = int(np.floor(np.log2(32768.0 / np.abs(W1).max())))          # the kernel's weight pre-scale 2^s
    wh, wl = split16(W1 * 2.0 ** s)
    w32, wbf = W1.astype(np.float32), to_bf16(W1)
    inv = 2.0 ** -s
    names = ["fp32 (plain float32 product)", "bf16 x bf16, 1 term", "fp16 split, 1 term (hi.hi)",
             "fp16 split, 2 terms (hi.hi + lo.hi)", "fp16 split, 2 terms (hi.hi + hi.lo)",
             "fp16 split, 3 terms (hi.hi + lo.hi + hi.lo) - what the kernel issues"]
    err = {n: 0.0 for n in names}
    flips = {n: 0 for n in names}
    margins, actions = [], []
    for c0 in range(0, states.shape[0], chunk):
        Z, W = pre_l1(sd, states[c0:c0 + chunk])
        b1 = W["l1.bias"]
        q_ref = post_l1(W, Z @ W["l1.weight"].T + b1, max_value)
        a_ref = q_ref.argmax(1)
        srt = np.sort(q_ref, 1)
        margins.append(srt[:, -1] - srt[:, -2])
        actions.append(a_ref)
        zh, zl = split16(Z)
        hh, lh, hl = zh @ wh.T, zl @ wh.T, zh @ wl.T
        hs = [Z.astype(np.float32) @ w32.T, to_bf16(Z) @ wbf.T, hh * inv, (hh + lh) * inv, (hh + hl) * inv, (hh + lh + hl) * inv]
        for n, h in zip(names, hs):
            qv = post_l1(W, h.astype(np.float64) + b1, max_value)
            err[n] = max(err[n], float(np.abs(qv - q_ref).max()))
            flips[n] += int((qv.argmax(1) != a_ref).sum())
    margin = np.concatenate(margins)
    res = {n: {"max_abs_dq_over_max_value": err[n] / max_value, "max_abs_dq": err[n], "flipped_actions": flips[n],
               "states_with_margin_below_2x_error": int((margin < 2 * err[n]).sum())} for n in names}
    return res, margin, np.concatenate(actions), s


def simulate(sd, rows, cols, episodes, max_value, seed):
    """`episodes` greedy episodes in lock step (float64 forward = the reference's actions): every state visited"""
    import random
    rng = random.Random(seed)
    seqs = ["".join(rng.choice("ACGT") for _ in range(60)) for _ in range(6)]
    pop = O.generate_population(seqs, episodes, 0.0, 0.05, rng)              # every board different
    envs = []
    for board in pop:
        grid = O.tile_grid(len(board), min(len(r) for r in board), rows, cols)
        fr, tr, fc, tc = grid[rng.randrange(len(grid))]
        envs.append(O.Env([row[fc:tc] for row in board[fr:tr]]))
    live = list(range(episodes))
    visited = []
    while live:
        st = np.array([envs[i].state() for i in live], np.int64)
        visited.append(st)
        acts = []
        for c0 in range(0, st.shape[0], 2048):
            acts.append(O.qnet_forward(sd, st[c0:c0 + 2048], max_value).argmax(1))
        acts = np.concatenate(acts)
        nxt = []
        for i, a in zip(live, acts):
            _, _, cont = envs[i].step(int(a))                                 # greedy: argmax (dqn.py:153-154)
            if cont:
                nxt.append(i)
        live = nxt
    return np.concatenate(visited)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=0)
    args = ap.parse_args()
    if args.episodes > 0:
        z = np.load(os.path.join(ROOT, "tests", "golden", "qnet_states_big.npz"))
        seed = int(z["weight_seed"])
        out = {"source": "%d greedy episodes per window simulated in lock step with the float64 forward (windows cut from a "
                         "population generated like GA.generate_population: 6x60 sequences, 5 %% gaps), substitute weights "
                         "seed %d" % (args.episodes, seed),
               "emulation": "l1 only; float64 before and after; fp16 split operands, float32 accumulation by BLAS", "windows": {}}
        for rows, cols in ((3, 30), (6, 30)):
            sd = O.synth_qnet_weights(rows, cols, seed)
            max_value = cols * rows * (rows - 1) / 2 * 4.0
            states = simulate(sd, rows, cols, args.episodes, max_value, 99 + rows)
            res, margin, _, s2 = measure(sd, states, max_value)
            out["windows"]["%dx%d" % (rows, cols)] = {
                "states": int(states.shape[0]), "episodes": args.episodes, "max_value": max_value, "weight_prescale_log2": s2,
                "smallest_top2_margin": float(margin.min()),
                "margin_percentiles_1_10_50": [float(np.percentile(margin, p)) for p in (1, 10, 50)], "variants": res}
            sys.stderr.write("%dx%d: %d states done\n" % (rows, cols, states.shape[0]))
        print(json.dumps(out, indent=1))
        return
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
