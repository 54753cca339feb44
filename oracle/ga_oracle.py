"""CPU restatement ("oracle") of the GA-DPAMSA generation loop.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
(``ga-dpamsa_b200/``) never does and fails loudly without its CUDA library.

Parity pin: every function is checked (tests/test_oracle_pins.py) against the
unmodified reference run in-process through ``oracle/ref_shim.py`` when the
reference checkout is present, and against the committed fixtures in
``tests/golden/`` (made by ``oracle/make_golden.py`` from the reference and
its 1340 published report records) everywhere else.
Exception: ``vertical_crossover`` has NO reference code
-> "parity unpinned" for that one function.

Citations are ``file:line`` into the reference checkout.  Boards are python
lists of rows of ints, exactly the reference's representation (GA.py:35:
A=1 T=2 C=3 G=4 gap=5; 0 is only the Q-net padding token).
"""
from __future__ import annotations

import ctypes
import math
import os
import random as _pyrandom
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NUC = {'A': 1, 'T': 2, 'C': 3, 'G': 4, 'a': 1, 't': 2, 'c': 3, 'g': 4, '-': 5}   # GA.py:35
LETTERS = "ATCG-"                                                              # utils.py:458
GAP = 5


class Weights:
    """config.py:21-23"""

    def __init__(self, match=4, mismatch=-4, gap=-4):
        self.match, self.mismatch, self.gap = match, mismatch, gap


DEFAULT_W = Weights()


# --------------------------------------------------------------------------- scoring
def encode(seqs):
    """GA.py:79 / env.py:67 - KeyError on an unknown letter, like the reference."""
    return [[NUC[ch] for ch in s] for s in seqs]


def decode(board):
    """utils.py:464-467"""
    return ["".join(LETTERS[v - 1] for v in row) for row in board]


def sum_of_pairs(board, fr, tr, fc, tc, w=DEFAULT_W):
    """utils.py:62-76 via per-column symbol counts: with counts c1..c4 and g
    gaps among n=tr-fr cells, match pairs = sum c(c-1)/2, gap pairs =
    g(g-1)/2 + g(n-g), the rest mismatch (same totals as the j<k loop)."""
    n = tr - fr
    total = 0
    for c in range(fc, tc):
        cnt = [0] * 6
        for r in range(fr, tr):
            cnt[board[r][c]] += 1
        g = cnt[GAP]
        m = sum(cnt[k] * (cnt[k] - 1) // 2 for k in (1, 2, 3, 4))
        gp = g * (g - 1) // 2 + g * (n - g)
        total += w.match * m + w.gap * gp + w.mismatch * (n * (n - 1) // 2 - m - gp)
    return total


def sum_of_pairs_literal(board, fr, tr, fc, tc, w=DEFAULT_W):
    """utils.py:62-76 as the literal pair loop (used to pin the count form)."""
    s = 0
    for c in range(fc, tc):
        for j in range(fr, tr):
            for k in range(j + 1, tr):
                a, b = board[j][c], board[k][c]
                s += w.gap if (a == GAP or b == GAP) else (w.match if a == b else w.mismatch)
    return s


def exact_columns(board, fr, tr, fc, tc):
    """utils.py:122-125 (numerator of the column score)."""
    return sum(1 for c in range(fc, tc) if all(board[r][c] == board[fr][c] for r in range(fr, tr)))


def column_score(board, fr, tr, fc, tc):
    """utils.py:119-128 - python true division, 0 when there are no columns."""
    ncol = tc - fc
    return exact_columns(board, fr, tr, fc, tc) / ncol if ncol > 0 else 0


def pad_rows(board):
    """GA.py:158-162"""
    width = max(map(len, board))
    for row in board:
        row.extend([GAP] * (width - len(row)))


def clean_unnecessary_gaps(board):
    """utils.py:404-428, in place.  Net effect: every all-gap column goes."""
    if not board or not board[0]:
        return
    def tail(row):
        i = len(row)
        while i > 0 and row[i - 1] == GAP:
            i -= 1
        return i
    cut = max(tail(row) for row in board)
    for row in board:
        del row[cut:]
    doomed = [c for c in range(len(board[0])) if all(row[c] == GAP for row in board)]
    for c in reversed(doomed):
        for row in board:
            del row[c]


def fitness_scores(population, mode, w=DEFAULT_W, keep_zero=True):
    """GA.py:154-178 (without the HoF update).  Mutates the boards in place.

    ``keep_zero=False`` reproduces GA.py:178 literally (``sp or cs`` turns an
    SP of 0 into None); the product keeps 0 - documented deviation."""
    scores = []
    for i, board in enumerate(population):
        pad_rows(board)
        clean_unnecessary_gaps(board)
        width = max(map(len, board))
        sp = sum_of_pairs(board, 0, len(board), 0, width, w) if mode in ('sp', 'mo') else None
        cs = column_score(board, 0, len(board), 0, width) if mode in ('cs', 'mo') else None
        if mode == 'mo':
            scores.append((i, sp, cs))
        elif keep_zero:
            scores.append((i, sp if mode == 'sp' else cs))
        else:
            scores.append((i, sp or cs))
    return scores


def fitness_scores_literal(population, mode, w=DEFAULT_W):
    """GA.py:154-178 with the reference's own loop structure (pair loop of utils.py:62-76,
    all() scan of utils.py:122-125) - the CPU cost model of the reference for bench.py's
    python-port baseline. Same results as ``fitness_scores``."""
    scores = []
    for i, board in enumerate(population):
        pad_rows(board)
        clean_unnecessary_gaps(board)
        width = max(map(len, board))
        sp = sum_of_pairs_literal(board, 0, len(board), 0, width, w) if mode in ('sp', 'mo') else None
        cs = column_score(board, 0, len(board), 0, width) if mode in ('cs', 'mo') else None
        scores.append((i, sp, cs) if mode == 'mo' else (i, sp if mode == 'sp' else cs))
    return scores


# --------------------------------------------------------------------------- ranking / HoF / selection
def _norm(v, lo, hi):
    """utils.py:343-344 / utils.py:302-303"""
    return (v - lo) / (hi - lo) if hi > lo else 0.5


def rank_best(scores, n):
    """utils.py:332-355 - stable descending sort (ties keep ascending position)."""
    if len(scores[0]) == 2:
        order = sorted(scores, key=lambda t: t[1], reverse=True)
        return [t[0] for t in order[:n]]
    sps = [t[1] for t in scores]
    css = [t[2] for t in scores]
    lo_s, hi_s, lo_c, hi_c = min(sps), max(sps), min(css), max(css)
    keyed = [(t[0], _norm(t[1], lo_s, hi_s), _norm(t[2], lo_c, hi_c)) for t in scores]
    keyed.sort(key=lambda t: t[1] + t[2], reverse=True)
    return [t[0] for t in keyed[:n]]


def hof_update(hof, population, scores):
    """GA.py:110-136.  ``hof`` is None/() or (board, score_tuple); returns the
    new (board_copy, score_tuple).  The HoF entrant is appended with the fresh
    index len(scores) (GA.py:127-132), so on equal keys a population member
    (lower position in the stable sort) replaces it."""
    import copy
    if not hof:
        b = rank_best(scores, 1)[0]
        return (copy.deepcopy(population[b]), tuple(scores[b]))
    pool = list(scores)
    if len(pool[0]) == 2:
        pool.append((len(pool), hof[1][1]))
    else:
        pool.append((len(pool), hof[1][1], hof[1][2]))
    b = rank_best(pool, 1)[0]
    board = hof[0] if b == len(scores) else population[b]
    return (copy.deepcopy(board), tuple(pool[b]))


def pareto_front(scores):
    """GA.py:195-211 - literal O(P^2) dominance test, output in input order."""
    front = []
    for i, (ia, sa, ca) in enumerate(scores):
        dominated = False
        for j, (_, sb, cb) in enumerate(scores):
            if i != j and (sb >= sa and cb >= ca) and (sb > sa or cb > ca):
                dominated = True
                break
        if not dominated:
            front.append((ia, sa, ca))
    return front


def pareto_front_fast(scores):
    """Same set as ``pareto_front`` in O(P log P): walk groups of equal SP from
    high to low; a point is dominated iff a strictly-higher-SP point has
    cs >= its cs, or a same-SP point has cs > its cs."""
    order = sorted(range(len(scores)), key=lambda i: -scores[i][1])
    keep = [False] * len(scores)
    best_cs_above = -math.inf
    k = 0
    while k < len(order):
        j = k
        sp = scores[order[k]][1]
        while j < len(order) and scores[order[j]][1] == sp:
            j += 1
        group_max = max(scores[i][2] for i in order[k:j])
        for i in order[k:j]:
            cs = scores[i][2]
            keep[i] = not (best_cs_above >= cs or group_max > cs)
        best_cs_above = max(best_cs_above, group_max)
        k = j
    return [scores[i] for i in range(len(scores)) if keep[i]]


def selection_keep(scores, mode, population_size, selection_rate, fast=False):
    """GA.py:232-256 - the SET of surviving positions (GA.py:259 keeps the
    boards whose position is in the set, in original order)."""
    top_n = int(population_size * selection_rate)
    if mode in ('sp', 'cs'):
        order = sorted(scores, key=lambda t: t[1], reverse=True)
        return {t[0] for t in order[:top_n]}
    front = (pareto_front_fast if fast else pareto_front)(scores)
    keep = {t[0] for t in front}
    if len(keep) < top_n:
        by_sp = sorted(scores, key=lambda t: t[1], reverse=True)
        by_cs = sorted(scores, key=lambda t: t[2], reverse=True)
        keep |= {t[0] for t in by_sp[:top_n]} | {t[0] for t in by_cs[:top_n]}
    else:
        best = sorted(front, key=lambda t: (t[1], t[2]), reverse=True)
        keep = {t[0] for t in best[:top_n]}
    return keep


# --------------------------------------------------------------------------- RNG-driven phases
def generate_population(seqs, population_size, clone_rate, gap_rate, rng=_pyrandom):
    """GA.py:71-94 - clones first, then rows with gaps inserted one at a time at
    rng.randint(0, len-1) (len grows with every insert)."""
    n_clone = round(population_size * clone_rate)
    pop = [encode(seqs) for _ in range(n_clone)]
    for _ in range(population_size - n_clone):
        board = []
        for s in seqs:
            row = [NUC[ch] for ch in s]
            for _ in range(max(1, round(len(row) * gap_rate))):
                row.insert(rng.randint(0, len(row) - 1), GAP)
            board.append(row)
        pop.append(board)
    return pop


def crossover_plan(n_parents, n_rows, n_children, rng=_pyrandom):
    """GA.py:281-293 - the (parent1, parent2, cut) triples, consuming the RNG
    exactly as the reference does: two choice() draws per attempt, retry when
    both land on the same parent, then randint(1, rows-1)."""
    if n_children > 0 and (n_parents < 2 or n_rows < 2):
        raise ValueError("reference loops forever here (GA.py:285-290)")
    plan = []
    parents = range(n_parents)
    while len(plan) < n_children:
        a = rng.choice(parents)
        b = rng.choice(parents)
        if a == b:
            continue
        plan.append((a, b, rng.randint(1, n_rows - 1)))
    return plan


def horizontal_child(p1, p2, cut):
    """GA.py:296 - whole rows: p1's rows above the cut, p2's from the cut down."""
    return [list(r) for r in p1[:cut]] + [list(r) for r in p2[cut:]]


def vertical_child(p1, p2, cut):
    """PARITY UNPINNED - the reference has no vertical crossover code (only
    img/vertical-crossover.png and supplement table A5).  Defined by analogy
    with GA.py:296: every row keeps p1's first ``cut`` cells and p2's rest."""
    return [list(a[:cut]) + list(b[cut:]) for a, b in zip(p1, p2)]


# --------------------------------------------------------------------------- tiles
def tile_grid(n_rows, min_len, rw, cw):
    """utils.py:230-252 restated in closed form: the scan accepts a window iff
    it is in bounds and overlaps nothing accepted before, which yields exactly
    the full tiles with origins on the (rw, cw) lattice, row-major."""
    return [(i, i + rw, j, j + cw)
            for i in range(0, n_rows - rw + 1, rw)
            for j in range(0, min_len - cw + 1, cw)]


def tile_grid_literal(board, rw, cw):
    """utils.py:230-252 literally (pins ``tile_grid``)."""
    m, n = len(board), min(len(r) for r in board)
    used = []
    def overlap(a, b):                                          # utils.py:155-162
        return a[0] < b[1] and a[1] > b[0] and a[2] < b[3] and a[3] > b[2]
    for i in range(m):
        for j in range(n):
            cand = (i, i + rw, j, j + cw)
            if not any(overlap(cand, u) for u in used) and cand[1] <= m and cand[3] <= n:
                used.append(cand)
    return used


def worst_tile(board, mode, rw, cw, w=DEFAULT_W):
    """utils.py:276-313 - (score, tile); first minimum wins."""
    tiles = tile_grid(len(board), min(len(r) for r in board), rw, cw)
    rec = [(sum_of_pairs(board, *t, w), column_score(board, *t), t) for t in tiles]
    if mode == 'sp':
        best = min(rec, key=lambda x: x[0])
    elif mode == 'cs':
        best = min(rec, key=lambda x: x[1])
    else:
        sps = [x[0] for x in rec]
        css = [x[1] for x in rec]
        lo_s, hi_s, lo_c, hi_c = min(sps), max(sps), min(css), max(css)
        nrm = [(_norm(s, lo_s, hi_s), _norm(c, lo_c, hi_c), t) for s, c, t in rec]
        best = min(nrm, key=lambda x: x[0] + x[1])
    return best[0], best[2]


# --------------------------------------------------------------------------- environment (DPAMSA/env.py)
class Env:
    """env.py:61-90,139-174,225-259,338-351 without the GUI."""

    def __init__(self, data, w=DEFAULT_W):
        self.data = [list(r) for r in data]
        self.row = len(data)
        self.max_len = max(len(r) for r in data)
        self.action_number = 2 ** self.row - 1                 # env.py:78
        self.max_reward = self.row * (self.row - 1) / 2 * w.match   # env.py:79
        self.w = w
        self.reset()

    def state(self):
        """env.py:147-153"""
        s = []
        for rest in self.not_aligned:
            s.extend(rest)
            s.append(GAP)
        s.extend([0] * (self.row * (self.max_len + 1) - len(s)))
        return s

    def reset(self):
        self.aligned = [[] for _ in range(self.row)]
        self.not_aligned = [list(r) for r in self.data]
        return self.state()

    def step(self, action):
        """env.py:246-259 -> (reward, state, done)"""
        for bit in range(self.row):
            if not (action >> bit) & 1 and not self.not_aligned[bit]:
                return -self.max_reward, self.state(), 0
        left = 0
        for bit in range(self.row):
            if not (action >> bit) & 1:
                self.aligned[bit].append(self.not_aligned[bit].pop(0))
            else:
                self.aligned[bit].append(GAP)
            left += len(self.not_aligned[bit])
        col = len(self.aligned[0]) - 1                          # env.py:163-172
        reward = sum_of_pairs(self.aligned, 0, self.row, col, col + 1, self.w)
        return reward, self.state(), 1 if left > 0 else 0

    def padding(self):
        """env.py:344-351"""
        longest = max(len(r) for r in self.not_aligned)
        for i in range(self.row):
            rest = self.not_aligned[i]
            self.aligned[i].extend(rest)
            self.aligned[i].extend([GAP] * (longest - len(rest)))
            rest.clear()


# --------------------------------------------------------------------------- Q-network (DPAMSA/dqn.py, DPAMSA/models.py), eval mode
def sinusoid_table(n_position, d_hid):
    """models.py:129-136"""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    ang = pos / np.power(10000.0, 2 * (j // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return ang.astype(np.float32)


QNET_KEYS = (
    "encoder.src_word_emb.weight", "encoder.position_enc.pos_table",
    "encoder.layer_norm.weight", "encoder.layer_norm.bias",
    "encoder.self_attention.w_qs.weight", "encoder.self_attention.w_ks.weight",
    "encoder.self_attention.w_vs.weight", "encoder.self_attention.fc.weight",
    "encoder.self_attention.layer_norm.weight", "encoder.self_attention.layer_norm.bias",
    "l1.weight", "l1.bias", "l2.weight", "l2.bias", "l3.weight", "l3.bias",
)


def synth_qnet_weights(rows, cols, seed, d_model=64, d_k=164, h1=1028, h2=512):
    """Deterministic substitute weights in the reference state_dict layout
    (dqn.py:54-63, models.py:158-166,235-239).  numpy-seeded so the same file
    can be rebuilt on the GPU box; scales follow torch's default inits."""
    rng = np.random.default_rng(seed)
    L = rows * (cols + 1)
    A = 2 ** rows - 1

    def lin(out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        return rng.uniform(-b, b, size=(out_f, in_f)).astype(np.float32)

    def bias(out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        return rng.uniform(-b, b, size=(out_f,)).astype(np.float32)

    emb = rng.standard_normal((6, d_model)).astype(np.float32)
    emb[0] = 0.0                                                # padding_idx=0
    sd = {
        "encoder.src_word_emb.weight": emb,
        "encoder.position_enc.pos_table": sinusoid_table(L, d_model)[None],
        "encoder.layer_norm.weight": (1 + 0.1 * rng.standard_normal(d_model)).astype(np.float32),
        "encoder.layer_norm.bias": (0.1 * rng.standard_normal(d_model)).astype(np.float32),
        "encoder.self_attention.w_qs.weight": lin(d_k, d_model),
        "encoder.self_attention.w_ks.weight": lin(d_k, d_model),
        "encoder.self_attention.w_vs.weight": lin(d_k, d_model),
        "encoder.self_attention.fc.weight": lin(d_model, d_k),
        "encoder.self_attention.layer_norm.weight": (1 + 0.1 * rng.standard_normal(d_model)).astype(np.float32),
        "encoder.self_attention.layer_norm.bias": (0.1 * rng.standard_normal(d_model)).astype(np.float32),
        "l1.weight": lin(h1, L * d_model), "l1.bias": bias(h1, L * d_model),
        "l2.weight": lin(h2, h1), "l2.bias": bias(h2, h1),
        "l3.weight": lin(A, h2), "l3.bias": bias(A, h2),
    }
    return sd


def _layer_norm(x, g, b, eps=1e-6):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * g + b


def qnet_forward(sd, states, max_value, dtype=np.float64):
    """dqn.py:69-79 + models.py:241-249,190-207,87-95 in eval mode (dropout =
    identity).  ``states``: int array (B, L).  Returns Q (B, A) in ``dtype``
    (float64 by default: the tolerance reference for the 1e-3 Q-value bar)."""
    W = {k: np.asarray(v, dtype=dtype) for k, v in sd.items()}
    tok = np.asarray(states, dtype=np.int64)
    B, L = tok.shape
    x = W["encoder.src_word_emb.weight"][tok] + W["encoder.position_enc.pos_table"][:, :L]
    x = _layer_norm(x, W["encoder.layer_norm.weight"], W["encoder.layer_norm.bias"])
    q = x @ W["encoder.self_attention.w_qs.weight"].T
    k = x @ W["encoder.self_attention.w_ks.weight"].T
    v = x @ W["encoder.self_attention.w_vs.weight"].T
    d_k = q.shape[-1]
    att = (q / dtype(d_k ** 0.5)) @ np.swapaxes(k, 1, 2)
    att = np.where((tok != 0)[:, None, :], att, dtype(-1e9))     # dqn.py:67, models.py:90-91,197
    att = att - att.max(-1, keepdims=True)
    att = np.exp(att)
    att = att / att.sum(-1, keepdims=True)
    o = (att @ v) @ W["encoder.self_attention.fc.weight"].T + x
    o = _layer_norm(o, W["encoder.self_attention.layer_norm.weight"],
                    W["encoder.self_attention.layer_norm.bias"])
    h = o.reshape(B, -1) @ W["l1.weight"].T + W["l1.bias"]
    h = np.where(h > 0, h, dtype(0.01) * h)
    h = h @ W["l2.weight"].T + W["l2.bias"]
    h = np.where(h > 0, h, dtype(0.01) * h)
    h = h @ W["l3.weight"].T + W["l3.bias"]
    return np.tanh(h) * dtype(max_value)


def qnet_forward_torch(sd_t, states, max_value):
    """The same forward with the torch CPU ops the reference itself executes (dqn.py:69-79,
    models.py:87-95,190-207,241-249), eval mode, fp32.  ``sd_t``: state_dict of torch tensors.
    This is the reference's real CPU cost model (bench.py's mutation baseline); results agree with
    ``qnet_forward`` to fp32 rounding."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        tok = torch.as_tensor(np.asarray(states), dtype=torch.long)
        L = tok.shape[1]
        x = F.embedding(tok, sd_t["encoder.src_word_emb.weight"], padding_idx=0)
        x = x + sd_t["encoder.position_enc.pos_table"][:, :L]
        x = F.layer_norm(x, (x.shape[-1],), sd_t["encoder.layer_norm.weight"], sd_t["encoder.layer_norm.bias"], 1e-6)
        q = F.linear(x, sd_t["encoder.self_attention.w_qs.weight"])
        k = F.linear(x, sd_t["encoder.self_attention.w_ks.weight"])
        v = F.linear(x, sd_t["encoder.self_attention.w_vs.weight"])
        att = torch.matmul(q / (q.shape[-1] ** 0.5), k.transpose(1, 2))
        att = att.masked_fill((tok != 0).unsqueeze(-2) == 0, -1e9)
        att = F.softmax(att, dim=-1)
        o = F.linear(torch.matmul(att, v), sd_t["encoder.self_attention.fc.weight"]) + x
        o = F.layer_norm(o, (o.shape[-1],), sd_t["encoder.self_attention.layer_norm.weight"],
                         sd_t["encoder.self_attention.layer_norm.bias"], 1e-6)
        h = F.leaky_relu(F.linear(o.reshape(o.shape[0], -1), sd_t["l1.weight"], sd_t["l1.bias"]))
        h = F.leaky_relu(F.linear(h, sd_t["l2.weight"], sd_t["l2.bias"]))
        h = torch.tanh(F.linear(h, sd_t["l3.weight"], sd_t["l3.bias"]))
        return torch.mul(h, max_value).numpy()


def greedy_action(qvalues):
    """dqn.py:153-154 - argmax, first index on ties."""
    return int(np.argmax(qvalues))


def mutate_board(board, tile, sd, w=DEFAULT_W, qdtype=np.float32, trace=None, forward=None):
    """GA.py:335-373 for one individual, eval-mode agent.  In place.  ``forward(states, max_value)``
    replaces the numpy forward (e.g. ``qnet_forward_torch`` bound to torch weights)."""
    fr, tr, fc, tc = tile
    sub = [row[fc:tc] for row in board[fr:tr]]
    env = Env(sub, w)
    max_value = env.max_len * env.max_reward                     # GA.py:354
    state = env.reset()
    while True:
        q = (forward([state], max_value) if forward else qnet_forward(sd, [state], max_value, dtype=qdtype))[0]
        a = greedy_action(q)
        if trace is not None:
            trace.append((list(state), a, q))
        _, state, done = env.step(a)
        if done == 0:
            break
    env.padding()
    for i, row in enumerate(env.aligned):
        board[fr + i][fc:tc] = row


def mutation(population, scores, mode, n_mutate, rw, cw, sd, w=DEFAULT_W, qdtype=np.float32,
             tile_choice="worst", probability=1.0, rng=_pyrandom):
    """GA.py:323-373.  Paper-style variants (PARITY UNPINNED, no reference code; supplement tables A3 / A5):
    ``probability`` < 1 - every selected individual mutates only when rng.random() < probability;
    ``tile_choice`` 'random' - the tile is drawn uniformly from the lattice of utils.py:230-252 with
    rng.randrange instead of being the worst-fitted one.  Draw order per individual: random(), then randrange."""
    for idx in rank_best(scores, n_mutate):
        if probability < 1.0 and rng.random() >= probability:
            continue
        board = population[idx]
        if tile_choice == "random":
            tiles = tile_grid(len(board), min(len(r) for r in board), rw, cw)
            if not tiles:
                raise ValueError("min() arg is an empty sequence")
            tile = tiles[rng.randrange(len(tiles))]
        else:
            _, tile = worst_tile(board, mode, rw, cw, w)
        mutate_board(board, tile, sd, w, qdtype)


def tournament_keep(scores, mode, population_size, selection_rate, tournament_rate, rng=_pyrandom):
    """Tournament selection (PARITY UNPINNED: supplement table A3 names it, the reference has no code).  top_n =
    int(population_size * selection_rate) tournaments without replacement: max(2, ceil(tournament_rate * pool))
    contestants drawn with rng.sample from the individuals still in the pool (ascending position), the best
    ranked one (utils.py:332-351 ordering, ties to the lower position) survives and leaves the pool."""
    import math
    top_n = min(int(population_size * selection_rate), len(scores))
    order = rank_best(scores, len(scores))
    rank = {pos: r for r, pos in enumerate(order)}
    pool = list(range(len(scores)))
    keep = set()
    for _ in range(top_n):
        t = min(len(pool), max(2, math.ceil(tournament_rate * len(pool))))
        winner = min(rng.sample(pool, t), key=rank.__getitem__)
        keep.add(winner)
        pool.remove(winner)
    return keep


def crossover_plan_variant(n_parents, n_rows, n_children, probability, copy_cut, rng=_pyrandom):
    """crossover_plan with a crossover probability (PARITY UNPINNED, supplement table A3 c_prob): after the
    reference's draws for a child (two choice(), retry on equal parents, randint) one rng.random(); at or above
    ``probability`` the child is a copy of parent1 (cut = copy_cut, i.e. every row / cell from parent1)."""
    if n_children > 0 and (n_parents < 2 or n_rows < 2):
        raise ValueError("reference loops forever here (GA.py:285-290)")
    plan = []
    parents = range(n_parents)
    while len(plan) < n_children:
        a = rng.choice(parents)
        b = rng.choice(parents)
        if a == b:
            continue
        cut = rng.randint(1, n_rows - 1)
        if rng.random() >= probability:
            cut = copy_cut
        plan.append((a, b, cut))
    return plan


# --------------------------------------------------------------------------- the loop (GA.py:443-545)
class Knobs:
    """config.py:39-46 defaults"""

    def __init__(self, **kw):
        self.AGENT_WINDOW_ROW = 3
        self.AGENT_WINDOW_COLUMN = 30
        self.GA_ITERATIONS = 3
        self.POPULATION_SIZE = 5
        self.CLONE_RATE = 0.25
        self.GAP_RATE = 0.05
        self.SELECTION_RATE = 0.5
        self.MUTATION_RATE = 0.25
        # paper-style variants (extensions without reference code; the defaults are the reference's behaviour)
        self.MUTATION_TILE = "worst"
        self.MUTATION_PROBABILITY = 1.0
        self.CROSSOVER_PROBABILITY = 1.0
        self.SELECTION_SCHEME = "truncation"
        self.TOURNAMENT_RATE = 0.25
        self.ELITISM_PROBABILITY = 0.0
        self.__dict__.update(kw)


def ga_run(seqs, mode, sd, knobs=None, w=DEFAULT_W, rng=_pyrandom, qdtype=np.float32, log=None, fast_pareto=False):
    """GA.py:471-545.  Returns (alignment strings, hof score tuple)."""
    kb = knobs or Knobs()
    P = kb.POPULATION_SIZE
    pop = generate_population(seqs, P, kb.CLONE_RATE, kb.GAP_RATE, rng)
    hof = None

    def evaluate():
        nonlocal hof
        sc = fitness_scores(pop, mode, w)
        hof = hof_update(hof, pop, sc)
        if log is not None:
            log.append(("fitness", [tuple(s) for s in sc], decode(hof[0]), tuple(hof[1])))
        return sc

    scores = evaluate()
    if kb.AGENT_WINDOW_ROW > len(pop[0]):
        raise ValueError("AGENT_WINDOW_ROW is larger than the sequence count.")       # GA.py:483-484
    if kb.AGENT_WINDOW_COLUMN > min(len(r) for r in pop[0]):
        raise ValueError("AGENT_WINDOW_COLUMN is larger than the shortest sequence length.")  # GA.py:486-488
    for _ in range(kb.GA_ITERATIONS):
        mutation(pop, scores, mode, round(P * kb.MUTATION_RATE), kb.AGENT_WINDOW_ROW,
                 kb.AGENT_WINDOW_COLUMN, sd, w, qdtype, kb.MUTATION_TILE, kb.MUTATION_PROBABILITY, rng)
        scores = evaluate()
        if kb.SELECTION_SCHEME == "tournament":
            keep = tournament_keep(scores, mode, P, kb.SELECTION_RATE, kb.TOURNAMENT_RATE, rng)
        else:
            keep = selection_keep(scores, mode, P, kb.SELECTION_RATE, fast=fast_pareto)
        pop[:] = [b for i, b in enumerate(pop) if i in keep]
        scores = evaluate()
        if kb.CROSSOVER_PROBABILITY < 1.0:
            plan = crossover_plan_variant(len(pop), len(pop[0]), P - len(pop), kb.CROSSOVER_PROBABILITY, len(pop[0]), rng)
        else:
            plan = crossover_plan(len(pop), len(pop[0]), P - len(pop), rng)
        pop.extend(horizontal_child(pop[a], pop[b], c) for a, b, c in plan)
        scores = evaluate()
        if kb.ELITISM_PROBABILITY > 0.0 and rng.random() < kb.ELITISM_PROBABILITY:
            # elitism (PARITY UNPINNED, supplement table A3 e_prob): the hall of fame replaces the worst-ranked individual
            import copy
            pop[rank_best(scores, len(scores))[-1]] = copy.deepcopy(hof[0])
            scores = evaluate()
    return decode(hof[0]), tuple(hof[1])


# --------------------------------------------------------------------------- C library (population-scale scoring)
_lib = None


def build_c(force=False):
    """Compile oracle/ga_oracle.c -> oracle/_build/libga_oracle.so (gcc)."""
    out_dir = os.path.join(HERE, "_build")
    so = os.path.join(out_dir, "libga_oracle.so")
    src = os.path.join(HERE, "ga_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", so, src])
    return so


def clib():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_c())
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i32p = ctypes.POINTER(ctypes.c_int32)
        lib.orc_sum_of_pairs.restype = ctypes.c_long
        lib.orc_sum_of_pairs.argtypes = [u8p] + [ctypes.c_int] * 8
        lib.orc_exact_columns.restype = ctypes.c_int
        lib.orc_exact_columns.argtypes = [u8p] + [ctypes.c_int] * 5
        lib.orc_pad_clean.restype = ctypes.c_int
        lib.orc_pad_clean.argtypes = [u8p, ctypes.c_int, ctypes.c_int, i32p]
        lib.orc_fitness_population.restype = None
        lib.orc_fitness_population.argtypes = [u8p, i32p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               i32p, i32p, i32p, ctypes.c_int]
        lib.orc_tile_scores.restype = ctypes.c_int
        lib.orc_tile_scores.argtypes = [u8p] + [ctypes.c_int] * 8 + [i32p, i32p]
        lib.orc_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def _u8(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def _i32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def fitness_population_c(boards, rowlen, w=DEFAULT_W, nthreads=1):
    """boards uint8 [P,R,stride] and rowlen int32 [P,R] are rewritten in place
    (pad + clean); returns (sp, em, al) int32 arrays.  GA.py:156-178."""
    assert boards.dtype == np.uint8 and rowlen.dtype == np.int32
    assert boards.flags.c_contiguous and rowlen.flags.c_contiguous
    P, R, stride = boards.shape
    sp = np.empty(P, np.int32)
    em = np.empty(P, np.int32)
    al = np.empty(P, np.int32)
    clib().orc_fitness_population(_u8(boards), _i32(rowlen), P, R, stride, w.match, w.mismatch, w.gap,
                                  _i32(sp), _i32(em), _i32(al), nthreads)
    return sp, em, al


def tile_scores_c(board, n, rw, cw, w=DEFAULT_W):
    """utils.py:282-289 for one rectangular uint8 board [R,stride]."""
    R, stride = board.shape
    nt = (R // rw) * (n // cw)
    tsp = np.empty(max(nt, 1), np.int32)
    tem = np.empty(max(nt, 1), np.int32)
    got = clib().orc_tile_scores(_u8(board), R, stride, n, rw, cw, w.match, w.mismatch, w.gap, _i32(tsp), _i32(tem))
    assert got == nt
    return tsp[:nt], tem[:nt]


# --------------------------------------------------------------------------- array <-> list helpers
def to_arrays(population, stride=None):
    """list-of-boards -> (uint8 [P,R,stride] filled with 0 beyond rowlen, int32 [P,R])."""
    P, R = len(population), len(population[0])
    width = max(len(r) for b in population for r in b)
    stride = stride or max(16, -(-width // 16) * 16)
    boards = np.zeros((P, R, stride), np.uint8)
    rowlen = np.zeros((P, R), np.int32)
    for p, b in enumerate(population):
        for r, row in enumerate(b):
            boards[p, r, :len(row)] = row
            rowlen[p, r] = len(row)
    return boards, rowlen


def to_lists(boards, rowlen):
    return [[boards[p, r, :rowlen[p, r]].tolist() for r in range(boards.shape[1])]
            for p in range(boards.shape[0])]
