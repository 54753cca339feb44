"""GA-DPAMSA generation loop with the reference's ``GA`` class surface (reference GA.py:38-545) on a
device-resident population.

Every method keeps the reference's name, arguments and observable result; the per-individual python
loops are replaced by one libgadpamsa launch per phase over all individuals (``population.py``).
``population`` / ``population_score`` / ``hall_of_fame`` are materialised from the device on read and
uploaded on assignment, so debug printing and callers that inspect them keep working.

Deliberate deviations:
  * the agent runs in eval mode (the reference leaves dropout on, which makes results random);
  * ``'sp'`` mode keeps an SP score of 0 as 0 (the reference's ``sp or cs`` turns it into None and the
    next sort raises TypeError, GA.py:178);
  * weights are loaded once per model instead of once per individual (GA.py:354-355);
  * crossover with fewer than 2 survivors / rows raises ValueError instead of looping forever
    (GA.py:281-290).
"""
from __future__ import annotations

import fractions
import functools
import random

import numpy as np
import torch

import config
import utils
import _native as nat
from DPAMSA import dqn as _dqn
from population import DevicePopulation, pack_boards, unpack_boards

# Map nucleotide characters to numerical values (gaps are represented by 5)  - reference GA.py:35
nucleotides_map = {'A': 1, 'T': 2, 'C': 3, 'G': 4, 'a': 1, 't': 2, 'c': 3, 'g': 4, '-': 5}


def _mt_state():
    """CPython's MT19937 state as a uint32[625] array (624 words + position)."""
    version, words, gauss = random.getstate()
    return np.array(words, dtype=np.uint32), (version, gauss)


def _mt_restore(arr, meta):
    random.setstate((meta[0], tuple(int(x) for x in arr), meta[1]))


class PlanJob:
    """The crossover plan of GA.py:281-293 drawn ahead of time on a worker thread (ctypes releases the GIL).

    The plan depends only on the number of survivors, the number of rows and the `random` stream - not on any
    score - so it can be drawn while the GPU runs the phases before the crossover. In 'sp' / 'cs' mode the survivor
    count is int(population_size * SELECTION_RATE) whatever the scores are, so the job starts with the generation;
    in 'mo' mode it starts when the selection has counted the survivors. Nothing else consumes `random` in
    between; the generator state is handed back in finish(). A job whose (parents, children) turn out wrong - a
    knob changed, a smaller population - or whose generator was touched meanwhile is discarded and the generator
    is left where it is."""

    def __init__(self, n_parents, n_rows, n_children):
        import threading
        self.key = (int(n_parents), int(n_rows), int(n_children))
        self.plan = np.empty((max(int(n_children), 0), 3), dtype=np.int32)
        self.state, self.meta = _mt_state()
        self.start_state = self.state.copy()
        self.error = None
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        try:
            nat.call("gad_mt_crossover_plan_host", self.state.ctypes.data, self.key[0], self.key[1], self.key[2], None,
                     self.plan.ctypes.data)
        except Exception as exc:              # re-raised by finish()
            self.error = exc

    @staticmethod
    def start(n_parents, n_rows, n_children):
        if n_children <= 0 or n_parents < 2 or n_rows < 2:
            return None                       # nothing to draw, or a case in which the inline path raises
        return PlanJob(n_parents, n_rows, n_children)

    def finish(self, n_parents, n_rows, n_children):
        """The plan, or None when the job does not fit (the caller then draws inline from the untouched generator)."""
        self.thread.join()
        if self.error is not None:
            raise self.error
        if self.key != (int(n_parents), int(n_rows), int(n_children)):
            return None
        if not np.array_equal(_mt_state()[0], self.start_state):
            return None                       # somebody reseeded or drew from `random` meanwhile: their stream wins
        _mt_restore(self.state, self.meta)
        return self.plan


def _variants_draw():
    """True when a paper-style variant (config.py) will draw from `random` inside mutation or selection: the crossover
    plan must then be drawn in its place in the generation, not ahead of time."""
    return (getattr(config, "MUTATION_TILE", "worst") != "worst" or float(getattr(config, "MUTATION_PROBABILITY", 1.0)) < 1.0
            or getattr(config, "SELECTION_SCHEME", "truncation") != "truncation"
            or float(getattr(config, "CROSSOVER_PROBABILITY", 1.0)) < 1.0)


def _on_device(fn):
    """Run a GA phase with the population's GPU current: kernels are launched on that device's current stream
    and the Q-net handle is created there, whatever device the caller (or another GA instance) selected last."""
    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        nat.require_device()                      # no GPU: fail with this build's own error, not torch's
        with torch.cuda.device(self._device):
            return fn(self, *args, **kwargs)
    return wrapped


class GA:
    def __init__(self, sequences, mode):
        self.sequences = sequences
        self.mode = mode
        self.population_size = config.POPULATION_SIZE
        self.current_iteration = 0
        self._dev = None                 # DevicePopulation
        self._device = torch.device(config.DEVICE)
        self._plan_job = None            # crossover plan being drawn ahead of time (PlanJob)

    # ------------------------------------------------------------------ device <-> python views
    def _weights(self):
        return int(config.MATCH_REWARD), int(config.MISMATCH_PENALTY), int(config.GAP_PENALTY)

    def _adopt(self, boards, rowlen):
        """Upload a host population (keeps the hall of fame of the previous device population)."""
        old = self._dev
        torch.cuda.set_device(self._device)
        cap = max(int(config.POPULATION_SIZE), self.population_size, boards.shape[0])
        extra = max(16, (int(config.AGENT_WINDOW_ROW) - 1) * int(config.AGENT_WINDOW_COLUMN))
        stride = -(-(boards.shape[2] + extra) // 16) * 16
        wide = np.full((boards.shape[0], boards.shape[1], stride), 5, dtype=np.uint8)
        wide[:, :, :boards.shape[2]] = boards
        self._dev = DevicePopulation.from_numpy(wide, rowlen, device=self._device, capacity=cap)
        self._keep_hof(old)

    def _keep_hof(self, old):
        """A new population keeps the hall of fame (and the counters) of the one it replaces (GA.py:71)."""
        if old is not None and old.hof_valid and old.R == self._dev.R:
            self._dev.ensure_stride(old.stride)
            old.ensure_stride(self._dev.stride)
            self._dev.hof_board.copy_(old.hof_board)
            self._dev.hof.copy_(old.hof)
            self._dev.hof_valid = True
            self._dev.evals, self._dev.forwards = old.evals, old.forwards
            self._dev.episodes, self._dev.unique_episodes, self._dev.states = old.episodes, old.unique_episodes, old.states

    @property
    def population(self):
        """Snapshot of the device population as the reference's list of boards of rows of ints.
        Assign a list to upload it; in-place edits of a returned snapshot are not seen by the device."""
        return [] if self._dev is None else self._dev.to_lists()

    @population.setter
    def population(self, boards):
        if not boards:                       # GA.py:71 `self.population = []`: the hall of fame is kept
            if self._dev is not None:
                self._dev.n = 0
            return
        self._adopt(*pack_boards(boards))

    @property
    def population_score(self):
        """(index, sp) / (index, cs) / (index, sp, cs) tuples like GA.py:175-178 (an SP of 0 stays 0)."""
        if self._dev is None:
            return []
        sp, em, al = self._dev.scores_numpy()
        out = []
        for i in range(self._dev.n):
            cs = int(em[i]) / int(al[i]) if al[i] > 0 else 0
            if self.mode == 'mo':
                out.append((i, int(sp[i]), cs))
            else:
                out.append((i, int(sp[i]) if self.mode == 'sp' else cs))
        return out

    @population_score.setter
    def population_score(self, scores):
        """Upload externally supplied score tuples (used by callers that drive single phases)."""
        if self._dev is None or len(scores) != self._dev.n:
            raise ValueError("population_score must have one tuple per individual of the current population")
        n = len(scores)
        sp = np.zeros(n, np.int32)
        em = np.zeros(n, np.int32)
        al = np.ones(n, np.int32)
        for i, t in enumerate(scores):
            if t[0] != i:
                raise ValueError("score tuple %d carries index %r" % (i, t[0]))
            cs = None
            if len(t) == 3:
                sp[i], cs = t[1], t[2]
            elif self.mode == 'sp':
                sp[i] = t[1]
            else:
                cs = t[1]
            if cs is not None:
                fr = fractions.Fraction(cs).limit_denominator(1 << 20)
                if fr.numerator / fr.denominator != cs:
                    raise ValueError("column score %r is not a ratio of small integers" % (cs,))
                em[i], al[i] = fr.numerator, fr.denominator
        b = self._dev.cur
        for name, arr in (("sp", sp), ("em", em), ("al", al)):
            getattr(b, name)[:n].copy_(torch.from_numpy(arr))

    def _score_tuple(self, idx, sp, em, al):
        cs = em / al if al > 0 else 0
        if self.mode == 'mo':
            return (idx, sp, cs)
        return (idx, sp if self.mode == 'sp' else cs)

    @property
    def hall_of_fame(self):
        if self._dev is None or not self._dev.hof_valid:
            return []
        board, (idx, sp, em, al) = self._dev.hof_numpy()
        return ([row.tolist() for row in board], self._score_tuple(idx, sp, em, al))

    @hall_of_fame.setter
    def hall_of_fame(self, value):
        if value:
            raise ValueError("the hall of fame lives on the device; only clearing it is supported")
        if self._dev is not None:
            self._dev.hof_valid = False

    # ------------------------------------------------------------------ a3
    def _generate_host(self, own_first=0, own_step=1, min_stride=0):
        """GA.py:71-94 through the CPython-compatible MT19937 replay (same draws as the reference):
        host arrays (boards uint8 [n,R,stride], rowlen int32 [n,R]) of the boards p with
        p % own_step == own_first (all of them by default)."""
        nat.require_device()
        seqs = [[nucleotides_map[nuc] for nuc in seq] for seq in self.sequences]      # KeyError like GA.py:79
        R = len(seqs)
        P = self.population_size
        n_clone = round(P * config.CLONE_RATE)
        ngaps = np.array([max(1, round(len(s) * config.GAP_RATE)) for s in seqs], dtype=np.int32)
        seqlen = np.array([len(s) for s in seqs], dtype=np.int32)
        if n_clone < P and int(seqlen.min()) < 1:
            raise ValueError("empty range in randrange(0, 0)")                        # random.randint(0, -1)
        cstride = max(1, int(seqlen.max()))
        seq_arr = np.full((R, cstride), 5, dtype=np.uint8)
        for r, s in enumerate(seqs):
            seq_arr[r, :len(s)] = s
        stride = max(-(-(cstride + int(ngaps.max())) // 16) * 16, -(-int(min_stride) // 16) * 16)
        n_own = len(range(own_first, P, own_step))
        boards = np.empty((n_own, R, stride), dtype=np.uint8)
        rowlen = np.empty((n_own, R), dtype=np.int32)
        state, meta = _mt_state()
        if own_step == 1:
            nat.call("gad_mt_population_host", state.ctypes.data, seq_arr.ctypes.data, seqlen.ctypes.data, cstride,
                     ngaps.ctypes.data, R, P, n_clone, boards.ctypes.data, rowlen.ctypes.data, stride)
        else:
            nat.call("gad_mt_population_shard_host", state.ctypes.data, seq_arr.ctypes.data, seqlen.ctypes.data, cstride,
                     ngaps.ctypes.data, R, P, n_clone, own_first, own_step, boards.ctypes.data, rowlen.ctypes.data,
                     stride)
        _mt_restore(state, meta)
        return boards, rowlen

    def _generate_device(self, own_first=0, own_step=1, min_stride=0):
        """GA.py:71-94 built ON the device (csrc/mt_device.cu: the same MT19937 draws, the same boards, no host
        copy and no upload). Returns device tensors (boards uint8 [n,R,stride], rowlen int32 [n,R]) of the boards p
        with p % own_step == own_first, or None for shapes the device generator does not cover / GAD_POPGEN=host."""
        import os
        nat.require_device()
        if os.environ.get("GAD_POPGEN", "device") == "host":
            return None
        seqs = [[nucleotides_map[nuc] for nuc in seq] for seq in self.sequences]      # KeyError like GA.py:79
        R = len(seqs)
        P = self.population_size
        n_clone = round(P * config.CLONE_RATE)
        ngaps = np.array([max(1, round(len(s) * config.GAP_RATE)) for s in seqs], dtype=np.int32)
        seqlen = np.array([len(s) for s in seqs], dtype=np.int32)
        if n_clone < P and int(seqlen.min()) < 1:
            raise ValueError("empty range in randrange(0, 0)")                        # random.randint(0, -1)
        if int(ngaps.max()) > 64:
            return None
        cstride = max(1, int(seqlen.max()))
        seq_arr = np.full((R, cstride), 5, dtype=np.uint8)
        for r, s in enumerate(seqs):
            seq_arr[r, :len(s)] = s
        stride = max(-(-(cstride + int(ngaps.max())) // 16) * 16, -(-int(min_stride) // 16) * 16)
        n_own = len(range(own_first, P, own_step))
        boards = torch.empty((n_own, R, stride), dtype=torch.uint8, device=self._device)
        rowlen = torch.empty((n_own, R), dtype=torch.int32, device=self._device)
        state, meta = _mt_state()
        try:
            nat.call("gad_population_device", state.ctypes.data, seq_arr.ctypes.data, seqlen.ctypes.data, cstride,
                     ngaps.ctypes.data, R, P, n_clone, own_first, own_step, nat.ptr(boards), nat.ptr(rowlen), stride,
                     nat.current_stream_ptr())
        except ValueError:
            return None                                    # GAD_EINVAL: a shape outside the device generator
        _mt_restore(state, meta)
        return boards, rowlen

    @_on_device
    def generate_population(self):
        """GA.py:71-94: n_clone exact copies, then gap-inserted individuals - generated on the device; the host
        replay (gad_mt_population_host + upload) covers the shapes the device generator does not."""
        extra = max(16, (int(config.AGENT_WINDOW_ROW) - 1) * int(config.AGENT_WINDOW_COLUMN))
        width = max((len(s) for s in self.sequences), default=0)
        width += max(1, round(width * config.GAP_RATE))
        made = self._generate_device(min_stride=width + extra)
        if made is None:
            self._adopt(*self._generate_host())
            return
        boards, rowlen = made
        old = self._dev
        cap = max(int(config.POPULATION_SIZE), self.population_size, boards.shape[0])
        self._dev = DevicePopulation(boards, rowlen, capacity=cap)
        self._dev.width_hint = width
        self._keep_hof(old)

    # ------------------------------------------------------------------ a4 + a8
    @_on_device
    def update_hall_of_fame(self):
        self._dev.hof_update(self.mode)

    @_on_device
    def calculate_fitness_score(self):
        """GA.py:154-181: pad, clean, score every board (one launch), then the hall of fame."""
        if self._dev is None or self._dev.n == 0:
            raise ValueError("empty population")
        self._dev.fitness_pass(*self._weights(), self.mode)        # scores + update_hall_of_fame, one launch

    # ------------------------------------------------------------------ a10 / a11
    @_on_device
    def _find_pareto_front(self):
        """GA.py:195-211 (computed in O(P log P) on the device, same set, index order)."""
        keep, _ = self._dev.select_flags('mo', -1)
        flags = keep.cpu().numpy()
        scores = self.population_score
        return [scores[i] for i in np.nonzero(flags)[0]]

    @_on_device
    def selection(self):
        """GA.py:232-260"""
        top_n = int(self.population_size * config.SELECTION_RATE)
        if getattr(config, "SELECTION_SCHEME", "truncation") == "tournament":
            self._tournament_selection(top_n)
            self.calculate_fitness_score()
            return
        n_keep = self._dev.select(self.mode, top_n)
        if (self._plan_job is None and getattr(config, "CROSSOVER_KIND", "horizontal") != "vertical"
                and not _variants_draw()):
            self._plan_job = PlanJob.start(n_keep, self._dev.R, int(config.POPULATION_SIZE) - n_keep)
        self.calculate_fitness_score()

    def _tournament_selection(self, top_n):
        """Extension (supplement table A3 'tourn'; no reference code): top_n tournaments without replacement. Each
        draws max(2, ceil(TOURNAMENT_RATE * pool)) contestants with random.sample from the individuals still in the
        pool; the best ranked one (the ordering of utils.py:332-351 on the device, ties to the lower position)
        survives and leaves the pool; survivors keep their original order (GA.py:259). Host loop over device
        ranks: meant for populations of the paper's size (N = 150)."""
        import math
        dev = self._dev
        n = dev.n
        top_n = min(top_n, n)
        order = dev.rank_order(self.mode, n).cpu().numpy()
        rank = np.empty(n, dtype=np.int64)
        rank[order] = np.arange(n)
        pool = list(range(n))
        keep = np.zeros(n, dtype=bool)
        rate = float(config.TOURNAMENT_RATE)
        for _ in range(top_n):
            t = min(len(pool), max(2, math.ceil(rate * len(pool))))
            winner = min(random.sample(pool, t), key=rank.__getitem__)
            keep[winner] = True
            pool.remove(winner)
        dev.compact_with(torch.from_numpy(keep).to(dev.device))

    # ------------------------------------------------------------------ a12 / a13
    @_on_device
    def _crossover(self, kind):
        dev = self._dev
        n_children = int(config.POPULATION_SIZE) - dev.n
        if n_children > 0:
            widths = None
            if kind == 1:
                widths = dev.cur.al[:dev.n].cpu().numpy().astype(np.int32)
            job, self._plan_job = self._plan_job, None
            plan = job.finish(dev.n, dev.R, n_children) if (job is not None and kind == 0) else None
            c_prob = float(getattr(config, "CROSSOVER_PROBABILITY", 1.0))
            if c_prob < 1.0:
                plan = self._plan_with_probability(dev.n, dev.R, n_children, c_prob, widths, kind)
            if plan is None:
                plan = np.empty((n_children, 3), dtype=np.int32)
                state, meta = _mt_state()
                nat.call("gad_mt_crossover_plan_host", state.ctypes.data, dev.n, dev.R, n_children,
                         None if widths is None else widths.ctypes.data, plan.ctypes.data)
                _mt_restore(state, meta)
            dev.crossover(plan, kind)
        else:
            self._plan_job = None
        self.calculate_fitness_score()

    @staticmethod
    def _plan_with_probability(n_parents, n_rows, n_children, probability, widths, kind):
        """Extension (supplement table A3 'c_prob'; no reference code): the reference's draws for every child
        (GA.py:281-293), then one random.random(); at or above `probability` the child is a copy of parent1 (a cut
        past the last row / cell). Drawn by the library's replay of CPython's generator
        (gad_mt_crossover_plan_prob_host), like the plain plan."""
        plan = np.empty((n_children, 3), dtype=np.int32)
        state, meta = _mt_state()
        nat.call("gad_mt_crossover_plan_prob_host", state.ctypes.data, n_parents, n_rows, n_children,
                 None if kind == 0 else widths.ctypes.data, float(probability), n_rows if kind == 0 else 1 << 20,
                 plan.ctypes.data)
        _mt_restore(state, meta)
        return plan

    def _elitism(self):
        """Extension (supplement table A3 'e_prob'; no reference code): with probability ELITISM_PROBABILITY per
        generation the hall of fame replaces the worst-ranked individual, and the population is scored again."""
        p = float(getattr(config, "ELITISM_PROBABILITY", 0.0))
        if p <= 0.0 or random.random() >= p:
            return
        dev = self._dev
        worst = int(dev.rank_order(self.mode, dev.n)[dev.n - 1].item())
        width = int(dev.hof[4].item())
        dev.cur.boards[worst].copy_(dev.hof_board)
        dev.cur.rowlen[worst].fill_(width)
        self.calculate_fitness_score()

    def horizontal_crossover(self):
        """GA.py:281-301"""
        self._crossover(0)

    def vertical_crossover(self):
        """Extension (supplement table A5; no reference code): every child row is
        parent1[:cut] + parent2[cut:], cut = randint(1, min(width1, width2) - 1)."""
        self._crossover(1)

    # ------------------------------------------------------------------ a14
    @_on_device
    def mutation(self, model_path):
        """GA.py:323-373 for all selected individuals at once."""
        dev = self._dev
        rw, cw = int(config.AGENT_WINDOW_ROW), int(config.AGENT_WINDOW_COLUMN)
        n = round(config.POPULATION_SIZE * config.MUTATION_RATE)
        n = min(n, dev.n)
        # 'sp' / 'cs': the selection will keep exactly top_n individuals, so the crossover plan can be drawn now,
        # under the whole mutation phase (PlanJob)
        self._plan_job = None
        if (self.mode in ("sp", "cs") and getattr(config, "CROSSOVER_KIND", "horizontal") != "vertical"
                and not _variants_draw()):
            top_n = int(self.population_size * config.SELECTION_RATE)
            if 2 <= top_n <= dev.n:
                self._plan_job = PlanJob.start(top_n, dev.R, int(config.POPULATION_SIZE) - top_n)
        if n <= 0:
            return
        if rw > dev.R:
            raise ValueError("AGENT_WINDOW_ROW is larger than the sequence count.")
        qnet = _dqn.load_qnet(model_path, rw, cw)
        idx = dev.rank_order(self.mode, n)
        m_prob = float(getattr(config, "MUTATION_PROBABILITY", 1.0))
        random_tile = getattr(config, "MUTATION_TILE", "worst") == "random"
        tile = None
        if m_prob < 1.0 or random_tile:
            idx, tile = self._mutation_variant(idx, rw, cw, m_prob, random_tile)
            if idx is None:
                return
        if tile is None:
            tile = dev.worst_tiles(idx, rw, cw, self.mode, *self._weights())
        if int(tile.min().item()) < 0:
            raise ValueError("min() arg is an empty sequence")          # no tile fits (utils.py:291)
        max_reward = rw * (rw - 1) / 2 * config.MATCH_REWARD            # env.py:79
        dev.mutate(qnet, idx, tile, cw * max_reward)                    # GA.py:354

    def _mutation_variant(self, idx, rw, cw, m_prob, random_tile):
        """Extensions (supplement tables A3 'm_prob' and A5 'random' mutation; no reference code). In ranking order
        every selected individual first draws random.random() (only when MUTATION_PROBABILITY < 1) and is skipped at
        or above the probability; with MUTATION_TILE = 'random' it then draws its tile with random.randrange over
        the lattice of utils.get_all_different_sub_range (row-major). Returns (device indices, device tiles or None
        for the worst-tile search); (None, None) when nobody mutates."""
        dev = self._dev
        order = idx.cpu().numpy()
        n = int(order.shape[0])
        ntiles = n_c = None
        if random_tile:
            minlen = dev.cur.rowlen[:dev.n].min(dim=1).values.cpu().numpy().astype(np.int64)[order]
            n_r = (dev.R - rw) // rw + 1 if dev.R >= rw else 0
            n_c = np.where(minlen >= cw, (minlen - cw) // cw + 1, 0)
            ntiles = np.ascontiguousarray(n_r * n_c, dtype=np.int32)
        keep = np.empty(n, dtype=np.uint8)
        tile = np.empty(n, dtype=np.int32)
        state, meta = _mt_state()                       # the library's replay of CPython's generator draws them
        nat.call("gad_mt_mutation_draws_host", state.ctypes.data, n, float(m_prob),
                 None if ntiles is None else ntiles.ctypes.data, keep.ctypes.data, tile.ctypes.data)
        _mt_restore(state, meta)
        sel = np.nonzero(keep)[0]
        if sel.size == 0:
            return None, None
        d_idx = torch.from_numpy(np.ascontiguousarray(order[sel], dtype=np.int32)).to(dev.device)
        d_tile = None
        if random_tile:
            t, nc = tile[sel].astype(np.int64), n_c[sel]
            coords = np.stack((t // nc * rw, t % nc * cw), axis=1).astype(np.int32)
            d_tile = torch.from_numpy(np.ascontiguousarray(coords)).to(dev.device)
        return d_idx, d_tile

    # ------------------------------------------------------------------ a22
    def print_hall_of_fame(self):
        if not self.hall_of_fame:
            print("Hall of Fame is empty! No best individual found yet.")
            return
        hof_individual, hof_score = self.hall_of_fame
        print("\nHALL OF FAME - BEST INDIVIDUAL SO FAR:")
        print("=" * 50)
        for sequence in utils.get_nucleotides_seqs(hof_individual):
            print("   %s" % sequence)
        if self.mode in {"sp", "cs"}:
            print("   Best %s Score: %s" % ("SP" if self.mode == "sp" else "CS", hof_score[1]))
        else:
            print("   SP Score: %s" % (hof_score[1],))
            print("   CS Score: %s" % (hof_score[2],))
        print("=" * 50)

    def print_population(self):
        print("\nCURRENT POPULATION (Iteration {}):".format(self.current_iteration))
        print("=" * 50)
        population = self.population
        if not population:
            print("Empty Population! Make sure your population was generated correctly.")
            return
        for i, (individual, scores) in enumerate(zip(population, self.population_score), start=1):
            print("Individual %d:" % i)
            for sequence in utils.get_nucleotides_seqs(individual):
                print("   %s" % sequence)
            if self.mode in {'sp', 'cs'}:
                print("   Fitness Score (%s): %s" % ("SP" if self.mode == "sp" else "CS", scores[1]))
            else:
                print("   SP Score: %s" % (scores[1],))
                print("   CS Score: %s" % (scores[2],))
            print("-" * 50)
        self.print_hall_of_fame()
        print("\nPopulation entirely printed.\n")

    # ------------------------------------------------------------------ a19
    def run(self, model_path, debug_mode=False):
        """GA.py:471-545"""
        self.generate_population()
        self.calculate_fitness_score()
        if debug_mode:
            self.print_population()
        dev = self._dev
        if config.AGENT_WINDOW_ROW > dev.R:
            raise ValueError("AGENT_WINDOW_ROW is larger than the sequence count.")
        sequence_min_length = int(dev.cur.rowlen[0].min().item())
        if config.AGENT_WINDOW_COLUMN > sequence_min_length:
            raise ValueError("AGENT_WINDOW_COLUMN is larger than the shortest sequence length.")
        for i in range(config.GA_ITERATIONS):
            if debug_mode:
                print("Iteration %d/%d..." % (i + 1, config.GA_ITERATIONS))
            self.current_iteration += 1
            self.mutation(model_path)
            self.calculate_fitness_score()
            if debug_mode:
                self.print_population()
            self.selection()
            if debug_mode:
                self.print_population()
            if getattr(config, "CROSSOVER_KIND", "horizontal") == "vertical":
                self.vertical_crossover()
            else:
                self.horizontal_crossover()
            self._elitism()
            if debug_mode:
                self.print_population()
        best_chromosome, _ = self.hall_of_fame
        return utils.get_nucleotides_seqs(best_chromosome)
