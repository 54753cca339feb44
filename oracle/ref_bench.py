"""Times the UNMODIFIED reference's own functions on a bench.py workload (CPU baseline / reference arm).

TEST / BASELINE INFRASTRUCTURE: executed only by bench.py (`--impl reference` and the `cpu_baseline`
leg), always in its own process because the reference's module names (GA, utils, config, DPAMSA.*)
collide with the product facade's. The reference comes from oracle/ref_shim.py: /root/reference in the
build container, the byte-identical copy oracle/_ref/ on the GPU box.

    python oracle/ref_bench.py --workload c3 --procs 16 --fitness-sample 8192 --mutation-sample 2

What is timed (per worker process, torch pinned to one thread - the reference is single-threaded):
  fitness   GA.calculate_fitness_score()  (GA.py:138-181: pad, clean, SP, CS, hall-of-fame update with
            its deep copy) on a GA object whose population is a slice of the synthetic workload;
  mutation  GA.mutation(model)            (GA.py:303-373: ranking, worst tile, Environment, DQN episode,
            splice) with config.POPULATION_SIZE / MUTATION_RATE set so that every individual of the
            slice is mutated; the agent forced into eval mode, timed twice:
            "as_shipped" (two Nets constructed and torch.load per individual, GA.py:354-355) and
            "hoisted" (the constructed and loaded agent re-used).
One JSON object on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)

_REF = None
_CTX = {}


def _ref():
    global _REF
    if _REF is None:
        import ref_shim
        _REF = ref_shim.load()
    return _REF


def _boards_to_lists(boards, rowlen):
    return [[boards[p, r, :rowlen[p, r]].tolist() for r in range(boards.shape[1])] for p in range(boards.shape[0])]


def _fitness_job(args):
    boards, rowlen, mode = args
    ref = _ref()
    ga = ref.GA.GA(["A"], mode)
    ga.population = _boards_to_lists(boards, rowlen)
    t0 = time.perf_counter()
    ga.calculate_fitness_score()
    return time.perf_counter() - t0, len(ga.population)


def _mutation_job(args):
    boards, rowlen, mode, rw, cw, model, hoist = args
    import torch
    torch.set_num_threads(1)
    ref = _ref()
    import ref_shim
    cfg = ref.config
    cfg.AGENT_WINDOW_ROW, cfg.AGENT_WINDOW_COLUMN = rw, cw
    n = boards.shape[0]
    cfg.POPULATION_SIZE, cfg.MUTATION_RATE = n, 1.0
    ga = ref.GA.GA(["A"], mode)
    ga.population = _boards_to_lists(boards, rowlen)
    ga.calculate_fitness_score()
    base = ref.dqn.DQN
    cache = {}

    class EvalDQN(base):                       # wrapper, not an edit of the reference: dropout off
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            ref_shim.force_eval(self)

    class HoistedDQN:                          # the agent constructed and loaded once (the artefact of GA.py:354-355 removed)
        def __new__(cls, *a, **k):
            if "agent" not in cache:
                cache["agent"] = EvalDQN(*a, **k)
                cache["loaded"] = False
                real_load = cache["agent"].load
                def load_once(name, *aa, **kk):
                    if not cache["loaded"]:
                        real_load(name, *aa, **kk)
                        cache["loaded"] = True
                cache["agent"].load = load_once
            return cache["agent"]

    ref.GA.DQN = HoistedDQN if hoist else EvalDQN
    steps = [0]
    real_predict = base.predict
    def counting_predict(self, state):
        steps[0] += 1
        return real_predict(self, state)
    base.predict = counting_predict
    try:
        t0 = time.perf_counter()
        ga.mutation(model)
        dt = time.perf_counter() - t0
    finally:
        ref.GA.DQN = base
        base.predict = real_predict
    return dt, n, steps[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--fitness-sample", type=int, default=0, help="boards scored in total (0 = skip)")
    ap.add_argument("--fitness-reps", type=int, default=1)
    ap.add_argument("--mutation-sample", type=int, default=0, help="individuals mutated per process and variant (0 = skip)")
    ap.add_argument("--one-core", action="store_true", help="also time the fitness sample on one process")
    args = ap.parse_args()

    import multiprocessing as mp
    import numpy as np
    import torch
    torch.set_num_threads(1)                       # before the fork: the reference is single-threaded
    ref = _ref()                                   # import once in the parent: workers are forked with it loaded;
    sys.path.insert(0, ROOT)                       # before bench.py puts the facade's directory on sys.path
    import bench                                   # synthetic inputs only (no facade import at module level)
    rows, cols, P, mode, desc = bench.WORKLOADS[args.workload]
    rw, cw = bench.AGENTS[args.workload]
    out = {"workload": args.workload, "procs": args.procs, "reference_root": os.path.basename(os.path.dirname(ref.GA.__file__)),
           "reference_files": "unmodified (oracle/ref_shim.py)", "mode": mode}
    n_src = max(4096, args.fitness_sample)
    boards, rowlen = bench.synth_population(rows, cols, min(P, n_src), 20251019)
    pool = mp.get_context("fork").Pool(args.procs) if args.procs > 1 else None
    run = (lambda fn, jobs: pool.map(fn, jobs)) if pool else (lambda fn, jobs: [fn(j) for j in jobs])

    if args.fitness_sample > 0:
        n = min(args.fitness_sample, boards.shape[0])
        idx = np.linspace(0, boards.shape[0] - 1, n).astype(np.int64)
        chunks = [c for c in np.array_split(idx, args.procs) if len(c)]
        jobs = [(boards[c].copy(), rowlen[c].copy(), mode) for c in chunks]
        run(_fitness_job, [(b[:8], r[:8], m) for b, r, m in jobs])          # warm the workers
        per_step = []
        for _ in range(args.fitness_reps):
            t0 = time.perf_counter()
            res = run(_fitness_job, jobs)
            per_step.append(time.perf_counter() - t0)
        wall = sum(per_step)
        out["fitness"] = {"evals_per_s": n * args.fitness_reps / wall, "boards_per_step": n, "steps": args.fitness_reps,
                          "seconds_per_step": per_step, "cores": args.procs,
                          "evals_per_s_per_core_busy": n / sum(r[0] for r in res),
                          "function": "GA.calculate_fitness_score (GA.py:138-181)"}
        if args.one_core:
            m = max(64, n // args.procs)
            dt, k = _fitness_job((boards[idx[:m]].copy(), rowlen[idx[:m]].copy(), mode))
            out["fitness_one_core"] = {"evals_per_s": k / dt, "boards": k, "seconds": dt}

    if args.mutation_sample > 0:
        import torch
        torch.manual_seed(7)
        model = "bench_ref_%dx%d" % (rw, cw)
        ref.write_weights(model, rw, cw, 7)
        k = args.mutation_sample
        rng = np.random.default_rng(3)
        # mutate non-clone individuals (distinct windows), spread over the population
        lo = round(boards.shape[0] * bench.CLONE_RATE)
        # warm the workers (first torch call in a process builds its thread pool and operator caches)
        run(_mutation_job, [(boards[lo:lo + 1].copy(), rowlen[lo:lo + 1].copy(), mode, rw, cw, model, True)] * args.procs)
        for name, hoist in (("hoisted", True), ("as_shipped", False)):
            jobs = []
            for w in range(args.procs):
                pick = rng.choice(np.arange(lo, boards.shape[0]), size=k, replace=False)
                jobs.append((boards[pick].copy(), rowlen[pick].copy(), mode, rw, cw, model, hoist))
            t0 = time.perf_counter()
            res = run(_mutation_job, jobs)
            wall = time.perf_counter() - t0
            n_ind = sum(r[1] for r in res)
            out["mutation_" + name] = {"seconds_per_individual_1core": sum(r[0] for r in res) / n_ind,
                                       "individuals": n_ind, "wall_seconds": wall, "cores": args.procs,
                                       "episode_steps_mean": sum(r[2] for r in res) / n_ind,
                                       "function": "GA.mutation (GA.py:303-373), eval mode forced"}
    if pool:
        pool.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
