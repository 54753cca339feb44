"""How many of a generation's mutation episodes start from a window that was already seen?
CPU probe with the oracle (c3 recipe scaled to --population): per generation the episodes
GA.py:331-373 would run, the distinct windows among them, and the windows never seen in any earlier
generation of the same run. Used to size gad_mutate's duplicate elimination."""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import ga_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--population", type=int, default=2048)
    ap.add_argument("--generations", type=int, default=7)
    ap.add_argument("--mode", default="mo")
    ap.add_argument("--rows", type=int, default=6)
    ap.add_argument("--cols", type=int, default=60)
    ap.add_argument("--window", type=int, nargs=2, default=(6, 30))
    ap.add_argument("--clone-rate", type=float, default=0.25)
    args = ap.parse_args()
    import torch
    import bench
    torch.set_num_threads(os.cpu_count() or 1)
    rw, cw = args.window
    sd = O.synth_qnet_weights(rw, cw, 7)
    sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
    forward = lambda states, mv: O.qnet_forward_torch(sd_t, states, mv)
    seqs = bench.seq_strings(args.rows, args.cols, 20251019)
    rng = random.Random(20251019)
    P = args.population
    pop = O.generate_population(seqs, P, args.clone_rate, 0.05, rng)
    scores = O.fitness_scores(pop, args.mode)
    seen = {}
    for g in range(args.generations):
        n = round(P * 0.25)
        in_gen = set()
        new = 0
        steps = 0
        for idx in O.rank_best(scores, n):
            b = pop[idx]
            _, (fr, tr, fc, tc) = O.worst_tile(b, args.mode, rw, cw)
            key = tuple(tuple(row[fc:tc]) for row in b[fr:tr])
            in_gen.add(key)
            if key not in seen:
                new += 1
                work = [list(r) for r in b]
                trace = []
                O.mutate_board(work, (fr, tr, fc, tc), sd, trace=trace, forward=forward)
                seen[key] = ([row[fc:fc + len(work[fr][fc:]) - len(b[fr][tc:])] for row in work[fr:tr]], len(trace))
            aligned, nsteps = seen[key]
            steps += nsteps
            for i, row in enumerate(aligned):
                b[fr + i][fc:tc] = row
        print("generation %d: episodes %d  distinct windows %d (%.2f%%)  never seen before %d (%.2f%%)  mean steps %.1f"
              % (g, n, len(in_gen), 100 * len(in_gen) / n, new, 100 * new / n, steps / n), flush=True)
        scores = O.fitness_scores(pop, args.mode)
        keep = O.selection_keep(scores, args.mode, P, 0.5, fast=True)
        pop[:] = [b for i, b in enumerate(pop) if i in keep]
        scores = O.fitness_scores(pop, args.mode)
        plan = O.crossover_plan(len(pop), len(pop[0]), P - len(pop), rng)
        pop.extend(O.horizontal_child(pop[a], pop[b], c) for a, b, c in plan)
        scores = O.fitness_scores(pop, args.mode)


if __name__ == "__main__":
    main()
