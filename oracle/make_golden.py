"""Generate tests/golden/* from the UNMODIFIED reference (container-only).

Run:  python oracle/make_golden.py [--big]

Needs /root/reference (imported through oracle/ref_shim.py from a scratch copy,
dropout forced off - SURVEY.md 0.4).  Writes small fixtures that travel to the
GPU box, where the reference does not exist:

  report_vectors.json.gz  the 1340 published (alignment -> QTY, AL, SP, EM, CS)
                          records of results/reports/*/*.txt, each re-scored
                          here with the reference's utils.get_sum_of_pairs /
                          get_column_score (only records that agree are kept;
                          the script asserts that all do)
  datasets.json           the first sequence sets of the inference datasets
                          used by BASELINE configs c1/c2
  ga_traces.json          per-fitness-pass (scores, hall of fame) traces of the
                          reference's GA.run for seeds/modes/population sizes,
                          with numpy-seeded substitute weights
  primitives.json         known answers of the reference primitives on seeded
                          random inputs (ranking, Pareto, selection, tiles,
                          worst tile, crossover, population, env episodes)
  qnet_vectors.npz        states and the reference Net's eval-mode Q-values
"""
from __future__ import annotations

import copy
import glob
import gzip
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ga_oracle as O          # noqa: E402
import ref_shim                # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
WEIGHT_SEED = 7


def parse_reports(ref):
    recs = []
    root = os.path.join(ref_shim.REFERENCE_ROOT, "results", "reports")
    for path in sorted(glob.glob(os.path.join(root, "*", "*.txt"))):
        tool = os.path.basename(os.path.dirname(path))
        with open(path) as fh:
            blocks = fh.read().split("File: ")[1:]
        for blk in blocks:
            lines = blk.strip("\n").split("\n")
            head = {}
            for ln in lines[1:6]:
                k, v = ln.rsplit(":", 1)
                head[k.strip()] = v.strip()
            i = lines.index("Alignment:")
            rows = [ln.strip() for ln in lines[i + 1:] if ln.strip()]
            recs.append({
                "tool": tool, "file": os.path.basename(path), "name": lines[0].strip(),
                "qty": int(head["Number of Sequences (QTY)"]), "al": int(head["Alignment Length (AL)"]),
                "sp": int(float(head["Sum of Pairs (SP)"])), "em": int(head["Exact Matches (EM)"]),
                "cs3": head["Column Score (CS)"], "rows": rows,
            })
    return recs


def make_report_vectors(ref):
    recs = parse_reports(ref)
    keep = []
    for r in recs:
        board = O.encode([s.upper() for s in r["rows"]])
        R, W = len(board), len(board[0])
        assert all(len(x) == W for x in board)
        sp = ref.utils.get_sum_of_pairs(board, 0, R, 0, W)
        cs = ref.utils.get_column_score(board, 0, R, 0, W)
        em = round(cs * W)
        assert (R, W, sp, em, "%.3f" % cs) == (r["qty"], r["al"], r["sp"], r["em"], r["cs3"]), r["name"]
        keep.append({"tool": r["tool"], "file": r["file"], "name": r["name"], "rows": r["rows"],
                     "sp": sp, "em": em, "al": W})
    with gzip.open(os.path.join(OUT, "report_vectors.json.gz"), "wt") as fh:
        json.dump(keep, fh, separators=(",", ":"))
    print("report vectors:", len(keep))


DATASETS = {
    "dataset1_6x30bp": 6, "dataset1_3x30bp": 4, "synthetic_dataset_3x30bp": 6,
    "synthetic_dataset_6x60bp": 6, "dataset1_6x60bp": 3, "encode_project_dataset_4x101bp": 2,
}


def make_datasets(ref):
    out = {}
    for name, n in DATASETS.items():
        mod = ref.dataset(name)
        out[name] = {k: getattr(mod, k) for k in mod.datasets[:n]}
    with open(os.path.join(OUT, "datasets.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    return out


def install_weights(ref, rows, cols):
    import torch
    sd = O.synth_qnet_weights(rows, cols, WEIGHT_SEED)
    name = "syn%dx%d" % (rows, cols)
    torch.save({k: torch.from_numpy(v) for k, v in sd.items()},
               os.path.join(ref.config.DPAMSA_WEIGHTS_PATH, name + ".pth"))
    return name, sd


def ref_ga_trace(ref, seqs, mode, seed, knobs, model):
    for k, v in knobs.items():
        setattr(ref.config, k, v)
    random.seed(seed)
    ga = ref.GA.GA(seqs, mode)
    log = []
    inner = ga.calculate_fitness_score

    def logged():
        inner()
        log.append({"scores": [list(s) for s in ga.population_score],
                    "hof": O.decode(ga.hall_of_fame[0]), "hof_score": list(ga.hall_of_fame[1])})
    ga.calculate_fitness_score = logged
    out = ga.run(model)
    return {"mode": mode, "seed": seed, "knobs": knobs, "seqs": seqs, "model": model,
            "result": out, "passes": log}


def make_ga_traces(ref, data, big):
    orig = ref.dqn.DQN.__init__

    def eval_init(self, *a, **k):
        orig(self, *a, **k)
        ref_shim.force_eval(self)
    ref.dqn.DQN.__init__ = eval_init
    m3, _ = install_weights(ref, 3, 30)
    m6, _ = install_weights(ref, 6, 30)
    base = {"AGENT_WINDOW_ROW": 3, "AGENT_WINDOW_COLUMN": 30, "GA_ITERATIONS": 3, "POPULATION_SIZE": 5,
            "CLONE_RATE": 0.25, "GAP_RATE": 0.05, "SELECTION_RATE": 0.5, "MUTATION_RATE": 0.25}
    jobs = []
    d630 = data["dataset1_6x30bp"]
    for mode in ("sp", "cs", "mo"):
        for seed in (0, 1, 2):
            jobs.append((d630["test0"], mode, seed, dict(base), m3))                      # c1
    for mode in ("sp", "cs", "mo"):
        jobs.append((d630["test1"], mode, 3, dict(base, POPULATION_SIZE=24, GA_ITERATIONS=2), m3))
        jobs.append((data["synthetic_dataset_3x30bp"]["test0"], mode, 4,
                     dict(base, POPULATION_SIZE=32, GA_ITERATIONS=2), m3))
        jobs.append((data["synthetic_dataset_6x60bp"]["test0"], mode, 5,
                     dict(base, POPULATION_SIZE=16, GA_ITERATIONS=2), m3))
        jobs.append((data["synthetic_dataset_6x60bp"]["test1"], mode, 6,
                     dict(base, POPULATION_SIZE=12, GA_ITERATIONS=2, AGENT_WINDOW_ROW=6), m6))
    jobs.append((data["encode_project_dataset_4x101bp"]["test0"], "mo", 8,
                 dict(base, POPULATION_SIZE=10, GA_ITERATIONS=2, SELECTION_RATE=0.4, MUTATION_RATE=0.3,
                      CLONE_RATE=0.3, GAP_RATE=0.08), m3))
    if big:                                                                               # c2
        jobs.append((data["synthetic_dataset_3x30bp"]["test1"], "mo", 9,
                     dict(base, POPULATION_SIZE=1024, GA_ITERATIONS=1), m3))
    traces = []
    for seqs, mode, seed, knobs, model in jobs:
        try:
            traces.append(ref_ga_trace(ref, seqs, mode, seed, knobs, model))
            print("trace", mode, seed, knobs["POPULATION_SIZE"], traces[-1]["passes"][-1]["hof_score"])
        except TypeError as exc:       # GA.py:178 quirk: SP==0 -> None -> sort fails
            print("reference raised TypeError (sp-or-cs quirk):", mode, seed, exc)
    for k, v in base.items():
        setattr(ref.config, k, v)
    ref.dqn.DQN.__init__ = orig
    with gzip.open(os.path.join(OUT, "ga_traces.json.gz"), "wt") as fh:
        json.dump({"weight_seed": WEIGHT_SEED, "traces": traces}, fh, separators=(",", ":"))


def make_primitives(ref):
    rng = random.Random(20251019)
    out = {"rank": [], "pareto": [], "selection": [], "tiles": [], "worst": [], "clean": [],
           "crossover": [], "population": [], "env": [], "hof": []}
    for _ in range(40):
        n = rng.randint(1, 60)
        single = [(i, rng.randint(-6, 6)) for i in range(n)]
        multi = [(i, rng.randint(-6, 6) * 4, rng.randint(0, 5) / rng.randint(5, 7)) for i in range(n)]
        k = rng.randint(1, n)
        out["rank"].append({"scores": single, "n": k,
                            "out": ref.utils.get_index_of_the_best_fitted_individuals(single, k)})
        out["rank"].append({"scores": multi, "n": k,
                            "out": ref.utils.get_index_of_the_best_fitted_individuals(multi, k)})
        g = ref.GA.GA(["AC", "AC"], "mo")
        g.population_score = multi
        out["pareto"].append({"scores": multi, "out": [t[0] for t in g._find_pareto_front()]})
        for mode, sc in (("sp", single), ("cs", [(i, c) for i, _, c in multi]), ("mo", multi)):
            rate = rng.choice([0.25, 0.5, 0.75])
            g = ref.GA.GA(["AC", "AC"], mode)
            g.population_size = n
            g.population = [[[1, 2], [1, 2]] for _ in range(n)]
            g.population_score = sc
            ref.config.SELECTION_RATE = rate
            tagged = [[[i], [i]] for i in range(n)]
            g.population = tagged
            g.calculate_fitness_score = lambda: None
            g.selection()
            out["selection"].append({"mode": mode, "scores": sc, "rate": rate, "n": n,
                                     "keep": [b[0][0] for b in g.population]})
        ref.config.SELECTION_RATE = 0.5
        # HoF
        mode = rng.choice(["sp", "mo"])
        sc = single if mode == "sp" else multi
        g = ref.GA.GA(["AC", "AC"], mode)
        g.population = [[[i + 1]] for i in range(n)]
        g.population_score = sc
        g.update_hall_of_fame()
        first = [g.hall_of_fame[0], list(g.hall_of_fame[1])]
        sc2 = [(i, rng.randint(-6, 6)) for i in range(n)] if mode == "sp" else \
            [(i, rng.randint(-6, 6) * 4, rng.randint(0, 5) / rng.randint(5, 7)) for i in range(n)]
        g.population = [[[100 + i]] for i in range(n)]
        g.population_score = sc2
        g.update_hall_of_fame()
        out["hof"].append({"mode": mode, "scores1": sc, "scores2": sc2, "first": first,
                           "second": [g.hall_of_fame[0], list(g.hall_of_fame[1])]})
    for _ in range(60):
        R, W = rng.randint(2, 9), rng.randint(1, 70)
        board = [[rng.choice([1, 2, 3, 4, 5, 5, 5]) for _ in range(W)] for _ in range(R)]
        if rng.random() < 0.3:
            for c in rng.sample(range(W), min(W, 3)):
                for r in range(R):
                    board[r][c] = 5
        ragged = [row[:rng.randint(max(0, W - 6), W)] for row in board]
        b2 = copy.deepcopy(ragged)
        width = max(map(len, b2))
        for row in b2:
            row.extend([5] * (width - len(row)))
        ref.utils.clean_unnecessary_gaps(b2)
        out["clean"].append({"in": ragged, "out": b2})
        rw, cw = rng.randint(1, R), rng.randint(1, max(1, W // 2))
        out["tiles"].append({"R": R, "W": W, "rw": rw, "cw": cw,
                             "out": ref.utils.get_all_different_sub_range(board, rw, cw)})
        if W >= cw and R >= rw:
            ref.config.AGENT_WINDOW_ROW, ref.config.AGENT_WINDOW_COLUMN = rw, cw
            for mode in ("sp", "cs", "mo"):
                score, tile = ref.utils.calculate_worst_fitted_sub_board(board, mode)
                out["worst"].append({"board": board, "mode": mode, "rw": rw, "cw": cw,
                                     "score": score, "tile": list(tile)})
    ref.config.AGENT_WINDOW_ROW, ref.config.AGENT_WINDOW_COLUMN = 3, 30
    for seed in range(6):
        seqs = ["".join(rng.choice("ACGT-") if rng.random() < 0.9 else "-" for _ in range(rng.randint(20, 40)))
                for _ in range(rng.randint(2, 6))]
        P = rng.randint(4, 12)
        ref.config.POPULATION_SIZE = P
        random.seed(seed)
        g = ref.GA.GA(seqs, "sp")
        g.generate_population()
        pop = copy.deepcopy(g.population)
        survivors = g.population[:max(2, P // 2)]
        g.population = copy.deepcopy(survivors)
        g.calculate_fitness_score = lambda: None
        g.horizontal_crossover()
        out["population"].append({"seed": seed, "seqs": seqs, "P": P, "population": pop,
                                  "n_survivors": len(survivors), "after_crossover": g.population,
                                  "next_random": random.random()})
    ref.config.POPULATION_SIZE = 5
    for _ in range(12):
        R, C = rng.randint(2, 4), rng.randint(3, 8)
        data = [[rng.choice([1, 2, 3, 4, 5]) for _ in range(rng.randint(C - 1, C))] for _ in range(R)]
        env = ref.env.Environment(copy.deepcopy(data), convert_data=False)
        steps = [{"state": env.reset()}]
        for _ in range(R * C + 2):
            a = rng.randint(0, 2 ** R - 2)
            r, s, d = env.step(a)
            steps.append({"action": a, "reward": r, "state": s, "done": d,
                          "aligned": copy.deepcopy(env.aligned)})
            if d == 0 and r != -env.max_reward:
                break
        env.padding()
        out["env"].append({"data": data, "steps": steps, "padded": env.aligned,
                           "max_reward": env.max_reward, "action_number": env.action_number})
    with gzip.open(os.path.join(OUT, "primitives.json.gz"), "wt") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("primitives:", {k: len(v) for k, v in out.items()})


def make_qnet(ref):
    import torch
    rng = random.Random(99)
    pack = {}
    for rows, cols in ((3, 30), (6, 30), (2, 5)):
        sd = O.synth_qnet_weights(rows, cols, WEIGHT_SEED)
        max_value = cols * rows * (rows - 1) / 2 * 4
        agent = ref.dqn.DQN(2 ** rows - 1, rows, cols, max_value)
        agent.eval_net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        ref_shim.force_eval(agent)
        states = []
        for ep in range(3):
            env = O.Env([[rng.randint(1, 5) for _ in range(cols)] for _ in range(rows)])
            s = env.reset()
            while True:
                states.append(s)
                with torch.no_grad():
                    a = int(agent.predict(s)) if ep else rng.randint(0, 2 ** rows - 2)
                _, s, d = env.step(a)
                if d == 0 or len(states) > 60 * (ep + 1):
                    break
        states = np.asarray(states[::3], dtype=np.int8)
        with torch.no_grad():
            q = agent.eval_net(torch.from_numpy(states.astype(np.int64))).numpy()
        tag = "%dx%d" % (rows, cols)
        pack["states_" + tag] = states
        pack["q_" + tag] = q.astype(np.float32)
        print("qnet", tag, states.shape, float(np.abs(q).max()))
    np.savez_compressed(os.path.join(OUT, "qnet_vectors.npz"), weight_seed=WEIGHT_SEED, **pack)


def main():
    big = "--big" in sys.argv
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()
    make_report_vectors(ref)
    data = make_datasets(ref)
    make_primitives(ref)
    make_qnet(ref)
    make_ga_traces(ref, data, big)


if __name__ == "__main__":
    main()
