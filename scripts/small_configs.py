"""BASELINE configs c1 / c2 through the facade: wall time of GA.run (after one warm-up run)."""
import json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200"), ROOT]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_project_root")
import torch
import bench, config
import GA as ga_mod

data = json.load(open(os.path.join(ROOT, "tests", "golden", "datasets.json")))
model, _ = bench.synth_weight_file(3, 30, tag="_small")
out = {}
for tag, seqs, mode, P in (("c1: dataset1_6x30bp test0, sp, P=5", data["dataset1_6x30bp"]["test0"], "sp", 5),
                           ("c2: synthetic_dataset_3x30bp test0, mo, P=1024", data["synthetic_dataset_3x30bp"]["test0"], "mo", 1024),
                           ("c2: synthetic_dataset_6x60bp test0, mo, P=1024", data["synthetic_dataset_6x60bp"]["test0"], "mo", 1024)):
    config.POPULATION_SIZE, config.GA_ITERATIONS = P, 3
    config.AGENT_WINDOW_ROW, config.AGENT_WINDOW_COLUMN = 3, 30
    times = []
    for rep in range(4):
        random.seed(rep)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        best = ga_mod.GA(seqs, mode).run(model)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    out[tag] = {"GA.run seconds (3 iterations), runs 2-4": [round(t, 4) for t in times[1:]], "first run": round(times[0], 4)}
print(json.dumps(out, indent=1))
