"""A bounded training run of a 3x30 DPAMSA agent with this package's trainer (reference recipe, DPAMSA/main.py:95-235,
dqn.py:123-185) and its effect on GA-DPAMSA's alignment quality - SURVEY.md 8(f) rank 1.

    python scripts/train_and_evaluate.py [--episodes 1500] [--train-set 0] [--eval dataset1_6x30bp]

Trains on ONE sequence set of the reference's datasets/training_dataset/synthetic_dataset_3x30bp.py (the reference
never resumes between sets: it looks for <model>.pth.pth, main.py:141-144, so more sets would not accumulate), saves
DPAMSA/weights/<model>.pth in the reference's layout, then runs mainGA.inference ('sp' and 'mo', default knobs:
P = 5, 3 generations, 3x30 window) on the evaluation dataset with (a) seeded random-init weights and (b) the trained
weights, same `random` seeds, and prints mean SP / CS per arm next to the reference's published GA-DPAMSA means.
The pretrained weights of the paper are not in the reference checkout; a few minutes of training do
not reproduce them - the table shows direction, the recipe is the deliverable.
"""
import argparse
import csv
import importlib.util
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200"), os.path.join(ROOT, "oracle"), ROOT]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_train_root")
REF = os.path.join(ROOT, "oracle", "_ref")


def load_module(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def evaluate(dataset, mode, model, seed):
    import config
    import mainGA
    random.seed(seed)
    mainGA.inference(mode, dataset=dataset, model_path=model)
    tag = os.path.splitext(dataset.file_name)[0]
    mode_tag = {"sp": "Max_SP", "cs": "Max_CS", "mo": "MO"}[mode]
    path = os.path.join(config.GA_DPAMSA_INF_CSV_PATH, "%s_%s_GA_DPAMSA_results.csv" % (tag, mode_tag))
    with open(path) as fh:
        rows = list(csv.DictReader(fh))
    sp = [float(r["Sum of Pairs (SP)"]) for r in rows]
    cs = [float(r["Column Score (CS)"]) for r in rows]
    return {"sets": len(rows), "mean_sp": sum(sp) / len(sp), "mean_cs": sum(cs) / len(cs)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=1500)
    ap.add_argument("--train-set", type=int, default=0)
    ap.add_argument("--eval", default="dataset1_6x30bp")
    ap.add_argument("--seed", type=int, default=20251020)
    args = ap.parse_args()
    import math
    import numpy as np
    import torch
    import config
    import ga_oracle as O
    import DPAMSA.main as dmain

    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    config.MAX_EPISODE = args.episodes
    config.DECREMENT_ITERATION = math.ceil(config.MAX_EPISODE * 0.8 / (config.EPSILON // config.DELTA))
    os.makedirs(config.DPAMSA_WEIGHTS_PATH, exist_ok=True)
    train_data = load_module(os.path.join(REF, "datasets", "training_dataset", "synthetic_dataset_3x30bp.py"), "_train")
    t0 = time.perf_counter()
    hist = dmain.train(train_data, start=args.train_set, end=args.train_set + 1, model_path="trained_3x30",
                       early_stopping_patience=10 ** 9)
    train_s = time.perf_counter() - t0
    log = next(iter(hist.values()))
    curve = [sum(e["reward"] for e in log[i:i + 100]) / len(log[i:i + 100]) for i in range(0, len(log), max(100, len(log) // 10))]

    sd = O.synth_qnet_weights(3, 30, 7)
    torch.save({k: torch.from_numpy(v) for k, v in sd.items()}, os.path.join(config.DPAMSA_WEIGHTS_PATH, "random_3x30.pth"))
    data = load_module(os.path.join(REF, "datasets", "inference_dataset", args.eval + ".py"), "_eval")
    out = {"recipe": {"trainer": "DPAMSA.main.train (reference recipe: eps-greedy 0.8 -> 0, Adam 1e-4, batch 128, replay 1000, "
                                 "target refresh every 128 updates)", "episodes": args.episodes,
                      "training_set": "synthetic_dataset_3x30bp." + train_data.datasets[args.train_set], "seed": args.seed,
                      "train_seconds": train_s, "steps": sum(e["steps"] for e in log)},
           "reward_curve_100_episode_means": curve, "final_sp_on_training_set": log[-1]["sp"], "evaluation": {}}
    for mode in ("sp", "mo"):
        for arm, model in (("random_init", "random_3x30"), ("trained", "trained_3x30")):
            out["evaluation"]["%s/%s" % (mode, arm)] = evaluate(data, mode, model, 11)
    pub = os.path.join(REF, "..", "..", "tests", "golden", "report_vectors.json.gz")
    out["note"] = ("evaluation = mainGA.inference on %s with the reference's default knobs (P = 5, 3 generations, 3x30 window); "
                   "published GA-DPAMSA means (pretrained weights that are not in the checkout) are in the reference's "
                   "results/benchmarks/csv/inference/GA-DPAMSA/*.csv" % args.eval)
    print(json.dumps(out))


main()
