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
    ap.add_argument("--train-sets", type=int, default=1, help="sequence sets trained on in turn, weights carried over")
    ap.add_argument("--eval", nargs="+", default=["synthetic_dataset_3x30bp", "synthetic_dataset_6x30bp"])
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
    log, curve = [], []
    weights = os.path.join(config.DPAMSA_WEIGHTS_PATH, "trained_3x30.pth")
    for k in range(args.train_sets):
        # the reference's trainer looks for "<model>.pth.pth" when it starts a set (main.py:141-144), so its sets never
        # build on each other; to carry the weights over, the previous result is put where it looks
        if k > 0:
            import shutil
            shutil.copyfile(weights, weights + ".pth")
        hist = dmain.train(train_data, start=k, end=k + 1, model_path="trained_3x30", early_stopping_patience=10 ** 9)
        part = next(iter(hist.values()))
        log += part
        curve.append({"set": train_data.datasets[k], "first_100_mean_reward": sum(e["reward"] for e in part[:100]) / 100,
                      "last_100_mean_reward": sum(e["reward"] for e in part[-100:]) / 100, "last_sp": part[-1]["sp"]})
    train_s = time.perf_counter() - t0
    sd = O.synth_qnet_weights(3, 30, 7)
    torch.save({k: torch.from_numpy(v) for k, v in sd.items()}, os.path.join(config.DPAMSA_WEIGHTS_PATH, "random_3x30.pth"))
    out = {"recipe": {"trainer": "DPAMSA.main.train (reference recipe: eps-greedy 0.8 -> 0, Adam 1e-4, batch 128, replay 1000, "
                                 "target refresh every 128 updates)", "episodes_per_set": args.episodes,
                      "training_sets": ["synthetic_dataset_3x30bp." + n for n in train_data.datasets[:args.train_sets]],
                      "weights_carried_over_between_sets": args.train_sets > 1, "seed": args.seed,
                      "train_seconds": train_s, "steps": sum(e["steps"] for e in log)},
           "per_set": curve, "evaluation": {},
           "published_GA_DPAMSA_MO_means": {"synthetic_dataset_3x30bp": {"mean_sp": 5.6, "mean_cs": 0.3467},
                                            "synthetic_dataset_6x30bp": {"mean_sp": 12.8, "mean_cs": 0.2793},
                                            "source": "reference results/benchmarks/csv/inference/GA-DPAMSA/*_MO_GA_DPAMSA_results.csv "
                                                      "(pretrained weights that are not in the checkout; GA settings of the paper)"}}
    for name in args.eval:
        data = load_module(os.path.join(REF, "datasets", "inference_dataset", name + ".py"), "_eval_" + name)
        for arm, model in (("random_init", "random_3x30"), ("trained", "trained_3x30")):
            out["evaluation"]["%s/mo/%s" % (name, arm)] = evaluate(data, "mo", model, 11)
    out["note"] = ("evaluation = mainGA.inference, 'mo' mode, the reference's default knobs (P = 5, 3 generations, 3x30 window), "
                   "same `random` seed for both arms")
    print(json.dumps(out))


main()
