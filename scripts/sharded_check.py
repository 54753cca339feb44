"""Sharded generation loop on N GPUs (torchrun) against the golden GA.run traces of the reference.
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           scripts/sharded_check.py
Every rank replays the recorded runs; the sharded result must equal the recorded per-pass scores, hall of
fame and final alignment for every world size. Prints SHARDED-OK on rank 0."""
import gzip
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200"), os.path.join(ROOT, "oracle")]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_project_root")

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import config
    config.DEVICE = "cuda:%d" % local
    import sharded
    import ga_oracle as O           # checker only: builds the substitute weights the traces were recorded with
    with gzip.open(os.path.join(ROOT, "tests", "golden", "ga_traces.json.gz"), "rt") as fh:
        traces = json.load(fh)
    os.makedirs(config.DPAMSA_WEIGHTS_PATH, exist_ok=True)
    names = {}
    for rows, cols in ((3, 30), (6, 30)):
        name = "syn%dx%d_r%d" % (rows, cols, rank)
        sd = O.synth_qnet_weights(rows, cols, traces["weight_seed"])
        torch.save({k: torch.from_numpy(v) for k, v in sd.items()}, os.path.join(config.DPAMSA_WEIGHTS_PATH, name + ".pth"))
        names[(rows, cols)] = name
    done = 0
    for tr in traces["traces"]:
        kb = tr["knobs"]
        if kb["POPULATION_SIZE"] < world * 2:
            continue
        for k, v in kb.items():
            setattr(config, k, v)
        random.seed(tr["seed"])
        g = sharded.ShardedGA(tr["seqs"], tr["mode"], sharded.DeviceBackend(config.DEVICE))
        log = []
        inner = g.calculate_fitness_score

        def logged():
            inner()
            hof = g.hall_of_fame
            log.append({"scores": [list(s) for s in g.population_score],
                        "hof": ["".join("ATCG-"[v - 1] for v in row) for row in hof[0]], "hof_score": list(hof[1])})
        g.calculate_fitness_score = logged
        out = g.run(names[(kb["AGENT_WINDOW_ROW"], kb["AGENT_WINDOW_COLUMN"])])
        assert len(log) == len(tr["passes"]), (len(log), len(tr["passes"]))
        for k, (mine, ref) in enumerate(zip(log, tr["passes"])):
            assert mine["scores"] == ref["scores"], (tr["mode"], tr["seed"], k)
            assert mine["hof"] == ref["hof"] and mine["hof_score"] == ref["hof_score"], (tr["mode"], tr["seed"], k)
        assert out == tr["result"]
        done += 1
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("SHARDED-OK world=%d traces=%d" % (world, done))


if __name__ == "__main__":
    main()
