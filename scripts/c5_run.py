"""BASELINE config c5: synthetic 64x1000 boards, a 1M-individual population sharded over 8 B200, a 3x30 agent
mutating 262 144 sub-boards per generation, per-generation NCCL all-gather of scores and survivor boards.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
      scripts/c5_run.py [--population 1048576] [--generations 2] [--rows 64] [--cols 1000] [--mode mo]

Prints one JSON line on rank 0 (times = max over ranks, barriers around every phase)."""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200"), ROOT]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_project_root")

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--population", type=int, default=1048576)
    ap.add_argument("--generations", type=int, default=2)
    ap.add_argument("--rows", type=int, default=64)
    ap.add_argument("--cols", type=int, default=1000)
    ap.add_argument("--mode", default="mo")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    import config
    import sharded
    config.DEVICE = "cuda:%d" % local
    config.POPULATION_SIZE, config.AGENT_WINDOW_ROW, config.AGENT_WINDOW_COLUMN = args.population, 3, 30
    model, _ = bench.synth_weight_file(3, 30, tag="_c5r%d" % rank)
    seqs = bench.seq_strings(args.rows, args.cols, 20251023)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    random.seed(20251023)
    g = sharded.ShardedGA(seqs, args.mode, sharded.DeviceBackend(config.DEVICE))
    t0 = time.perf_counter()
    g.generate_population()
    sync()
    t_pop = time.perf_counter() - t0
    t0 = time.perf_counter()
    g.calculate_fitness_score()
    sync()
    t_first = time.perf_counter() - t0
    gens = []
    for it in range(args.generations):
        rec = {}
        f0 = g.backend.pop.forwards
        c0 = g.comm_bytes
        for name, fn in (("mutation", lambda: g.mutation(model)), ("fitness", g.calculate_fitness_score),
                         ("selection", g.selection), ("crossover", g.horizontal_crossover)):
            sync()
            t0 = time.perf_counter()
            fn()
            sync()
            rec[name] = time.perf_counter() - t0
        t = torch.tensor([rec[k] for k in sorted(rec)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rec = {k: round(1e3 * float(v), 2) for k, v in zip(sorted(rec), t)}
        rec["total_ms"] = round(sum(rec.values()), 2)
        rec["forwards"] = g.backend.pop.forwards - f0
        rec["collective_GB_received_rank0"] = round((g.comm_bytes - c0) / 1e9, 3)
        rec["survivors_plus_children"] = g.n_global
        gens.append(rec)
    hof = g.hall_of_fame
    mem = torch.cuda.max_memory_allocated() / 1e9
    if rank == 0:
        last = gens[-1]
        print(json.dumps({
            "workload": "c5: synthetic %dx%d boards, population %d over %d GPU(s), mode %s, 3x30 agent"
                        % (args.rows, args.cols, args.population, world, args.mode),
            "generations_per_s": 1e3 / last["total_ms"], "individuals_per_s": args.population * 1e3 / last["total_ms"],
            "population_generation_s": round(t_pop, 2), "first_fitness_pass_ms": round(1e3 * t_first, 2),
            "per_generation": gens, "hall_of_fame_score": list(hof[1]), "max_memory_allocated_GB_rank0": round(mem, 2),
            "stride": g.backend.pop.stride}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
