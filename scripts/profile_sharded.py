"""Per-step wall times of the sharded mutation phase (GAD_SHARDED_TIMING=1; every step synchronised).
torchrun --nproc-per-node N scripts/profile_sharded.py [c3|c4|c5]"""
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200"), ROOT]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_project_root")
os.environ["GAD_SHARDED_TIMING"] = "1"
import torch
import torch.distributed as dist


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    import config
    import sharded
    config.DEVICE = "cuda:%d" % local
    bench.WORKLOADS["c5"] = (64, 1000, 131072 * world, "mo", "c5 shape, 131072 individuals per rank")
    bench.AGENTS["c5"] = (3, 30)
    rows, cols, P, mode, _ = bench.WORKLOADS[wl]
    if wl == "c3":
        P *= world
    rw, cw = bench.AGENTS[wl]
    config.POPULATION_SIZE, config.AGENT_WINDOW_ROW, config.AGENT_WINDOW_COLUMN = P, rw, cw
    model, _ = bench.synth_weight_file(rw, cw, tag="_p%d" % rank)
    random.seed(20251019)
    g = sharded.ShardedGA(bench.seq_strings(rows, cols, 20251019), mode, sharded.DeviceBackend(config.DEVICE))
    g.generate_population()
    g.calculate_fitness_score()
    phases = {}
    def timed(name, fn):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        phases[name] = phases.get(name, 0.0) + time.perf_counter() - t0
    for it in range(5):
        if it == 2:
            phases.clear(); g.timing.clear()
        timed("mutation", lambda: g.mutation(model))
        timed("fitness", g.calculate_fitness_score)
        timed("selection", g.selection)
        timed("crossover", g.horizontal_crossover)
    for r in range(world):
        dist.barrier()
        if rank == r:
            print(json.dumps({"workload": wl, "world": world, "rank": rank,
                              "phases_ms": {k: round(1e3 * v / 3, 3) for k, v in phases.items()},
                              "mutation_steps_ms": {k: round(1e3 * v / 3, 3) for k, v in g.timing.items()}}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


main()
