"""Host-side profile of the sharded generation loop (cProfile on rank 0): where the wall time of the non-kernel
part goes. torchrun --nproc-per-node N scripts/profile_sharded.py"""
import cProfile
import os
import pstats
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200"), ROOT]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_project_root")
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    import config
    import sharded
    config.DEVICE = "cuda:%d" % local
    rows, cols, P, mode, _ = bench.WORKLOADS["c3"]
    config.POPULATION_SIZE, config.AGENT_WINDOW_ROW, config.AGENT_WINDOW_COLUMN = P * world, 6, 30
    model, _ = bench.synth_weight_file(6, 30, tag="_p%d" % rank)
    random.seed(20251019)
    g = sharded.ShardedGA(bench.seq_strings(rows, cols, 20251019), mode, sharded.DeviceBackend(config.DEVICE))
    g.generate_population()
    g.calculate_fitness_score()
    def gen():
        g.mutation(model)
        g.calculate_fitness_score()
        g.selection()
        g.horizontal_crossover()
    for _ in range(3):
        gen()
    torch.cuda.synchronize()
    dist.barrier()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(4):
        gen()
    torch.cuda.synchronize()
    pr.disable()
    if rank == 0:
        st = pstats.Stats(pr)
        st.sort_stats("cumulative").print_stats(45)
        st.sort_stats("tottime").print_stats(25)
    dist.barrier()
    dist.destroy_process_group()


main()
