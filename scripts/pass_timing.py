"""Where the time of GA.calculate_fitness_score goes at the c3 shape: the bare fitness kernel, the one-launch pass
(decision in the kernel's tail) and the two-launch pass, per scoring mode, from one CUDA graph of K passes over fresh
replicas; then the %globaltimer stamps the tail leaves (gad_debug_stamps).    python scripts/pass_timing.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200")]
import _native as nat                      # noqa: E402
from population import DevicePopulation    # noqa: E402


def synthetic(R, C, P, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(1, 5, size=(R, C), dtype=np.uint8)
    gaps = max(1, round(C * 0.05))
    stride = -(-(C + gaps) // 16) * 16
    boards = np.full((P, R, stride), 5, np.uint8)
    rowlen = np.full((P, R), C + gaps, np.int32)
    for p0 in range(0, P, 8192):
        n = min(8192, P - p0)
        keep = np.argsort(rng.random((n, R, C + gaps)), axis=2)[:, :, :C]
        keep.sort(axis=2)
        np.put_along_axis(boards[p0:p0 + n, :, :C + gaps], keep, np.broadcast_to(base, (n, R, C)), axis=2)
    return boards, rowlen


def timed(fn_list, side):
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for f in fn_list:
            f()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def main():
    nat.require_device()
    torch.cuda.set_device(0)
    R, C, P, K, NREP = 6, 60, 65536, 20, 8
    boards, rowlen = synthetic(R, C, P, 1)
    reps = [DevicePopulation.from_numpy(boards.copy(), rowlen.copy()) for _ in range(NREP)]
    for r in reps:
        r._workspace()
    side = torch.cuda.Stream()
    stamps = torch.zeros(16, dtype=torch.int64, device="cuda")
    for mode in ("sp", "mo"):
        for r in reps:
            r.hof_valid = False
            r.hof.zero_()
        with torch.cuda.stream(side):
            for r in reps:
                r.fitness_pass(4, -4, -4, mode)
                r.fitness(4, -4, -4)
                r.hof_update(mode)
        torch.cuda.synchronize()
        t_k = timed([(lambda r=reps[i % NREP]: r.fitness(4, -4, -4)) for i in range(K)], side)
        t_1 = timed([(lambda r=reps[i % NREP]: r.fitness_pass(4, -4, -4, mode)) for i in range(K)], side)
        t_2 = timed([(lambda r=reps[i % NREP]: (r.fitness(4, -4, -4), r.hof_update(mode))) for i in range(K)], side)
        t_h = timed([(lambda r=reps[i % NREP]: r.hof_update(mode)) for i in range(K)], side)
        print("mode %s: kernel %.2f us, one-launch pass %.2f us, two-launch pass %.2f us, decision kernel alone %.2f us"
              % (mode, 1e3 * t_k / K, 1e3 * t_1 / K, 1e3 * t_2 / K, 1e3 * t_h / K))
        nat.call("gad_debug_stamps", stamps.data_ptr())
        with torch.cuda.stream(side):
            for i in range(3):
                reps[i].fitness_pass(4, -4, -4, mode)
        torch.cuda.synchronize()
        nat.call("gad_debug_stamps", 0)
        s = stamps.cpu().tolist()
        names = ["kernel entry", "tail entry", "before arrival", "barrier passed", "fenced", "ranges folded", "CTA best known",
                 "ticket taken", "last CTA: entry", "last: fenced", "last: winner known", "last: board copied"]
        t0 = s[0]
        print("  stamps (us after kernel entry of CTA 0): " + ", ".join("%s %.2f" % (n, (v - t0) / 1e3) for n, v in zip(names, s) if v))
        stamps.zero_()


main()
