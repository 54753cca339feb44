"""crossover_peer_kernel (the fused exchange + crossover of the sharded loop) in ONE process on two GPUs, so that it
can be timed with CUDA events and captured by ncu with NVLink counters (a torchrun job cannot be profiled that way).
GPU 0 plays "this rank": it builds children locally from parents that live on GPU 0 and on GPU 1; the rows of
remote parents cross NVLink inside the kernel.

  python scripts/peer_crossover_probe.py [--shape c5|c4|c3] [--remote 0.5]
  ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum -k regex:crossover_peer python scripts/...
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "ga-dpamsa_b200"), ROOT]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_project_root")
import numpy as np
import torch

SHAPES = {"c3": (6, 63, 224, 32768, 32768), "c4": (20, 210, 288, 16384, 16384), "c5": (64, 1050, 1120, 8192, 8192)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="c5")
    ap.add_argument("--remote", type=float, default=0.875, help="fraction of parents that live on the peer (7/8 on 8 GPUs)")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import _native as nat
    R, W, stride, S, C = SHAPES[args.shape]
    torch.cuda.set_device(0)
    nat.call("gad_enable_peer_access", 1)
    rng = np.random.default_rng(1)
    def parents(dev):
        b = torch.from_numpy(rng.integers(1, 6, size=(S, R, stride), dtype=np.uint8)).to(dev)
        r = torch.full((S, R), W, dtype=torch.int32, device=dev)
        return b, r
    b0, r0 = parents("cuda:0")
    b1, r1 = parents("cuda:1")
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    peer_b = torch.tensor([b0.data_ptr(), b1.data_ptr()], dtype=torch.int64, device="cuda:0")
    peer_r = torch.tensor([r0.data_ptr(), r1.data_ptr()], dtype=torch.int64, device="cuda:0")
    owner = (rng.random((C, 2)) < args.remote).astype(np.int64)
    slot = rng.integers(0, S, size=(C, 2))
    plan = np.empty((C, 3), np.int32)
    plan[:, :2] = owner * S + slot
    plan[:, 2] = rng.integers(1, R, size=C)
    d_plan = torch.from_numpy(plan).to("cuda:0")
    dst_b = torch.full((C, R, stride), 5, dtype=torch.uint8, device="cuda:0")
    dst_r = torch.zeros((C, R), dtype=torch.int32, device="cuda:0")
    def run():
        nat.call("gad_crossover_peer", peer_b.data_ptr(), peer_r.data_ptr(), S, dst_b.data_ptr(), dst_r.data_ptr(), 0,
                 d_plan.data_ptr(), C, R, stride, 0, torch.cuda.current_stream().cuda_stream)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    # check against a host restatement on a sample
    hb = [b0.cpu().numpy(), b1.cpu().numpy()]
    got = dst_b.cpu().numpy()
    for c in rng.choice(C, 64, replace=False):
        cut = plan[c, 2]
        for r in range(R):
            o, s_ = divmod(int(plan[c, 0 if r < cut else 1]), S)
            assert np.array_equal(got[c, r, :W], hb[o][s_, r, :W])
    rows_units = -(-W // 16) * 16
    remote_rows = int((owner[:, 0] * plan[:, 2] + owner[:, 1] * (R - plan[:, 2])).sum())
    remote_bytes = remote_rows * rows_units
    total_bytes = C * R * rows_units
    print(json.dumps({"shape": args.shape, "children": C, "rows": R, "row_bytes_read": rows_units, "ms": ms,
                      "remote_fraction": remote_bytes / total_bytes, "nvlink_read_GBps": remote_bytes / ms / 1e6,
                      "child_bytes_GBps": total_bytes / ms / 1e6}))


main()
