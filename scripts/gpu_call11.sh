cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_population_device_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_popgen.log
cat gpurun_out/r2_popgen.log
python - <<'PY' 2>&1 | tail -8
import os, sys, time, random
sys.path[:0] = ["ga-dpamsa_b200", "."]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_project_root")
import torch, bench, config
import GA as ga_mod
for name,(rows,cols,P) in {"c3":(6,60,65536),"c4":(20,200,262144),"c5_rank":(64,1000,1048576)}.items():
    config.POPULATION_SIZE = P
    seqs = bench.seq_strings(rows, cols, 20251019)
    g = ga_mod.GA(seqs, "sp")
    own = (0, 8) if name == "c5_rank" else (0, 1)
    for it in range(2):
        random.seed(1); torch.cuda.synchronize(); t0 = time.perf_counter()
        out = g._generate_device(*own); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, "device %.3f s" % dt, tuple(out[0].shape))
    if name != "c5_rank":
        random.seed(1); t0 = time.perf_counter(); h = g._generate_host(*own); print(name, "host %.3f s" % (time.perf_counter() - t0))
PY
