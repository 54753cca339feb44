"""Times one Q-network forward per encoder variant (GAD_QNET_ENCODER = tc default | fa) with bench.py's own
CUDA-event timer. Usage: python scripts/encoder_timing.py [rows cols states] ; prints ms per forward."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = """
import os, sys
sys.path[:0] = [%r, %r]
os.environ.setdefault("GADPAMSA_PROJECT_ROOT", "/tmp/gadpamsa_project_root")
import bench
r, c, n = %d, %d, %d
model, _ = bench.synth_weight_file(r, c, tag="_enc")
bench.qnet_forward_ms(model, r, c, n)
print(os.environ.get("GAD_QNET_ENCODER", "tc"), r, c, n, bench.qnet_forward_ms(model, r, c, n))
"""


def main():
    shapes = [(6, 30, 16384), (3, 30, 32768)]
    if len(sys.argv) == 4:
        shapes = [tuple(int(a) for a in sys.argv[1:4])]
    for r, c, n in shapes:
        for enc in ("fa", "tc"):
            env = dict(os.environ)
            env["GAD_QNET_ENCODER"] = enc
            code = CHILD % (os.path.join(ROOT, "ga-dpamsa_b200"), ROOT, r, c, n)
            out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
            print(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:])


if __name__ == "__main__":
    main()
