"""Recipe for ``oracle/_ref/``: a byte-for-byte copy of the reference's python files (and the small
inference data they read), taken from where they lie under /root/reference.

TEST / BASELINE INFRASTRUCTURE. ``oracle/_ref/`` is git-ignored (reference sources never enter this
repository's history) but not gpurun-ignored, so the copy travels to the GPU box, where
/root/reference does not exist. Consumers:
  * ``bench.py --impl reference`` and bench.py's ``cpu_baseline`` leg: time the reference's own
    ``GA.calculate_fitness_score`` / ``GA.mutation`` (through ``oracle/ref_shim.py``);
  * ``tests/test_dropin_gpu.py``: runs the reference's unmodified ``mainGA.py`` / ``run_benchmarks.py``
    on top of this build through ``ga-dpamsa_b200/gadpamsa_dropin.py``.
The reference is pure Python: nothing is compiled. ``python oracle/make_ref.py`` (also called by
``__graft_entry__.build()`` when /root/reference is present) is idempotent.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("GADPAMSA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")

# (directory relative to the reference root, file suffixes) - nothing else is copied
WHAT = [
    ("", (".py",)),
    ("DPAMSA", (".py",)),
    ("datasets/inference_dataset", (".py",)),
    ("datasets/fasta_files/encode_project_dataset_4x101bp", None),      # what run_benchmarks.py lists (run_benchmarks.py:50-51)
    ("datasets/training_dataset", (".py",)),                             # DPAMSA training sets (scripts/train_and_evaluate.py)
]


def make(verbose=False):
    """Returns the number of files copied (0 = up to date); raises when the reference is absent."""
    if not os.path.isfile(os.path.join(SRC, "GA.py")):
        raise RuntimeError("reference checkout not found at %s" % SRC)
    copied = 0
    for rel, suffixes in WHAT:
        src_dir = os.path.join(SRC, rel)
        dst_dir = os.path.join(DST, rel)
        os.makedirs(dst_dir, exist_ok=True)
        for fn in sorted(os.listdir(src_dir)):
            s = os.path.join(src_dir, fn)
            if not os.path.isfile(s) or (suffixes is not None and not fn.endswith(suffixes)):
                continue
            d = os.path.join(dst_dir, fn)
            if os.path.exists(d) and filecmp.cmp(s, d, shallow=False):
                continue
            shutil.copyfile(s, d)
            copied += 1
            if verbose:
                print("copied", os.path.join(rel, fn))
    return copied


def available():
    return os.path.isfile(os.path.join(DST, "GA.py"))


if __name__ == "__main__":
    n = make(verbose="-v" in sys.argv)
    print("oracle/_ref: %d file(s) copied from %s" % (n, SRC))
