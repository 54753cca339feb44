"""The reference's own scripts, byte-identical copies under oracle/_ref (oracle/make_ref.py), run on the GPU
build through the drop-in launcher, checked against the oracle's whole GA.run on the same `random` stream.
Covers the north star's "mainGA.py and run_benchmarks.py drop in unchanged" (mainGA.py:115-153,
run_benchmarks.py:36-73)."""
import csv
import importlib.util
import os
import random
import stat
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import ga_oracle as O
from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu

LAUNCHER = os.path.join(PKG, "gadpamsa_dropin.py")
REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def ref_tree(native_lib):
    native_lib.require_device()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_ref
    if os.path.isfile(os.path.join(make_ref.SRC, "GA.py")):
        make_ref.make()
    if not make_ref.available():
        pytest.skip("oracle/_ref not built (python oracle/make_ref.py in the container)")
    return REF


def dataset_module(name):
    path = os.path.join(REF, "datasets", "inference_dataset", name + ".py")
    spec = importlib.util.spec_from_file_location("_refdata_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def project_root(tmp_path, models):
    """A scratch project root: substitute weights under DPAMSA/weights (the pretrained files are not in the
    reference checkout), the reference's datasets tree linked in."""
    import torch
    root = tmp_path / "root"
    (root / "DPAMSA" / "weights").mkdir(parents=True)
    os.symlink(os.path.join(REF, "datasets"), root / "datasets")
    sds = {}
    for name, (rows, cols, seed) in models.items():
        sd = O.synth_qnet_weights(rows, cols, seed)
        torch.save({k: torch.from_numpy(v) for k, v in sd.items()}, str(root / "DPAMSA" / "weights" / (name + ".pth")))
        sds[name] = sd
    return root, sds


def launch(script, root, seed, stdin=None, path_prefix=None, timeout=900):
    env = dict(os.environ, GADPAMSA_PROJECT_ROOT=str(root))
    env.pop("PYTHONPATH", None)
    if path_prefix:
        env["PATH"] = path_prefix + os.pathsep + env.get("PATH", "")
    return subprocess.run([sys.executable, LAUNCHER, "--seed", str(seed), os.path.join(REF, script)], cwd=str(root),
                          env=env, input=stdin, capture_output=True, text=True, timeout=timeout)


def parse_report(path):
    blocks = {}
    with open(path) as fh:
        for block in fh.read().strip().split("\n\n"):
            lines = block.strip().splitlines()
            if not lines:
                continue
            name = lines[0].split(": ", 1)[1]
            fields = dict(l.split(": ", 1) for l in lines[1:6])
            blocks[name] = (fields, lines[7:])
    return blocks


def oracle_runs(dataset, mode, sd, seed, knobs=None):
    """What the reference's mainGA.inference computes set after set on ONE `random` stream (mainGA.py:115-128)."""
    import torch
    sd_t = {k: torch.from_numpy(v) for k, v in sd.items()}
    rng = random.Random(seed)
    out = {}
    for name in dataset.datasets:
        seqs = getattr(dataset, name)
        best, score = O.ga_run(seqs, mode, sd, knobs or O.Knobs(), rng=rng, fast_pareto=True)
        out[name] = (best, score)
    return out


def test_reference_mainGA_unmodified(ref_tree, tmp_path):
    """`python mainGA.py` of the reference as shipped: dataset1_3x30bp, 'sp' mode, model 'new_model_3x30', the
    default knobs of config.py:39-46 (P=5, 3 iterations, 3x30 window)."""
    root, sds = project_root(tmp_path, {"new_model_3x30": (3, 30, 7)})
    r = launch("mainGA.py", root, seed=11)
    assert r.returncode == 0, r.stderr[-2000:]
    report = root / "results" / "reports" / "GA-DPAMSA" / "dataset1_3x30bp_Max_SP.txt"
    table = root / "results" / "benchmarks" / "csv" / "inference" / "GA-DPAMSA" / "dataset1_3x30bp_Max_SP_GA_DPAMSA_results.csv"
    assert report.exists() and table.exists()
    got = parse_report(str(report))
    data = dataset_module("dataset1_3x30bp")
    want = oracle_runs(data, "sp", sds["new_model_3x30"], 11)
    assert list(got) == list(data.datasets)
    for name, (best, score) in want.items():
        fields, alignment = got[name]
        assert alignment == best, name
        board = O.encode(best)
        n = len(board[0])
        assert int(fields["Sum of Pairs (SP)"]) == O.sum_of_pairs(board, 0, len(board), 0, n) == score[1], name
        assert int(fields["Alignment Length (AL)"]) == n
        assert int(fields["Exact Matches (EM)"]) == O.exact_columns(board, 0, len(board), 0, n)
        assert fields["Column Score (CS)"] == "%.3f" % O.column_score(board, 0, len(board), 0, n)
    with open(table) as fh:
        rows = list(csv.reader(fh))
    assert len(rows) == 1 + len(data.datasets) and rows[1][0] == data.datasets[0]


FAKE_TOOL = textwrap.dedent('''\
    #!%s
    """Stand-in for an external MSA tool: "aligns" by copying the input FASTA (equal-length sequences)."""
    import os, sys
    a = sys.argv[1:]
    name = os.path.basename(sys.argv[0])
    def after(flag):
        return a[a.index(flag) + 1]
    if name == "clustalo": src, dst = after("-i"), after("-o")
    elif name == "msaprobs": src, dst = a[0], after("-o")
    elif name == "clustalw":
        src = [x for x in a if x.startswith("-INFILE=")][0].split("=", 1)[1]
        dst = [x for x in a if x.startswith("-OUTFILE=")][0].split("=", 1)[1]
    elif name == "mafft": src, dst = a[-1], None
    elif name == "muscle5": src, dst = after("-align"), after("-output")
    elif name == "run_upp.py":
        src = after("-s"); os.makedirs(after("-d"), exist_ok=True); dst = os.path.join(after("-d"), "output_alignment.fasta")
    elif name == "run_pasta.py":
        src = after("-i"); os.makedirs(after("-o"), exist_ok=True)
        dst = os.path.join(after("-o"), "pastajob.marker001.%%s.aln" %% os.path.splitext(os.path.basename(src))[0])
    text = open(src).read()
    if dst is None: sys.stdout.write(text)
    else: open(dst, "w").write(text)
''')


def test_reference_run_benchmarks_unmodified(ref_tree, tmp_path):
    """`python run_benchmarks.py` of the reference as shipped, menu answer 2 (GA-DPAMSA vs the external tools;
    the tools are stand-ins on PATH that echo their input). GA-DPAMSA runs in 'sp' mode on
    encode_project_dataset_4x101bp with model 'model_3x30' (run_benchmarks.py:27-33,57) through the reference's
    own mainGA.inference, imported lazily by utils.run_ga_dpamsa_inference."""
    root, sds = project_root(tmp_path, {"model_3x30": (3, 30, 7)})
    bindir = tmp_path / "bin"
    bindir.mkdir()
    for tool in ("clustalo", "msaprobs", "clustalw", "mafft", "muscle5", "run_upp.py", "run_pasta.py"):
        f = bindir / tool
        f.write_text(FAKE_TOOL % sys.executable)
        f.chmod(f.stat().st_mode | stat.S_IEXEC)
    r = launch("run_benchmarks.py", root, seed=5, stdin="2\n", path_prefix=str(bindir))
    assert r.returncode == 0, r.stderr[-2000:]
    report = root / "results" / "reports" / "GA-DPAMSA" / "encode_project_dataset_4x101bp_Max_SP.txt"
    assert report.exists(), r.stdout[-2000:]
    got = parse_report(str(report))
    data = dataset_module("encode_project_dataset_4x101bp")
    want = oracle_runs(data, "sp", sds["model_3x30"], 5)
    for name, (best, score) in want.items():
        assert got[name][1] == best, name
        assert int(got[name][0]["Sum of Pairs (SP)"]) == score[1]
    # the tools' rows were scored on the device too (an unaligned copy of the input) and saved per tool
    for tool in ("ClustalOmega", "MSAProbs", "ClustalW", "MAFFT", "MUSCLE5", "UPP", "PASTA"):
        path = root / "results" / "benchmarks" / "csv" / tool / ("encode_project_dataset_4x101bp_%s_results.csv" % tool)
        assert path.exists(), tool
        with open(path) as fh:
            rows = list(csv.reader(fh))
        assert len(rows) == 1 + len(data.datasets), tool
