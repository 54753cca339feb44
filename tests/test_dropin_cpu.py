"""The drop-in launcher (ga-dpamsa_b200/gadpamsa_dropin.py) must make the reference's UNMODIFIED
scripts import this build's GA / utils / config / DPAMSA.* even though the reference's own files of
the same names sit next to the script (sys.path[0]); data modules and mainGA stay the script tree's.
No compute here (no GPU): only where names resolve."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import PKG, ROOT

LAUNCHER = os.path.join(PKG, "gadpamsa_dropin.py")


def run_launcher(args, cwd, env_extra=None):
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    env.pop("GADPAMSA_PROJECT_ROOT", None)              # conftest's default; the launcher's own default is under test
    env.update(env_extra or {})
    return subprocess.run([sys.executable, LAUNCHER] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=300)


def test_launcher_beats_the_script_directory(tmp_path):
    """A tree that looks like the reference checkout: decoy GA.py / utils.py / config.py / DPAMSA/env.py that
    would raise if imported, a namespace `datasets` package, and a script importing all of them."""
    tree = tmp_path / "GA-DPAMSA"
    (tree / "DPAMSA").mkdir(parents=True)
    (tree / "datasets" / "inference_dataset").mkdir(parents=True)
    for name in ("GA.py", "utils.py", "config.py", "DPAMSA/env.py", "DPAMSA/dqn.py"):
        (tree / name).write_text("raise ImportError('the reference CPU module %s was imported')\n" % name)
    (tree / "mainGA.py").write_text("WHO = 'script tree'\n")
    (tree / "datasets" / "inference_dataset" / "toy.py").write_text("file_name = 'toy.py'\ndatasets = ['t0']\nt0 = ['ACGT', 'ACGA']\n")
    (tree / "probe.py").write_text(textwrap.dedent("""
        import config
        from DPAMSA.env import Environment
        from DPAMSA.dqn import DQN
        from GA import GA
        import utils
        import mainGA
        import datasets.inference_dataset.toy as toy
        import GA as ga_mod, DPAMSA.env as env_mod
        print("GA", ga_mod.__file__)
        print("utils", utils.__file__)
        print("config", config.__file__)
        print("env", env_mod.__file__)
        print("mainGA", mainGA.WHO)
        print("toy", toy.t0, toy.__file__)
        print("root", config.PROJECT_ROOT)
        import sys
        print("argv", sys.argv[1:])
        print("main", __name__)
    """))
    r = run_launcher([str(tree / "probe.py"), "a", "b"], cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    out = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines())
    assert out["GA"] == os.path.join(PKG, "GA.py")
    assert out["utils"] == os.path.join(PKG, "utils.py")
    assert out["config"] == os.path.join(PKG, "config.py")
    assert out["env"] == os.path.join(PKG, "DPAMSA", "env.py")
    assert out["mainGA"] == "script tree"
    assert out["toy"].startswith("['ACGT', 'ACGA']") and str(tree) in out["toy"]
    assert out["root"] == str(tree)                      # weights / results live in the script's tree
    assert out["argv"] == "['a', 'b']" and out["main"] == "__main__"


def test_plain_pythonpath_does_not_drop_in(tmp_path):
    """The recipe INTEGRATION.md used to give (PYTHONPATH only) picks the script directory's modules:
    documents why the launcher exists."""
    tree = tmp_path / "ref"
    tree.mkdir()
    (tree / "GA.py").write_text("WHO = 'reference'\n")
    (tree / "probe.py").write_text("import GA\nprint(GA.WHO)\n")
    env = dict(os.environ, PYTHONPATH=PKG)
    r = subprocess.run([sys.executable, str(tree / "probe.py")], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "reference"


def _ref_tree():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_ref
    if os.path.isfile(os.path.join(make_ref.SRC, "GA.py")):
        make_ref.make()
    return make_ref.DST if make_ref.available() else None


@pytest.mark.parametrize("script", ["mainGA.py", "run_benchmarks.py"])
def test_reference_scripts_resolve_to_the_facade(script, tmp_path):
    """The reference's own mainGA.py / run_benchmarks.py (oracle/_ref copy, byte-identical): their imports
    (mainGA.py:5-10, run_benchmarks.py:4-6) land on this build, their dataset modules on the reference tree."""
    ref = _ref_tree()
    if ref is None:
        pytest.skip("no reference checkout and no oracle/_ref copy")
    r = run_launcher(["--check-imports", os.path.join(ref, script)], cwd=str(tmp_path),
                     env_extra={"GADPAMSA_PROJECT_ROOT": str(tmp_path)})
    assert r.returncode == 0, r.stderr
    out = dict(line.split(" -> ", 1) for line in r.stdout.strip().splitlines() if " -> " in line)
    assert out["utils"] == os.path.join(PKG, "utils.py")
    assert out["config"] == os.path.join(PKG, "config.py")
    if script == "mainGA.py":
        assert out["GA"] == os.path.join(PKG, "GA.py")
        assert out["DPAMSA.env"] == os.path.join(PKG, "DPAMSA", "env.py")
    assert out["datasets"] == str([os.path.join(ref, "datasets")])


def test_make_ref_is_a_faithful_copy():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import filecmp
    import make_ref
    if not os.path.isfile(os.path.join(make_ref.SRC, "GA.py")):
        pytest.skip("reference checkout not present")
    make_ref.make()
    for rel in ("GA.py", "utils.py", "config.py", "mainGA.py", "run_benchmarks.py", "DPAMSA/env.py", "DPAMSA/dqn.py",
                "DPAMSA/models.py", "datasets/inference_dataset/dataset1_6x30bp.py"):
        assert filecmp.cmp(os.path.join(make_ref.SRC, rel), os.path.join(make_ref.DST, rel), shallow=False), rel
    assert make_ref.make() == 0                          # idempotent
