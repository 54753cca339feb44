"""Run one of the reference's OWN scripts, unmodified, on top of this build.

    python /path/to/ga-dpamsa_b200/gadpamsa_dropin.py /path/to/GA-DPAMSA/mainGA.py
    python /path/to/ga-dpamsa_b200/gadpamsa_dropin.py [--seed N] /path/to/GA-DPAMSA/run_benchmarks.py

Why a launcher and not ``PYTHONPATH``: Python puts the script's directory at ``sys.path[0]``, ahead of
``PYTHONPATH``, so ``import GA`` / ``import utils`` / ``import config`` inside the reference's
``mainGA.py`` (``mainGA.py:5-10``, ``run_benchmarks.py:4-6``) would find the reference's own CPU
modules next to the script. The launcher installs a meta-path finder IN FRONT of the path-based one that
resolves exactly the hot-path module names

    GA  utils  config  DPAMSA  DPAMSA.*           (+ this build's private helpers)

to this directory, and leaves everything else to the normal import system with the script's directory at
``sys.path[0]`` - so ``mainGA`` (imported lazily by ``utils.run_ga_dpamsa_inference``, reference
``utils.py:773``) and the data modules ``datasets.inference_dataset.*`` come from the reference tree.
``datasets`` gets one more fix: the reference's ``datasets/`` has no ``__init__.py`` (a namespace
package) and is shadowed by any installed distribution of that name (HuggingFace ``datasets``); when
the script's directory has a ``datasets/`` folder, the finder binds the name to it.

``GADPAMSA_PROJECT_ROOT`` defaults to the script's directory, so ``DPAMSA/weights``, ``results/...``
and ``datasets/fasta_files`` are the reference tree's (reference ``config.py:61-109``).

``--seed N`` seeds ``random`` before the script starts (the reference never seeds; reproducible runs
need it). ``--check-imports`` executes the script's imports only (``__name__`` is not ``__main__``) and
prints where each hot-path module came from.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import runpy
import sys

PKG = os.path.dirname(os.path.abspath(__file__))

# module names of the reference's surface that this build replaces
FACADE_TOP = ("GA", "utils", "config", "DPAMSA")
# this build's own helpers, imported by the facade modules by bare name
PRIVATE_TOP = ("_native", "population", "sharded", "build_native")


class FacadeFinder(importlib.abc.MetaPathFinder):
    """Resolve the facade's module names to this directory, whatever sys.path says."""

    def __init__(self, script_dir):
        self.script_dir = script_dir

    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split(".")[0]
        if top in FACADE_TOP or top in PRIVATE_TOP:
            parts = fullname.split(".")
            base = os.path.join(PKG, *parts)
            if os.path.isdir(base):
                init = os.path.join(base, "__init__.py")
                if os.path.isfile(init):
                    return importlib.util.spec_from_file_location(fullname, init, submodule_search_locations=[base])
                spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)      # namespace package
                spec.submodule_search_locations = [base]
                return spec
            if os.path.isfile(base + ".py"):
                return importlib.util.spec_from_file_location(fullname, base + ".py")
            if top in FACADE_TOP:
                # a name of the replaced surface that this build does not have must not fall through to the
                # reference's CPU file of the same name
                raise ModuleNotFoundError("%s is not provided by the B200 build (%s)" % (fullname, PKG), name=fullname)
            return None
        if fullname == "datasets" and self.script_dir:
            d = os.path.join(self.script_dir, "datasets")
            if os.path.isdir(d) and not os.path.isfile(os.path.join(d, "__init__.py")):
                spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
                spec.submodule_search_locations = [d]
                return spec
        return None


def install(script_dir=None):
    """Put the finder first on sys.meta_path (idempotent). Returns it."""
    for f in sys.meta_path:
        if isinstance(f, FacadeFinder):
            f.script_dir = script_dir or f.script_dir
            return f
    finder = FacadeFinder(script_dir)
    sys.meta_path.insert(0, finder)
    return finder


def run(script, argv=(), seed=None, check_imports=False):
    script = os.path.abspath(script)
    if not os.path.isfile(script):
        raise FileNotFoundError(script)
    script_dir = os.path.dirname(script)
    os.environ.setdefault("GADPAMSA_PROJECT_ROOT", script_dir)
    for name in FACADE_TOP:
        if name in sys.modules:
            raise RuntimeError("module %r was imported before the launcher ran" % name)
    install(script_dir)
    if sys.path and os.path.abspath(sys.path[0] or os.getcwd()) == PKG:
        sys.path.pop(0)                      # `python gadpamsa_dropin.py` put this directory there
    sys.path.insert(0, script_dir)           # what `python script.py` would have done
    sys.argv = [script] + list(argv)
    if seed is not None:
        import random
        random.seed(seed)
    if check_imports:
        runpy.run_path(script, run_name="__dropin_check__")
        for name in ("GA", "utils", "config", "DPAMSA.env", "DPAMSA.dqn", "mainGA"):
            mod = sys.modules.get(name)
            print("%s -> %s" % (name, getattr(mod, "__file__", None) if mod else "(not imported)"))
        ds = sys.modules.get("datasets")
        print("datasets -> %s" % (list(ds.__path__) if ds is not None else "(not imported)"))
        return
    runpy.run_path(script, run_name="__main__")


def main(args=None):
    args = list(sys.argv[1:] if args is None else args)
    seed, check = None, False
    while args and args[0].startswith("--"):
        opt = args.pop(0)
        if opt == "--seed":
            seed = int(args.pop(0))
        elif opt.startswith("--seed="):
            seed = int(opt.split("=", 1)[1])
        elif opt == "--check-imports":
            check = True
        elif opt == "--":
            break
        else:
            raise SystemExit("unknown option %s\n%s" % (opt, __doc__))
    if not args:
        raise SystemExit(__doc__)
    run(args[0], args[1:], seed, check)


if __name__ == "__main__":
    main()
