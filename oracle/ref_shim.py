"""Import shim for the UNMODIFIED reference (container-only test infrastructure).

TEST INFRASTRUCTURE - never imported by the product package. Only
``oracle/make_golden.py`` and the ``-m "not gpu"`` pin tests (which skip when
``/root/reference`` is absent, i.e. on the GPU box) use it.

Why a shim is needed:
  * ``DPAMSA/env.py:5-6`` imports tkinter, ``utils.py:2`` imports matplotlib -
    neither is installed here -> stub modules.
  * ``config.py:91-109`` creates result directories next to itself at import
    time -> the reference ``*.py`` files are first copied to a scratch
    directory under /tmp (never into this repository).
  * the reference's ``datasets/`` namespace package is shadowed by the
    HuggingFace ``datasets`` distribution -> data modules are loaded by path.
  * pretrained weights are missing from the checkout -> ``write_weights``
    creates seeded substitute weights with the reference's own ``DQN.save``.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import shutil
import sys
import tempfile
import types

def _default_root():
    """/root/reference in the build container; on the GPU box the byte-identical copy oracle/_ref/
    (oracle/make_ref.py) that travels with the repository."""
    if os.path.isfile("/root/reference/GA.py"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = os.environ.get("GADPAMSA_REFERENCE_ROOT") or _default_root()

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "GA.py"))


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


class RefModules:
    """Handle on the imported reference modules (all unmodified)."""

    def __init__(self, root, config, utils, GA, env, dqn, models):
        self.root = root
        self.config = config
        self.utils = utils
        self.GA = GA
        self.env = env
        self.dqn = dqn
        self.models = models

    def dataset(self, name: str):
        """Load ``datasets/inference_dataset/<name>.py`` by path."""
        path = os.path.join(REFERENCE_ROOT, "datasets", "inference_dataset", name + ".py")
        spec = importlib.util.spec_from_file_location("_refdata_" + name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    def write_weights(self, name: str, rows: int, cols: int, seed: int) -> str:
        """Seeded substitute weights in the reference's own state_dict layout
        (``DPAMSA/dqn.py:187-199``). Returns the file path."""
        import torch

        torch.manual_seed(seed)
        match = self.config.MATCH_REWARD
        max_reward = rows * (rows - 1) / 2 * match
        agent = self.dqn.DQN(2 ** rows - 1, rows, cols, cols * max_reward)
        path = os.path.join(self.config.DPAMSA_WEIGHTS_PATH, name + ".pth")
        torch.save(agent.eval_net.state_dict(), path)
        return path


def load() -> RefModules:
    """Copy the reference python files to a scratch dir and import them.

    The reference module names (config, utils, GA, DPAMSA.*) are imported
    under their own names from the scratch dir; callers that also import the
    product facade (same module names) must do so in a different process.
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    scratch = tempfile.mkdtemp(prefix="gadpamsa_ref_")
    for fn in ("GA.py", "utils.py", "config.py"):
        shutil.copy(os.path.join(REFERENCE_ROOT, fn), os.path.join(scratch, fn))
    os.makedirs(os.path.join(scratch, "DPAMSA"))
    for fn in os.listdir(os.path.join(REFERENCE_ROOT, "DPAMSA")):
        if fn.endswith(".py"):
            shutil.copy(os.path.join(REFERENCE_ROOT, "DPAMSA", fn), os.path.join(scratch, "DPAMSA", fn))

    tk = _stub("tkinter")
    tkf = _stub("tkinter.font")
    tk.font = tkf
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")

    for name in ("config", "utils", "GA", "DPAMSA", "DPAMSA.env", "DPAMSA.dqn", "DPAMSA.models",
                 "DPAMSA.replay_memory"):
        if name in sys.modules:
            raise RuntimeError("module %r already imported; load the reference in a fresh process" % name)
    sys.path.insert(0, scratch)
    try:
        config = importlib.import_module("config")
        env = importlib.import_module("DPAMSA.env")
        models = importlib.import_module("DPAMSA.models")
        dqn = importlib.import_module("DPAMSA.dqn")
        utils = importlib.import_module("utils")
        GA = importlib.import_module("GA")
    finally:
        sys.path.remove(scratch)
    _loaded = RefModules(scratch, config, utils, GA, env, dqn, models)
    return _loaded


def force_eval(agent) -> None:
    """Dropout off: the reference never calls .eval(), so parity
    for Q-values is only definable with the net forced into eval mode."""
    agent.eval_net.eval()
    agent.target_net.eval()
