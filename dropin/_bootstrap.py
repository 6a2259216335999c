"""Shared bootstrap of the drop-in shim: locate the reference checkout, make it importable AFTER
this directory, and paper over the two environment problems the reference has on a modern stack
(SURVEY 8c): `h5py` is imported by feature_utils.py:7 but only used by an unrelated loader, and
loss.py:134 uses the removed `np.bool`."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _reference_dir():
    """$DRB_REFERENCE_DIR, else a checkout at /root/reference, else the byte-for-byte copy oracle/make_ref.py puts
    in oracle/_ref (what exists on the GPU box)."""
    for cand in (os.environ.get("DRB_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "model_cl.py")):
            return cand
    return "/root/reference"


REF = _reference_dir()


def setup():
    if not hasattr(np, "bool"):
        np.bool = bool
    try:
        import h5py  # noqa: F401
    except ImportError:
        sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    for p in (ROOT, REF):
        if p not in sys.path:
            sys.path.append(p)


def load_reference_module(name):
    """Import /root/reference/<name>.py under the alias _ref_<name> (so that `import <name>` keeps
    resolving to the shim in this directory)."""
    import importlib.util

    setup()
    alias = "_ref_" + name
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod
