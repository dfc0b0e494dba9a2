"""oracle/ref_loader.py -- TEST INFRASTRUCTURE (checker / CPU baseline), not product code.

The reference (SergioRAgostinho/cvxpnpl) is a single pure-Python module.  `stage()` -- called by
`__graft_entry__.build()` in the build container, where /root/reference is mounted -- places an
UNMODIFIED copy of it under the git-ignored oracle/_ref/ so that it travels to the GPU box with
the snapshot (the box has no /root/reference).  `load()` imports that copy with oracle/shim/scs
on the path: real SCS is neither installed nor installable (no network), so the `scs.solve` call
inside the reference (cvxpnpl.py:485-489) is answered by oracle/scs_port.c, our restatement of the
published SCS algorithm.  Everything else that runs -- constraint builders, vech, eigh, rank test,
multi-solution recovery, SVD projection, the Python glue -- is the reference's own code.

Nothing under oracle/_ref/ is tracked by git; no reference source is part of this repository.
"""
import importlib.util
import os
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REF_FILE = os.path.join(REF_DIR, "cvxpnpl.py")
SOURCE = "/root/reference/cvxpnpl.py"

_mod = None


def stage(force=False):
    """Copy the reference module into oracle/_ref/ (build container only).  Returns True if staged."""
    if not os.path.exists(SOURCE):
        return os.path.exists(REF_FILE)
    if force or not os.path.exists(REF_FILE) or os.path.getmtime(SOURCE) > os.path.getmtime(REF_FILE):
        os.makedirs(REF_DIR, exist_ok=True)
        shutil.copyfile(SOURCE, REF_FILE)
    return True


def available():
    return os.path.exists(REF_FILE)


def load():
    """The verbatim reference module (imported from oracle/_ref/, scs = oracle/shim/scs), or None."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        return None
    shim = os.path.join(_HERE, "shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    spec = importlib.util.spec_from_file_location("cvxpnpl_reference", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _mod = mod
    return mod
