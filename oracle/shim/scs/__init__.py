"""oracle/shim/scs -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A module *named* `scs` so that the reference's cvxpnpl.py (which does
`import scs` at cvxpnpl.py:7 and calls `scs.solve` at cvxpnpl.py:485-489) can be
imported and run VERBATIM in the build container, where real SCS is neither
installed nor installable.  Behind it sits oracle/scs_port.c, our restatement of
the published SCS algorithm -- NOT the SCS code.  Used only to generate the
golden fixtures under tests/golden/ (script: tests/golden/make_golden.py).

Interface honoured (what cvxpnpl.py touches): `__version__` (parsed at
cvxpnpl.py:12; >= 3.0.0 selects the `eps_abs` kwarg and cone key "z"),
`solve(data, cone, eps_abs=, max_iters=, verbose=)` returning
{"x": ndarray(55), "info": {"dobj": float, ...}}.

Deliberate difference from real SCS 3.x: eps_rel defaults to 0 here (SCS 3.x
defaults to 1e-4, which the reference does not override), i.e. the shim solves
to the *absolute* tolerance the reference asks for; max_iters is honoured.
Environment variable ORACLE_SCS_MAX_ITERS overrides max_iters (used when
generating goldens so that every fixture is fully converged).
"""
import os
import sys

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(_here, "..", "..", "..")))
from oracle import scs_port as _port  # noqa: E402

__version__ = "3.2.0"


def solve(data, cone, eps_abs=1e-9, eps_rel=0.0, max_iters=2500, verbose=False, eps=None, **kw):
    # `eps=` is the SCS 2.x spelling still used by benchmarks/toolkit/methods/rc.py:90-96
    if eps is not None:
        eps_abs = eps
    assert cone.get("z", cone.get("f")) in (22, 16) and list(cone["s"]) == [10]
    A = data["A"]
    A = A.toarray() if hasattr(A, "toarray") else np.asarray(A)
    mi = int(os.environ.get("ORACLE_SCS_MAX_ITERS", max_iters))
    res = _port.solve(A, np.asarray(data["b"], float), np.asarray(data["c"], float),
                      eps_abs=eps_abs, eps_rel=eps_rel, max_iters=mi)
    if verbose:
        print("[oracle scs shim]", res["info"])
    return res
