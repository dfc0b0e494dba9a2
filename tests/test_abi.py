"""CPU test: the C-ABI library loads and exports every symbol include/*.h declares
(no compute calls without a GPU), and the product fails loudly without CUDA."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_exports_match_header():
    from cvxpnpl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cvxpnpl_b200.h")).read()
    declared = set(re.findall(r"\b(cvxpnpl_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.cvxpnpl_b200_version()
    assert lib.cvxpnpl_b200_workspace_bytes(0) == 0
    assert lib.cvxpnpl_b200_workspace_bytes(100000) >= 0


def test_desc_layout_matches_header():
    """ctypes mirror vs the C struct: same field order and a plausible size."""
    import ctypes
    from cvxpnpl_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cvxpnpl_b200.h")).read()
    body = hdr[hdr.index("typedef struct cvxpnpl_b200_desc"):hdr.index("} cvxpnpl_b200_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b([A-Za-z_0-9]+);", body)
    assert names == [f[0] for f in _lib.Desc._fields_]
    # 1 int64 + 4 int32 + 5 ptr + eps + 4 int32 + 3 double + 7 ptr + ptr + size_t + 4 int32 + record ptr
    assert ctypes.sizeof(_lib.Desc) == 8 + 16 + 40 + 8 + 16 + 24 + 56 + 8 + 8 + 16 + 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import cvxpnpl_b200 as cb
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cb.pnp(np.zeros((4, 2)), np.zeros((4, 3)), np.eye(3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cvxpnpl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f == "synth.py" or "import oracle" not in src
                assert "from oracle" not in src and "import oracle" not in src
