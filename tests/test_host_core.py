"""CPU tests of the per-problem device routines compiled for the host
(tests/host/host_harness.cpp): same source as the CUDA kernels, so algorithmic
bugs show up here without a GPU.  The GPU parity tests proper are in
tests/test_gpu_parity.py."""
import warnings

import numpy as np
import pytest

from cvxpnpl_b200 import synth
from oracle import cvxpnpl_oracle as orc
from oracle import kkt
from tests.host import harness


@pytest.mark.parametrize("n_pts,n_lines", [(8, 4), (8, 0), (0, 6)])
def test_host_core_vs_oracle(n_pts, n_lines):
    B = 6
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=21)
    r = harness.solve(d)
    assert (r["status"] & 0xFF == 0).all() and (r["n_poses"] == 1).all()
    for i in range(B):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            C, N = orc._stack(d["pts_2d"][i] if n_pts else None, d["pts_3d"][i] if n_pts else None,
                              d["line_2d"][i] if n_lines else None, d["line_3d"][i] if n_lines else None, d["K"])
            A, Bm = orc.reduce_translation(C, N)
            poses, aux = orc.solve_relaxation(A, Bm, max_iters=200000, return_aux=True)
        Ro, to = poses[0]
        assert synth.rotation_angle(Ro, r["R"][i, 0]) < 1e-6
        assert np.linalg.norm(to - r["t"][i, 0]) / np.linalg.norm(to) < 1e-6
        cert = kkt.certificate(aux["Q"], r["Z"][i], aux["info"]["y"])
        assert cert["eq_res"] < 1e-7 and cert["psd_res"] < 1e-9 and cert["gap"] < 1e-7
        assert abs(r["obj"][i, 0] - r["obj"][i, 1]) < 1e-8


def test_host_extraction_degenerate(golden):
    g = golden["degenerate"]
    for name in ("pts4", "pts3", "lines3", "lines4", "p2l1", "coplanar8"):
        for i in range(len(g[name + "_Z"])):
            if not g[name + "_ok"][i]:
                continue
            n, R, t, st = harness.extract(g[name + "_Z"][i], g[name + "_AtA"][i], g[name + "_B"][i])
            ne = int(g[name + "_n"][i])
            assert n == ne
            got = np.concatenate([R[:n].reshape(n, 9), t[:n]], axis=1)
            exp = np.concatenate([g[name + "_R"][i, :n].reshape(n, 9), g[name + "_t"][i, :n]], axis=1)
            dist = np.sort(np.abs(got[:, None, :] - exp[None, :, :]).max(-1).min(1))
            assert dist[(n - 1) // 2] < 1e-6 and np.all(dist[: max(n - 1, 1)] < 1e-4), (name, i, dist)


def test_quartic_real_parts():
    """np.real(np.roots(.)) of cvxpnpl.py:185-186, complex pairs included."""
    rng = np.random.default_rng(0)
    for k in range(500):
        c = rng.standard_normal(5)
        if k % 3 == 0:
            c = np.poly(rng.standard_normal(4))[::-1] * rng.standard_normal()
        ref = np.sort(np.real(np.roots(c[::-1])))
        got = np.sort(harness.quartic(c))
        assert np.abs(got - ref).max() / max(1.0, np.abs(ref).max()) < 1e-8
