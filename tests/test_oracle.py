"""CPU tests: the oracle restatement (oracle/) against the golden vectors produced
by the verbatim reference (tests/golden/make_golden.py).  No GPU, no reference."""
import warnings

import numpy as np
import pytest

from cvxpnpl_b200 import synth
from oracle import cvxpnpl_oracle as orc
from oracle import kkt


def _sort_rows(a):
    a = a[~np.isnan(a).any(axis=1)]
    return a[np.lexsort(np.round(a, 6).T)]


def test_static_sdp_data(golden):
    u = golden["units"]
    assert np.array_equal(orc._A, u["A_sdp"])
    assert np.array_equal(orc._b, u["b_sdp"])
    assert np.allclose(orc.vech10(u["S"], 2), u["vech_S_2"], rtol=0, atol=0)
    assert np.allclose(orc.vech10(u["S"]), u["vech_S_1"], rtol=0, atol=0)
    assert np.array_equal(orc.vech10_inv(orc.vech10(u["S"])), u["S"])


def test_constraint_builders(golden):
    u = golden["units"]
    C, N = orc.point_constraints(u["p2"], u["p3"], u["K"])
    assert np.allclose(C, u["Cp"], rtol=0, atol=1e-15)
    assert np.allclose(N, u["Np"], rtol=0, atol=1e-15)
    C, N = orc.line_constraints(u["l2"], u["l3"], u["K"])
    assert np.allclose(C, u["Cl"], rtol=0, atol=1e-15)
    assert np.allclose(N, u["Nl"], rtol=0, atol=1e-15)


def test_e6q3(golden):
    u = golden["units"]
    for A, ref in zip(u["e6_in"], u["e6_out"]):
        a, b, c = orc._e6q3(A)
        got = _sort_rows(np.stack([a, b, c], axis=1))
        exp = _sort_rows(ref.T)
        assert np.allclose(got, exp, rtol=1e-8, atol=1e-9)


def test_constraint_ortho_det(golden):
    u = golden["units"]
    for V, rank, ref in zip(u["cod_V"], u["cod_rank"], u["cod_out"]):
        got = _sort_rows(orc.constraint_ortho_det(V, int(rank)))
        exp = _sort_rows(ref)
        assert got.shape == exp.shape
        assert np.allclose(got, exp, rtol=1e-7, atol=1e-8)


@pytest.mark.parametrize("name", ["pnp", "pnl", "pnpl"])
def test_examples_known_answer(golden, name):
    """examples/*.py are noise free: the estimate must equal the hard-coded ground
    truth (printed there with 8 decimals) and the reference's own output."""
    e = golden["examples"]
    kw = {k[len(name) + 1:]: e[k] for k in e.files if k.startswith(name + "_")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if name == "pnp":
            poses = orc.pnp(kw["pts_2d"], kw["pts_3d"], kw["K"], max_iters=50000)
        elif name == "pnl":
            poses = orc.pnl(kw["line_2d"], kw["line_3d"], kw["K"], max_iters=50000)
        else:
            poses = orc.pnpl(kw["pts_2d"], kw["line_2d"], kw["pts_3d"], kw["line_3d"], kw["K"],
                             max_iters=50000)
    assert len(poses) == int(kw["n"]) == 1
    R, t = poses[0]
    assert synth.rotation_angle(kw["R_gt"], R) < 2e-7
    assert np.linalg.norm(t - kw["t_gt"]) / np.linalg.norm(kw["t_gt"]) < 2e-7
    assert synth.rotation_angle(kw["R"][0], R) < 1e-6
    assert np.linalg.norm(t - kw["t"][0]) / np.linalg.norm(kw["t"][0]) < 1e-6


@pytest.mark.parametrize("name,n_pts,n_lines", [("pnp8", 8, 0), ("pnpl8_4", 8, 4), ("pnl6", 0, 6)])
def test_synth_against_reference(golden, name, n_pts, n_lines):
    g = golden["synth"]
    for noise in (0, 1, 2):
        key = f"{name}_s{noise}"
        for i in range(3):  # a subset keeps the CPU suite short
            C, N = orc._stack(g[key + "_pts_2d"][i] if n_pts else None,
                              g[key + "_pts_3d"][i] if n_pts else None,
                              g[key + "_line_2d"][i] if n_lines else None,
                              g[key + "_line_3d"][i] if n_lines else None, g[key + "_K"])
            A, B = orc.reduce_translation(C, N)
            assert np.allclose(A.T @ A, g[key + "_AtA"][i], rtol=0, atol=1e-13)
            assert np.allclose(B, g[key + "_B"][i], rtol=1e-12, atol=1e-13)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                poses, aux = orc.solve_relaxation(A, B, max_iters=100000, return_aux=True)
            assert len(poses) == int(g[key + "_n"][i])
            R, t = poses[0]
            assert synth.rotation_angle(g[key + "_R"][i, 0], R) < 1e-6
            tr = g[key + "_t"][i, 0]
            assert np.linalg.norm(t - tr) / np.linalg.norm(tr) < 1e-6
            # solver-independent certificate of the SDP optimum
            cert = kkt.certificate(aux["Q"], aux["Z"], aux["info"]["y"])
            assert cert["eq_res"] < 1e-7 and cert["psd_res"] < 1e-7
            assert cert["dual_psd_res"] < 1e-7 and cert["gap"] < 1e-7


@pytest.mark.parametrize("name", ["pts4", "pts3", "lines3", "lines4", "p2l1", "coplanar8"])
def test_degenerate_extraction(golden, name):
    """Replay the reference's multi-solution extraction on the stored Z."""
    g = golden["degenerate"]
    for i in range(len(g[name + "_Z"])):
        if not g[name + "_ok"][i]:
            with pytest.raises(np.linalg.LinAlgError):
                orc.extract(g[name + "_Z"][i], np.zeros((1, 9)), g[name + "_B"][i])
            continue
        poses = orc.extract(g[name + "_Z"][i], np.zeros((1, 9)), g[name + "_B"][i])
        n = int(g[name + "_n"][i])
        assert len(poses) == n
        got = np.array([np.concatenate([R.ravel(), t]) for R, t in poses])
        exp = np.concatenate([g[name + "_R"][i, :n].reshape(n, 9), g[name + "_t"][i, :n]], axis=1)
        # order-free comparison of the candidate sets.  The rank-4 recovery is ill
        # conditioned (normal equations + quartic roots; a near-double root turns
        # into a complex pair under 1e-16 perturbations), and this restatement
        # evaluates the quartic as det M(a) where the reference expands it, so one
        # candidate per problem is allowed to disagree.
        dist = np.sort(np.abs(got[:, None, :] - exp[None, :, :]).max(-1).min(1))
        assert dist[(n - 1) // 2] < 1e-6
        assert np.all(dist[: max(n - 1, 1)] < 1e-4)


@pytest.mark.parametrize("name", ["pts4", "pts3", "lines3", "lines4", "p2l1", "coplanar8", "pts5"])
def test_degenerate_extraction_big(name):
    """tests/golden/degenerate_big.npz (generator make_degenerate_big.py, verbatim reference): 40 problems
    per family.  The oracle raises LinAlgError exactly where the reference did, returns the same number
    of candidates, and matches every candidate the reference itself reproduces under a 1e-14
    perturbation of Z (oracle/candidate_sets.py) to 1e-6."""
    import os
    from oracle import candidate_sets as du
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "degenerate_big.npz"))
    total = stable = 0
    for i in range(len(g[name + "_err"])):
        if g[name + "_err"][i] == 1:
            with pytest.raises(np.linalg.LinAlgError), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                orc.extract(g[name + "_Z"][i], np.zeros((1, 9)), g[name + "_B"][i])
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            poses = orc.extract(g[name + "_Z"][i], np.zeros((1, 9)), g[name + "_B"][i])
        n = int(g[name + "_n"][i])
        assert len(poses) == n
        got = np.array([np.concatenate([R.ravel(), t]) for R, t in poses])
        exp = du.flat(g[name + "_R"][i], g[name + "_t"][i], n)
        m = du.stable_mask(exp, g[name + "_Rp"][i], g[name + "_tp"][i], g[name + "_np"][i])
        total, stable = total + n, stable + int(m.sum())
        assert du.compare(got, exp, m) < 1e-6, (name, i)
    assert stable >= 0.85 * total, (name, stable, total)


def test_rc_variant_against_reference(golden):
    """benchmarks/toolkit/methods/rc.py: static data and poses of the verbatim reference."""
    g = golden["rc"]
    assert np.array_equal(orc._A_rc, g["A_rc"]) and np.array_equal(orc._b_rc, g["b_rc"])
    for i in range(3):
        C, N = orc.point_constraints(g["pts_2d"][i], g["pts_3d"][i], g["K"])
        A, B = orc.reduce_translation(C, N)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            poses = orc.solve_relaxation(A, B, max_iters=200000, variant="rc")
        assert len(poses) == int(g["n"][i]) == 1
        assert synth.rotation_angle(g["R"][i, 0], poses[0][0]) < 1e-6
        assert np.linalg.norm(poses[0][1] - g["t"][i, 0]) / np.linalg.norm(g["t"][i, 0]) < 1e-6
