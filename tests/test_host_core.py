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


def test_host_plateau_jump_hard_problems():
    """tests/golden/hard_pnpl.npz: the slowest PnPL problems of three 1e5 batches (plateau walkers, 247-799
    iterations without the jump, generator tests/golden/make_hard.py).  With the plateau jump of pass_dr
    they take less than half the iterations and land on the same optimum as the oracle."""
    import os
    h = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "hard_pnpl.npz")))
    r = harness.solve(h)
    assert (r["status"] & 0xFF == 0).all() and (r["n_poses"] == 1).all()
    assert (r["iters"] <= 0.7 * h["iters_without_jump"]).all() and r["iters"].max() <= 320, r["iters"]
    assert r["iters"].sum() <= 0.5 * h["iters_without_jump"].sum()
    for i in range(len(r["iters"])):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Ro, to = orc.pnpl(h["pts_2d"][i], h["line_2d"][i], h["pts_3d"][i], h["line_3d"][i], h["K"],
                              max_iters=400000)[0]
        assert synth.rotation_angle(Ro, r["R"][i, 0]) < 1e-6
        assert np.linalg.norm(to - r["t"][i, 0]) / np.linalg.norm(to) < 1e-6


def test_plateau_detector_trust_region():
    """plateau_update (pnpl_solve.cuh): a jump after 5 flat residuals; tau doubles while the jump kicks the
    residual by less than 1.5x, stays above that, halves above 4x, and switches off when it shrinks to nothing."""
    def run_flat(st, n, r=1e-6):
        taus = []
        for _ in range(n):
            st, tau = harness.plateau_update(st, r, r)
            taus.append(tau)
        return st, taus
    st, taus = run_flat(0, 5)
    assert taus == [0, 0, 0, 0, 8]                      # first jump: 8 steps, after 5 flat iterations
    st, tau = harness.plateau_update(st, 1.2e-6, 1e-6)  # kick 1.1x (squared 1.2x): tau may double
    assert tau == 0
    st, taus = run_flat(st, 5, 1.2e-6)
    assert taus[-1] == 16 and sum(taus) == 16
    st, tau = harness.plateau_update(st, 4e-6, 1.2e-6)  # kick ~1.8x: tau stays
    st, taus = run_flat(st, 5, 4e-6)
    assert taus[-1] == 16
    st, tau = harness.plateau_update(st, 4e-4, 4e-6)    # kick 10x: tau halved
    st, taus = run_flat(st, 5, 4e-4)
    assert taus[-1] == 8
    for _ in range(4):                                  # every jump kicks hard: 8 -> 4 -> 2 -> 1 -> off
        st, _ = harness.plateau_update(st, 1.0, 1e-6)
        st, taus = run_flat(st, 5, 1.0)
    st, _ = harness.plateau_update(st, 1.0, 1e-6)
    st, taus = run_flat(st, 20, 1.0)
    assert sum(taus) == 0
    # a residual that keeps falling never jumps
    st, r = 0, 1.0
    for _ in range(50):
        st, tau = harness.plateau_update(st, r * 0.8, r)
        r *= 0.8
        assert tau == 0
    # the state never touches the two phase bits of LaneState::phase
    assert st & 3 == 0


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


@pytest.mark.parametrize("n_pts,n_lines", [(0, 6), (8, 4)])
def test_host_fp32_first_phase(n_pts, n_lines):
    """"fp32 ADMM + fp64 extraction" (BASELINE.json configs[3]): FP32 iterations into the
    tail, FP64 re-orthonormalisation, FP64 iterations to eps.  Same poses as the pure
    FP64 solve (1e-6 rad / 1e-6 relative is the north-star tolerance; the two agree far
    better), and the SDP optimum is still certified by the KKT conditions."""
    B = 40
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=33)
    a = harness.solve(d)
    m = harness.solve(d, fp32_iters=400)
    ok = (a["status"] & 0xFF == 0) & (m["status"] & 0xFF == 0)
    assert ok.mean() > 0.9 and (a["n_poses"][ok] == 1).all() and (m["n_poses"][ok] == 1).all()
    ang = synth.rotation_angle(a["R"][ok, 0], m["R"][ok, 0])
    terr = np.linalg.norm(a["t"][ok, 0] - m["t"][ok, 0], axis=1) / np.linalg.norm(a["t"][ok, 0], axis=1)
    assert ang.max() < 1e-7 and terr.max() < 1e-7
    # part of the iterations really ran in FP32, and the total is not inflated
    assert np.median(m["iters"][ok]) <= 1.15 * np.median(a["iters"][ok]) + 5
    # equality residual and PSD-ness of the returned Z
    for i in np.flatnonzero(ok)[:5]:
        Z = m["Z"][i]
        assert abs(Z[9, 9] - 1.0) < 1e-8 and np.linalg.eigvalsh(Z).min() > -1e-9
        assert abs(np.trace(Z[:9, :9]) - 3.0) < 1e-8


@pytest.mark.parametrize("name", ["pts4", "pts3", "lines3", "lines4", "p2l1", "coplanar8", "pts5"])
def test_host_extraction_degenerate_big(name):
    """The device extraction routines (host build) on tests/golden/degenerate_big.npz: ST_SINGULAR exactly
    where the verbatim reference raised LinAlgError (cvxpnpl.py:165 / 212 / 510), the reference's number of
    candidates, and every reproducible candidate (oracle/candidate_sets.py) to 1e-6."""
    import os
    from oracle import candidate_sets as du
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "degenerate_big.npz"))
    total = stable = 0
    for i in range(len(g[name + "_err"])):
        n, R, t, st = harness.extract(g[name + "_Z"][i], g[name + "_AtA"][i], g[name + "_B"][i])
        if g[name + "_err"][i] == 1:
            assert st & 0xFF == 3 and n == 0, (name, i, st, n)
            continue
        ne = int(g[name + "_n"][i])
        assert n == ne and st & 0xFF == 0, (name, i, n, ne, st)
        exp, got = du.flat(g[name + "_R"][i], g[name + "_t"][i], ne), du.flat(R, t, n)
        m = du.stable_mask(exp, g[name + "_Rp"][i], g[name + "_tp"][i], g[name + "_np"][i])
        total, stable = total + ne, stable + int(m.sum())
        assert du.compare(got, exp, m) < 1e-6, (name, i, du.compare(got, exp, m))
    assert stable >= 0.85 * total


def test_host_quartic_matches_np_roots():
    """quartic_real_parts (Ferrari start + Aberth-Ehrlich polish) against np.real(np.roots(..)) (cvxpnpl.py:185-186),
    including the badly scaled quartics the rank-4 branch produces (one root at 1e5, three near 1e-2) and
    complex pairs."""
    rng = np.random.default_rng(0)
    cases = [np.array([2.13399037e-30, 3.19277247e-27, -2.21505703e-25, 8.37712288e-26, 5.51503240e-31]),
             np.array([3.89663680e-18, -2.85182644e-14, -2.32927774e-11, -1.97020291e-09, 2.67766077e-14])]
    for k in range(600):
        if k % 3 == 0:
            cases.append(rng.standard_normal(5))
        elif k % 3 == 1:
            cases.append(np.poly(rng.standard_normal(4) * 10.0 ** rng.integers(-3, 4, 4))[::-1] * rng.standard_normal())
        else:
            a = rng.standard_normal() + 1j * rng.standard_normal()
            cases.append(np.real(np.poly([a, np.conj(a), 100 * rng.standard_normal(), 1e-3 * rng.standard_normal()]))[::-1])
    for c in cases:
        ref = np.sort(np.real(np.roots(c[::-1])))
        got = np.sort(harness.quartic(c))
        assert len(ref) == len(got)
        assert np.all(np.abs(ref - got) <= 1e-7 * np.maximum(np.abs(ref), 1e-9)), (c, ref, got)


@pytest.mark.parametrize("n_pts,n_lines,B", [(8, 4, 300), (8, 0, 300), (0, 6, 300), (4, 0, 100)])
def test_host_tracked_psd_matches_full_decomposition(n_pts, n_lines, B):
    """pnpl_track.cuh (two tracked eigenpairs + Cholesky certificate per DR iteration; problems whose certificate
    fails for good continue with the full decomposition) against the full 10x10 decomposition every iteration:
    same statuses, poses and -- within a few percent -- iteration counts; and against the oracle on a few problems."""
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=77)
    a = harness.solve(d)
    w = harness.solve_track(d)
    sa, sw = a["status"] & 0xFF, w["status"] & 0xFF
    assert (sa != sw).mean() <= 0.02
    ok = (sa == 0) & (sw == 0) & (a["n_poses"] == 1) & (w["n_poses"] == 1)
    assert ok.mean() > (0.95 if n_pts + n_lines > 4 else 0.4)
    ang = synth.rotation_angle(a["R"][ok, 0], w["R"][ok, 0])
    terr = np.linalg.norm(a["t"][ok, 0] - w["t"][ok, 0], axis=1) / np.linalg.norm(a["t"][ok, 0], axis=1)
    assert ang.max() < 1e-6 and terr.max() < 1e-6
    assert abs(w["iters"][ok].mean() - a["iters"][ok].mean()) <= 0.05 * a["iters"][ok].mean() + 1
    # well-posed families are tracked from start to end (no hand-back); minimal ones often need a third eigenpair
    if n_pts >= 8:
        assert w["fallbacks"].mean() <= 0.03
    if n_pts + n_lines <= 4:
        return   # minimal sets: the optimal face need not be a point, another solver may land elsewhere on it
                 # (those families are compared as candidate SETS on the same Z: test_extraction_degenerate)
    checked = 0
    for i in np.flatnonzero(ok & (w["fallbacks"] == 0))[:3]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if n_pts and n_lines:
                Ro, to = orc.pnpl(d["pts_2d"][i], d["line_2d"][i], d["pts_3d"][i], d["line_3d"][i], d["K"], max_iters=200000)[0]
            elif n_pts:
                Ro, to = orc.pnp(d["pts_2d"][i], d["pts_3d"][i], d["K"], max_iters=200000)[0]
            else:
                Ro, to = orc.pnl(d["line_2d"][i], d["line_3d"][i], d["K"], max_iters=200000)[0]
        assert synth.rotation_angle(Ro, w["R"][i, 0]) < 1e-6
        assert np.linalg.norm(to - w["t"][i, 0]) / np.linalg.norm(to) < 1e-6
        checked += 1
    assert checked == 3


@pytest.mark.parametrize("n_pts,n_lines,B", [(8, 4, 300), (0, 6, 300), (4, 0, 100)])
def test_host_two_threads_per_problem_matches_one(n_pts, n_lines, B):
    """pnpl_track2.cuh (the tracked solver with every step of an iteration split between two threads, here two host
    threads and a spinning barrier where the kernel has its barriers) against pnpl_track.cuh: same statuses, same
    hand-backs, same poses; iteration counts differ by the summation order of the split dot products only."""
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=78)
    a = harness.solve_track(d)
    w = harness.solve_track(d, two=True)
    sa, sw = a["status"] & 0xFF, w["status"] & 0xFF
    assert (sa != sw).mean() <= 0.02
    assert abs(int(a["fallbacks"].sum()) - int(w["fallbacks"].sum())) <= max(2, 0.1 * a["fallbacks"].sum())
    ok = (sa == 0) & (sw == 0) & (a["n_poses"] == 1) & (w["n_poses"] == 1)
    assert ok.mean() > (0.95 if n_pts + n_lines > 4 else 0.4)
    ang = synth.rotation_angle(a["R"][ok, 0], w["R"][ok, 0])
    terr = np.linalg.norm(a["t"][ok, 0] - w["t"][ok, 0], axis=1) / np.linalg.norm(a["t"][ok, 0], axis=1)
    assert ang.max() < 1e-6 and terr.max() < 1e-6
    assert abs(w["iters"][ok].mean() - a["iters"][ok].mean()) <= 0.03 * a["iters"][ok].mean() + 1


@pytest.mark.parametrize("n_pts,n_lines,B", [(8, 4, 400), (0, 6, 400), (8, 0, 300), (4, 0, 150)])
def test_host_stale_angle_sweep_matches_exact_sweep(n_pts, n_lines, B):
    """The sweep the warp-per-problem kernel runs inside a DR iteration (pnpl_warp.cuh: warp_sweep_stale) takes all 45
    rotation angles from the matrix as it is when the sweep begins.  The same change in the thread-form sweep
    (jacobi_sweep_reg behind -DCVX_STALE_ANGLES, host build only -- there it replaces EVERY sweep: cold start,
    iterations and polishing) against the exact round-by-round sweep: same statuses and poses, and the DR iteration does
    not notice (mean iteration counts within 2 % + 1 on well-posed families, within 10 % on 4 points)."""
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=79)
    a = harness.solve(d)
    with harness.load_variant("stale", ["CVX_STALE_ANGLES"])():
        w = harness.solve(d)
    sa, sw = a["status"] & 0xFF, w["status"] & 0xFF
    assert (sa != sw).mean() <= 0.02
    ok = (sa == 0) & (sw == 0) & (a["n_poses"] == 1) & (w["n_poses"] == 1)
    assert ok.mean() > (0.95 if n_pts + n_lines > 4 else 0.4)
    ang = synth.rotation_angle(a["R"][ok, 0], w["R"][ok, 0])
    terr = np.linalg.norm(a["t"][ok, 0] - w["t"][ok, 0], axis=1) / np.linalg.norm(a["t"][ok, 0], axis=1)
    assert ang.max() < 1e-6 and terr.max() < 1e-6
    tol = 0.02 if n_pts + n_lines > 4 else 0.10
    assert abs(w["iters"][ok].mean() - a["iters"][ok].mean()) <= tol * a["iters"][ok].mean() + 1
