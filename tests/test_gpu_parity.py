"""GPU parity tests (`-m gpu`): the CUDA path, called through the C ABI, against the
CPU oracle and the golden vectors from the verbatim reference.

Tolerance (BASELINE.json north_star): rotation <= 1e-6 rad, relative translation
<= 1e-6 against the oracle on identical inputs."""
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROT_TOL = 1e-6
T_TOL = 1e-6


@pytest.fixture(scope="module")
def cb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cvxpnpl_b200
    return cvxpnpl_b200


def _cuda(d, *keys):
    return [torch.from_numpy(np.ascontiguousarray(d[k])).cuda() for k in keys]


def _oracle_call(orc, d, i, n_pts, n_lines, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if n_pts and n_lines:
            return orc.pnpl(d["pts_2d"][i], d["line_2d"][i], d["pts_3d"][i], d["line_3d"][i], d["K"], **kw)
        if n_pts:
            return orc.pnp(d["pts_2d"][i], d["pts_3d"][i], d["K"], **kw)
        return orc.pnl(d["line_2d"][i], d["line_3d"][i], d["K"], **kw)


def _solve(cb, d, n_pts, n_lines, **kw):
    K = torch.from_numpy(d["K"]).cuda()
    args = {}
    if n_pts:
        args.update(pts_2d=_cuda(d, "pts_2d")[0], pts_3d=_cuda(d, "pts_3d")[0])
    if n_lines:
        args.update(line_2d=_cuda(d, "line_2d")[0], line_3d=_cuda(d, "line_3d")[0])
    res = cb.solve_batched(K, **args, **kw)
    torch.cuda.synchronize()
    return res


def test_library_loaded(cb):
    from cvxpnpl_b200 import _lib
    assert b"sm_100a" in _lib.load().cvxpnpl_b200_version()


def test_assembly_matches_reference(cb, golden):
    """cvxpnpl.py:20-153, 623-624, 475 -- against A'A and B computed by the verbatim
    reference (tests/golden/synth.npz)."""
    g = golden["synth"]
    for name, n_pts, n_lines in (("pnp8", 8, 0), ("pnpl8_4", 8, 4), ("pnl6", 0, 6)):
        for noise in (0, 1, 2):
            key = f"{name}_s{noise}"
            d = {k: g[f"{key}_{k}"] for k in ("pts_2d", "pts_3d", "line_2d", "line_3d", "K")}
            Q, Bm = cb.assemble_batched(d["K"], d["pts_2d"] if n_pts else None, d["pts_3d"] if n_pts else None,
                                        d["line_2d"] if n_lines else None, d["line_3d"] if n_lines else None)
            torch.cuda.synchronize()
            assert np.allclose(Q.cpu().numpy(), g[key + "_AtA"], rtol=0, atol=1e-13)
            assert np.allclose(Bm.cpu().numpy(), g[key + "_B"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("name", ["pnp", "pnl", "pnpl"])
def test_examples_known_answer(cb, golden, name):
    """examples/pnp.py, pnl.py, pnpl.py through the reference-shaped scalar API."""
    from cvxpnpl_b200 import synth
    e = golden["examples"]
    kw = {k[len(name) + 1:]: e[k] for k in e.files if k.startswith(name + "_")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if name == "pnp":
            poses = cb.pnp(pts_2d=kw["pts_2d"], pts_3d=kw["pts_3d"], K=kw["K"])
        elif name == "pnl":
            poses = cb.pnl(line_2d=kw["line_2d"], line_3d=kw["line_3d"], K=kw["K"])
        else:
            poses = cb.pnpl(pts_2d=kw["pts_2d"], line_2d=kw["line_2d"], pts_3d=kw["pts_3d"],
                            line_3d=kw["line_3d"], K=kw["K"])
    assert len(poses) == 1
    R, t = poses[0]
    assert R.shape == (3, 3) and t.shape == (3,)
    # ground truth is printed with 8 decimals in the examples
    assert synth.rotation_angle(kw["R_gt"], R) < 2e-7
    assert np.linalg.norm(t - kw["t_gt"]) / np.linalg.norm(kw["t_gt"]) < 2e-7
    # and the verbatim reference's output (golden)
    assert synth.rotation_angle(kw["R"][0], R) < ROT_TOL
    assert np.linalg.norm(t - kw["t"][0]) / np.linalg.norm(kw["t"][0]) < T_TOL


@pytest.mark.parametrize("name,n_pts,n_lines", [("pnp8", 8, 0), ("pnpl8_4", 8, 4), ("pnl6", 0, 6)])
def test_golden_synth(cb, golden, name, n_pts, n_lines):
    """Poses returned by the verbatim reference (SDP solved by the oracle shim)."""
    from cvxpnpl_b200 import synth
    g = golden["synth"]
    for noise in (0, 1, 2):
        key = f"{name}_s{noise}"
        d = {k: g[f"{key}_{k}"] for k in ("pts_2d", "pts_3d", "line_2d", "line_3d", "K")}
        res = _solve(cb, d, n_pts, n_lines)
        assert (res.n_poses.cpu().numpy() == g[key + "_n"]).all()
        R, t = res.R.cpu().numpy()[:, 0], res.t.cpu().numpy()[:, 0]
        ang = synth.rotation_angle(g[key + "_R"][:, 0], R)
        terr = np.linalg.norm(t - g[key + "_t"][:, 0], axis=1) / np.linalg.norm(g[key + "_t"][:, 0], axis=1)
        assert ang.max() < ROT_TOL, (key, ang)
        assert terr.max() < T_TOL, (key, terr)


@pytest.mark.parametrize("n_pts,n_lines,noise", [(8, 4, 1.0), (8, 0, 2.0), (8, 4, 0.0), (6, 0, 1.0), (0, 8, 1.0),
                                                 (0, 6, 1.0), (8, 4, 2.0)])
def test_parity_vs_oracle_seeded(cb, n_pts, n_lines, noise):
    """Same seeded inputs through the CUDA path and the CPU oracle (restated SCS).  EVERY problem of
    the batch is compared -- a problem that stops at the reference's iteration cap (status
    MAX_ITERS, SCS "solved_inaccurate") is held to the same tolerance, not dropped."""
    from cvxpnpl_b200 import synth
    from oracle import cvxpnpl_oracle as orc
    from oracle import kkt
    B = 400
    d = synth.make_batch(B, n_pts, n_lines, noise=noise, seed=11)
    res = _solve(cb, d, n_pts, n_lines, return_Z=True)
    R, t, Z = res.R.cpu().numpy(), res.t.cpu().numpy(), res.Z.cpu().numpy()
    st = res.status.cpu().numpy() & 0xFF
    assert np.isin(st, (0, 1)).all(), np.bincount(st)
    rot, tr = np.zeros(B), np.zeros(B)
    for i in range(B):
        poses, aux = _oracle_call(orc, d, i, n_pts, n_lines, max_iters=200000, return_aux=True)
        assert len(poses) == int(res.n_poses[i]) == 1, (i, len(poses), int(res.n_poses[i]), st[i])
        Ro, to = poses[0]
        rot[i] = synth.rotation_angle(Ro, R[i, 0])
        tr[i] = np.linalg.norm(to - t[i, 0]) / np.linalg.norm(to)
        # solver-independent certificate on the CUDA Z, using the oracle's multipliers
        cert = kkt.certificate(aux["Q"], Z[i], aux["info"]["y"])
        assert cert["eq_res"] < 1e-7 and cert["psd_res"] < 1e-9 and cert["gap"] < 1e-7, (i, st[i], cert)
    assert rot.max() < ROT_TOL and tr.max() < T_TOL, (rot.max(), tr.max(), int(rot.argmax()), np.bincount(st))


def test_full_size_properties(cb):
    """BASELINE config 3 at full size (1e5 x PnPL 8+4): size-independent checks --
    every R orthonormal, SDP objective == dual objective, noise-free ground truth
    recovered, and the batch result identical to a re-solve of a random subset."""
    from cvxpnpl_b200 import synth
    B = 100_000
    d = synth.make_batch(B, 8, 4, noise=0.0, seed=5)
    res = _solve(cb, d, 8, 4)
    R, t = res.R[:, 0], res.t[:, 0]
    assert int((res.n_poses != 1).sum()) == 0
    assert int(((res.status & 0xFF) != 0).sum()) == 0
    I = torch.eye(3, dtype=torch.float64, device=R.device)
    assert float((R @ R.transpose(1, 2) - I).abs().max()) < 1e-12
    assert float((res.obj[:, 0] - res.obj[:, 1]).abs().max()) < 1e-8
    ang, terr = synth.pose_error(d["R_gt"], d["t_gt"], R.cpu().numpy(), t.cpu().numpy())
    assert ang.max() < 1e-6 and terr.max() < 1e-6, (ang.max(), terr.max())
    idx = np.random.default_rng(0).choice(B, 257, replace=False)
    sub = {k: (v[idx] if k != "K" else v) for k, v in d.items()}
    res2 = _solve(cb, sub, 8, 4)
    # (not bit-identical: the Anderson history column a lane writes to is warp-uniform,
    # so rounding depends on which problems share a warp)
    assert float((res2.R[:, 0] - res.R[idx, 0]).abs().max()) < 1e-8
    assert float((res2.t[:, 0] - res.t[idx, 0]).abs().max()) < 1e-8


def test_edge_cases(cb):
    """Empty batch, batch not a multiple of the CTA size, per-problem K, NaN input,
    fewer than 3 elements through the plugin class."""
    from cvxpnpl_b200 import synth
    d = synth.make_batch(0, 8, 4, seed=1)
    res = _solve(cb, d, 8, 4)
    assert res.R.shape == (0, 4, 3, 3)
    d = synth.make_batch(131, 8, 4, noise=1.0, seed=2)
    res = _solve(cb, d, 8, 4)
    dK = dict(d)
    dK["K"] = np.repeat(d["K"][None], 131, axis=0)
    resK = _solve(cb, dK, 8, 4)
    assert float((res.R[:, 0] - resK.R[:, 0]).abs().max()) < 1e-8 and float((res.t[:, 0] - resK.t[:, 0]).abs().max()) < 1e-8
    bad = {k: v.copy() for k, v in d.items()}
    bad["pts_2d"][5, 0, 0] = np.nan
    resb = _solve(cb, bad, 8, 4)
    assert int(resb.status[5]) & 0xFF == 2 and int(resb.n_poses[5]) == 1
    assert torch.isnan(resb.R[5]).all()
    ok = np.ones(131, bool)
    ok[5] = False
    assert float((resb.R[ok][:, 0] - res.R[ok][:, 0]).abs().max()) < 1e-8
    poses = cb.CvxPnPL.estimate_pose(d["K"], pts_2d=d["pts_2d"][0, :2], pts_3d=d["pts_3d"][0, :2])
    assert len(poses) == 1 and np.isnan(poses[0][0]).all()


@pytest.mark.parametrize("name", ["pts4", "pts3", "lines3", "lines4", "p2l1", "coplanar8"])
def test_extraction_degenerate(cb, golden, name):
    """Multi-solution extraction (cvxpnpl.py:221-343, 156-218) replayed on the Z the
    reference saw (golden), compared with the reference's candidate poses as sets; where the
    reference raised LinAlgError the CUDA path must report ST_SINGULAR."""
    g = golden["degenerate"]
    Z, Q, Bm = g[name + "_Z"], g[name + "_AtA"], g[name + "_B"]
    res = cb.extract_batched(Z, Q, Bm)
    torch.cuda.synchronize()
    R, t = res.R.cpu().numpy(), res.t.cpu().numpy()
    npo, st = res.n_poses.cpu().numpy(), res.status.cpu().numpy()
    for i in range(len(Z)):
        if not g[name + "_ok"][i]:
            assert st[i] & 0xFF == 3 and npo[i] == 0, (name, i, st[i], npo[i])   # ST_SINGULAR <-> LinAlgError
            continue
        n = int(g[name + "_n"][i])
        assert npo[i] == n, (name, i, npo[i], n, st[i])
        got = np.concatenate([R[i, :n].reshape(n, 9), t[i, :n]], axis=1)
        exp = np.concatenate([g[name + "_R"][i, :n].reshape(n, 9), g[name + "_t"][i, :n]], axis=1)
        dist = np.sort(np.abs(got[:, None, :] - exp[None, :, :]).max(-1).min(1))
        # see tests/test_oracle.py::test_degenerate_extraction for the tolerance
        assert dist[(n - 1) // 2] < 1e-6, (name, i, dist)
        assert np.all(dist[: max(n - 1, 1)] < 1e-4), (name, i, dist)


@pytest.mark.parametrize("name", ["pts4", "pts3", "lines3", "lines4", "p2l1", "coplanar8", "pts5"])
def test_extraction_degenerate_big(cb, name):
    """tests/golden/degenerate_big.npz (40 problems per family, verbatim reference): the reference's
    exceptions (LinAlgError at cvxpnpl.py:165 / 212 / 510) map to ST_SINGULAR, the number of candidates
    is the reference's, and EVERY candidate the reference itself reproduces under a 1e-14 perturbation
    of Z is matched to 1e-6 (the others sit on near-double roots of the resultant quartic and are
    only counted; see oracle/candidate_sets.py)."""
    import os
    from oracle import candidate_sets as du
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "degenerate_big.npz"))
    res = cb.extract_batched(g[name + "_Z"], g[name + "_AtA"], g[name + "_B"])
    torch.cuda.synchronize()
    R, t = res.R.cpu().numpy(), res.t.cpu().numpy()
    npo, st = res.n_poses.cpu().numpy(), res.status.cpu().numpy() & 0xFF
    total = stable = 0
    for i in range(len(npo)):
        if g[name + "_err"][i] == 1:
            assert st[i] == 3 and npo[i] == 0, (name, i, st[i], npo[i])
            continue
        n = int(g[name + "_n"][i])
        assert npo[i] == n and st[i] == 0, (name, i, npo[i], n, st[i])
        exp, got = du.flat(g[name + "_R"][i], g[name + "_t"][i], n), du.flat(R[i], t[i], n)
        m = du.stable_mask(exp, g[name + "_Rp"][i], g[name + "_tp"][i], g[name + "_np"][i])
        total, stable = total + n, stable + int(m.sum())
        assert du.compare(got, exp, m) < 1e-6, (name, i, du.compare(got, exp, m))
    assert stable >= 0.85 * total, (name, stable, total)     # the filter must not hollow the test out


@pytest.mark.parametrize("n_pts,n_lines,coplanar", [(3, 0, False), (4, 0, False), (0, 3, False), (2, 1, False),
                                                    (0, 4, False), (8, 0, True)])
def test_degenerate_candidates_vs_oracle_same_Z(cb, n_pts, n_lines, coplanar):
    """BASELINE config 5 on 600 FRESH problems per family: the CUDA path's candidate poses against the
    oracle's extraction (pinned to the verbatim reference by tests/test_oracle.py) fed the SAME Z the
    CUDA solver returned, compared as sets.  Candidates the oracle itself does not reproduce under a
    1e-14 perturbation of Z (near-double quartic roots) are excluded and counted."""
    from cvxpnpl_b200 import synth
    from oracle import cvxpnpl_oracle as orc
    from oracle import candidate_sets as du
    B = 600
    d = synth.make_batch(B, n_pts, n_lines, noise=0.0, seed=19, coplanar=coplanar)
    res = _solve(cb, d, n_pts, n_lines, return_Z=True)
    R, t, Z = res.R.cpu().numpy(), res.t.cpu().numpy(), res.Z.cpu().numpy()
    npo, st = res.n_poses.cpu().numpy(), res.status.cpu().numpy() & 0xFF
    rng = np.random.default_rng(3)
    total = stable = compared = errors = borderline = err_mismatch = 0
    for i in range(B):
        C, N = orc._stack(d["pts_2d"][i] if n_pts else None, d["pts_3d"][i] if n_pts else None,
                          d["line_2d"][i] if n_lines else None, d["line_3d"][i] if n_lines else None, d["K"])
        A, Bm = orc.reduce_translation(C, N)
        lam = np.linalg.eigvalsh(Z[i])
        if np.any(np.abs(lam - 1e-3) < 1e-6):
            borderline += 1                     # an eigenvalue sits ON the rank threshold (cvxpnpl.py:502)
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                poses = orc.extract(Z[i], A, Bm)
            except np.linalg.LinAlgError:
                # an EXACTLY singular system in numpy (LAPACK info > 0).  Exact zeros are not stable under a
                # change of eigensolver (the CUDA path extracts from its own eigenvectors, numpy from eigh(Z)),
                # so single disagreements are tolerated and bounded below, not asserted one by one
                errors += 1
                err_mismatch += int(not (st[i] == 3 and npo[i] == 0))
                continue
            pert = []
            for _ in range(3):
                try:
                    pert.append(orc.extract(Z[i] * (1.0 + 1e-14 * rng.standard_normal((10, 10))), A, Bm))
                except np.linalg.LinAlgError:
                    pert.append([])
        n = len(poses)
        assert npo[i] == n and st[i] in (0, 1), (i, npo[i], n, st[i])
        exp = np.array([np.concatenate([Rr.ravel(), tt]) for Rr, tt in poses])
        got = du.flat(R[i], t[i], n)
        Rp = np.full((3, 4, 3, 3), np.nan)
        tp = np.full((3, 4, 3), np.nan)
        for k, pp in enumerate(pert):
            for c, (Rr, tt) in enumerate(pp):
                Rp[k, c], tp[k, c] = Rr, tt
        m = du.stable_mask(exp, Rp, tp, np.array([len(pp) for pp in pert]))
        total, stable, compared = total + n, stable + int(m.sum()), compared + 1
        assert du.compare(got, exp, m) < 1e-6, (i, n, du.compare(got, exp, m))
    assert compared + errors >= 0.98 * B and stable >= 0.8 * total, (compared, errors, borderline, stable, total)
    assert err_mismatch <= max(1, 0.01 * B), (err_mismatch, errors)


@pytest.mark.parametrize("n_pts,n_lines,coplanar,min_found", [
    (3, 0, False, 0.75), (4, 0, False, 0.95), (0, 3, False, 0.5), (2, 1, False, 0.6), (0, 4, False, 0.95),
    (8, 0, True, 0.25)])
def test_degenerate_sweep(cb, n_pts, n_lines, coplanar, min_found):
    """BASELINE config 5 (minimal / planar configurations, rank > 1 SDP optima).  The
    optimal face is not a point, so poses are not compared with the oracle's; checked
    instead: status codes are legal, every finite candidate is orthonormal, and on
    these NOISE-FREE problems the ground truth is among the candidates for at least
    the fraction the reference's own extraction achieves (measured with the oracle;
    its rank-2 averaging, cvxpnpl.py:303-315, degenerates to NaN on about half of the
    planar cases)."""
    from cvxpnpl_b200 import synth
    B = 1700
    d = synth.make_batch(B, n_pts, n_lines, noise=0.0, seed=9, coplanar=coplanar)
    res = _solve(cb, d, n_pts, n_lines)
    st = res.status.cpu().numpy() & 0xFF
    n = res.n_poses.cpu().numpy()
    assert np.isin(st, (0, 1, 3, 4)).all()
    assert np.isin(n, (0, 1, 2, 4)).all()
    R = res.R.cpu().numpy()
    fin = np.isfinite(R).all(axis=(2, 3))
    assert (fin.sum(1) <= n).all()
    RtR = np.einsum("bkij,bklj->bkil", R, R)
    assert np.abs(RtR[fin] - np.eye(3)).max() < 1e-9
    ang = synth.rotation_angle(d["R_gt"][:, None], R)          # [B, 4]
    ang = np.where(fin, ang, np.inf)
    conv = st == 0
    found = (ang.min(axis=1) < 1e-4)[conv].mean()
    assert found >= min_found, found


def test_full_size_pnp8_and_pnl6(cb):
    """BASELINE configs 2 and 4 at reduced-but-large size through size-independent
    properties (fp64 path): noise-free ground truth recovered wherever the solver
    reports convergence with one pose, rotations orthonormal, duality gap closed."""
    from cvxpnpl_b200 import synth
    for n_pts, n_lines, B in ((8, 0, 100_000), (0, 6, 50_000)):
        d = synth.make_batch(B, n_pts, n_lines, noise=0.0, seed=3)
        res = _solve(cb, d, n_pts, n_lines)
        st = (res.status & 0xFF).cpu().numpy()
        n = res.n_poses.cpu().numpy()
        ok = (st == 0) & (n == 1)
        assert ok.mean() > (0.99 if n_pts else 0.85), ok.mean()
        R, t = res.R[:, 0].cpu().numpy(), res.t[:, 0].cpu().numpy()
        ang, terr = synth.pose_error(d["R_gt"], d["t_gt"], R, t)
        assert np.nanmax(ang[ok]) < 1e-5 and np.nanmax(terr[ok]) < 1e-5, (np.nanmax(ang[ok]), np.nanmax(terr[ok]))
        gap = (res.obj[:, 0] - res.obj[:, 1]).abs().cpu().numpy()
        assert np.nanmax(gap[ok]) < 1e-7


def test_stage_entry_points_compose(cb):
    """assemble -> solve_sdp -> extract through the three stage entry points equals
    the fused kernel."""
    from cvxpnpl_b200 import synth
    d = synth.make_batch(300, 8, 4, noise=1.0, seed=13)
    fused = _solve(cb, d, 8, 4)
    Q, Bm = cb.assemble_batched(d["K"], d["pts_2d"], d["pts_3d"], d["line_2d"], d["line_3d"])
    Z, dobj, iters, status = cb.solve_sdp_batched(Q)
    res = cb.extract_batched(Z, Q, Bm, dobj)
    torch.cuda.synchronize()
    # same algorithm; the Anderson history differs (a batch this small runs in the warp kernel: FP32 history)
    assert float((iters - fused.iters).abs().float().mean()) < 6.0
    assert float((res.R[:, 0] - fused.R[:, 0]).abs().max()) < 1e-9
    assert float((res.t[:, 0] - fused.t[:, 0]).abs().max()) < 1e-9


def test_optimality_flag_and_scalar_warnings(cb):
    """cvxpnpl.py:516-519: |r'Qr - dual objective| > eps raises the 'not certifiably
    optimal' warning; here it is a status flag in the batch API and a warning in the
    scalar drop-in.  An iteration cap of 15 leaves the problem unconverged."""
    from cvxpnpl_b200 import synth
    d = synth.make_batch(64, 8, 4, noise=1.0, seed=33)
    res = _solve(cb, d, 8, 4, max_iters=15)
    st = res.status.cpu().numpy()
    assert ((st & 0xFF) == 1).all()                 # MAX_ITERS (SCS: solved_inaccurate)
    assert ((st & 0x100) != 0).mean() > 0.9         # flagged as not certifiably optimal
    assert (res.iters.cpu().numpy() == 15).all()
    with pytest.warns(UserWarning, match="not certifiably optimal"):
        poses = cb.pnpl(d["pts_2d"][0], d["line_2d"][0], d["pts_3d"][0], d["line_3d"][0], d["K"], max_iters=15)
    assert len(poses) >= 1
    # converged solve: no flag, no warning
    res = _solve(cb, d, 8, 4)
    assert (res.status.cpu().numpy() == 0).all()
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        cb.pnpl(d["pts_2d"][0], d["line_2d"][0], d["pts_3d"][0], d["line_3d"][0], d["K"])


def test_plugin_class_matches_scalar_api(cb):
    """benchmarks/toolkit/methods/p*.py: K-first estimate_pose with the harness's
    keyword names (suite.py:80)."""
    from cvxpnpl_b200 import synth
    d = synth.make_batch(1, 8, 4, noise=1.0, seed=34)
    kw = dict(pts_2d=d["pts_2d"][0], line_2d=d["line_2d"][0], pts_3d=d["pts_3d"][0], line_3d=d["line_3d"][0])
    a = cb.CvxPnPL.estimate_pose(d["K"], **kw)
    b = cb.pnpl(kw["pts_2d"], kw["line_2d"], kw["pts_3d"], kw["line_3d"], d["K"])
    assert len(a) == len(b) == 1
    assert np.allclose(a[0][0], b[0][0], atol=1e-9) and np.allclose(a[0][1], b[0][1], atol=1e-9)
    assert cb.CvxPnPL.name == "CvxPnPL" and cb.CvxPnPL.loaded
    p = cb.CvxPnPL.estimate_pose(d["K"], pts_2d=kw["pts_2d"], pts_3d=kw["pts_3d"])
    q = cb.CvxPnPL.estimate_pose(d["K"], line_2d=d["line_2d"][0], line_3d=d["line_3d"][0])
    assert len(p) >= 1 and len(q) >= 1


def test_batched_synth_suite(cb):
    """the reference's synthetic grid as whole-cell batches on the
    GPU.  Noise free => ground truth; error grows with noise and shrinks with n."""
    from cvxpnpl_b200 import suite
    grid = suite.run_grid("pnp", n_elements=(6, 12), noises=(0.0, 2.0), runs=4000, seed=1)
    assert grid[(6, 0.0)].ang_median_deg < 1e-4 and grid[(12, 0.0)].trans_median < 1e-6
    assert grid[(12, 2.0)].ang_median_deg < grid[(6, 2.0)].ang_median_deg < 5.0
    assert grid[(12, 2.0)].failed == 0.0
    cell = suite.run_cell("pnpl", 8, 1.0, 4000, seed=2)
    assert cell.ang_median_deg < 2.0 and cell.failed < 0.01
    cell = suite.run_cell("pnl", 4, 0.0, 2000, seed=3)          # minimal-ish: several candidates
    # (the share of multi-candidate results shrinks as convergence improves: most rank-2 results
    # of 4-line problems are iterates that stopped at the cap)
    # (0.045 in round 1; 0.0045 with the dual guess kappa = 0.9 for small problems, which took the 4-line problems
    # at the cap from 2 % to 0.5 %)
    assert cell.multi > 0.001 and cell.ang_median_deg < 1e-3


def test_rc_variant(cb, golden):
    """the "rc" operator (benchmarks/toolkit/methods/rc.py, six row
    orthonormality equalities removed) against the verbatim reference's poses."""
    from cvxpnpl_b200 import synth
    g = golden["rc"]
    res = cb.solve_batched(g["K"], pts_2d=g["pts_2d"], pts_3d=g["pts_3d"], variant="rc")
    torch.cuda.synchronize()
    assert (res.n_poses.cpu().numpy() == g["n"]).all() and ((res.status & 0xFF) == 0).all()
    R, t = res.R.cpu().numpy()[:, 0], res.t.cpu().numpy()[:, 0]
    assert synth.rotation_angle(g["R"][:, 0], R).max() < ROT_TOL
    assert (np.linalg.norm(t - g["t"][:, 0], axis=1) / np.linalg.norm(g["t"][:, 0], axis=1)).max() < T_TOL
    poses = cb.rc(g["pts_2d"][0], g["pts_3d"][0], g["K"])
    assert len(poses) == 1 and synth.rotation_angle(g["R"][0, 0], poses[0][0]) < ROT_TOL


def test_null_baseline(cb):
    """benchmarks/toolkit/methods/pnp.py:24-55 restated in numpy
    (smallest right singular vector of A, SVD projection, det sign fix, t = -B r)."""
    from cvxpnpl_b200 import synth
    from oracle import cvxpnpl_oracle as orc
    d = synth.make_batch(50, 8, 0, noise=1.0, seed=41)
    res = cb.null_batched(d["pts_2d"], d["pts_3d"], d["K"])
    torch.cuda.synchronize()
    R, t = res.R.cpu().numpy()[:, 0], res.t.cpu().numpy()[:, 0]
    for i in range(50):
        C, N = orc.point_constraints(d["pts_2d"][i], d["pts_3d"][i], d["K"])
        A, B = orc.reduce_translation(C, N)
        Rn = np.linalg.svd(A)[2][-1].reshape((3, 3)).T
        U, _, Vt = np.linalg.svd(Rn)
        Rn = U @ Vt
        Rn *= np.sign(np.linalg.det(Rn))
        tn = -B @ Rn.ravel("F")
        assert synth.rotation_angle(Rn, R[i]) < 1e-7, i
        assert np.linalg.norm(tn - t[i]) / np.linalg.norm(tn) < 1e-7, i


def test_large_n_assembly(cb):
    """SURVEY 8f rank 3 (benchmarks/scalability/pnp.py:37-40, n up to 10 000): the
    chunked streaming assembly against the oracle's A'A and B, and the full path
    against ground truth."""
    from cvxpnpl_b200 import synth
    from oracle import cvxpnpl_oracle as orc
    # staging="tma": the point slabs go through shared memory with bulk-asynchronous copies (accumulate_tma_kernel:
    # even point counts; 1537 = one full tile ring + a ragged last tile, odd -> the plain-load kernel takes the points);
    # "loads": the plain-load kernel.  Both against the oracle, and against each other.
    for n_pts, n_lines in ((1000, 0), (5000, 300), (0, 9000), (1537, 0), (3074, 0), (512, 256), (20000, 0)):
        d = synth.make_batch(3, n_pts, n_lines, noise=1.0, seed=51)
        got = {}
        for staging in ("tma", "loads"):
            Q, Bm = cb.assemble_batched(d["K"], d["pts_2d"] if n_pts else None, d["pts_3d"] if n_pts else None,
                                        d["line_2d"] if n_lines else None, d["line_3d"] if n_lines else None,
                                        staging=staging)
            torch.cuda.synchronize()
            got[staging] = (Q.cpu().numpy(), Bm.cpu().numpy())
            for i in range(3):
                C, N = orc._stack(d["pts_2d"][i] if n_pts else None, d["pts_3d"][i] if n_pts else None,
                                  d["line_2d"][i] if n_lines else None, d["line_3d"][i] if n_lines else None, d["K"])
                A, B = orc.reduce_translation(C, N)
                ref = A.T @ A
                assert np.abs(got[staging][0][i] - ref).max() < 1e-11 * np.abs(ref).max()
                assert np.allclose(got[staging][1][i], B, rtol=1e-9, atol=1e-11)
        assert np.abs(got["tma"][0] - got["loads"][0]).max() <= 1e-13 * np.abs(got["loads"][0]).max()
    d = synth.make_batch(40, 2000, 0, noise=1.0, seed=52)
    res = _solve(cb, d, 2000, 0)
    assert ((res.status & 0xFF) == 0).all() and (res.n_poses == 1).all()
    ang, terr = synth.pose_error(d["R_gt"], d["t_gt"], res.R[:, 0].cpu().numpy(), res.t[:, 0].cpu().numpy())
    assert ang.max() < 2e-3 and terr.max() < 2e-3      # 2000 points average the 1 px noise down
    poses = cb.pnp(d["pts_2d"][0], d["pts_3d"][0], d["K"])
    assert len(poses) == 1 and synth.rotation_angle(poses[0][0], res.R[0, 0].cpu().numpy()) < 1e-9


@pytest.mark.parametrize("n_pts,n_lines", [(8, 4), (0, 6), (4, 0)])
def test_straggler_warp_path_matches_thread_path(cb, n_pts, n_lines):
    """The warp-per-problem straggler kernel (csrc/pnpl_warp.cuh) runs the same DR
    iteration as the thread-per-problem solver.  handoff=1 pushes every problem
    through it after one pass (the batch is smaller than the grid, so the queue is
    empty at once); handoff=-1 never hands over.  Same poses, same statuses."""
    from cvxpnpl_b200 import synth
    B = 2000
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=77)
    a = _solve(cb, d, n_pts, n_lines, handoff=-1)
    w = _solve(cb, d, n_pts, n_lines, handoff=1)
    assert a.launches == 5 and w.launches == 7   # pre + 2 sort + solver + finish (+ straggler + resume)
    sa, sw = (a.status & 0xFF).cpu().numpy(), (w.status & 0xFF).cpu().numpy()
    ok = (sa == 0) & (sw == 0) & (a.n_poses.cpu().numpy() == 1) & (w.n_poses.cpu().numpy() == 1)
    assert ok.mean() > (0.95 if n_pts + n_lines > 4 else 0.5)
    # a problem converges on both paths or on neither, up to the few that sit at the cap
    assert (sa != sw).mean() < (0.02 if n_pts + n_lines > 4 else 0.06)
    Ra, Rw = a.R[:, 0].cpu().numpy()[ok], w.R[:, 0].cpu().numpy()[ok]
    ta, tw = a.t[:, 0].cpu().numpy()[ok], w.t[:, 0].cpu().numpy()[ok]
    ang = synth.rotation_angle(Ra, Rw)
    terr = np.linalg.norm(ta - tw, axis=1) / np.linalg.norm(ta, axis=1)
    assert ang.max() <= ROT_TOL and terr.max() <= T_TOL, (ang.max(), terr.max())
    ia, iw = a.iters.cpu().numpy()[ok], w.iters.cpu().numpy()[ok]
    # the hand-over only restarts the Anderson history: similar iteration counts
    assert np.median(iw) <= 1.3 * np.median(ia) + 10


@pytest.mark.parametrize("handoff", [-1, 1])
def test_plateau_jump_hard_problems(cb, handoff):
    """tests/golden/hard_pnpl.npz: plateau walkers (247-799 DR iterations without the plateau jump of
    pass_dr / the warp kernel).  Thread path (handoff=-1) and warp path (handoff=1): fewer than half the
    iterations, poses equal to the oracle's."""
    import os
    import warnings
    from cvxpnpl_b200 import synth
    from oracle import cvxpnpl_oracle as orc
    h = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "hard_pnpl.npz")))
    res = _solve(cb, h, 8, 4, handoff=handoff)
    it = res.iters.cpu().numpy()
    assert ((res.status & 0xFF) == 0).all() and (res.n_poses == 1).all()
    assert it.max() <= 400 and it.sum() <= 0.5 * h["iters_without_jump"].sum(), it
    R, t = res.R[:, 0].cpu().numpy(), res.t[:, 0].cpu().numpy()
    for i in range(len(it)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Ro, to = orc.pnpl(h["pts_2d"][i], h["line_2d"][i], h["pts_3d"][i], h["line_3d"][i], h["K"],
                              max_iters=400000)[0]
        assert synth.rotation_angle(Ro, R[i]) <= ROT_TOL
        assert np.linalg.norm(to - t[i]) / np.linalg.norm(to) <= T_TOL


@pytest.mark.parametrize("n_pts,n_lines", [(0, 6), (8, 4)])
def test_fp32_admm_matches_oracle(cb, n_pts, n_lines):
    """BASELINE.json configs[3]: fp32 ADMM + fp64 extraction (PnL, 6 lines) -- and the
    same switch on the PnPL configuration.  Against the CPU oracle on identical inputs,
    north-star tolerance; and against the pure FP64 path on a larger batch."""
    from cvxpnpl_b200 import synth
    from oracle import cvxpnpl_oracle as orc
    B = 4000
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=91)
    m = _solve(cb, d, n_pts, n_lines, admm_dtype="f32")
    a = _solve(cb, d, n_pts, n_lines)
    assert m.launches == a.launches + 1     # admm32_kernel + ortho_kernel really ran (and no early_kernel before them)
    sm, sa = (m.status & 0xFF).cpu().numpy(), (a.status & 0xFF).cpu().numpy()
    ok = (sm == 0) & (sa == 0) & (m.n_poses.cpu().numpy() == 1) & (a.n_poses.cpu().numpy() == 1)
    assert ok.mean() > 0.95
    Rm, tm = m.R[:, 0].cpu().numpy(), m.t[:, 0].cpu().numpy()
    Ra, ta = a.R[:, 0].cpu().numpy(), a.t[:, 0].cpu().numpy()
    ang = synth.rotation_angle(Ra[ok], Rm[ok])
    terr = np.linalg.norm(ta[ok] - tm[ok], axis=1) / np.linalg.norm(ta[ok], axis=1)
    assert ang.max() <= ROT_TOL and terr.max() <= T_TOL, (ang.max(), terr.max())
    checked = 0
    for i in np.flatnonzero(ok)[:6]:
        Ro, to = _oracle_call(orc, d, i, n_pts, n_lines, max_iters=100000)[0]
        assert synth.rotation_angle(Ro, Rm[i]) <= ROT_TOL
        assert np.linalg.norm(to - tm[i]) / np.linalg.norm(to) <= T_TOL
        checked += 1
    assert checked == 6


def test_kernel_times_and_launch_count(cb):
    """cvxpnpl_b200_kernel_times: CUDA-event time of every kernel of a timed solve; the
    eleven (twelve with the FP32 first phase, which has no early iterations) launches of the tracked path -- ten on the
    caller's stream and the concurrent service kernel on the library's side stream -- are all there and add up to the
    step; the full-decomposition path (psd="full") has seven."""
    from cvxpnpl_b200 import synth
    d = synth.make_batch(20000, 8, 4, noise=1.0, seed=5)
    for admm, n_launch, extra in (("f64", 11, ()), ("f32", 12, ("admm32_kernel", "ortho_kernel"))):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _solve(cb, d, 8, 4, admm_dtype=admm)        # warm-up
        s.record()
        res = _solve(cb, d, 8, 4, admm_dtype=admm, timing=True)
        e.record()
        torch.cuda.synchronize()
        assert res.launches == n_launch
        t = cb.last_kernel_times()
        for k in ("pre_kernel", "solve_track_kernel", "redecomp_kernel", "solve_fused_kernel", "straggler_kernel",
                  "solve_fused_kernel<resume>", "finish_kernel") + extra:
            assert t[k] > 0.0, (k, t)
        if admm == "f64":
            assert t["admm32_kernel"] == 0.0 and t["ortho_kernel"] == 0.0
        assert sum(t.values()) <= s.elapsed_time(e) * 1.05
    full = _solve(cb, d, 8, 4, psd="full", timing=True)
    assert full.launches == 7
    t = cb.last_kernel_times()
    assert t["solve_track_kernel"] == 0.0 and t["redecomp_kernel"] == 0.0 and t["solve_fused_kernel"] > 0.0


@pytest.mark.parametrize("psd", ["track", "track1"])
@pytest.mark.parametrize("n_pts,n_lines,B", [(8, 4, 30000), (8, 0, 30000), (0, 6, 20000), (4, 0, 6000), (5, 3, 10000)])
def test_tracked_psd_matches_full_decomposition(cb, n_pts, n_lines, B, psd):
    """The tracked-eigenpair PSD projection (pnpl_track.cuh: two eigenpairs refined per iteration + Cholesky
    certificate; problems that fail it are finished with the full decomposition) against the full 10x10
    decomposition every iteration (psd="full", the round-1 solver) on the same batch: same statuses, same number
    of poses, same poses, about the same iteration counts.  psd="track": two threads per problem (pnpl_track2.cuh, the
    default); "track1": one thread per problem."""
    from cvxpnpl_b200 import synth
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=1234)
    a = _solve(cb, d, n_pts, n_lines, psd="full")
    w = _solve(cb, d, n_pts, n_lines, psd=psd)
    # early_kernel + solve_track_kernel + redecomp_kernel (+ the concurrent service kernel beside the two-thread solver)
    # really ran
    assert w.launches == a.launches + (4 if psd == "track" else 3)
    sa, sw = (a.status & 0xFF).cpu().numpy(), (w.status & 0xFF).cpu().numpy()
    na, nw = a.n_poses.cpu().numpy(), w.n_poses.cpu().numpy()
    well_posed = n_pts + n_lines >= 8
    # a problem converges on both paths or on neither, up to the few that sit at the iteration cap
    assert (sa != sw).mean() <= (1e-3 if well_posed else 5e-3)
    ok = (sa == 0) & (sw == 0) & (na == 1) & (nw == 1)
    assert ok.mean() > (0.98 if well_posed else 0.45)
    assert ((na != nw) & (sa == 0) & (sw == 0)).mean() <= (1e-4 if well_posed else 5e-3)
    Ra, Rw = a.R[:, 0].cpu().numpy()[ok], w.R[:, 0].cpu().numpy()[ok]
    ta, tw = a.t[:, 0].cpu().numpy()[ok], w.t[:, 0].cpu().numpy()[ok]
    ang = synth.rotation_angle(Ra, Rw)
    terr = np.linalg.norm(ta - tw, axis=1) / np.linalg.norm(ta, axis=1)
    assert ang.max() <= ROT_TOL and terr.max() <= T_TOL, (ang.max(), terr.max())
    ia, iw = a.iters.cpu().numpy()[ok], w.iters.cpu().numpy()[ok]
    assert abs(iw.mean() - ia.mean()) <= 0.03 * ia.mean() + 1


def test_host_pipeline_matches_direct_solve(cb):
    """HostPipeline (double-buffered pinned-host staging) returns the same records as a
    direct solve of the same batches, in order."""
    from cvxpnpl_b200 import synth
    from cvxpnpl_b200.distributed import unpack_record
    batches = [synth.make_batch(3000, 8, 4, noise=1.0, seed=100 + i) for i in range(3)]
    hosts = [{k: torch.from_numpy(b[k]).pin_memory() for k in ("pts_2d", "pts_3d", "line_2d", "line_3d")} for b in batches]
    pipe = cb.HostPipeline(batches[0]["K"], "cuda:0")
    for i, b in enumerate(batches):
        rec = pipe.step(hosts[i], hosts[i + 1] if i + 1 < len(hosts) else None)
        torch.cuda.synchronize()
        R, t, n, st, it = unpack_record(rec.clone())
        ref = _solve(cb, b, 8, 4)
        ok = ((ref.status & 0xFF) == 0).cpu() & ((st & 0xFF) == 0)
        assert ok.float().mean() > 0.99
        assert float((R[ok] - ref.R[:, 0].cpu()[ok]).abs().max()) < 1e-7
        assert float((t[ok] - ref.t[:, 0].cpu()[ok]).abs().max()) < 1e-7


def test_host_stager_matches_direct_solve(cb):
    """HostStager (chunked H2D with the pre-pass of each slice under the next copies, through
    cvxpnpl_b200_prepass + desc.skip_prepass) gives the same result as a direct solve."""
    from cvxpnpl_b200 import synth
    for n_pts, n_lines, B in ((8, 4, 30001), (0, 6, 5000), (8, 0, 3)):
        d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=321)
        host = {k: torch.from_numpy(d[k]).pin_memory() for k in ("pts_2d", "pts_3d", "line_2d", "line_3d")}
        stager = cb.HostStager(d["K"], "cuda:0", chunks=4)
        for _ in range(2):                       # second call reuses the buffers
            res = stager.solve(host)
        torch.cuda.synchronize()
        ref = _solve(cb, d, n_pts, n_lines)
        ok = (((ref.status & 0xFF) == 0) & ((res.status & 0xFF) == 0)).cpu().numpy()
        assert ok.mean() > 0.9
        assert float((res.R[:, 0] - ref.R[:, 0]).abs().cpu()[ok].max()) < 1e-7
        assert float((res.t[:, 0] - ref.t[:, 0]).abs().cpu()[ok].max()) < 1e-7
        n_slices = len({(c * B) // 4 for c in range(5)}) - 1     # non-empty slices
        n_pre = 2 if B > 2368 else 1                              # pre-pass kernels of the tracked solver: two
        assert res.launches == ref.launches + (n_slices - 1) * n_pre   # the pre-pass once per slice instead of once


def test_host_stager_parts_matches_direct_solve(cb):
    """HostStager(parts=3): the batch as three independent sub-batch solves on their own streams (sub-batch p starts
    when its slice has arrived; its rows go back to the host under the kernels of the next ones) -- same poses as a
    direct solve, rows in the pinned host record equal to the device fields."""
    from cvxpnpl_b200 import synth
    n_pts, n_lines, B = 8, 4, 30001
    d = synth.make_batch(B, n_pts, n_lines, noise=1.0, seed=322)
    host = {k: torch.from_numpy(d[k]).pin_memory() for k in ("pts_2d", "pts_3d", "line_2d", "line_3d")}
    stager = cb.HostStager(d["K"], "cuda:0", parts=3)
    host_rec = torch.empty((B, 15), dtype=torch.float64).pin_memory()
    for _ in range(2):                           # second call reuses buffers, streams and workspaces
        host_rec.fill_(-7.0)
        res = stager.solve(host, host_record=host_rec)
        torch.cuda.synchronize()
    ref = _solve(cb, d, n_pts, n_lines)
    assert ((res.status & 0xFF) == 0).all() and ((ref.status & 0xFF) == 0).all()
    assert float((res.R[:, 0] - ref.R[:, 0]).abs().max()) < 1e-7
    assert float((res.t[:, 0] - ref.t[:, 0]).abs().max()) < 1e-7
    assert torch.equal(host_rec[:, :9], res.R[:, 0].reshape(B, 9).cpu())
    assert torch.equal(host_rec[:, 9:12], res.t[:, 0].cpu())
    assert torch.equal(host_rec[:, 14].to(torch.int32), res.iters.cpu())
    # a batch too small to split goes through the single-solve path of the same object
    small = {k: v[:100] for k, v in host.items()}
    r2 = stager.solve(small)
    torch.cuda.synchronize()
    assert float((r2.R[:, 0] - ref.R[:100, 0]).abs().max()) < 1e-7


def test_record_output_matches_fields(cb):
    """desc.record: the packed [B,15] row the finish kernel writes (R0 | t0 | n_poses | status | iters) equals the
    separately written outputs (what pack_record builds from them)."""
    from cvxpnpl_b200 import synth
    from cvxpnpl_b200.distributed import pack_record
    d = synth.make_batch(5000, 8, 4, noise=1.0, seed=8)
    d["pts_2d"][7, 0, 0] = np.nan            # one NaN problem: NaN pose, status 2
    rec = torch.empty((5000, 15), dtype=torch.float64, device="cuda")
    res = _solve(cb, d, 8, 4, record=rec)
    ref = pack_record(res.R[:, 0], res.t[:, 0], res.n_poses, res.status, res.iters)
    assert torch.equal(torch.nan_to_num(rec, nan=-7.0), torch.nan_to_num(ref, nan=-7.0))
    assert int(rec[7, 13]) & 0xFF == 2 and bool(torch.isnan(rec[7, :12]).all())


def test_singular_normal_system_raises_linalgerror(cb):
    """cvxpnpl.py:548 / 579 / 623: np.linalg.solve(N'N, N'C) raises LinAlgError on an exactly singular 3x3
    normal matrix (all bearings identical: here every pixel is the principal point of K = I, so
    N'N = n diag(1,1,0)).  Batch API: ST_SINGULAR, no pose; scalar API: the exception.  Non-finite input
    stays the NaN pose (cvxpnpl.py:493-498)."""
    from cvxpnpl_b200 import synth
    d = synth.make_batch(40, 8, 0, noise=1.0, seed=4)
    d["K"] = np.eye(3)
    d["pts_2d"][3] = 0.0
    d["pts_2d"][5, 1, 1] = np.inf
    res = _solve(cb, d, 8, 0)
    st, n = res.status.cpu().numpy() & 0xFF, res.n_poses.cpu().numpy()
    assert st[3] == 3 and n[3] == 0 and bool(torch.isnan(res.R[3]).all())
    assert st[5] == 2 and n[5] == 1
    with pytest.raises(np.linalg.LinAlgError):
        cb.pnp(d["pts_2d"][3], d["pts_3d"][3], d["K"])
    # the oracle (restated reference) raises on the same input
    from oracle import cvxpnpl_oracle as orc
    with pytest.raises(np.linalg.LinAlgError):
        orc.pnp(d["pts_2d"][3], d["pts_3d"][3], d["K"])


def test_rc_variant_lines_and_large_n(cb):
    """benchmarks/toolkit/methods/pnl.py:11-34, pnpl.py:12-46 (rc for lines / points + lines) against the
    oracle's rc variant; and variant="rc" on the large-n (stage-kernel) path, which used to drop it."""
    from cvxpnpl_b200 import synth
    from oracle import cvxpnpl_oracle as orc
    d = synth.make_batch(12, 6, 5, noise=1.0, seed=61)
    for i in range(4):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            A, Bm = orc.reduce_translation(*orc._stack(None, None, d["line_2d"][i], d["line_3d"][i], d["K"]))
            exp = orc.solve_relaxation(A, Bm, max_iters=200000, variant="rc")
            got = cb.rc_pnl(d["line_2d"][i], d["line_3d"][i], d["K"])
            assert len(got) == len(exp) == 1 and synth.rotation_angle(exp[0][0], got[0][0]) < ROT_TOL
            A, Bm = orc.reduce_translation(*orc._stack(d["pts_2d"][i], d["pts_3d"][i], d["line_2d"][i], d["line_3d"][i], d["K"]))
            exp = orc.solve_relaxation(A, Bm, max_iters=200000, variant="rc")
            got = cb.rc_pnpl(d["pts_2d"][i], d["line_2d"][i], d["pts_3d"][i], d["line_3d"][i], d["K"])
            assert len(got) == len(exp) == 1 and synth.rotation_angle(exp[0][0], got[0][0]) < ROT_TOL
            assert np.linalg.norm(exp[0][1] - got[0][1]) / np.linalg.norm(exp[0][1]) < T_TOL
    # large n: the rc and the full SDP have different optima on noisy data -> the variant must reach the stage kernel
    dl = synth.make_batch(6, 300, 0, noise=2.0, seed=62)
    full = _solve(cb, dl, 300, 0)
    rc = _solve(cb, dl, 300, 0, variant="rc")
    small = {k: (v[:, :200] if k in ("pts_2d", "pts_3d") else v) for k, v in dl.items()}
    for i in range(3):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            A, Bm = orc.reduce_translation(*orc._stack(dl["pts_2d"][i], dl["pts_3d"][i], None, None, dl["K"]))
            exp = orc.solve_relaxation(A, Bm, max_iters=200000, variant="rc")
        assert synth.rotation_angle(exp[0][0], rc.R[i, 0].cpu().numpy()) < ROT_TOL
    with pytest.raises(NotImplementedError):
        _solve(cb, dl, 300, 0, admm_dtype="f32")
    assert small["pts_2d"].shape[1] == 200 and full.R.shape == rc.R.shape


def test_null_scalar_and_plugin_vs_oracle(cb):
    """The plugin class (K-first estimate_pose, benchmarks/toolkit/methods/p*.py) and the scalar functions against
    the ORACLE (not against each other) on the harness's keyword layout (suite.py:80)."""
    from cvxpnpl_b200 import synth
    from oracle import cvxpnpl_oracle as orc
    d = synth.make_batch(6, 7, 3, noise=1.0, seed=35)
    for i in range(6):
        kw = dict(pts_2d=d["pts_2d"][i], line_2d=d["line_2d"][i], pts_3d=d["pts_3d"][i], line_3d=d["line_3d"][i])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            exp = orc.pnpl(kw["pts_2d"], kw["line_2d"], kw["pts_3d"], kw["line_3d"], d["K"], max_iters=200000)
            got = cb.CvxPnPL.estimate_pose(d["K"], **kw)
            assert len(got) == len(exp) == 1
            assert synth.rotation_angle(exp[0][0], got[0][0]) < ROT_TOL
            assert np.linalg.norm(exp[0][1] - got[0][1]) / np.linalg.norm(exp[0][1]) < T_TOL
            exp = orc.pnp(kw["pts_2d"], kw["pts_3d"], d["K"], max_iters=200000)
            got = cb.CvxPnPL.estimate_pose(d["K"], pts_2d=kw["pts_2d"], pts_3d=kw["pts_3d"])
            assert len(got) == len(exp) == 1 and synth.rotation_angle(exp[0][0], got[0][0]) < ROT_TOL
    R, t = cb.null(d["pts_2d"][0], d["pts_3d"][0], d["K"])[0]
    assert R.shape == (3, 3) and abs(np.linalg.det(R) - 1.0) < 1e-9


def test_suite_cell_matches_oracle_on_same_problems(cb):
    """one cell of the batched synthetic suite (cvxpnpl_b200/suite.py) -- the problems it
    generates ON THE DEVICE are handed to the CPU oracle one by one, and the disambiguated pose / error metric
    of the suite (suite.py:22-33, 95-108) is recomputed from the oracle's poses."""
    from cvxpnpl_b200 import suite, synth
    from oracle import cvxpnpl_oracle as orc
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    batch = suite.generate(96, 6, 0, 1.0, gen, torch.device("cuda"))
    res = cb.solve_batched(batch["K"], pts_2d=batch["pts_2d"], pts_3d=batch["pts_3d"])
    R, t = suite.disambiguate(res, batch, gen)
    ang = suite.rotation_angle_deg(batch["R_gt"], R).cpu().numpy()
    p2, p3, Kn = batch["pts_2d"].cpu().numpy(), batch["pts_3d"].cpu().numpy(), batch["K"].cpu().numpy()
    Rg = batch["R_gt"].cpu().numpy()
    for i in range(96):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            poses = orc.pnp(p2[i], p3[i], Kn, max_iters=200000)
        assert len(poses) == int(res.n_poses[i])
        if len(poses) == 1:
            assert synth.rotation_angle(poses[0][0], R[i].cpu().numpy()) < ROT_TOL
            assert abs(np.rad2deg(synth.rotation_angle(Rg[i], poses[0][0])) - ang[i]) < 1e-4


def test_two_devices_one_process(cb):
    """Kernel attributes (> 48 KB dynamic shared memory) are opted in per device: a second GPU driven from the
    same process solves too.  Skipped on a single-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from cvxpnpl_b200 import synth
    d = synth.make_batch(3000, 8, 4, noise=1.0, seed=17)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        K = torch.from_numpy(d["K"]).to(dev)
        r = cb.solve_batched(K, pts_2d=torch.from_numpy(d["pts_2d"]).to(dev), pts_3d=torch.from_numpy(d["pts_3d"]).to(dev),
                             line_2d=torch.from_numpy(d["line_2d"]).to(dev), line_3d=torch.from_numpy(d["line_3d"]).to(dev))
        torch.cuda.synchronize(dev)
        outs.append(r)
    assert ((outs[0].status & 0xFF) == 0).all() and ((outs[1].status & 0xFF) == 0).all()
    assert float((outs[0].R[:, 0].cpu() - outs[1].R[:, 0].cpu()).abs().max()) < 1e-8


def _nccl_worker_all(rank, world, port, q):
    import os
    import sys
    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
    import torch.distributed as dist
    import cvxpnpl_b200 as cb
    from cvxpnpl_b200 import synth
    from cvxpnpl_b200.distributed import solve_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    total = 20_001
    d = synth.make_batch(total, 8, 4, noise=1.0, seed=23)
    dev = torch.device("cuda", rank)
    full = {k: torch.from_numpy(d[k]).to(dev) for k in ("pts_2d", "pts_3d", "line_2d", "line_3d")}
    K = torch.from_numpy(d["K"]).to(dev)
    R, t, n, st, it = solve_sharded(K, **full)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        one = cb.solve_batched(K, **full)
        ok = (bool(((st & 0xFF) == 0).all()) and R.shape == (total, 3, 3)
              and float((R - one.R[:, 0]).abs().max()) < 1e-8 and float((t - one.t[:, 0]).abs().max()) < 1e-8)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_solve_sharded_nccl_all_gpus(cb):
    """cvxpnpl_b200.distributed.solve_sharded over NCCL: the batch sharded by contiguous ranges over every GPU of
    the box (one process per GPU), rows all-gathered in place, result equal to the single-GPU solve.  On a
    one-GPU box this still runs the NCCL path with world size 1."""
    import socket
    import torch.multiprocessing as mp
    world = max(1, min(torch.cuda.device_count(), 8))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker_all, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(r, True) for r in range(world)]


def _nccl_worker(rank, world, port, B, q):
    import os
    import sys
    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
    import torch.distributed as dist
    import cvxpnpl_b200 as cb
    from cvxpnpl_b200 import synth
    from cvxpnpl_b200.distributed import solve_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    d = synth.make_batch(B, 8, 4, noise=1.0, seed=2024)      # every rank holds the FULL batch (strong scaling)
    full = {k: torch.from_numpy(d[k]).to(dev) for k in ("pts_2d", "pts_3d", "line_2d", "line_3d")}
    K = torch.from_numpy(d["K"]).to(dev)
    R, t, n, st, it = solve_sharded(K, **full)
    torch.cuda.synchronize()
    single = cb.solve_batched(K, **full)                      # the same batch on this rank alone
    torch.cuda.synchronize()
    ang = synth.rotation_angle(single.R[:, 0].cpu().numpy(), R.cpu().numpy())
    terr = (single.t[:, 0] - t).norm(dim=1) / single.t[:, 0].norm(dim=1)
    ok = (R.shape == (B, 3, 3) and bool((n == 1).all()) and bool(((st & 0xFF) == 0).all())
          and float(np.nanmax(ang)) <= 1e-8 and float(terr.max()) <= 1e-8 and bool((it > 0).all()))
    q.put((rank, bool(ok), float(np.nanmax(ang))))
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [20001])
def test_solve_sharded_nccl(cb, B):
    """distributed.solve_sharded on real GPUs over NCCL ("rank g gets problems [gB/G, (g+1)B/G)"): two
    ranks, ragged shards (odd batch), in-place all-gather of the packed records; every rank ends up with the poses of
    the whole batch, equal to a single-GPU solve of the same batch.  Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
