"""Generates tests/golden/*.npz by running the VERBATIM reference
(/root/reference/cvxpnpl.py) in the build container, with oracle/shim/scs standing
in for the (absent) third-party SCS package.  Run from the repo root:

    ORACLE_SCS_MAX_ITERS=200000 python tests/golden/make_golden.py

The reference cannot travel to the GPU box; these small fixtures can.
What each file pins:
  units.npz     inputs/outputs of the reference's own helper functions
                (_point_constraints, _line_constraints, _vech10, _sdp_constraints,
                _re6q3, _constraint_ortho_det) -- solver independent.
  examples.npz  examples/pnp.py, pnl.py, pnpl.py inputs, ground truth (hard coded
                in the examples) and the reference's output.
  synth.npz     seeded synthetic PnP-8 / PnPL-8+4 / PnL-6 problems (noise 0,1,2 px):
                inputs, the reference's A and B, and the poses the reference
                returns (SDP solved by the shim to eps_abs 1e-9, fully converged).
  rc.npz        the "rc" ablation (benchmarks/toolkit/methods/rc.py): static data + poses.
  degenerate.npz minimal / planar configurations that take the rank-2 / rank-4
                branches: inputs, Z returned by the shim, and the reference's
                candidate poses for that Z.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
os.environ.setdefault("ORACLE_SCS_MAX_ITERS", "200000")

import cvxpnpl as ref  # noqa: E402  (the verbatim reference)
import scs  # noqa: E402  (the shim)
from cvxpnpl_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.simplefilter("ignore")


def pack_poses(poses):
    R = np.full((4, 3, 3), np.nan)
    t = np.full((4, 3), np.nan)
    for i, (Ri, ti) in enumerate(poses):
        R[i], t[i] = Ri, ti
    return R, t, len(poses)


def units():
    rng = np.random.default_rng(2024)
    K = synth.K_KINECT
    p2 = rng.random((7, 2)) * 400
    p3 = rng.random((7, 3)) - 0.5
    (C1, C2, C3), (N1, N2, N3) = ref._point_constraints(p2, p3, K)
    l2 = rng.random((5, 2, 2)) * 400
    l3 = rng.random((5, 2, 3)) - 0.5
    Cl, Nl = ref._line_constraints(l2, l3, K)
    S = rng.random((10, 10))
    S = S + S.T
    e6_in = rng.standard_normal((6, 21, 10))
    e6_out = np.array([np.stack(ref._re6q3(A)) for A in e6_in])  # (6, 3, 4)
    cod_V = np.array([np.linalg.qr(rng.standard_normal((10, 10)))[0] for _ in range(6)])
    cod_rank = np.array([2, 2, 3, 3, 4, 4])
    cod_out = np.full((6, 4, 9), np.nan)
    for i in range(6):
        r = ref._constraint_ortho_det(cod_V[i], cod_rank[i])
        cod_out[i, : len(r)] = r
    np.savez(os.path.join(OUT, "units.npz"),
             K=K, p2=p2, p3=p3, Cp=np.vstack((C1, C2, C3)), Np=np.vstack((N1, N2, N3)),
             l2=l2, l3=l3, Cl=Cl, Nl=Nl,
             S=S, vech_S_2=ref._vech10(S, 2), vech_S_1=ref._vech10(S),
             A_sdp=ref._A.toarray(), b_sdp=ref._b,
             e6_in=e6_in, e6_out=e6_out, cod_V=cod_V, cod_rank=cod_rank, cod_out=cod_out)


def examples():
    """Inputs rebuilt exactly as examples/*.py build them (same seeds, same
    expressions: pnp.py:5-26, pnl.py:5-32, pnpl.py:5-38)."""
    out = {}
    # pnp
    np.random.seed(0)
    np.random.seed(42)
    pts = 0.6 * (np.random.random((6, 3)) - 0.5)
    K = np.array([[160, 0, 320], [0, 120, 240], [0, 0, 1]])
    R_gt = np.array([[-0.48048015, 0.1391384, -0.86589799],
                     [-0.0333282, -0.98951829, -0.14050899],
                     [-0.8763721, -0.03865296, 0.48008113]])
    t_gt = np.array([-0.10266772, 0.25450789, 1.70391109])
    p2 = (pts @ R_gt.T + t_gt) @ K.T
    p2 = (p2 / p2[:, -1, None])[:, :-1]
    R, t, n = pack_poses(ref.pnp(pts_2d=p2, pts_3d=pts, K=K))
    out.update(pnp_K=K.astype(float), pnp_pts_2d=p2, pnp_pts_3d=pts, pnp_R_gt=R_gt, pnp_t_gt=t_gt,
               pnp_R=R, pnp_t=t, pnp_n=n)
    # pnl
    np.random.seed(0)
    np.random.seed(42)
    line_3d = 0.6 * (np.random.random((6, 2, 3)) - 0.5)
    R_gt = np.array([[0.89802142, -0.41500101, 0.14605372],
                     [0.24509948, 0.7476071, 0.61725997],
                     [-0.36535431, -0.51851499, 0.77308372]])
    t_gt = np.array([-0.0767557, 0.13917375, 1.9708239])
    pl = line_3d.reshape((-1, 3))
    l2 = (pl @ R_gt.T + t_gt) @ K.T
    l2 = (l2 / l2[:, -1, None])[:, :-1].reshape((-1, 2, 2))
    R, t, n = pack_poses(ref.pnl(line_2d=l2, line_3d=line_3d, K=K))
    out.update(pnl_K=K.astype(float), pnl_line_2d=l2, pnl_line_3d=line_3d, pnl_R_gt=R_gt,
               pnl_t_gt=t_gt, pnl_R=R, pnl_t=t, pnl_n=n)
    # pnpl
    np.random.seed(0)
    np.random.seed(42)
    pts = 0.6 * (np.random.random((4, 3)) - 0.5)
    line_3d = 0.6 * (np.random.random((4, 2, 3)) - 0.5)
    pall = np.vstack((pts, line_3d.reshape((-1, 3))))
    a2 = (pall @ R_gt.T + t_gt) @ K.T
    a2 = (a2 / a2[:, -1, None])[:, :-1]
    p2, l2 = a2[:4], a2[4:].reshape((-1, 2, 2))
    R, t, n = pack_poses(ref.pnpl(pts_2d=p2, line_2d=l2, pts_3d=pts, line_3d=line_3d, K=K))
    out.update(pnpl_K=K.astype(float), pnpl_pts_2d=p2, pnpl_line_2d=l2, pnpl_pts_3d=pts,
               pnpl_line_3d=line_3d, pnpl_R_gt=R_gt, pnpl_t_gt=t_gt, pnpl_R=R, pnpl_t=t, pnpl_n=n)
    np.savez(os.path.join(OUT, "examples.npz"), **out)


def _ref_AB(d, i, n_pts, n_lines):
    Cs, Ns = [], []
    if n_pts:
        (C1, C2, C3), (N1, N2, N3) = ref._point_constraints(d["pts_2d"][i], d["pts_3d"][i], d["K"])
        Cs += [C1, C2, C3]
        Ns += [N1, N2, N3]
    if n_lines:
        Cl, Nl = ref._line_constraints(d["line_2d"][i], d["line_3d"][i], d["K"])
        Cs.append(Cl)
        Ns.append(Nl)
    C, N = np.vstack(Cs), np.vstack(Ns)
    B = np.linalg.solve(N.T @ N, N.T @ C)
    return C - N @ B, B


def _call(d, i, n_pts, n_lines):
    if n_pts and n_lines:
        return ref.pnpl(d["pts_2d"][i], d["line_2d"][i], d["pts_3d"][i], d["line_3d"][i], d["K"])
    if n_pts:
        return ref.pnp(d["pts_2d"][i], d["pts_3d"][i], d["K"])
    return ref.pnl(d["line_2d"][i], d["line_3d"][i], d["K"])


def synth_set():
    out = {}
    for name, (n_pts, n_lines) in {"pnp8": (8, 0), "pnpl8_4": (8, 4), "pnl6": (0, 6)}.items():
        for noise in (0, 1, 2):
            B = 8
            d = synth.make_batch(B, n_pts, n_lines, noise=float(noise), seed=1000 + noise)
            key = f"{name}_s{noise}"
            Rs, ts, ns, Qs, Bs = [], [], [], [], []
            for i in range(B):
                A, Bm = _ref_AB(d, i, n_pts, n_lines)
                R, t, n = pack_poses(_call(d, i, n_pts, n_lines))
                Rs.append(R), ts.append(t), ns.append(n), Qs.append(A.T @ A), Bs.append(Bm)
            for k in ("pts_2d", "pts_3d", "line_2d", "line_3d", "K", "R_gt", "t_gt"):
                out[f"{key}_{k}"] = d[k]
            out[f"{key}_R"], out[f"{key}_t"], out[f"{key}_n"] = np.array(Rs), np.array(ts), np.array(ns)
            out[f"{key}_AtA"], out[f"{key}_B"] = np.array(Qs), np.array(Bs)
    np.savez(os.path.join(OUT, "synth.npz"), **out)


def degenerate_set():
    """Rank-2 / rank-4 branches.  For each problem the shim's Z is stored with the
    reference's candidate poses *for that Z*, so the extraction restatement can be
    replayed without depending on which point of a non-unique optimal face a
    solver lands on."""
    out = {}
    cfgs = {"pts4": (4, 0, False), "pts3": (3, 0, False), "lines3": (0, 3, False),
            "lines4": (0, 4, False), "p2l1": (2, 1, False), "coplanar8": (8, 0, True)}
    captured = {}
    real_solve = scs.solve

    def spy(data, cone, **kw):
        res = real_solve(data, cone, **kw)
        captured["x"] = res["x"].copy()
        captured["dobj"] = res["info"]["dobj"]
        return res

    scs.solve = spy
    ref.scs.solve = spy
    for name, (n_pts, n_lines, cop) in cfgs.items():
        B = 6
        d = synth.make_batch(B, n_pts, n_lines, noise=0.0, seed=77, coplanar=cop)
        Rs, ts, ns, Zs, As, Bs, ok = [], [], [], [], [], [], []
        for i in range(B):
            A, Bm = _ref_AB(d, i, n_pts, n_lines)
            try:
                R, t, n = pack_poses(_call(d, i, n_pts, n_lines))
                ok.append(1)
            except np.linalg.LinAlgError:
                R, t, n = np.full((4, 3, 3), np.nan), np.full((4, 3), np.nan), 0
                ok.append(0)
            Rs.append(R), ts.append(t), ns.append(n)
            Zs.append(ref._vech10_inv(captured["x"]))
            As.append(A.T @ A), Bs.append(Bm)
        for k in ("pts_2d", "pts_3d", "line_2d", "line_3d", "K", "R_gt", "t_gt"):
            out[f"{name}_{k}"] = d[k]
        out[f"{name}_R"], out[f"{name}_t"], out[f"{name}_n"] = np.array(Rs), np.array(ts), np.array(ns)
        out[f"{name}_Z"], out[f"{name}_AtA"], out[f"{name}_B"] = np.array(Zs), np.array(As), np.array(Bs)
        out[f"{name}_ok"] = np.array(ok)
        print(name, "n_poses", ns, "ok", ok)
    scs.solve = real_solve
    ref.scs.solve = real_solve
    np.savez(os.path.join(OUT, "degenerate.npz"), **out)


def rc_set():
    """benchmarks/toolkit/methods/rc.py (the 16-equality ablation) run verbatim: its static
    data and its poses on seeded PnP-8 problems."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_rc", "/root/reference/benchmarks/toolkit/methods/rc.py")
    rcm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rcm)
    d = synth.make_batch(6, 8, 0, noise=1.0, seed=2100)
    Rs, ts, ns = [], [], []
    for i in range(6):
        A, Bm = _ref_AB(d, i, 8, 0)
        R, t, n = pack_poses(rcm._solve_relaxation_rc(A, Bm))
        Rs.append(R), ts.append(t), ns.append(n)
    np.savez(os.path.join(OUT, "rc.npz"), A_rc=rcm._A_rc.toarray(), b_rc=rcm._b_rc, K=d["K"], pts_2d=d["pts_2d"],
             pts_3d=d["pts_3d"], R=np.array(Rs), t=np.array(ts), n=np.array(ns))


if __name__ == "__main__":
    rc_set()
    units()
    examples()
    synth_set()
    degenerate_set()
    print("golden fixtures written to", OUT)
