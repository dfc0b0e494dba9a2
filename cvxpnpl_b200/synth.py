"""Synthetic PnP / PnL / PnPL workloads (numpy, host side).

Restates the generators of the reference's benchmark toolkit so that bench.py,
the tests and the CPU baseline all draw the same kind of problem:
  * random_pose            benchmarks/toolkit/suites/synth.py:13-42
  * Kinect-v1 intrinsics   benchmarks/toolkit/suites/synth.py:49-51
  * correspondences        benchmarks/toolkit/suites/synth.py:276-346
  * pinhole projection     benchmarks/toolkit/suites/suite.py:17-19
  * pose error             benchmarks/toolkit/suites/suite.py:8-33
The reference draws one problem at a time from numpy's global RNG; here a whole
batch is drawn from a seeded numpy Generator (the global-RNG stream itself is not
reproduced).  BASELINE.json fixes the point/line split (8 + 4) where the
reference's PnPLSynth draws it at random (synth.py:323-324).
"""
import numpy as np

#: Kinect v1 intrinsics (synth.py:49-51)
K_KINECT = np.array([[572.41140, 0.0, 325.26110], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]])
#: side of the cube the 3D points are drawn from (synth.py:55)
LENGTH = 0.6


def random_poses(rng, B):
    """B random poses: axis ~ normalised U(-.5,.5)^3, angle ~ U(0, 2 pi), Rodrigues;
    t = (U(-.5,.5), U(-.5,.5), U(.6, 2.2)).  synth.py:13-42."""
    axis = rng.random((B, 3)) - 0.5
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    ang = 2 * np.pi * rng.random(B)
    Kx = np.zeros((B, 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2] = -axis[:, 2], axis[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = axis[:, 2], -axis[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -axis[:, 1], axis[:, 0]
    R = (np.eye(3) + np.sin(ang)[:, None, None] * Kx
         + (1 - np.cos(ang))[:, None, None] * (Kx @ Kx))
    t = np.concatenate([rng.random((B, 2)) - 0.5, 1.6 * rng.random((B, 1)) + 0.6], axis=1)
    return R, t


def project_points(pts, K, R, t):
    """Batched pinhole projection, suite.py:17-19.  pts (B,n,3) -> (B,n,2)."""
    pc = (pts @ np.swapaxes(R, -1, -2) + t[..., None, :]) @ K.T
    return pc[..., :2] / pc[..., 2:]


def make_batch(B, n_pts, n_lines, noise=0.0, seed=42, K=K_KINECT, coplanar=False):
    """A batch of B problems with n_pts points and n_lines lines each.

    Returns a dict with pts_2d (B,n,2), pts_3d (B,n,3), line_2d (B,m,2,2),
    line_3d (B,m,2,3), K (3,3) and the ground truth R_gt (B,3,3), t_gt (B,3).
    Lines are consecutive point pairs (synth.py:296-309); pixel noise is
    N(0, noise^2) on every projected point (synth.py:283)."""
    rng = np.random.default_rng(seed)
    R, t = random_poses(rng, B)
    P = LENGTH * (rng.random((B, n_pts + 2 * n_lines, 3)) - 0.5)
    if coplanar:
        P[..., 2] = 0.0
    p2 = project_points(P, K, R, t)
    if noise > 0:
        p2 = p2 + rng.normal(scale=noise, size=p2.shape)
    return {
        "pts_2d": np.ascontiguousarray(p2[:, :n_pts]),
        "pts_3d": np.ascontiguousarray(P[:, :n_pts]),
        "line_2d": np.ascontiguousarray(p2[:, n_pts:].reshape(B, n_lines, 2, 2)),
        "line_3d": np.ascontiguousarray(P[:, n_pts:].reshape(B, n_lines, 2, 3)),
        "K": np.array(K, dtype=np.float64),
        "R_gt": R,
        "t_gt": t,
    }


def rotation_angle(Ra, Rb):
    """Geodesic angle (rad) between rotations, batched.  Uses
    atan2(|skew|, trace - 1) rather than the reference's arccos (suite.py:14),
    whose noise floor near identity is ~2e-8 rad."""
    E = np.swapaxes(Ra, -1, -2) @ Rb
    sk = np.stack([E[..., 2, 1] - E[..., 1, 2], E[..., 0, 2] - E[..., 2, 0],
                   E[..., 1, 0] - E[..., 0, 1]], axis=-1)
    s = np.linalg.norm(sk, axis=-1)
    c = np.trace(E, axis1=-2, axis2=-1) - 1.0
    return np.arctan2(s, c)


def pose_error(R_gt, t_gt, R, t):
    """(angle [rad], |t - t_gt| / |t_gt|), the metric of suite.py:22-33 (which
    reports the angle in degrees)."""
    return rotation_angle(R_gt, R), (np.linalg.norm(t - t_gt, axis=-1)
                                     / np.linalg.norm(t_gt, axis=-1))
