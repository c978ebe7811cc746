"""CPU restatement (numpy) of the reference's pose -> BVH-channel post-step:
ZEGGS/generate.py:389-406 -> anim/txform.py:23-34 (xform_orthogonalize_from_xy), anim/quat.py:166-206 (from_xform),
ZEGGS/utils.py:47-87 (write_bvh: re-base the root to start_position / start_rotation, fold the root transform into joint 0),
anim/quat.py:111-127 (to_euler 'zyx') and np.degrees.

TEST INFRASTRUCTURE ONLY.  Pinned: tests/test_oracle_vs_reference.py runs it against the imported reference functions, and
tests/golden/pose_post.npz holds outputs of the reference itself (oracle/make_golden.py)."""
import numpy as np


def orthogonalize_from_xy(xy, eps=1e-10):
    """txform.py:23-34 in float32: xy [..., 2, 3] -> rotation matrices [..., 3, 3] whose COLUMNS are the x, y, z axes."""
    xy = np.asarray(xy, dtype=np.float32)
    x = xy[..., 0, :]
    z = np.cross(x, xy[..., 1, :])
    y = np.cross(z, x)
    n = lambda v: v / (np.linalg.norm(v, axis=-1, keepdims=True).astype(np.float32) + np.float32(eps))
    return np.stack([n(x), n(y), n(z)], axis=-1).astype(np.float32)


def quat_from_xform(ts, eps=1e-10):
    """quat.py:166-206: w-first quaternion of a rotation matrix, four-branch form, float32."""
    ts = np.asarray(ts, dtype=np.float32)
    eps = np.float32(eps)
    m = lambda i, j: ts[..., i, j]
    t = m(0, 0) + m(1, 1) + m(2, 2)
    out = np.zeros(ts.shape[:-2] + (4,), dtype=np.float32)
    s = np.float32(0.5) / np.sqrt(np.maximum(t + 1, eps))
    q0 = np.stack([np.float32(0.25) / s, s * (m(2, 1) - m(1, 2)), s * (m(0, 2) - m(2, 0)), s * (m(1, 0) - m(0, 1))], -1)
    c0 = (m(0, 0) > m(1, 1)) & (m(0, 0) > m(2, 2))
    s0 = np.float32(2.0) * np.sqrt(np.maximum(1 + m(0, 0) - m(1, 1) - m(2, 2), eps))
    q1 = np.stack([(m(2, 1) - m(1, 2)) / s0, s0 * np.float32(0.25), (m(0, 1) + m(1, 0)) / s0, (m(0, 2) + m(2, 0)) / s0], -1)
    c1 = (~c0) & (m(1, 1) > m(2, 2))
    s1 = np.float32(2.0) * np.sqrt(np.maximum(1 + m(1, 1) - m(0, 0) - m(2, 2), eps))
    q2 = np.stack([(m(0, 2) - m(2, 0)) / s1, (m(0, 1) + m(1, 0)) / s1, s1 * np.float32(0.25), (m(1, 2) + m(2, 1)) / s1], -1)
    s2 = np.float32(2.0) * np.sqrt(np.maximum(1 + m(2, 2) - m(0, 0) - m(1, 1), eps))
    q3 = np.stack([(m(1, 0) - m(0, 1)) / s2, (m(0, 2) + m(2, 0)) / s2, (m(1, 2) + m(2, 1)) / s2, s2 * np.float32(0.25)], -1)
    pos = (t > 0)[..., None]
    out = np.where(pos, q0, np.where(c0[..., None], q1, np.where(c1[..., None], q2, q3)))
    return out.astype(np.float32)


def quat_mul(x, y):
    """quat.py:17-25."""
    x0, x1, x2, x3 = x[..., 0:1], x[..., 1:2], x[..., 2:3], x[..., 3:4]
    y0, y1, y2, y3 = y[..., 0:1], y[..., 1:2], y[..., 2:3], y[..., 3:4]
    return np.concatenate([y0 * x0 - y1 * x1 - y2 * x2 - y3 * x3, y0 * x1 + y1 * x0 - y2 * x3 + y3 * x2,
                           y0 * x2 + y1 * x3 + y2 * x0 - y3 * x1, y0 * x3 - y1 * x2 + y2 * x1 + y3 * x0], axis=-1)


def quat_mul_vec(q, v):
    """quat.py:36-38 (the cross products land in float64 there: np.empty default dtype)."""
    u = q[..., 1:].astype(np.float64)
    t = 2.0 * np.cross(u, v.astype(np.float64))
    return v + q[..., 0][..., None] * t + np.cross(u, t)


def quat_inv(q):
    return np.array([1, -1, -1, -1], dtype=np.float32) * q


def to_euler_zyx(x):
    """quat.py:111-119."""
    x0, x1, x2, x3 = x[..., 0:1], x[..., 1:2], x[..., 2:3], x[..., 3:4]
    return np.concatenate([np.arctan2(2.0 * (x0 * x3 + x1 * x2), 1.0 - 2.0 * (x2 * x2 + x3 * x3)),
                           np.arcsin(np.clip(2.0 * (x0 * x2 - x3 * x1), -1.0, 1.0)),
                           np.arctan2(2.0 * (x0 * x1 + x2 * x3), 1.0 - 2.0 * (x1 * x1 + x2 * x2))], axis=-1)


def pose_to_bvh_channels(root_pos, root_rot, lpos, ltxy, start_position=(0, 0, 0), start_rotation=(1, 0, 0, 0)):
    """One clip: root_pos [T,3], root_rot [T,4], lpos [T,J,3], ltxy [T,J,2,3] -> (positions [T,J,3], euler degrees [T,J,3])
    exactly as generate.py:389-406 + utils.write_bvh hand them to bvh.save."""
    lrot = quat_from_xform(orthogonalize_from_xy(ltxy))
    sp = np.asarray(start_position, dtype=np.float64)
    sr = np.asarray(start_rotation, dtype=np.float32)
    off_p, off_r = root_pos[0:1].copy(), root_rot[0:1].copy()
    rp = quat_mul_vec(quat_inv(off_r), root_pos - off_p)
    rr = quat_mul(quat_inv(off_r), root_rot)
    rp = quat_mul_vec(sr[None], rp) + sp[None]
    rr = quat_mul(sr[None], rr)
    pos = lpos.copy()                    # float32 like V_lpos: the float64 root composition is rounded on assignment
    rot = lrot.copy()
    pos[:, 0] = quat_mul_vec(rr, lpos[:, 0]) + rp
    rot[:, 0] = quat_mul(rr, lrot[:, 0])
    return pos, np.degrees(to_euler_zyx(rot))
