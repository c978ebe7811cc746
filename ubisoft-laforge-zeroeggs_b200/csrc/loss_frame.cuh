// Per-frame math of the training loss (ZEGGS/train.py:277-421): local->world transforms, 75-joint forward
// kinematics with velocities (anim/txform.py:10-34) and their adjoints.  All functions are
// __host__ __device__ and work on structure-of-arrays buffers `p[ch * stride + idx]` so that the CUDA kernels
// (one thread per frame, lanes = consecutive frames -> coalesced) and the host-side unit harness share
// one implementation.
#pragma once
#include "common.cuh"

namespace zeggs {

// channel map of the world-space buffer Q (per frame), in the order the loss terms consume it
constexpr int Q_ROOT_POS = 0;                 // 3
constexpr int Q_ROOT_MAT = 3;                 // 9   quat_to_xform(root_rot)
constexpr int Q_ROOT_VEL = 12;                // 3   world root velocity (train.py:281-286)
constexpr int Q_ROOT_VRT = 15;                // 3
constexpr int Q_LPOS = 18;                    // 225 local positions, joint 0 in world space (train.py:296-305)
constexpr int Q_LTXY = Q_LPOS + NJ * 3;       // 450 raw two-axis rotations
constexpr int Q_LVEL = Q_LTXY + NJ * 6;       // 225
constexpr int Q_LVRT = Q_LVEL + NJ * 3;       // 225
constexpr int Q_CPOS = Q_LVRT + NJ * 3;       // 225 FK global positions
constexpr int Q_CMAT = Q_CPOS + NJ * 3;       // 675 FK global rotations
constexpr int Q_CVEL = Q_CMAT + NJ * 9;       // 225
constexpr int Q_CVRT = Q_CVEL + NJ * 3;       // 225
constexpr int Q_GAZE = Q_CVRT + NJ * 3;       // 3
constexpr int Q_CH = Q_GAZE + 3;              // 2496

struct M3 { float m[9]; };   // row-major

__host__ __device__ inline V3 mv(const M3& a, V3 v) {
  return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
__host__ __device__ inline V3 mtv(const M3& a, V3 v) {   // a^T v
  return v3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z, a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}
__host__ __device__ inline M3 mm(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
  return r;
}
__host__ __device__ inline M3 mtm(const M3& a, const M3& b) {  // a^T b
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[i] * b.m[j] + a.m[3 + i] * b.m[3 + j] + a.m[6 + i] * b.m[6 + j];
  return r;
}
__host__ __device__ inline M3 mmt(const M3& a, const M3& b) {  // a b^T
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[i * 3] * b.m[j * 3] + a.m[i * 3 + 1] * b.m[j * 3 + 1] + a.m[i * 3 + 2] * b.m[j * 3 + 2];
  return r;
}
__host__ __device__ inline void add_outer(M3& d, V3 a, V3 b) {   // d += a b^T
  d.m[0] += a.x * b.x; d.m[1] += a.x * b.y; d.m[2] += a.x * b.z;
  d.m[3] += a.y * b.x; d.m[4] += a.y * b.y; d.m[5] += a.y * b.z;
  d.m[6] += a.z * b.x; d.m[7] += a.z * b.y; d.m[8] += a.z * b.z;
}
__host__ __device__ inline M3 quat_to_xform(Q4 q) {  // tquat.py:53-67
  float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
  float xx = q.x * x2, yy = q.y * y2, wx = q.w * x2;
  float xy = q.x * y2, yz = q.y * z2, wy = q.w * y2;
  float xz = q.x * z2, zz = q.z * z2, wz = q.w * z2;
  M3 r;
  r.m[0] = 1.0f - (yy + zz); r.m[1] = xy - wz; r.m[2] = xz + wy;
  r.m[3] = xy + wz; r.m[4] = 1.0f - (xx + zz); r.m[5] = yz - wx;
  r.m[6] = xz - wy; r.m[7] = yz + wx; r.m[8] = 1.0f - (xx + yy);
  return r;
}
__host__ __device__ inline Q4 quat_to_xform_bwd(Q4 q, const M3& d) {
  Q4 g;
  g.w = 2.0f * (-q.z * d.m[1] + q.y * d.m[2] + q.z * d.m[3] - q.x * d.m[5] - q.y * d.m[6] + q.x * d.m[7]);
  g.x = 2.0f * (q.y * d.m[1] + q.z * d.m[2] + q.y * d.m[3] - 2.0f * q.x * d.m[4] - q.w * d.m[5] + q.z * d.m[6] + q.w * d.m[7] - 2.0f * q.x * d.m[8]);
  g.y = 2.0f * (-2.0f * q.y * d.m[0] + q.x * d.m[1] + q.w * d.m[2] + q.x * d.m[3] + q.z * d.m[5] - q.w * d.m[6] + q.z * d.m[7] - 2.0f * q.y * d.m[8]);
  g.z = 2.0f * (-2.0f * q.z * d.m[0] - q.w * d.m[1] + q.x * d.m[2] + q.w * d.m[3] - 2.0f * q.z * d.m[4] + q.y * d.m[5] + q.x * d.m[6] + q.y * d.m[7]);
  return g;
}
// u = v / (|v| + eps) and its adjoint
__host__ __device__ inline V3 unit_eps(V3 v, float eps) { float n = sqrtf(dot(v, v)); return (1.0f / (n + eps)) * v; }
__host__ __device__ inline V3 unit_eps_bwd(V3 v, float eps, V3 du) {
  float n = sqrtf(dot(v, v));
  V3 r = (1.0f / (n + eps)) * du;
  if (n > 0.0f) r = r - (dot(du, v) / ((n + eps) * (n + eps) * n)) * v;
  return r;
}
// txform.py:23-34: columns of the result are the normalised x, y = z cross x, z = x cross xy[1] axes
__host__ __device__ inline M3 orthogonalize_xy(V3 x, V3 yv) {
  V3 z = cross(x, yv), y = cross(z, x);
  V3 X = unit_eps(x, 1e-10f), Y = unit_eps(y, 1e-10f), Z = unit_eps(z, 1e-10f);
  M3 r;
  r.m[0] = X.x; r.m[1] = Y.x; r.m[2] = Z.x;
  r.m[3] = X.y; r.m[4] = Y.y; r.m[5] = Z.y;
  r.m[6] = X.z; r.m[7] = Y.z; r.m[8] = Z.z;
  return r;
}
__host__ __device__ inline void orthogonalize_xy_bwd(V3 x, V3 yv, const M3& d, V3& dx, V3& dyv) {
  V3 z = cross(x, yv), y = cross(z, x);
  V3 dX = v3(d.m[0], d.m[3], d.m[6]), dY = v3(d.m[1], d.m[4], d.m[7]), dZ = v3(d.m[2], d.m[5], d.m[8]);
  dx = unit_eps_bwd(x, 1e-10f, dX);
  V3 dy = unit_eps_bwd(y, 1e-10f, dY);
  V3 dz = unit_eps_bwd(z, 1e-10f, dZ);
  // y = z x x
  dz = dz + cross(x, dy);
  dx = dx + cross(dy, z);
  // z = x x yv
  dx = dx + cross(yv, dz);
  dyv = cross(dz, x);
}

__host__ __device__ inline V3 ld3(const float* p, size_t stride, size_t idx, int ch) {
  return v3(p[(size_t)ch * stride + idx], p[(size_t)(ch + 1) * stride + idx], p[(size_t)(ch + 2) * stride + idx]);
}
__host__ __device__ inline void st3(float* p, size_t stride, size_t idx, int ch, V3 v) {
  p[(size_t)ch * stride + idx] = v.x; p[(size_t)(ch + 1) * stride + idx] = v.y; p[(size_t)(ch + 2) * stride + idx] = v.z;
}
__host__ __device__ inline M3 ldm(const float* p, size_t stride, size_t idx, int ch) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 9; ++i) r.m[i] = p[(size_t)(ch + i) * stride + idx];
  return r;
}
__host__ __device__ inline void stm(float* p, size_t stride, size_t idx, int ch, const M3& v) {
#pragma unroll
  for (int i = 0; i < 9; ++i) p[(size_t)(ch + i) * stride + idx] = v.m[i];
}

// ------------------------------------------------------------------ forward: world-space quantities of one frame
// Ys: pose vector in SoA [1131][stride]; q / qp: root rotation of this frame / of the previous frame (= q for t = 0);
// pos: root position; gaze: gaze target.  Writes Q[Q_CH][stride] at column idx.
// Ys / Q never alias (separate workspace buffers): __restrict__ lets the compiler batch a joint's 15 local loads with the
// parent's 18 instead of ordering every load behind the previous joint's stores
__host__ __device__ inline void loss_frame_forward(const float* __restrict__ Ys, size_t stride, size_t idx, Q4 q, Q4 qp, V3 pos, V3 gaze,
                                                   const int* __restrict__ parents, float* __restrict__ Q) {
  st3(Q, stride, idx, Q_ROOT_POS, pos);
  const M3 R = quat_to_xform(q);
  stm(Q, stride, idx, Q_ROOT_MAT, R);
  const V3 velw = quat_mul_vec(qp, ld3(Ys, stride, idx, 0));          // train.py:281-286
  const V3 vrtw = quat_mul_vec(qp, ld3(Ys, stride, idx, 3));
  st3(Q, stride, idx, Q_ROOT_VEL, velw);
  st3(Q, stride, idx, Q_ROOT_VRT, vrtw);
  // gaze (train.py:336-337): R(q)^-1 normalize(gaze - pos), eps 1e-8
  st3(Q, stride, idx, Q_GAZE, quat_mul_vec(quat_inv(q), unit_eps(gaze - pos, 1e-8f)));
  // joint 0 to world space (train.py:296-303)
  {
    V3 lp = ld3(Ys, stride, idx, OFF_LPOS), lv = ld3(Ys, stride, idx, OFF_LVEL), lr = ld3(Ys, stride, idx, OFF_LVRT);
    V3 x = ld3(Ys, stride, idx, OFF_LTXY), yv = ld3(Ys, stride, idx, OFF_LTXY + 3);
    V3 rp0 = quat_mul_vec(q, lp);
    V3 p0 = rp0 + pos;
    M3 m0 = mm(R, orthogonalize_xy(x, yv));
    V3 v0 = velw + quat_mul_vec(q, lv) + cross(vrtw, rp0);
    V3 t0 = vrtw + quat_mul_vec(q, lr);
    st3(Q, stride, idx, Q_LPOS, p0); st3(Q, stride, idx, Q_LVEL, v0); st3(Q, stride, idx, Q_LVRT, t0);
    st3(Q, stride, idx, Q_CPOS, p0); stm(Q, stride, idx, Q_CMAT, m0); st3(Q, stride, idx, Q_CVEL, v0); st3(Q, stride, idx, Q_CVRT, t0);
  }
  for (int c = 0; c < NJ * 6; ++c) Q[(size_t)(Q_LTXY + c) * stride + idx] = Ys[(size_t)(OFF_LTXY + c) * stride + idx];
  // FK (txform.py:10-20)
  for (int i = 1; i < NJ; ++i) {
    const int p = parents[i];
    V3 lp = ld3(Ys, stride, idx, OFF_LPOS + 3 * i), lv = ld3(Ys, stride, idx, OFF_LVEL + 3 * i), lr = ld3(Ys, stride, idx, OFF_LVRT + 3 * i);
    st3(Q, stride, idx, Q_LPOS + 3 * i, lp); st3(Q, stride, idx, Q_LVEL + 3 * i, lv); st3(Q, stride, idx, Q_LVRT + 3 * i, lr);
    M3 lm = orthogonalize_xy(ld3(Ys, stride, idx, OFF_LTXY + 6 * i), ld3(Ys, stride, idx, OFF_LTXY + 6 * i + 3));
    M3 grp = ldm(Q, stride, idx, Q_CMAT + 9 * p);
    V3 gpp = ld3(Q, stride, idx, Q_CPOS + 3 * p), gtp = ld3(Q, stride, idx, Q_CVRT + 3 * p), gvp = ld3(Q, stride, idx, Q_CVEL + 3 * p);
    V3 rp = mv(grp, lp);
    st3(Q, stride, idx, Q_CPOS + 3 * i, gpp + rp);
    stm(Q, stride, idx, Q_CMAT + 9 * i, mm(grp, lm));
    st3(Q, stride, idx, Q_CVRT + 3 * i, gtp + mv(grp, lr));
    st3(Q, stride, idx, Q_CVEL + 3 * i, gvp + mv(grp, lv) + cross(gtp, rp));
  }
}

// ------------------------------------------------------------------ backward of one frame
// Q: forward world-space values of this frame (the "O" side); G: dLoss/dQ on entry (FK channels are used as
// accumulators and clobbered).  Outputs: gY[1131][stride] (pose-vector gradient, SoA), *dpos, *dq (this frame's root
// rotation), *dqp (previous frame's root rotation).
__host__ __device__ inline void loss_frame_backward(const float* __restrict__ Ys, const float* __restrict__ Q, float* __restrict__ G,
                                                    size_t stride, size_t idx, Q4 q, Q4 qp, V3 pos, V3 gaze,
                                                    const int* __restrict__ parents, float* __restrict__ gY, V3* dpos_out, Q4* dq_out,
                                                    Q4* dqp_out) {
  // reverse FK: children push into their parents' accumulators
  for (int i = NJ - 1; i >= 1; --i) {
    const int p = parents[i];
    V3 lp = ld3(Ys, stride, idx, OFF_LPOS + 3 * i), lv = ld3(Ys, stride, idx, OFF_LVEL + 3 * i), lr = ld3(Ys, stride, idx, OFF_LVRT + 3 * i);
    V3 x = ld3(Ys, stride, idx, OFF_LTXY + 6 * i), yv = ld3(Ys, stride, idx, OFF_LTXY + 6 * i + 3);
    M3 lm = orthogonalize_xy(x, yv);
    M3 grp = ldm(Q, stride, idx, Q_CMAT + 9 * p);
    V3 gtp = ld3(Q, stride, idx, Q_CVRT + 3 * p);
    V3 rp = mv(grp, lp);
    V3 dgp = ld3(G, stride, idx, Q_CPOS + 3 * i), dgt = ld3(G, stride, idx, Q_CVRT + 3 * i), dgv = ld3(G, stride, idx, Q_CVEL + 3 * i);
    M3 dgr = ldm(G, stride, idx, Q_CMAT + 9 * i);
    M3 acc = ldm(G, stride, idx, Q_CMAT + 9 * p);
    V3 drp = dgp + cross(dgv, gtp);                       // gp[i] = gp[p] + rp ;  gv[i] += gt[p] x rp
    add_outer(acc, dgv, lv);                               // gv[i] += gr[p] lvel
    add_outer(acc, dgt, lr);                               // gt[i]  = gt[p] + gr[p] lvrt
    add_outer(acc, drp, lp);                               // rp = gr[p] lpos
    M3 t = mmt(dgr, lm);                                   // gr[i] = gr[p] lmat
#pragma unroll
    for (int k = 0; k < 9; ++k) acc.m[k] += t.m[k];
    stm(G, stride, idx, Q_CMAT + 9 * p, acc);
    st3(G, stride, idx, Q_CPOS + 3 * p, ld3(G, stride, idx, Q_CPOS + 3 * p) + dgp);
    st3(G, stride, idx, Q_CVRT + 3 * p, ld3(G, stride, idx, Q_CVRT + 3 * p) + dgt + cross(rp, dgv));
    st3(G, stride, idx, Q_CVEL + 3 * p, ld3(G, stride, idx, Q_CVEL + 3 * p) + dgv);
    // local quantities of joint i: direct L1 terms + FK
    st3(gY, stride, idx, OFF_LPOS + 3 * i, ld3(G, stride, idx, Q_LPOS + 3 * i) + mtv(grp, drp));
    st3(gY, stride, idx, OFF_LVEL + 3 * i, ld3(G, stride, idx, Q_LVEL + 3 * i) + mtv(grp, dgv));
    st3(gY, stride, idx, OFF_LVRT + 3 * i, ld3(G, stride, idx, Q_LVRT + 3 * i) + mtv(grp, dgt));
    V3 dx, dyv;
    orthogonalize_xy_bwd(x, yv, mtm(grp, dgr), dx, dyv);
    st3(gY, stride, idx, OFF_LTXY + 6 * i, ld3(G, stride, idx, Q_LTXY + 6 * i) + dx);
    st3(gY, stride, idx, OFF_LTXY + 6 * i + 3, ld3(G, stride, idx, Q_LTXY + 6 * i + 3) + dyv);
  }
  // joint 0 + root terms
  const M3 R = quat_to_xform(q);
  V3 lp = ld3(Ys, stride, idx, OFF_LPOS), lv = ld3(Ys, stride, idx, OFF_LVEL), lr = ld3(Ys, stride, idx, OFF_LVRT);
  V3 x = ld3(Ys, stride, idx, OFF_LTXY), yv = ld3(Ys, stride, idx, OFF_LTXY + 3);
  V3 vel = ld3(Ys, stride, idx, 0), vrt = ld3(Ys, stride, idx, 3);
  V3 vrtw = quat_mul_vec(qp, vrt);
  V3 rp0 = quat_mul_vec(q, lp);
  // joint 0 receives the local (Q_L*) and the FK-root (Q_C*) gradients
  V3 dp0 = ld3(G, stride, idx, Q_LPOS) + ld3(G, stride, idx, Q_CPOS);
  V3 dv0 = ld3(G, stride, idx, Q_LVEL) + ld3(G, stride, idx, Q_CVEL);
  V3 dt0 = ld3(G, stride, idx, Q_LVRT) + ld3(G, stride, idx, Q_CVRT);
  M3 dm0 = ldm(G, stride, idx, Q_CMAT);
  M3 dR = ldm(G, stride, idx, Q_ROOT_MAT);
  V3 dpos = ld3(G, stride, idx, Q_ROOT_POS) + dp0;                 // p0 = rp0 + pos
  V3 dvelw = ld3(G, stride, idx, Q_ROOT_VEL) + dv0;                // v0 = velw + rot(q, lv) + vrtw x rp0
  V3 dvrtw = ld3(G, stride, idx, Q_ROOT_VRT) + dt0 + cross(rp0, dv0);   // t0 = vrtw + rot(q, lr)
  V3 drp0 = dp0 + cross(dv0, vrtw);
  Q4 dq; dq.w = dq.x = dq.y = dq.z = 0.f;
  Q4 gq; V3 gv_;
  quat_mul_vec_bwd(q, lp, drp0, gq, gv_);  dq.w += gq.w; dq.x += gq.x; dq.y += gq.y; dq.z += gq.z;
  st3(gY, stride, idx, OFF_LPOS, gv_);
  quat_mul_vec_bwd(q, lv, dv0, gq, gv_);   dq.w += gq.w; dq.x += gq.x; dq.y += gq.y; dq.z += gq.z;
  st3(gY, stride, idx, OFF_LVEL, gv_);
  quat_mul_vec_bwd(q, lr, dt0, gq, gv_);   dq.w += gq.w; dq.x += gq.x; dq.y += gq.y; dq.z += gq.z;
  st3(gY, stride, idx, OFF_LVRT, gv_);
  // m0 = R lmat0
  M3 lm0 = orthogonalize_xy(x, yv);
  M3 t = mmt(dm0, lm0);
#pragma unroll
  for (int k = 0; k < 9; ++k) dR.m[k] += t.m[k];
  V3 dx, dyv;
  orthogonalize_xy_bwd(x, yv, mtm(R, dm0), dx, dyv);
  st3(gY, stride, idx, OFF_LTXY, ld3(G, stride, idx, Q_LTXY) + dx);
  st3(gY, stride, idx, OFF_LTXY + 3, ld3(G, stride, idx, Q_LTXY + 3) + dyv);
  gq = quat_to_xform_bwd(q, dR);           dq.w += gq.w; dq.x += gq.x; dq.y += gq.y; dq.z += gq.z;
  // gaze = rot(q^-1, unit(gaze - pos))
  V3 u = gaze - pos;
  V3 un = unit_eps(u, 1e-8f);
  quat_mul_vec_bwd(quat_inv(q), un, ld3(G, stride, idx, Q_GAZE), gq, gv_);
  dq.w += gq.w; dq.x -= gq.x; dq.y -= gq.y; dq.z -= gq.z;
  dpos = dpos - unit_eps_bwd(u, 1e-8f, gv_);
  // world root velocities use the PREVIOUS frame's rotation
  Q4 dqp; dqp.w = dqp.x = dqp.y = dqp.z = 0.f;
  quat_mul_vec_bwd(qp, vel, dvelw, gq, gv_); dqp.w += gq.w; dqp.x += gq.x; dqp.y += gq.y; dqp.z += gq.z;
  st3(gY, stride, idx, 0, gv_);
  quat_mul_vec_bwd(qp, vrt, dvrtw, gq, gv_); dqp.w += gq.w; dqp.x += gq.x; dqp.y += gq.y; dqp.z += gq.z;
  st3(gY, stride, idx, 3, gv_);
  *dpos_out = dpos; *dq_out = dq; *dqp_out = dqp;
}

}  // namespace zeggs
