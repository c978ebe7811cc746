// Host-side harness: runs the SAME __host__ __device__ per-frame loss math (csrc/loss_frame.cuh) on the CPU so
// the FK forward/adjoint can be unit-tested without a GPU (tests/test_loss_frame_host.py).  Test tooling only.
#include "../../ubisoft-laforge-zeroeggs_b200/csrc/loss_frame.cuh"
namespace zeggs { void set_error(const char*, ...) {} const char* get_error() { return ""; } }
using namespace zeggs;

static void frame_state(const float* rq, const float* rp, const float* gaze, int T, size_t idx, Q4& q, Q4& qp, V3& pos, V3& gz) {
  const float* q4 = rq + idx * 4;
  q.w = q4[0]; q.x = q4[1]; q.y = q4[2]; q.z = q4[3];
  qp = q;
  if (idx % T) { qp.w = q4[-4]; qp.x = q4[-3]; qp.y = q4[-2]; qp.z = q4[-1]; }
  pos = v3(rp[idx * 3], rp[idx * 3 + 1], rp[idx * 3 + 2]);
  gz = v3(gaze[idx * 3], gaze[idx * 3 + 1], gaze[idx * 3 + 2]);
}

extern "C" void loss_forward_host(const float* Ys, size_t stride, size_t nframes, int T, const float* rq, const float* rp,
                                  const float* gaze, const int* parents, float* Q) {
  for (size_t idx = 0; idx < nframes; ++idx) {
    Q4 q, qp; V3 pos, gz;
    frame_state(rq, rp, gaze, T, idx, q, qp, pos, gz);
    loss_frame_forward(Ys, stride, idx, q, qp, pos, gz, parents, Q);
  }
}
extern "C" void loss_backward_host(const float* Ys, const float* Q, float* G, size_t stride, size_t nframes, int T, const float* rq,
                                   const float* rp, const float* gaze, const int* parents, float* gY, float* dpos, float* dq, float* dqp) {
  for (size_t idx = 0; idx < nframes; ++idx) {
    Q4 q, qp; V3 pos, gz;
    frame_state(rq, rp, gaze, T, idx, q, qp, pos, gz);
    V3 a; Q4 b, c;
    loss_frame_backward(Ys, Q, G, stride, idx, q, qp, pos, gz, parents, gY, &a, &b, &c);
    dpos[idx * 3] = a.x; dpos[idx * 3 + 1] = a.y; dpos[idx * 3 + 2] = a.z;
    dq[idx * 4] = b.w; dq[idx * 4 + 1] = b.x; dq[idx * 4 + 2] = b.y; dq[idx * 4 + 3] = b.z;
    dqp[idx * 4] = c.w; dqp[idx * 4 + 1] = c.x; dqp[idx * 4 + 2] = c.y; dqp[idx * 4 + 3] = c.z;
  }
}
