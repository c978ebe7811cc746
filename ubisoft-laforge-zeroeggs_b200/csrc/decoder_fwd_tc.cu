// Decoder window forward on the 5th-gen tensor cores (engine 1): the same persistent, weight-stationary
// partition as decoder_fwd.cu (CTA c owns U hidden units of layer0 / GRU0 / GRU1 and ceil(1131/G) layer2 rows,
// four stages per step separated by a grid barrier) but every stage GEMM is a tcgen05.mma chain:
//
//   D[128 x N] (f32, TMEM) = X[128 x K] (bf16, smem)  *  Wslice[N x K]^T (bf16, smem)
//
//   * X = the stage's activation vector for the 32 samples of the batch tile.  Only rows 0..31 of the M = 128
//     operand are real; the k-block tiles are packed 4 KB apart, so rows 32..127 of every tile alias the following
//     k-blocks (any finite or non-finite garbage there only reaches accumulator rows that are never read).
//   * Activations live in global memory as bf16 *shared-memory images* (128-byte rows, 16-byte chunks XOR-swizzled
//     by row&7, i.e. the SWIZZLE_128B K-major canonical layout), written by the producing epilogue, so one
//     cp.async.bulk per vector brings them in -- no tensor maps, no conversion passes.
//   * Weight slices are pre-packed once per optimizer step into the same image format per (CTA, chain, k-block) and
//     streamed through a 16-slot smem ring by a dedicated producer warp that runs ahead across grid barriers.
//   * Warp roles (128 threads): warp 0 = epilogue (TMEM lanes 0..31 = the 32 samples: gates / ELU / pose integration
//     in fp32, writes next activations as bf16 images + fp32 state/history), warp 1 = MMA issuer (one lane),
//     warp 2 = weight producer, warp 3 = activation loader (waits on the grid barrier, then bulk-copies X).
//   GRU state, gates, pose integration and all saved-for-backward tensors stay fp32; only the MMA operands are bf16.
#include "decoder_common.cuh"
#include "tc_common.cuh"
#include "tc_dec_common.cuh"

namespace zeggs {

constexpr int TC_RING = 8;              // weight ring slots (two k-block tiles each)
constexpr int TC_XCH = 5;               // chunks (of 4 k-blocks) of an XA load
constexpr int TC_XKB = 18;              // k-blocks of the widest activation vector (x_pose: 1136 -> 1152)
constexpr int TC_SLOT_BYTES = 8192;     // 2 k-blocks x (32 rows x 128 B) (N <= 32)

struct TcGeom {
  int NP;          // padded rows of a gate chain: round_up(3U,16)
  int N1;          // rows of the stage-1 chain: 4U
  int kbH, kbX;    // k-blocks of an H-vector / of x_pose
  int n4t;         // 16-row tiles of layer2 per CTA
  int nacc;        // independent TMEM accumulators per chain (breaks the MMA->MMA accumulate dependency)
  size_t chain_off[6];   // byte offset of chain c inside one CTA's packed block (chain 5 = first layer2 tile)
  size_t cta_bytes;
};

__host__ __device__ inline size_t tc_tile_bytes(int N) { return (size_t)N * 128; }

inline TcGeom make_tcgeom(const DecGeom& g) {
  TcGeom t;
  t.NP = round_up(3 * g.U, 16);
  t.N1 = 4 * g.U;
  t.kbH = ceil_div(g.H, 64);
  t.kbX = ceil_div(K1P, 64);
  t.n4t = g.n4t;
  t.nacc = 4;
  size_t off = 0;
  t.chain_off[0] = off; off += (size_t)t.kbX * tc_tile_bytes(t.N1);     // S1   X = x_pose
  t.chain_off[1] = off; off += (size_t)t.kbH * tc_tile_bytes(t.NP);     // gh0  X = h0(t-1)
  t.chain_off[2] = off; off += (size_t)t.kbH * tc_tile_bytes(t.NP);     // gi0a X = a(t)
  t.chain_off[3] = off; off += (size_t)t.kbH * tc_tile_bytes(t.NP);     // gh1  X = h1(t-1)
  t.chain_off[4] = off; off += (size_t)t.kbH * tc_tile_bytes(t.NP);     // gi1  X = h0(t)
  t.chain_off[5] = off; off += (size_t)t.n4t * t.kbH * tc_tile_bytes(16);  // y tiles X = h1(t)
  t.cta_bytes = off;
  return t;
}

// ------------------------------------------------------------------ packing (bf16 images of the weight slices)
__global__ void pack_decoder_tc_kernel(DecGeom g, TcGeom tg, const float* __restrict__ W0, const float* __restrict__ Wih0,
                                       const float* __restrict__ Whh0, const float* __restrict__ Wih1,
                                       const float* __restrict__ Whh1, const float* __restrict__ W2, uint8_t* __restrict__ out) {
  const int H = g.H, U = g.U, A = g.A;
  const size_t total_elems = (size_t)g.G * tg.cta_bytes / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total_elems; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / (tg.cta_bytes / 2));
    size_t b = (i % (tg.cta_bytes / 2)) * 2;           // byte offset inside the CTA block
    int chain = 5;
    for (int q = 0; q < 5; ++q) if (b < tg.chain_off[q + 1]) { chain = q; break; }
    b -= tg.chain_off[chain];
    int N = chain == 0 ? tg.N1 : (chain == 5 ? 16 : tg.NP);
    int tile = 0;
    if (chain == 5) { tile = (int)(b / ((size_t)tg.kbH * tc_tile_bytes(16))); b -= (size_t)tile * tg.kbH * tc_tile_bytes(16); }
    const int kb = (int)(b / tc_tile_bytes(N));
    const int rb = (int)(b % tc_tile_bytes(N));
    const int row = rb / 128, chunk_phys = (rb % 128) / 16, e = (rb % 16) / 2;
    const int k = kb * 64 + ((chunk_phys ^ (row & 7)) << 3) + e;
    float v = 0.f;
    if (chain == 0) {
      const int gi = row / U, j = c * U + row % U;
      if (k < P_IN) v = gi == 0 ? W0[(size_t)j * A + k] : Wih0[(size_t)((gi - 1) * H + j) * (A + H) + H + k];
    } else if (chain <= 4) {
      if (row < 3 * U && k < H) {
        const int gi = row / U, j = c * U + row % U;
        const size_t r = (size_t)(gi * H + j);
        v = chain == 1 ? Whh0[r * H + k] : chain == 2 ? Wih0[r * (A + H) + k] : chain == 3 ? Whh1[r * H + k] : Wih1[r * H + k];
      }
    } else {
      const int lr = tile * 16 + row, n = c * g.rpc + lr;
      if (lr < g.rpc && n < P_OUT && k < H) v = W2[(size_t)n * H + k];
    }
    reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
  }
}

// fp32 k-major vector [K][32] -> bf16 image (32 rows); K padded with zeros to kbs*64
__global__ void image_from_kmajor_kernel(const float* __restrict__ src, int K, int kbs, uint8_t* __restrict__ img) {
  const int total = kbs * 64 * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i & 31, k = i >> 5;
    const float v = k < K ? src[(size_t)k * 32 + b] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(img + img_off(32, b, k)) = __float2bfloat16_rn(v);
  }
}

struct TcWs {
  uint8_t *xpb[2], *ab, *h0b[2], *h1b[2];   // bf16 activation images
  long long* dbg;                            // optional per-step phase timestamps of CTA 0 (clock64), [T][32]
  size_t bytes;
};
inline TcWs make_tcws(void* base, const DecGeom& g) {
  TcWs w; size_t off = 0;
  auto take = [&](size_t n) { uint8_t* p = base ? (uint8_t*)base + off : nullptr; off += ((n + 1023) / 1024) * 1024; return p; };
  const size_t xb = (size_t)ceil_div(K1P, 64) * 4096, hb = (size_t)ceil_div(g.H, 64) * 4096;
  w.xpb[0] = take(xb); w.xpb[1] = take(xb); w.ab = take(hb);
  w.h0b[0] = take(hb); w.h0b[1] = take(hb); w.h1b[0] = take(hb); w.h1b[1] = take(hb);
  w.dbg = nullptr;
  w.bytes = off; return w;
}

#define TCDBG(ev) do { if (tw.dbg && c == 0 && lane == 0 && t < 64) tw.dbg[t * 32 + (ev)] = clock64(); } while (0)

template <int U>
__global__ void __launch_bounds__(128, 1)
decoder_fwd_tc_kernel(zeggs_decoder_fwd_args a, DecGeom g, TcGeom tg, DecWs w, TcWs tw, const uint8_t* __restrict__ packed) {
  constexpr int NP = (3 * U + 15) / 16 * 16;     // gate-chain rows (32 for U=8, 16 for U=4)
  constexpr int N1 = 4 * U;                      // stage-1 rows
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // layout: XA | XB | ring | 12 KB slack (operand rows 32..127 of the last k-blocks alias whatever follows) | barriers | constants
  uint8_t* XA = smem;
  uint8_t* XB = XA + TC_XKB * 4096;
  uint8_t* ring = XB + tg.kbH * 4096;
  uint8_t* tail = ring + TC_RING * TC_SLOT_BYTES + 12288;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = bars;                 // [TC_RING]
  uint64_t* empty = bars + TC_RING;      // [TC_RING]
  uint64_t* xa_full = bars + 2 * TC_RING;   // [TC_XCH] one per 4-k-block chunk of XA: the MMA chain starts on chunk 0
  uint64_t* xb_full = xa_full + TC_XCH;
  uint64_t* xa_free = xb_full + 1;
  uint64_t* xb_free = xb_full + 2;
  uint64_t* d_full = xb_full + 3;        // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 4);
  float* cst = reinterpret_cast<float*>(tail + 512);   // per-CTA constants (biases, normalisation rows)
  const int R4 = tg.n4t * 16;
  float* c_bhh0 = cst;            // [3U]
  float* c_bih1 = cst + 3 * U;    // [3U]
  float* c_bhh1 = cst + 6 * U;    // [3U]
  float* c_b2 = cst + 9 * U;      // [R4]  then out_std, out_mean, in_mean, in_std (R4 each), then gaze in_mean/in_std (3+3)
  float* c_os = c_b2 + R4; float* c_om = c_os + R4; float* c_im = c_om + R4; float* c_is = c_im + R4; float* c_gz = c_is + R4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x, H = a.H, T = a.T;
  const int kbH = tg.kbH, kbX = tg.kbX, n4t = tg.n4t;
  const uint8_t* pk = packed + (size_t)c * tg.cta_bytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < TC_XCH; ++i) mbar_init(&xa_full[i], 1);
    mbar_init(xb_full, 1); mbar_init(xa_free, 1); mbar_init(xb_free, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&d_full[i], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 3 * U; i += blockDim.x) {
    const int j = (i / U) * H + c * U + (i % U);
    c_bhh0[i] = a.b_hh0[j]; c_bih1[i] = a.b_ih1[j]; c_bhh1[i] = a.b_hh1[j];
  }
  for (int i = threadIdx.x; i < R4; i += blockDim.x) {
    const int n = c * g.rpc + i;
    const bool ok = i < g.rpc && n < P_OUT;
    c_b2[i] = ok ? a.b2[n] : 0.f; c_os[i] = ok ? a.out_std[n] : 0.f; c_om[i] = ok ? a.out_mean[n] : 0.f;
    c_im[i] = ok ? a.in_mean[n] : 0.f; c_is[i] = ok ? a.in_std[n] : 1.f;
  }
  if (threadIdx.x < 3) { c_gz[threadIdx.x] = a.in_mean[P_OUT + threadIdx.x]; c_gz[3 + threadIdx.x] = a.in_std[P_OUT + threadIdx.x]; }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // The CTA allocates all 512 TMEM columns (1 CTA/SM), so the allocation base is column 0 / lane 0.  Using the literal
  // keeps every tcgen05 address operand provably warp-uniform: otherwise the compiler wraps each UTCHMMA in an
  // ELECT / R2UR.BROADCAST waterfall loop (~90 cycles per MMA, the issue-bound regime seen in the first trace).
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem = 0u;
  // TMEM regions (columns): gh0 [0,128) | gh1 [128,256) | S1 / gi0a / gi1 / y [256,512); accumulator q of a chain at +q*N
  constexpr uint32_t R_GH0 = 0, R_GH1 = 128, R_MAIN = 256;
  const size_t actH = (size_t)g.nbt * H * 32, actX = (size_t)g.nbt * K1P * 32;

  if (warp == 2) {
    // ================= weight producer: streams every chain's tiles in the MMA warp's consumption order
    if (lane == 0) {
      uint32_t it = 0;
      auto stream = [&](int chain, int tile) {
        const int N = chain == 0 ? N1 : (chain == 5 ? 16 : NP);
        const int nkb = chain == 0 ? kbX : kbH;
        const uint8_t* src = pk + tg.chain_off[chain] + (size_t)tile * kbH * tc_tile_bytes(16);
        for (int kb = 0; kb < nkb; kb += 2, ++it) {
          const int s = it & (TC_RING - 1); const uint32_t ph = (it / TC_RING) & 1;
          const uint32_t bytes = (uint32_t)((nkb - kb >= 2 ? 2 : 1) * tc_tile_bytes(N));
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], bytes);
          bulk_g2s(ring + s * TC_SLOT_BYTES, src + (size_t)kb * tc_tile_bytes(N), bytes, &full[s]);
        }
      };
      for (int t = 1; t < T; ++t) {   // order: gh0(1), S1(0), gh1(3), gi0a(2), gi1(4), y tiles(5)
        stream(1, 0); stream(0, 0); stream(3, 0); stream(2, 0); stream(4, 0);
        for (int tile = 0; tile < n4t; ++tile) stream(5, tile);
      }
    }
  } else if (warp == 3) {
    // ================= activation loader
    if (lane == 0) {
      uint32_t xa_n = 0, xb_n = 0;        // loads issued so far into XA / XB
      auto load_xa = [&](const uint8_t* img, int nkb) {
        if (xa_n > 0) mbar_wait(xa_free, (xa_n - 1) & 1);
        fence_proxy_async();
        for (int ch = 0; ch * 4 < nkb; ++ch) {
          const uint32_t bytes = (uint32_t)((nkb - ch * 4 >= 4 ? 4 : nkb - ch * 4) * 4096);
          mbar_arrive_expect_tx(&xa_full[ch], bytes);
          bulk_g2s(XA + ch * 16384, img + (size_t)ch * 16384, bytes, &xa_full[ch]);
        }
        for (int ch = (nkb + 3) / 4; ch < TC_XCH; ++ch) mbar_arrive(&xa_full[ch]);   // keep every chunk barrier in phase
        ++xa_n;
      };
      auto load_xb = [&](const uint8_t* img) {
        if (xb_n > 0) mbar_wait(xb_free, (xb_n - 1) & 1);
        fence_proxy_async();
        mbar_arrive_expect_tx(xb_full, (uint32_t)kbH * 4096);
        bulk_g2s(XB, img, (uint32_t)kbH * 4096, xb_full);
        ++xb_n;
      };
      unsigned epoch = 0;
      for (int t = 1; t < T; ++t) {
        load_xb(tw.h0b[(t - 1) & 1]);                               // h0(t-1): complete since barrier B2 of step t-1
        if (t > 1) grid_wait(w.bar, (++epoch) * gridDim.x);          // B4(t-1): x_pose(t) complete
        TCDBG(0);
        load_xa(tw.xpb[t & 1], kbX);
        load_xb(tw.h1b[(t - 1) & 1]);                               // h1(t-1)
        grid_wait(w.bar, (++epoch) * gridDim.x);                    // B1
        TCDBG(8);
        load_xa(tw.ab, kbH);
        grid_wait(w.bar, (++epoch) * gridDim.x);                    // B2
        TCDBG(14);
        load_xa(tw.h0b[t & 1], kbH);
        grid_wait(w.bar, (++epoch) * gridDim.x);                    // B3
        TCDBG(20);
        load_xa(tw.h1b[t & 1], kbH);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer.  The warp runs the loops converged; one elected lane issues.  Descriptors advance by
    // constants and the four k-steps of a k-block go to four independent TMEM accumulators.
    {
      uint32_t it = 0, xa_n = 0, xb_n = 0;
      const uint64_t dXA = make_smem_desc_sw128(XA), dXB = make_smem_desc_sw128(XB), dRing = make_smem_desc_sw128(ring);
      auto chain_mma = [&](uint64_t dx, int nkb, int N, uint32_t d0, int xph) {   // xph >= 0: X arrives in chunks (XA)
        const uint32_t idesc = make_idesc_bf16_f32(128, N);
        const uint64_t bstep = (uint64_t)(N * 8);            // one k-block tile of the weight slice, in 16-byte units
        for (int kb = 0; kb < nkb; kb += 2, ++it) {
          const uint32_t s = it & (TC_RING - 1), ph = (it / TC_RING) & 1;
          if (xph >= 0 && (kb & 3) == 0) mbar_wait(&xa_full[kb >> 2], (uint32_t)xph);
          mbar_wait(&full[s], ph);
          tc_fence_after_sync();
          const uint64_t da = dx + (uint64_t)kb * 256, db = dRing + (uint64_t)s * (TC_SLOT_BYTES >> 4);
          const bool acc0 = kb > 0, two = kb + 1 < nkb;
          if (elect_one_sync()) {
            umma_bf16(d0 + 0 * N, da + 0, db + 0, idesc, acc0);
            umma_bf16(d0 + 1 * N, da + 2, db + 2, idesc, acc0);
            umma_bf16(d0 + 2 * N, da + 4, db + 4, idesc, acc0);
            umma_bf16(d0 + 3 * N, da + 6, db + 6, idesc, acc0);
            if (two) {
              umma_bf16(d0 + 0 * N, da + 256 + 0, db + bstep + 0, idesc, true);
              umma_bf16(d0 + 1 * N, da + 256 + 2, db + bstep + 2, idesc, true);
              umma_bf16(d0 + 2 * N, da + 256 + 4, db + bstep + 4, idesc, true);
              umma_bf16(d0 + 3 * N, da + 256 + 6, db + bstep + 6, idesc, true);
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
      };
      auto commit2 = [&](uint64_t* b0, uint64_t* b1) {
        if (elect_one_sync()) { umma_commit(b0); if (b1) umma_commit(b1); }
        __syncwarp();
      };
      for (int t = 1; t < T; ++t) {
        mbar_wait(xb_full, xb_n & 1); ++xb_n; tc_fence_after_sync();
        chain_mma(dXB, kbH, NP, tmem + R_GH0, -1);                        // gh0
        commit2(xb_free, nullptr);
        TCDBG(3);
        chain_mma(dXA, kbX, N1, tmem + R_MAIN, xa_n & 1); ++xa_n;         // S1
        commit2(xa_free, &d_full[0]);
        TCDBG(4);
        mbar_wait(xb_full, xb_n & 1); ++xb_n; tc_fence_after_sync();
        chain_mma(dXB, kbH, NP, tmem + R_GH1, -1);                        // gh1
        commit2(xb_free, nullptr);
        TCDBG(10);
        chain_mma(dXA, kbH, NP, tmem + R_MAIN, xa_n & 1); ++xa_n;         // gi0a
        commit2(xa_free, &d_full[1]);
        TCDBG(11);
        TCDBG(15);
        chain_mma(dXA, kbH, NP, tmem + R_MAIN, xa_n & 1); ++xa_n;         // gi1
        commit2(xa_free, &d_full[2]);
        TCDBG(16);
        TCDBG(21);
        for (int tile = 0; tile < n4t; ++tile) chain_mma(dXA, kbH, 16, tmem + R_MAIN + (uint32_t)(tile * 64), tile == 0 ? (int)(xa_n & 1) : -1);   // y
        ++xa_n;
        commit2(xa_free, &d_full[3]);
      }
    }
  } else {
    // ================= epilogue warp (TMEM lanes 0..31 = samples)
    const int b = lane;
    const bool live = b < a.B;
    const int j0 = c * U;
    constexpr int na_main = 4, na_side = 4;
    auto ld_sum = [&](uint32_t col, int na, float (&v)[NP]) {     // sum of a gate chain's accumulators
      tmem_ld_cols<NP>(tmem + col, v);
      for (int q = 1; q < na; ++q) {
        float u_[NP];
        tmem_ld_cols<NP>(tmem + col + (uint32_t)(q * NP), u_);
#pragma unroll
        for (int i = 0; i < NP; ++i) v[i] += u_[i];
      }
    };
    float gi0p[3 * U];
    for (int t = 1; t < T; ++t) {
      const uint32_t ph = (t - 1) & 1;
      const int ts = w.save ? t : (t & 1), tp = w.save ? t - 1 : ((t - 1) & 1), tn = w.save ? t + 1 : ((t + 1) & 1);
      // ---------------- stage 1   (operands that do not depend on the MMA are fetched before the wait)
      float sv[4 * U];
      {
        const float* S = w.S01 + ((size_t)t * 32 + b) * 4 * H + j0;      // [t][b][4H]: U consecutive floats per gate block
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
          for (int u4 = 0; u4 < U; u4 += 4) {
            const float4 v4 = __ldg(reinterpret_cast<const float4*>(S + (size_t)q * H + u4));
            sv[q * U + u4 + 0] = v4.x; sv[q * U + u4 + 1] = v4.y; sv[q * U + u4 + 2] = v4.z; sv[q * U + u4 + 3] = v4.w;
          }
        }
      }
      mbar_wait(&d_full[0], ph);
      tc_fence_after_sync();
      TCDBG(5);
      {
        float v[N1];
        tmem_ld_cols<N1>(tmem + R_MAIN, v);
        for (int q = 1; q < na_main; ++q) {
          float u_[N1];
          tmem_ld_cols<N1>(tmem + R_MAIN + (uint32_t)(q * N1), u_);
#pragma unroll
          for (int i = 0; i < N1; ++i) v[i] += u_[i];
        }
        float av[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          av[u] = elu_f(v[u] + sv[u]);
#pragma unroll
          for (int q = 0; q < 3; ++q) gi0p[q * U + u] = v[(1 + q) * U + u] + sv[(1 + q) * U + u];
        }
        store_img_units<U>(tw.ab, b, j0, av);
        tc_fence_before_sync();
        TCDBG(6);
        grid_arrive(w.bar);                     // only the bf16 image feeds other CTAs: publish it first ...
        TCDBG(7);
#pragma unroll
        for (int u = 0; u < U; ++u) w.A[ts * actH + (size_t)(j0 + u) * 32 + b] = av[u];   // ... fp32 history afterwards
      }
      // ---------------- stage 2 (GRU layer 0)
      float hp[U];
#pragma unroll
      for (int u = 0; u < U; ++u) hp[u] = w.H0[tp * actH + (size_t)(j0 + u) * 32 + b];
      mbar_wait(&d_full[1], ph);
      tc_fence_after_sync();
      TCDBG(12);
      {
        float gh[NP], gi[NP];
        ld_sum(R_GH0, na_side, gh);
        ld_sum(R_MAIN, na_main, gi);
        float hv[U], rr[U], zz[U], nn[U], gn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          rr[u] = fast_sigmoid(gi[u] + gi0p[u] + gh[u] + c_bhh0[u]);
          zz[u] = fast_sigmoid(gi[U + u] + gi0p[U + u] + gh[U + u] + c_bhh0[U + u]);
          gn[u] = gh[2 * U + u] + c_bhh0[2 * U + u];
          nn[u] = fast_tanh(gi[2 * U + u] + gi0p[2 * U + u] + rr[u] * gn[u]);
          hv[u] = (1.f - zz[u]) * nn[u] + zz[u] * hp[u];
        }
        store_img_units<U>(tw.h0b[t & 1], b, j0, hv);
        tc_fence_before_sync();
        TCDBG(13);
        grid_arrive(w.bar);
#pragma unroll
        for (int u = 0; u < U; ++u) w.H0[ts * actH + (size_t)(j0 + u) * 32 + b] = hv[u];
        if (w.save) {
          float* G = w.G0 + ((size_t)t * g.nbt) * 4 * H * 32;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = j0 + u;
            G[(size_t)(0 * H + j) * 32 + b] = rr[u]; G[(size_t)(1 * H + j) * 32 + b] = zz[u];
            G[(size_t)(2 * H + j) * 32 + b] = nn[u]; G[(size_t)(3 * H + j) * 32 + b] = gn[u];
          }
        }
      }
      // ---------------- stage 3 (GRU layer 1)
#pragma unroll
      for (int u = 0; u < U; ++u) hp[u] = w.H1[tp * actH + (size_t)(j0 + u) * 32 + b];
      mbar_wait(&d_full[2], ph);
      tc_fence_after_sync();
      TCDBG(17);
      {
        float gh[NP], gi[NP];
        ld_sum(R_GH1, na_side, gh);
        ld_sum(R_MAIN, na_main, gi);
        float hv[U], rr[U], zz[U], nn[U], gn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          rr[u] = fast_sigmoid(gi[u] + c_bih1[u] + gh[u] + c_bhh1[u]);
          zz[u] = fast_sigmoid(gi[U + u] + c_bih1[U + u] + gh[U + u] + c_bhh1[U + u]);
          gn[u] = gh[2 * U + u] + c_bhh1[2 * U + u];
          nn[u] = fast_tanh(gi[2 * U + u] + c_bih1[2 * U + u] + rr[u] * gn[u]);
          hv[u] = (1.f - zz[u]) * nn[u] + zz[u] * hp[u];
        }
        store_img_units<U>(tw.h1b[t & 1], b, j0, hv);
        tc_fence_before_sync();
        TCDBG(18);
        grid_arrive(w.bar);
        TCDBG(19);
#pragma unroll
        for (int u = 0; u < U; ++u) w.H1[ts * actH + (size_t)(j0 + u) * 32 + b] = hv[u];
        if (w.save) {
          float* G = w.G1 + ((size_t)t * g.nbt) * 4 * H * 32;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = j0 + u;
            G[(size_t)(0 * H + j) * 32 + b] = rr[u]; G[(size_t)(1 * H + j) * 32 + b] = zz[u];
            G[(size_t)(2 * H + j) * 32 + b] = nn[u]; G[(size_t)(3 * H + j) * 32 + b] = gn[u];
          }
        }
      }
      // ---------------- stage 4 (layer2, de-normalise, pose integration, next x_pose)
      V3 pos = v3(0, 0, 0), gzp = v3(0, 0, 0);
      Q4 q; q.w = 1.f; q.x = q.y = q.z = 0.f;
      if (c == 0 && live) {
        const float* rp = a.root_pos + ((size_t)b * T + (t - 1)) * 3;
        const float* rq = a.root_rot + ((size_t)b * T + (t - 1)) * 4;
        pos = v3(rp[0], rp[1], rp[2]);
        q.w = rq[0]; q.x = rq[1]; q.y = rq[2]; q.z = rq[3];
        if (t + 1 < T) { const float* gp = a.gaze_pos + ((size_t)b * T + (t + 1)) * 3; gzp = v3(gp[0], gp[1], gp[2]); }
      }
      mbar_wait(&d_full[3], ph);
      tc_fence_after_sync();
      TCDBG(22);
      {
        uint8_t* xpn = tw.xpb[(t + 1) & 1];
        float* xpf = w.XP + tn * actX;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f, r4 = 0.f, r5 = 0.f;
        for (int tile = 0; tile < n4t; ++tile) {
          float y[16];
          tmem_ld_cols<16>(tmem + R_MAIN + (uint32_t)(tile * na_side * 16), y);
          for (int qa = 1; qa < na_side; ++qa) {
            float u_[16];
            tmem_ld_cols<16>(tmem + R_MAIN + (uint32_t)((tile * na_side + qa) * 16), u_);
#pragma unroll
            for (int i = 0; i < 16; ++i) y[i] += u_[i];
          }
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const int lr = tile * 16 + r, n = c * g.rpc + lr;
            if (lr < g.rpc && n < P_OUT) {
              const float p = (y[r] + c_b2[lr]) * c_os[lr] + c_om[lr];
              if (live) a.Y[((size_t)b * T + t) * P_OUT + n] = p;
              if (t + 1 < T) {
                const float xn = (p - c_im[lr]) / c_is[lr];
                *reinterpret_cast<__nv_bfloat16*>(xpn + img_off(32, b, n)) = __float2bfloat16_rn(xn);
                if (w.save) xpf[(size_t)n * 32 + b] = xn;
              }
              if (tile == 0) { if (r == 0) r0 = p; if (r == 1) r1 = p; if (r == 2) r2 = p; if (r == 3) r3 = p; if (r == 4) r4 = p; if (r == 5) r5 = p; }
            }
          }
        }
        if (c == 0 && live) {
          V3 npos = quat_mul_vec(q, a.dt * v3(r0, r1, r2)) + pos;
          Q4 nq = quat_mul(quat_from_helical(quat_mul_vec(q, a.dt * v3(r3, r4, r5))), q);
          float* op = a.root_pos + ((size_t)b * T + t) * 3;
          float* oq = a.root_rot + ((size_t)b * T + t) * 4;
          op[0] = npos.x; op[1] = npos.y; op[2] = npos.z;
          oq[0] = nq.w; oq[1] = nq.x; oq[2] = nq.y; oq[3] = nq.z;
          if (t + 1 < T) {
            V3 gd = quat_mul_vec(quat_inv(nq), gzp - npos);
            const float gx[3] = {gd.x, gd.y, gd.z};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const float xn = (gx[d] - c_gz[d]) / c_gz[3 + d];
              *reinterpret_cast<__nv_bfloat16*>(xpn + img_off(32, b, P_OUT + d)) = __float2bfloat16_rn(xn);
              if (w.save) xpf[(size_t)(P_OUT + d) * 32 + b] = xn;
            }
          }
        }
      }
      tc_fence_before_sync();
      TCDBG(23);
      if (t + 1 < T) grid_arrive(w.bar);
      TCDBG(24);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------ host
extern "C" size_t zeggs_decoder_packed_tc_bytes(int H, int S, int Z) {
  if (H % 16 != 0 || pick_U(H) <= 0) return 0;
  DecGeom g = make_geom(1, H, S, Z);
  return (size_t)g.G * make_tcgeom(g).cta_bytes;
}
extern "C" size_t zeggs_decoder_tc_workspace_bytes(int H, int S, int Z) {
  if (H % 16 != 0 || pick_U(H) <= 0) return 0;
  DecGeom g = make_geom(1, H, S, Z);
  return make_tcws(nullptr, g).bytes;
}
extern "C" int zeggs_decoder_pack_weights_tc(const zeggs_decoder_fwd_args* a, void* packed, void* stream_) {
  ZCHECK_ARG(a && packed && a->H % 16 == 0 && pick_U(a->H) > 0, "decoder tc pack: bad arguments");
  DecGeom g = make_geom(a->B, a->H, a->S, a->Z);
  TcGeom tg = make_tcgeom(g);
  ZCHECK_ARG(tg.n4t <= 4, "decoder tc: hidden size %d too small for the tensor-core engine", a->H);
  pack_decoder_tc_kernel<<<592, 256, 0, (cudaStream_t)stream_>>>(g, tg, a->W0, a->W_ih0, a->W_hh0, a->W_ih1, a->W_hh1, a->W2, (uint8_t*)packed);
  count_launch();
  ZCHECK_LAUNCH();
  return ZEGGS_OK;
}

template <int U>
static int launch_tc(const zeggs_decoder_fwd_args& a, const DecGeom& g, const TcGeom& tg, const DecWs& w, const TcWs& tw,
                     const uint8_t* packed, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)TC_XKB * 4096 + (size_t)tg.kbH * 4096 + TC_RING * TC_SLOT_BYTES + 12288 + 512 +
                      (size_t)(9 * U + 5 * tg.n4t * 16 + 8) * sizeof(float);
  ZCHECK_CUDA(cudaFuncSetAttribute(decoder_fwd_tc_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, nsm = 0, occ = 0;
  ZCHECK_CUDA(cudaGetDevice(&dev));
  ZCHECK_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  ZCHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, decoder_fwd_tc_kernel<U>, 128, smem));
  ZCHECK_ARG(occ * nsm >= g.G, "decoder tc: cooperative grid of %d CTAs does not fit", g.G);
  void* args[] = {(void*)&a, (void*)&g, (void*)&tg, (void*)&w, (void*)&tw, (void*)&packed};
  ZCHECK_CUDA(cudaLaunchCooperativeKernel((void*)decoder_fwd_tc_kernel<U>, dim3(g.G), dim3(128), args, smem, stream));
  count_launch();
  return ZEGGS_OK;
}

static long long* g_tc_dbg = nullptr;
static int g_tc_nacc = 0;
extern "C" void zeggs_debug_set_tc_nacc(int n) { g_tc_nacc = n; }
long long* tc_debug_buffer() { return g_tc_dbg; }
extern "C" void zeggs_debug_set_tc_trace(void* p) { g_tc_dbg = (long long*)p; }

// called by zeggs_decoder_window_fwd after the prologue / CellStateEncoder / cond pre-pass when engine == 1
int decoder_fwd_tc_run(const zeggs_decoder_fwd_args& a, const DecGeom& g, const DecWs& w, cudaStream_t stream) {
  TcGeom tg = make_tcgeom(g);
  ZCHECK_ARG(g.nbt == 1, "decoder tc engine handles one 32-sample batch tile (B <= 32); got B=%d", a.B);
  ZCHECK_ARG(a.packed_tc && a.workspace_tc, "decoder tc: packed_tc / workspace_tc missing");
  ZCHECK_ARG(tg.n4t <= 4 && tg.kbH <= 16, "decoder tc: unsupported hidden size %d", a.H);
  if (g_tc_nacc > 0) tg.nacc = g_tc_nacc;
  TcWs tw = make_tcws(a.workspace_tc, g);
  tw.dbg = g_tc_dbg;
  const size_t xb = (size_t)tg.kbX * 4096, hb = (size_t)tg.kbH * 4096;
  ZCHECK_CUDA(cudaMemsetAsync(tw.xpb[0], 0, xb, stream));
  ZCHECK_CUDA(cudaMemsetAsync(tw.xpb[1], 0, xb, stream));
  // images of x_pose(1), h0(0), h1(0) from the fp32 k-major buffers the prologue / CellStateEncoder wrote
  image_from_kmajor_kernel<<<64, 256, 0, stream>>>(w.XP + (size_t)1 * g.nbt * K1P * 32, K1P, tg.kbX, tw.xpb[1]); count_launch();
  image_from_kmajor_kernel<<<64, 256, 0, stream>>>(w.H0, a.H, tg.kbH, tw.h0b[0]); count_launch();
  image_from_kmajor_kernel<<<64, 256, 0, stream>>>(w.H1, a.H, tg.kbH, tw.h1b[0]); count_launch();
  ZCHECK_LAUNCH();
  (void)hb;
  ScopedTimer tm("decoder_fwd", stream);
  return g.U == 4 ? launch_tc<4>(a, g, tg, w, tw, (const uint8_t*)a.packed_tc, stream)
                  : launch_tc<8>(a, g, tg, w, tw, (const uint8_t*)a.packed_tc, stream);
}

}  // namespace zeggs
