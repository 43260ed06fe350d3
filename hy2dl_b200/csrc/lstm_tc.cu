// K1 + K2 on tcgen05 (HY2DL_PREC_TF32X3): persistent recurrent LSTM forward for H = 64.
//
// One CTA owns 128 sequences for all T steps.  Per step the gate pre-activations
//     D[128 x 256] = [x_t | h_{t-1}] [128 x K]  *  [W_ih | W_hh]^T [K x 256],   K = IP + 64
// are one tcgen05 accumulation in TMEM (input projection fused by K-concatenation, so no
// Gx[B,T,4H] ever reaches HBM).  fp32-class accuracy comes from the 3-term TF32 split
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo      (hi = a rounded to tf32, lo = exact remainder)
// i.e. 3 x K/8 MMAs of shape M128 N256 K8 per step.
//
// Data placement (why this maps well onto Blackwell):
//   * W_hi, W_lo (B operands) sit in shared memory for the whole kernel (2 x K x 256 x 4 B = 192 KB
//     for K = 96), in the no-swizzle K-major canonical layout [K/4][256][4].
//   * the A operand lives in TENSOR MEMORY: thread r of the 4 row warps owns TMEM lane r = sequence r.
//     It reads its 256 pre-activations with tcgen05.ld, applies the gate nonlinearities, keeps c in
//     registers, and writes h_t (hi and lo parts) straight back into the A columns with tcgen05.st.
//     No shared-memory round trip, no swizzle, no cross-thread exchange on the recurrent path.
//   * all time-major streams (x, h, c, gates) use the row-interleaved RI32 layout
//     [t][B/32][F/32][32 rows][32 floats] (tc_common.cuh): a row thread moves whole 32-byte sectors /
//     128-byte lines of its own row, and a (group, 32-feature block) slab is exactly the MN-major
//     SWIZZLE_128B_BASE32B UMMA operand tile that the weight-gradient GEMM consumes via bulk copies.
//
// Gate column order inside D is n = 4*u + g (unit-major), so one 32-column TMEM load gives the four
// gates of 8 hidden units.
#include "common.cuh"
#include "tc_common.cuh"

namespace hy2dl {
namespace tc {

constexpr int kH = 64;
constexpr int kN = 256;            // 4H
constexpr int kRows = 128;         // sequences per CTA
constexpr int kFwdThreads = 160;   // 4 row warps + 1 MMA warp
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColD = 0, kColAhi = 256, kColAlo = 384;   // K <= 128 columns each

// ---- weight packing for the tc path ------------------------------------------------------------
// Wt_hi / Wt_lo: [K/4][256][4] with n = 4u + g; k < IP: W_ih (zero padded), else W_hh.  bias_n[256].
__global__ void lstm_tc_pack_w_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                      const float* __restrict__ b_ih, const float* __restrict__ b_hh, int I, int IP,
                                      float* __restrict__ w_hi, float* __restrict__ w_lo, float* __restrict__ bias_n) {
  const int K = IP + kH;
  const int total = K * kN;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kk = e & 3, n = (e >> 2) & 255, kc = e >> 10;
    const int k = kc * 4 + kk, u = n >> 2, g = n & 3, row = g * kH + u;
    float v;
    if (k < IP) v = k < I ? w_ih[row * I + k] : 0.f;
    else v = w_hh[row * kH + (k - IP)];
    uint32_t hi, lo;
    split_tf32_rr(v, hi, lo);
    w_hi[e] = __uint_as_float(hi);
    w_lo[e] = __uint_as_float(lo);
    if (k == 0) bias_n[n] = b_ih[row] + b_hh[row];
  }
}

struct TcFwdArgs {
  const float4* x;      // RI32, F = IP = 32
  const float* w_hi;    // [K/4][256][4]
  const float* w_lo;
  const float* bias_n;  // [256]
  float4* h_seq;        // RI32 F = 64   (nullable)
  float4* c_seq;        // RI32 F = 64   (nullable)
  float4* gates;        // RI32 F = 256, n = 4u + g (nullable)
  float* h_last;        // row-major [B][64] (nullable)
  int64_t B, T, G;      // G = ceil(B / 32) groups
  int IP;
};

__global__ void __launch_bounds__(kFwdThreads, 1) lstm_tc_fwd_kernel(TcFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int IP = a.IP, K = IP + kH, KC = K / 4;
  float* s_whi = reinterpret_cast<float*>(smem_raw);
  float* s_wlo = s_whi + K * kN;
  float* s_bias = s_wlo + K * kN;
  uint64_t* bar_a_ready = reinterpret_cast<uint64_t*>(s_bias + kN);   // 128 row threads -> MMA warp
  uint64_t* bar_mma_done = bar_a_ready + 1;                           // MMA commit -> row threads
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_mma_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: weights -> smem, barriers, TMEM
  {
    const float4* g_hi = reinterpret_cast<const float4*>(a.w_hi);
    const float4* g_lo = reinterpret_cast<const float4*>(a.w_lo);
    float4* d_hi = reinterpret_cast<float4*>(s_whi);
    float4* d_lo = reinterpret_cast<float4*>(s_wlo);
    for (int e = tid; e < KC * kN; e += kFwdThreads) { d_hi[e] = __ldg(g_hi + e); d_lo[e] = __ldg(g_lo + e); }
    for (int e = tid; e < kN; e += kFwdThreads) s_bias[e] = __ldg(a.bias_n + e);
  }
  if (tid == 0) {
    mbar_init(bar_a_ready, kRows);
    mbar_init(bar_mma_done, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(s_tmem, kTmemCols);
  fence_async_smem();          // generic-proxy smem writes (weights) -> visible to the MMA's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < 4) {
    // =========================== row-owning threads ===========================================
    const int64_t gidx = (int64_t)blockIdx.x * 4 + warp;      // this warp's 32-row group
    const bool live = gidx < a.G;                              // warp-uniform
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    const int XC = IP / 4;                                     // x chunks (float4) per row, <= 8
    float c[kH];
#pragma unroll
    for (int u = 0; u < kH; ++u) c[u] = 0.f;

    // A(t = 0) = [x_0 | 0]
    {
      uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int q = 0; q < kH / 8; ++q) {
        tmem_st8(lane_base + kColAhi + IP + 8 * q, z);
        tmem_st8(lane_base + kColAlo + IP + 8 * q, z);
      }
      for (int j2 = 0; j2 < XC; j2 += 2) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live) v = ri32_ld(a.x, 0, a.G, gidx, 1, j2 + jj, lane);
          split_tf32(v.x, hi[4 * jj + 0], lo[4 * jj + 0]);
          split_tf32(v.y, hi[4 * jj + 1], lo[4 * jj + 1]);
          split_tf32(v.z, hi[4 * jj + 2], lo[4 * jj + 2]);
          split_tf32(v.w, hi[4 * jj + 3], lo[4 * jj + 3]);
        }
        tmem_st8(lane_base + kColAhi + 4 * j2, hi);
        tmem_st8(lane_base + kColAlo + 4 * j2, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_a_ready);
    }

    for (int64_t t = 0; t < a.T; ++t) {
      // prefetch x_{t+1} while the MMAs of step t run
      float4 xn[8];
      const bool more = t + 1 < a.T;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xn[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < XC && more && live) xn[j] = ri32_ld(a.x, t + 1, a.G, gidx, 1, j, lane);
      }
      mbar_wait(bar_mma_done, (uint32_t)(t & 1));
      tc_fence_after();

#pragma unroll
      for (int q = 0; q < kH / 8; ++q) {          // 8 hidden units (32 D columns) per iteration
        uint32_t v[32];
        tmem_ld32(lane_base + kColD + 32 * q, v);
        tmem_ld_wait();
        float hv[8], cv[8], gv[32];
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          const int u = 8 * q + uu;
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + 4 * u);
          const float gi = fast_sigmoid(__uint_as_float(v[4 * uu + 0]) + bb.x);
          const float gf = fast_sigmoid(__uint_as_float(v[4 * uu + 1]) + bb.y);
          const float gg = fast_tanh(__uint_as_float(v[4 * uu + 2]) + bb.z);
          const float go = fast_sigmoid(__uint_as_float(v[4 * uu + 3]) + bb.w);
          const float cn = gf * c[u] + gi * gg;
          c[u] = cn;
          hv[uu] = go * fast_tanh(cn);
          cv[uu] = cn;
          gv[4 * uu + 0] = gi; gv[4 * uu + 1] = gf; gv[4 * uu + 2] = gg; gv[4 * uu + 3] = go;
        }
        // h_t back into the A operand (hi / lo), columns IP + 8q .. IP + 8q + 7
        {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int uu = 0; uu < 8; ++uu) split_tf32(hv[uu], hi[uu], lo[uu]);
          tmem_st8(lane_base + kColAhi + IP + 8 * q, hi);
          tmem_st8(lane_base + kColAlo + IP + 8 * q, lo);
        }
        if (live) {
          if (a.h_seq) {
            ri32_st(a.h_seq, t, a.G, gidx, 2, 2 * q, lane, make_float4(hv[0], hv[1], hv[2], hv[3]));
            ri32_st(a.h_seq, t, a.G, gidx, 2, 2 * q + 1, lane, make_float4(hv[4], hv[5], hv[6], hv[7]));
          }
          if (a.c_seq) {
            ri32_st(a.c_seq, t, a.G, gidx, 2, 2 * q, lane, make_float4(cv[0], cv[1], cv[2], cv[3]));
            ri32_st(a.c_seq, t, a.G, gidx, 2, 2 * q + 1, lane, make_float4(cv[4], cv[5], cv[6], cv[7]));
          }
          if (a.gates) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              ri32_st(a.gates, t, a.G, gidx, 8, 8 * q + j, lane, make_float4(gv[4 * j], gv[4 * j + 1], gv[4 * j + 2], gv[4 * j + 3]));
          }
          if (a.h_last && t == a.T - 1) {
            const int64_t b = gidx * 32 + lane;
            if (b < a.B) {
#pragma unroll
              for (int uu = 0; uu < 8; ++uu) a.h_last[b * kH + 8 * q + uu] = hv[uu];
            }
          }
        }
      }
      // x_{t+1} into the A operand
      if (more) {
#pragma unroll
        for (int j2 = 0; j2 < 8; j2 += 2) {
          if (j2 < XC) {
            uint32_t hi[8], lo[8];
            split_tf32(xn[j2].x, hi[0], lo[0]); split_tf32(xn[j2].y, hi[1], lo[1]);
            split_tf32(xn[j2].z, hi[2], lo[2]); split_tf32(xn[j2].w, hi[3], lo[3]);
            split_tf32(xn[j2 + 1].x, hi[4], lo[4]); split_tf32(xn[j2 + 1].y, hi[5], lo[5]);
            split_tf32(xn[j2 + 1].z, hi[6], lo[6]); split_tf32(xn[j2 + 1].w, hi[7], lo[7]);
            tmem_st8(lane_base + kColAhi + 4 * j2, hi);
            tmem_st8(lane_base + kColAlo + 4 * j2, lo);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_a_ready);
      }
    }
  } else {
    // =========================== MMA issuer (one thread) ======================================
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(kRows, kN, 0, 0);
      const uint32_t whi = smem_u32(s_whi), wlo = smem_u32(s_wlo);
      const uint32_t lbo = kN * 16, sbo = 128;     // K-major no-swizzle: chunk-to-chunk 4096 B, 8-row groups 128 B
      const int KS = K / 8;
      for (int64_t t = 0; t < a.T; ++t) {
        mbar_wait(bar_a_ready, (uint32_t)(t & 1));
        tc_fence_after();
        // all correction terms first, while the accumulator is still ~2^-11 of its final size: the
        // tensor core's fp32 accumulation loses up to an ulp of the ACCUMULATOR per MMA, so only the
        // K/8 hi*hi instructions round at full magnitude
        for (int j = 0; j < KS; ++j) {
          const uint64_t d_hi = smem_desc(whi + (uint32_t)j * 2 * lbo, lbo, sbo);
          const uint64_t d_lo = smem_desc(wlo + (uint32_t)j * 2 * lbo, lbo, sbo);
          mma_tf32_ts(tmem + kColD, tmem + kColAlo + 8 * j, d_hi, idesc, j > 0);
          mma_tf32_ts(tmem + kColD, tmem + kColAhi + 8 * j, d_lo, idesc, 1);
        }
        for (int j = 0; j < KS; ++j) {
          const uint64_t d_hi = smem_desc(whi + (uint32_t)j * 2 * lbo, lbo, sbo);
          mma_tf32_ts(tmem + kColD, tmem + kColAhi + 8 * j, d_hi, idesc, 1);
        }
        mma_commit(bar_mma_done);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

static size_t tc_fwd_smem(int K) { return (size_t)K * kN * 4 * 2 + kN * 4 + 64; }

int64_t lstm_tc_wgrad_workspace(int64_t B, int64_t T, int64_t I);

int64_t lstm_tc_workspace(int kind, int64_t B, int64_t T, int64_t I, int64_t H) {
  (void)I;
  const int64_t IP = 32, K = IP + H;      // the RI32 input stream is always 32 columns wide
  if (kind == HY2DL_WS_LSTM_FWD) return sizeof(float) * (2 * K * 4 * H + 4 * H);
  if (kind == HY2DL_WS_LSTM_BWD) return sizeof(float) * 2 * 4 * H * H;
  if (kind == HY2DL_WS_LSTM_WGRAD) return lstm_tc_wgrad_workspace(B, T, I);
  return 16;
}

bool lstm_tc_supported(int64_t I, int64_t H) { return H == 64 && I <= 32; }

int lstm_tc_fwd(const float* x_ri, int64_t B, int64_t T, int64_t I, int64_t IP, const float* w_ih, const float* w_hh,
                const float* b_ih, const float* b_hh, float* h_seq, float* c_seq, float* gates, float* h_last, void* ws,
                int64_t ws_bytes, cudaStream_t s) {
  const int K = (int)(IP + kH);
  const int64_t need = lstm_tc_workspace(HY2DL_WS_LSTM_FWD, B, T, I, kH);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_fwd(tf32x3): workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  float* w_hi = reinterpret_cast<float*>(ws);
  float* w_lo = w_hi + K * kN;
  float* bias_n = w_lo + K * kN;
  lstm_tc_pack_w_kernel<<<96, 256, 0, s>>>(w_ih, w_hh, b_ih, b_hh, (int)I, (int)IP, w_hi, w_lo, bias_n);
  TcFwdArgs a{};
  a.x = reinterpret_cast<const float4*>(x_ri);
  a.w_hi = w_hi; a.w_lo = w_lo; a.bias_n = bias_n;
  a.h_seq = reinterpret_cast<float4*>(h_seq); a.c_seq = reinterpret_cast<float4*>(c_seq);
  a.gates = reinterpret_cast<float4*>(gates); a.h_last = h_last;
  a.B = B; a.T = T; a.G = cdiv(B, 32); a.IP = (int)IP;
  const size_t smem = tc_fwd_smem(K);
  cudaFuncSetAttribute(lstm_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lstm_tc_fwd_kernel<<<(unsigned)cdiv(B, kRows), kFwdThreads, smem, s>>>(a);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

}  // namespace tc
}  // namespace hy2dl

// ================================================================================================
// K2' on tcgen05: back-propagation through time, H = 64.
//
// Per step (reverse time) and per 128-sequence tile:
//   row threads : gate gradients dG[row][4u+g] from the tape (gates, c_t, c_{t-1}) and
//                 dh = dh_ext + dh_rec; dG goes to HBM (for the weight-gradient GEMM) and, split
//                 hi/lo, into TMEM as the A operand
//   MMA thread  : dh_rec[128 x 64] = dG[128 x 256] * W_hh[256 x 64]  (3xTF32, M128 N64 K8 MMAs)
// The K = 256 reduction is pipelined in four quarters of 64 columns (16 hidden units): while the
// tensor core multiplies quarter k, the row threads are already computing quarter k+1.  A-operand
// quarters live in a 2-slot TMEM ring, the dh accumulator is double buffered across steps.
// TMEM columns: D0 [0,64) D1 [64,128) slot0 [128,256) slot1 [256,384)  (hi = first 64, lo = last 64).
// ================================================================================================
namespace hy2dl {
namespace tc {

constexpr int kBwdThreads = 160;
constexpr uint32_t kColDh0 = 0, kColSlot0 = 128;

// W_hh as the B operand of dh = dG * W_hh: B[n = k_out][kdim = gate column 4u+g], K-major no-swizzle
// layout [256/4][64][4]; element = W_hh[g*64 + u][k_out].
__global__ void lstm_tc_pack_whh_kernel(const float* __restrict__ w_hh, float* __restrict__ hi_out, float* __restrict__ lo_out) {
  const int total = kN * kH;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kk = e & 3, n = (e >> 2) & 63, kc = e >> 8;
    const int col = kc * 4 + kk, u = col >> 2, g = col & 3;
    uint32_t hi, lo;
    split_tf32_rr(w_hh[(g * kH + u) * kH + n], hi, lo);
    hi_out[e] = __uint_as_float(hi);
    lo_out[e] = __uint_as_float(lo);
  }
}

struct TcBwdArgs {
  const float* w_hi;      // [64][64][4] packed W_hh (hi)
  const float* w_lo;
  const float4* gates;    // RI32 F = 256 (n = 4u + g)
  const float4* c_seq;    // RI32 F = 64
  const float4* dh_seq;   // RI32 F = 64 or null
  const float* dh_last;   // row-major [B][64] or null
  float4* dG;             // RI32 F = 256; may alias gates
  int64_t B, T, G, dh_t0;
};

struct BwdChunk {   // tape of 4 hidden units of one row at one step
  float4 g[4], ct, cp, dh;
};

__device__ __forceinline__ void bwd_load_chunk(const TcBwdArgs& a, int64_t t, int q, int64_t gidx, int lane, bool live, BwdChunk& c) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool on = live && t >= 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) c.g[j] = on ? ri32_ld(a.gates, t, a.G, gidx, 8, 4 * q + j, lane) : z;
  c.ct = on ? ri32_ld(a.c_seq, t, a.G, gidx, 2, q, lane) : z;
  c.cp = (on && t > 0) ? ri32_ld(a.c_seq, t - 1, a.G, gidx, 2, q, lane) : z;
  c.dh = (on && a.dh_seq && t >= a.dh_t0) ? ri32_ld(a.dh_seq, t, a.G, gidx, 2, q, lane) : z;
  if (on && a.dh_last && t == a.T - 1) {
    const int64_t b = gidx * 32 + lane;
    if (b < a.B) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.dh_last + b * kH + 4 * q));
      c.dh.x += v.x; c.dh.y += v.y; c.dh.z += v.z; c.dh.w += v.w;
    }
  }
}

__global__ void __launch_bounds__(kBwdThreads, 1) lstm_tc_bwd_kernel(TcBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* s_whi = reinterpret_cast<float*>(smem_raw);
  float* s_wlo = s_whi + kN * kH;
  float* s_dc = s_wlo + kN * kH;                                        // [64][128] cell-state adjoint, per row thread
  uint64_t* bar_ready = reinterpret_cast<uint64_t*>(s_dc + kH * kRows);  // [2] rows -> MMA, per slot
  uint64_t* bar_free = bar_ready + 2;                                   // [2] MMA commit -> rows, per slot
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_free + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  {
    const float4* g_hi = reinterpret_cast<const float4*>(a.w_hi);
    const float4* g_lo = reinterpret_cast<const float4*>(a.w_lo);
    float4* d_hi = reinterpret_cast<float4*>(s_whi);
    float4* d_lo = reinterpret_cast<float4*>(s_wlo);
    for (int e = tid; e < kN * kH / 4; e += kBwdThreads) { d_hi[e] = __ldg(g_hi + e); d_lo[e] = __ldg(g_lo + e); }
  }
  if (tid == 0) {
    mbar_init(bar_ready + 0, kRows); mbar_init(bar_ready + 1, kRows);
    mbar_init(bar_free + 0, 1); mbar_init(bar_free + 1, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(s_tmem, kTmemCols);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < 4) {
    const int64_t gidx = (int64_t)blockIdx.x * 4 + warp;
    const bool live = gidx < a.G;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    float* dcp = s_dc + tid;     // dc[u] at dcp[u * 128]: conflict-free, keeps the chunk loop rolled and small
    for (int u = 0; u < kH; ++u) dcp[u * kRows] = 0.f;

    // tape chunks travel through a 4-deep register ring: chunk q + 4 is requested as soon as chunk q has
    // been consumed, so ~4 x 112 B per thread (57 KB per SM) are in flight -- one warp per scheduler has
    // no other way to cover the HBM latency
    BwdChunk ring[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) bwd_load_chunk(a, a.T - 1, kk, gidx, lane, live, ring[kk]);
    for (int64_t s = 0; s < a.T; ++s) {
      const int64_t t = a.T - 1 - s;
      const uint32_t d_prev = lane_base + kColDh0 + 64 * (uint32_t)((s & 1) ^ 1);   // dh_rec from step s-1
      if (s > 0) {
        mbar_wait(bar_free + 1, 1);     // last quarter of step s-1 committed => dh_rec complete, both slots free
        tc_fence_after();
      }
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {      // quarter k = 16 hidden units = 64 gate columns, lives in slot k & 1
        const int slot = k & 1;
        if (k >= 2) {                    // the second use of a slot in a step waits for its first MMAs
          mbar_wait(bar_free + slot, 0);
          tc_fence_after();
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) { // 4 hidden units (16 gate columns) per chunk
          const int q = 4 * k + kk;
          BwdChunk& cur = ring[kk];
          uint32_t r[4] = {0, 0, 0, 0};
          if (s > 0) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                         : "r"(d_prev + 4 * q));
            tmem_ld_wait();
          }
          const float dhe[4] = {cur.dh.x, cur.dh.y, cur.dh.z, cur.dh.w};
          const float ctv[4] = {cur.ct.x, cur.ct.y, cur.ct.z, cur.ct.w};
          const float cpv[4] = {cur.cp.x, cur.cp.y, cur.cp.z, cur.cp.w};
          float dg[16];
#pragma unroll
          for (int uu = 0; uu < 4; ++uu) {
            const int u = 4 * q + uu;
            const float gi = cur.g[uu].x, gf = cur.g[uu].y, gg = cur.g[uu].z, go = cur.g[uu].w;
            const float dh = dhe[uu] + __uint_as_float(r[uu]);
            const float tcv = fast_tanh(ctv[uu]);
            const float dct = dcp[u * kRows] + dh * go * (1.f - tcv * tcv);
            dg[4 * uu + 0] = dct * gg * gi * (1.f - gi);
            dg[4 * uu + 1] = dct * cpv[uu] * gf * (1.f - gf);
            dg[4 * uu + 2] = dct * gi * (1.f - gg * gg);
            dg[4 * uu + 3] = dh * tcv * go * (1.f - go);
            dcp[u * kRows] = dct * gf;
          }
          // refill this ring slot: chunk q + 4 of this step, or chunk kk of the next (earlier) step
          if (k < 3) bwd_load_chunk(a, t, q + 4, gidx, lane, live, cur);
          else bwd_load_chunk(a, t - 1, kk, gidx, lane, live, cur);
          if (live) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              ri32_st(a.dG, t, a.G, gidx, 8, 4 * q + j, lane, make_float4(dg[4 * j], dg[4 * j + 1], dg[4 * j + 2], dg[4 * j + 3]));
          }
          const uint32_t sbase = lane_base + kColSlot0 + 128 * slot + 16 * kk;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_tf32(dg[8 * j + e], hi[e], lo[e]);
            tmem_st8(sbase + 8 * j, hi);
            tmem_st8(sbase + 64 + 8 * j, lo);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_ready + slot);
      }
    }
  } else {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(kRows, kH, 0, 0);
      const uint32_t whi = smem_u32(s_whi), wlo = smem_u32(s_wlo);
      const uint32_t lbo = kH * 16, sbo = 128;
      for (int64_t s = 0; s < a.T; ++s) {
        const uint32_t d = tmem + kColDh0 + 64 * (uint32_t)(s & 1);
        for (int k = 0; k < 4; ++k) {
          const int slot = k & 1;
          mbar_wait(bar_ready + slot, (uint32_t)(k >> 1));
          tc_fence_after();
          const uint32_t abase = tmem + kColSlot0 + 128 * slot;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t J = 8 * k + j;
            const uint64_t d_hi = smem_desc(whi + J * 2 * lbo, lbo, sbo);
            const uint64_t d_lo = smem_desc(wlo + J * 2 * lbo, lbo, sbo);
            mma_tf32_ts(d, abase + 64 + 8 * j, d_hi, idesc, (k | j) != 0);
            mma_tf32_ts(d, abase + 8 * j, d_lo, idesc, 1);
            mma_tf32_ts(d, abase + 8 * j, d_hi, idesc, 1);
          }
          mma_commit(bar_free + slot);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kTmemCols);
}

int lstm_tc_bwd(int64_t B, int64_t T, const float* w_hh, const float* gates, const float* c_seq, const float* dh_seq,
                int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, cudaStream_t s) {
  const int64_t need = sizeof(float) * 2 * kN * kH;
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_bwd(tf32x3): workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  float* w_hi = reinterpret_cast<float*>(ws);
  float* w_lo = w_hi + kN * kH;
  lstm_tc_pack_whh_kernel<<<64, 256, 0, s>>>(w_hh, w_hi, w_lo);
  TcBwdArgs a{};
  a.w_hi = w_hi; a.w_lo = w_lo;
  a.gates = reinterpret_cast<const float4*>(gates); a.c_seq = reinterpret_cast<const float4*>(c_seq);
  a.dh_seq = reinterpret_cast<const float4*>(dh_seq); a.dh_last = dh_last; a.dG = reinterpret_cast<float4*>(dG);
  a.B = B; a.T = T; a.G = cdiv(B, 32); a.dh_t0 = dh_t0;
  const size_t smem = (size_t)kN * kH * 4 * 2 + (size_t)kH * kRows * 4 + 64;
  cudaFuncSetAttribute(lstm_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lstm_tc_bwd_kernel<<<(unsigned)cdiv(B, kRows), kBwdThreads, smem, s>>>(a);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

}  // namespace tc
}  // namespace hy2dl
