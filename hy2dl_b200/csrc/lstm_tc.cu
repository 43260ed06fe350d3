// K1 + K2 on tcgen05 (HY2DL_PREC_TF32X3): persistent recurrent LSTM forward for H = 64.
//
// One CTA owns 128 sequences for all T steps.  Per step the gate pre-activations
//     D[128 x 256] = [x_t | h_{t-1}] [128 x K]  *  [W_ih | W_hh]^T [K x 256],   K = IP + 64
// are one tcgen05 accumulation in TMEM (input projection fused by K-concatenation, so no
// Gx[B,T,4H] ever reaches HBM).  fp32-class accuracy comes from the 3-term TF32 split
//     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo      (hi = top 19 bits, lo = exact remainder)
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
//     [t][B/32][F/4][32 rows][4 floats], so a warp-wide 16-byte-per-lane access by the row-owning
//     threads is one fully coalesced 512-byte transaction, and a 32-row group is also exactly the
//     MN-major no-swizzle UMMA operand layout that the weight-gradient GEMM consumes via bulk copies.
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

// ---- RI32 addressing: element (row, f) of a [rows][F] slab at step t --------------------------
// float4 index of chunk j (4 floats) of row r in group gidx: ((t*G + gidx) * (F/4) + j) * 32 + (r & 31)
__device__ __forceinline__ int64_t ri32_chunk(int64_t t, int64_t G, int64_t gidx, int F4, int j, int lane) {
  return ((t * G + gidx) * F4 + j) * 32 + lane;
}

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
    split_tf32(v, hi, lo);
    w_hi[e] = __uint_as_float(hi);
    w_lo[e] = __uint_as_float(lo);
    if (k == 0) bias_n[n] = b_ih[row] + b_hh[row];
  }
}

struct TcFwdArgs {
  const float4* x;      // RI32 [T][G][IP/4][32][4]
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
          if (live) v = __ldg(a.x + ri32_chunk(0, a.G, gidx, XC, j2 + jj, lane));
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
        if (j < XC && more && live) xn[j] = __ldg(a.x + ri32_chunk(t + 1, a.G, gidx, XC, j, lane));
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
            a.h_seq[ri32_chunk(t, a.G, gidx, 16, 2 * q, lane)] = make_float4(hv[0], hv[1], hv[2], hv[3]);
            a.h_seq[ri32_chunk(t, a.G, gidx, 16, 2 * q + 1, lane)] = make_float4(hv[4], hv[5], hv[6], hv[7]);
          }
          if (a.c_seq) {
            a.c_seq[ri32_chunk(t, a.G, gidx, 16, 2 * q, lane)] = make_float4(cv[0], cv[1], cv[2], cv[3]);
            a.c_seq[ri32_chunk(t, a.G, gidx, 16, 2 * q + 1, lane)] = make_float4(cv[4], cv[5], cv[6], cv[7]);
          }
          if (a.gates) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              a.gates[ri32_chunk(t, a.G, gidx, 64, 8 * q + j, lane)] = make_float4(gv[4 * j], gv[4 * j + 1], gv[4 * j + 2], gv[4 * j + 3]);
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
        for (int j = 0; j < KS; ++j) {
          const uint64_t d_hi = smem_desc(whi + (uint32_t)j * 2 * lbo, lbo, sbo);
          const uint64_t d_lo = smem_desc(wlo + (uint32_t)j * 2 * lbo, lbo, sbo);
          mma_tf32_ts(tmem + kColD, tmem + kColAlo + 8 * j, d_hi, idesc, j > 0);   // small terms first
          mma_tf32_ts(tmem + kColD, tmem + kColAhi + 8 * j, d_lo, idesc, 1);
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

int64_t lstm_tc_workspace(int kind, int64_t B, int64_t T, int64_t I, int64_t H) {
  (void)B; (void)T;
  const int64_t IP = round_up(I, 8), K = IP + H;
  if (kind == HY2DL_WS_LSTM_FWD) return sizeof(float) * (2 * K * 4 * H + 4 * H);
  return 16;
}

bool lstm_tc_supported(int64_t I, int64_t H) { return H == 64 && round_up(I, 8) <= 32; }

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
