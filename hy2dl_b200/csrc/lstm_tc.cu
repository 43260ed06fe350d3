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
#include <stdlib.h>

namespace hy2dl {
namespace tc {

constexpr int kH = 64;
constexpr int kN = 256;            // 4H
constexpr int kRows = 128;         // sequences per CTA
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColD = 0, kColHead = 256;                 // gates accumulator [0,256), fused-head accumulator [256,288)
constexpr uint32_t kColAhi = 288, kColAlo = 384;              // A operand hi / lo, K <= 96 columns each

// ---- weight packing for the tc path ------------------------------------------------------------
// Wt_hi / Wt_lo: [K/4][256][4] with n = 4u + g; k < IP: W_ih (zero padded), else W_hh.
// Two things are folded into the packed weights so that the epilogue gets exp2 arguments directly:
//   * every row is multiplied by -log2(e) (gates i, f, o) or -2 log2(e) (gate g): the accumulator is
//     z' with e^-z = 2^z' (sigmoid) and e^-2z = 2^z' (tanh);
//   * the bias (b_ih + b_hh, scaled the same way) sits in input column IP-1, which the forward kernel
//     feeds with the constant 1 (input_size <= 31 leaves that column free);
//   * the tensor core's fp32 accumulation truncates toward zero; over the 36 chained MMAs of a step the
//     pre-activations come out 2.75e-7 (relative, mean; std 2.8e-7 -- tools/dbg_zbias.py) too small, a
//     systematic shrink that compounds over hundreds of recurrent steps.  The packed weights carry the
//     factor 1 + 2.75e-7 that removes the mean of that error.
__global__ void lstm_tc_pack_w_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                      const float* __restrict__ b_ih, const float* __restrict__ b_hh, int I, int IP, int NT,
                                      float* __restrict__ w_hi, float* __restrict__ w_lo) {
  constexpr float kScaleIFO = (float)(-1.4426950408889634 * (1.0 + 2.75e-7));        // -log2(e) x truncation compensation
  constexpr float kScaleG = (float)(-2.0 * 1.4426950408889634 * (1.0 + 2.75e-7));
  const int K = IP + kH;
  const int total = K * kN;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kk = e & 3, n = (e >> 2) & 255, kc = e >> 10;
    const int k = kc * 4 + kk, u = n >> 2, g = n & 3, row = g * kH + u;
    float v;
    if (k < IP) v = k < I ? w_ih[row * I + k] : (k == IP - 1 ? b_ih[row] + b_hh[row] : 0.f);
    else v = w_hh[row * kH + (k - IP)];
    v *= (g == 2) ? kScaleG : kScaleIFO;
    uint32_t hi, lo;
    split_tf32_rr(v, hi, lo);
    const int o = (kc * NT + n) * 4 + kk;        // chunk kc holds NT = 256 (+ NH head) rows
    w_hi[o] = __uint_as_float(hi);
    w_lo[o] = __uint_as_float(lo);
  }
}

// Head Linear(H, C) + sigmoid range mapping fused into the recurrent kernels (C <= 32 columns).
// Column c of the head: kind 0 = dynamic parameter, 1 = static parameter, 2 = routing parameter.
constexpr int kHeadMax = 32;
struct HeadCols {
  int32_t C, NH;                 // columns, columns rounded up to 16 (MMA N)
  int32_t m, warmup;
  int64_t N;                     // B * m series
  int64_t off[kHeadMax];         // plane offset in par_sim
  float lo[kHeadMax], span[kHeadMax];
  int8_t kind[kHeadMax];
  int16_t par[kHeadMax], mem[kHeadMax];
};

// Head weights are NH extra rows (n = 256 ..) of the gates' B operand [K/4][256 + NH][4], rows >= C zero: the
// last epilogue set's MMAs simply run with N = 128 + NH, so the head pre-activations of time t-1 come out of
// step t's accumulation at no extra instruction.  The Linear bias sits in input column IP-1 (fed with 1),
// everything scaled by -log2(e) x truncation compensation.
__global__ void lstm_tc_pack_head_kernel(const float* __restrict__ w, const float* __restrict__ bias, int C, int NH, int IP,
                                         float* __restrict__ wh_hi, float* __restrict__ wh_lo) {
  constexpr float kScale = (float)(-1.4426950408889634 * (1.0 + 2.75e-7));
  const int K = IP + kH;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < K * NH; e += gridDim.x * blockDim.x) {
    const int kk = e & 3, n = (e >> 2) % NH, kc = e / (4 * NH), k = 4 * kc + kk;
    float v = 0.f;
    if (n < C) v = k < IP ? (k == IP - 1 ? bias[n] : 0.f) : w[n * kH + (k - IP)];
    v *= kScale;
    uint32_t hi, lo;
    split_tf32_rr(v, hi, lo);
    const int o = (kc * (kN + NH) + kN + n) * 4 + kk;
    wh_hi[o] = __uint_as_float(hi);
    wh_lo[o] = __uint_as_float(lo);
  }
}

struct TcFwdArgs {
  const float* x;       // RI32, F = IP = 32
  const float* w_hi;    // [K/4][256][4]
  const float* w_lo;
  float* h_seq;         // RI32 F = 64   (nullable)
  float* c_seq;         // RI32 F = 64   (nullable)
  float* gates;         // RI32 F = 256, n = 4u + g (nullable)
  float* h_last;        // row-major [B][64] (nullable)
  int64_t B, T, G;      // G = ceil(B / 32) groups
  int IP;
  int single;           // HY2DL_PREC_TF32: single-pass TF32 (the hi*hi MMAs only), the throughput mode
  // fused head (has_head == 0: none); its weights are rows 256.. of w_hi / w_lo
  int has_head;
  float* par_sim;
  float* par_warm;
  HeadCols hc;
};

// NSETS = 2 or 4 sets of four row warps: set s owns hidden units [s*64/NSETS, (s+1)*64/NSETS) of all 128
// rows (warp w of a set owns TMEM lanes / rows 32 (w & 3) ..), and the step's MMAs are issued per set
// (N = 256/NSETS gate columns each, committed separately), so the epilogue of set s runs while the tensor
// core is still busy with sets s+1.. -- and there are NSETS warps per scheduler to hide each other's latencies.
// The A operand [x_t | h_{t-1}] is shared by all MMAs of a step, so every set holds its new h values in
// registers until the LAST commit of the step before writing them into TMEM.
template <int NSETS>
__global__ void __launch_bounds__(128 * NSETS + 32, 1) lstm_tc_fwd_kernel(TcFwdArgs a) {
  constexpr int kUPT = kH / NSETS;          // hidden units per thread
  constexpr int kQ = kUPT / 8;              // 8-unit blocks per thread
  constexpr int kNS = kN / NSETS;           // gate columns per set
  constexpr int kMmaWarp = 4 * NSETS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int IP = a.IP, K = IP + kH, KC = K / 4;
  const bool head = a.has_head != 0;
  const int NH = head ? a.hc.NH : 0, NT = kN + NH;                     // B-operand rows per K chunk
  float* s_whi = reinterpret_cast<float*>(smem_raw);
  float* s_wlo = s_whi + K * NT;
  uint64_t* bar_a_ready = reinterpret_cast<uint64_t*>(s_wlo + K * NT);   // all row threads -> MMA warp
  uint64_t* bar_done = bar_a_ready + 1;                               // [NSETS] MMA commit of set s -> row threads
  uint64_t* bar_head = bar_done + NSETS;                              // head MMAs committed -> row threads
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_head + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kThreads = 128 * NSETS + 32;

  // ---- one-time setup: weights -> smem, barriers, TMEM
  {
    const float4* g_hi = reinterpret_cast<const float4*>(a.w_hi);
    const float4* g_lo = reinterpret_cast<const float4*>(a.w_lo);
    float4* d_hi = reinterpret_cast<float4*>(s_whi);
    float4* d_lo = reinterpret_cast<float4*>(s_wlo);
    for (int e = tid; e < KC * NT; e += kThreads) { d_hi[e] = __ldg(g_hi + e); d_lo[e] = __ldg(g_lo + e); }
  }
  if (tid == 0) {
    mbar_init(bar_a_ready, 128 * NSETS);
    for (int s = 0; s < NSETS; ++s) mbar_init(bar_done + s, 1);
    mbar_init(bar_head, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(s_tmem, kTmemCols);
  fence_async_smem();          // generic-proxy smem writes (weights) -> visible to the MMA's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < kMmaWarp) {
    // =========================== row-owning threads ===========================================
    const int set = warp >> 2;
    const int64_t gidx = (int64_t)blockIdx.x * 4 + (warp & 3);   // this warp's 32-row group
    const bool live = gidx < a.G;                                 // warp-uniform
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int kXU = 4 / NSETS;                                // x units (8 columns) this thread moves
    // this thread's row in the RI32 streams: base pointers at step 0 (advanced once per step) + the four
    // swizzled unit offsets of its row
    const int64_t row_off = (int64_t)gidx * 1024 + lane * 32;
    const float* xp = a.x + row_off;
    float* hp = a.h_seq ? a.h_seq + 2 * (int64_t)gidx * 1024 + lane * 32 : nullptr;
    float* cp = a.c_seq ? a.c_seq + 2 * (int64_t)gidx * 1024 + lane * 32 : nullptr;
    float* gp = a.gates ? a.gates + 8 * (int64_t)gidx * 1024 + lane * 32 : nullptr;
    const int64_t xs = a.G * 1024;                                // floats per step of a 32-column stream
    int swz[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) swz[j] = (j ^ (lane & 3)) << 3;
    auto put_x = [&](const f8& v, int U) {                        // x unit U -> A operand; column IP-1 carries the bias 1
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_tf32(v.v[e], hi[e], lo[e]);
      if (U == 3) { hi[7] = 0x3F800000u; lo[7] = 0u; }
      tmem_st8(lane_base + kColAhi + 8 * U, hi);
      tmem_st8(lane_base + kColAlo + 8 * U, lo);
    };
    // fused head: set 0 turns the head pre-activations of time th (already in registers) into parameter values
    const int64_t brow = gidx * 32 + lane;
    auto head_store = [&](const uint32_t (&z)[kHeadMax], int64_t th) {
      if (!live || brow >= a.B) return;
      const HeadCols& hc = a.hc;
#pragma unroll
      for (int cc = 0; cc < kHeadMax; ++cc) {
        if (cc < hc.C) {
          const float val = fmaf(hc.span[cc], fast_rcp(1.f + fast_ex2(__uint_as_float(z[cc]))), hc.lo[cc]);
          const int64_t n = brow * hc.m + hc.mem[cc];
          if (hc.kind[cc] == 0) {
            if (th == hc.warmup - 1) a.par_warm[(int64_t)hc.par[cc] * hc.N + n] = val;
            else a.par_sim[hc.off[cc] + (th - hc.warmup) * hc.N + n] = val;
          } else if (th == a.T - 1) {
            if (hc.kind[cc] == 2) a.par_sim[hc.off[cc] + brow] = val;
            else { a.par_sim[hc.off[cc] + n] = val; a.par_warm[(int64_t)hc.par[cc] * hc.N + n] = val; }
          }
        }
      }
    };
    auto head_load = [&](uint32_t (&z)[kHeadMax]) {          // D_head -> registers (must precede the next a_ready arrive)
      uint32_t v16[16];
      tmem_ld16(lane_base + kColHead, v16);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = v16[j];
      if (NH > 16) {
        tmem_ld16(lane_base + kColHead + 16, v16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) z[16 + j] = v16[j];
      }
    };
    float c[kUPT];
#pragma unroll
    for (int u = 0; u < kUPT; ++u) c[u] = 0.f;

    // A(t = 0) = [x_0 | 0]
    {
      uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int q = 0; q < kQ; ++q) {
        tmem_st8(lane_base + kColAhi + IP + 8 * (kQ * set + q), z);
        tmem_st8(lane_base + kColAlo + IP + 8 * (kQ * set + q), z);
      }
#pragma unroll
      for (int j = 0; j < kXU; ++j) {               // the 32 input columns = 4 units of 8
        const int U = kXU * set + j;
        f8 v = f8_zero();
        if (live) v = ldg256(xp + swz[U]);
        put_x(v, U);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_a_ready);
    }

    for (int64_t t = 0; t < a.T; ++t) {
      // prefetch x_{t+1} while the MMAs of step t run
      f8 xn[kXU];
      const bool more = t + 1 < a.T;
#pragma unroll
      for (int j = 0; j < kXU; ++j) {
        xn[j] = f8_zero();
        if (more && live) xn[j] = ldg256(xp + (t + 1) * xs + swz[kXU * set + j]);
      }
      mbar_wait(bar_done + set, (uint32_t)(t & 1));
      tc_fence_after();

      f8 hkeep[kQ];                                 // h_t of this thread's units, written to TMEM after the last commit
#pragma unroll
      for (int q = 0; q < kQ; ++q) {                // 8 hidden units (32 D columns) per iteration
        const int qg = kQ * set + q;                // unit block 0..7
        uint32_t v[32];
        tmem_ld32(lane_base + kColD + 32 * qg, v);
        tmem_ld_wait();
        f8 hv, cv, gv[4];
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          float gi, gf, gg, go;
          lstm_gates(__uint_as_float(v[4 * uu + 0]), __uint_as_float(v[4 * uu + 1]), __uint_as_float(v[4 * uu + 2]),
                     __uint_as_float(v[4 * uu + 3]), gi, gf, gg, go);
          const float cn = fmaf(gf, c[8 * q + uu], gi * gg);
          c[8 * q + uu] = cn;
          hv.v[uu] = go * tanh_mufu(cn);
          cv.v[uu] = cn;
          gv[uu >> 1].v[4 * (uu & 1) + 0] = gi; gv[uu >> 1].v[4 * (uu & 1) + 1] = gf;
          gv[uu >> 1].v[4 * (uu & 1) + 2] = gg; gv[uu >> 1].v[4 * (uu & 1) + 3] = go;
        }
        hkeep[q] = hv;
        if (live) {
          // one 256-bit store = one 32-byte sector of this row; the four gate units are its 128-byte line
          if (hp) stg256(hp + 2 * t * xs + (qg >> 2) * 1024 + swz[qg & 3], hv);
          if (cp) stg256(cp + 2 * t * xs + (qg >> 2) * 1024 + swz[qg & 3], cv);
          if (gp) {
#pragma unroll
            for (int j = 0; j < 4; ++j) stg256(gp + 8 * t * xs + qg * 1024 + swz[j], gv[j]);
          }
          if (a.h_last && t == a.T - 1) {
            const int64_t b = gidx * 32 + lane;
            if (b < a.B) {
              *reinterpret_cast<float4*>(a.h_last + b * kH + 8 * qg) = make_float4(hv.v[0], hv.v[1], hv.v[2], hv.v[3]);
              *reinterpret_cast<float4*>(a.h_last + b * kH + 8 * qg + 4) = make_float4(hv.v[4], hv.v[5], hv.v[6], hv.v[7]);
            }
          }
        }
      }
      // a head round (pre-activations of time t-1) was issued with this step's MMAs when t-1 >= warmup-1
      const bool head_round = head && t >= a.hc.warmup;
      uint32_t zh[kHeadMax];
      if (more || head) {
        // the A operand may be overwritten once EVERY MMA of this step has retired
        if (set != NSETS - 1) {
          mbar_wait(bar_done + (NSETS - 1), (uint32_t)(t & 1));
          tc_fence_after();
        }
        if (head_round && set == 0) head_load(zh);   // rides on the last set's accumulation (columns 256..)
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int uu = 0; uu < 8; ++uu) split_tf32(hkeep[q].v[uu], hi[uu], lo[uu]);
          tmem_st8(lane_base + kColAhi + IP + 8 * (kQ * set + q), hi);
          tmem_st8(lane_base + kColAlo + IP + 8 * (kQ * set + q), lo);
        }
        if (more) {
#pragma unroll
          for (int j = 0; j < kXU; ++j) put_x(xn[j], kXU * set + j);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_a_ready);
      }
      if (head_round && set == 0) head_store(zh, t - 1);
    }
    if (head) {                                     // last round: pre-activations of time T-1
      mbar_wait(bar_head, 0);
      tc_fence_after();
      if (set == 0) {
        uint32_t zh[kHeadMax];
        head_load(zh);
        head_store(zh, a.T - 1);
      }
    }
  } else {
    // =========================== MMA issuer (one thread) ======================================
    if (elect_one()) {
      const uint32_t whi = smem_u32(s_whi), wlo = smem_u32(s_wlo);
      const uint32_t sbo = 128;                    // K-major no-swizzle: 8-row groups 128 B apart
      const int KS = K / 8;
      const uint32_t lbo_t = (uint32_t)NT * 16;    // chunk-to-chunk stride of the [K/4][NT][4] operand
      for (int64_t t = 0; t < a.T; ++t) {
        mbar_wait(bar_a_ready, (uint32_t)(t & 1));
        tc_fence_after();
        for (int s = 0; s < NSETS; ++s) {
          // the last set also accumulates the NH head columns (D columns 256..)
          const uint32_t id = idesc_tf32(kRows, kNS + (s == NSETS - 1 ? NH : 0), 0, 0);
          const uint32_t d = tmem + kColD + kNS * s, boff = (uint32_t)(kNS * s) * 16;   // B rows n = kNS*s ..
          // all correction terms first, while the accumulator is still ~2^-11 of its final size: the
          // tensor core's fp32 accumulation loses up to an ulp of the ACCUMULATOR per MMA, so only the
          // K/8 hi*hi instructions round at full magnitude
          if (!a.single)
            for (int j = 0; j < KS; ++j) {
              const uint64_t d_hi = smem_desc(whi + boff + (uint32_t)j * 2 * lbo_t, lbo_t, sbo);
              const uint64_t d_lo = smem_desc(wlo + boff + (uint32_t)j * 2 * lbo_t, lbo_t, sbo);
              mma_tf32_ts(d, tmem + kColAlo + 8 * j, d_hi, id, j > 0);
              mma_tf32_ts(d, tmem + kColAhi + 8 * j, d_lo, id, 1);
            }
          for (int j = 0; j < KS; ++j) {
            const uint64_t d_hi = smem_desc(whi + boff + (uint32_t)j * 2 * lbo_t, lbo_t, sbo);
            mma_tf32_ts(d, tmem + kColAhi + 8 * j, d_hi, id, !a.single || j > 0);
          }
          mma_commit(bar_done + s);
        }
      }
      if (head) {     // pre-activations of the last time step: one head-only round on A = [. | h_{T-1}]
        mbar_wait(bar_a_ready, (uint32_t)(a.T & 1));
        tc_fence_after();
        const uint32_t id = idesc_tf32(kRows, NH, 0, 0), d = tmem + kColHead, boff = (uint32_t)kN * 16;
        if (!a.single)
          for (int j = 0; j < KS; ++j) {
            const uint64_t d_hi = smem_desc(whi + boff + (uint32_t)j * 2 * lbo_t, lbo_t, sbo);
            const uint64_t d_lo = smem_desc(wlo + boff + (uint32_t)j * 2 * lbo_t, lbo_t, sbo);
            mma_tf32_ts(d, tmem + kColAlo + 8 * j, d_hi, id, j > 0);
            mma_tf32_ts(d, tmem + kColAhi + 8 * j, d_lo, id, 1);
          }
        for (int j = 0; j < KS; ++j) {
          const uint64_t d_hi = smem_desc(whi + boff + (uint32_t)j * 2 * lbo_t, lbo_t, sbo);
          mma_tf32_ts(d, tmem + kColAhi + 8 * j, d_hi, id, !a.single || j > 0);
        }
        mma_commit(bar_head);
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

static size_t tc_fwd_smem(int K, int NH) { return (size_t)K * (kN + NH) * 4 * 2 + 128; }

// column table of the fused head from the parameter map (same column order as the unfused head kernels:
// conceptual parameter p, member j -> column p*m + j; routing parameters follow)
int head_cols_from_parmap(const hy2dl_parmap* pm, int64_t B, HeadCols* hc) {
  const int C = pm->n_par * pm->n_members + pm->n_routing;
  if (C > kHeadMax) return -1;
  hc->C = C; hc->NH = C <= 16 ? 16 : 32;
  hc->m = pm->n_members; hc->warmup = pm->warmup; hc->N = B * pm->n_members;
  for (int p = 0; p < pm->n_par + pm->n_routing; ++p) {
    const bool routing = p >= pm->n_par;
    const int members = routing ? 1 : pm->n_members;
    for (int j = 0; j < members; ++j) {
      const int c = routing ? pm->n_par * pm->n_members + (p - pm->n_par) : p * pm->n_members + j;
      hc->off[c] = pm->sim_off[p];
      hc->lo[c] = pm->lo[p];
      hc->span[c] = pm->hi[p] - pm->lo[p];
      hc->kind[c] = routing ? 2 : (pm->dynamic[p] ? 0 : 1);
      hc->par[c] = (int16_t)p;
      hc->mem[c] = (int16_t)j;
    }
  }
  return 0;
}
constexpr int kFwdSets = 2;   // measured on B200 (B = 18944, T = 365): 2 sets 3.01 ms, 4 sets 3.40 ms, 1 set 4.06 ms

int64_t lstm_tc_wgrad_workspace(int64_t B, int64_t T, int64_t I);

int64_t lstm_tc_workspace(int kind, int64_t B, int64_t T, int64_t I, int64_t H) {
  (void)I;
  const int64_t IP = 32, K = IP + H;      // the RI32 input stream is always 32 columns wide
  if (kind == HY2DL_WS_LSTM_FWD) return sizeof(float) * (2 * K * 4 * H + 2 * K * kHeadMax);   // packed W (+ packed head)
  if (kind == HY2DL_WS_LSTM_BWD) return sizeof(float) * (2 * 4 * H * H + 2 * kHeadMax * H);   // packed W_hh (+ packed head)
  if (kind == HY2DL_WS_LSTM_WGRAD) return lstm_tc_wgrad_workspace(B, T, I);
  return 16;
}

bool lstm_tc_supported(int64_t I, int64_t H) { return H == 64 && I <= 31; }   // column 31 of the input stream carries the bias

int lstm_tc_fwd(const float* x_ri, int64_t B, int64_t T, int64_t I, int64_t IP, const float* w_ih, const float* w_hh,
                const float* b_ih, const float* b_hh, float* h_seq, float* c_seq, float* gates, float* h_last, void* ws,
                int64_t ws_bytes, const hy2dl_head_fwd* head, int single, cudaStream_t s) {
  const int K = (int)(IP + kH);
  const int64_t need = lstm_tc_workspace(HY2DL_WS_LSTM_FWD, B, T, I, kH);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_fwd(tf32x3): workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  TcFwdArgs a{};
  int NH = 0;
  if (head) {
    HY_REQUIRE(head->w && head->bias && head->pm && head->par_sim && head->par_warm, HY2DL_ERR_ALIGN, "hy2dl_lstm_fwd: incomplete head description");
    HY_REQUIRE(head->pm->warmup >= 1 && head->pm->warmup < T, HY2DL_ERR_SHAPE, "hy2dl_lstm_fwd: head warmup=%d must satisfy 1 <= warmup < T=%lld",
               head->pm->warmup, (long long)T);
    HY_REQUIRE(head_cols_from_parmap(head->pm, B, &a.hc) == 0, HY2DL_ERR_SHAPE,
               "hy2dl_lstm_fwd: the fused head takes at most %d columns (use hy2dl_head_map_fwd)", kHeadMax);
    NH = a.hc.NH;
  }
  const int NT = kN + NH;
  float* w_hi = reinterpret_cast<float*>(ws);
  float* w_lo = w_hi + K * NT;
  lstm_tc_pack_w_kernel<<<96, 256, 0, s>>>(w_ih, w_hh, b_ih, b_hh, (int)I, (int)IP, NT, w_hi, w_lo);
  if (head) {
    lstm_tc_pack_head_kernel<<<16, 256, 0, s>>>(head->w, head->bias, a.hc.C, NH, (int)IP, w_hi, w_lo);
    a.has_head = 1; a.par_sim = head->par_sim; a.par_warm = head->par_warm;
  }
  a.x = x_ri;
  a.w_hi = w_hi; a.w_lo = w_lo;
  a.h_seq = h_seq; a.c_seq = c_seq; a.gates = gates; a.h_last = h_last;
  a.B = B; a.T = T; a.G = cdiv(B, 32); a.IP = (int)IP; a.single = single;
  const size_t smem = tc_fwd_smem(K, NH);
  static const int sets = [] { const char* e = getenv("HY2DL_FWD_SETS"); return e ? (atoi(e) == 4 ? 4 : 2) : kFwdSets; }();   // tuning knob
  if (sets == 2) {
    cudaFuncSetAttribute(lstm_tc_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lstm_tc_fwd_kernel<2><<<(unsigned)cdiv(B, kRows), 128 * 2 + 32, smem, s>>>(a);
  } else {
    cudaFuncSetAttribute(lstm_tc_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lstm_tc_fwd_kernel<4><<<(unsigned)cdiv(B, kRows), 128 * 4 + 32, smem, s>>>(a);
  }
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
// quarters live in a 2-slot TMEM ring, the dh accumulators are double buffered across steps.
// The tensor core's fp32 accumulation truncates (-2.8e-8 relative per accumulating MMA, tools/dbg_rz.py), and a
// bias in dh_rec compounds coherently over the recurrence.  So the 96 MMAs of a step do NOT share one
// accumulator: quarters {0,1} and {2,3} accumulate separately, each quarter issues its 16 correction MMAs
// before its 8 hi*hi ones, and the row threads add the two partial sums in registers (round to nearest).
// TMEM columns: step buffer b at [128 b, 128 b + 128): accumulators of quarters {0,1} | {2,3};
//               slot0 [256,384) slot1 [384,512)  (hi = first 64, lo = last 64).
// ================================================================================================
namespace hy2dl {
namespace tc {

constexpr int kBwdRowWarps = 8;        // two sets of four row warps: set s owns hidden units 8Q + 4s .. 8Q + 4s + 3 of every block Q
constexpr int kBwdThreads = 32 * kBwdRowWarps + 32;   // + MMA warp
constexpr int kBwdThreadsHead = kBwdThreads + 96;     // + 3 warps that turn dpar into the head's dz operand
constexpr int kDzThreads = 96;
constexpr uint32_t kColDh0 = 0, kColSlot0 = 256;

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

// W_head^T as the B operand of dh_head = dz * W_head: B[n = k_out][k = head column], K-major no-swizzle
// [NH/4][64][4]; columns >= C zero.
__global__ void lstm_tc_pack_headT_kernel(const float* __restrict__ w, int C, int NH, float* __restrict__ hi_out, float* __restrict__ lo_out) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < NH * kH; e += gridDim.x * blockDim.x) {
    const int kk = e & 3, n = (e >> 2) & 63, kc = e >> 8, k = 4 * kc + kk;
    uint32_t hi, lo;
    split_tf32_rr(k < C ? w[k * kH + n] : 0.f, hi, lo);
    hi_out[e] = __uint_as_float(hi);
    lo_out[e] = __uint_as_float(lo);
  }
}

struct TcBwdArgs {
  const float* w_hi;      // [64][64][4] packed W_hh (hi)
  const float* w_lo;
  const float* gates;     // RI32 F = 256 (n = 4u + g)
  const float* c_seq;     // RI32 F = 64
  const float* dh_seq;    // RI32 F = 64 or null
  const float* dh_last;   // row-major [B][64] or null
  float* dG;              // RI32 F = 256; may alias gates
  int64_t B, T, G, dh_t0;
  int single;             // HY2DL_PREC_TF32: hi*hi MMAs only
  // fused head adjoint (has_head == 0: none): dh_ext(t) = dz_t * W_head with dz from dpar_sim / par_sim
  long long* stamps;      // -DHY2DL_US_STAMPS: clock64 timeline of CTA 0, reverse steps 100-101
  int has_head;
  const float* wh_hi;     // [NH/4][64][4]
  const float* wh_lo;
  const float* par_sim;
  const float* dpar_sim;
  float* dz_out;          // RI32 F = 32 [T][G][32 rows][32]: dz of time th (columns >= C untouched), for the weight gradient
  HeadCols hc;
};

struct BwdChunk {   // tape of 4 hidden units of one row at one step: half a 128-byte line of gates + three half sectors
  f8 g[2];
  float4 ct, cp, dh;
};

// chunk (Q, set): hidden units [8Q + 4 set, 8Q + 4 set + 4) = 32-byte units 4Q + 2 set, 4Q + 2 set + 1 of the gates stream
// (unit j holds 2 hidden units x 4 gates) and half `set` of unit Q of the c / dh streams
__device__ __forceinline__ void bwd_load_chunk(const TcBwdArgs& a, int64_t t, int Q, int set, int64_t gidx, int lane, bool live,
                                               BwdChunk& c) {
  const bool on = live && t >= 0;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 2; ++j) c.g[j] = on ? ldg256(a.gates + ri32_unit(t, a.G, gidx, 8, 4 * Q + 2 * set + j, lane)) : f8_zero();
  c.ct = on ? __ldg(reinterpret_cast<const float4*>(a.c_seq + ri32_unit(t, a.G, gidx, 2, Q, lane) + 4 * set)) : z4;
  c.cp = (on && t > 0) ? __ldg(reinterpret_cast<const float4*>(a.c_seq + ri32_unit(t - 1, a.G, gidx, 2, Q, lane) + 4 * set)) : z4;
  c.dh = (on && a.dh_seq && t >= a.dh_t0) ? __ldg(reinterpret_cast<const float4*>(a.dh_seq + ri32_unit(t, a.G, gidx, 2, Q, lane) + 4 * set)) : z4;
  if (on && a.dh_last && t == a.T - 1) {
    const int64_t b = gidx * 32 + lane;
    if (b < a.B) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(a.dh_last + b * kH + 8 * Q + 4 * set));
      c.dh.x += v0.x; c.dh.y += v0.y; c.dh.z += v0.z; c.dh.w += v0.w;
    }
  }
}

#ifdef HY2DL_US_STAMPS
#define TB_STAMP(id) do { if (a.stamps && blockIdx.x == 0 && s >= 100 && s < 102) a.stamps[(s - 100) * 32 + (id)] = clock64(); } while (0)
#else
#define TB_STAMP(id) do { } while (0)
#endif
__global__ void __launch_bounds__(kBwdThreadsHead, 1) lstm_tc_bwd_kernel(TcBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const bool head = a.has_head != 0;
  const int NH = head ? a.hc.NH : 0;
  // one dz operand buffer (hi | lo): its MMAs open the chain, so it is free again for most of the step -- and
  // shared memory is precious here: the tape loads need L1 (228 KB minus the carve-out) for their in-flight
  // lines; measured with the same kernel: 164 KB of shared memory 3.8 ms, 196 KB 4.3 ms, 204+ KB 5.8 ms
  constexpr int n_dzbuf = 1;
  float* s_whi = reinterpret_cast<float*>(smem_raw);
  float* s_wlo = s_whi + kN * kH;
  float* s_dc = s_wlo + kN * kH;                                        // [64][128] cell-state adjoint, per row thread
  float* s_hw = s_dc + kH * kRows;                                      // head: W_head^T hi | lo, [NH/4][64][4] each
  float* s_dz = s_hw + 2 * NH * kH;                                     // head: n_dzbuf x (hi | lo) [NH/4][128 rows][4]
  uint64_t* bar_ready = reinterpret_cast<uint64_t*>(s_dz + (head ? n_dzbuf * 2 * NH * kRows : 0));  // [2] rows -> MMA, per slot
  uint64_t* bar_free = bar_ready + 2;                                   // [2] MMA commit -> rows, per slot
  uint64_t* bar_dz_full = bar_free + 2;                                 // [2] dz warps -> MMA
  uint64_t* bar_dz_free = bar_dz_full + 2;                              // [2] MMA commit -> dz warps
  uint64_t* bar_pre = bar_dz_free + 2;                                  // dh(T-1) = dz_{T-1} W_head ready -> rows
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_pre + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nthreads = blockDim.x;

  {
    const float4* g_hi = reinterpret_cast<const float4*>(a.w_hi);
    const float4* g_lo = reinterpret_cast<const float4*>(a.w_lo);
    float4* d_hi = reinterpret_cast<float4*>(s_whi);
    float4* d_lo = reinterpret_cast<float4*>(s_wlo);
    for (int e = tid; e < kN * kH / 4; e += nthreads) { d_hi[e] = __ldg(g_hi + e); d_lo[e] = __ldg(g_lo + e); }
    if (head) {
      for (int e = tid; e < NH * kH; e += nthreads) { s_hw[e] = __ldg(a.wh_hi + e); s_hw[NH * kH + e] = __ldg(a.wh_lo + e); }
      for (int e = tid; e < n_dzbuf * 2 * NH * kRows; e += nthreads) s_dz[e] = 0.f;     // padding columns stay zero
    }
  }
  if (tid == 0) {
    mbar_init(bar_ready + 0, 32 * kBwdRowWarps); mbar_init(bar_ready + 1, 32 * kBwdRowWarps);
    mbar_init(bar_free + 0, 1); mbar_init(bar_free + 1, 1);
    mbar_init(bar_dz_full + 0, kDzThreads); mbar_init(bar_dz_full + 1, kDzThreads);
    mbar_init(bar_dz_free + 0, 1); mbar_init(bar_dz_free + 1, 1);
    mbar_init(bar_pre, 1);
    mbar_fence_init();
  }
  if (warp == kBwdRowWarps) tmem_alloc(s_tmem, kTmemCols);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < kBwdRowWarps) {
    // two warps per scheduler (set 0 / set 1 of the same 32 rows) hide each other's latencies: the chunk loop is
    // a chain of tape loads, TMEM loads, ~40 dependent flops per unit, stores and TMEM stores
    const int set = warp >> 2, wq = warp & 3;
    const int64_t gidx = (int64_t)blockIdx.x * 4 + wq;
    const bool live = gidx < a.G;
    const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
    float* dcp = s_dc + wq * 32 + lane;     // dc[u] at dcp[u * 128]: conflict-free
    for (int u = 4 * set; u < kH; u += 8)
      for (int uu = 0; uu < 4; ++uu) dcp[(u + uu) * kRows] = 0.f;
    if (head && set == 0) {      // dh(T-1) comes from the head alone: only accumulator {0,1} of buffer 1 is written
      uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int cc = 0; cc < 64; cc += 8) tmem_st8(lane_base + kColDh0 + 128 + 64 + cc, z);
      tmem_st_wait();
      tc_fence_before();
    }
    if (head) asm volatile("bar.sync 1, 256;" ::: "memory");   // row warps only: set 1 also reads those columns at s = 0
    if (head) tc_fence_after();

    // tape chunks (4 hidden units = 112 B per row) travel through a 2-deep register ring: the chunk of block Q + 2
    // is requested as soon as the one of block Q has been consumed.  Gate loads are whole sectors; the two sets
    // share the c / dh sectors of a block (16 bytes each, requested at about the same time).
    BwdChunk ring[2];
    bwd_load_chunk(a, a.T - 1, 0, set, gidx, lane, live, ring[0]);
    bwd_load_chunk(a, a.T - 1, 1, set, gidx, lane, live, ring[1]);
    for (int64_t s = 0; s < a.T; ++s) {
      const int64_t t = a.T - 1 - s;
      const uint32_t d_prev = lane_base + kColDh0 + 128 * (uint32_t)((s & 1) ^ 1);  // dh_rec partial sums from step s-1
      if (s > 0) {
        mbar_wait(bar_free + 1, 1);     // last quarter of step s-1 committed => dh_rec complete, both slots free
        tc_fence_after();
      } else if (head) {
        mbar_wait(bar_pre, 0);
        tc_fence_after();
      }
      const bool have_prev = s > 0 || head;
      if (tid == 0) TB_STAMP(0);
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {      // quarter k = 16 hidden units = 64 gate columns, lives in slot k & 1
        const int slot = k & 1;
        if (k >= 2) {                    // the second use of a slot in a step waits for its first MMAs
          mbar_wait(bar_free + slot, 0);
          tc_fence_after();
        }
        if (tid == 0) TB_STAMP(1 + 4 * k);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) { // blocks Q = 2k, 2k+1; this thread's 4 hidden units of each
          const int Q = 2 * k + kk;
          BwdChunk& cur = ring[kk];
          uint32_t r[4] = {0, 0, 0, 0};
          if (have_prev) {
            uint32_t r2[4];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                         : "r"(d_prev + 8 * Q + 4 * set));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3])
                         : "r"(d_prev + 64 + 8 * Q + 4 * set));
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 4; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
          }
          const float dhe[4] = {cur.dh.x, cur.dh.y, cur.dh.z, cur.dh.w};
          const float ctv[4] = {cur.ct.x, cur.ct.y, cur.ct.z, cur.ct.w};
          const float cpv[4] = {cur.cp.x, cur.cp.y, cur.cp.z, cur.cp.w};
          f8 dg[2];
#pragma unroll
          for (int uu = 0; uu < 4; ++uu) {
            const int u = 8 * Q + 4 * set + uu, j = uu >> 1, o = 4 * (uu & 1);
            const float gi = cur.g[j].v[o], gf = cur.g[j].v[o + 1], gg = cur.g[j].v[o + 2], go = cur.g[j].v[o + 3];
            const float dh = dhe[uu] + __uint_as_float(r[uu]);
            const float tcv = fast_tanh(ctv[uu]);
            const float dct = dcp[u * kRows] + dh * go * (1.f - tcv * tcv);
            dg[j].v[o + 0] = dct * gg * gi * (1.f - gi);
            dg[j].v[o + 1] = dct * cpv[uu] * gf * (1.f - gf);
            dg[j].v[o + 2] = dct * gi * (1.f - gg * gg);
            dg[j].v[o + 3] = dh * tcv * go * (1.f - go);
            dcp[u * kRows] = dct * gf;
          }
          // refill this ring slot: block Q + 2 of this step, or block kk of the next (earlier) step
          if (k < 3) bwd_load_chunk(a, t, Q + 2, set, gidx, lane, live, cur);
          else bwd_load_chunk(a, t - 1, kk, set, gidx, lane, live, cur);
          if (live) {
#pragma unroll
            for (int j = 0; j < 2; ++j) stg256(a.dG + ri32_unit(t, a.G, gidx, 8, 4 * Q + 2 * set + j, lane), dg[j]);
          }
          const uint32_t sbase = lane_base + kColSlot0 + 128 * slot + 32 * kk + 16 * set;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_tf32(dg[j].v[e], hi[e], lo[e]);
            tmem_st8(sbase + 8 * j, hi);
            tmem_st8(sbase + 64 + 8 * j, lo);
          }
          if (tid == 0) TB_STAMP(2 + 4 * k + kk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_ready + slot);
        if (tid == 0) TB_STAMP(4 + 4 * k);
      }
    }
  } else if (warp == kBwdRowWarps) {
    if (elect_one()) {
      const uint32_t idesc = idesc_tf32(kRows, kH, 0, 0);
      const uint32_t whi = smem_u32(s_whi), wlo = smem_u32(s_wlo);
      const uint32_t lbo = kH * 16, sbo = 128;
      // head: dh_ext(th) = dz_th [128 x NH] * W_head [NH x 64], A and B from shared memory, opens the chain of time th
      const uint32_t hw_hi = smem_u32(s_hw), hw_lo = hw_hi + (uint32_t)NH * kH * 4;
      const uint32_t dz_buf_bytes = 2u * (uint32_t)NH * kRows * 4, dz_lo_off = (uint32_t)NH * kRows * 4;
      int64_t dz_i = 0;
      auto issue_dz = [&](uint32_t d) {
        const int buf = (int)(dz_i % n_dzbuf);
        mbar_wait(bar_dz_full + buf, (uint32_t)((dz_i / n_dzbuf) & 1));
        tc_fence_after();
        const uint32_t a_hi = smem_u32(s_dz) + buf * dz_buf_bytes, a_lo = a_hi + dz_lo_off;
        if (!a.single)
          for (int j = 0; j < NH / 8; ++j) {
            const uint64_t ah = smem_desc(a_hi + j * 2 * 2048, 2048, 128), al = smem_desc(a_lo + j * 2 * 2048, 2048, 128);
            const uint64_t bh = smem_desc(hw_hi + j * 2 * lbo, lbo, sbo), bl = smem_desc(hw_lo + j * 2 * lbo, lbo, sbo);
            mma_tf32_ss(d, al, bh, idesc, j > 0);
            mma_tf32_ss(d, ah, bl, idesc, 1);
          }
        for (int j = 0; j < NH / 8; ++j) {
          const uint64_t ah = smem_desc(a_hi + j * 2 * 2048, 2048, 128), bh = smem_desc(hw_hi + j * 2 * lbo, lbo, sbo);
          mma_tf32_ss(d, ah, bh, idesc, !a.single || j > 0);
        }
        mma_commit(bar_dz_free + buf);
        ++dz_i;
      };
      if (head) {                                  // chain of time T-1: the head term only
        issue_dz(tmem + kColDh0 + 128);
        mma_commit(bar_pre);
      }
      for (int64_t s = 0; s < a.T; ++s) {
        const int64_t th = a.T - 2 - s;            // this chain accumulates dh of time th
        const bool has_dz = head && th >= a.hc.warmup;
        for (int k = 0; k < 4; ++k) {
          const int slot = k & 1;
          const uint32_t d = tmem + kColDh0 + 128 * (uint32_t)(s & 1) + 64 * (uint32_t)(k >> 1);   // accumulator of quarters {0,1} / {2,3}
          mbar_wait(bar_ready + slot, (uint32_t)(k >> 1));
          tc_fence_after();
          TB_STAMP(20 + k);
          if (k == 0 && has_dz) issue_dz(d);       // (the row threads have finished reading this buffer by now)
          const uint32_t abase = tmem + kColSlot0 + 128 * slot;
          const bool acc0 = k == 0 && has_dz;
          if (!a.single) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {          // correction terms first (see header)
              const uint32_t J = 8 * k + j;
              const uint64_t d_hi = smem_desc(whi + J * 2 * lbo, lbo, sbo);
              const uint64_t d_lo = smem_desc(wlo + J * 2 * lbo, lbo, sbo);
              mma_tf32_ts(d, abase + 64 + 8 * j, d_hi, idesc, (((k & 1) | j) != 0) || acc0);
              mma_tf32_ts(d, abase + 8 * j, d_lo, idesc, 1);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint64_t d_hi = smem_desc(whi + (uint32_t)(8 * k + j) * 2 * lbo, lbo, sbo);
            mma_tf32_ts(d, abase + 8 * j, d_hi, idesc, !a.single || (((k & 1) | j) != 0) || acc0);
          }
          mma_commit(bar_free + slot);
          TB_STAMP(24 + k);
        }
      }
    }
    __syncwarp();
  } else if (head) {
    // =========================== dz warps ========================================================
    // dz[row][c] = dpar * dval/dz (sigma recovered from the stored parameter value), for the steps that feed the
    // head: dynamic columns on th >= warmup, static + routing columns on th = T-1 (baseconceptualmodel.py:61-67)
    const int dt = tid - kBwdThreads;             // 0 .. 95
    const HeadCols& hc = a.hc;
    const int64_t row0 = (int64_t)blockIdx.x * kRows;
    const int n_items = kRows * hc.C;
    for (int64_t th = a.T - 1, i = 0; th >= hc.warmup; --th, ++i) {
      const int buf = (int)(i % n_dzbuf);
      if (i >= n_dzbuf) mbar_wait_relaxed(bar_dz_free + buf, (uint32_t)((i / n_dzbuf - 1) & 1));
      float* dz_hi = s_dz + buf * 2 * NH * kRows;
      float* dz_lo = dz_hi + NH * kRows;
      // batches of 16 items per thread: all loads of a batch are issued before the first use (one HBM latency
      // per batch instead of one per item)
#pragma unroll 1
      for (int bb = 0; bb < 3; ++bb) {             // 3 x 16 x 96 >= 128 rows x 32 columns
        const int e0 = dt + bb * 16 * kDzThreads;
        if (e0 >= n_items) break;
        float val[16], dp[16];
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
          const int e = e0 + i2 * kDzThreads;
          val[i2] = 0.f; dp[i2] = 0.f;
          if (e < n_items) {
            const int c = e >> 7, r = e & 127;
            const int64_t b = row0 + r;
            const int kind = hc.kind[c];
            if (b < a.B && (kind == 0 || th == a.T - 1)) {
              const int64_t idx = hc.off[c] + (kind == 0 ? (th - hc.warmup) * hc.N : 0) + (kind == 2 ? b : b * hc.m + hc.mem[c]);
              val[i2] = __ldg(a.par_sim + idx);
              dp[i2] = __ldg(a.dpar_sim + idx);
            }
          }
        }
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
          const int e = e0 + i2 * kDzThreads;
          if (e < n_items) {
            const int c = e >> 7, r = e & 127;
            const float lo = hc.lo[c], sp = hc.span[c];
            const float dz = dp[i2] * (val[i2] - lo) * ((lo + sp) - val[i2]) / sp;    // dp = 0 where the column has no gradient
            uint32_t hi, lw;
            split_tf32(dz, hi, lw);
            const int o = ((c >> 2) * kRows + r) * 4 + (c & 3);
            dz_hi[o] = __uint_as_float(hi);
            dz_lo[o] = __uint_as_float(lw);
            // for the weight-gradient GEMM: row r of this tile is row r & 31 of group 4 blockIdx + r / 32
            const int64_t g = (int64_t)blockIdx.x * 4 + (r >> 5);
            if (a.dz_out && g < a.G)
              a.dz_out[(th * a.G + g) * 1024 + (r & 31) * 32 + ((((c >> 3) ^ r) & 3) << 3) + (c & 7)] = dz;
          }
        }
      }
      fence_async_smem();
      mbar_arrive(bar_dz_full + buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBwdRowWarps) tmem_dealloc(tmem, kTmemCols);
}

int lstm_tc_bwd(int64_t B, int64_t T, const float* w_hh, const float* gates, const float* c_seq, const float* dh_seq,
                int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, const hy2dl_head_bwd* head,
                int single, cudaStream_t s) {
  const int64_t need = lstm_tc_workspace(HY2DL_WS_LSTM_BWD, B, T, 1, kH);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_bwd(tf32x3): workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  float* w_hi = reinterpret_cast<float*>(ws);
  float* w_lo = w_hi + kN * kH;
  lstm_tc_pack_whh_kernel<<<64, 256, 0, s>>>(w_hh, w_hi, w_lo);
  TcBwdArgs a{};
  a.w_hi = w_hi; a.w_lo = w_lo;
  a.gates = gates; a.c_seq = c_seq; a.dh_seq = dh_seq; a.dh_last = dh_last; a.dG = dG;
  a.B = B; a.T = T; a.G = cdiv(B, 32); a.dh_t0 = dh_t0; a.single = single;
  size_t smem = (size_t)kN * kH * 4 * 2 + (size_t)kH * kRows * 4 + 128;
  int threads = kBwdThreads;
  if (head) {
    HY_REQUIRE(head->w && head->pm && head->par_sim && head->dpar_sim, HY2DL_ERR_ALIGN, "hy2dl_lstm_bwd: incomplete head description");
    HY_REQUIRE(head->pm->warmup >= 1 && head->pm->warmup < T, HY2DL_ERR_SHAPE, "hy2dl_lstm_bwd: head warmup=%d must satisfy 1 <= warmup < T=%lld",
               head->pm->warmup, (long long)T);
    HY_REQUIRE(head_cols_from_parmap(head->pm, B, &a.hc) == 0, HY2DL_ERR_SHAPE,
               "hy2dl_lstm_bwd: the fused head takes at most %d columns (use hy2dl_head_map_bwd)", kHeadMax);
    const int NH = a.hc.NH;
    float* wh_hi = w_lo + kN * kH;
    float* wh_lo = wh_hi + NH * kH;
    lstm_tc_pack_headT_kernel<<<8, 256, 0, s>>>(head->w, a.hc.C, NH, wh_hi, wh_lo);
    a.has_head = 1; a.wh_hi = wh_hi; a.wh_lo = wh_lo; a.par_sim = head->par_sim; a.dpar_sim = head->dpar_sim;
    a.dz_out = head->dz;
    smem += (size_t)2 * NH * kH * 4 + (size_t)2 * NH * kRows * 4;
    threads = kBwdThreadsHead;
  }
  cudaFuncSetAttribute(lstm_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#ifdef HY2DL_US_STAMPS
  cudaMalloc(&a.stamps, 64 * sizeof(long long)); cudaMemset(a.stamps, 0, 64 * sizeof(long long));
#endif
  lstm_tc_bwd_kernel<<<(unsigned)cdiv(B, kRows), threads, smem, s>>>(a);
#ifdef HY2DL_US_STAMPS
  {
    long long h[64];
    cudaDeviceSynchronize();
    cudaMemcpy(h, a.stamps, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(a.stamps);
    for (int i = 0; i < 64; ++i) if (h[i]) fprintf(stderr, "TB_STAMP step %d id %2d  %8lld\n", 100 + i / 32, i % 32, h[i] - h[0]);
  }
#endif
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

}  // namespace tc
}  // namespace hy2dl
