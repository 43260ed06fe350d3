// Back-propagation through time for H = 64 / 128 / 256 at SMALL batches with the hidden units of a 128-sequence tile split
// over S = H / 16 co-resident CTAs; the counterpart of lstm_us_tc_fwd.cu (read its header first).
//
// CTA s owns hidden units [16 s, 16 s + 16) of the tile.  Per reverse step it
//   * forms the gate gradients of ITS units (4 x 16 columns of dG) from the tapes, dc (registers) and
//     dh = dh_ext + dh_rec, writes them to the dG tape and, split hi / lo, to shared memory as the A operand [128 x 64];
//   * multiplies them by its 64 rows of W_hh (hi + lo resident in shared memory for the whole kernel, 128 KB at
//     H = 256): a PARTIAL dh_rec[128 x H] over its slice of the contraction, 24 MMAs (M = 128, N = H), the correction terms
//     first and then one chain of 8 hi*hi MMAs into the same accumulator (lstm_rm_tc_fwd.cu explains the order);
//   * pulls the partial and writes it to the exchange buffer in L2; a release / acquire counter per group says when
//     all S partials of a step are there; every CTA then adds the S partials of its own 16 units in a fixed order
//     (bit-reproducible) -- that is dh_rec of the next (earlier) time step.
// The contraction is split rather than the output columns: an output split would need all of dG (128 x 4H) in every CTA.
// Per step a CTA moves 2 x 128 KB through L2 and runs 1/16 of the tile's MMAs; the chain is ~7 us where the
// one-CTA-per-tile kernel (lstm_rm_tc_bwd.cu) needs ~60 us whatever the batch.
//
// Semantics and tapes as lstm_rm_tc_bwd.cu: RB8 gates / dG (may alias) and c_seq, row-major dh_seq / dh_last.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace hy2dl {
namespace tc {

namespace {
constexpr int kRowsV = 128;
constexpr int kUnitsV = 16;                  // hidden units per CTA
constexpr int kKV = 4 * kUnitsV;             // K values per CTA: k = 16 g + u
constexpr int kThreadsV = 288;               // warps 0-7: rows (two threads per sequence: 8 | 8 units, H/2 | H/2 columns), warp 8: MMA issuer
constexpr int kAImgV = kKV * kRowsV * 4;     // hi image of A, 32 KB; lo follows
constexpr uint32_t kSpinLimitV = 1u << 26;

struct UsBwdArgs {
  const float* wpk;     // [S][hi | lo][16 k4][H][4]: rows g*H + 16 s + u of W_hh
  const float* gates;   // RB8 tape
  const float* c_seq;   // RB8 tape
  const float* dh_seq;  // [T][B][H] row-major or null
  const float* dh_last; // [B][H] or null
  float* dG;            // RB8 tape, may alias gates
  float* px;            // [2][groups][S][H/8][128 rows][8] partials of dh_rec
  uint32_t* flags;      // [groups], zero at launch
  int64_t B, T, dh_t0;
  int H, S, groups, tiles;
  int single;
  long long* stamps;    // HY2DL_US_DBG bit 128: clock64 timeline of CTA 0, reverse steps 100-101
};

__global__ void us_pack_whh_kernel(const float* __restrict__ w_hh, int H, float comp, float* __restrict__ out) {
  const int S = H / kUnitsV;
  const int64_t total = (int64_t)S * kKV * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(e & 3), n = (int)((e >> 2) % H);
    const int64_t r2 = (e >> 2) / H;
    const int k4 = (int)(r2 % (kKV / 4)), s = (int)(r2 / (kKV / 4));
    const int k = 4 * k4 + kk, g = k >> 4, u = k & 15;
    const double vd = (double)w_hh[(int64_t)(g * H + kUnitsV * s + u) * H + n] * (1.0 + (double)comp);
    const uint32_t hi = (__float_as_uint((float)vd) + 0x1000u) & 0xFFFFE000u;
    const uint32_t lo = (__float_as_uint((float)(vd - (double)__uint_as_float(hi))) + 0x1000u) & 0xFFFFE000u;
    float* dst = out + (int64_t)s * 2 * kKV * H + ((int64_t)k4 * H + n) * 4 + kk;
    dst[0] = __uint_as_float(hi);
    dst[(int64_t)kKV * H] = __uint_as_float(lo);
  }
}

__device__ __forceinline__ uint32_t ld_relaxed_gpu_v(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_v(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ f8 ld256_l2v(const float* p) {          // never an L1 hit: the buffer is rewritten by other SMs
  f8 r;
  asm volatile("ld.relaxed.gpu.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}

#ifdef HY2DL_US_STAMPS    // timeline instrumentation, see lstm_us_tc_fwd.cu
#define USB_STAMP(id) do { if (a.stamps && blockIdx.x == 0 && tid == 0 && st >= 100 && st < 102) a.stamps[(st - 100) * 16 + (id)] = clock64(); } while (0)
#else
#define USB_STAMP(id) do { } while (0)
#endif
__global__ void __launch_bounds__(kThreadsV, 1) lstm_us_tc_bwd_kernel(UsBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int H = a.H, S = a.S;
  const int wbytes = kKV * H * 4;                                    // one image (hi or lo) of the W_hh slice
  uint8_t* s_w = smem_raw;
  uint8_t* s_a = s_w + 2 * (size_t)wbytes;                           // [hi | lo][16 k4][128 rows][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + 2 * (size_t)kAImgV);
  uint64_t* bar_a = bars;               // rows -> MMA: A image in place (and the accumulator of the previous step read)
  uint64_t* bar_acc = bars + 1;         // [2] MMA commit -> rows: column half hh of the accumulator is complete
  uint64_t* bar_w = bars + 3;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int group = blockIdx.x / S, s = blockIdx.x % S;
  if (tid == 0) {
    mbar_init(bar_a, 256);
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc + 1, 1);
    mbar_init(bar_w, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  uint32_t* flag = a.flags + group;

  if (warp < 8) {
    // =========================== row threads ===================================================
    const int q = warp & 3, hh = warp >> 2;
    const int rloc = q * 32 + lane;
    const uint32_t lane_tm = tmem + ((uint32_t)(q * 32) << 16);
    const int64_t G = (a.B + 31) >> 5;
    const int HB = H / 8, ub = 2 * s + hh;                           // this thread's unit block (8 units) of the tapes
    const int HH = H / 2;
    int64_t gs = 0;                                                  // step counter over all tiles of the group
    for (int tile = group; tile < a.tiles; tile += a.groups) {
      const int64_t b = (int64_t)tile * kRowsV + rloc;
      const bool live = b < a.B;
      f8 dc = f8_zero();
      for (int64_t st = 0; st < a.T; ++st, ++gs) {
        const int64_t t = a.T - 1 - st;
        // ---- tapes of this step: no dependence on the other CTAs, issued before the wait
        f8 gi = f8_zero(), gf = gi, gg = gi, go = gi, ct = gi, cp = gi, dh = gi;
        if (live) {
          const float* gp = a.gates + rb8_off(t, G, b, 4 * H, ub);
          gi = ld256(gp); gf = ld256(gp + HB * 256); gg = ld256(gp + 2 * HB * 256); go = ld256(gp + 3 * HB * 256);
          ct = ld256(a.c_seq + rb8_off(t, G, b, H, ub));
          if (t > 0) cp = ld256(a.c_seq + rb8_off(t - 1, G, b, H, ub));
          if (a.dh_seq && t >= a.dh_t0) dh = ld256(a.dh_seq + (t * a.B + b) * H + kUnitsV * s + 8 * hh);
          if (a.dh_last && st == 0) { const f8 e = ld256(a.dh_last + b * H + kUnitsV * s + 8 * hh); for (int k = 0; k < 8; ++k) dh.v[k] += e.v[k]; }
        }
        USB_STAMP(0);
        // ---- dh_rec(t): the S partials of the previous step, added in a fixed order
        if (gs > 0) {                               // acquire polls by every thread: the fastest variant measured (lstm_us_tc_fwd.cu)
          const uint32_t target = (uint32_t)(gs * S);
          uint32_t spins = 0;
          while ((int32_t)(ld_relaxed_gpu_v(flag) - target) < 0) {
            if (++spins > kSpinLimitV) __trap();
          }
          (void)ld_acquire_gpu_v(flag);
        }
        USB_STAMP(1);
        if (st > 0) {
          const float* pp = a.px + (((((st - 1) & 1) * (int64_t)a.groups + group) * S * HB + ub) * kRowsV + rloc) * 8;   // partial j: + j * HB * 1024
          f8 acc = f8_zero();
          for (int j0 = 0; j0 < S; j0 += 8) {     // 8 loads in flight (S is 8 or 16), then a fixed-order sum
            f8 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (j0 + j < S) ? ld256_l2v(pp + (int64_t)(j0 + j) * HB * (kRowsV * 8)) : f8_zero();   // S = 4 at H = 64
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
              for (int k = 0; k < 8; ++k) acc.v[k] += v[j].v[k];
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) dh.v[k] += acc.v[k];
        }
        USB_STAMP(2);
        // ---- gate gradients of this thread's 8 units
        f8 di, df, dg, dO;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float tcv = fast_tanh(ct.v[k]);
          const float dct = dc.v[k] + dh.v[k] * go.v[k] * (1.f - tcv * tcv);
          dO.v[k] = dh.v[k] * tcv * go.v[k] * (1.f - go.v[k]);
          di.v[k] = dct * gg.v[k] * gi.v[k] * (1.f - gi.v[k]);
          df.v[k] = dct * cp.v[k] * gf.v[k] * (1.f - gf.v[k]);
          dg.v[k] = dct * gi.v[k] * (1.f - gg.v[k] * gg.v[k]);
          dc.v[k] = dct * gf.v[k];
        }
        if (live) {
          float* dgp = a.dG + rb8_off(t, G, b, 4 * H, ub);
          stg256(dgp, di); stg256(dgp + HB * 256, df); stg256(dgp + 2 * HB * 256, dg); stg256(dgp + 3 * HB * 256, dO);
        }
        if (t == 0) {                               // nothing flows further back; keep the counter uniform
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (tid == 0) atomicAdd(flag, 1u);
          continue;
        }
        USB_STAMP(3);
        // ---- A image [128 x 64]: K index 16 g + 8 hh + k
        {
          float4* img_hi = reinterpret_cast<float4*>(s_a);
          float4* img_lo = img_hi + kAImgV / 16;
          const f8* src[4] = {&di, &df, &dg, &dO};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) split_tf32(src[g]->v[k], hi[k], lo[k]);
            const int k40 = (kUnitsV * g + 8 * hh) >> 2;
            img_hi[k40 * kRowsV + rloc] = make_float4(__uint_as_float(hi[0]), __uint_as_float(hi[1]), __uint_as_float(hi[2]), __uint_as_float(hi[3]));
            img_hi[(k40 + 1) * kRowsV + rloc] = make_float4(__uint_as_float(hi[4]), __uint_as_float(hi[5]), __uint_as_float(hi[6]), __uint_as_float(hi[7]));
            img_lo[k40 * kRowsV + rloc] = make_float4(__uint_as_float(lo[0]), __uint_as_float(lo[1]), __uint_as_float(lo[2]), __uint_as_float(lo[3]));
            img_lo[(k40 + 1) * kRowsV + rloc] = make_float4(__uint_as_float(lo[4]), __uint_as_float(lo[5]), __uint_as_float(lo[6]), __uint_as_float(lo[7]));
          }
        }
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(bar_a);
        USB_STAMP(4);
        // ---- this CTA's partial of dh_rec(t-1): columns [hh H/2, (hh+1) H/2) of the accumulator -> exchange buffer
        mbar_wait(bar_acc + hh, (uint32_t)((gs - (tile - group) / a.groups) & 1));      // one MMA round per step but the last of a tile
        tc_fence_after();
        USB_STAMP(5);
        // exchange layout [slot][group][s][H/8 column blocks][128 rows][8]: a warp's 256-bit accesses cover 1 KB contiguous (row-major
        // partials cost 4.1 us per step here and 1.9 us in the sum above: 32 lines per instruction)
        float* po = a.px + (((((st & 1) * (int64_t)a.groups + group) * S + s) * HB + hh * (HB / 2)) * kRowsV + rloc) * 8;
        for (int c0 = 0; c0 < HH; c0 += 64) {        // two 32-column pulls in flight
          uint32_t z[32], y[32];
          const bool second = c0 + 32 < HH;          // H = 64: 32 columns per thread
          tmem_ld32(lane_tm + hh * HH + c0, z);
          if (second) tmem_ld32(lane_tm + hh * HH + c0 + 32, y);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f8 v;
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] = __uint_as_float(z[8 * i + k]);
            st256(po + (c0 / 8 + i) * (kRowsV * 8), v);
          }
          if (second)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f8 v;
#pragma unroll
            for (int k = 0; k < 8; ++k) v.v[k] = __uint_as_float(y[8 * i + k]);
            st256(po + (c0 / 8 + 4 + i) * (kRowsV * 8), v);
          }
        }
        USB_STAMP(6);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) { __threadfence(); atomicAdd(flag, 1u); }
        USB_STAMP(7);
      }
    }
  } else {
    // =========================== MMA issuer =====================================================
    if (elect_one()) {
      {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wpk) + (size_t)s * 2 * wbytes;
        mbar_arrive_expect_tx(bar_w, 2u * (uint32_t)wbytes);
        for (int off = 0; off < 2 * wbytes; off += 16384) bulk_g2s(s_w + off, src + off, (uint32_t)min(16384, 2 * wbytes - off), bar_w);
        mbar_wait(bar_w, 0);
      }
      // the accumulator is computed in two column halves (N = H/2 each) with a commit in between: the row threads of
      // half 0 pull and publish their columns while the tensor core still multiplies half 1 (the pull -- 128 KB through the
      // 64 B/clk TMEM read port -- and the MMAs took 4.2 k + 3.3 k cycles one after the other, measured with HY2DL_US_STAMPS)
      const uint32_t idesc = idesc_tf32(kRowsV, H / 2, 0, 0);
      const uint64_t whi0 = smem_desc(smem_u32(s_w), (uint32_t)H * 16, 128), wlo0 = smem_desc(smem_u32(s_w + wbytes), (uint32_t)H * 16, 128);
      const uint64_t ahi0 = smem_desc(smem_u32(s_a), kRowsV * 16, 128), alo0 = smem_desc(smem_u32(s_a + kAImgV), kRowsV * 16, 128);
      const uint64_t kBStep = (uint64_t)(2 * H * 16 / 16);
      constexpr uint64_t kAStep = 2 * kRowsV * 16 / 16;
      int64_t n = 0;                                 // MMA rounds so far
      for (int tile = group; tile < a.tiles; tile += a.groups) {
        for (int64_t st = 0; st + 1 < a.T; ++st, ++n) {
          mbar_wait(bar_a, (uint32_t)(n & 1));
          tc_fence_after();
          for (int half = 0; half < 2; ++half) {
            const uint32_t d = tmem + (uint32_t)(half * (H / 2));
            const uint64_t bo = (uint64_t)(half * (H / 2));            // column n of B: n * 16 bytes
            if (!a.single) {
#pragma unroll
              for (int j = 0; j < kKV / 8; ++j) {        // correction terms first, while the accumulator is small
                mma_tf32_ss(d, alo0 + j * kAStep, whi0 + bo + j * kBStep, idesc, j != 0);
                mma_tf32_ss(d, ahi0 + j * kAStep, wlo0 + bo + j * kBStep, idesc, 1);
              }
            }
#pragma unroll
            for (int j = 0; j < kKV / 8; ++j) mma_tf32_ss(d, ahi0 + j * kAStep, whi0 + bo + j * kBStep, idesc, (a.single && j == 0) ? 0 : 1);
            mma_commit(bar_acc + half);
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}
}  // namespace

static int usb_groups(int64_t B, int64_t H) {
  const int S = (int)(H / kUnitsV);
  return (int)std::min<int64_t>(cdiv(B, (int64_t)kRowsV), sm_count() / S);
}

bool lstm_us_tc_bwd_supported(int64_t H) {
  return (H == 64 || H == 128 || H == 256) && sm_count() >= H / kUnitsV;
}

// as lstm_us_tc_fwd_preferred: 2.9 ms per round against 25-29 ms (H = 256), 1.9 ms against 8-10 ms (H = 128); HY2DL_RM_US(_BWD)=0 / 1 forces
bool lstm_us_tc_bwd_preferred(int64_t B, int64_t H) {
  if (!lstm_us_tc_bwd_supported(H)) return false;
  static const int force = [] { const char* e = getenv("HY2DL_RM_US_BWD"); if (!e) e = getenv("HY2DL_RM_US"); return e ? atoi(e) : -1; }();
  if (force == 0) return false;
  if (force == 1) return true;
  const int64_t tiles = cdiv(B, (int64_t)kRowsV);
  const int64_t rounds = cdiv(tiles, (int64_t)usb_groups(B, H));
  return rounds <= (H == 256 ? 8 : H == 128 ? 4 : 1);
}

int64_t lstm_us_tc_bwd_workspace(int64_t B, int64_t H) {
  return sizeof(float) * (8 * H * H + 2 * (int64_t)usb_groups(B, H) * kRowsV * H * (H / kUnitsV)) + 1024 + 256;
}

int lstm_us_tc_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates, const float* c_seq, const float* dh_seq,
                   int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, int single, cudaStream_t s) {
  const int64_t need = lstm_us_tc_bwd_workspace(B, H);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_bwd: workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  UsBwdArgs a{};
  a.S = (int)(H / kUnitsV);
  a.tiles = (int)cdiv(B, (int64_t)kRowsV);
  a.groups = usb_groups(B, H);
  float* wpk = reinterpret_cast<float*>(ws);
  float* px = wpk + 8 * H * H;
  uint32_t* flags = reinterpret_cast<uint32_t*>(px + 2 * (int64_t)a.groups * kRowsV * H * a.S);
  // chains of 8 hi*hi MMAs over gate gradients of random sign: no measurable mean bias (lstm_rm_tc_bwd.cu); HY2DL_RM_BCOMP overrides
  static const float comp = [] { const char* e = getenv("HY2DL_RM_BCOMP"); return e ? (float)atof(e) : 0.f; }();
  us_pack_whh_kernel<<<256, 256, 0, s>>>(w_hh, (int)H, comp, wpk);
  cudaMemsetAsync(flags, 0, 1024, s);
  a.wpk = wpk; a.gates = gates; a.c_seq = c_seq; a.dh_seq = dh_seq; a.dh_last = dh_last; a.dG = dG; a.px = px; a.flags = flags;
  a.B = B; a.T = T; a.dh_t0 = dh_t0; a.H = (int)H; a.single = single;
  const size_t smem = (size_t)kKV * H * 8 + 2 * kAImgV + 256 + 1024;
  cudaFuncSetAttribute(lstm_us_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.groups * a.S)); cfg.blockDim = dim3(kThreadsV); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;       // all CTAs co-resident, or the launch fails: they wait for each other
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
#ifdef HY2DL_US_STAMPS
  static const int dbg = [] { const char* e = getenv("HY2DL_US_DBG"); return e ? atoi(e) : 0; }();
  if (dbg & 128) { cudaMalloc(&a.stamps, 32 * sizeof(long long)); cudaMemset(a.stamps, 0, 32 * sizeof(long long)); }
#endif
  cudaLaunchKernelEx(&cfg, lstm_us_tc_bwd_kernel, a);
#ifdef HY2DL_US_STAMPS
  if (a.stamps) {                                   // debugging only: timeline of CTA 0, thread 0, reverse steps 100-101
    long long h[32];
    cudaDeviceSynchronize();
    cudaMemcpy(h, a.stamps, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(a.stamps);
    for (int i = 0; i < 32; ++i) if (h[i]) fprintf(stderr, "USB_STAMP step %d id %2d  %8lld\n", 100 + i / 16, i % 16, h[i] - h[0]);
  }
#endif
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

}  // namespace tc
}  // namespace hy2dl
