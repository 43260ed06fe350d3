// Weight gradients of the LSTM on tcgen05 (HY2DL_PREC_TF32X3), H = 64: one time-parallel GEMM
//
//     dW[256 x 112] = sum over (t, sequence)  dG[t, seq, 256]^T  *  [h_{t-1} | x_t | 1 | 0..][t, seq, 112]
//
// i.e. dW_hh, dW_ih and db in one accumulation (the "1" column produces the bias gradient).  The
// contraction index is the sequence, which is the row index of every RI32 stream: a (t, group,
// 32-feature block) slab, [32 rows][32 floats] = 4 KB, is bit-for-bit the MN-major
// SWIZZLE_128B_BASE32B UMMA operand tile for K = 32 (the only MN-major layout kind::tf32 takes), so
// operands go HBM -> shared memory with plain bulk async copies (TMA engine, no tensor map) and are
// consumed by tcgen05.mma (A and B from shared memory) without any reshuffle.
//
// fp32-class accuracy: 3-term TF32 split, hi = x rounded to tf32, lo = exact remainder.  The converter
// warps turn each landed slab into (hi in place, lo in a second buffer); the MMA thread then issues
// lo*hi + hi*lo + hi*hi.
// The tensor core's fp32 accumulation TRUNCATES (measured, tools/dbg_rz.py: relative bias -2.8e-8 per
// accumulating MMA, -8.5e-5 after 3072), so a TMEM accumulator is only trusted for kChain = 8 slabs
// (96 MMAs): after every chain the converter warps add the chain accumulator D (2 halves x [128 lanes x
// 112 columns]) into a running sum S kept in a second TMEM region, in registers with round-to-nearest
// (tcgen05.ld D, S -> FADD -> tcgen05.st S), and the next chain starts by overwriting D.  The MMA pipe
// idles for the ~0.5 us of a drain, once per 8 slabs.  Every CTA writes its sum S once and a
// small kernel adds the partials in a fixed order (deterministic) and scatters them into PyTorch's
// [4H, I] / [4H, H] / [4H] layouts.
//
// Warp roles (320 threads): warp 0 = bulk-copy producer, warp 1 = MMA issuer, warps 2-9 = split
// converters + accumulator drain + epilogue.  3 raw stages (48 KB each) + A_lo + 2 x B_lo.  The pipeline unit is
// half a slab (128 gate columns): while the tensor core works on one half the converters split the next, so
// the MMA stream is not interrupted by the split.  The kernel is shared-memory-bandwidth bound (SS-mode MMAs
// read 7.5 KB per 56-cycle instruction), which is why the landed fp32 words double as the hi operand.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace hy2dl {
namespace tc {

constexpr int kWgThreads = 320;
constexpr int kWgStages = 3;
constexpr int kWgConv = 256;                // converter / drain threads (warps 2..9)
constexpr int kChain = 8;                   // slabs accumulated in TMEM before a drain
constexpr int kBlk = 4096;                  // bytes of one [32 rows][32 floats] block
constexpr int kABlocks = 8;                 // 256 gate columns
constexpr int kABytes = kABlocks * kBlk;    // 32 KB: one (t, group) slab of dG
constexpr int kBBlocks = 4;                 // h (2 blocks) | x (1) | constant (1)
constexpr int kBBytes = kBBlocks * kBlk;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kND = 112;                    // D columns: [0,64) dW_hh, [64,96) dW_ih, 96 db, rest 0
constexpr int kH64 = 64;

struct TcWgradArgs {
  const float* dG;   // RI32 F = 256 (n = 4u + g)
  const float* h;    // RI32 F = 64
  const float* x;    // RI32 F = 32
  float* partial;    // [grid][256][kND]
  int64_t T, G;
  int raw_hi;        // leave the landed fp32 words in place as the "hi" operand: kind::tf32 ignores the low 13 bits
                     // (verified, tools/dbg_rz.py with HY2DL_WG_RAWHI=0/1), which saves a third of the split's smem traffic
};

__global__ void __launch_bounds__(kWgThreads, 1) lstm_tc_wgrad_kernel(TcWgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t* s_lo = smem_raw + kWgStages * kStageBytes;          // [A_lo (two halves) | B_lo[0] | B_lo[1]]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lo + kABytes + 2 * kBBytes);
  uint64_t* bar_full = bars;                       // [stages] producer (tx bytes) -> converters
  uint64_t* bar_raw_free = bars + kWgStages;       // [stages] MMA commit -> producer
  uint64_t* bar_conv = bars + 2 * kWgStages;       // [2] converters -> MMA, per half of the gate columns
  uint64_t* bar_lo_free = bar_conv + 2;            // [2] MMA commit -> converters, per half
  uint64_t* bar_done = bar_conv + 4;               // all MMAs retired -> epilogue
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_conv + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_groups = a.T * a.G;
  const int64_t n_my = blockIdx.x < n_groups ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // ---- setup: barriers, TMEM, constant block (hi: column 0 of every row = 1; lo: 0)
  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_raw_free + s, 1); }
    for (int hf = 0; hf < 2; ++hf) { mbar_init(bar_conv + hf, kWgConv); mbar_init(bar_lo_free + hf, 1); }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  for (int e = tid; e < kBlk / 16; e += kWgThreads) {           // float4 e = row * 8 + chunk
    const float4 one = make_float4((e & 7) == 2 * ((e >> 3) & 3) ? 1.f : 0.f, 0.f, 0.f, 0.f);   // unit 0 of row r sits at unit r & 3
    for (int s = 0; s < kWgStages; ++s)
      reinterpret_cast<float4*>(smem_raw + s * kStageBytes + kABytes + 3 * kBlk)[e] = one;
    reinterpret_cast<float4*>(s_lo + kABytes + 3 * kBlk)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(s_lo + kABytes + kBBytes + 3 * kBlk)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    // =========================== producer ======================================================
    if (lane == 0) {
      for (int64_t i = 0; i < n_my; ++i) {
        const int s = (int)(i % kWgStages);
        const int64_t id = blockIdx.x + i * gridDim.x, t = id / a.G, g = id - t * a.G;
        if (i >= kWgStages) mbar_wait(bar_raw_free + s, (uint32_t)((i / kWgStages - 1) & 1));
        uint8_t* st = smem_raw + s * kStageBytes;
        const uint32_t bytes = kABytes + kBlk + (t > 0 ? 2 * kBlk : 0);
        mbar_arrive_expect_tx(bar_full + s, bytes);
        bulk_g2s(st, a.dG + (t * a.G + g) * (int64_t)(kABytes / 4), kABytes, bar_full + s);
        bulk_g2s(st + kABytes + 2 * kBlk, a.x + (t * a.G + g) * (int64_t)(kBlk / 4), kBlk, bar_full + s);
        if (t > 0) bulk_g2s(st + kABytes, a.h + ((t - 1) * a.G + g) * (int64_t)(2 * kBlk / 4), 2 * kBlk, bar_full + s);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer =====================================================
    if (lane == 0) {
      const uint32_t lo_a = smem_u32(s_lo);
      const uint32_t idesc = idesc_tf32(128, kND, 1, 1);
      for (int64_t i = 0; i < n_my; ++i) {
        const int s = (int)(i % kWgStages);
        const uint32_t hi_a = smem_u32(smem_raw + s * kStageBytes), hi_b = hi_a + kABytes;
        const uint32_t lo_b = lo_a + kABytes + (uint32_t)(i & 1) * kBBytes;
        const bool first = (i % kChain) == 0;                               // first slab of a chain overwrites D
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {           // gate columns [128 hf, 128 hf + 128): the pipeline unit
          mbar_wait(bar_conv + hf, (uint32_t)(i & 1));
          tc_fence_after();
          const uint32_t d = tmem + 128 * hf;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {         // 8 sequences (K = 8 = two 4-row groups) per MMA
            const uint64_t b_hi = smem_desc_mn32(hi_b + kb * 1024, kBlk, 512), b_lo = smem_desc_mn32(lo_b + kb * 1024, kBlk, 512);
            const uint64_t a_hi = smem_desc_mn32(hi_a + hf * 4 * kBlk + kb * 1024, kBlk, 512);
            const uint64_t a_lo = smem_desc_mn32(lo_a + hf * 4 * kBlk + kb * 1024, kBlk, 512);
            mma_tf32_ss(d, a_lo, b_hi, idesc, !(first && kb == 0));
            mma_tf32_ss(d, a_hi, b_lo, idesc, 1);
            mma_tf32_ss(d, a_hi, b_hi, idesc, 1);
          }
          mma_commit(bar_lo_free + hf);
        }
        mma_commit(bar_raw_free + s);
      }
      mma_commit(bar_done);
    }
    __syncwarp();
  } else {
    // =========================== converters, then epilogue ======================================
    const int ct = tid - 64;                                  // 0 .. 255
    const int hf = ct >> 7;                                   // drain: warps 2-5 own half 0, warps 6-9 half 1
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128 * hf;
    // S = 0
    {
      uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < kND; c += 8) tmem_st8(lane_base + 256 + c, z);
      tmem_st_wait();
    }
    auto drain = [&]() {                                      // S += D (every MMA into D has retired)
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kND; c += 16) {
        uint32_t d[16], sacc[16];
        tmem_ld16(lane_base + c, d);
        tmem_ld16(lane_base + 256 + c, sacc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) sacc[j] = __float_as_uint(__uint_as_float(sacc[j]) + __uint_as_float(d[j]));
        tmem_st8(lane_base + 256 + c, *reinterpret_cast<uint32_t(*)[8]>(&sacc[0]));
        tmem_st8(lane_base + 256 + c + 8, *reinterpret_cast<uint32_t(*)[8]>(&sacc[8]));
      }
      tmem_st_wait();
      tc_fence_before();
    };
    float4* lo_a4 = reinterpret_cast<float4*>(s_lo);
    for (int64_t i = 0; i < n_my; ++i) {
      const int s = (int)(i % kWgStages);
      const int64_t id = blockIdx.x + i * gridDim.x, t = id / a.G;
      float4* raw = reinterpret_cast<float4*>(smem_raw + s * kStageBytes);
      float4* lo_b4 = reinterpret_cast<float4*>(s_lo + kABytes + (i & 1) * kBBytes);
      constexpr int kA4 = kABytes / 16, kBlk4 = kBlk / 16;
      // hi/lo split of float4 `e` of the stage: A part -> lo_a4, B part -> lo_b4 (B_lo is double buffered by slab)
      auto split4 = [&](int e) {
        const float4 v = raw[e];
        float4* lo = e < kA4 ? lo_a4 + e : lo_b4 + (e - kA4);
        float4 hv, lv;
        if (a.raw_hi) {   // hi = truncation (what the tensor core reads from the raw word), lo = exact remainder
          hv.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); hv.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          hv.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); hv.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          *lo = make_float4(v.x - hv.x, v.y - hv.y, v.z - hv.z, v.w - hv.w);
          return;
        }
        hv.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u); lv.x = v.x - hv.x;
        hv.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u); lv.y = v.y - hv.y;
        hv.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u); lv.z = v.z - hv.z;
        hv.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u); lv.w = v.w - hv.w;
        raw[e] = hv;
        *lo = lv;
      };
      mbar_wait(bar_full + s, (uint32_t)((i / kWgStages) & 1));
      // ---- half 0: the B operand (shared by both halves) + gate columns [0, 128)
      if (i > 0) mbar_wait(bar_lo_free + 0, (uint32_t)((i - 1) & 1));    // MMAs of (slab i-1, half 0) and everything before retired
      if (i > 0 && (i % kChain) == 0) {                                   // chain complete: S += D before slab i overwrites D
        mbar_wait(bar_lo_free + 1, (uint32_t)((i - 1) & 1));
        drain();
      }
      for (int e = ct; e < kBlk4; e += kWgConv) split4(kA4 + 2 * kBlk4 + e);    // x
      if (t > 0) {
        for (int e = ct; e < 2 * kBlk4; e += kWgConv) split4(kA4 + e);          // h_{t-1}
      } else {                                                                  // h_{-1} = 0
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = ct; e < 2 * kBlk4; e += kWgConv) { raw[kA4 + e] = z; lo_b4[e] = z; }
      }
#pragma unroll 4
      for (int e = ct; e < kA4 / 2; e += kWgConv) split4(e);
      fence_async_smem();
      mbar_arrive(bar_conv + 0);
      // ---- half 1: gate columns [128, 256)
      if (i > 0) mbar_wait(bar_lo_free + 1, (uint32_t)((i - 1) & 1));
#pragma unroll 4
      for (int e = ct; e < kA4 / 2; e += kWgConv) split4(kA4 / 2 + e);
      fence_async_smem();
      mbar_arrive(bar_conv + 1);
    }
    // last chain, then this CTA's sums
    mbar_wait(bar_done, 0);
    if (n_my > 0) drain();
    tc_fence_after();
    const int row = hf * 128 + (warp & 3) * 32 + lane;
    float* dst = a.partial + ((int64_t)blockIdx.x * 256 + row) * kND;
#pragma unroll 1
    for (int c = 0; c < kND; c += 16) {
      uint32_t v[16];
      tmem_ld16(lane_base + 256 + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(dst + c + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// partial [n_part][256][kND] -> dw_hh [256][64], dw_ih [256][I], db [256] (row = g*64 + u from n = 4u + g)
__global__ void lstm_tc_wgrad_finish_kernel(const float* __restrict__ partial, int n_part, int I,
                                            float* __restrict__ dw_ih, float* __restrict__ dw_hh, float* __restrict__ db) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 256 * kND) return;
  const int n = e / kND, col = e - n * kND;
  if (col > 96) return;
  float acc = 0.f;
  for (int p = 0; p < n_part; ++p) acc += __ldg(partial + (int64_t)p * 256 * kND + e);
  const int u = n >> 2, g = n & 3, row = g * kH64 + u;
  if (col < kH64) dw_hh[row * kH64 + col] = acc;
  else if (col < 96) { if (col - kH64 < I) dw_ih[row * I + (col - kH64)] = acc; }
  else db[row] = acc;
}

static int wg_grid(int64_t T, int64_t G) { return (int)std::min<int64_t>(sm_count(), T * G); }

int64_t lstm_tc_wgrad_workspace(int64_t B, int64_t T, int64_t I) {
  (void)B; (void)T; (void)I;
  return sizeof(float) * 256 * kND * std::max<int64_t>(sm_count(), 1);
}

int lstm_tc_wgrad(const float* dG, const float* h_seq, const float* x_ri, int64_t B, int64_t T, int64_t I, int64_t IP,
                  float* dw_ih, float* dw_hh, float* db, void* ws, int64_t ws_bytes, cudaStream_t s) {
  const int64_t need = lstm_tc_wgrad_workspace(B, T, I);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_wgrad(tf32x3): workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  HY_REQUIRE(IP == 32, HY2DL_ERR_SHAPE, "hy2dl_lstm_wgrad(tf32x3): the RI32 input stream is 32 columns wide (IP=%lld)", (long long)IP);
  TcWgradArgs a{};
  a.dG = dG; a.h = h_seq; a.x = x_ri; a.partial = reinterpret_cast<float*>(ws);
  a.T = T; a.G = cdiv(B, 32);
  static const int raw_hi = [] { const char* e = getenv("HY2DL_WG_RAWHI"); return e ? atoi(e) : 1; }();   // 0: write the rounded hi back in place
  a.raw_hi = raw_hi;
  const size_t smem = (size_t)kWgStages * kStageBytes + kABytes + 2 * kBBytes + 128 + 1024;
  const int grid = wg_grid(T, a.G);
  cudaFuncSetAttribute(lstm_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lstm_tc_wgrad_kernel<<<grid, kWgThreads, smem, s>>>(a);
  lstm_tc_wgrad_finish_kernel<<<(256 * kND + 255) / 256, 256, 0, s>>>(a.partial, grid, (int)I, dw_ih, dw_hh, db);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

}  // namespace tc
}  // namespace hy2dl
