// Weight gradients of the LSTM (and of the fused head) on tcgen05 (HY2DL_PREC_TF32X3), H = 64: one
// time-parallel GEMM
//
//     dW[256 x 96] = sum over (t, sequence)  dG[t, seq, 256]^T  *  [h_{t-1} | x_t (column 31 := 1)][t, seq, 96]
//
// i.e. dW_hh, dW_ih and db in one accumulation (x's unused 32nd column is set to 1 and produces the bias
// gradient), plus, when the head is fused,
//
//     [dW_head | . | db_head]^T [96 x NHz] = sum over (t, sequence)  [h_{t-1} | x_t (column 31 = 1)]^T  *  dz[t-1, seq, NHz]
//
// on the same h slabs (the head acts on h_t, i.e. the slab that step t+1 loads as h_{t-1}); the ones column
// yields the head bias gradient, the other x rows are computed and ignored.  The contraction
// index is the sequence, which is the row index of every RI32 stream: a (t, group, 32-feature block) slab,
// [32 rows][32 floats] = 4 KB, is bit-for-bit the MN-major SWIZZLE_128B_BASE32B UMMA operand tile for K = 32
// (the only MN-major layout kind::tf32 takes), so operands go HBM -> shared memory with plain bulk async
// copies (TMA engine, no tensor map) and are consumed by tcgen05.mma (A and B from shared memory) untouched.
//
// fp32-class accuracy: 3-term TF32 split a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi.  kind::tf32 ignores the low
// 13 bits of its operands (verified, tools/dbg_rz.py with HY2DL_WG_RAWHI=0/1), so the landed fp32 words
// themselves are the hi operand and the converter warps only produce lo = x - trunc(x) in a second buffer.
// The tensor core's fp32 accumulation TRUNCATES (measured: relative bias -2.8e-8 per accumulating MMA, -8.5e-5
// after 3072), so a TMEM accumulator is only trusted for kChain = 8 slabs (96 MMAs): after every chain the
// converter warps add the chain accumulator D into a running sum S kept in a second TMEM region, in registers
// with round-to-nearest (tcgen05.ld D, S -> FADD -> tcgen05.st S), and the next chain starts by overwriting D.
// Every CTA writes its sum S once and a small kernel adds the partials in a fixed order (deterministic).
//
// Warp roles (320 threads): warp 0 = bulk-copy producer, warp 1 = MMA issuer, warps 2-9 = split converters +
// accumulator drain + epilogue.  3 raw stages (48 KB each) + A_lo + 2 x B_lo.  The pipeline unit is half a slab
// (128 gate columns): while the tensor core works on one half the converters split the next.
//
// Measured (B = 18944, T = 365, no head): 2.23 ms = 4.4 TB/s.  The MMA warp issues under `elect_one()`: guarded by
// `lane == 0` the same loop paid an ELECT + 4 x R2UR.BROADCAST waterfall per tcgen05.mma (divergent code cannot use
// the uniform datapath) and the kernel ran 2.69 ms, issue-bound at ~130 cycles per 48-cycle MMA.  Timing
// experiments on that version: hi*hi MMAs only 2.01 ms, no dG split 2.57 ms, both 1.68 ms (88 % of HBM peak = the
// floor of this pipeline).  A variant with dG^T transposed through registers into TMEM (TS-mode MMAs, no dG lo image
// in shared memory, running sum in global memory) measured the same 2.20 ms and was dropped; K-major instead of
// MN-major descriptors made no difference either (MN-major kind::tf32 operands are not slower).
//
// Stage layout: A part = [dz block | dG blocks 0..7] (36 KB), B part = [h block 0 | h block 1 | x block] (12 KB).
// The head MMA takes the stage's B part as its A operand (M = 128: rows 96.. are whatever follows in shared memory,
// computed and ignored) and the first NHz = 16 or 32 columns of the dz block as its B operand.
// TMEM: chain D = gate columns 0..127 -> [0,96), 128..255 -> [96,192), head^T -> [192,192+NHz); running sum S = D + 256.
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
constexpr int kBlk4 = kBlk / 16;            // float4 per block
constexpr int kABytes = 9 * kBlk;           // dz block + 8 dG blocks
constexpr int kBBytes = 3 * kBlk;           // h (2 blocks) | x (1)
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kND = 96;                     // main D columns: [0,64) dW_hh, [64,95) dW_ih, 95 db
constexpr int kHeadRows = 128;              // head^T accumulator rows kept per CTA (64 h units, 32 x columns incl. ones, 32 junk)
constexpr int kH64 = 64;
constexpr uint32_t kColHeadD = 192;

struct TcWgradArgs {
  const float* dG;    // RI32 F = 256 (n = 4u + g)
  const float* h;     // RI32 F = 64
  const float* x;     // RI32 F = 32
  const float* dz;    // RI32 F = 32, valid for t >= dz_t0 (null: no head)
  float* partial;     // [grid][256][kND] then [grid][kHeadRows][32]
  int64_t T, G, dz_t0;
  int NHz;            // head columns rounded up to 16 (MMA N of the head part)
  int raw_hi;         // 1: landed words are the hi operand; 0: write the rounded hi back in place
  int single;         // HY2DL_PREC_TF32: hi*hi MMAs only
};

__global__ void __launch_bounds__(kWgThreads, 1) lstm_tc_wgrad_kernel(TcWgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t* s_lo = smem_raw + kWgStages * kStageBytes;          // [A_lo (dz | dG halves) | B_lo[0] | B_lo[1]]
  // (+ one block of slack: the head MMA's lo A operand reads 4 blocks starting at a 3-block B_lo buffer)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lo + kABytes + 2 * kBBytes + kBlk);
  uint64_t* bar_full = bars;                       // [stages] producer (tx bytes) -> converters
  uint64_t* bar_raw_free = bars + kWgStages;       // [stages] MMA commit -> producer
  uint64_t* bar_conv = bars + 2 * kWgStages;       // [2] converters -> MMA, per half of the gate columns
  uint64_t* bar_lo_free = bar_conv + 2;            // [2] MMA commit -> converters, per half
  uint64_t* bar_done = bar_conv + 4;               // all MMAs retired -> epilogue
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_conv + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool head = a.dz != nullptr;
  // time slots: t = 0 .. T-1 pair dG(t), x(t) with h(t-1); with a head, slot t also pairs dz(t-1) with h(t-1),
  // which adds the slot t = T (dz(T-1), h(T-1) only)
  const int64_t n_groups = (a.T + (head ? 1 : 0)) * a.G;
  const int64_t n_my = blockIdx.x < n_groups ? (n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto slot_of = [&](int64_t i, int64_t& t, int64_t& g) { const int64_t id = blockIdx.x + i * gridDim.x; t = id / a.G; g = id - t * a.G; };
  auto has_dz = [&](int64_t t) { return head && t - 1 >= a.dz_t0; };

  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_raw_free + s, 1); }
    for (int hf = 0; hf < 2; ++hf) { mbar_init(bar_conv + hf, kWgConv); mbar_init(bar_lo_free + hf, 1); }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    // =========================== producer ======================================================
    if (elect_one()) {
      for (int64_t i = 0; i < n_my; ++i) {
        const int s = (int)(i % kWgStages);
        int64_t t, g;
        slot_of(i, t, g);
        if (i >= kWgStages) mbar_wait(bar_raw_free + s, (uint32_t)((i / kWgStages - 1) & 1));
        uint8_t* st = smem_raw + s * kStageBytes;
        const bool main = t < a.T, hprev = t > 0, dzb = has_dz(t);
        const uint32_t bytes = (main ? 9 * kBlk : 0) + (hprev ? 2 * kBlk : 0) + (dzb ? kBlk : 0);
        mbar_arrive_expect_tx(bar_full + s, bytes);
        if (dzb) bulk_g2s(st, a.dz + ((t - 1) * a.G + g) * (int64_t)(kBlk / 4), kBlk, bar_full + s);
        if (main) {
          bulk_g2s(st + kBlk, a.dG + (t * a.G + g) * (int64_t)(8 * kBlk / 4), 8 * kBlk, bar_full + s);
          bulk_g2s(st + kABytes + 2 * kBlk, a.x + (t * a.G + g) * (int64_t)(kBlk / 4), kBlk, bar_full + s);
        }
        if (hprev) bulk_g2s(st + kABytes, a.h + ((t - 1) * a.G + g) * (int64_t)(2 * kBlk / 4), 2 * kBlk, bar_full + s);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer =====================================================
    if (elect_one()) {
      const uint32_t lo_a = smem_u32(s_lo);
      const uint32_t idesc = idesc_tf32(128, kND, 1, 1), idesc_h = idesc_tf32(128, a.NHz, 1, 1);
      bool head_open = false;                       // the head accumulator already holds a term of this chain
      for (int64_t i = 0; i < n_my; ++i) {
        const int s = (int)(i % kWgStages);
        int64_t t, g;
        slot_of(i, t, g);
        const uint32_t hi_a = smem_u32(smem_raw + s * kStageBytes), hi_b = hi_a + kABytes;
        const uint32_t lo_b = lo_a + kABytes + (uint32_t)(i & 1) * kBBytes;
        const bool first = (i % kChain) == 0;                               // first slab of a chain overwrites D
        if (first) head_open = false;
        const bool main = t < a.T;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {           // gate columns [128 hf, 128 hf + 128): the pipeline unit
          mbar_wait(bar_conv + hf, (uint32_t)(i & 1));
          tc_fence_after();
          const uint32_t d = tmem + kND * hf;
          if (main) {
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {       // 8 sequences (K = 8 = two 4-row groups) per MMA
              const uint64_t b_hi = smem_desc_mn32(hi_b + kb * 1024, kBlk, 512), b_lo = smem_desc_mn32(lo_b + kb * 1024, kBlk, 512);
              const uint64_t a_hi = smem_desc_mn32(hi_a + kBlk + hf * 4 * kBlk + kb * 1024, kBlk, 512);
              const uint64_t a_lo = smem_desc_mn32(lo_a + kBlk + hf * 4 * kBlk + kb * 1024, kBlk, 512);
              if (!a.single) {
                mma_tf32_ss(d, a_lo, b_hi, idesc, !(first && kb == 0));
                mma_tf32_ss(d, a_hi, b_lo, idesc, 1);
              }
              mma_tf32_ss(d, a_hi, b_hi, idesc, !a.single || !(first && kb == 0));
            }
          }
          if (hf == 0 && has_dz(t)) {              // head^T: A = [h0 | h1 | x | junk] (the B part), B = dz block
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
              const uint64_t a_hi = smem_desc_mn32(hi_b + kb * 1024, kBlk, 512), a_lo = smem_desc_mn32(lo_b + kb * 1024, kBlk, 512);
              const uint64_t b_hi = smem_desc_mn32(hi_a + kb * 1024, kBlk, 512), b_lo = smem_desc_mn32(lo_a + kb * 1024, kBlk, 512);
              if (!a.single) {
                mma_tf32_ss(tmem + kColHeadD, a_lo, b_hi, idesc_h, head_open || kb > 0);
                mma_tf32_ss(tmem + kColHeadD, a_hi, b_lo, idesc_h, 1);
              }
              mma_tf32_ss(tmem + kColHeadD, a_hi, b_hi, idesc_h, !a.single || head_open || kb > 0);
            }
            head_open = true;
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
    const uint32_t lane_tm = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t dcol = kND * hf;                           // this thread's main columns [dcol, dcol + 96)
    const uint32_t hcol = kColHeadD;                          // head^T columns (drained by the half-0 threads)
    // S = 0
    {
      uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < kND; c += 8) tmem_st8(lane_tm + 256 + dcol + c, z);
      if (hf == 0) for (int c = 0; c < 32; c += 8) tmem_st8(lane_tm + 256 + hcol + c, z);
      tmem_st_wait();
    }
    auto add16 = [&](uint32_t col) {                          // S[col .. col+16) += D[col .. col+16)
      uint32_t d[16], sacc[16];
      tmem_ld16(lane_tm + col, d);
      tmem_ld16(lane_tm + 256 + col, sacc);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) sacc[j] = __float_as_uint(__uint_as_float(sacc[j]) + __uint_as_float(d[j]));
      tmem_st8(lane_tm + 256 + col, *reinterpret_cast<uint32_t(*)[8]>(&sacc[0]));
      tmem_st8(lane_tm + 256 + col + 8, *reinterpret_cast<uint32_t(*)[8]>(&sacc[8]));
    };
    auto drain = [&](bool main_any, bool head_any) {          // S += D (every MMA into D has retired)
      tc_fence_after();
      if (main_any) {
#pragma unroll 1
        for (int c = 0; c < kND; c += 16) add16(dcol + c);
      }
      if (head_any && hf == 0) { add16(hcol); if (a.NHz > 16) add16(hcol + 16); }
      tmem_st_wait();
      tc_fence_before();
    };
    // what the chain that ends with slab `last` accumulated: main terms unless all of its slabs are t = T
    // slots; head terms iff its last slab has one (has_dz is monotone in t, t is monotone in the slab index)
    auto chain_flags = [&](int64_t first_i, int64_t last_i, bool& main_any, bool& head_any) {
      int64_t t, g;
      slot_of(first_i, t, g); main_any = t < a.T;
      slot_of(last_i, t, g); head_any = has_dz(t);
    };
    float4* lo_a4 = reinterpret_cast<float4*>(s_lo);
    for (int64_t i = 0; i < n_my; ++i) {
      const int s = (int)(i % kWgStages);
      int64_t t, g;
      slot_of(i, t, g);
      const bool main = t < a.T, dzb = has_dz(t);
      float4* raw = reinterpret_cast<float4*>(smem_raw + s * kStageBytes);
      float4* lo_b4 = reinterpret_cast<float4*>(s_lo + kABytes + (i & 1) * kBBytes);
      constexpr int kA4 = kABytes / 16;
      // lo = x - trunc(x) of float4 `e` of the stage: A part -> lo_a4, B part -> lo_b4 (double buffered by slab)
      // (the loads of a half are issued first, the arithmetic and stores follow: the shared-memory pipe is shared with the tensor
      // core's operand reads, and a dependent LDS -> STS chain per float4 cost ~230 cycles each in the H = 256 kernel's timeline)
      auto split4v = [&](int e, float4 v, bool one_w) {
        if (one_w) v.w = 1.f;                      // x column 31 := 1 (bias gradient)
        float4* lo = e < kA4 ? lo_a4 + e : lo_b4 + (e - kA4);
        float4 hv, lv;
        if (a.raw_hi) {
          hv.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); hv.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
          hv.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); hv.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
          *lo = make_float4(v.x - hv.x, v.y - hv.y, v.z - hv.z, v.w - hv.w);
          if (one_w) raw[e] = v;
          return;
        }
        hv.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xFFFFE000u); lv.x = v.x - hv.x;
        hv.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xFFFFE000u); lv.y = v.y - hv.y;
        hv.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xFFFFE000u); lv.z = v.z - hv.z;
        hv.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xFFFFE000u); lv.w = v.w - hv.w;
        raw[e] = hv;
        *lo = lv;
      };
      static_assert(kBlk4 == kWgConv, "one float4 per thread and block");
      mbar_wait(bar_full + s, (uint32_t)((i / kWgStages) & 1));
      // ---- half 0: dz, the B operand (shared by both halves) and gate columns [0, 128)
      if (i > 0) mbar_wait(bar_lo_free + 0, (uint32_t)((i - 1) & 1));    // MMAs of (slab i-1, half 0) and everything before retired
      if (i > 0 && (i % kChain) == 0) {                                   // chain complete: S += D before slab i overwrites D
        mbar_wait(bar_lo_free + 1, (uint32_t)((i - 1) & 1));
        bool m_any, h_any;
        chain_flags(i - kChain, i - 1, m_any, h_any);
        drain(m_any, h_any);
      }
      {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool one_w = (ct & 7) == 2 * (3 ^ ((ct >> 3) & 3)) + 1;     // x column 31; row r's unit 3 sits at unit 3 ^ (r & 3)
        float4 vd = z, vx = z, vh0 = z, vh1 = z, va[4];
        if (dzb) vd = raw[ct];                                            // dz(t-1)
        if (main) vx = raw[kA4 + 2 * kBlk4 + ct];                         // x(t)
        if (t > 0) { vh0 = raw[kA4 + ct]; vh1 = raw[kA4 + kBlk4 + ct]; }  // h(t-1)
        if (main) {
#pragma unroll
          for (int k = 0; k < 4; ++k) va[k] = raw[kBlk4 + ct + k * kWgConv];
        }
        if (dzb) split4v(ct, vd, false);
        if (main) {
          split4v(kA4 + 2 * kBlk4 + ct, vx, one_w);
        } else {                                                          // slot t = T: no x, only the ones column
          raw[kA4 + 2 * kBlk4 + ct] = make_float4(0.f, 0.f, 0.f, one_w ? 1.f : 0.f);
          lo_b4[2 * kBlk4 + ct] = z;
        }
        if (t > 0) {
          split4v(kA4 + ct, vh0, false);
          split4v(kA4 + kBlk4 + ct, vh1, false);
        } else {                                                          // h(-1) = 0
          raw[kA4 + ct] = z; lo_b4[ct] = z;
          raw[kA4 + kBlk4 + ct] = z; lo_b4[kBlk4 + ct] = z;
        }
        if (main) {
#pragma unroll
          for (int k = 0; k < 4; ++k) split4v(kBlk4 + ct + k * kWgConv, va[k], false);
        }
      }
      fence_async_smem();
      mbar_arrive(bar_conv + 0);
      // ---- half 1: gate columns [128, 256)
      if (i > 0) mbar_wait(bar_lo_free + 1, (uint32_t)((i - 1) & 1));
      if (main) {
        float4 va[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) va[k] = raw[5 * kBlk4 + ct + k * kWgConv];
#pragma unroll
        for (int k = 0; k < 4; ++k) split4v(5 * kBlk4 + ct + k * kWgConv, va[k], false);
      }
      fence_async_smem();
      mbar_arrive(bar_conv + 1);
    }
    // last chain, then this CTA's sums
    mbar_wait(bar_done, 0);
    if (n_my > 0) {
      bool m_any, h_any;
      chain_flags((n_my - 1) / kChain * kChain, n_my - 1, m_any, h_any);
      drain(m_any, h_any);
    }
    tc_fence_after();
    const int row = hf * 128 + (warp & 3) * 32 + lane;
    float* dst = a.partial + ((int64_t)blockIdx.x * 256 + row) * kND;
#pragma unroll 1
    for (int c = 0; c < kND; c += 16) {
      uint32_t v[16];
      tmem_ld16(lane_tm + 256 + dcol + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(dst + c + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
    if (head && hf == 0) {                                   // head^T row (h unit / x column) = TMEM lane
      float* dsth = a.partial + (int64_t)gridDim.x * 256 * kND + ((int64_t)blockIdx.x * kHeadRows + (warp & 3) * 32 + lane) * 32;
#pragma unroll 1
      for (int c = 0; c < a.NHz; c += 16) {
        uint32_t v[16];
        tmem_ld16(lane_tm + 256 + hcol + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(dsth + c + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// partial [n_part][256][kND] -> dw_hh [256][64], dw_ih [256][I], db [256] (row = g*64 + u from n = 4u + g);
// head partial [n_part][kHeadRows][32] (transposed: row = h unit k, row 95 = ones) -> dw_head [C][64], db_head [C]
__global__ void lstm_tc_wgrad_finish_kernel(const float* __restrict__ partial, int n_part, int I, int C,
                                            float* __restrict__ dw_ih, float* __restrict__ dw_hh, float* __restrict__ db,
                                            float* __restrict__ dw_head, float* __restrict__ db_head) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < 256 * kND) {
    const int n = e / kND, col = e - n * kND;
    if (col >= kH64 + I && col != kND - 1) return;
    float acc = 0.f;
    for (int p = 0; p < n_part; ++p) acc += __ldg(partial + (int64_t)p * 256 * kND + e);
    const int u = n >> 2, g = n & 3, row = g * kH64 + u;
    if (col < kH64) dw_hh[row * kH64 + col] = acc;
    else if (col < kH64 + I) dw_ih[row * I + (col - kH64)] = acc;
    else db[row] = acc;
  } else if (dw_head != nullptr) {
    const int eh = e - 256 * kND;                   // (c, k) with k = 64: the bias gradient
    if (eh >= C * (kH64 + 1)) return;
    const int c = eh / (kH64 + 1), k = eh - c * (kH64 + 1);
    const int row = k < kH64 ? k : kND - 1;
    const float* ph = partial + (int64_t)n_part * 256 * kND + row * 32 + c;
    float acc = 0.f;
    for (int p = 0; p < n_part; ++p) acc += __ldg(ph + (int64_t)p * kHeadRows * 32);
    if (k < kH64) dw_head[c * kH64 + k] = acc;
    else if (db_head) db_head[c] = acc;
  }
}

static int wg_grid(int64_t slots) { return (int)std::min<int64_t>(sm_count(), slots); }

int64_t lstm_tc_wgrad_workspace(int64_t B, int64_t T, int64_t I) {
  (void)B; (void)T; (void)I;
  return sizeof(float) * (256 * kND + kHeadRows * 32) * std::max<int64_t>(sm_count(), 1);
}

int lstm_tc_wgrad(const float* dG, const float* h_seq, const float* x_ri, int64_t B, int64_t T, int64_t I, int64_t IP,
                  float* dw_ih, float* dw_hh, float* db, void* ws, int64_t ws_bytes, const hy2dl_head_wgrad* head,
                  int single, cudaStream_t s) {
  const int64_t need = lstm_tc_wgrad_workspace(B, T, I);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_wgrad(tf32x3): workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  HY_REQUIRE(IP == 32 && I <= 31, HY2DL_ERR_SHAPE, "hy2dl_lstm_wgrad(tf32x3): 32-column RI32 input stream with input_size <= 31 (IP=%lld I=%lld)",
             (long long)IP, (long long)I);
  TcWgradArgs a{};
  a.dG = dG; a.h = h_seq; a.x = x_ri; a.partial = reinterpret_cast<float*>(ws);
  a.T = T; a.G = cdiv(B, 32);
  int C = 0;
  if (head) {
    HY_REQUIRE(head->dz && head->dw && head->n_columns > 0 && head->n_columns <= 32 && head->t0 >= 0 && head->t0 < T, HY2DL_ERR_SHAPE,
               "hy2dl_lstm_wgrad: bad head description (columns=%d t0=%lld)", head->n_columns, (long long)head->t0);
    HY_REQUIRE(aligned16(head->dz), HY2DL_ERR_ALIGN, "hy2dl_lstm_wgrad: head dz is not 16-byte aligned");
    a.dz = head->dz; a.dz_t0 = head->t0; C = head->n_columns;
  }
  a.NHz = C <= 16 ? 16 : 32;
  static const int raw_hi = [] { const char* e = getenv("HY2DL_WG_RAWHI"); return e ? atoi(e) : 1; }();   // 0: write the rounded hi back in place
  a.raw_hi = raw_hi; a.single = single;
  const size_t smem = (size_t)kWgStages * kStageBytes + kABytes + 2 * kBBytes + kBlk + 128 + 1024;
  const int grid = wg_grid((T + (head ? 1 : 0)) * a.G);
  cudaFuncSetAttribute(lstm_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lstm_tc_wgrad_kernel<<<grid, kWgThreads, smem, s>>>(a);
  const int n_out = 256 * kND + (head ? C * (kH64 + 1) : 0);
  lstm_tc_wgrad_finish_kernel<<<(n_out + 255) / 256, 256, 0, s>>>(a.partial, grid, (int)I, C, dw_ih, dw_hh, db,
                                                                 head ? head->dw : nullptr, head ? head->db : nullptr);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

}  // namespace tc
}  // namespace hy2dl
