// Weight gradients of the fp32 (row-major tape) recurrent path on the tensor cores, H = 128 / 256:
//   dW[4H x (H + I + 1)] = sum over rows r = (t, b) of dG[r][:]^T * [h_{t-1}[b] | x_t[b] | 1]
// i.e. dW_hh, dW_ih and db in one time-parallel GEMM whose contraction index is the flattened (t, b) row.
// The forward / BPTT kernels of this path are CUDA-core kernels (lstm_fp32.cu); their weight-gradient GEMM was a
// third of the step (67.7 ms of 188 ms at H = 256, B = 9472, T = 365: 30 TFLOP/s of fp32 FMA).  This kernel is the
// tcgen05 pipeline of lstm_tc_wgrad.cu (3-term TF32 split, MN-major SWIZZLE_128B_BASE32B operands, chained TMEM
// accumulators with a round-to-nearest running sum because the tensor core's fp32 accumulation truncates) fed from
// this chain's tapes -- dG in the RB8 layout [T][ceil(B/32)][4H/8][32 rows][8] (tc_common.cuh), h_seq [T][B][H] and xT
// [T][B][IP] row-major -- instead of RI32 slabs; a slab is the 32 sequences of one (t, group):
//   * the output is cut into tiles of 256 gate rows x 96 or 128 feature columns (3 or 4 blocks of 32, whichever
//     needs fewer tiles); tile type (mt, nt) is
//     owned by `parts` CTAs that share its 32-row slabs round-robin, so all 148 SMs work and every CTA keeps its
//     tile in TMEM for the whole kernel;
//   * the converter threads themselves copy a slab two iterations ahead with 16-byte cp.async (zero-fill for rows
//     before t = 0, rows past the end and feature columns past IP) straight into the swizzled operand image: thread
//     ct owns chunk (row ct/8, 16-byte column ct%8) of each of the 11 blocks, so its addresses are loop constants;
//     completion is tracked by an mbarrier (cp.async.mbarrier.arrive.noinc).  A single producer warp computing 88
//     chunk addresses per lane and slab made the kernel producer-bound (49 ms instead of 13 at H = 256);
//   * the ones column for the bias gradient is feature H + I, patched by the converter warps.
// Gate columns keep this path's order n = g*H + u, which is the row order of dW, so no permutation is needed.
// Deterministic: per-CTA sums are combined in a fixed order by the finish kernel.
#include "common.cuh"
#include "tc_common.cuh"

namespace hy2dl {
namespace tc {

namespace {
constexpr int kThreads = 320;               // warp 0 idle, warp 1 MMA issuer, warps 2-9 loaders / converters / drain
constexpr int kStages = 3;
constexpr int kConv = 256;
constexpr int kChain = 8;                   // slabs accumulated in TMEM before a drain (96 MMAs per accumulator)
constexpr int kBlk = 4096;                  // one [32 rows][32 floats] block
constexpr int kBlk4 = kBlk / 16;
constexpr int kABytes = 8 * kBlk;           // 256 gate columns
// feature columns per tile: NB = 3 or 4 blocks of 32 (N = 96 / 128), whichever needs fewer tiles for H + I + 1 columns

struct Args {
  const float* dG;    // RB8, F = 4H, column g*H + u
  const float* h;     // [T][B][H]; the h operand of (t, b) is h[t-1][b] (zero for t = 0)
  const float* x;     // [T][B][IP]
  float* partial;     // [parts][MT*NT][256][32 NB]
  int64_t T, B, G;
  int H, I, IP, MT, NT, NB, parts;
  int single;         // HY2DL_PREC_TF32: hi*hi MMAs only
  long long* stamps;  // -DHY2DL_US_STAMPS + HY2DL_US_DBG=128: clock64 timeline of CTA 0, slabs 100-101
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;             // src-size 0: the 16 destination bytes are zero-filled, src is not read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {   // arrives on bar when this thread's prior cp.async are done
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace

#ifdef HY2DL_US_STAMPS
#define WG_STAMP(id) do { if (a.stamps && blockIdx.x == 0 && i >= 100 && i < 102) a.stamps[(i - 100) * 16 + (id)] = clock64(); } while (0)
#else
#define WG_STAMP(id) do { } while (0)
#endif
__global__ void __launch_bounds__(kThreads, 1) lstm_wgrad_rm_tc_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int NB = a.NB, ND = 32 * NB, bbytes = NB * kBlk, stage_bytes = kABytes + bbytes;
  uint8_t* s_lo = smem_raw + kStages * stage_bytes;              // [A_lo | B_lo[0] | B_lo[1]]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_lo + kABytes + 2 * bbytes);
  uint64_t* bar_full = bars;                        // [stages] cp.async completion of all converter threads
  uint64_t* bar_raw_free = bars + kStages;          // (unused slots)
  uint64_t* bar_conv = bars + 2 * kStages;          // [2] converters -> MMA, per half of the gate columns
  uint64_t* bar_lo_free = bar_conv + 2;             // [2] MMA commit -> converters, per half
  uint64_t* bar_done = bar_conv + 4;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_conv + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int TT = a.MT * a.NT;
  const int tt = blockIdx.x % TT, part = blockIdx.x / TT;
  const int mt = tt / a.NT, nt = tt - mt * a.NT;
  const int64_t n_slabs = a.T * a.G;
  const int64_t n_my = part < n_slabs ? (n_slabs - part + a.parts - 1) / a.parts : 0;   // slabs part, part + parts, ...
  const int HB = a.H / 32;                           // feature blocks: [0, HB) h, then x blocks, then nothing
  const int ones_f = a.H + a.I;                      // feature index of the ones column

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bar_full + s, kConv);
    for (int hf = 0; hf < 2; ++hf) { mbar_init(bar_conv + hf, kConv); mbar_init(bar_lo_free + hf, 1); }
    mbar_init(bar_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    // idle: the loads are issued by the converter threads
  } else if (warp == 1) {
    // =========================== MMA issuer =====================================================
    if (elect_one()) {
      const uint32_t lo_a = smem_u32(s_lo);
      const uint32_t idesc = idesc_tf32(128, ND, 1, 1);
      for (int64_t i = 0; i < n_my; ++i) {
        const int s = (int)(i % kStages);
        const uint32_t hi_a = smem_u32(smem_raw + s * stage_bytes), hi_b = hi_a + kABytes;
        const uint32_t lo_b = lo_a + kABytes + (uint32_t)(i & 1) * bbytes;
        const bool first = (i % kChain) == 0;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(bar_conv + hf, (uint32_t)(i & 1));
          tc_fence_after();
          WG_STAMP(8 + 2 * hf);
          const uint32_t d = tmem + ND * hf;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t b_hi = smem_desc_mn32(hi_b + kb * 1024, kBlk, 512), b_lo = smem_desc_mn32(lo_b + kb * 1024, kBlk, 512);
            const uint64_t a_hi = smem_desc_mn32(hi_a + hf * 4 * kBlk + kb * 1024, kBlk, 512);
            const uint64_t a_lo = smem_desc_mn32(lo_a + hf * 4 * kBlk + kb * 1024, kBlk, 512);
            if (!a.single) {
              mma_tf32_ss(d, a_lo, b_hi, idesc, !(first && kb == 0));
              mma_tf32_ss(d, a_hi, b_lo, idesc, 1);
            }
            mma_tf32_ss(d, a_hi, b_hi, idesc, !a.single || !(first && kb == 0));
          }
          mma_commit(bar_lo_free + hf);
          WG_STAMP(9 + 2 * hf);
        }
      }
      mma_commit(bar_done);
    }
    __syncwarp();
  } else {
    // =========================== converters, drain, epilogue ====================================
    const int ct = tid - 64;
    const int hf = ct >> 7;
    const uint32_t lane_tm = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t dcol = ND * hf;
    {
      uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < ND; c += 8) tmem_st8(lane_tm + 256 + dcol + c, z);
      tmem_st_wait();
    }
    auto drain = [&]() {                                     // S += D, round to nearest, in registers
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < ND; c += 16) {
        uint32_t d[16], sacc[16];
        tmem_ld16(lane_tm + dcol + c, d);
        tmem_ld16(lane_tm + 256 + dcol + c, sacc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) sacc[j] = __float_as_uint(__uint_as_float(sacc[j]) + __uint_as_float(d[j]));
        tmem_st8(lane_tm + 256 + dcol + c, *reinterpret_cast<uint32_t(*)[8]>(&sacc[0]));
        tmem_st8(lane_tm + 256 + dcol + c + 8, *reinterpret_cast<uint32_t(*)[8]>(&sacc[8]));
      }
      tmem_st_wait();
      tc_fence_before();
    };
    // where the ones column sits in this CTA's B part (block, float4 within the block row, component), if at all
    const int ones_blk = ones_f / 32 - NB * nt;              // block of the B part (0 .. NB-1) or outside
    const int ones_col = ones_f & 31;
    // loads: this thread's chunk of every block of slab j -> stage j % kStages
    const int lrow = ct >> 3, lch = ct & 7;
    const uint32_t ldst = (uint32_t)(lrow * 128 + ((((lch >> 1) ^ (lrow & 3)) << 5) | ((lch & 1) << 4)));
    const int F4 = 4 * a.H;
    // The loader threads are this kernel's critical resource (ncu source view: they are busy, not waiting), so everything
    // that does not depend on the slab is hoisted: slab j of this CTA is (t, grp) = divmod(part + j * parts, G), advanced
    // incrementally (a 64-bit division per slab and thread was 6 % of the kernel's samples), the per-thread offsets inside a
    // slab are constants, and the B-part source of every block is decided once.
    const int64_t step_t = a.parts / a.G, step_g = a.parts - step_t * a.G;       // divmod(parts, G), once
    int64_t ld_j = 0, ld_t = part / a.G, ld_g = part - (part / a.G) * a.G;       // next slab to load
    const int64_t g_off = ((int64_t)(32 * mt + (lch >> 1))) * 256 + lrow * 8 + (lch & 1) * 4;
    const int64_t slab_floats = (int64_t)(F4 >> 3) * 256;                        // dG floats per (t, group) slab
    int64_t boff[4];                                                             // B block k: offset inside row r of h (< 0: x), or unused
    bool bx[4], bvalid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int fb = NB * nt + k;
      bx[k] = fb >= HB;
      const int col = bx[k] ? 32 * (fb - HB) + 4 * lch : 32 * fb + 4 * lch;
      bvalid[k] = k < NB && (!bx[k] || col < a.IP);
      boff[k] = col;
    }
    auto load_slab = [&]() {
      const int s = (int)(ld_j % kStages);
      if (ld_j < n_my) {
        const int64_t slab = ld_t * a.G + ld_g;
        const int64_t b = ld_g * 32 + lrow;
        const bool in = b < a.B;
        const int64_t r = ld_t * a.B + b;              // row of the row-major streams
        const uint32_t st = smem_u32(smem_raw + s * stage_bytes) + ldst;
        // RB8: 8-float block (256 mt + 32 k) / 8 + lch / 2 of sequence lrow, 16-byte half lch & 1
        const float* g = a.dG + slab * slab_floats + g_off;
#pragma unroll
        for (int k = 0; k < 8; ++k) cp_async16_zfill(st + k * kBlk, in ? g + 4 * k * 256 : a.dG, in);
        const float* hrow = a.h + (r - a.B) * a.H;
        const float* xrow = a.x + r * a.IP;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < NB) {
            const bool valid = in && bvalid[k] && (bx[k] || ld_t > 0);
            cp_async16_zfill(st + (8 + k) * kBlk, valid ? (bx[k] ? xrow : hrow) + boff[k] : a.dG, valid);
          }
        }
      }
      cp_async_arrive(bar_full + s);     // also for slabs past the end: nobody waits for those phases
      ++ld_j;
      ld_t += step_t; ld_g += step_g;
      if (ld_g >= a.G) { ld_g -= a.G; ++ld_t; }
    };
    load_slab();
    load_slab();
    float4* lo_a4 = reinterpret_cast<float4*>(s_lo);
    // float4 index (A part first) of the ones column among THIS thread's float4s of a stage, or -1: thread ct handles the
    // B-part float4s eb = ct, ct + kConv, ...; float4 `pos` of row `row` of block `blk` holds features
    // 8 ((pos >> 1) ^ (row & 3)) + 4 (pos & 1) .. + 3 of the block
    int ones_e = -1;
    for (int eb = ct; eb < NB * kBlk4; eb += kConv) {
      const int blk = eb / kBlk4, w = eb - blk * kBlk4, row = w >> 3, pos = w & 7;
      if (blk == ones_blk && (((pos >> 1) ^ (row & 3)) * 8 + (pos & 1) * 4) == (ones_col & ~3)) ones_e = kABytes / 16 + eb;
    }
    for (int64_t i = 0; i < n_my; ++i) {
      const int s = (int)(i % kStages);
      float4* raw = reinterpret_cast<float4*>(smem_raw + s * stage_bytes);
      float4* lo_b4 = reinterpret_cast<float4*>(s_lo + kABytes + (i & 1) * bbytes);
      constexpr int kA4 = kABytes / 16;
      // lo = v - trunc(v); the landed words are the hi operand.  Loads first, then arithmetic and stores: the shared-memory
      // pipe is shared with the tensor core's operand reads, a dependent LDS -> STS chain per float4 cost ~230 cycles each
      // (clock64 timeline: 1640 cycles for the 7 float4 of half 0).
      auto lo_of = [&](const float4& v) {
        const float4 hv = make_float4(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                                      __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
        return make_float4(v.x - hv.x, v.y - hv.y, v.z - hv.z, v.w - hv.w);
      };
      auto split_a = [&](int half) {                         // this thread's 4 float4 of gate-column half `half`
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = raw[half * 4 * kBlk4 + ct + k * kConv];
#pragma unroll
        for (int k = 0; k < 4; ++k) lo_a4[half * 4 * kBlk4 + ct + k * kConv] = lo_of(v[k]);
      };
      auto split_b = [&]() {                                 // ... and its NB float4 of the B part (ones column patched)
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) if (k < NB) v[k] = raw[kA4 + ct + k * kConv];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < NB) {
            const int e = kA4 + ct + k * kConv;
            if (e == ones_e) {                               // the ones column of the bias gradient (decided once, above)
              const int q = ones_col & 3;
              if (q == 0) v[k].x = 1.f; else if (q == 1) v[k].y = 1.f; else if (q == 2) v[k].z = 1.f; else v[k].w = 1.f;
              raw[e] = v[k];
            }
            lo_b4[ct + k * kConv] = lo_of(v[k]);
          }
        }
      };
      if (ct == 0) WG_STAMP(0);
      mbar_wait(bar_full + s, (uint32_t)((i / kStages) & 1));
      if (ct == 0) WG_STAMP(1);
      // ---- half 0: the B operand (shared by both halves) and gate columns [0, 128)
      if (i > 0) mbar_wait(bar_lo_free + 0, (uint32_t)((i - 1) & 1));
      if (i > 0 && (i % kChain) == 0) {
        mbar_wait(bar_lo_free + 1, (uint32_t)((i - 1) & 1));
        drain();
      }
      if (ct == 0) WG_STAMP(2);
      split_b();
      split_a(0);
      fence_async_smem();
      mbar_arrive(bar_conv + 0);
      if (ct == 0) WG_STAMP(3);
      // ---- half 1: gate columns [128, 256)
      if (i > 0) mbar_wait(bar_lo_free + 1, (uint32_t)((i - 1) & 1));
      if (ct == 0) WG_STAMP(4);
      load_slab();           // slab i + 2: its stage held slab i - 1, whose MMAs (hi operands) have just been seen to retire
      if (ct == 0) WG_STAMP(5);
      split_a(1);
      fence_async_smem();
      mbar_arrive(bar_conv + 1);
      if (ct == 0) WG_STAMP(6);
    }
    mbar_wait(bar_done, 0);
    if (n_my > 0) drain();
    tc_fence_after();
    const int row = hf * 128 + (warp & 3) * 32 + lane;
    float* dst = a.partial + (((int64_t)part * TT + tt) * 256 + row) * ND;
#pragma unroll 1
    for (int c = 0; c < ND; c += 16) {
      uint32_t v[16];
      tmem_ld16(lane_tm + 256 + dcol + c, v);
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

// partial [parts][MT*NT][256][96] -> dw_hh [4H][H], dw_ih [4H][I], db [4H]
__global__ void lstm_wgrad_rm_tc_finish_kernel(const float* __restrict__ partial, int parts, int MT, int NT, int ND, int H, int I,
                                               float* __restrict__ dw_ih, float* __restrict__ dw_hh, float* __restrict__ db) {
  const int nf = H + I + 1;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)4 * H * nf) return;
  const int n = (int)(e / nf), f = (int)(e - (int64_t)n * nf);
  const int mt = n >> 8, nt = f / ND, TT = MT * NT;
  const float* p = partial + (((int64_t)(mt * NT + nt)) * 256 + (n & 255)) * ND + (f - nt * ND);
  float acc = 0.f;
  for (int q = 0; q < parts; ++q) acc += __ldg(p + (int64_t)q * TT * 256 * ND);
  if (f < H) dw_hh[(int64_t)n * H + f] = acc;
  else if (f < H + I) dw_ih[(int64_t)n * I + (f - H)] = acc;
  else db[n] = acc;
}

static void rm_tc_shape(int64_t H, int64_t I, int& MT, int& NT, int& NB, int& parts) {
  MT = (int)(4 * H / 256);
  const int64_t blocks = cdiv(H + I + 1, (int64_t)32);
  NB = cdiv(blocks, (int64_t)4) < cdiv(blocks, (int64_t)3) ? 4 : 3;
  NT = (int)cdiv(blocks, (int64_t)NB);
  parts = std::max(1, sm_count() / (MT * NT));
}

bool lstm_wgrad_rm_tc_supported(int64_t H) { return H == 64 || H == 128 || H == 256; }

int64_t lstm_wgrad_rm_tc_workspace(int64_t H, int64_t I) {
  int MT, NT, NB, parts;
  rm_tc_shape(H, I, MT, NT, NB, parts);
  return sizeof(float) * (int64_t)parts * MT * NT * 256 * 32 * NB;
}

int lstm_wgrad_rm_tc(const float* dG, const float* h_seq, const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H,
                     float* dw_ih, float* dw_hh, float* db, void* ws, int64_t ws_bytes, int single, cudaStream_t s) {
  const int64_t need = lstm_wgrad_rm_tc_workspace(H, I);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_wgrad: workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  Args a{};
  a.dG = dG; a.h = h_seq; a.x = xT; a.partial = reinterpret_cast<float*>(ws);
  a.T = T; a.B = B; a.G = cdiv(B, (int64_t)32); a.H = (int)H; a.I = (int)I; a.IP = (int)IP;
  rm_tc_shape(H, I, a.MT, a.NT, a.NB, a.parts);
  a.single = single;
  const size_t bbytes = (size_t)a.NB * kBlk;
  const size_t smem = (size_t)kStages * (kABytes + bbytes) + kABytes + 2 * bbytes + 128 + 1024;
  cudaFuncSetAttribute(lstm_wgrad_rm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#ifdef HY2DL_US_STAMPS
  cudaMalloc(&a.stamps, 32 * sizeof(long long)); cudaMemset(a.stamps, 0, 32 * sizeof(long long));
#endif
  lstm_wgrad_rm_tc_kernel<<<a.MT * a.NT * a.parts, kThreads, smem, s>>>(a);
#ifdef HY2DL_US_STAMPS
  {
    long long h[32];
    cudaDeviceSynchronize();
    cudaMemcpy(h, a.stamps, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(a.stamps);
    for (int i = 0; i < 32; ++i) if (h[i]) fprintf(stderr, "WG_STAMP slab %d id %2d  %8lld\n", 100 + i / 16, i % 16, h[i] - h[0]);
  }
#endif
  const int64_t n_out = 4 * H * (H + I + 1);
  lstm_wgrad_rm_tc_finish_kernel<<<(unsigned)cdiv(n_out, (int64_t)256), 256, 0, s>>>(a.partial, a.parts, a.MT, a.NT, 32 * a.NB, (int)H, (int)I,
                                                                                 dw_ih, dw_hh, db);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

}  // namespace tc
}  // namespace hy2dl
