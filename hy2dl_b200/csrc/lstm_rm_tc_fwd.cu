// Recurrent LSTM forward on the tensor cores for H = 128 / 256 (row-major tapes of the fp32 path).
//
// For H = 64 the packed weights (196 KB hi + lo) stay in shared memory and the A operand [x | h] (hi + lo) in TMEM
// (lstm_tc.cu).  Neither fits here: [W_ih | W_hh] hi + lo is 2.4 MB at H = 256, and [x | h] hi + lo for 128 sequences
// is 2 x 288 TMEM columns > 512 (295 KB in shared memory).  This kernel therefore
//   * keeps the h part of A_hi in TMEM (H columns, TS-mode MMAs for hi*hi and hi*lo), the x part of A_hi and all of
//     A_lo in shared memory (K-major images, SS-mode MMAs);
//   * streams the packed weights every step from L2 through a ring of 8 KB bulk copies (one copy = one K step of 8
//     rows x 128 gate columns, hi | lo), chunk after chunk of 32 hidden units.  N = 128 per MMA matters: a first
//     version with 32-column chunks ran 60+ cycles per MMA where 16 were budgeted (small-N MMAs pay a fixed cost);
//   * accumulates a chunk in SHORT chains.  The tensor core's fp32 accumulation truncates; the loss of a chain of L
//     accumulating hi*hi MMAs (~0.25 * 2^-23 * L of the sum, toward zero) repeats from step to step for a slowly varying
//     state, so it compounds linearly over the recurrence (tools/dbg_step_noise.py, tools/sim_trunc.py: one chain of
//     K/8 = 36 gave a parameter-gradient error of 1.2e-4 on the reference golden of BASELINE configs[2], B = 8, T = 365).
//     The K steps of a chunk are cut into splits of kSplitJ = 4: a split first issues its 8 correction MMAs
//     (a_lo*b_hi, a_hi*b_lo: 2^-11 of the result, their truncation costs nothing) and then its 4 hi*hi MMAs into ONE
//     128-column accumulator; the two accumulator regions of TMEM (H + 2 x 128 columns = all of it at H = 256) alternate
//     between splits, so a region is drained by the row threads -- who add the partial sums in registers, round to
//     nearest -- while the tensor core fills the other one.  A split's 4 weight slices stay in the ring until its hi*hi
//     MMAs have been issued (they are read twice);
//   * keeps no recurrent state in registers: the row thread re-reads c_{t-1} and, at the end of the step, its own
//     h_t values from a two-slot scratch it has just written (L2 hits) to rebuild A for the next step.  A row is owned by
//     two threads (TMEM lane = row; the two warps of a lane quarter split every chunk's 32 units 16 | 16), and a
//     thread only ever re-reads what it wrote itself, so no CTA-wide synchronisation is needed inside a step.
// Timing experiments on the first complete version (H = 256, B = 256, 30.8 ms): without tape stores 20.7 ms (they
// were row-strided 16-byte stores then; whole sectors now), without the activation arithmetic 30.6 ms, a third of
// the MMAs 27.1 ms, all three 12.3 ms; ring depth 2 / 4 / 8 and one or two K steps per commit: no difference.
// One CTA = 128 sequences for all T steps; 10 warps: 8 row warps, the MMA issuer, the bulk-copy producer.
// Per step the tensor core needs 128 x 4H x K x 3 / 2048 cycles (55 k at H = 256) while 2.4 MB of weights cross
// L2 -> shared memory per CTA: at a full wave of 148 CTAs the L2 stream, not the tensor pipe, bounds the step.
//
// Semantics: torch.nn.LSTM, one layer, h0 = c0 = 0 (modelzoo/cudalstm.py:43-50, hybrid.py:84-92); tapes as
// h_seq [T][B][H] row-major (it is an API output); c_seq and gates (post-activation, column g*H + u) are tapes only
// this chain reads and are stored in the RB8 layout (tc_common.cuh), T x ceil(B/32)*32 rows.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace hy2dl {
namespace tc {

namespace {
constexpr int kRowsF = 128;
constexpr int kThreadsF = 320;              // warps 0-7 rows, 8 MMA issuer, 9 producer
constexpr int kChunkN = 128;                // gate columns per chunk = 32 hidden units x 4 gates
constexpr int kStepBytes = 8 * kChunkN * 4 * 2;    // one K step of the packed weights: hi | lo = 8 KB
constexpr int kSliceJ = 1;                         // K steps per ring slice (one bulk copy, one commit); 2 measured the same
constexpr int kSliceBytes = kSliceJ * kStepBytes;
constexpr uint32_t kColAcc = 256;                   // accumulator regions [256, 384) and [384, 512), alternating between splits
constexpr int kSplitJ = 4;                         // K steps per split = accumulating hi*hi MMAs per chain (2 if the ring is short)
constexpr int kMaxKX = 80;

struct FwdArgs {
  const float* x;       // [T][B][IP]
  const float* wpk;     // packed weights [H/32 chunks][K/8 steps][hi [2][128][4] | lo [2][128][4]]
  float* h_seq;         // [T][B][H] row-major API output, or null
  float* h_own;         // RB8 [2 slots][ceil(B/32)*32][H]: h of the last two steps, what the row threads re-read
  float* c_seq;         // RB8 tape [T][..][H], or two slots of scratch (c_pingpong)
  float* gates;         // RB8 tape, F = 4H, or null
  float* h_last;        // [B][H] or null
  int64_t B, T;
  int H, I, IP, KX, K;  // KX = x columns incl. the ones column, rounded to 8; K = KX + H
  int stages;           // ring depth (what shared memory allows)
  int c_pingpong;
  int act_sep;          // experiment: one reciprocal per gate
  int splitj;           // K steps per split (a split's weight slices stay in the ring until its hi*hi MMAs are issued)
  int single;           // HY2DL_PREC_TF32: hi*hi MMAs only (one chain per chunk), only the hi half of the weights is streamed
  long long* stamps;    // -DHY2DL_US_STAMPS: clock64 timeline of CTA 0, steps 100-101
  int dbg;              // timing experiments (HY2DL_RM_DBG): 1 no tape stores, 2 no correction MMAs, 4 no drains but the last, 8 no A rebuild
};

// packed weight element (chunk c, K index k, column n = 4 uu + g of the chunk): rows scaled by -log2(e) (x 2 for g)
// and the truncation compensation; k < I: W_ih, k == I: bias, k in [KX, KX + H): W_hh, else 0.
__global__ void rm_tc_pack_w_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                    const float* __restrict__ b_hh, int H, int I, int KX, int K, float comp, int pair,
                                    float* __restrict__ out) {
  const int NJ = K / 8, NC = H / 32;
  const int64_t total = (int64_t)NC * K * kChunkN;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(e & 3), n = (int)((e >> 2) & 127), kc = (int)((e >> 9) & 1);
    const int64_t rest = e >> 10;
    const int J = (int)(rest % NJ), c = (int)(rest / NJ);
    const int k = J * 8 + kc * 4 + kk, uu = n >> 2, g = n & 3, row = g * H + c * 32 + uu;
    float v = 0.f;
    if (k < I) v = w_ih[(int64_t)row * I + k];
    else if (k == I) v = b_ih[row] + b_hh[row];
    else if (k >= KX) v = w_hh[(int64_t)row * H + (k - KX)];
    v *= (g == 2 ? -2.f : -1.f) * 1.4426950408889634f * comp;
    uint32_t hi, lo;
    split_tf32_rr(v, hi, lo);
    if (pair) {      // CTA-pair layout: per K step [rank 0: hi [2][64][4] | lo [2][64][4]] [rank 1: hi | lo]; rank r holds columns 64 r ..
      float* dst = out + ((int64_t)c * NJ + J) * (kStepBytes / 4) + (n >> 6) * (kStepBytes / 8) + (kc * 64 + (n & 63)) * 4 + kk;
      dst[0] = __uint_as_float(hi);
      dst[kStepBytes / 16] = __uint_as_float(lo);
    } else {
      float* dst = out + ((int64_t)c * NJ + J) * (kStepBytes / 4) + (kc * kChunkN + n) * 4 + kk;
      dst[0] = __uint_as_float(hi);
      dst[kStepBytes / 8] = __uint_as_float(lo);
    }
  }
}
}  // namespace

// PAIR: the kernel runs as clusters of two CTAs on the SMs of one TPC (cta_group::2).  The leader's thread issues every MMA with
// M = 256 -- 128 sequences from each CTA -- and each CTA streams only ITS half of every weight slice (N/2 = 64 gate columns), so
// the L2 -> shared-memory weight stream per sequence halves (it bounds the single-CTA kernel at a full wave: 148 CTAs x 2.4 MB
// per step).  Row threads, tapes and the accumulator hand-shake are per CTA as before; completions are multicast to both CTAs,
// the peer's arrivals reach the leader's mbarriers through the cluster's shared-memory window.
#ifdef HY2DL_US_STAMPS
#define RF_STAMP(id) do { if (a.stamps && blockIdx.x == 0 && t >= 100 && t < 102) a.stamps[(t - 100) * 64 + (id)] = clock64(); } while (0)
#else
#define RF_STAMP(id) do { } while (0)
#endif
template <bool PAIR>
__global__ void __launch_bounds__(kThreadsF, 1) lstm_rm_tc_fwd_kernel(FwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int K = a.K, H = a.H, KX = a.KX;
  constexpr int kSlot = PAIR ? kSliceBytes / 2 : kSliceBytes;    // bytes of one K step in THIS CTA's ring: hi | lo of its columns
  const int NS = PAIR ? 2 * a.stages : a.stages;                 // ring slots (<= 16)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  float* s_alo = reinterpret_cast<float*>(smem_raw);                       // [K/4][128 rows][4]
  float* s_xhi = s_alo + (size_t)K * kRowsF;                               // [KX/4][128 rows][4]
  uint8_t* s_ring = reinterpret_cast<uint8_t*>(s_xhi + (size_t)KX * kRowsF);   // NS slots of kSlot bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)a.stages * kSliceBytes);
  uint64_t* bar_full = bars;                       // [NS <= 16] producer (+ the peer's relay) -> MMA
  uint64_t* bar_free = bars + 16;                  // [NS] MMA commit -> producer
  uint64_t* bar_dfull = bars + 32;                 // [2] MMA commit -> rows: the split's accumulator region is complete
  uint64_t* bar_dfree = bars + 34;                 // [2] rows -> MMA: it has been read
  uint64_t* bar_a = bars + 36;                     // rows -> MMA: A operand of the step is in place
  uint64_t* bar_step = bars + 37;                  // MMA commit: every MMA of the step retired, A may be rebuilt
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 38);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NC = H / 32, NJ = K / 8, JX = KX / 8;
  const int SJ = a.splitj;
  const int NSP = (NJ + SJ - 1) / SJ;               // splits per chunk (the last one may be shorter)
  // arrivals: every local row thread arrives on bar_a / bar_dfree itself; the peer's eight row warps add one arrival each
  // (lane 0, after a warp barrier) through the cluster window.  The leader's bar_full takes its own producer's
  // arrive.expect_tx plus the peer's relay (its copy of the slice has landed).
  if (tid == 0) {
    const uint32_t rows_cnt = PAIR ? 256 + 8 : 256;
    for (int s = 0; s < NS; ++s) { mbar_init(bar_full + s, (PAIR && rank == 0) ? 2 : 1); mbar_init(bar_free + s, 1); }
    for (int r = 0; r < 2; ++r) { mbar_init(bar_dfull + r, 1); mbar_init(bar_dfree + r, rows_cnt); }
    mbar_init(bar_a, rows_cnt);
    mbar_init(bar_step, 1);
    mbar_fence_init();
  }
  if (warp == 8) { if (PAIR) tmem_alloc2(s_tmem, 512); else tmem_alloc(s_tmem, 512); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();                        // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  // a row thread's arrival on a barrier the MMA issuer waits for
  auto rows_arrive = [&](uint64_t* bar) {
    if (!PAIR || rank == 0) { mbar_arrive(bar); return; }
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(mapa_rank(smem_u32(bar), 0));
  };

  if (warp < 8) {
    // =========================== row threads ===================================================
    const int q = warp & 3, hh = warp >> 2;                       // lane quarter, which 16 of every 32 units
    const int rloc = q * 32 + lane;
    const int64_t b = (int64_t)blockIdx.x * kRowsF + rloc;
    const bool live = b < a.B;
    const uint32_t lane_tm = tmem + ((uint32_t)(q * 32) << 16);
    float4* alo4 = reinterpret_cast<float4*>(s_alo);
    float4* xhi4 = reinterpret_cast<float4*>(s_xhi);
    const int64_t BH = a.B * H;
    const int64_t G = (a.B + 31) >> 5;
    auto hown = [&](int64_t t) { return a.h_own + rb8_off(t & 1, G, b, H, 2 * hh); };                        // RB8, block 2 hh
    auto cbuf = [&](int64_t t) { return a.c_seq + rb8_off(a.c_pingpong ? (t & 1) : t, G, b, H, 2 * hh); };   // RB8, block 2 hh
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto split4 = [](const float4& v, uint32_t (&hi)[4], float4& lo) {
      uint32_t l[4];
      split_tf32(v.x, hi[0], l[0]); split_tf32(v.y, hi[1], l[1]); split_tf32(v.z, hi[2], l[2]); split_tf32(v.w, hi[3], l[3]);
      lo = make_float4(__uint_as_float(l[0]), __uint_as_float(l[1]), __uint_as_float(l[2]), __uint_as_float(l[3]));
    };

    for (int64_t t = 0; t < a.T; ++t) {
      // ---- A operand of step t: x_t (threads of half 0) and this thread's units of h_{t-1}
      if (t > 0) {
        mbar_wait(bar_step, (uint32_t)((t - 1) & 1));            // all MMAs of step t-1 retired
        tc_fence_after();
      }
      if (tid == 0) RF_STAMP(0);
      if (hh == 0) {
        const float* xr = a.x + (t * a.B + b) * a.IP;
        for (int k0 = 0; k0 < KX; k0 += 4) {
          float4 v = (live && k0 < a.IP) ? __ldg(reinterpret_cast<const float4*>(xr + k0)) : z4;
          if ((a.I >> 2) == (k0 >> 2)) {                          // the ones column (bias) is input column I
            const int kq = a.I & 3;
            if (kq == 0) v.x = 1.f; else if (kq == 1) v.y = 1.f; else if (kq == 2) v.z = 1.f; else v.w = 1.f;
          }
          uint32_t hi[4]; float4 lo;
          split4(v, hi, lo);
          xhi4[(k0 >> 2) * kRowsF + rloc] = make_float4(__uint_as_float(hi[0]), __uint_as_float(hi[1]), __uint_as_float(hi[2]), __uint_as_float(hi[3]));
          alo4[(k0 >> 2) * kRowsF + rloc] = lo;
        }
      }
      {
        const float* hp = (t > 0 && live && !(a.dbg & 8)) ? hown(t - 1) : nullptr;
        // Eight 256-bit loads in flight, then their splits and stores.  The loads are volatile asm with a memory clobber (they
        // must see this thread's own earlier stores), so the compiler keeps each behind the previous block's stores: one
        // load -> split -> store per 8 units serialised 16 L2 / DRAM round trips, 22.9 k of the step's 163 k cycles (clock64
        // timeline, -DHY2DL_US_STAMPS; 13.9 k batched); the scratch (39 MB at 18 944 sequences) does not survive 133 MB of tape
        // stores in L2.  The kernel's time did not change (36.4 ms): at a full wave the step is bound elsewhere.
        if (!((a.dbg & 8) && t > 1))
        for (int c0 = 0; c0 < NC; c0 += 4) {                      // NC = H / 32 = 4 or 8
          f8 v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] = hp ? ld256(hp + (4 * (c0 + (k >> 1)) + (k & 1)) * 256) : f8_zero();
#pragma unroll
          for (int k = 0; k < 8; ++k) {                           // 8 units: one tcgen05.st, two lo chunks
            const int u0 = 32 * (c0 + (k >> 1)) + 16 * hh + 8 * (k & 1);
            uint32_t hi8[8], lo8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_tf32(v[k].v[e], hi8[e], lo8[e]);
            tmem_st8(lane_tm + u0, hi8);
            alo4[((KX + u0) >> 2) * kRowsF + rloc] =
                make_float4(__uint_as_float(lo8[0]), __uint_as_float(lo8[1]), __uint_as_float(lo8[2]), __uint_as_float(lo8[3]));
            alo4[((KX + u0) >> 2) * kRowsF + kRowsF + rloc] =
                make_float4(__uint_as_float(lo8[4]), __uint_as_float(lo8[5]), __uint_as_float(lo8[6]), __uint_as_float(lo8[7]));
          }
        }
      }
      tmem_st_wait();
      fence_async_smem();
      tc_fence_before();
      rows_arrive(bar_a);
      if (tid == 0) RF_STAMP(1);

      // ---- epilogue of every chunk: 16 hidden units of this row
      const float* cprev = (t > 0 && live) ? cbuf(t - 1) : nullptr;
      float* hout = (live && a.h_seq) ? a.h_seq + t * BH + b * H + 16 * hh : nullptr;      // row-major API output
      float* hself = live ? hown(t) : nullptr;
      float* cout = live ? cbuf(t) : nullptr;
      float* gout = (live && a.gates) ? a.gates + rb8_off(t, G, b, 4 * H, 2 * hh) : nullptr;            // RB8, F = 4H
      float* hlast = (live && a.h_last && t == a.T - 1) ? a.h_last + b * H + 16 * hh : nullptr;
      for (int c = 0; c < NC; ++c) {
        const int64_t gch = t * NC + c;
        f8 cp[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) cp[i] = cprev ? ld256(cprev + (4 * c + i) * 256) : f8_zero();
        // partial sums of the chunk's splits, added in registers (round to nearest); see the header
        float z[64];                                              // this thread's 16 units x 4 gates
        for (int sp = 0; sp < NSP; ++sp) {
          const int64_t gsp = gch * NSP + sp;                     // split counter over the whole kernel
          const int r = (int)(gsp & 1);
          mbar_wait(bar_dfull + r, (uint32_t)((gsp >> 1) & 1));
          tc_fence_after();
          uint32_t z0[32], z1[32];
          if ((a.dbg & 4) && sp != NSP - 1) { tc_fence_before(); rows_arrive(bar_dfree + r); continue; }
          tmem_ld32(lane_tm + kColAcc + 128 * r + 64 * hh, z0);
          tmem_ld32(lane_tm + kColAcc + 128 * r + 64 * hh + 32, z1);
          tmem_ld_wait();
          tc_fence_before();
          rows_arrive(bar_dfree + r);                             // the region may be overwritten
          if (sp == 0 || (a.dbg & 4)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { z[j] = __uint_as_float(z0[j]); z[32 + j] = __uint_as_float(z1[j]); }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { z[j] += __uint_as_float(z0[j]); z[32 + j] += __uint_as_float(z1[j]); }
          }
        }
        if (tid == 0) RF_STAMP(2 + c);
#pragma unroll
        for (int i = 0; i < 2; ++i) {                             // 8 units per round: whole 32-byte sectors per access
          f8 hv, cv, gi, gf, gg, go;
#pragma unroll
          for (int uu = 0; uu < 8; ++uu) {
            const float* zz = z + 32 * i + 4 * uu;
            if (a.act_sep) lstm_gates_sep(zz[0], zz[1], zz[2], zz[3], gi.v[uu], gf.v[uu], gg.v[uu], go.v[uu]);
            else lstm_gates(zz[0], zz[1], zz[2], zz[3], gi.v[uu], gf.v[uu], gg.v[uu], go.v[uu]);
            cv.v[uu] = fmaf(gf.v[uu], cp[i].v[uu], gi.v[uu] * gg.v[uu]);
            hv.v[uu] = go.v[uu] * tanh_mufu(cv.v[uu]);
          }
          if (hself && !(a.dbg & 1)) {
            const int o = 32 * c + 8 * i;
            const int blk = (4 * c + i) * 256;                    // unit block (32 c + 16 hh + 8 i) / 8 relative to block 2 hh
            st256(hself + blk, hv);                               // re-read by this thread at the end of the step
            if (hout) stg256(hout + o, hv);
            st256(cout + blk, cv);                                // re-read in the next step
            if (gout) {
              stg256(gout + blk, gi);
              stg256(gout + (H / 8) * 256 + blk, gf);
              stg256(gout + 2 * (H / 8) * 256 + blk, gg);
              stg256(gout + 3 * (H / 8) * 256 + blk, go);
            }
            if (hlast) stg256(hlast + o, hv);
          }
        }
        if (tid == 0) RF_STAMP(12 + c);
      }
    }
  } else if (warp == 8) {
    // =========================== MMA issuer (the leader's, in a pair) ===========================
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = idesc_tf32(PAIR ? 2 * kRowsF : kRowsF, kChunkN, 0, 0);
      const uint32_t a_lbo = kRowsF * 16, b_lbo = (PAIR ? kChunkN / 2 : kChunkN) * 16, sbo = 128;
      // descriptors advance by constants (the address field counts 16-byte units): no per-slice rebuilding
      const uint64_t ring_hi0 = smem_desc(smem_u32(s_ring), b_lbo, sbo), alo0 = smem_desc(smem_u32(s_alo), a_lbo, sbo);
      const uint64_t xhi0 = smem_desc(smem_u32(s_xhi), a_lbo, sbo);
      constexpr uint64_t kStageStep = kSlot / 16, kLoOff = kSlot / 32, kAStep = 2 * kRowsF * 16 / 16;
      auto mma_ss = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) { if (PAIR) mma2_tf32_ss(d, ad, bd, idesc, acc); else mma_tf32_ss(d, ad, bd, idesc, acc); };
      auto mma_ts = [&](uint32_t d, uint32_t at, uint64_t bd, uint32_t acc) { if (PAIR) mma2_tf32_ts(d, at, bd, idesc, acc); else mma_tf32_ts(d, at, bd, idesc, acc); };
      auto commit = [&](uint64_t* bar) { if (PAIR) mma2_commit(bar); else mma_commit(bar); };
      static_assert(true, "");
      // (a try_wait with .acquire.cluster costs several hundred cycles per poll: 140 us per step against 42 for the plain
      // form, measured; the plain form is what CUTLASS's ClusterBarrier uses for cluster-wide arrivals as well)
      auto wait_rows = [&](uint64_t* bar, uint32_t par) { mbar_wait(bar, par); };
      int st = 0;                                   // ring stage of the split's first K step and its phase parity
      uint32_t ph = 0;
      int64_t gsp = 0;                              // split counter
      for (int64_t t = 0; t < a.T; ++t) {
        wait_rows(bar_a, (uint32_t)(t & 1));
        tc_fence_after();
        for (int c = 0; c < NC; ++c) {
          RF_STAMP(24 + c);
          for (int Jb = 0; Jb < NJ; Jb += SJ, ++gsp) {
            const int n = min(SJ, NJ - Jb);
            const int r = (int)(gsp & 1);
            const uint32_t d = tmem + kColAcc + 128 * r;
            if (gsp >= 2) {                         // the rows have drained this region's previous split
              wait_rows(bar_dfree + r, (uint32_t)(((gsp >> 1) - 1) & 1));
              tc_fence_after();
            }
            if (a.single) {                         // single pass: every slice is read once
              for (int j = 0; j < n; ++j) {
                const int J = Jb + j;
                wait_rows(bar_full + st, ph);
                tc_fence_after();
                const uint64_t bh = ring_hi0 + (uint64_t)st * kStageStep;
                if (J < JX) mma_ss(d, xhi0 + (uint64_t)J * kAStep, bh, j > 0);
                else mma_ts(d, tmem + 8 * (uint32_t)(J - JX), bh, j > 0);
                commit(bar_free + st);
                if (++st == NS) { st = 0; ph ^= 1; }
              }
              commit(bar_dfull + r);
              continue;
            }
            // correction terms first, while the accumulator is ~2^-11 of its final size
            for (int j = 0; j < n; ++j) {
              const int J = Jb + j;
              int sj = st + j; uint32_t pj = ph;
              if (sj >= NS) { sj -= NS; pj ^= 1; }
              wait_rows(bar_full + sj, pj);
              tc_fence_after();
              const uint64_t bh = ring_hi0 + (uint64_t)sj * kStageStep, bl = bh + kLoOff;
              if (a.dbg & 2) { if (j == 0) mma_ss(d, alo0 + (uint64_t)J * kAStep, bh, 0); continue; }
              mma_ss(d, alo0 + (uint64_t)J * kAStep, bh, j > 0);
              if (J < JX) mma_ss(d, xhi0 + (uint64_t)J * kAStep, bl, 1);     // x part: A_hi from shared memory
              else mma_ts(d, tmem + 8 * (uint32_t)(J - JX), bl, 1);           // h part: A_hi in TMEM columns [0, H)
            }
            // the split's hi*hi chain; a slice is released once its second reader has been issued
            for (int j = 0; j < n; ++j) {
              const int J = Jb + j;
              int sj = st + j;
              if (sj >= NS) sj -= NS;
              const uint64_t bh = ring_hi0 + (uint64_t)sj * kStageStep;
              if (J < JX) mma_ss(d, xhi0 + (uint64_t)J * kAStep, bh, 1);
              else mma_ts(d, tmem + 8 * (uint32_t)(J - JX), bh, 1);
              commit(bar_free + sj);
            }
            commit(bar_dfull + r);
            st += n;
            if (st >= NS) { st -= NS; ph ^= 1; }
          }
        }
        commit(bar_step);
        RF_STAMP(40);
      }
    } else if (PAIR && rank == 1 && lane == 0) {
      // =========================== the peer's relay ============================================
      // tells the leader that this CTA's half of a weight slice has landed (a bulk copy can only signal a barrier of the
      // CTA it writes to)
      int st = 0;
      uint32_t ph = 0;
      const int64_t n_slices = a.T * NC * NJ;
      for (int64_t i = 0; i < n_slices; ++i) {
        mbar_wait(bar_full + st, ph);
        mbar_arrive_remote(mapa_rank(smem_u32(bar_full + st), 0));
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =========================== weight-slice producer ==========================================
    if (lane == 0) {
      const int per_step = NC * NJ / kSliceJ;
      int st = 0;
      uint32_t ph = 1;                              // parity of the previous use of the stage (first round: nothing to wait for)
      bool first_round = true;
      const uint32_t nbytes = a.single ? kSlot / 2 : kSlot;        // hi | lo
      for (int64_t t = 0; t < a.T; ++t) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wpk) + (PAIR ? rank * kSlot : 0);
        for (int i = 0; i < per_step; ++i, src += kSliceBytes) {
          if (!first_round) mbar_wait(bar_free + st, ph);
          mbar_arrive_expect_tx(bar_full + st, nbytes);
          bulk_g2s(s_ring + (size_t)st * kSlot, src, nbytes, bar_full + st);
          if (++st == NS) { st = 0; ph ^= 1; first_round = false; }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) {
    cluster_sync();                                 // both CTAs are done with TMEM and with each other's barriers
    if (warp == 8) tmem_dealloc2(tmem, 512);
  } else if (warp == 8) {
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
static void rm_fwd_shape(int64_t I, int64_t H, int& KX, int& K) {
  KX = (int)round_up(I + 1, 8 * kSliceJ);      // a ring slice holds kSliceJ K steps: K / 8 must be a multiple of it
  K = KX + (int)H;
}
// ring depth: what is left of the 227 KB after the A images
static int rm_fwd_stages(int KX, int K) {
  const int64_t fixed = (int64_t)(K + KX) * kRowsF * 4 + 512 + 1024;
  return (int)std::min<int64_t>(8, (232448 - fixed) / kSliceBytes);   // 2, 4 and 8 stages measured the same
}

bool lstm_rm_tc_fwd_supported(int64_t I, int64_t H) {
  if (!(H == 128 || H == 256)) return false;
  int KX, K;
  rm_fwd_shape(I, H, KX, K);
  return KX <= kMaxKX && rm_fwd_stages(KX, K) >= 3;
}

// the unit-split kernel for small batches (lstm_us_tc_fwd.cu)
bool lstm_us_tc_fwd_supported(int64_t I, int64_t H);
bool lstm_us_tc_fwd_preferred(int64_t B, int64_t I, int64_t H);
int64_t lstm_us_tc_fwd_workspace(int64_t B, int64_t I, int64_t H);
int lstm_us_tc_fwd(const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H, const float* w_ih, const float* w_hh,
                   const float* b_ih, const float* b_hh, float* h_seq, float* c_seq, float* gates, float* h_last, void* ws,
                   int64_t ws_bytes, int single, cudaStream_t s);

// packed weights + (if no tape is requested) ping-pong buffers for h and c
int64_t lstm_rm_tc_fwd_workspace(int64_t B, int64_t I, int64_t H) {
  int KX, K;
  rm_fwd_shape(I, H, KX, K);
  const int64_t own = sizeof(float) * ((int64_t)(H / 32) * K * kChunkN * 2 + 4 * round_up(B, (int64_t)32) * H) + 256;   // weights + h_own + c ping-pong
  return lstm_us_tc_fwd_supported(I, H) ? std::max(own, lstm_us_tc_fwd_workspace(B, I, H)) : own;
}

int lstm_rm_tc_fwd(const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H, const float* w_ih, const float* w_hh,
                   const float* b_ih, const float* b_hh, float* h_seq, float* c_seq, float* gates, float* h_last, void* ws,
                   int64_t ws_bytes, int single, cudaStream_t s) {
  if (lstm_us_tc_fwd_preferred(B, I, H))
    return lstm_us_tc_fwd(xT, B, T, I, IP, H, w_ih, w_hh, b_ih, b_hh, h_seq, c_seq, gates, h_last, ws, ws_bytes, single, s);
  const int64_t need = lstm_rm_tc_fwd_workspace(B, I, H);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_fwd: workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  FwdArgs a{};
  rm_fwd_shape(I, H, a.KX, a.K);
  float* wpk = reinterpret_cast<float*>(ws);
  const int64_t wfl = (int64_t)(H / 32) * a.K * kChunkN * 2;
  float* scratch = wpk + round_up(wfl, (int64_t)64);
  // Mean truncation bias of the hi*hi accumulator (the tensor core's fp32 accumulation rounds toward zero; the K/8
  // MMAs of a chunk are chained).  tools/dbg_zbias.py without compensation (N(0,1) inputs, default-initialised
  // weights): -1.22e-6 (H = 256, I = 31), -1.16e-6 (H = 256, I = 41), -5.3e-7 (H = 128, I = 31); std 5.8e-7 in all
  // cases, as for the CUDA-core kernel.  The mean must be removed to ~1e-7: on a 701-step sequence (H = 128) the
  // parameter-gradient error is 3.6e-4 uncompensated, 1.0e-4 with 5e-7, 2.5e-4 with 7e-7.  It depends on how the
  // partial sums grow, i.e. mildly on the data; HY2DL_RM_COMP overrides the constant.
  static const float comp = [] { const char* e = getenv("HY2DL_RM_COMP"); return e ? (float)atof(e) : 0.f; }();
  // chains of kSplitJ = 4 accumulating hi*hi MMAs leave a mean truncation loss of ~1.3e-7 of the sum (0.25 * 2^-23 * 4),
  // removed here; what remains is data dependent and an order of magnitude below fp32 rounding.
  const float factor = 1.f + (comp != 0.f ? comp : 1.3e-7f);
  // CTA pairs (cta_group::2, M = 256): HY2DL_RM_PAIR=1.  Built, parity-green (the whole -m gpu suite passes with it) and NOT
  // the default: at H = 256, B = 18 944, T = 365 it runs 36.4 ms against 36.8 ms single-CTA (H = 128: 12.3 / 11.8), i.e. halving
  // the weight stream per SM buys nothing -- what bounds this kernel is the row threads' chain (accumulator pulls through the
  // 64 B/clk TMEM read port, activations, 0.9 MB of tape stores per step and CTA), see the HY2DL_RM_DBG ablation in DESIGN.md.
  // With twice the ring slots a pair can run longer splits (HY2DL_RM_PAIR_SPLITJ=12: 28.7 ms) at the accuracy of chains of 12.
  static const bool pair = [] { const char* e = getenv("HY2DL_RM_PAIR"); return e && e[0] == '1'; }();
  rm_tc_pack_w_kernel<<<256, 256, 0, s>>>(w_ih, w_hh, b_ih, b_hh, (int)H, (int)I, a.KX, a.K, factor, pair ? 1 : 0, wpk);
  a.x = xT; a.wpk = wpk; a.gates = gates; a.h_last = h_last;
  a.B = B; a.T = T; a.H = (int)H; a.I = (int)I; a.IP = (int)IP;
  a.h_seq = h_seq; a.h_own = scratch;
  a.c_seq = c_seq ? c_seq : scratch + 2 * round_up(B, (int64_t)32) * H; a.c_pingpong = c_seq ? 0 : 1;
  a.stages = rm_fwd_stages(a.KX, a.K);
  static const int stages_env = [] { const char* e = getenv("HY2DL_RM_STAGES"); return e ? atoi(e) : 0; }();   // experiment: shallower ring
  if (stages_env >= 3 && stages_env < a.stages) a.stages = stages_env;
  // K steps per split: its slices stay in the ring until the hi*hi pass, so it must leave the producer a free slot.
  // Measured at H = 256, I = 31, B = 18 944, T = 365 (8 stages), forward time / parameter-gradient error vs fp64 on the
  // reference golden of BASELINE configs[2]:  2: 44.9 ms / 2.3e-5,  4: 38.4 / 9.8e-6,  6: 35.5 / 2.2e-5,  7: 35.2 / 3.1e-5;
  // the old kernel (one chain of 36, separate correction accumulator) 27.3 ms / 1.2e-4, cut into 2 / 3 chains with the
  // accumulators pulled in between 31.8 ms / 4.9e-5 and 34.6 ms / 2.6e-5.  The cost is TMEM read bandwidth: every split
  // is one more 64 KB pull of its accumulator region into registers.  (A variant with separate hi and lo rings of 4 KB
  // slots, so that only the hi halves stay resident, measured the same: the L2 stream is not what binds here.)
  a.splitj = a.stages >= 8 ? 6 : a.stages >= 6 ? kSplitJ : 2;
  static const int sj_env = [] { const char* e = getenv("HY2DL_RM_SPLITJ"); return e ? atoi(e) : 0; }();
  if (sj_env > 0 && sj_env < a.stages) a.splitj = sj_env;
  a.single = single;
  if (single) a.splitj = a.K / 8;              // one chain, one accumulator pull per chunk
  static const int act_sep = [] { const char* e = getenv("HY2DL_RM_ACT"); return e ? atoi(e) : 0; }();
  a.act_sep = act_sep;
  static const int dbg = [] { const char* e = getenv("HY2DL_RM_DBG"); return e ? atoi(e) : 0; }();
  a.dbg = dbg;
  const size_t smem = (size_t)(a.K + a.KX) * kRowsF * 4 + (size_t)a.stages * kSliceBytes + 512 + 1024;
  const unsigned tiles = (unsigned)cdiv(B, (int64_t)kRowsF);
  if (pair) {
    static const int psj_env = [] { const char* e = getenv("HY2DL_RM_PAIR_SPLITJ"); return e ? atoi(e) : 0; }();
    if (!single && psj_env > 0 && psj_env <= 2 * a.stages - 2) a.splitj = psj_env;    // twice the slots would allow longer splits
    cudaFuncSetAttribute(lstm_rm_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((tiles + 1) / 2 * 2); cfg.blockDim = dim3(kThreadsF); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, lstm_rm_tc_fwd_kernel<true>, a);
  } else {
    cudaFuncSetAttribute(lstm_rm_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#ifdef HY2DL_US_STAMPS
    cudaMalloc(&a.stamps, 128 * sizeof(long long)); cudaMemset(a.stamps, 0, 128 * sizeof(long long));
#endif
    lstm_rm_tc_fwd_kernel<false><<<tiles, kThreadsF, smem, s>>>(a);
#ifdef HY2DL_US_STAMPS
    {
      long long h[128];
      cudaDeviceSynchronize();
      cudaMemcpy(h, a.stamps, sizeof(h), cudaMemcpyDeviceToHost);
      cudaFree(a.stamps);
      for (int i = 0; i < 128; ++i) if (h[i]) fprintf(stderr, "RF_STAMP step %d id %2d  %8lld\n", 100 + i / 64, i % 64, h[i] - h[0]);
    }
#endif
  }
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

}  // namespace tc
}  // namespace hy2dl
