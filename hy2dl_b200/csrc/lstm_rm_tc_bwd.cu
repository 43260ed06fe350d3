// Back-propagation through time on the tensor cores for H = 128 / 256 (row-major tapes of the fp32 path); the
// counterpart of lstm_rm_tc_fwd.cu.
//
// Per reverse step and 128-sequence tile:  row threads form the gate gradients dG[row][g*H + u] from the tape
// (gates, c_t, c_{t-1}) and dh = dh_ext + dh_rec, write dG over the gate tape (for the weight-gradient GEMM) and,
// split hi / lo, into shared memory as the A operand of
//     dh_rec[128 x H] = dG[128 x 4H] * W_hh[4H x H]        (3xTF32, M128 N=H K8 MMAs, A and B from shared memory).
// What does not fit as in the H = 64 kernel, and what is done instead:
//   * dh_rec needs an accumulator for the hi*hi terms and one for the correction terms (see lstm_rm_tc_fwd.cu), H
//     columns each: at H = 256 that is all of TMEM, so both halves of the A operand come from shared memory;
//   * the contraction (K = 4H) is cut into chunks of 16 hidden units x 4 gates = 64 K values (8 MMA K steps); the A
//     image of a chunk (hi | lo, 64 KB) is double buffered, so the row threads produce chunk c+1 while the tensor
//     core multiplies chunk c;
//   * W_hh (hi | lo, 2 MB at H = 256) is streamed every step from L2 through a ring of 16 KB bulk copies (one K step);
//   * the hi*hi accumulator is only trusted for 4 chunks (32 chained MMAs: the accumulation truncates, and a bias in
//     dh_rec compounds over the recurrence): after every 4 chunks the row threads add main + corrections into the
//     dh_rec buffer in global scratch (round to nearest; 3 x [B][H] of scratch in all, which stays L2 resident --
//     one buffer per group made 78 MB and went to DRAM) and the next group starts by overwriting the accumulators;
//     the group's mean bias is compensated in the packed weights;
//   * dc and dh live in global scratch ([B][H] each, two time slots for dh), not in shared memory or registers.
// Timing experiments on the row-major-tape version (H = 256, B = 256, 52.1 ms): hi*hi MMAs only 50.7 ms, no accumulator
// pull 46.0 ms, no tape traffic 18.1 ms -- the row-strided sector accesses were the cost, hence the RB8 tapes.
// A row is owned by two threads (TMEM lane = row; the two warps of a lane quarter take 8 | 8 of a chunk's 16 units);
// a thread only re-reads scratch it wrote itself.  10 warps: 8 row warps, the MMA issuer, the bulk-copy producer.
//
// Semantics: autograd of torch.nn.LSTM, h0 = c0 = 0; dG may alias gates (each element is read before it is written).
// gates / dG, c_seq and the scratch buffers are in the RB8 layout (tc_common.cuh); dh_seq and dh_last are row-major.
#include "common.cuh"
#include "tc_common.cuh"
#include <stdlib.h>

namespace hy2dl {
namespace tc {

namespace {
constexpr int kRowsB = 128;
constexpr int kThreadsB = 320;
constexpr int kUnitsC = 16;                 // hidden units per chunk
constexpr int kKC = 4 * kUnitsC;            // K values per chunk (gate-major: [g][16 units])
constexpr int kABufBytes = kKC * kRowsB * 4 * 2;   // hi | lo image of a chunk: 64 KB
constexpr int kGroup = 4;                   // chunks per accumulator group (32 chained hi*hi MMAs)

struct BwdArgs {
  const float* wpk;     // packed W_hh: [H/16 chunks][8 K steps][hi [2][H][4] | lo [2][H][4]]
  const float* gates;   // [T][B][4H] post-activation i f g o, column g*H + u
  const float* c_seq;   // [T][B][H]
  const float* dh_seq;  // [T][B][H] or null
  const float* dh_last; // [B][H] or null
  float* dG;            // [T][B][4H], may alias gates
  float* dc;            // scratch [B][H], zero-initialised
  float* dh;            // scratch RB8 [2 time slots][B][H], zero-initialised: dh_rec, accumulated group by group
  int64_t B, T, dh_t0;
  int H, stages;
  int single;           // HY2DL_PREC_TF32: hi*hi MMAs only, only the hi half of W_hh is streamed
};

// packed element (chunk c, K step j = 2 g + half, kc, k_out n, kk): W_hh[g*H + 16 c + 8 half + 4 kc + kk][n]
__global__ void rm_tc_pack_whh_kernel(const float* __restrict__ w_hh, int H, float comp, int pair, float* __restrict__ out) {
  const int step_floats = 2 * H * 4 * 2;                      // hi | lo of one K step
  const int64_t total = (int64_t)4 * H * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(e & 3);
    const int n = (int)((e >> 2) % H);
    const int64_t r2 = (e >> 2) / H;
    const int kc = (int)(r2 & 1), j = (int)((r2 >> 1) & 7), c = (int)(r2 >> 4);
    const int g = j >> 1, half = j & 1;
    const int row = g * H + kUnitsC * c + 8 * half + 4 * kc + kk;
    uint32_t hi, lo;
    split_tf32_rr(w_hh[(int64_t)row * H + n] * comp, hi, lo);
    if (pair) {    // CTA-pair layout: per K step [rank 0: hi [2][H/2][4] | lo] [rank 1: hi | lo]; rank r holds output columns r H/2 ..
      const int hn = H / 2;
      float* dst = out + ((int64_t)c * 8 + j) * step_floats + (n / hn) * (step_floats / 2) + (kc * hn + (n % hn)) * 4 + kk;
      dst[0] = __uint_as_float(hi);
      dst[step_floats / 4] = __uint_as_float(lo);
    } else {
      float* dst = out + ((int64_t)c * 8 + j) * step_floats + (kc * H + n) * 4 + kk;
      dst[0] = __uint_as_float(hi);
      dst[step_floats / 2] = __uint_as_float(lo);
    }
  }
}
}  // namespace

// PAIR: clusters of two CTAs (cta_group::2): the leader issues every MMA with M = 256 (128 sequences of each CTA), each CTA
// streams only its half of every W_hh slice (H/2 output columns); see lstm_rm_tc_fwd.cu.
template <bool PAIR>
__global__ void __launch_bounds__(kThreadsB, 1) lstm_rm_tc_bwd_kernel(BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem_raw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int H = a.H, NCH = H / kUnitsC;
  const int NS = PAIR ? 2 * a.stages : a.stages;              // ring slots (<= 16)
  const int full_step_bytes = 2 * H * 4 * 4 * 2;              // one K step of packed W_hh, both CTAs' columns: hi | lo (16 KB at H = 256)
  const int step_bytes = PAIR ? full_step_bytes / 2 : full_step_bytes;   // this CTA's part = one ring slot
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  uint8_t* s_a = smem_raw;                                    // [2 buffers][hi 32 KB | lo 32 KB], image [16 kc][128 rows][4]
  uint8_t* s_ring = s_a + 2 * kABufBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)a.stages * full_step_bytes);
  uint64_t* bar_full = bars;                       // [NS <= 16] producer (+ the peer's relay) -> MMA
  uint64_t* bar_free = bars + 16;                  // [NS] MMA commit -> producer
  uint64_t* bar_afull = bars + 32;                 // [2] rows -> MMA: A image of a chunk is in place
  uint64_t* bar_afree = bars + 34;                 // [2] MMA commit -> rows: the chunk's MMAs retired
  uint64_t* bar_group = bars + 36;                 // MMA commit: a group's MMAs retired, accumulators complete
  uint64_t* bar_pulled = bars + 37;                // rows -> MMA: accumulators have been read
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 38);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    const uint32_t rows_cnt = PAIR ? 256 + 8 : 256;           // + one arrival per row warp of the peer
    for (int s = 0; s < NS; ++s) { mbar_init(bar_full + s, (PAIR && rank == 0) ? 2 : 1); mbar_init(bar_free + s, 1); }
    for (int p = 0; p < 2; ++p) { mbar_init(bar_afull + p, rows_cnt); mbar_init(bar_afree + p, 1); }
    mbar_init(bar_group, 1);
    mbar_init(bar_pulled, rows_cnt);
    mbar_fence_init();
  }
  if (warp == 8) { if (PAIR) tmem_alloc2(s_tmem, 512); else tmem_alloc(s_tmem, 512); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  auto rows_arrive = [&](uint64_t* bar) {            // a row thread's arrival on a barrier the MMA issuer waits for
    if (!PAIR || rank == 0) { mbar_arrive(bar); return; }
    __syncwarp();
    if (lane == 0) mbar_arrive_remote(mapa_rank(smem_u32(bar), 0));
  };
  const uint32_t col_main = 0, col_corr = (uint32_t)H;        // H = 256: [0,256) | [256,512)

  if (warp < 8) {
    // =========================== row threads ===================================================
    const int q = warp & 3, hh = warp >> 2;
    const int rloc = q * 32 + lane;
    const int64_t b = (int64_t)blockIdx.x * kRowsB + rloc;
    const bool live = b < a.B;
    const uint32_t lane_tm = tmem + ((uint32_t)(q * 32) << 16);
    const int64_t BH = a.B * H;
    const int64_t G = (a.B + 31) >> 5, GH = G * 32 * H;      // RB8: rows padded to whole groups
    const int HB = H / 8;
    float* dcp = a.dc + rb8_off(0, G, b, H, hh);               // this thread's unit blocks: hh, hh + 2, ... (chunk c: 2 c + hh)
    const int n_groups = NCH / kGroup;
    struct Tape { f8 gi, gf, gg, go, ct, cp, dh, dc; };
    int64_t gch = 0;                               // chunk counter over the whole kernel
    int64_t ggr = 0;                               // group counter
    for (int64_t s = 0; s < a.T; ++s) {
      const int64_t t = a.T - 1 - s;
      // dh_rec(t) = sum of the group partials written during the previous step; dh_rec(t-1) is written group by group
      const float* dh_rd = a.dh + (s & 1) * GH + rb8_off(0, G, b, H, hh);
      float* dh_wr = a.dh + ((s + 1) & 1) * GH + rb8_off(0, G, b, H, hh);
      const float* gp = a.gates + rb8_off(t, G, b, 4 * H, hh);
      float* dgp = a.dG + rb8_off(t, G, b, 4 * H, hh);
      const float* ctp = a.c_seq + rb8_off(t, G, b, H, hh);
      const float* cpp = t > 0 ? a.c_seq + rb8_off(t - 1, G, b, H, hh) : nullptr;
      const float* dhe = (a.dh_seq && t >= a.dh_t0) ? a.dh_seq + (t * a.B + b) * H + 8 * hh : nullptr;   // row-major
      const float* dhl = (a.dh_last && t == a.T - 1) ? a.dh_last + b * H + 8 * hh : nullptr;
      auto load = [&](int c, Tape& tp) {           // everything chunk c needs; issued one chunk ahead
        const int o = kUnitsC * c;                 // row-major offset of the chunk (units o + 8 hh ..)
        const int ob = 2 * c * 256;                // RB8 offset of unit block 2 c (+ hh in the base pointers)
        tp.gi = tp.gf = tp.gg = tp.go = tp.ct = tp.cp = tp.dh = tp.dc = f8_zero();
        if (!live) return;
        tp.gi = ld256(gp + ob); tp.gf = ld256(gp + HB * 256 + ob); tp.gg = ld256(gp + 2 * HB * 256 + ob); tp.go = ld256(gp + 3 * HB * 256 + ob);
        tp.ct = ld256(ctp + ob);
        if (cpp) tp.cp = ld256(cpp + ob);
        tp.dc = ld256(dcp + ob);
        tp.dh = ld256(dh_rd + ob);
        if (dhe) { const f8 e = ld256(dhe + o); for (int k = 0; k < 8; ++k) tp.dh.v[k] += e.v[k]; }
        if (dhl) { const f8 e = ld256(dhl + o); for (int k = 0; k < 8; ++k) tp.dh.v[k] += e.v[k]; }
      };
      auto produce = [&](int c, const Tape& tp) {  // gate gradients of chunk c -> tape, dc scratch and the A image
        const int buf = (int)(gch & 1);
        const int ob = 2 * c * 256;
        f8 di, df, dg, dO, dcn;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float tcv = fast_tanh(tp.ct.v[k]);
          const float dct = tp.dc.v[k] + tp.dh.v[k] * tp.go.v[k] * (1.f - tcv * tcv);
          dO.v[k] = tp.dh.v[k] * tcv * tp.go.v[k] * (1.f - tp.go.v[k]);
          di.v[k] = dct * tp.gg.v[k] * tp.gi.v[k] * (1.f - tp.gi.v[k]);
          df.v[k] = dct * tp.cp.v[k] * tp.gf.v[k] * (1.f - tp.gf.v[k]);
          dg.v[k] = dct * tp.gi.v[k] * (1.f - tp.gg.v[k] * tp.gg.v[k]);
          dcn.v[k] = dct * tp.gf.v[k];
        }
        if (live) {
          stg256(dgp + ob, di); stg256(dgp + HB * 256 + ob, df); stg256(dgp + 2 * HB * 256 + ob, dg); stg256(dgp + 3 * HB * 256 + ob, dO);
          st256(dcp + ob, dcn);
        }
        // A image of the chunk: K index = 16 g + (8 hh + k); this thread writes K step 2 g + hh of every gate
        if (gch >= 2) mbar_wait(bar_afree + buf, (uint32_t)(((gch - 2) >> 1) & 1));
        float4* img_hi = reinterpret_cast<float4*>(s_a + buf * kABufBytes);
        float4* img_lo = img_hi + kABufBytes / 32;
        const f8* src[4] = {&di, &df, &dg, &dO};
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) split_tf32(src[g]->v[k], hi[k], lo[k]);
          const int kc0 = (kUnitsC * g + 8 * hh) >> 2;
          img_hi[kc0 * kRowsB + rloc] = make_float4(__uint_as_float(hi[0]), __uint_as_float(hi[1]), __uint_as_float(hi[2]), __uint_as_float(hi[3]));
          img_hi[(kc0 + 1) * kRowsB + rloc] = make_float4(__uint_as_float(hi[4]), __uint_as_float(hi[5]), __uint_as_float(hi[6]), __uint_as_float(hi[7]));
          img_lo[kc0 * kRowsB + rloc] = make_float4(__uint_as_float(lo[0]), __uint_as_float(lo[1]), __uint_as_float(lo[2]), __uint_as_float(lo[3]));
          img_lo[(kc0 + 1) * kRowsB + rloc] = make_float4(__uint_as_float(lo[4]), __uint_as_float(lo[5]), __uint_as_float(lo[6]), __uint_as_float(lo[7]));
        }
        fence_async_smem();
        rows_arrive(bar_afull + buf);
        ++gch;
      };
      auto pull = [&](int g) {                     // group g of this step: main + corrections -> its partial of dh_rec(t-1)
        mbar_wait(bar_group, (uint32_t)(ggr & 1));
        tc_fence_after();
        float* dst = dh_wr;                        // one buffer per time slot: 3 x [B][H] of scratch stay L2 resident
#pragma unroll 4
        for (int u0 = 8 * hh; u0 < H; u0 += 16) {  // this thread's columns: units with (u >> 3) & 1 == hh
          uint32_t zm[8], zc[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(zm[0]), "=r"(zm[1]), "=r"(zm[2]), "=r"(zm[3]), "=r"(zm[4]), "=r"(zm[5]), "=r"(zm[6]), "=r"(zm[7])
                       : "r"(lane_tm + col_main + u0));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(zc[0]), "=r"(zc[1]), "=r"(zc[2]), "=r"(zc[3]), "=r"(zc[4]), "=r"(zc[5]), "=r"(zc[6]), "=r"(zc[7])
                       : "r"(lane_tm + col_corr + u0));
          tmem_ld_wait();
          f8 acc;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc.v[k] = a.single ? __uint_as_float(zm[k]) : __uint_as_float(zm[k]) + __uint_as_float(zc[k]);
          if (live) {                                        // unit block u0 / 8 = 2 (u0 / 16) + hh
            float* p = dst + (u0 >> 4) * 512;
            if (g > 0) { const f8 old = ld256(p); for (int k = 0; k < 8; ++k) acc.v[k] += old.v[k]; }
            st256(p, acc);
          }
        }
        tc_fence_before();
        rows_arrive(bar_pulled);
        ++ggr;
      };
      Tape cur, nxt;
      load(0, cur);
      for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) load(c + 1, nxt);
        produce(c, cur);
        // a group's accumulators are read after the first image of the next group has been queued, so that the
        // tensor core resumes the moment they are released
        if (c > 0 && (c % kGroup) == 0) pull(c / kGroup - 1);
        cur = nxt;
      }
      pull(n_groups - 1);
    }
  } else if (warp == 8) {
    // =========================== MMA issuer (the leader's, in a pair) ===========================
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = idesc_tf32(PAIR ? 2 * kRowsB : kRowsB, H, 0, 0);
      const uint32_t a_lbo = kRowsB * 16, b_lbo = (uint32_t)(PAIR ? H / 2 : H) * 16, sbo = 128;
      const uint32_t d_main = tmem + col_main, d_corr = tmem + col_corr;
      const uint64_t ring0 = smem_desc(smem_u32(s_ring), b_lbo, sbo);
      const uint64_t a0[2] = {smem_desc(smem_u32(s_a), a_lbo, sbo), smem_desc(smem_u32(s_a + kABufBytes), a_lbo, sbo)};
      const uint64_t kStageStep = (uint64_t)step_bytes / 16, kBLo = (uint64_t)step_bytes / 32;
      constexpr uint64_t kALo = kABufBytes / 32, kAStep = 2 * kRowsB * 16 / 16;
      auto mma_ss = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) { if (PAIR) mma2_tf32_ss(d, ad, bd, idesc, acc); else mma_tf32_ss(d, ad, bd, idesc, acc); };
      auto commit = [&](uint64_t* bar) { if (PAIR) mma2_commit(bar); else mma_commit(bar); };
      int st = 0;
      uint32_t ph = 0;
      uint64_t dbh = ring0;
      int64_t gch = 0, ggr = 0;
      for (int64_t s = 0; s < a.T; ++s) {
        for (int c = 0; c < NCH; ++c, ++gch) {
          const int buf = (int)(gch & 1);
          const bool open = (c % kGroup) == 0;     // first chunk of a group overwrites the accumulators
          if (open && ggr > 0) {
            mbar_wait(bar_pulled, (uint32_t)((ggr - 1) & 1));
            tc_fence_after();
          }
          mbar_wait(bar_afull + buf, (uint32_t)((gch >> 1) & 1));
          tc_fence_after();
          uint64_t dah = a0[buf], dal = a0[buf] + kALo;
#pragma unroll 1
          for (int j = 0; j < 8; ++j) {
            mbar_wait(bar_full + st, ph);
            tc_fence_after();
            const uint64_t bl = dbh + kBLo;
            const uint32_t acc = !(open && j == 0);
            if (!a.single) {
              mma_ss(d_corr, dal, dbh, acc);
              mma_ss(d_corr, dah, bl, 1);
            }
            mma_ss(d_main, dah, dbh, acc);
            dah += kAStep; dal += kAStep;
            commit(bar_free + st);
            dbh += kStageStep;
            if (++st == NS) { st = 0; ph ^= 1; dbh = ring0; }
          }
          commit(bar_afree + buf);
          if ((c % kGroup) == kGroup - 1) { commit(bar_group); ++ggr; }
        }
      }
    } else if (PAIR && rank == 1 && lane == 0) {
      // the peer's relay: tells the leader that this CTA's half of a W_hh slice has landed
      int st = 0;
      uint32_t ph = 0;
      const int64_t n_slices = a.T * NCH * 8;
      for (int64_t i = 0; i < n_slices; ++i) {
        mbar_wait(bar_full + st, ph);
        mbar_arrive_remote(mapa_rank(smem_u32(bar_full + st), 0));
        if (++st == NS) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =========================== W_hh slice producer ============================================
    if (lane == 0) {
      const int per_step = NCH * 8;
      int st = 0;
      uint32_t ph = 1;
      bool first_round = true;
      const uint32_t nbytes = a.single ? (uint32_t)step_bytes / 2 : (uint32_t)step_bytes;    // hi | lo
      for (int64_t s = 0; s < a.T; ++s) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wpk) + (PAIR ? rank * step_bytes : 0);
        for (int i = 0; i < per_step; ++i, src += full_step_bytes) {
          if (!first_round) mbar_wait(bar_free + st, ph);
          mbar_arrive_expect_tx(bar_full + st, nbytes);
          bulk_g2s(s_ring + (size_t)st * step_bytes, src, nbytes, bar_full + st);
          if (++st == NS) { st = 0; ph ^= 1; first_round = false; }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) {
    cluster_sync();
    if (warp == 8) tmem_dealloc2(tmem, 512);
  } else if (warp == 8) {
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
static int rm_bwd_stages(int64_t H) {
  const int64_t step_bytes = 2 * H * 4 * 4 * 2;
  return (int)std::min<int64_t>(8, (232448 - 2 * kABufBytes - 512 - 1024) / step_bytes);
}

bool lstm_rm_tc_bwd_supported(int64_t H) {
  return H == 128 || H == 256;
}

// the unit-split kernel for small batches (lstm_us_tc_bwd.cu)
bool lstm_us_tc_bwd_supported(int64_t H);
bool lstm_us_tc_bwd_preferred(int64_t B, int64_t H);
int64_t lstm_us_tc_bwd_workspace(int64_t B, int64_t H);
int lstm_us_tc_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates, const float* c_seq, const float* dh_seq,
                   int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, int single, cudaStream_t s);

// packed W_hh (hi | lo) + dc [B][H] + dh [2][B][H]
int64_t lstm_rm_tc_bwd_workspace(int64_t B, int64_t H) {
  const int64_t own = sizeof(float) * (8 * H * H + 3 * round_up(B, (int64_t)32) * H) + 256;
  return lstm_us_tc_bwd_preferred(B, H) ? std::max(own, lstm_us_tc_bwd_workspace(B, H)) : own;
}

int lstm_rm_tc_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates, const float* c_seq, const float* dh_seq,
                   int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, int single, cudaStream_t s) {
  if (lstm_us_tc_bwd_preferred(B, H))
    return lstm_us_tc_bwd(B, T, H, w_hh, gates, c_seq, dh_seq, dh_t0, dh_last, dG, ws, ws_bytes, single, s);
  const int64_t need = lstm_rm_tc_bwd_workspace(B, H);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_bwd: workspace %lld < %lld bytes (H = %lld needs the "
             "packed W_hh and the dc / dh scratch)", (long long)ws_bytes, (long long)need, (long long)H);
  BwdArgs a{};
  float* wpk = reinterpret_cast<float*>(ws);
  float* scratch = wpk + 8 * H * H;
  // mean truncation bias of a group's hi*hi accumulator (kGroup chunks x 8 = 32 chained MMAs).  The gate gradients
  // have random signs, so the partial sums grow like a random walk and the bias is smaller than in the forward
  // kernel: parameter-gradient error vs an fp64 evaluation (H = 256, T = 365) 4.1e-5 / 4.0e-5 / 4.3e-5 / 4.6e-5 for a
  // compensation of 0 / 5e-7 / 1.1e-6 / 1.7e-6 (CUDA-core chain: 2.0e-5).  HY2DL_RM_BCOMP overrides.
  static const float comp = [] { const char* e = getenv("HY2DL_RM_BCOMP"); return e ? (float)atof(e) : 5e-7f; }();
  // CTA pairs (HY2DL_RM_PAIR=1): parity-green, not faster (30.9 ms against 30.6 ms at H = 256; 10.4 / 9.7 at H = 128), see lstm_rm_tc_fwd.cu
  static const bool pair = [] { const char* e = getenv("HY2DL_RM_PAIR"); return e && e[0] == '1'; }();
  rm_tc_pack_whh_kernel<<<256, 256, 0, s>>>(w_hh, (int)H, 1.f + comp, pair ? 1 : 0, wpk);
  const int64_t Bp = round_up(B, (int64_t)32);
  cudaMemsetAsync(scratch, 0, sizeof(float) * 3 * Bp * H, s);
  a.wpk = wpk; a.gates = gates; a.c_seq = c_seq; a.dh_seq = dh_seq; a.dh_last = dh_last; a.dG = dG;
  a.dc = scratch; a.dh = scratch + Bp * H;
  a.B = B; a.T = T; a.dh_t0 = dh_t0; a.H = (int)H; a.single = single;
  a.stages = rm_bwd_stages(H);
  const size_t smem = (size_t)2 * kABufBytes + (size_t)a.stages * (2 * H * 4 * 4 * 2) + 512 + 1024;
  const unsigned tiles = (unsigned)cdiv(B, (int64_t)kRowsB);
  if (pair) {
    cudaFuncSetAttribute(lstm_rm_tc_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((tiles + 1) / 2 * 2); cfg.blockDim = dim3(kThreadsB); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, lstm_rm_tc_bwd_kernel<true>, a);
  } else {
    cudaFuncSetAttribute(lstm_rm_tc_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lstm_rm_tc_bwd_kernel<false><<<tiles, kThreadsB, smem, s>>>(a);
  }
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

}  // namespace tc
}  // namespace hy2dl
