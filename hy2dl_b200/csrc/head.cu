// Head: Linear(H, P*m + R) evaluated only where the conceptual model needs it, fused with the
// sigmoid range mapping (modelzoo/hybrid.py:93-97,110-113; baseconceptualmodel.py:53-76).
//
//   dynamic parameter columns : steps t = warmup-1 (-> warm-up value) and t >= warmup (-> simulation)
//   static columns + routing   : step t = T-1 only
//
// The reference computes the Linear on all T steps and all columns and then slices; the values it
// actually uses are exactly the ones above, so skipping the rest changes no result.
// Output goes straight into the time-major planes the conceptual kernels read.
#include "common.cuh"

namespace hy2dl {

constexpr int kRows = 32;     // rows (samples of one step) per tile
constexpr int kChunk = 32;    // head columns per launch
constexpr int kThreads = 256;
constexpr int kDz = kChunk + 4;   // row stride of the dz tile in shared memory

// RI32 slab of one (step, 32-row group): float e of [F/32][32 rows][32 floats] holds (row, feature)
__device__ __forceinline__ int ri32_row(int e) { return (e >> 5) & 31; }
__device__ __forceinline__ int ri32_feat(int e) { return (e >> 10) * 32 + ((((e >> 3) ^ (e >> 5)) & 3) << 3) + (e & 7); }

struct HeadChunk {
  // one launch handles <= kChunk head columns of the same kind (all dynamic or all static)
  int32_t ncol;
  int32_t col[kChunk];      // column index into w / bias
  int32_t par[kChunk];      // parameter index (>= n_par: routing)
  int32_t mem[kChunk];      // member index j
  float lo[kChunk], hi[kChunk];
  int64_t off[kChunk];      // plane offset in par_sim
  int32_t is_dynamic;
};

struct HeadArgs {
  const float* h_seq;   // [T][B][H]
  const float* w;       // [C][H]
  const float* bias;    // [C]
  float* par_sim;       // fwd out / bwd: values (const)
  float* par_warm;      // fwd out
  const float* dpar_sim;
  float* dh_seq;
  float* dw;
  float* dbias;
  float* part;          // bwd: per-block partial sums [grid][kChunk][H] | [grid][kChunk] (workspace)
  int64_t B, T, H;
  int32_t m, n_par, warmup;
  int32_t accumulate_dh;
  int32_t ri32;   // h_seq / dh_seq in RI32 [T][G][H/4][32][4] instead of [T][B][H]
};

// ---- forward ------------------------------------------------------------------------------------
// Tile = one time slot x 32 rows.  256 threads: lane = row, warp = group of 4 head columns, so a thread
// accumulates 4 dot products over H with one conflict-free LDS (h, row stride H+1) and one broadcast
// LDS.128 (W^T[k][4 columns]) per k.  smem: hs[32][H+1] | wT[H][kChunk]
__global__ void __launch_bounds__(kThreads) head_fwd_kernel(HeadArgs a, HeadChunk ch) {
  extern __shared__ __align__(16) float sm[];
  const int H = (int)a.H;
  float* wT = sm;                       // [H][kChunk]
  float* hs = wT + H * kChunk;          // [kRows][H+1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t N = a.B * a.m;
  for (int e = tid; e < H * kChunk; e += kThreads) {
    const int k = e / kChunk, c = e % kChunk;
    wT[e] = c < ch.ncol ? __ldg(a.w + (int64_t)ch.col[c] * H + k) : 0.f;
  }
  // time slots: dynamic: warmup-1 .. T-1 (T_sim+1 slots); static: T-1
  const int64_t t_first = ch.is_dynamic ? a.warmup - 1 : a.T - 1;
  const int64_t n_slots = a.T - t_first;
  const int64_t tiles_b = cdiv(a.B, kRows);
  const int64_t n_tiles = n_slots * tiles_b;
  const int c0 = 4 * warp;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t t = t_first + tile / tiles_b;
    const int64_t b0 = (tile % tiles_b) * kRows;
    const int nb = (int)min((int64_t)kRows, a.B - b0);
    __syncthreads();
    if (a.ri32) {   // tile = one 32-row group, contiguous as [H/32][32 rows][32 floats], units swizzled
      const float* src = a.h_seq + (t * tiles_b + b0 / kRows) * kRows * H;
      for (int e = tid; e < kRows * H; e += kThreads) hs[ri32_row(e) * (H + 1) + ri32_feat(e)] = __ldg(src + e);
    } else {        // the [nb][H] slab of step t is contiguous in h_seq
      const float* src = a.h_seq + (t * a.B + b0) * H;
      for (int e = tid; e < nb * H; e += kThreads) hs[(e / H) * (H + 1) + (e % H)] = __ldg(src + e);
    }
    __syncthreads();
    if (c0 < ch.ncol && lane < nb) {
      const float* hr = hs + lane * (H + 1);
      float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int k = 0; k < H; ++k) {
        const float hk = hr[k];
        const float4 w4 = *reinterpret_cast<const float4*>(wT + k * kChunk + c0);
        z[0] = fmaf(hk, w4.x, z[0]); z[1] = fmaf(hk, w4.y, z[1]); z[2] = fmaf(hk, w4.z, z[2]); z[3] = fmaf(hk, w4.w, z[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        if (c >= ch.ncol) break;
        const float val = ch.lo[c] + sigmoidf_acc(z[j] + __ldg(a.bias + ch.col[c])) * (ch.hi[c] - ch.lo[c]);
        const int64_t b = b0 + lane;
        const bool routing = ch.par[c] >= a.n_par;
        if (ch.is_dynamic) {
          if (t == a.warmup - 1) a.par_warm[(int64_t)ch.par[c] * N + b * a.m + ch.mem[c]] = val;
          else a.par_sim[ch.off[c] + (t - a.warmup) * N + b * a.m + ch.mem[c]] = val;
        } else if (routing) {
          a.par_sim[ch.off[c] + b] = val;
        } else {
          a.par_sim[ch.off[c] + b * a.m + ch.mem[c]] = val;
          a.par_warm[(int64_t)ch.par[c] * N + b * a.m + ch.mem[c]] = val;
        }
      }
    }
  }
}

// ---- backward -----------------------------------------------------------------------------------
// Tile = one time slot x 32 rows.  256 threads = 64 hidden columns x 4 row quarters: a thread keeps its
// column of W (ncol registers per 64-column block) and, for each of its 8 rows, reads the row of dz
// with broadcast LDS.128 and h[r][k] with one conflict-free LDS; it produces dh[r][k] (written as whole
// 128-byte lines) and accumulates dW[c][k] in registers across all tiles of the CTA (H = 64) or in
// shared memory (larger H).  smem: ws[kChunk][H] | hs[32][H] | dzs[32][kChunk] | dws[kChunk][H] | dbs[kChunk]
template <int NC>   // NC: head columns of the launch rounded up (8, 16, 32)
__global__ void __launch_bounds__(kThreads) head_bwd_kernel(HeadArgs a, HeadChunk ch) {
  extern __shared__ __align__(16) float sm[];
  const int H = (int)a.H;
  float* ws = sm;                          // [kChunk][H]
  float* hs = ws + kChunk * H;             // [kRows][H]
  float* dzs = hs + kRows * H;             // [kRows][kDz] (padded: conflict-free column writes, 16-byte rows)
  float* dws = dzs + kRows * kDz;          // [kChunk][H]
  float* dbs = dws + kChunk * H;           // [kChunk]
  // thread -> (hidden column k, row slice): 256 threads = 64 columns x 4; with KB = H / 64 column blocks the four
  // groups are KB column blocks x RS = 4 / KB row slices of 32 / RS rows.  Every thread keeps its column of W and
  // its NC weight-gradient sums in registers for the whole kernel -- no shared-memory atomics, fixed summation order.
  const int tid = threadIdx.x, kk = tid & 63, rq = tid >> 6;
  const int KB = H / 64, RS = 4 / KB, nrows = kRows / RS;
  const int k = (rq % KB) * 64 + kk, r0 = (rq / KB) * nrows;
  const int64_t N = a.B * a.m;
  const bool want_dh = a.dh_seq != nullptr;      // null: weight / bias gradients only

  for (int e = tid; e < kChunk * H; e += kThreads) {
    ws[e] = (e / H) < ch.ncol ? __ldg(a.w + (int64_t)ch.col[e / H] * H + (e % H)) : 0.f;
    dws[e] = 0.f;
  }
  if (tid < kChunk) dbs[tid] = 0.f;
  __syncthreads();
  float wreg[NC], acc_dw[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) { wreg[c] = ws[c * H + k]; acc_dw[c] = 0.f; }
  float acc_db = 0.f;

  const int64_t t_first = ch.is_dynamic ? a.warmup : a.T - 1;
  const int64_t n_slots = a.T - t_first;
  const int64_t tiles_b = cdiv(a.B, kRows);
  const int64_t n_tiles = n_slots * tiles_b;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t t = t_first + tile / tiles_b;
    const int64_t b0 = (tile % tiles_b) * kRows;
    const int nb = (int)min((int64_t)kRows, a.B - b0);
    __syncthreads();
    // h tile -> hs[r][k]
    if (a.ri32) {
      const float* src = a.h_seq + (t * tiles_b + b0 / kRows) * kRows * H;
      for (int e = tid; e < kRows * H; e += kThreads) hs[ri32_row(e) * H + ri32_feat(e)] = __ldg(src + e);
    } else {
      const float* src = a.h_seq + (t * a.B + b0) * H;
      for (int e = tid; e < kRows * H; e += kThreads) hs[e] = e < nb * H ? __ldg(src + e) : 0.f;
    }
    // dz[r][c] = dpar * dval/dz, with sigma recovered from the stored value (0 for rows >= nb, columns >= ncol)
    for (int e = tid; e < kChunk * kRows; e += kThreads) {
      const int c = e / kRows, r = e % kRows;
      float dz = 0.f;
      if (r < nb && c < ch.ncol) {
        const int64_t b = b0 + r;
        const bool routing = ch.par[c] >= a.n_par;
        const int64_t idx = ch.off[c] + (ch.is_dynamic ? (t - a.warmup) * N : 0) + (routing ? b : b * a.m + ch.mem[c]);
        const float val = __ldg(a.par_sim + idx);
        dz = __ldg(a.dpar_sim + idx) * (val - ch.lo[c]) * (ch.hi[c] - val) / (ch.hi[c] - ch.lo[c]);
      }
      dzs[r * kDz + c] = dz;
    }
    __syncthreads();
    if (tid < ch.ncol) {
      float s = 0.f;
      for (int r = 0; r < kRows; ++r) s += dzs[r * kDz + tid];
      acc_db += s;
    }
#pragma unroll 2
    for (int rr = 0; rr < nrows; ++rr) {
      const int r = r0 + rr;
      const float hk = hs[r * H + k];
      float dh = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < NC; c4 += 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(dzs + r * kDz + c4);
        dh = fmaf(d4.x, wreg[c4], dh); dh = fmaf(d4.y, wreg[c4 + 1], dh);
        dh = fmaf(d4.z, wreg[c4 + 2], dh); dh = fmaf(d4.w, wreg[c4 + 3], dh);
        acc_dw[c4] = fmaf(d4.x, hk, acc_dw[c4]); acc_dw[c4 + 1] = fmaf(d4.y, hk, acc_dw[c4 + 1]);
        acc_dw[c4 + 2] = fmaf(d4.z, hk, acc_dw[c4 + 2]); acc_dw[c4 + 3] = fmaf(d4.w, hk, acc_dw[c4 + 3]);
      }
      // dh[r][k]: a warp covers 32 consecutive k of one row = one 128-byte line in either layout
      int64_t o;
      if (a.ri32) o = (t * tiles_b + b0 / kRows) * kRows * H + (int64_t)(k >> 5) * 1024 + r * 32 + ((((k >> 3) ^ r) & 3) << 3) + (k & 7);
      else o = (t * a.B + b0 + r) * H + k;
      if (want_dh && (a.ri32 || r < nb)) a.dh_seq[o] = a.accumulate_dh ? a.dh_seq[o] + dh : dh;
    }
  }
  __syncthreads();
  // the RS row slices of a column add their sums in turn (fixed order)
  for (int turn = 0; turn < RS; ++turn) {
    if (rq / KB == turn) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < ch.ncol) dws[c * H + k] += acc_dw[c];
    }
    __syncthreads();
  }
  // per-block sums -> workspace; head_bwd_finish_kernel adds them in block order (no floating-point atomics)
  float* pw = a.part + (int64_t)blockIdx.x * kChunk * H;
  for (int e = tid; e < ch.ncol * H; e += kThreads) pw[e] = dws[e];
  if (tid < ch.ncol) a.part[(int64_t)gridDim.x * kChunk * H + blockIdx.x * kChunk + tid] = acc_db;
}

// dw[col[c]][k] = sum over blocks (in block order) of the partial sums; one thread per element
__global__ void __launch_bounds__(256) head_bwd_finish_kernel(const float* __restrict__ part, int nblocks, int H, HeadChunk ch,
                                                              float* __restrict__ dw, float* __restrict__ dbias) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e < ch.ncol * H) {
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += part[(int64_t)b * kChunk * H + e];
    dw[(int64_t)ch.col[e / H] * H + (e % H)] = s;
  } else if (e < ch.ncol * H + ch.ncol) {
    const int c = e - ch.ncol * H;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += part[(int64_t)nblocks * kChunk * H + b * kChunk + c];
    dbias[ch.col[c]] = s;
  }
}

constexpr int kHeadBwdMaxBlocks = 296;

static size_t head_smem(int64_t H, bool bwd) {
  size_t f = bwd ? (size_t)kChunk * H * 2 + (size_t)kRows * H + (size_t)kRows * kDz + kChunk
                 : (size_t)H * kChunk + (size_t)kRows * (H + 1);
  return f * sizeof(float);
}

// enumerate the launches: dynamic columns first, then static (+ routing)
static int build_chunks(const hy2dl_parmap* pm, HeadChunk* out, int max_chunks) {
  int n = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int want_dynamic = pass == 0;
    HeadChunk cur{};
    cur.is_dynamic = want_dynamic;
    const int total_par = pm->n_par + (want_dynamic ? 0 : pm->n_routing);
    for (int p = 0; p < total_par; ++p) {
      const bool routing = p >= pm->n_par;
      const bool dyn = !routing && pm->dynamic[p];
      if ((int)dyn != want_dynamic) continue;
      const int members = routing ? 1 : pm->n_members;
      for (int j = 0; j < members; ++j) {
        const int c = cur.ncol++;
        cur.col[c] = routing ? pm->n_par * pm->n_members + (p - pm->n_par) : p * pm->n_members + j;
        cur.par[c] = p;
        cur.mem[c] = j;
        cur.lo[c] = pm->lo[p];
        cur.hi[c] = pm->hi[p];
        cur.off[c] = pm->sim_off[p];
        if (cur.ncol == kChunk) {
          if (n == max_chunks) return -1;
          out[n++] = cur;
          cur = HeadChunk{};
          cur.is_dynamic = want_dynamic;
        }
      }
    }
    if (cur.ncol) {
      if (n == max_chunks) return -1;
      out[n++] = cur;
    }
  }
  return n;
}

static int check_parmap(const hy2dl_parmap* pm, int64_t T, const char* who) {
  HY_REQUIRE(pm != nullptr, HY2DL_ERR_ALIGN, "%s: parmap is null", who);
  HY_REQUIRE(pm->n_par > 0 && pm->n_par <= HY2DL_MAX_PAR && pm->n_members > 0 && (pm->n_routing == 0 || pm->n_routing == 2),
             HY2DL_ERR_SHAPE, "%s: bad parmap (n_par=%d m=%d n_routing=%d)", who, pm->n_par, pm->n_members, pm->n_routing);
  HY_REQUIRE(pm->warmup >= 1 && pm->warmup < T, HY2DL_ERR_SHAPE, "%s: warmup=%d must satisfy 1 <= warmup < T=%lld", who,
             pm->warmup, (long long)T);
  return HY2DL_OK;
}

}  // namespace hy2dl
using namespace hy2dl;

extern "C" int64_t hy2dl_parmap_layout(hy2dl_parmap* pm, int64_t B, int64_t T) {
  if (!pm || pm->n_par <= 0 || pm->n_par > HY2DL_MAX_PAR) return -1;
  const int64_t N = B * pm->n_members, T_sim = T - pm->warmup;
  int64_t off = 0;
  for (int p = 0; p < pm->n_par + pm->n_routing; ++p) {
    pm->sim_off[p] = off;
    const int64_t sz = p >= pm->n_par ? B : (pm->dynamic[p] ? T_sim * N : N);
    off += round_up(sz, 4);
  }
  return off;
}

extern "C" int hy2dl_head_map_fwd(const float* h_seq, int64_t B, int64_t T, int64_t H, const float* w,
                                  const float* bias, const hy2dl_parmap* pm, float* par_sim, float* par_warm,
                                  int layout, hy2dl_stream_t stream) {
  int rc = check_parmap(pm, T, "hy2dl_head_map_fwd");
  if (rc) return rc;
  HY_REQUIRE(B > 0 && H > 0 && H <= 512, HY2DL_ERR_SHAPE, "hy2dl_head_map_fwd: bad sizes B=%lld H=%lld", (long long)B, (long long)H);
  HY_NONNULL(h_seq); HY_NONNULL(w); HY_NONNULL(bias); HY_NONNULL(par_sim); HY_NONNULL(par_warm);
  HeadChunk chunks[32];
  const int n = build_chunks(pm, chunks, 32);
  HY_REQUIRE(n > 0, HY2DL_ERR_SHAPE, "hy2dl_head_map_fwd: too many head columns");
  HeadArgs a{};
  a.h_seq = h_seq; a.w = w; a.bias = bias; a.par_sim = par_sim; a.par_warm = par_warm;
  a.B = B; a.T = T; a.H = H; a.m = pm->n_members; a.n_par = pm->n_par; a.warmup = pm->warmup;
  HY_REQUIRE(layout == HY2DL_LAYOUT_ROWMAJOR || (layout == HY2DL_LAYOUT_RI32 && H % 32 == 0), HY2DL_ERR_SHAPE,
             "hy2dl_head_map_fwd: unknown layout %d (RI32 needs H %% 32 == 0)", layout);
  a.ri32 = layout == HY2DL_LAYOUT_RI32;
  const size_t smem = head_smem(H, false);
  cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int i = 0; i < n; ++i) {
    const int64_t slots = chunks[i].is_dynamic ? (T - pm->warmup + 1) : 1;
    const unsigned grid = (unsigned)std::min<int64_t>(slots * cdiv(B, kRows), 8 * (int64_t)sm_count());
    head_fwd_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(a, chunks[i]);
  }
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int64_t hy2dl_head_bwd_workspace_bytes(int64_t H) {
  return (int64_t)sizeof(float) * kHeadBwdMaxBlocks * kChunk * (H + 1);
}

extern "C" int hy2dl_head_map_bwd(const float* h_seq, const float* par_sim, const float* dpar_sim, int64_t B, int64_t T,
                                  int64_t H, const float* w, const hy2dl_parmap* pm, float* dh_seq, float* dw,
                                  float* dbias, void* ws, int64_t ws_bytes, int layout, hy2dl_stream_t stream) {
  int rc = check_parmap(pm, T, "hy2dl_head_map_bwd");
  if (rc) return rc;
  HY_REQUIRE(B > 0 && (H == 64 || H == 128 || H == 256), HY2DL_ERR_SHAPE, "hy2dl_head_map_bwd: bad sizes B=%lld H=%lld (H must be 64, 128 or 256)", (long long)B, (long long)H);
  HY_NONNULL(h_seq); HY_NONNULL(par_sim); HY_NONNULL(dpar_sim); HY_NONNULL(w); HY_NONNULL(dw); HY_NONNULL(dbias);
  HY_REQUIRE(ws != nullptr && ws_bytes >= hy2dl_head_bwd_workspace_bytes(H), HY2DL_ERR_WORKSPACE,
             "hy2dl_head_map_bwd: workspace %lld < %lld bytes", (long long)ws_bytes, (long long)hy2dl_head_bwd_workspace_bytes(H));
  HeadChunk chunks[32];
  const int n = build_chunks(pm, chunks, 32);
  HY_REQUIRE(n > 0, HY2DL_ERR_SHAPE, "hy2dl_head_map_bwd: too many head columns");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t C = (int64_t)pm->n_par * pm->n_members + pm->n_routing;
  cudaMemsetAsync(dw, 0, sizeof(float) * C * H, s);
  cudaMemsetAsync(dbias, 0, sizeof(float) * C, s);
  HeadArgs a{};
  a.h_seq = h_seq; a.w = w; a.par_sim = const_cast<float*>(par_sim);
  a.B = B; a.T = T; a.H = H; a.m = pm->n_members; a.n_par = pm->n_par; a.warmup = pm->warmup;
  a.dpar_sim = dpar_sim; a.dh_seq = dh_seq; a.dw = dw; a.dbias = dbias; a.part = reinterpret_cast<float*>(ws);
  HY_REQUIRE(layout == HY2DL_LAYOUT_ROWMAJOR || (layout == HY2DL_LAYOUT_RI32 && H % 32 == 0), HY2DL_ERR_SHAPE,
             "hy2dl_head_map_bwd: unknown layout %d (RI32 needs H %% 32 == 0)", layout);
  a.ri32 = layout == HY2DL_LAYOUT_RI32;
  const int64_t Bp = a.ri32 ? cdiv(B, 32) * 32 : B;   // rows per step in dh_seq
  const size_t smem = head_smem(H, true);
  auto launch = [&](auto kern, unsigned grid, const HeadChunk& chk) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, kThreads, smem, s>>>(a, chk);
  };
  bool dyn_written = false;
  for (int i = 0; i < n; ++i) {
    if (!chunks[i].is_dynamic && !dyn_written && dh_seq) {
      // no dynamic launch has initialised rows warmup .. T-1: zero them once
      cudaMemsetAsync(dh_seq + (int64_t)pm->warmup * Bp * H, 0, sizeof(float) * (T - pm->warmup) * Bp * H, s);
      dyn_written = true;
    }
    a.accumulate_dh = chunks[i].is_dynamic ? (dyn_written ? 1 : 0) : 1;
    if (chunks[i].is_dynamic) dyn_written = true;
    const int64_t slots = chunks[i].is_dynamic ? (T - pm->warmup) : 1;
    const unsigned grid = (unsigned)std::min<int64_t>(slots * cdiv(B, kRows), std::min<int64_t>(kHeadBwdMaxBlocks, 2 * (int64_t)sm_count()));
    const int nc = chunks[i].ncol;
    if (nc <= 8) launch(head_bwd_kernel<8>, grid, chunks[i]);
    else if (nc <= 16) launch(head_bwd_kernel<16>, grid, chunks[i]);
    else launch(head_bwd_kernel<32>, grid, chunks[i]);
    head_bwd_finish_kernel<<<(unsigned)cdiv((int64_t)nc * H + nc, 256), 256, 0, s>>>(a.part, (int)grid, (int)H, chunks[i], dw, dbias);
  }
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}
