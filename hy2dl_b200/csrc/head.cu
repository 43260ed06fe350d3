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
constexpr int kThreads = 128;

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
  int64_t B, T, H;
  int32_t m, n_par, warmup;
  int32_t accumulate_dh;
  int32_t ri32;   // h_seq / dh_seq in RI32 [T][G][H/4][32][4] instead of [T][B][H]
};

// smem: hs[kRows][H+1] | ws[kChunk][H] | (bwd) dzs[kChunk][kRows+1] | (bwd) dws[kChunk][H] | dbs[kChunk]
template <bool BWD>
__global__ void __launch_bounds__(kThreads) head_kernel(HeadArgs a, HeadChunk ch) {
  extern __shared__ float sm[];
  const int H = (int)a.H;
  float* hs = sm;
  float* ws = hs + kRows * (H + 1);
  float* dzs = ws + kChunk * H;
  float* dws = dzs + kChunk * (kRows + 1);
  float* dbs = dws + kChunk * H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t N = a.B * a.m;

  for (int e = tid; e < ch.ncol * H; e += kThreads) ws[e] = __ldg(a.w + (int64_t)ch.col[e / H] * H + (e % H));
  if (BWD) {
    for (int e = tid; e < kChunk * H; e += kThreads) dws[e] = 0.f;
    if (tid < kChunk) dbs[tid] = 0.f;
  }
  // time slots: dynamic fwd: warmup-1 .. T-1 (T_sim+1 slots); dynamic bwd: warmup .. T-1; static: T-1
  const int64_t t_first = ch.is_dynamic ? (BWD ? a.warmup : a.warmup - 1) : a.T - 1;
  const int64_t n_slots = a.T - t_first;
  const int64_t tiles_b = cdiv(a.B, kRows);
  const int64_t n_tiles = n_slots * tiles_b;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t t = t_first + tile / tiles_b;
    const int64_t b0 = (tile % tiles_b) * kRows;
    const int nb = (int)min((int64_t)kRows, a.B - b0);
    __syncthreads();
    // the [nb][H] slab of step t is contiguous in h_seq
    if (a.ri32) {   // tile = one 32-row group, contiguous as [H/32][32 rows][32 floats], units swizzled
      const float* src = a.h_seq + (t * tiles_b + b0 / kRows) * kRows * H;
      for (int e = tid; e < kRows * H; e += kThreads) hs[ri32_row(e) * (H + 1) + ri32_feat(e)] = __ldg(src + e);
    } else {
      const float* src = a.h_seq + (t * a.B + b0) * H;
      for (int e = tid; e < nb * H; e += kThreads) hs[(e / H) * (H + 1) + (e % H)] = __ldg(src + e);
    }
    __syncthreads();

    if (!BWD) {
      // lanes = rows, warps = column subsets
      if (lane < nb) {
        const float* hr = hs + lane * (H + 1);
        for (int c = warp; c < ch.ncol; c += kThreads / 32) {
          const float* wr = ws + c * H;
          float z = 0.f;
#pragma unroll 8
          for (int k = 0; k < H; ++k) z = fmaf(hr[k], wr[k], z);
          z += __ldg(a.bias + ch.col[c]);
          const float val = ch.lo[c] + sigmoidf_acc(z) * (ch.hi[c] - ch.lo[c]);
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
    } else {
      // phase A: dz[c][r] = dpar * dval/dz, with sigma recovered from the stored value
      for (int e = tid; e < ch.ncol * kRows; e += kThreads) {
        const int c = e / kRows, r = e % kRows;
        float dz = 0.f;
        if (r < nb) {
          const int64_t b = b0 + r;
          const bool routing = ch.par[c] >= a.n_par;
          const int64_t idx = ch.off[c] + (ch.is_dynamic ? (t - a.warmup) * N : 0) + (routing ? b : b * a.m + ch.mem[c]);
          const float val = __ldg(a.par_sim + idx);
          dz = __ldg(a.dpar_sim + idx) * (val - ch.lo[c]) * (ch.hi[c] - val) / (ch.hi[c] - ch.lo[c]);
        }
        dzs[c * (kRows + 1) + r] = dz;
      }
      __syncthreads();
      // phase B: dh[r][k] (+)= sum_c dz[c][r] w[c][k]; consecutive threads -> consecutive k
      if (a.ri32) {   // all 32 rows of the group are written (dz = 0 for rows >= nb)
        float* dst = a.dh_seq + (t * tiles_b + b0 / kRows) * kRows * H;
        for (int e = tid; e < kRows * H; e += kThreads) {
          const int r = ri32_row(e), k = ri32_feat(e);
          float acc = 0.f;
          for (int c = 0; c < ch.ncol; ++c) acc = fmaf(dzs[c * (kRows + 1) + r], ws[c * H + k], acc);
          dst[e] = a.accumulate_dh ? dst[e] + acc : acc;
        }
      } else {
        float* dst = a.dh_seq + (t * a.B + b0) * H;
        for (int e = tid; e < nb * H; e += kThreads) {
          const int r = e / H, k = e % H;
          float acc = 0.f;
          for (int c = 0; c < ch.ncol; ++c) acc = fmaf(dzs[c * (kRows + 1) + r], ws[c * H + k], acc);
          dst[e] = a.accumulate_dh ? dst[e] + acc : acc;
        }
      }
      // phase C: dw[c][k] += sum_r dz[c][r] h[r][k]; each thread owns fixed (c,k) entries
      for (int e = tid; e < ch.ncol * H; e += kThreads) {
        const int c = e / H, k = e % H;
        float acc = 0.f;
        for (int r = 0; r < nb; ++r) acc = fmaf(dzs[c * (kRows + 1) + r], hs[r * (H + 1) + k], acc);
        dws[e] += acc;
      }
      if (tid < ch.ncol) {
        float acc = 0.f;
        for (int r = 0; r < nb; ++r) acc += dzs[tid * (kRows + 1) + r];
        dbs[tid] += acc;
      }
    }
  }
  if (BWD) {
    __syncthreads();
    for (int e = tid; e < ch.ncol * H; e += kThreads) atomicAdd(a.dw + (int64_t)ch.col[e / H] * H + (e % H), dws[e]);
    if (tid < ch.ncol) atomicAdd(a.dbias + ch.col[tid], dbs[tid]);
  }
}

static size_t head_smem(int64_t H, bool bwd) {
  size_t f = (size_t)kRows * (H + 1) + (size_t)kChunk * H;
  if (bwd) f += (size_t)kChunk * (kRows + 1) + (size_t)kChunk * H + kChunk;
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
  cudaFuncSetAttribute(head_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int i = 0; i < n; ++i) {
    const int64_t slots = chunks[i].is_dynamic ? (T - pm->warmup + 1) : 1;
    const unsigned grid = (unsigned)std::min<int64_t>(slots * cdiv(B, kRows), 8 * (int64_t)sm_count());
    head_kernel<false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(a, chunks[i]);
  }
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int hy2dl_head_map_bwd(const float* h_seq, const float* par_sim, const float* dpar_sim, int64_t B, int64_t T,
                                  int64_t H, const float* w, const hy2dl_parmap* pm, float* dh_seq, float* dw,
                                  float* dbias, int layout, hy2dl_stream_t stream) {
  int rc = check_parmap(pm, T, "hy2dl_head_map_bwd");
  if (rc) return rc;
  HY_REQUIRE(B > 0 && H > 0 && H <= 512, HY2DL_ERR_SHAPE, "hy2dl_head_map_bwd: bad sizes B=%lld H=%lld", (long long)B, (long long)H);
  HY_NONNULL(h_seq); HY_NONNULL(par_sim); HY_NONNULL(dpar_sim); HY_NONNULL(w); HY_NONNULL(dh_seq); HY_NONNULL(dw); HY_NONNULL(dbias);
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
  a.dpar_sim = dpar_sim; a.dh_seq = dh_seq; a.dw = dw; a.dbias = dbias;
  HY_REQUIRE(layout == HY2DL_LAYOUT_ROWMAJOR || (layout == HY2DL_LAYOUT_RI32 && H % 32 == 0), HY2DL_ERR_SHAPE,
             "hy2dl_head_map_bwd: unknown layout %d (RI32 needs H %% 32 == 0)", layout);
  a.ri32 = layout == HY2DL_LAYOUT_RI32;
  const int64_t Bp = a.ri32 ? cdiv(B, 32) * 32 : B;   // rows per step in dh_seq
  const size_t smem = head_smem(H, true);
  cudaFuncSetAttribute(head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  bool dyn_written = false;
  for (int i = 0; i < n; ++i) {
    if (!chunks[i].is_dynamic && !dyn_written) {
      // no dynamic launch has initialised rows warmup .. T-1: zero them once
      cudaMemsetAsync(dh_seq + (int64_t)pm->warmup * Bp * H, 0, sizeof(float) * (T - pm->warmup) * Bp * H, s);
      dyn_written = true;
    }
    a.accumulate_dh = chunks[i].is_dynamic ? (dyn_written ? 1 : 0) : 1;
    if (chunks[i].is_dynamic) dyn_written = true;
    const int64_t slots = chunks[i].is_dynamic ? (T - pm->warmup) : 1;
    const unsigned grid = (unsigned)std::min<int64_t>(slots * cdiv(B, kRows), 2 * (int64_t)sm_count());
    head_kernel<true><<<grid, kThreads, smem, s>>>(a, chunks[i]);
  }
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}
