// K4: fused NaN-masked losses with warp-shuffle reductions
// (aux_functions/functions_training.py:32-41 nse_basin_averaged, :72-81 weighted_rmse).
// Two launches: *_reduce accumulates masked partial sums (so that ranks can all-reduce them and
// the data-parallel loss equals the single-process mean), *_finish forms the scalar and dL/dy.
#include "common.cuh"

namespace hy2dl {

constexpr int kLossBlock = 256;

// Deterministic grid reduction (no floating-point atomics: the reference seeds everything, functions_aux.py
// set_random_seed, so two runs must give bit-identical weights).  acc layout, kLossAccFloats floats, zero-initialised
// by the caller:  [0,4) the sums (written by the last block to finish), [4] arrival ticket (integer atomic, reset to 0
// by the last block), [8 + 4 b, 8 + 4 b + NACC) the partial sums of block b.  The last block adds the partials in a
// fixed order (lane-strided, then the xor-shuffle tree), so the result depends on the grid size only.
constexpr int kLossMaxBlocks = 592;
constexpr int kLossAccFloats = 8 + 4 * kLossMaxBlocks;

template <int NACC>
__device__ __forceinline__ void block_accumulate(float (&v)[NACC], float* acc) {
  __shared__ float sm[NACC][kLossBlock / 32];
  __shared__ int is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    const float s = warp_sum(v[k]);
    if (lane == 0) sm[k][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float* mine = acc + 8 + 4 * blockIdx.x;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      float s = lane < kLossBlock / 32 ? sm[k][lane] : 0.f;
      s = warp_sum(s);
      if (lane == 0) mine[k] = s;
    }
    if (lane == 0) {
      __threadfence();
      const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(acc + 4), 1u);
      is_last = ticket == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (is_last && warp == 0) {
    __threadfence();
    const volatile float* part = acc + 8;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      float s = 0.f;
      for (unsigned b = lane; b < gridDim.x; b += 32) s += part[4 * b + k];
      s = warp_sum(s);
      if (lane == 0) acc[k] = s;
    }
    if (lane == 0) *reinterpret_cast<unsigned*>(acc + 4) = 0u;
  }
}

__device__ __forceinline__ float logflow(float v) { return log10f(sqrtf(v + 1e-6f) + 0.1f); }

__global__ void __launch_bounds__(kLossBlock) nse_reduce_kernel(const float* __restrict__ ys, const float* __restrict__ yo,
                                                                const float* __restrict__ sd, int64_t n, float* acc) {
  float v[2] = {0.f, 0.f};
  for (int64_t i = (int64_t)blockIdx.x * kLossBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kLossBlock) {
    const float o = yo[i];
    if (o == o) {
      const float e = ys[i] - o, d = sd[i] + 0.1f;
      v[0] += (1.f / (d * d)) * (e * e);
      v[1] += 1.f;
    }
  }
  block_accumulate<2>(v, acc);
}

__global__ void __launch_bounds__(kLossBlock) nse_finish_kernel(const float* __restrict__ ys, const float* __restrict__ yo,
                                                                const float* __restrict__ sd, int64_t n,
                                                                const float* __restrict__ acc, const float* gscale,
                                                                float* loss, float* __restrict__ dy) {
  const float cnt = acc[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) *loss = acc[0] / cnt;
  if (!dy) return;
  const float gs = (gscale ? *gscale : 1.f) * 2.f / cnt;
  for (int64_t i = (int64_t)blockIdx.x * kLossBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kLossBlock) {
    const float o = yo[i];
    float g = 0.f;
    if (o == o) {
      const float d = sd[i] + 0.1f;
      g = gs * (1.f / (d * d)) * (ys[i] - o);
    }
    dy[i] = g;
  }
}

__global__ void __launch_bounds__(kLossBlock) wrmse_reduce_kernel(const float* __restrict__ ys, const float* __restrict__ yo,
                                                                  int64_t n, float* acc) {
  float v[3] = {0.f, 0.f, 0.f};
  for (int64_t i = (int64_t)blockIdx.x * kLossBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kLossBlock) {
    const float o = yo[i];
    if (o == o) {
      const float s = ys[i], e = s - o, le = logflow(s) - logflow(o);
      v[0] += e * e;
      v[1] += le * le;
      v[2] += 1.f;
    }
  }
  block_accumulate<3>(v, acc);
}

__global__ void __launch_bounds__(kLossBlock) wrmse_finish_kernel(const float* __restrict__ ys, const float* __restrict__ yo,
                                                                  int64_t n, const float* __restrict__ acc,
                                                                  const float* gscale, float* loss, float* __restrict__ dy) {
  const float cnt = acc[2];
  const float r1 = sqrtf(acc[0] / cnt), r2 = sqrtf(acc[1] / cnt);
  if (blockIdx.x == 0 && threadIdx.x == 0) *loss = 0.75f * r1 + 0.25f * r2;
  if (!dy) return;
  const float gs = gscale ? *gscale : 1.f;
  // d sqrt(S/N) / d y_i = (1 / (2 sqrt(S/N))) * (2 e_i / N)
  const float c1 = gs * 0.75f / (r1 * cnt), c2 = gs * 0.25f / (r2 * cnt);
  const float inv_ln10 = 0.43429448190325176f;
  for (int64_t i = (int64_t)blockIdx.x * kLossBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kLossBlock) {
    const float o = yo[i];
    float g = 0.f;
    if (o == o) {
      const float s = ys[i];
      const float rt = sqrtf(s + 1e-6f);
      const float dl = inv_ln10 / ((rt + 0.1f) * (2.f * rt));
      g = c1 * (s - o) + c2 * (logflow(s) - logflow(o)) * dl;
    }
    dy[i] = g;
  }
}

static unsigned loss_grid(int64_t n) { return (unsigned)std::min<int64_t>(cdiv(n, kLossBlock), std::min<int64_t>(kLossMaxBlocks, 4 * (int64_t)sm_count())); }

}  // namespace hy2dl
using namespace hy2dl;

extern "C" int64_t hy2dl_loss_acc_floats(void) { return kLossAccFloats; }

extern "C" int hy2dl_loss_nse_reduce(const float* y_sim, const float* y_obs, const float* basin_std, int64_t n,
                                     float* acc, hy2dl_stream_t stream) {
  HY_REQUIRE(n > 0, HY2DL_ERR_SHAPE, "hy2dl_loss_nse_reduce: empty input");
  HY_NONNULL(y_sim); HY_NONNULL(y_obs); HY_NONNULL(basin_std); HY_NONNULL(acc);
  nse_reduce_kernel<<<loss_grid(n), kLossBlock, 0, (cudaStream_t)stream>>>(y_sim, y_obs, basin_std, n, acc);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int hy2dl_loss_nse_finish(const float* y_sim, const float* y_obs, const float* basin_std, int64_t n,
                                     const float* acc, const float* gscale, float* loss, float* dy,
                                     hy2dl_stream_t stream) {
  HY_REQUIRE(n > 0, HY2DL_ERR_SHAPE, "hy2dl_loss_nse_finish: empty input");
  HY_NONNULL(y_sim); HY_NONNULL(y_obs); HY_NONNULL(basin_std); HY_NONNULL(acc); HY_NONNULL(loss);
  const unsigned grid = dy ? loss_grid(n) : 1u;
  nse_finish_kernel<<<grid, kLossBlock, 0, (cudaStream_t)stream>>>(y_sim, y_obs, basin_std, n, acc, gscale, loss, dy);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int hy2dl_loss_wrmse_reduce(const float* y_sim, const float* y_obs, int64_t n, float* acc,
                                       hy2dl_stream_t stream) {
  HY_REQUIRE(n > 0, HY2DL_ERR_SHAPE, "hy2dl_loss_wrmse_reduce: empty input");
  HY_NONNULL(y_sim); HY_NONNULL(y_obs); HY_NONNULL(acc);
  wrmse_reduce_kernel<<<loss_grid(n), kLossBlock, 0, (cudaStream_t)stream>>>(y_sim, y_obs, n, acc);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int hy2dl_loss_wrmse_finish(const float* y_sim, const float* y_obs, int64_t n, const float* acc,
                                       const float* gscale, float* loss, float* dy, hy2dl_stream_t stream) {
  HY_REQUIRE(n > 0, HY2DL_ERR_SHAPE, "hy2dl_loss_wrmse_finish: empty input");
  HY_NONNULL(y_sim); HY_NONNULL(y_obs); HY_NONNULL(acc); HY_NONNULL(loss);
  const unsigned grid = dy ? loss_grid(n) : 1u;
  wrmse_finish_kernel<<<grid, kLossBlock, 0, (cudaStream_t)stream>>>(y_sim, y_obs, n, acc, gscale, loss, dy);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}


// ------------------------------------------------------------------------------------------------
// evaluation: per-basin Nash-Sutcliffe efficiency of whole-period series on the device
// (aux_functions/functions_evaluation.py:28-48).  y_sim, y_obs are time-major [T][B] (the native layout of
// the discharge the hybrid chain produces); a block owns 32 neighbouring series and 8 time phases, so every
// load is a full 128-byte line.  Two passes (mean of the valid observations, then the two sums of squares),
// float64 accumulation like the reference's numpy arithmetic.
// ------------------------------------------------------------------------------------------------
namespace hy2dl {

constexpr int kEvalPhases = 8;

__global__ void __launch_bounds__(32 * kEvalPhases) nse_basins_kernel(const float* __restrict__ ys, const float* __restrict__ yo,
                                                                      int64_t T, int64_t B, double* __restrict__ out) {
  __shared__ double sm[2][kEvalPhases][32];
  const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * 32 + lane;
  const bool live = b < B;
  double so = 0.0, n = 0.0;
  if (live)
    for (int64_t t = ph; t < T; t += kEvalPhases) {
      const float s = ys[t * B + b], o = yo[t * B + b];
      if (s == s && o == o) { so += (double)o; n += 1.0; }
    }
  sm[0][ph][lane] = so; sm[1][ph][lane] = n;
  __syncthreads();
  so = 0.0; n = 0.0;
#pragma unroll
  for (int k = 0; k < kEvalPhases; ++k) { so += sm[0][k][lane]; n += sm[1][k][lane]; }
  __syncthreads();
  const double mean = n > 0.0 ? so / n : 0.0;
  double sse = 0.0, sso = 0.0;
  if (live)
    for (int64_t t = ph; t < T; t += kEvalPhases) {
      const float s = ys[t * B + b], o = yo[t * B + b];
      if (s == s && o == o) {
        const double e = (double)s - (double)o, d = (double)o - mean;
        sse += e * e; sso += d * d;
      }
    }
  sm[0][ph][lane] = sse; sm[1][ph][lane] = sso;
  __syncthreads();
  if (ph == 0 && live) {
    sse = 0.0; sso = 0.0;
#pragma unroll
    for (int k = 0; k < kEvalPhases; ++k) { sse += sm[0][k][lane]; sso += sm[1][k][lane]; }
    out[b] = n > 1.0 ? 1.0 - sse / sso : nan("");     // fewer than two valid pairs: NaN (functions_evaluation.py:45-48)
  }
}

}  // namespace hy2dl

extern "C" int hy2dl_nse_basins(const float* y_sim, const float* y_obs, int64_t T, int64_t B, double* nse,
                                hy2dl_stream_t stream) {
  HY_REQUIRE(T > 0 && B > 0, HY2DL_ERR_SHAPE, "hy2dl_nse_basins: empty input");
  HY_NONNULL(y_sim); HY_NONNULL(y_obs); HY_NONNULL(nse);
  nse_basins_kernel<<<(unsigned)cdiv(B, 32), 32 * kEvalPhases, 0, (cudaStream_t)stream>>>(y_sim, y_obs, T, B, nse);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}
