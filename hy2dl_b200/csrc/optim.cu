// Optimiser tail on one flat fp32 buffer: global-norm clip + Adam (torch defaults), fused
// (experiments/LSTM_CAMELS_GB.ipynb:298-299: clip_grad_norm_(params, 1); optimizer.step()).
#include "common.cuh"

namespace hy2dl {

// Deterministic (no floating-point atomics).  out layout, kNormFloats floats, zero-initialised once by the caller:
// [0] the squared norm (written by the last block to finish), [1] arrival ticket (integer atomic, reset by the last
// block), [2 + b] partial sum of block b; the last block adds them in a fixed order.
constexpr int kNormMaxBlocks = 296;
constexpr int kNormFloats = 2 + kNormMaxBlocks;

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, int64_t n, float* out) {
  float v = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) v += g[i] * g[i];
  __shared__ float sm[8];
  __shared__ int is_last;
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < 8 ? sm[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      out[2 + blockIdx.x] = s;
      __threadfence();
      is_last = atomicAdd(reinterpret_cast<unsigned*>(out + 1), 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    __threadfence();
    const volatile float* part = out + 2;
    float s = 0.f;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) s += part[b];
    s = warp_sum(s);
    if (threadIdx.x == 0) { out[0] = s; *reinterpret_cast<unsigned*>(out + 1) = 0u; }
  }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        const float* __restrict__ norm_sq, float grad_scale,
                                                        float max_norm, const float* __restrict__ lr, float b1, float b2,
                                                        float eps, const int64_t* __restrict__ step_count) {
  const int64_t step = *step_count + 1;
  // clip_grad_norm_: coef = min(1, max_norm / (||g|| + 1e-6)) applied to the (scaled) gradient
  float coef = grad_scale;
  if (max_norm > 0.f) {
    const float total = sqrtf(*norm_sq) * grad_scale;
    coef *= fminf(max_norm / (total + 1e-6f), 1.f);
  }
  const float bc1 = 1.f - powf(b1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, (float)step));
  const float step_size = *lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float gi = g[i] * coef;
    const float mi = m[i] * b1 + (1.f - b1) * gi;
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

__global__ void bump_step_kernel(int64_t* step_count) { *step_count += 1; }

}  // namespace hy2dl
using namespace hy2dl;

extern "C" int64_t hy2dl_sqnorm_floats(void) { return kNormFloats; }

extern "C" int hy2dl_grad_sqnorm(const float* grad, int64_t n, float* norm_sq, hy2dl_stream_t stream) {
  HY_REQUIRE(n > 0, HY2DL_ERR_SHAPE, "hy2dl_grad_sqnorm: empty input");
  HY_NONNULL(grad); HY_NONNULL(norm_sq);
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), std::min<int64_t>(kNormMaxBlocks, 2 * (int64_t)sm_count()));
  sqnorm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(grad, n, norm_sq);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int hy2dl_clip_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                               const float* norm_sq, float grad_scale, float max_norm, const float* lr, float beta1,
                               float beta2, float eps, int64_t* step_count, hy2dl_stream_t stream) {
  HY_REQUIRE(n > 0, HY2DL_ERR_SHAPE, "hy2dl_clip_adam: empty input");
  HY_NONNULL(param); HY_NONNULL(grad); HY_NONNULL(exp_avg); HY_NONNULL(exp_avg_sq); HY_NONNULL(lr);
  HY_REQUIRE(step_count != nullptr, HY2DL_ERR_ALIGN, "hy2dl_clip_adam: step_count is null");
  HY_REQUIRE(max_norm <= 0.f || norm_sq != nullptr, HY2DL_ERR_ALIGN, "hy2dl_clip_adam: norm_sq is null but clipping requested");
  const unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), 2 * (int64_t)sm_count());
  clip_adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, norm_sq, grad_scale,
                                                           max_norm, lr, beta1, beta2, eps, step_count);
  bump_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_count);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}
