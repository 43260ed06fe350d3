// K3b: gamma unit-hydrograph routing (modelzoo/uh_routing.py:42-44, 64-74, 92-109).
// uh[k][b] = normalised gamma pdf at k + 0.5 (15 taps), y[t][b] = sum_k uh[k][b] q[t-k][b].
// All arrays are time-major [T][B]: a warp touches 32 consecutive samples of one step.
#include "common.cuh"

namespace hy2dl {

constexpr int kUH = 15;

__global__ void uh_weights_kernel(const float* __restrict__ alpha, const float* __restrict__ beta, int64_t B,
                                  float* __restrict__ uh) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float a = alpha[b], be = beta[b];
  // same factorisation as the reference: coeff * x^(a-1) * exp(-x/beta), coeff = 1/(beta^a * Gamma(a))
  const float coeff = 1.f / (powf(be, a) * expf(lgammaf(a)));
  float pdf[kUH], sum = 0.f;
#pragma unroll
  for (int k = 0; k < kUH; ++k) {
    const float x = 0.5f + (float)k;
    pdf[k] = coeff * powf(x, a - 1.f) * expf(-x / be);
    sum += pdf[k];
  }
#pragma unroll
  for (int k = 0; k < kUH; ++k) uh[k * B + b] = pdf[k] / sum;
}

// y[t][b] = sum_k w[k][b] * x[t - sign*k][b]; sign=+1: causal convolution (forward), sign=-1:
// correlation (adjoint w.r.t. q).
template <int SIGN>
__global__ void uh_fir_kernel(const float* __restrict__ x, const float* __restrict__ w, int64_t B, int64_t T,
                              float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int64_t t = i / B, b = i - t * B;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < kUH; ++k) {
    const int64_t tt = t - SIGN * k;
    if (tt >= 0 && tt < T) acc += __ldg(w + k * B + b) * __ldg(x + tt * B + b);
  }
  y[i] = acc;
}

// per sample: G[k] = sum_t dy[t] q[t-k]; then through the normalisation and the gamma pdf.
// The -ln(beta) - digamma(alpha) and alpha/beta terms are constant over k and cancel exactly
// against the normalisation, so d uh_k / d alpha = uh_k (ln x_k - sum_j uh_j ln x_j) and
// d uh_k / d beta = uh_k (x_k - sum_j uh_j x_j) / beta^2.
__global__ void uh_param_grad_kernel(const float* __restrict__ q, const float* __restrict__ beta,
                                     const float* __restrict__ uh, const float* __restrict__ dy, int64_t B, int64_t T,
                                     float* __restrict__ dalpha, float* __restrict__ dbeta) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float G[kUH], win[kUH];
#pragma unroll
  for (int k = 0; k < kUH; ++k) { G[k] = 0.f; win[k] = 0.f; }
  for (int64_t t = 0; t < T; ++t) {
#pragma unroll
    for (int k = kUH - 1; k > 0; --k) win[k] = win[k - 1];
    win[0] = __ldg(q + t * B + b);
    const float g = __ldg(dy + t * B + b);
#pragma unroll
    for (int k = 0; k < kUH; ++k) G[k] += g * win[k];
  }
  float w[kUH], gbar = 0.f;
#pragma unroll
  for (int k = 0; k < kUH; ++k) { w[k] = uh[k * B + b]; gbar += G[k] * w[k]; }
  float da = 0.f, db = 0.f;
#pragma unroll
  for (int k = 0; k < kUH; ++k) {
    const float x = 0.5f + (float)k;
    const float c = (G[k] - gbar) * w[k];
    da += c * logf(x);
    db += c * x;
  }
  const float be = beta[b];
  dalpha[b] = da;
  dbeta[b] = db / (be * be);
}

}  // namespace hy2dl
using namespace hy2dl;

extern "C" int hy2dl_uh_fwd(const float* q, const float* alpha, const float* beta, int64_t B, int64_t T,
                            float* uh, float* y, hy2dl_stream_t stream) {
  HY_REQUIRE(B > 0 && T > 0, HY2DL_ERR_SHAPE, "hy2dl_uh_fwd: empty input");
  HY_NONNULL(q); HY_NONNULL(alpha); HY_NONNULL(beta); HY_NONNULL(uh); HY_NONNULL(y);
  cudaStream_t s = (cudaStream_t)stream;
  uh_weights_kernel<<<(unsigned)cdiv(B, 128), 128, 0, s>>>(alpha, beta, B, uh);
  uh_fir_kernel<+1><<<(unsigned)cdiv(B * T, 256), 256, 0, s>>>(q, uh, B, T, y);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int hy2dl_uh_bwd(const float* q, const float* alpha, const float* beta, const float* uh, const float* dy,
                            int64_t B, int64_t T, float* dq, float* dalpha, float* dbeta, hy2dl_stream_t stream) {
  (void)alpha;
  HY_REQUIRE(B > 0 && T > 0, HY2DL_ERR_SHAPE, "hy2dl_uh_bwd: empty input");
  HY_NONNULL(q); HY_NONNULL(beta); HY_NONNULL(uh); HY_NONNULL(dy); HY_NONNULL(dq); HY_NONNULL(dalpha); HY_NONNULL(dbeta);
  cudaStream_t s = (cudaStream_t)stream;
  uh_fir_kernel<-1><<<(unsigned)cdiv(B * T, 256), 256, 0, s>>>(dy, uh, B, T, dq);
  uh_param_grad_kernel<<<(unsigned)cdiv(B, 128), 128, 0, s>>>(q, beta, uh, dy, B, T, dalpha, dbeta);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}
