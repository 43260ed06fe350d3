// K5: GPU-resident sliding-window sampler (datasetzoo/basedataset.py:168-195 + default collate).
// The whole data set lives in HBM as concatenated per-basin series; one launch turns a batch of
// (basin, end index) pairs into the time-major native layouts the LSTM / conceptual kernels read,
// so no per-sample host work and no [B,T,I] host->device copy is left on the training path.
#include "common.cuh"

namespace hy2dl {

struct GatherArgs {
  const float *x_d, *x_s, *x_conc, *y_obs, *basin_std;
  const int64_t *row0, *basin, *end;
  int64_t B, L, n_last, nd, ns, nc, IP;
  float *xT, *forc, *y_win, *std_win;
  int32_t ri32;   // xT in RI32 [L][G][IP/4][32][4] (rows >= B of the last group are zero filled)
};

// grid: (ceil(B/32), ceil(L / kStepsPerBlock)).  A block owns 32 samples for a run of time steps: the
// (sample, column) a thread writes does not depend on the step, so it resolves its sources once -- a
// pointer into the resident dynamic series (advancing nd floats per step), a static attribute value
// (constant over the window, basedataset.py:175-177) or zero padding -- and then streams one
// contiguous [32][IP] slab per step.
constexpr int kStepsPerBlock = 73;
constexpr int kGatherThreads = 256;
constexpr int kMaxPerThread = 8;   // 32 rows x IP <= 64 columns / 256 threads

__global__ void __launch_bounds__(kGatherThreads) window_gather_kernel(GatherArgs a) {
  const int64_t t0 = (int64_t)blockIdx.y * kStepsPerBlock;
  const int64_t t1 = min(a.L, t0 + kStepsPerBlock);
  const int64_t b0 = (int64_t)blockIdx.x * 32;
  __shared__ int64_t s_row[32], s_basin[32];
  if (threadIdx.x < 32) {
    const int64_t b = b0 + threadIdx.x;
    if (b < a.B) {
      const int64_t bs = a.basin[b];
      s_basin[threadIdx.x] = bs;
      s_row[threadIdx.x] = a.row0[bs] + a.end[b] - (a.L - 1);   // resident row of step 0
    }
  }
  __syncthreads();
  const int64_t nb = min((int64_t)32, a.B - b0);
  const int64_t I = a.nd + a.ns;
  const int per_slab = 32 * (int)a.IP;                          // floats of one step's slab (RI32: IP = 32)
  const int n_mine = (per_slab + kGatherThreads - 1) / kGatherThreads;

  if (a.ri32) {
    // RI32 slab = 1024 floats: thread i owns words 4i .. 4i+3 (4 consecutive features of one row, inside one
    // swizzled 8-float unit) and writes them with one 16-byte store per step
    const int e0 = 4 * threadIdx.x;
    const int64_t lb = e0 >> 5;
    const int c0 = ((((e0 >> 3) & 3) ^ (int)(lb & 3)) << 3) | (e0 & 7);
    const float* src[4];
    float cst[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      src[i] = nullptr; cst[i] = 0.f;
      const int c = c0 + i;
      if (lb < nb) {
        if (c < a.nd) src[i] = a.x_d + s_row[lb] * a.nd + c;
        else if (c < I) cst[i] = __ldg(a.x_s + s_basin[lb] * a.ns + (c - a.nd));
      }
    }
    const bool any_dyn = src[0] || src[1] || src[2] || src[3];
    float4* dst = reinterpret_cast<float4*>(a.xT + (t0 * gridDim.x + blockIdx.x) * 1024) + threadIdx.x;
    for (int64_t t = t0; t < t1; ++t, dst += (int64_t)gridDim.x * 256) {
      float4 v = make_float4(cst[0], cst[1], cst[2], cst[3]);
      if (any_dyn) {
        if (src[0]) v.x = __ldg(src[0] + t * a.nd);
        if (src[1]) v.y = __ldg(src[1] + t * a.nd);
        if (src[2]) v.z = __ldg(src[2] + t * a.nd);
        if (src[3]) v.w = __ldg(src[3] + t * a.nd);
      }
      *dst = v;
    }
  } else {
    const float* src[kMaxPerThread];   // dynamic source at step 0 (nullptr: constant)
    float cst[kMaxPerThread];
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) {
      src[i] = nullptr; cst[i] = 0.f;
      const int e = threadIdx.x + i * kGatherThreads;
      if (i < n_mine && e < per_slab) {
        const int64_t lb = e / a.IP, c = e - lb * a.IP;
        if (lb < nb) {
          if (c < a.nd) src[i] = a.x_d + s_row[lb] * a.nd + c;
          else if (c < I) cst[i] = __ldg(a.x_s + s_basin[lb] * a.ns + (c - a.nd));
        }
      }
    }
    const int live = (int)(nb * a.IP);                           // rows >= nb do not exist in the row-major layout
    for (int64_t t = t0; t < t1; ++t) {
      float* dst = a.xT + (t * a.B + b0) * a.IP;
#pragma unroll
      for (int i = 0; i < kMaxPerThread; ++i) {
        const int e = threadIdx.x + i * kGatherThreads;
        if (i < n_mine && e < live) dst[e] = src[i] ? __ldg(src[i] + t * a.nd) : cst[i];
      }
    }
  }
  if (a.forc) {
    for (int64_t e = threadIdx.x; e < (t1 - t0) * 3 * nb; e += kGatherThreads) {
      const int64_t lb = e % nb, c = (e / nb) % 3, t = t0 + e / (3 * nb);
      const float* r = a.x_conc + (s_row[lb] + t) * a.nc;
      float v;
      if (c < 2) v = __ldg(r + c);
      else v = (a.nc == 4) ? (__ldg(r + 2) + __ldg(r + 3)) / 2.f : __ldg(r + 2);
      a.forc[(t * 3 + c) * a.B + b0 + lb] = v;
    }
  }
  const int64_t tw = a.L - a.n_last;                            // first step of the target window
  for (int64_t e = threadIdx.x; e < (t1 - t0) * nb; e += kGatherThreads) {
    const int64_t lb = e % nb, t = t0 + e / nb;
    if (t >= tw) {
      a.y_win[(t - tw) * a.B + b0 + lb] = __ldg(a.y_obs + s_row[lb] + t);
      a.std_win[(t - tw) * a.B + b0 + lb] = __ldg(a.basin_std + s_basin[lb]);
    }
  }
}

// [B][T][I] -> [T][B][IP] (zero padded).  Tile transpose through shared memory: reads are
// contiguous along (t, i) of one sample, writes contiguous along (b, i) of one step.
template <bool RI32>
__global__ void __launch_bounds__(256) pack_x_kernel(const float* __restrict__ x, int64_t B, int64_t T, int64_t I,
                                                      int64_t IP, float* __restrict__ xT) {
  extern __shared__ float tile[];  // [8 samples][8 steps * I]
  const int64_t b0 = (int64_t)blockIdx.x * 8, t0 = (int64_t)blockIdx.y * 8;
  const int64_t nb = min((int64_t)8, B - b0), nt = min((int64_t)8, T - t0);
  const int64_t run = nt * I;
  for (int64_t e = threadIdx.x; e < nb * run; e += blockDim.x) {
    const int64_t lb = e / run, r = e - lb * run;
    tile[lb * 8 * I + r] = __ldg(x + ((b0 + lb) * T + t0) * I + r);
  }
  __syncthreads();
  for (int64_t e = threadIdx.x; e < nt * nb * IP; e += blockDim.x) {
    const int64_t lt = e / (nb * IP), r = e - lt * nb * IP;
    const int64_t lb = r / IP, c = r - lb * IP;
    const float v = c < I ? tile[lb * 8 * I + lt * I + c] : 0.f;
    if (RI32) {   // IP = 32: one block per group
      const int64_t b = b0 + lb, G = (B + 31) / 32;
      xT[((t0 + lt) * G + (b >> 5)) * 1024 + (b & 31) * 32 + ((((c >> 3) ^ b) & 3) << 3) + (c & 7)] = v;
    } else {
      xT[((t0 + lt) * B + b0) * IP + r] = v;
    }
  }
}

}  // namespace hy2dl
using namespace hy2dl;

extern "C" int hy2dl_window_gather(const float* x_d, const float* x_s, const float* x_conc, const float* y_obs,
                                   const float* basin_std, const int64_t* row0, const int64_t* basin,
                                   const int64_t* end, int64_t B, int64_t L, int64_t n_last, int64_t nd, int64_t ns,
                                   int64_t nc, int64_t IP, float* xT, float* forc, float* y_win, float* std_win,
                                   int layout, hy2dl_stream_t stream) {
  HY_REQUIRE(B > 0 && L > 0 && n_last > 0 && n_last <= L && nd > 0 && ns >= 0, HY2DL_ERR_SHAPE,
             "hy2dl_window_gather: bad sizes B=%lld L=%lld n_last=%lld nd=%lld ns=%lld", (long long)B, (long long)L,
             (long long)n_last, (long long)nd, (long long)ns);
  HY_REQUIRE(IP >= nd + ns && IP % 4 == 0, HY2DL_ERR_SHAPE, "hy2dl_window_gather: IP=%lld must be >= nd+ns and a multiple of 4", (long long)IP);
  HY_REQUIRE(forc == nullptr || nc == 3 || nc == 4, HY2DL_ERR_SHAPE, "hy2dl_window_gather: nc must be 3 or 4");
  HY_REQUIRE(L <= 65535, HY2DL_ERR_SHAPE, "hy2dl_window_gather: L=%lld exceeds 65535", (long long)L);
  HY_NONNULL(x_d); HY_NONNULL(y_obs); HY_NONNULL(basin_std); HY_NONNULL(row0); HY_NONNULL(basin); HY_NONNULL(end);
  HY_PTR(xT); HY_NONNULL(y_win); HY_NONNULL(std_win);
  HY_REQUIRE(ns == 0 || x_s != nullptr, HY2DL_ERR_ALIGN, "hy2dl_window_gather: x_s is null but ns > 0");
  HY_REQUIRE(forc == nullptr || x_conc != nullptr, HY2DL_ERR_ALIGN, "hy2dl_window_gather: x_conc is null");
  HY_REQUIRE(layout == HY2DL_LAYOUT_ROWMAJOR || layout == HY2DL_LAYOUT_RI32, HY2DL_ERR_SHAPE, "hy2dl_window_gather: unknown layout %d", layout);
  HY_REQUIRE(layout != HY2DL_LAYOUT_RI32 || IP == 32, HY2DL_ERR_SHAPE, "hy2dl_window_gather: RI32 streams are 32 columns wide (IP=%lld)", (long long)IP);
  GatherArgs a{x_d, x_s, x_conc, y_obs, basin_std, row0, basin, end, B, L, n_last, nd, ns, nc, IP, xT, forc, y_win, std_win,
               layout == HY2DL_LAYOUT_RI32};
  HY_REQUIRE(IP <= 64, HY2DL_ERR_SHAPE, "hy2dl_window_gather: IP=%lld exceeds 64", (long long)IP);
  dim3 grid((unsigned)cdiv(B, 32), (unsigned)cdiv(L, kStepsPerBlock));
  window_gather_kernel<<<grid, kGatherThreads, 0, (cudaStream_t)stream>>>(a);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int hy2dl_pack_x(const float* x, int64_t B, int64_t T, int64_t I, int64_t IP, float* xT, int layout,
                            hy2dl_stream_t stream) {
  HY_REQUIRE(B > 0 && T > 0 && I > 0 && IP >= I && IP % 4 == 0, HY2DL_ERR_SHAPE,
             "hy2dl_pack_x: bad sizes B=%lld T=%lld I=%lld IP=%lld", (long long)B, (long long)T, (long long)I, (long long)IP);
  HY_REQUIRE(I <= 1024, HY2DL_ERR_SHAPE, "hy2dl_pack_x: I=%lld exceeds 1024", (long long)I);
  HY_NONNULL(x); HY_PTR(xT);
  dim3 grid((unsigned)cdiv(B, 8), (unsigned)cdiv(T, 8));
  HY_REQUIRE(grid.y <= 65535, HY2DL_ERR_SHAPE, "hy2dl_pack_x: T=%lld too long", (long long)T);
  const size_t smem = (size_t)(8 * 8 * I) * sizeof(float);
  if (layout == HY2DL_LAYOUT_RI32) {
    HY_REQUIRE(IP == 32, HY2DL_ERR_SHAPE, "hy2dl_pack_x: RI32 streams are 32 columns wide (IP=%lld)", (long long)IP);
    if (B % 32) cudaMemsetAsync(xT, 0, sizeof(float) * T * cdiv(B, 32) * 32 * IP, (cudaStream_t)stream);   // padded rows = 0
    pack_x_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(x, B, T, I, IP, xT);
  } else {
    HY_REQUIRE(layout == HY2DL_LAYOUT_ROWMAJOR, HY2DL_ERR_SHAPE, "hy2dl_pack_x: unknown layout %d", layout);
    pack_x_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(x, B, T, I, IP, xT);
  }
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}
