// K2 / K2' (fp32-exact path): persistent recurrent LSTM forward and back-propagation through time
// on CUDA-core FMAs, plus the time-parallel weight-gradient GEMM.  This is the parity path
// (HY2DL_PREC_FP32, rtol 1e-4 against torch.nn.LSTM through 365+ steps); the tcgen05 path lives
// in lstm_tc.cu.
//
// Work decomposition (forward and backward): one CTA owns a tile of BT = 8*SPT sequences for ALL
// time steps (the recurrence never leaves the SM: h_{t-1}, c and dh/dc stay in shared memory /
// registers).  Warp w owns sequences [w*SPT, (w+1)*SPT) of the tile, lane l owns hidden units
// {l, l+32} of each 64-unit chunk with all four gates, so the cell update is thread-local.
// The input projection is fused by K-concatenation: the step operand is z_t = [x_t | h_{t-1}],
// K = IP + H, so no Gx[B,T,4H] is ever written to HBM.
// Weights are streamed from L2 in 8 KB slices through a 4-stage cp.async ring (they are shared by
// every CTA, so after the first touch they are L2 hits); arithmetic intensity against that stream
// is 2*BT flop/4 B, far above the L2 roofline for BT >= 16.
//
// Reference semantics: torch.nn.LSTM, one layer, h0 = c0 = 0 (modelzoo/cudalstm.py:43-50,
// hybrid.py:84-92; equations torch/nn/modules/rnn.py:842-847; weight layout :935-941).
#include "common.cuh"
#include <stdlib.h>

namespace hy2dl {

constexpr int kT = 256;        // threads per CTA
constexpr int kStages = 4;     // cp.async ring depth
constexpr int kSliceFloats = 2048;  // 8 KB per weight slice

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int SPT>
struct Tile {
  static constexpr int BT = 8 * SPT;
  static constexpr int LD = BT + (SPT >= 4 ? 4 : 2);  // padded so that vector stores along s are conflict-free
};

// ------------------------------------------------------------------------------------------------
// weight packing: Wp [H/64][K = IP+H][2][32][4], bp [H/64][2][32][4]; unit u = ch*64 + uu*32 + lane
// ------------------------------------------------------------------------------------------------
__global__ void lstm_pack_w_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                   const float* __restrict__ b_ih, const float* __restrict__ b_hh, int I, int IP, int H,
                                   float* __restrict__ Wp, float* __restrict__ bp) {
  const int K = IP + H;
  const int64_t total = (int64_t)(H / 64) * K * 256;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int g = e & 3, lane = (e >> 2) & 31, uu = (e >> 7) & 1;
    const int64_t rest = e >> 8;
    const int k = (int)(rest % K), ch = (int)(rest / K);
    const int row = g * H + ch * 64 + uu * 32 + lane;
    float v;
    if (k < IP) v = k < I ? w_ih[(int64_t)row * I + k] : 0.f;
    else v = w_hh[(int64_t)row * H + (k - IP)];
    Wp[e] = v;
    if (k == 0) bp[ch * 256 + (e & 255)] = b_ih[row] + b_hh[row];
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
struct LstmFwdArgs {
  const float* xT;   // [T][B][IP]
  const float* Wp;   // packed weights
  const float* bp;   // packed bias
  float *h_seq, *c_seq, *gates, *h_last;
  int64_t B, T;
  int IP;
};

template <int H, int SPT>
__global__ void __launch_bounds__(kT, 1) lstm_fwd_kernel(LstmFwdArgs a) {
  using TL = Tile<SPT>;
  constexpr int BT = TL::BT, LD = TL::LD, NCH = H / 64;
  constexpr int KS = kSliceFloats / 256;  // k rows per slice (8)
  constexpr int XMAX = (BT * 64 + kT - 1) / kT;  // x elements per thread, IP <= 64
  extern __shared__ __align__(16) float smem[];
  const int IP = a.IP, K = IP + H;
  float* xs = smem;                       // [IP][LD]
  float* hs = xs + 64 * LD;               // [2][H][LD]
  float* wb = hs + 2 * H * LD;            // [kStages][2048]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * BT;
  const int nb = (int)min((int64_t)BT, a.B - b0);
  const int n_slices_step = NCH * (K / KS);
  const int64_t n_slices = (int64_t)a.T * n_slices_step;

  auto issue = [&](int64_t gs) {
    if (gs < n_slices) {
      const int sl = (int)(gs % n_slices_step);
      const float* src = a.Wp + (int64_t)sl * kSliceFloats;
      float* dst = wb + (gs % kStages) * kSliceFloats;
      cp_async16(dst + tid * 4, src + tid * 4);
      cp_async16(dst + (tid + kT) * 4, src + (tid + kT) * 4);
    }
    cp_async_commit();
  };

  // h_{-1} = 0, x_0 into xs
  for (int e = tid; e < H * LD; e += kT) hs[e] = 0.f;
  const int x_elems = nb * IP;
  for (int e = tid; e < BT * IP; e += kT) {
    const int s = e / IP, k = e - s * IP;
    xs[k * LD + s] = e < x_elems ? __ldg(a.xT + b0 * IP + e) : 0.f;
  }
  for (int i = 0; i < kStages - 1; ++i) issue(i);

  float c[NCH][SPT][2];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int s = 0; s < SPT; ++s) c[ch][s][0] = c[ch][s][1] = 0.f;

  int64_t gs = 0;
  int cur = 0;
  __syncthreads();
  for (int64_t t = 0; t < a.T; ++t) {
    // prefetch x_{t+1} into registers
    float xr[XMAX];
    const bool more = t + 1 < a.T;
    {
      const float* src = a.xT + ((t + 1) * a.B + b0) * IP;
#pragma unroll
      for (int i = 0; i < XMAX; ++i) {
        const int e = tid + i * kT;
        xr[i] = (more && e < x_elems) ? __ldg(src + e) : 0.f;
      }
    }
    const float* hcur = hs + cur * H * LD;
    float* hnxt = hs + (cur ^ 1) * H * LD;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float acc[SPT][2][4];
      {
        const float4 b0v = __ldg(reinterpret_cast<const float4*>(a.bp + ch * 256 + lane * 4));
        const float4 b1v = __ldg(reinterpret_cast<const float4*>(a.bp + ch * 256 + 128 + lane * 4));
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
          acc[s][0][0] = b0v.x; acc[s][0][1] = b0v.y; acc[s][0][2] = b0v.z; acc[s][0][3] = b0v.w;
          acc[s][1][0] = b1v.x; acc[s][1][1] = b1v.y; acc[s][1][2] = b1v.z; acc[s][1][3] = b1v.w;
        }
      }
      for (int k0 = 0; k0 < K; k0 += KS, ++gs) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        issue(gs + kStages - 1);
        const float* wsl = wb + (gs % kStages) * kSliceFloats;
        const float* zrow = (k0 < IP) ? xs + k0 * LD : hcur + (k0 - IP) * LD;  // IP % KS == 0: a slice never straddles
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
          float av[SPT];
          const float* zp = zrow + kk * LD + w * SPT;
          if constexpr (SPT == 2) {
            const float2 v = *reinterpret_cast<const float2*>(zp);
            av[0] = v.x; av[1] = v.y;
          } else {
#pragma unroll
            for (int q = 0; q < SPT / 4; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(zp + 4 * q);
              av[4 * q] = v.x; av[4 * q + 1] = v.y; av[4 * q + 2] = v.z; av[4 * q + 3] = v.w;
            }
          }
          const float4 w0 = *reinterpret_cast<const float4*>(wsl + kk * 256 + lane * 4);
          const float4 w1 = *reinterpret_cast<const float4*>(wsl + kk * 256 + 128 + lane * 4);
#pragma unroll
          for (int s = 0; s < SPT; ++s) {
            acc[s][0][0] = fmaf(av[s], w0.x, acc[s][0][0]);
            acc[s][0][1] = fmaf(av[s], w0.y, acc[s][0][1]);
            acc[s][0][2] = fmaf(av[s], w0.z, acc[s][0][2]);
            acc[s][0][3] = fmaf(av[s], w0.w, acc[s][0][3]);
            acc[s][1][0] = fmaf(av[s], w1.x, acc[s][1][0]);
            acc[s][1][1] = fmaf(av[s], w1.y, acc[s][1][1]);
            acc[s][1][2] = fmaf(av[s], w1.z, acc[s][1][2]);
            acc[s][1][3] = fmaf(av[s], w1.w, acc[s][1][3]);
          }
        }
      }
      // ---- cell update for (SPT sequences) x (units lane, lane+32 of chunk ch)
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        const int u = ch * 64 + uu * 32 + lane;
        float hv[SPT];
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
          const float gi = sigmoidf_acc(acc[s][uu][0]);
          const float gf = sigmoidf_acc(acc[s][uu][1]);
          const float gg = tanhf(acc[s][uu][2]);
          const float go = sigmoidf_acc(acc[s][uu][3]);
          const float cn = gf * c[ch][s][uu] + gi * gg;
          c[ch][s][uu] = cn;
          const float hn = go * tanhf(cn);
          hv[s] = hn;
          const int sl = w * SPT + s;
          if (sl < nb) {
            const int64_t row = t * a.B + b0 + sl;
            if (a.h_seq) a.h_seq[row * H + u] = hn;
            if (a.c_seq) a.c_seq[row * H + u] = cn;
            if (a.gates) {
              float* gp = a.gates + row * 4 * H + u;
              gp[0] = gi; gp[H] = gf; gp[2 * H] = gg; gp[3 * H] = go;
            }
            if (a.h_last && t == a.T - 1) a.h_last[(b0 + sl) * H + u] = hn;
          }
        }
        float* hp = hnxt + u * LD + w * SPT;
        if constexpr (SPT == 2) {
          *reinterpret_cast<float2*>(hp) = make_float2(hv[0], hv[1]);
        } else {
#pragma unroll
          for (int q = 0; q < SPT / 4; ++q)
            *reinterpret_cast<float4*>(hp + 4 * q) = make_float4(hv[4 * q], hv[4 * q + 1], hv[4 * q + 2], hv[4 * q + 3]);
        }
      }
    }
    __syncthreads();  // every warp is done reading xs / hs[cur]; hs[cur^1] is complete
#pragma unroll
    for (int i = 0; i < XMAX; ++i) {
      const int e = tid + i * kT;
      if (e < BT * IP) {
        const int s = e / IP, k = e - s * IP;
        xs[k * LD + s] = xr[i];
      }
    }
    cur ^= 1;
    __syncthreads();
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// backward through time
// ------------------------------------------------------------------------------------------------
struct LstmBwdArgs {
  const float* w_hh;   // [4H][H] PyTorch layout
  const float *gates, *c_seq, *dh_seq, *dh_last;
  float* dG;
  int64_t B, T, dh_t0;
};

template <int H, int SPT>
__global__ void __launch_bounds__(kT, 1) lstm_bwd_kernel(LstmBwdArgs a) {
  using TL = Tile<SPT>;
  constexpr int BT = TL::BT, LD = TL::LD, NCH = H / 64;
  constexpr int KPT = H / 32;                 // dh outputs (k) per thread
  constexpr int VEC = KPT >= 4 ? 4 : 2;       // floats per vector load of a W_hh row
  constexpr int NQ = KPT / VEC;
  constexpr int RS = kSliceFloats / H;        // W_hh rows per slice
  constexpr int SL_PER_CH = 256 / RS;
  extern __shared__ __align__(16) float smem[];
  float* dgs = smem;                 // [256][LD]  gate gradients of the current chunk, col = g*64 + ul
  float* dhs = dgs + 256 * LD;       // [H][LD]    recurrent dh
  float* wb = dhs + H * LD;          // [kStages][2048]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t b0 = (int64_t)blockIdx.x * BT;
  const int nb = (int)min((int64_t)BT, a.B - b0);
  const int n_slices_step = NCH * SL_PER_CH;
  const int64_t n_slices = a.T * (int64_t)n_slices_step;

  auto issue = [&](int64_t gs) {
    if (gs < n_slices) {
      const int sl = (int)(gs % n_slices_step);
      const int ch = sl / SL_PER_CH, c0 = (sl % SL_PER_CH) * RS;   // first local column of the slice
      const int g = c0 / 64, ul0 = c0 % 64;
      const float* src = a.w_hh + ((int64_t)g * H + ch * 64 + ul0) * H;
      float* dst = wb + (gs % kStages) * kSliceFloats;
      cp_async16(dst + tid * 4, src + tid * 4);
      cp_async16(dst + (tid + kT) * 4, src + (tid + kT) * 4);
    }
    cp_async_commit();
  };

  for (int e = tid; e < H * LD; e += kT) dhs[e] = 0.f;
  for (int i = 0; i < kStages - 1; ++i) issue(i);

  float dc[NCH][SPT][2];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int s = 0; s < SPT; ++s) dc[ch][s][0] = dc[ch][s][1] = 0.f;

  int64_t gs = 0;
  __syncthreads();
  for (int64_t t = a.T - 1; t >= 0; --t) {
    float dh_acc[SPT][KPT];
#pragma unroll
    for (int s = 0; s < SPT; ++s)
#pragma unroll
      for (int j = 0; j < KPT; ++j) dh_acc[s][j] = 0.f;

#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      // ---- phase 1: gate gradients for (SPT sequences) x (units lane, lane+32 of chunk ch)
#pragma unroll
      for (int uu = 0; uu < 2; ++uu) {
        const int ul = uu * 32 + lane, u = ch * 64 + ul;
        float dhv[SPT], di[SPT], df[SPT], dg[SPT], dO[SPT];
        {
          const float* hp = dhs + u * LD + w * SPT;
          if constexpr (SPT == 2) {
            const float2 v = *reinterpret_cast<const float2*>(hp);
            dhv[0] = v.x; dhv[1] = v.y;
          } else {
#pragma unroll
            for (int q = 0; q < SPT / 4; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(hp + 4 * q);
              dhv[4 * q] = v.x; dhv[4 * q + 1] = v.y; dhv[4 * q + 2] = v.z; dhv[4 * q + 3] = v.w;
            }
          }
        }
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
          const int sl = w * SPT + s;
          di[s] = df[s] = dg[s] = dO[s] = 0.f;
          if (sl < nb) {
            const int64_t row = t * a.B + b0 + sl;
            const float* gp = a.gates + row * 4 * H + u;
            const float gi = gp[0], gf = gp[H], gg = gp[2 * H], go = gp[3 * H];
            const float ct = a.c_seq[row * H + u];
            const float cp = t > 0 ? a.c_seq[(row - a.B) * H + u] : 0.f;
            float dh = dhv[s];
            if (a.dh_seq && t >= a.dh_t0) dh += a.dh_seq[row * H + u];
            if (a.dh_last && t == a.T - 1) dh += a.dh_last[(b0 + sl) * H + u];
            const float tc = tanhf(ct);
            const float dct = dc[ch][s][uu] + dh * go * (1.f - tc * tc);
            dO[s] = dh * tc * go * (1.f - go);
            di[s] = dct * gg * gi * (1.f - gi);
            df[s] = dct * cp * gf * (1.f - gf);
            dg[s] = dct * gi * (1.f - gg * gg);
            dc[ch][s][uu] = dct * gf;
            float* op = a.dG + row * 4 * H + u;   // may alias gates: this thread has read them already
            op[0] = di[s]; op[H] = df[s]; op[2 * H] = dg[s]; op[3 * H] = dO[s];
          }
        }
        float* d0 = dgs + (0 * 64 + ul) * LD + w * SPT;
        float* d1 = dgs + (1 * 64 + ul) * LD + w * SPT;
        float* d2 = dgs + (2 * 64 + ul) * LD + w * SPT;
        float* d3 = dgs + (3 * 64 + ul) * LD + w * SPT;
        if constexpr (SPT == 2) {
          *reinterpret_cast<float2*>(d0) = make_float2(di[0], di[1]);
          *reinterpret_cast<float2*>(d1) = make_float2(df[0], df[1]);
          *reinterpret_cast<float2*>(d2) = make_float2(dg[0], dg[1]);
          *reinterpret_cast<float2*>(d3) = make_float2(dO[0], dO[1]);
        } else {
#pragma unroll
          for (int q = 0; q < SPT / 4; ++q) {
            *reinterpret_cast<float4*>(d0 + 4 * q) = make_float4(di[4 * q], di[4 * q + 1], di[4 * q + 2], di[4 * q + 3]);
            *reinterpret_cast<float4*>(d1 + 4 * q) = make_float4(df[4 * q], df[4 * q + 1], df[4 * q + 2], df[4 * q + 3]);
            *reinterpret_cast<float4*>(d2 + 4 * q) = make_float4(dg[4 * q], dg[4 * q + 1], dg[4 * q + 2], dg[4 * q + 3]);
            *reinterpret_cast<float4*>(d3 + 4 * q) = make_float4(dO[4 * q], dO[4 * q + 1], dO[4 * q + 2], dO[4 * q + 3]);
          }
        }
      }
      // ---- phase 2: dh[s][k] += sum_col dG[s][col] * W_hh[col][k]
      for (int sc = 0; sc < SL_PER_CH; ++sc, ++gs) {
        cp_async_wait<kStages - 2>();
        __syncthreads();   // slice landed; also orders phase-1 writes of dgs before the reads below
        issue(gs + kStages - 1);
        const float* wsl = wb + (gs % kStages) * kSliceFloats;
        const float* drow = dgs + (sc * RS) * LD + w * SPT;
#pragma unroll 4
        for (int r = 0; r < RS; ++r) {
          float av[SPT];
          const float* zp = drow + r * LD;
          if constexpr (SPT == 2) {
            const float2 v = *reinterpret_cast<const float2*>(zp);
            av[0] = v.x; av[1] = v.y;
          } else {
#pragma unroll
            for (int q = 0; q < SPT / 4; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(zp + 4 * q);
              av[4 * q] = v.x; av[4 * q + 1] = v.y; av[4 * q + 2] = v.z; av[4 * q + 3] = v.w;
            }
          }
          float wv[KPT];
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const float* wp = wsl + r * H + q * 32 * VEC + lane * VEC;
            if constexpr (VEC == 2) {
              const float2 v = *reinterpret_cast<const float2*>(wp);
              wv[q * VEC] = v.x; wv[q * VEC + 1] = v.y;
            } else {
              const float4 v = *reinterpret_cast<const float4*>(wp);
              wv[q * VEC] = v.x; wv[q * VEC + 1] = v.y; wv[q * VEC + 2] = v.z; wv[q * VEC + 3] = v.w;
            }
          }
#pragma unroll
          for (int s = 0; s < SPT; ++s)
#pragma unroll
            for (int j = 0; j < KPT; ++j) dh_acc[s][j] = fmaf(av[s], wv[j], dh_acc[s][j]);
        }
      }
      __syncthreads();   // all reads of dgs done before the next chunk overwrites it
    }
    // ---- recurrent dh for step t-1: thread owns k = q*32*VEC + lane*VEC + j
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int k = q * 32 * VEC + lane * VEC + j;
        float* hp = dhs + k * LD + w * SPT;
        if constexpr (SPT == 2) {
          *reinterpret_cast<float2*>(hp) = make_float2(dh_acc[0][q * VEC + j], dh_acc[1][q * VEC + j]);
        } else {
#pragma unroll
          for (int p = 0; p < SPT / 4; ++p)
            *reinterpret_cast<float4*>(hp + 4 * p) =
                make_float4(dh_acc[4 * p][q * VEC + j], dh_acc[4 * p + 1][q * VEC + j], dh_acc[4 * p + 2][q * VEC + j],
                            dh_acc[4 * p + 3][q * VEC + j]);
        }
      }
    __syncthreads();
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// weight gradients: C[m][n] = sum_r dG[r][m] * Z[r][n],  r = (t, b), Z = [h_{t-1} | x_t | 1]
// split-K over r with deterministic two-stage reduction
// ------------------------------------------------------------------------------------------------
struct WgradArgs {
  const float *dG, *h_seq, *xT;
  float* partial;   // [S][M][NP]
  int64_t B, R;     // R = T*B rows
  int H, IP, M, NP, rows_per_slab;
};

__global__ void __launch_bounds__(kT) lstm_wgrad_kernel(WgradArgs a) {
  constexpr int KB = 8;
  __shared__ __align__(16) float As[2][KB][128];
  __shared__ __align__(16) float Bs[2][KB][128];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
  const int64_t r_begin = (int64_t)blockIdx.z * a.rows_per_slab;
  const int64_t r_end = min(a.R, r_begin + a.rows_per_slab);
  // thread tile 8x8 as two 4-wide halves 64 apart (keeps the float4 smem reads conflict-free):
  // rows m0 + {ty*4 + i, 64 + ty*4 + i}, cols n0 + {tx*4 + j, 64 + tx*4 + j}
  const int ty = tid / 16, tx = tid % 16;
  const int lr = tid / 32, lc = (tid % 32) * 4;  // loader: row lr of the stage, float4 at column lc
  const int H = a.H, IP = a.IP, Ntot = H + IP + 1;

  auto load_a = [&](int64_t r) -> float4 {
    if (r < r_end) return __ldg(reinterpret_cast<const float4*>(a.dG + r * a.M + m0 + lc));
    return make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto load_b = [&](int64_t r) -> float4 {
    const int n = n0 + lc;
    if (r >= r_end || n >= Ntot) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < H) {
      if (r < a.B) return make_float4(0.f, 0.f, 0.f, 0.f);  // t = 0: h_{-1} = 0
      return __ldg(reinterpret_cast<const float4*>(a.h_seq + (r - a.B) * H + n));
    }
    if (n < H + IP) return __ldg(reinterpret_cast<const float4*>(a.xT + r * IP + (n - H)));
    return make_float4(1.f, 0.f, 0.f, 0.f);  // bias column
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra = load_a(r_begin + lr), rb = load_b(r_begin + lr);
  int buf = 0;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += KB) {
    *reinterpret_cast<float4*>(&As[buf][lr][lc]) = ra;
    *reinterpret_cast<float4*>(&Bs[buf][lr][lc]) = rb;
    __syncthreads();
    if (r0 + KB < r_end) { ra = load_a(r0 + KB + lr); rb = load_b(r0 + KB + lr); }
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    float* out = a.partial + ((int64_t)blockIdx.z * a.M + m) * a.NP + n0 + tx * 4;
    *reinterpret_cast<float4*>(out) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(out + 64) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

__global__ void lstm_wgrad_reduce_kernel(const float* __restrict__ partial, int S, int M, int NP, int H, int I, int IP,
                                         float* __restrict__ dw_ih, float* __restrict__ dw_hh, float* __restrict__ db) {
  const int Ntot = H + IP + 1;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)M * Ntot) return;
  const int m = (int)(e / Ntot), n = (int)(e % Ntot);
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += partial[((int64_t)k * M + m) * NP + n];
  if (n < H) dw_hh[(int64_t)m * H + n] = s;
  else if (n < H + IP) { if (n - H < I) dw_ih[(int64_t)m * I + (n - H)] = s; }
  else db[m] = s;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int SPT>
static size_t fwd_smem(int H) { return sizeof(float) * ((size_t)64 * Tile<SPT>::LD + 2 * (size_t)H * Tile<SPT>::LD + kStages * kSliceFloats); }
template <int SPT>
static size_t bwd_smem(int H) { return sizeof(float) * ((size_t)256 * Tile<SPT>::LD + (size_t)H * Tile<SPT>::LD + kStages * kSliceFloats); }

// sequences per thread: fill the SMs first (small batches are latency bound), then grow the tile
static int pick_spt(int64_t B) {
  const int sms = sm_count();
  if (B >= (int64_t)64 * sms) return 8;
  if (B >= (int64_t)32 * sms) return 4;
  return 2;
}

template <int H, int SPT>
static int launch_fwd(const LstmFwdArgs& a, cudaStream_t s) {
  const size_t smem = fwd_smem<SPT>(H);
  cudaFuncSetAttribute(lstm_fwd_kernel<H, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lstm_fwd_kernel<H, SPT><<<(unsigned)cdiv(a.B, Tile<SPT>::BT), kT, smem, s>>>(a);
  return 0;
}
template <int H, int SPT>
static int launch_bwd(const LstmBwdArgs& a, cudaStream_t s) {
  const size_t smem = bwd_smem<SPT>(H);
  cudaFuncSetAttribute(lstm_bwd_kernel<H, SPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lstm_bwd_kernel<H, SPT><<<(unsigned)cdiv(a.B, Tile<SPT>::BT), kT, smem, s>>>(a);
  return 0;
}

#define HY_DISPATCH_H_SPT(H_, SPT_, FN, ...)                                   \
  do {                                                                         \
    if ((H_) == 64) {                                                          \
      if ((SPT_) == 8) FN<64, 8>(__VA_ARGS__);                                 \
      else if ((SPT_) == 4) FN<64, 4>(__VA_ARGS__);                            \
      else FN<64, 2>(__VA_ARGS__);                                             \
    } else if ((H_) == 128) {                                                  \
      if ((SPT_) == 8) FN<128, 8>(__VA_ARGS__);                                \
      else if ((SPT_) == 4) FN<128, 4>(__VA_ARGS__);                           \
      else FN<128, 2>(__VA_ARGS__);                                            \
    } else {                                                                   \
      if ((SPT_) == 8) FN<256, 8>(__VA_ARGS__);                                \
      else if ((SPT_) == 4) FN<256, 4>(__VA_ARGS__);                           \
      else FN<256, 2>(__VA_ARGS__);                                            \
    }                                                                          \
  } while (0)

static int wgrad_splits(int64_t R, int M, int NP) {
  const int tiles = (M / 128) * (NP / 128);
  int64_t S = cdiv(4 * (int64_t)sm_count(), tiles);
  S = std::min<int64_t>(S, cdiv(R, 64));
  return (int)std::max<int64_t>(S, 1);
}

int64_t lstm_fp32_workspace(int kind, int64_t B, int64_t T, int64_t I, int64_t H);

// H = 128 / 256: the tensor-core chain (lstm_rm_tc_fwd.cu, lstm_rm_tc_bwd.cu, lstm_wgrad_rm_tc.cu).  Its tapes c_seq and
// gates / dG are stored in the blocked RB8 layout (tc_common.cuh), so the three kernels are selected together:
// HY2DL_FP32_TC=0 switches the whole chain back to the CUDA-core kernels of this file (row-major tapes).
namespace tc {
bool lstm_rm_tc_fwd_supported(int64_t I, int64_t H);
int64_t lstm_rm_tc_fwd_workspace(int64_t B, int64_t I, int64_t H);
int lstm_rm_tc_fwd(const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H, const float* w_ih, const float* w_hh,
                   const float* b_ih, const float* b_hh, float* h_seq, float* c_seq, float* gates, float* h_last, void* ws,
                   int64_t ws_bytes, int single, cudaStream_t s);
}  // namespace tc

namespace tc {
bool lstm_rm_tc_bwd_supported(int64_t H);
bool lstm_wgrad_rm_tc_supported(int64_t H);
}  // namespace tc
static bool rm_tc_chain(int64_t I, int64_t H) {
  static const bool off = [] { const char* e = getenv("HY2DL_FP32_TC"); return e && e[0] == '0'; }();
  return !off && tc::lstm_rm_tc_fwd_supported(I, H) && tc::lstm_rm_tc_bwd_supported(H) && tc::lstm_wgrad_rm_tc_supported(H);
}

// H = 64 at small batches: the unit-split forward / BPTT kernels (lstm_us_tc_*.cu, 4 CTAs per 128-sequence tile) and the
// tensor-core weight-gradient GEMM, on RB8 tapes like the H = 128 / 256 chain -- instead of the CUDA-core kernels below,
// whose tile of sequences sits on one SM for all T steps.  One round of 37 groups (4 736 sequences) takes 2.9 ms forward +
// BPTT at T = 365 where the CUDA-core kernels need 5.7 ms and the resident-weights kernels (lstm_tc.cu) 5.5 ms.  The three
// kernels share the tapes, so the rule depends only on what all three calls know: B and H.
namespace tc {
bool lstm_us_tc_fwd_preferred(int64_t B, int64_t I, int64_t H);
bool lstm_us_tc_bwd_preferred(int64_t B, int64_t H);
int64_t lstm_us_tc_fwd_workspace(int64_t B, int64_t I, int64_t H);
int64_t lstm_us_tc_bwd_workspace(int64_t B, int64_t H);
int lstm_us_tc_fwd(const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H, const float* w_ih, const float* w_hh,
                   const float* b_ih, const float* b_hh, float* h_seq, float* c_seq, float* gates, float* h_last, void* ws,
                   int64_t ws_bytes, int single, cudaStream_t s);
int lstm_us_tc_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates, const float* c_seq, const float* dh_seq,
                   int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, int single, cudaStream_t s);
int64_t lstm_wgrad_rm_tc_workspace(int64_t H, int64_t I);
int lstm_wgrad_rm_tc(const float* dG, const float* h_seq, const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H,
                     float* dw_ih, float* dw_hh, float* db, void* ws, int64_t ws_bytes, int single, cudaStream_t s);
}  // namespace tc
static bool us64_chain(int64_t B, int64_t H) {
  static const bool off = [] { const char* e = getenv("HY2DL_FP32_TC"); return e && e[0] == '0'; }();
  return !off && H == 64 && tc::lstm_us_tc_fwd_preferred(B, 64, 64) && tc::lstm_us_tc_bwd_preferred(B, 64) &&
         tc::lstm_wgrad_rm_tc_supported(64);
}

int lstm_fp32_fwd(const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H, const float* w_ih,
                  const float* w_hh, const float* b_ih, const float* b_hh, float* h_seq, float* c_seq, float* gates,
                  float* h_last, void* ws, int64_t ws_bytes, int single, cudaStream_t s) {
  if (rm_tc_chain(I, H))
    return tc::lstm_rm_tc_fwd(xT, B, T, I, IP, H, w_ih, w_hh, b_ih, b_hh, h_seq, c_seq, gates, h_last, ws, ws_bytes, single, s);
  if (us64_chain(B, H))
    return tc::lstm_us_tc_fwd(xT, B, T, I, IP, H, w_ih, w_hh, b_ih, b_hh, h_seq, c_seq, gates, h_last, ws, ws_bytes, single, s);
  HY_REQUIRE(!single, HY2DL_ERR_SHAPE, "hy2dl_lstm_fwd: precision tf32 runs on the tensor-core kernels only (H = 64 with input_size <= 31, "
             "H = 128 / 256); got H=%lld I=%lld", (long long)H, (long long)I);
  const int64_t need = lstm_fp32_workspace(HY2DL_WS_LSTM_FWD, B, T, I, H);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_fwd: workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  float* Wp = reinterpret_cast<float*>(ws);
  float* bp = Wp + (IP + H) * 4 * H;
  lstm_pack_w_kernel<<<(unsigned)std::min<int64_t>(cdiv((IP + H) * 4 * H, 256), 1024), 256, 0, s>>>(
      w_ih, w_hh, b_ih, b_hh, (int)I, (int)IP, (int)H, Wp, bp);
  LstmFwdArgs a{xT, Wp, bp, h_seq, c_seq, gates, h_last, B, T, (int)IP};
  const int spt = pick_spt(B);
  HY_DISPATCH_H_SPT(H, spt, launch_fwd, a, s);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

// tensor-core BPTT with streamed W_hh on the same row-major tapes (lstm_rm_tc_bwd.cu), H = 128 / 256
namespace tc {
bool lstm_rm_tc_bwd_supported(int64_t H);
int64_t lstm_rm_tc_bwd_workspace(int64_t B, int64_t H);
int lstm_rm_tc_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates, const float* c_seq, const float* dh_seq,
                   int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, int single, cudaStream_t s);
}  // namespace tc

int lstm_fp32_bwd(int64_t B, int64_t T, int64_t H, const float* w_hh, const float* gates, const float* c_seq,
                  const float* dh_seq, int64_t dh_t0, const float* dh_last, float* dG, void* ws, int64_t ws_bytes, int single,
                  cudaStream_t s) {
  if (rm_tc_chain(8, H))     // the chain's condition on the input size (I <= 64) holds for every accepted call
    return tc::lstm_rm_tc_bwd(B, T, H, w_hh, gates, c_seq, dh_seq, dh_t0, dh_last, dG, ws, ws_bytes, single, s);
  if (us64_chain(B, H))
    return tc::lstm_us_tc_bwd(B, T, H, w_hh, gates, c_seq, dh_seq, dh_t0, dh_last, dG, ws, ws_bytes, single, s);
  HY_REQUIRE(!single, HY2DL_ERR_SHAPE, "hy2dl_lstm_bwd: precision tf32 runs on the tensor-core kernels only (H=%lld)", (long long)H);
  LstmBwdArgs a{w_hh, gates, c_seq, dh_seq, dh_last, dG, B, T, dh_t0};
  const int spt = pick_spt(B);
  HY_DISPATCH_H_SPT(H, spt, launch_bwd, a, s);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

// tensor-core GEMM on the same row-major tapes (lstm_wgrad_rm_tc.cu), H = 128 / 256
namespace tc {
bool lstm_wgrad_rm_tc_supported(int64_t H);
int64_t lstm_wgrad_rm_tc_workspace(int64_t H, int64_t I);
int lstm_wgrad_rm_tc(const float* dG, const float* h_seq, const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP, int64_t H,
                     float* dw_ih, float* dw_hh, float* db, void* ws, int64_t ws_bytes, int single, cudaStream_t s);
}  // namespace tc

int lstm_fp32_wgrad(const float* dG, const float* h_seq, const float* xT, int64_t B, int64_t T, int64_t I, int64_t IP,
                    int64_t H, float* dw_ih, float* dw_hh, float* db, void* ws, int64_t ws_bytes, int single, cudaStream_t s) {
  if (rm_tc_chain(I, H) || us64_chain(B, H))
    return tc::lstm_wgrad_rm_tc(dG, h_seq, xT, B, T, I, IP, H, dw_ih, dw_hh, db, ws, ws_bytes, single, s);
  HY_REQUIRE(!single, HY2DL_ERR_SHAPE, "hy2dl_lstm_wgrad: precision tf32 runs on the tensor-core kernels only (H=%lld)", (long long)H);
  const int64_t need = lstm_fp32_workspace(HY2DL_WS_LSTM_WGRAD, B, T, I, H);
  HY_REQUIRE(ws != nullptr && ws_bytes >= need, HY2DL_ERR_WORKSPACE, "hy2dl_lstm_wgrad: workspace %lld < %lld bytes",
             (long long)ws_bytes, (long long)need);
  WgradArgs a{};
  a.dG = dG; a.h_seq = h_seq; a.xT = xT; a.partial = reinterpret_cast<float*>(ws);
  a.B = B; a.R = T * B; a.H = (int)H; a.IP = (int)IP; a.M = (int)(4 * H);
  a.NP = (int)round_up(H + IP + 1, 128);
  const int S = wgrad_splits(a.R, a.M, a.NP);
  a.rows_per_slab = (int)round_up(cdiv(a.R, S), 8);
  dim3 grid(a.M / 128, a.NP / 128, S);
  lstm_wgrad_kernel<<<grid, kT, 0, s>>>(a);
  const int64_t n_out = (int64_t)a.M * (H + IP + 1);
  lstm_wgrad_reduce_kernel<<<(unsigned)cdiv(n_out, 256), 256, 0, s>>>(a.partial, S, a.M, a.NP, (int)H, (int)I, (int)IP,
                                                                      dw_ih, dw_hh, db);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

int64_t lstm_fp32_workspace(int kind, int64_t B, int64_t T, int64_t I, int64_t H) {
  const int64_t IP = round_up(I, 8);
  if (kind == HY2DL_WS_LSTM_FWD) {
    if (rm_tc_chain(I, H)) return tc::lstm_rm_tc_fwd_workspace(B, I, H);
    if (us64_chain(B, H)) return tc::lstm_us_tc_fwd_workspace(B, I, H);
    return sizeof(float) * ((IP + H) * 4 * H + 4 * H);
  }
  if (kind == HY2DL_WS_LSTM_BWD) return rm_tc_chain(I, H) ? tc::lstm_rm_tc_bwd_workspace(B, H) : us64_chain(B, H) ? tc::lstm_us_tc_bwd_workspace(B, H) : 16;
  if (kind == HY2DL_WS_LSTM_WGRAD) {
    if (rm_tc_chain(I, H) || us64_chain(B, H)) return tc::lstm_wgrad_rm_tc_workspace(H, I);
    const int M = (int)(4 * H), NP = (int)round_up(H + IP + 1, 128);
    return sizeof(float) * (int64_t)wgrad_splits(T * B, M, NP) * M * NP;
  }
  return -1;
}

}  // namespace hy2dl
