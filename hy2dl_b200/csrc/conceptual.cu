// K3: per-series conceptual bucket models (SHM, HBV, linear reservoir, NonSense) stepped through time, and their hand-derived
// adjoints.  One thread owns one series n = b*m + j (sample b, ensemble member j) for all steps;
// every load/store at a step is a contiguous run over n (time-major planes), i.e. fully coalesced.
// HBM-bound integer-free stream work: no tensor cores here by design.
//
// Reference semantics: modelzoo/shm.py:74-93,135-182, modelzoo/hbv.py:73-86,127-188,
// modelzoo/linear_reservoir.py:64-92 and modelzoo/nonsense.py:82-168; PyTorch
// sub-gradient conventions for min/max/clamp/pow are reproduced (see common.cuh and SURVEY.md 7).
#include "common.cuh"
#include "conceptual_models.cuh"

namespace hy2dl {

constexpr int kBlock = 128;
constexpr int kNS = HY2DL_N_STATES;

struct ConceptualArgs {
  const float* forc;    // [T_all][3][B]
  const float* par[HY2DL_MAX_PAR];
  int64_t ts[HY2DL_MAX_PAR];
  float* dpar[HY2DL_MAX_PAR];
  const float* init;    // [5][N] or null
  float* states;        // [Tn][5][N]   (const in bwd)
  float* q;             // [Tn][B]
  const float* dq;      // [Tn][B]
  const float* dstates; // [Tn][5][N] or null
  float* dinit;         // [5][N] or null
  int64_t B, t0, Tn, m, N;
  int has_betaet;
};

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(kBlock) conceptual_fwd_kernel(ConceptualArgs a, int tpb) {
  // tpb = floor(kBlock / m) * m threads of each block carry series, so that the members of one
  // sample never straddle two blocks (needed by the in-block member mean)
  __shared__ float red[kBlock];
  const int64_t n = (int64_t)blockIdx.x * tpb + threadIdx.x;
  const bool live = (int)threadIdx.x < tpb && n < a.N;
  const int64_t nn = live ? n : a.N - 1;   // tail threads shadow the last series, write nothing
  const int64_t b = nn / a.m;
  const int j = (int)(nn - b * a.m);
  const bool pow2 = (a.m & (a.m - 1)) == 0 && a.m <= 32;

  constexpr int NS = M::NS;   // planes NS..4 of the states buffer are not touched
  float s[NS], p[M::NP];
#pragma unroll
  for (int k = 0; k < NS; ++k) s[k] = a.init ? a.init[k * a.N + nn] : M::init(k);
#pragma unroll
  for (int k = 0; k < M::NP; ++k) p[k] = a.par[k] ? __ldg(a.par[k] + nn) : 0.f;  // static value / t = 0

  // software prefetch of the next step's inputs hides the load latency behind the dependent chain
  const float* f = a.forc + a.t0 * 3 * a.B + b;
  float P = __ldg(f), E = __ldg(f + a.B), Tm = __ldg(f + 2 * a.B);
  for (int64_t t = 0; t < a.Tn; ++t) {
    float Pn = 0.f, En = 0.f, Tn_ = 0.f, pn[M::NP];
    const bool more = t + 1 < a.Tn;
    if (more) {
      const float* fn = f + (t + 1) * 3 * a.B;
      Pn = __ldg(fn); En = __ldg(fn + a.B); Tn_ = __ldg(fn + 2 * a.B);
    }
#pragma unroll
    for (int k = 0; k < M::NP; ++k) pn[k] = (more && a.ts[k]) ? __ldg(a.par[k] + (t + 1) * a.ts[k] + nn) : p[k];

    const float qm = M::step(s, p, P, E, Tm, a.has_betaet != 0);

    if (live) {
      float* st = a.states + t * kNS * a.N + n;
#pragma unroll
      for (int k = 0; k < NS; ++k) st[k * a.N] = s[k];
    }
    // member mean (shm.py:182, hbv.py:188)
    float qs = qm;
    if (pow2) {
      for (int o = (int)a.m >> 1; o > 0; o >>= 1) qs += __shfl_xor_sync(0xffffffffu, qs, o);
      if (live && j == 0) a.q[t * a.B + b] = qs / (float)a.m;
    } else {
      red[threadIdx.x] = qm;
      __syncthreads();
      if (live && j == 0) {
        float acc = 0.f;
        for (int k = 0; k < (int)a.m; ++k) acc += red[threadIdx.x + k];
        a.q[t * a.B + b] = acc / (float)a.m;
      }
      __syncthreads();
    }
    P = Pn; E = En; Tm = Tn_;
#pragma unroll
    for (int k = 0; k < M::NP; ++k) p[k] = pn[k];
  }
}

template <class M>
__global__ void __launch_bounds__(kBlock) conceptual_bwd_kernel(ConceptualArgs a) {
  const int64_t n = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (n >= a.N) return;
  const int64_t b = n / a.m;
  const float inv_m = 1.f / (float)a.m;

  constexpr int NS = M::NS;
  float g[NS], gacc[M::NP], p[M::NP];
#pragma unroll
  for (int k = 0; k < NS; ++k) g[k] = 0.f;
#pragma unroll
  for (int k = 0; k < M::NP; ++k) { gacc[k] = 0.f; p[k] = a.par[k] ? __ldg(a.par[k] + n) : 0.f; }

  // inputs of one step: forcings, dynamic parameters, the state BEFORE the step, incoming gradients.  The
  // next (earlier) step's inputs are requested before this step's adjoint chain starts, so their HBM latency
  // hides behind ~100 dependent flops (one thread per series leaves nothing else to hide it with).
  struct In { float P, E, Tm, gq, p[M::NP], s[NS], ds[NS]; };
  auto load = [&](int64_t t, In& in) {
    const float* f = a.forc + (a.t0 + t) * 3 * a.B + b;
    in.P = __ldg(f); in.E = __ldg(f + a.B); in.Tm = __ldg(f + 2 * a.B);
#pragma unroll
    for (int k = 0; k < M::NP; ++k) in.p[k] = a.ts[k] ? __ldg(a.par[k] + t * a.ts[k] + n) : p[k];
    if (t > 0) {
      const float* st = a.states + (t - 1) * kNS * a.N + n;
#pragma unroll
      for (int k = 0; k < NS; ++k) in.s[k] = __ldg(st + k * a.N);
    } else {
#pragma unroll
      for (int k = 0; k < NS; ++k) in.s[k] = a.init ? a.init[k * a.N + n] : M::init(k);
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) in.ds[k] = a.dstates ? __ldg(a.dstates + (t * kNS + k) * a.N + n) : 0.f;
    in.gq = __ldg(a.dq + t * a.B + b) * inv_m;
  };
  In cur, nxt;
  load(a.Tn - 1, cur);
  for (int64_t t = a.Tn - 1; t >= 0; --t) {
    if (t > 0) load(t - 1, nxt);
#pragma unroll
    for (int k = 0; k < NS; ++k) g[k] += cur.ds[k];
    float gp[M::NP];
    M::adjoint(cur.s, cur.p, cur.P, cur.E, cur.Tm, a.has_betaet != 0, g, cur.gq, gp);
#pragma unroll
    for (int k = 0; k < M::NP; ++k) {
      if (a.ts[k]) {
        a.dpar[k][t * a.ts[k] + n] = gp[k];
      } else {
        gacc[k] += gp[k];
      }
    }
    cur = nxt;
  }
#pragma unroll
  for (int k = 0; k < M::NP; ++k)
    if (a.dpar[k] && !a.ts[k]) a.dpar[k][n] = gacc[k];
  if (a.dinit) {
#pragma unroll
    for (int k = 0; k < NS; ++k) a.dinit[k * a.N + n] = g[k];
  }
}

__global__ void pack_forcing_kernel(const float* __restrict__ xc, int64_t B, int64_t T, int nc, float* __restrict__ forc) {
  // one thread per (t, b): reads nc adjacent floats, writes three coalesced planes
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T) return;
  const int64_t t = i / B, b = i - t * B;
  const float* r = xc + (b * T + t) * nc;
  const float tm = (nc == 4) ? (r[2] + r[3]) / 2.f : (nc == 3 ? r[2] : 0.f);  // nc == 2: P and ET only (linear reservoir)
  float* o = forc + t * 3 * B + b;
  o[0] = r[0];
  o[B] = r[1];
  o[2 * B] = tm;
}

static int n_par_of(int model) {
  switch (model) {
    case HY2DL_MODEL_SHM: return SHM::NP;
    case HY2DL_MODEL_HBV: return HBV::NP;
    case HY2DL_MODEL_LINRES: return LinRes::NP;
    case HY2DL_MODEL_NONSENSE: return NonSense::NP;
    default: return 0;
  }
}

// runs `stmt` with M bound to the model's step/adjoint struct
#define HY_MODEL_DISPATCH(model, stmt)                                       \
  switch (model) {                                                           \
    case HY2DL_MODEL_SHM: { using M = SHM; stmt; } break;                    \
    case HY2DL_MODEL_HBV: { using M = HBV; stmt; } break;                    \
    case HY2DL_MODEL_LINRES: { using M = LinRes; stmt; } break;              \
    default: { using M = NonSense; stmt; } break;                            \
  }

static int fill_args(ConceptualArgs& a, int model, const float* forc, int64_t B, int64_t t0, int64_t Tn, int64_t m,
                     const float* const* par, const int64_t* tstride, const float* init) {
  const int np = n_par_of(model);
  HY_REQUIRE(np > 0, HY2DL_ERR_SHAPE, "conceptual: unknown model %d", model);
  HY_REQUIRE(B > 0 && Tn > 0 && t0 >= 0 && m > 0 && m <= kBlock, HY2DL_ERR_SHAPE,
             "conceptual: bad sizes B=%lld t0=%lld Tn=%lld m=%lld (m <= %d)", (long long)B, (long long)t0,
             (long long)Tn, (long long)m, kBlock);
  HY_NONNULL(forc);
  a = ConceptualArgs{};
  a.forc = forc; a.init = init; a.B = B; a.t0 = t0; a.Tn = Tn; a.m = m; a.N = B * m;
  a.has_betaet = 1;
  for (int k = 0; k < np; ++k) {
    if (par[k] == nullptr) {
      HY_REQUIRE(model == HY2DL_MODEL_HBV && k == HBV::BETAET, HY2DL_ERR_ALIGN, "conceptual: parameter %d is null", k);
      a.has_betaet = 0;
    }
    HY_REQUIRE(tstride[k] == 0 || tstride[k] == a.N, HY2DL_ERR_SHAPE, "conceptual: tstride[%d]=%lld must be 0 or N=%lld",
               k, (long long)tstride[k], (long long)a.N);
    a.par[k] = par[k];
    a.ts[k] = par[k] ? tstride[k] : 0;
  }
  return HY2DL_OK;
}

}  // namespace hy2dl

using namespace hy2dl;

extern "C" int hy2dl_pack_forcing(const float* xc, int64_t B, int64_t T, int64_t nc, float* forc, hy2dl_stream_t stream) {
  HY_REQUIRE(nc >= 2 && nc <= 4, HY2DL_ERR_SHAPE, "hy2dl_pack_forcing: nc must be 2, 3 or 4, got %lld", (long long)nc);
  HY_REQUIRE(B > 0 && T > 0, HY2DL_ERR_SHAPE, "hy2dl_pack_forcing: empty input");
  HY_NONNULL(xc); HY_NONNULL(forc);
  const int64_t n = B * T;
  pack_forcing_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(xc, B, T, (int)nc, forc);
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int hy2dl_conceptual_fwd(int model, const float* forc, int64_t B, int64_t t0, int64_t Tn, int64_t m,
                                    const float* const* par, const int64_t* tstride, const float* init,
                                    float* states, float* q, hy2dl_stream_t stream) {
  ConceptualArgs a;
  int rc = fill_args(a, model, forc, B, t0, Tn, m, par, tstride, init);
  if (rc) return rc;
  HY_NONNULL(states); HY_NONNULL(q);
  a.states = states; a.q = q;
  const int tpb = (int)((kBlock / m) * m);
  const unsigned grid = (unsigned)cdiv(a.N, tpb);
  HY_MODEL_DISPATCH(model, (conceptual_fwd_kernel<M><<<grid, kBlock, 0, (cudaStream_t)stream>>>(a, tpb)));
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}

extern "C" int hy2dl_conceptual_bwd(int model, const float* forc, int64_t B, int64_t t0, int64_t Tn, int64_t m,
                                    const float* const* par, const int64_t* tstride, const float* init,
                                    const float* states, const float* dq, const float* dstates,
                                    float* const* dpar, float* dinit, hy2dl_stream_t stream) {
  ConceptualArgs a;
  int rc = fill_args(a, model, forc, B, t0, Tn, m, par, tstride, init);
  if (rc) return rc;
  HY_NONNULL(states); HY_NONNULL(dq);
  a.states = const_cast<float*>(states); a.dq = dq; a.dstates = dstates; a.dinit = dinit;
  const int np = n_par_of(model);
  for (int k = 0; k < np; ++k) {
    HY_REQUIRE(dpar[k] != nullptr || a.par[k] == nullptr, HY2DL_ERR_ALIGN, "hy2dl_conceptual_bwd: dpar[%d] is null", k);
    a.dpar[k] = dpar[k];
  }
  const unsigned grid = (unsigned)cdiv(a.N, kBlock);
  HY_MODEL_DISPATCH(model, (conceptual_bwd_kernel<M><<<grid, kBlock, 0, (cudaStream_t)stream>>>(a)));
  HY_LAUNCH_CHECK();
  return HY2DL_OK;
}
