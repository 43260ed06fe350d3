// TEST HARNESS (not product code): drives the bucket models' step and adjoint functions of
// hy2dl_b200/csrc/conceptual_models.cuh on the host, with the same loop structure as the K3
// kernels, so that `pytest -m "not gpu"` can check the hand-derived adjoints against the golden
// vectors of the reference on the CPU-only build box.  Built by tests/test_conceptual_host.py.
#include <stdint.h>
#include <vector>
#include "../../hy2dl_b200/csrc/conceptual_models.cuh"

using namespace hy2dl;

template <class M>
static void run(int64_t B, int64_t T, int64_t m, const float* forc, const float* const* par, const int64_t* ts,
                const float* init, int has_bet, float* states, float* q, const float* dq, const float* dstates,
                float* const* dpar, float* dinit) {
  const int64_t N = B * m;
  for (int64_t t = 0; t < T; ++t)
    for (int64_t b = 0; b < B; ++b) q[t * B + b] = 0.f;
  for (int64_t n = 0; n < N; ++n) {
    const int64_t b = n / m;
    constexpr int NS = M::NS;   // states / init / dinit here are [.][NS][N]
    float s[NS], p[M::NP];
    for (int k = 0; k < NS; ++k) s[k] = init ? init[k * N + n] : M::init(k);
    for (int64_t t = 0; t < T; ++t) {
      for (int k = 0; k < M::NP; ++k) p[k] = par[k] ? par[k][t * ts[k] + n] : 0.f;
      const float* f = forc + t * 3 * B + b;
      const float qm = M::step(s, p, f[0], f[B], f[2 * B], has_bet != 0);
      for (int k = 0; k < NS; ++k) states[(t * NS + k) * N + n] = s[k];
      q[t * B + b] += qm / (float)m;
    }
    if (!dq) continue;
    float g[NS];
    for (int k = 0; k < NS; ++k) g[k] = 0.f;
    std::vector<float> gacc(M::NP, 0.f);
    for (int64_t t = T - 1; t >= 0; --t) {
      for (int k = 0; k < M::NP; ++k) p[k] = par[k] ? par[k][t * ts[k] + n] : 0.f;
      for (int k = 0; k < NS; ++k) s[k] = t > 0 ? states[((t - 1) * NS + k) * N + n] : (init ? init[k * N + n] : M::init(k));
      if (dstates) for (int k = 0; k < NS; ++k) g[k] += dstates[(t * NS + k) * N + n];
      const float* f = forc + t * 3 * B + b;
      float gp[M::NP];
      M::adjoint(s, p, f[0], f[B], f[2 * B], has_bet != 0, g, dq[t * B + b] / (float)m, gp);
      for (int k = 0; k < M::NP; ++k) {
        if (!par[k]) continue;
        if (ts[k]) dpar[k][t * ts[k] + n] = gp[k];
        else gacc[k] += gp[k];
      }
    }
    for (int k = 0; k < M::NP; ++k) if (par[k] && !ts[k]) dpar[k][n] = gacc[k];
    if (dinit) for (int k = 0; k < NS; ++k) dinit[k * N + n] = g[k];
  }
}

extern "C" void conceptual_host(int model, int64_t B, int64_t T, int64_t m, const float* forc, const float* const* par,
                                const int64_t* ts, const float* init, int has_bet, float* states, float* q,
                                const float* dq, const float* dstates, float* const* dpar, float* dinit) {
  if (model == 0) run<SHM>(B, T, m, forc, par, ts, init, has_bet, states, q, dq, dstates, dpar, dinit);
  else if (model == 1) run<HBV>(B, T, m, forc, par, ts, init, has_bet, states, q, dq, dstates, dpar, dinit);
  else if (model == 2) run<LinRes>(B, T, m, forc, par, ts, init, has_bet, states, q, dq, dstates, dpar, dinit);
  else run<NonSense>(B, T, m, forc, par, ts, init, has_bet, states, q, dq, dstates, dpar, dinit);
}
