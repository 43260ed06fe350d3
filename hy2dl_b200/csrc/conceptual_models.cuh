// Step functions and hand-derived adjoints of the bucket models (SHM, HBV, linear reservoir, NonSense).
// Plain C++ (no CUDA intrinsics) so that the very same code is compiled into the K3 kernels
// (conceptual.cu) and into the host-side test harness (tests/host/conceptual_host.cpp), which
// checks the adjoint algebra against the reference's autograd on the CPU-only build box.
//
// Reference semantics: modelzoo/shm.py:135-182, modelzoo/hbv.py:127-188, modelzoo/linear_reservoir.py:77-92,
// modelzoo/nonsense.py:82-168.  Every model declares NP parameters, NS <= HY2DL_N_STATES storages (planes
// 0..NS-1 of the states buffer, in the order of the reference's _initial_states dict) and their defaults.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HY_HD __host__ __device__ __forceinline__
#else
#define HY_HD inline
#endif

namespace hy2dl {

// sub-gradient weights with PyTorch's tie conventions ("Adjoint fidelity"):
// min/max(a,b): the smaller/larger operand gets the gradient, ties split 1/2 - 1/2.
HY_HD float w_lt(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }
// clamp(x, min=lo): gradient passes where x >= lo (bound included).
HY_HD float w_ge(float x, float lo) { return x >= lo ? 1.f : 0.f; }

// pow adjoint pieces with torch semantics (pow_backward_self / pow_backward_exponent)
HY_HD float pow_dbase(float base, float ex) { return ex == 0.f ? 0.f : ex * powf(base, ex - 1.f); }
HY_HD float pow_dexp(float base, float ex, float res) {
  const bool zero = (ex == 0.f) ? (base >= 0.f) : (base == 0.f && ex >= 0.f);
  return zero ? 0.f : res * logf(base);
}

// ------------------------------------------------------------------------------------------------
// SHM
// ------------------------------------------------------------------------------------------------
struct SHM {
  static constexpr int NP = 8;
  static constexpr int NS = 5;   // ss sf su si sb
  enum { DD, FTHR, SUMAX, BETA, PERC, KF, KI, KB };
  HY_HD static float init(int) { return 0.001f; }
  static constexpr float klu = 0.90f;

  // forward step: s[] is updated in place, returns member outflow
  HY_HD static float step(float* s, const float* p, float P, float E, float Tm, bool) {
    const bool cold = Tm < 0.f;
    const float mp = cold ? 0.f : Tm * p[DD];
    const float rain = cold ? 0.f : P, snowf = cold ? P : 0.f;
    const float qs = fminf(s[0], mp);
    s[0] = s[0] - qs + snowf;
    const float qsp = qs + rain;
    const float qf_in = fmaxf(0.f, qsp - p[FTHR]);
    const float qu_in = fminf(qsp, p[FTHR]);
    const float sf1 = s[1] + qf_in;
    const float qf = sf1 * p[KF];
    s[1] = sf1 - qf;
    const float psi = powf(s[2] / p[SUMAX], p[BETA]);
    const float su_t = s[2] + qu_in * (1.f - psi);
    const float su1 = fminf(su_t, p[SUMAX]);
    const float qu = qu_in * psi + fmaxf(0.f, su_t - p[SUMAX]);
    const float pwp = 0.8f * p[SUMAX];
    const float kth = (su1 <= pwp) ? su1 / p[SUMAX] : 1.f;
    const float ret = E * klu * kth;
    s[2] = fmaxf(0.f, su1 - ret);
    const float si1 = s[3] + qu * p[PERC];
    const float qi = si1 * p[KI];
    s[3] = si1 - qi;
    const float sb1 = s[4] + qu * (1.f - p[PERC]);
    const float qb = sb1 * p[KB];
    s[4] = sb1 - qb;
    return qf + qi + qb;
  }

  // adjoint of one step. s = storages BEFORE the step; g = adjoint of storages AFTER the step on
  // entry, adjoint of storages before the step on exit; gq = adjoint of the member outflow;
  // gp[] receives this step's parameter adjoints.
  HY_HD static void adjoint(const float* s, const float* p, float P, float E, float Tm, bool,
                                                 float* g, float gq, float* gp) {
    // ---- recompute the forward intermediates
    const bool cold = Tm < 0.f;
    const float mp = cold ? 0.f : Tm * p[DD];
    const float rain = cold ? 0.f : P;
    const float qs = fminf(s[0], mp);
    const float qsp = qs + rain;
    const float xthr = qsp - p[FTHR];
    const float qu_in = fminf(qsp, p[FTHR]);
    const float sf1 = s[1] + fmaxf(0.f, xthr);
    const float r = s[2] / p[SUMAX];
    const float psi = powf(r, p[BETA]);
    const float su_t = s[2] + qu_in * (1.f - psi);
    const float su1 = fminf(su_t, p[SUMAX]);
    const float ovx = su_t - p[SUMAX];
    const float qu = qu_in * psi + fmaxf(0.f, ovx);
    const bool wet = su1 <= 0.8f * p[SUMAX];
    const float kth = wet ? su1 / p[SUMAX] : 1.f;
    const float d = su1 - E * klu * kth;
    const float si1 = s[3] + qu * p[PERC];
    const float sb1 = s[4] + qu * (1.f - p[PERC]);
    // ---- baseflow reservoir
    const float g_qb = gq - g[4];
    const float g_sb1 = g[4] + g_qb * p[KB];
    gp[KB] = g_qb * sb1;
    float g_qu = g_sb1 * (1.f - p[PERC]);
    float g_perc = -g_sb1 * qu;
    g[4] = g_sb1;
    // ---- interflow reservoir
    const float g_qi = gq - g[3];
    const float g_si1 = g[3] + g_qi * p[KI];
    gp[KI] = g_qi * si1;
    g_qu += g_si1 * p[PERC];
    g_perc += g_si1 * qu;
    gp[PERC] = g_perc;
    g[3] = g_si1;
    // ---- evapotranspiration: su2 = max(0, d)
    const float g_d = g[2] * (d > 0.f ? 1.f : (d == 0.f ? 0.5f : 0.f));
    float g_su1 = g_d;
    float g_smax = 0.f;
    if (wet) {
      const float g_kth = -g_d * E * klu;
      g_su1 += g_kth / p[SUMAX];
      g_smax -= g_kth * su1 / (p[SUMAX] * p[SUMAX]);
    }
    // ---- unsaturated zone
    float g_qu_in = g_qu * psi;
    float g_psi = g_qu * qu_in;
    const float g_ovx = g_qu * (ovx > 0.f ? 1.f : (ovx == 0.f ? 0.5f : 0.f));
    float g_su_t = g_ovx;
    g_smax -= g_ovx;
    g_su_t += g_su1 * w_lt(su_t, p[SUMAX]);
    g_smax += g_su1 * w_lt(p[SUMAX], su_t);
    float g_su = g_su_t;
    g_qu_in += g_su_t * (1.f - psi);
    g_psi -= g_su_t * qu_in;
    const float g_r = g_psi * pow_dbase(r, p[BETA]);
    gp[BETA] = g_psi * pow_dexp(r, p[BETA], psi);
    g_su += g_r / p[SUMAX];
    g_smax -= g_r * s[2] / (p[SUMAX] * p[SUMAX]);
    gp[SUMAX] = g_smax;
    g[2] = g_su;
    // ---- fast reservoir
    const float g_qf = gq - g[1];
    const float g_sf1 = g[1] + g_qf * p[KF];
    gp[KF] = g_qf * sf1;
    g[1] = g_sf1;
    // ---- threshold split
    float g_qsp = g_qu_in * w_lt(qsp, p[FTHR]);
    float g_fthr = g_qu_in * w_lt(p[FTHR], qsp);
    const float g_x = g_sf1 * (xthr > 0.f ? 1.f : (xthr == 0.f ? 0.5f : 0.f));
    g_qsp += g_x;
    g_fthr -= g_x;
    gp[FTHR] = g_fthr;
    // ---- snow
    const float g_qs = g_qsp - g[0];
    g[0] = g[0] + g_qs * w_lt(s[0], mp);
    const float g_mp = g_qs * w_lt(mp, s[0]);
    gp[DD] = cold ? 0.f : g_mp * Tm;
  }
};

// ------------------------------------------------------------------------------------------------
// HBV
// ------------------------------------------------------------------------------------------------
struct HBV {
  static constexpr int NP = 13;
  static constexpr int NS = 5;   // SNOWPACK MELTWATER SM SUZ SLZ
  enum { BETA, FC, K0, K1, K2, LP, PERC, UZL, TT, CFMAX, CFR, CWH, BETAET };
  HY_HD static float init(int) { return 0.001f; }

  HY_HD static float step(float* s, const float* p, float P, float E, float Tm, bool has_bet) {
    const bool cold = Tm < p[TT];
    const float rain = cold ? 0.f : P, snowf = cold ? P : 0.f;
    float SP = s[0] + snowf;
    float melt = fminf(fmaxf(p[CFMAX] * (Tm - p[TT]), 0.f), SP);
    float MW = s[1] + melt;
    SP = SP - melt;
    float refr = fminf(fmaxf(p[CFR] * p[CFMAX] * (p[TT] - Tm), 0.f), MW);
    SP = SP + refr;
    MW = MW - refr;
    const float tosoil = fmaxf(MW - p[CWH] * SP, 0.f);
    MW = MW - tosoil;
    const float wetn = fminf(fmaxf(powf(s[2] / p[FC], p[BETA]), 0.f), 1.f);
    const float recharge = (rain + tosoil) * wetn;
    float SM = s[2] + rain + tosoil - recharge;
    const float excess = fmaxf(SM - p[FC], 0.f);
    SM = SM - excess;
    const float re = SM / (p[LP] * p[FC]);
    const float e0 = has_bet ? powf(re, p[BETAET]) : re;
    const float ef = fminf(fmaxf(e0, 0.f), 1.f);
    const float eta = fminf(SM, E * ef);
    SM = fmaxf(SM - eta, 1e-5f);
    float SUZ = s[3] + recharge + excess;
    const float perc = fminf(SUZ, p[PERC]);
    SUZ = SUZ - perc;
    const float Q0 = p[K0] * fmaxf(SUZ - p[UZL], 0.f);
    SUZ = SUZ - Q0;
    const float Q1 = p[K1] * SUZ;
    SUZ = SUZ - Q1;
    const float SLZ1 = s[4] + perc;
    const float Q2 = p[K2] * SLZ1;
    s[0] = SP; s[1] = MW; s[2] = SM; s[3] = SUZ; s[4] = SLZ1 - Q2;
    return Q0 + Q1 + Q2;
  }

  HY_HD static void adjoint(const float* s, const float* p, float P, float E, float Tm,
                                                 bool has_bet, float* g, float gq, float* gp) {
    // ---- recompute forward
    const bool cold = Tm < p[TT];
    const float rain = cold ? 0.f : P, snowf = cold ? P : 0.f;
    const float SP1 = s[0] + snowf;
    const float dT = Tm - p[TT];
    const float m0 = p[CFMAX] * dT;
    const float m1 = fmaxf(m0, 0.f);
    const float melt = fminf(m1, SP1);
    const float MW1 = s[1] + melt;
    const float SP2 = SP1 - melt;
    const float cc = p[CFR] * p[CFMAX];
    const float r0 = cc * (p[TT] - Tm);
    const float r1 = fmaxf(r0, 0.f);
    const float refr = fminf(r1, MW1);
    const float SP3 = SP2 + refr;
    const float MW2 = MW1 - refr;
    const float ts0 = MW2 - p[CWH] * SP3;
    const float tosoil = fmaxf(ts0, 0.f);
    const float rw = s[2] / p[FC];
    const float w0 = powf(rw, p[BETA]);
    const float wetn = fminf(fmaxf(w0, 0.f), 1.f);
    const float inflow = rain + tosoil;
    const float recharge = inflow * wetn;
    const float SM1 = s[2] + rain + tosoil - recharge;
    const float ex0 = SM1 - p[FC];
    const float excess = fmaxf(ex0, 0.f);
    const float SM2 = SM1 - excess;
    const float lf = p[LP] * p[FC];
    const float re = SM2 / lf;
    const float e0 = has_bet ? powf(re, p[BETAET]) : re;
    const float ef = fminf(fmaxf(e0, 0.f), 1.f);
    const float et0 = E * ef;
    const float eta = fminf(SM2, et0);
    const float sm3 = SM2 - eta;
    const float SUZ1 = s[3] + recharge + excess;
    const float perc = fminf(SUZ1, p[PERC]);
    const float SUZ2 = SUZ1 - perc;
    const float u0 = SUZ2 - p[UZL];
    const float c0 = fmaxf(u0, 0.f);
    const float SUZ3 = SUZ2 - p[K0] * c0;
    const float SLZ1 = s[4] + perc;
    // ---- lower zone
    const float gQ2 = gq - g[4];
    const float gSLZ1 = g[4] + gQ2 * p[K2];
    gp[K2] = gQ2 * SLZ1;
    g[4] = gSLZ1;
    // ---- upper zone
    const float gQ1 = gq - g[3];
    const float gSUZ3 = g[3] + gQ1 * p[K1];
    gp[K1] = gQ1 * SUZ3;
    const float gQ0 = gq - gSUZ3;
    gp[K0] = gQ0 * c0;
    const float gu0 = gQ0 * p[K0] * w_ge(u0, 0.f);
    const float gSUZ2 = gSUZ3 + gu0;
    gp[UZL] = -gu0;
    const float gperc = gSLZ1 - gSUZ2;
    const float gSUZ1 = gSUZ2 + gperc * w_lt(SUZ1, p[PERC]);
    gp[PERC] = gperc * w_lt(p[PERC], SUZ1);
    g[3] = gSUZ1;
    float gexcess = gSUZ1;
    // ---- soil moisture / evaporation
    const float gsm3 = g[2] * w_ge(sm3, 1e-5f);
    float gSM2 = gsm3 - gsm3 * w_lt(SM2, et0);        // sm3 = SM2 - eta, eta = min(SM2, et0)
    const float get0 = -gsm3 * w_lt(et0, SM2);
    const float ge0 = get0 * E * ((e0 >= 0.f && e0 <= 1.f) ? 1.f : 0.f);
    float gre;
    if (has_bet) {
      gre = ge0 * pow_dbase(re, p[BETAET]);
      gp[BETAET] = ge0 * pow_dexp(re, p[BETAET], e0);
    } else {
      gre = ge0;
      gp[BETAET] = 0.f;
    }
    gSM2 += gre / lf;
    const float glf = -gre * SM2 / (lf * lf);
    gp[LP] = glf * p[FC];
    float gFC = glf * p[LP];
    float gSM1 = gSM2;
    gexcess -= gSM2;
    const float gex0 = gexcess * w_ge(ex0, 0.f);
    gSM1 += gex0;
    gFC -= gex0;
    float gtosoil = gSM1;
    const float grecharge = gSUZ1 - gSM1;
    gtosoil += grecharge * wetn;
    const float gw0 = grecharge * inflow * ((w0 >= 0.f && w0 <= 1.f) ? 1.f : 0.f);
    const float grw = gw0 * pow_dbase(rw, p[BETA]);
    gp[BETA] = gw0 * pow_dexp(rw, p[BETA], w0);
    g[2] = gSM1 + grw / p[FC];
    gFC -= grw * s[2] / (p[FC] * p[FC]);
    gp[FC] = gFC;
    // ---- snow routine
    gtosoil -= g[1];
    const float gts0 = gtosoil * w_ge(ts0, 0.f);
    const float gMW2 = g[1] + gts0;
    gp[CWH] = -gts0 * SP3;
    const float gSP3 = g[0] - gts0 * p[CWH];
    const float grefr = gSP3 - gMW2;
    const float gr0 = grefr * w_lt(r1, MW1) * w_ge(r0, 0.f);
    const float gMW1 = gMW2 + grefr * w_lt(MW1, r1);
    gp[CFR] = gr0 * p[CFMAX] * (p[TT] - Tm);
    float gCFMAX = gr0 * p[CFR] * (p[TT] - Tm);
    float gTT = gr0 * cc;
    g[1] = gMW1;
    const float gmelt = gMW1 - gSP3;
    const float gm0 = gmelt * w_lt(m1, SP1) * w_ge(m0, 0.f);
    g[0] = gSP3 + gmelt * w_lt(SP1, m1);
    gCFMAX += gm0 * dT;
    gTT -= gm0 * p[CFMAX];
    gp[CFMAX] = gCFMAX;
    gp[TT] = gTT;
  }
};

// sub-gradient of max(0, x) written as torch.maximum(zero, x): ties split 1/2 - 1/2
HY_HD float w_pos(float x) { return x > 0.f ? 1.f : (x == 0.f ? 0.5f : 0.f); }

// ------------------------------------------------------------------------------------------------
// linear reservoir (modelzoo/linear_reservoir.py:77-92): one storage, uses P and ET only
// ------------------------------------------------------------------------------------------------
struct LinRes {
  static constexpr int NP = 2;
  static constexpr int NS = 1;   // si
  enum { KI, AUXET };
  HY_HD static float init(int) { return 0.001f; }

  HY_HD static float step(float* s, const float* p, float P, float E, float, bool) {
    const float d = (s[0] + P) - E * p[AUXET];
    const float s1 = fmaxf(0.f, d);
    const float qi = s1 * p[KI];
    s[0] = s1 - qi;
    return qi;
  }

  HY_HD static void adjoint(const float* s, const float* p, float P, float E, float, bool, float* g, float gq,
                            float* gp) {
    const float d = (s[0] + P) - E * p[AUXET];
    const float s1 = fmaxf(0.f, d);
    const float g_qi = gq - g[0];
    const float g_s1 = g[0] + g_qi * p[KI];
    gp[KI] = g_qi * s1;
    const float g_d = g_s1 * w_pos(d);
    gp[AUXET] = -g_d * E;
    g[0] = g_d;
  }
};

// ------------------------------------------------------------------------------------------------
// NonSense (modelzoo/nonsense.py:82-168): snow -> baseflow -> interflow -> unsaturated zone -> outlet.
// Reservoir constants divide (qb = sb / kb); the outflow is the unsaturated-zone outflow alone.
// ------------------------------------------------------------------------------------------------
struct NonSense {
  static constexpr int NP = 5;
  static constexpr int NS = 4;   // ss su si sb  (nonsense.py:172)
  enum { DD, SUMAX, BETA, KI, KB };
  enum { SS, SU, SI, SB };
  static constexpr float klu = 0.90f;
  HY_HD static float init(int k) { return k == SS ? 0.f : (k == SU ? 5.f : (k == SI ? 10.f : 15.f)); }

  HY_HD static float step(float* s, const float* p, float P, float E, float Tm, bool) {
    const bool cold = Tm < 0.f;
    const float mp = cold ? 0.f : Tm * p[DD];
    const float rain = cold ? 0.f : P, snowf = cold ? P : 0.f;
    const float qs = fminf(s[SS], mp);
    s[SS] = s[SS] - qs + snowf;
    const float qsp = qs + rain;
    const float sb1 = s[SB] + qsp;
    const float qb = sb1 / p[KB];
    s[SB] = sb1 - qb;
    const float si1 = s[SI] + qb;
    const float qi = si1 / p[KI];
    s[SI] = si1 - qi;
    const float psi = powf(s[SU] / p[SUMAX], p[BETA]);
    const float su_t = s[SU] + qi * (1.f - psi);
    const float su1 = fminf(su_t, p[SUMAX]);
    const float qu = qi * psi + fmaxf(0.f, su_t - p[SUMAX]);
    const float kth = (su1 <= 0.8f * p[SUMAX]) ? su1 / p[SUMAX] : 1.f;
    s[SU] = fmaxf(0.f, su1 - E * klu * kth);
    return qu;
  }

  HY_HD static void adjoint(const float* s, const float* p, float P, float E, float Tm, bool, float* g, float gq,
                            float* gp) {
    // ---- recompute the forward intermediates
    const bool cold = Tm < 0.f;
    const float mp = cold ? 0.f : Tm * p[DD];
    const float rain = cold ? 0.f : P;
    const float qs = fminf(s[SS], mp);
    const float sb1 = s[SB] + (qs + rain);
    const float qb = sb1 / p[KB];
    const float si1 = s[SI] + qb;
    const float qi = si1 / p[KI];
    const float r = s[SU] / p[SUMAX];
    const float psi = powf(r, p[BETA]);
    const float su_t = s[SU] + qi * (1.f - psi);
    const float su1 = fminf(su_t, p[SUMAX]);
    const float ovx = su_t - p[SUMAX];
    const bool wet = su1 <= 0.8f * p[SUMAX];
    const float kth = wet ? su1 / p[SUMAX] : 1.f;
    const float d = su1 - E * klu * kth;
    // ---- evapotranspiration: su2 = max(0, d); the wilting point enters the comparison only
    const float g_d = g[SU] * w_pos(d);
    float g_su1 = g_d;
    float g_smax = 0.f;
    if (wet) {
      const float g_kth = -g_d * E * klu;
      g_su1 += g_kth / p[SUMAX];
      g_smax -= g_kth * su1 / (p[SUMAX] * p[SUMAX]);
    }
    // ---- unsaturated zone (its outflow is the model outflow)
    float g_qi = gq * psi;
    float g_psi = gq * qi;
    const float g_ovx = gq * w_pos(ovx);
    float g_su_t = g_ovx;
    g_smax -= g_ovx;
    g_su_t += g_su1 * w_lt(su_t, p[SUMAX]);
    g_smax += g_su1 * w_lt(p[SUMAX], su_t);
    float g_su = g_su_t;
    g_qi += g_su_t * (1.f - psi);
    g_psi -= g_su_t * qi;
    const float g_r = g_psi * pow_dbase(r, p[BETA]);
    gp[BETA] = g_psi * pow_dexp(r, p[BETA], psi);
    g_su += g_r / p[SUMAX];
    g_smax -= g_r * s[SU] / (p[SUMAX] * p[SUMAX]);
    gp[SUMAX] = g_smax;
    g[SU] = g_su;
    // ---- interflow: qi = si1 / ki, si' = si1 - qi
    g_qi -= g[SI];
    const float g_si1 = g[SI] + g_qi / p[KI];
    gp[KI] = -g_qi * (qi / p[KI]);      // torch: -grad * ((self / other) / other)
    g[SI] = g_si1;
    // ---- baseflow: qb = sb1 / kb, sb' = sb1 - qb
    const float g_qb = g_si1 - g[SB];
    const float g_sb1 = g[SB] + g_qb / p[KB];
    gp[KB] = -g_qb * (qb / p[KB]);
    g[SB] = g_sb1;
    // ---- snow
    const float g_qs = g_sb1 - g[SS];
    g[SS] = g[SS] + g_qs * w_lt(s[SS], mp);
    const float g_mp = g_qs * w_lt(mp, s[SS]);
    gp[DD] = cold ? 0.f : g_mp * Tm;
  }
};


}  // namespace hy2dl
