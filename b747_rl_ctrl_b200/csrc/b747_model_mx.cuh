// b747_model_mx.cuh -- throughput ("f32") path of the B747 model step: one environment per thread,
// K fused RK4 substeps per launch with the state held in registers.
//
// Same block diagram as b747_model_f64.cuh (model_simple_step dll@0x16d0),
// re-formulated for the FP32/XU pipes of an SM:
//  * aerodynamics, atmosphere, trigonometry, table look-ups, actuator and PIDs in float32;
//  * the attitude is carried as the pitch angle itself (theta' = wz) instead of the quaternion the
//    DLL integrates and re-normalises: for a pitch-only rotation the two are the same ODE
//    (q = (cos th/2, 0, 0, sin th/2)), and RK4 on either differs by O((wz h)^5) ~ 1e-20;
//    sin/cos of the body rotation then need no asin;
//  * the pitch-error chain -- theta, dvartheta = ref - theta, its integral, ITSE and the two
//    finite-difference Derivative blocks -- is kept in float64, because the observation and the
//    CLASSIC reward difference it over 5..10 ms (hard part 3); B200 issues DFMA at half
//    the FFMA rate, so these ~10 operations per pass are nearly free;
//  * every integrator state is accumulated in float64 across steps (y += h/6 * sum), the stage
//    values inside a step are float32;
//  * sin(alpha), cos(alpha) come from the body-axis velocity components (-wb/V, ub/V);
//    density uses exp2/log2; table fractions use precomputed reciprocal breakpoint spacings;
//    breakpoint searches are branch-free compare chains against immediates;
//  * the transport delay (0.03 s = 3 steps), the Derivative and rate-limiter stamps are resolved
//    from the integer tick, so no time-stamp arithmetic is left in floating point;
//  * states that no observation/reward reads (ITAE, IAE, ISE; x and the altitude-loop PID unless the
//    configuration needs them) are not integrated.
#pragma once
#include <math.h>

#include "b747_common.cuh"

namespace b747 {

__host__ __device__ constexpr double Pc(int i) {
  constexpr double a[kNP] = B747_P_INIT;
  return a[i];
}
template <int I>
struct PF { static constexpr float v = (float)Pc(I); };
#define PCF(i) (PF<(i)>::v)

// float copies of the uniform tunables + folded constants (built on the host, b747_kernels_f32.cu)
struct MP32 {
  float PID_SS[4], PID_CS[4];
  float P, g, inv_m0, half_S, half_Sc_over_Iz, use_RP, use_RL, use_PID_SS;
};

struct RegsMx {
  double h, th, Vx, Vy, wz, ssi, ssf, dvi, itse;  // integrator states, accumulated in float64
  double csi, csf, x;                              // GEN only
  double d1_u;                                     // dvartheta at the last major step
  double vref, href, ep_return;
  double vartheta;                                 // current pitch reference (the DLL's `vartheta` param); registers only
  float df_x, df_y, rl_prev, deltaz, uh[4], sig_upid, d2_u, sig_vzh, tf_tp;
  float sumA[5];                                   // 1 + aero_err[k]
  double oscA[3], oscf[3];
  int tick, flags;
  uint32_t ep_idx;
};

struct PassMx {
  double dv, dv_dt;
  float th, V, alpha, Mach, CXa, CYa, mz, K_alpha, dCm, U_com, U_com_PID, deltaz_RP, vartheta_zh, td, rl_out;
  bool and_ss, and_cs;
};

// stage-4 (predictor) state values = what the DLL's `state`/integral signals show after a step
struct Stage4Mx { float h, Vx, Vy, wz, x; double dvi, itse; };

__device__ __forceinline__ float satf(float u, float lo, float hi) { return fminf(fmaxf(u, lo), hi); }
__device__ __forceinline__ int sgnf(float x) { return (x > 0.f) - (x < 0.f); }

// index of the breakpoint interval containing u (look2_binlx prelookup semantics: clamp to the end
// intervals, which extrapolate) + fraction; breakpoints are immediates, spacing reciprocals in smem.
template <int B, int M>
__device__ __forceinline__ int prelook32(float u, const float* __restrict__ sP, const float* __restrict__ sR, float& frac) {
  int idx = 0;
#pragma unroll
  for (int j = 1; j < M; j++) idx += (u >= (float)Pc(B + j)) ? 1 : 0;
  frac = (u - sP[B + idx]) * sR[B + idx];
  return idx;
}

__device__ __forceinline__ float bilin32(const float* __restrict__ tab, int i0, float f0, int i1, float f1, int stride) {
  const float* p = tab + i1 * stride + i0;
  float a = p[0], b = p[1], c = p[stride], d = p[stride + 1];
  float yL = fmaf(b - a, f0, a);
  float yR = fmaf(d - c, f0, c);
  return fmaf(yR - yL, f1, yL);
}

// One pass over the diagram at a stage state.  dt_last = t - (stamp of the last update) as an exact
// constant (0.01 or 0.005), stage: 0 major, 1/2 half steps, 3 full step.
template <bool GEN>
__device__ __forceinline__ void pass32(const float* __restrict__ sP, const float* __restrict__ sR, const MP32& mp,
                                       const DevCfg& c, int stage, int n, double th_d, double t_d, float h, double h_d,
                                       float Vx, float Vy, float wz, float ssi, float ssf, double csi, double csf,
                                       RegsMx& r, bool& memout_ss, bool& memout_cs, PassMx& o, float& f_h, float& f_Vx,
                                       float& f_Vy, float& f_wz, float& f_ssi, float& f_ssf, double& f_csi,
                                       double& f_csf, double& f_itse) {
  const bool major = stage == 0;
  // attitude: the DLL's th = asin(sin(theta)) folds beyond +-90 deg; keep that (rare) behaviour
  float thf = (float)th_d;
  float sn, cs;
  sincosf(thf, &sn, &cs);
  double th_fold = th_d;
  if (fabsf(thf) > 1.57079632679f) {  // rare: keep the DLL's principal-value pitch
    th_fold = asin(sin(th_d));
    thf = (float)th_fold;
    cs = fabsf(cs);
  }
  o.th = thf;
  float ub = fmaf(cs, Vx, sn * Vy);
  float wb = fmaf(cs, Vy, -sn * Vx);
  float V2 = fmaf(ub, ub, wb * wb);
  float rV = rsqrtf(V2);
  float V = V2 * rV;
  float alpha = -atan2f(wb, ub);
  float sa = -wb * rV, ca = ub * rV;
  o.V = V; o.alpha = alpha;
  // ISA atmosphere
  float hs = fminf(fmaxf(h, PCF(18)), PCF(17));
  float T = fmaf(-hs, PCF(19), PCF(16));
  float Mach = V * rsqrtf(T * PCF(20));
  float ad = alpha * PCF(21);
  o.Mach = Mach;
  // look-ups (Mach axis of CYa and mz share breakpoints P[42..45] == P[276..279])
  float fMa, fMb, fMc, fAa, fAb, fAc, fH, fC;
  int iMa = prelook32<42, 3>(Mach, sP, sR, fMa);
  int iAa = prelook32<46, 4>(ad, sP, sR, fAa);
  float CYa = bilin32(sP + 22, iMa, fMa, iAa, fAa, 4) * r.sumA[1];
  int iMb = prelook32<108, 3>(Mach, sP, sR, fMb);
  int iC = prelook32<112, 13>(CYa, sP, sR, fC);
  float CXa = bilin32(sP + 52, iMb, fMb, iC, fC, 4) * r.sumA[0];
  o.CYa = CYa; o.CXa = CXa;
  int iH = prelook32<201, 4>(h, sP, sR, fH);
  int iMc = prelook32<206, 9>(Mach, sP, sR, fMc);
  float dCm = bilin32(sP + 151, iH, fH, iMc, fMc, 5) * r.sumA[3];
  int iAc = prelook32<225, 6>(ad, sP, sR, fAc);
  float Ka = fmaf(sP[218 + iAc + 1] - sP[218 + iAc], fAc, sP[218 + iAc]) * r.sumA[4];
  int iAb = prelook32<280, 10>(ad, sP, sR, fAb);
  float mz = bilin32(sP + 232, iMa, fMa, iAb, fAb, 4) * r.sumA[2];
  o.dCm = dCm; o.K_alpha = Ka; o.mz = mz;
  // density: rho0 * (T/T0)^(g/(LR)-1) * exp(g/R * sat(11000-h) / T)
  float Tr = T * PCF(127);
  float rho = PCF(129) * exp2f((PCF(128) - 1.0f) * log2f(Tr));
  float dh = PCF(130) - h;
  if (dh < PCF(131)) {  // above the tropopause (rare for this envelope)
    float xs = fmaxf(dh, PCF(132));
    rho *= __expf(xs * PCF(133) / T);
  }
  float rV2 = rho * V2;
  float qS = rV2 * mp.half_S;
  float mD = PCF(126) * CXa * qS;
  float Lf = qS * CYa;
  float Fx = fmaf(mD, ca, fmaf(sa, Lf, mp.P));
  float Fy = fmaf(ca, Lf, -mD * sa);
  // actuator: transport delay (3 steps) -> discrete filter (every 5th tick) -> rate limiter -> saturation
  float td;
  if (stage == 0) td = n > 3 ? r.uh[1] : PCF(137);
  else if (stage == 3) td = n >= 3 ? r.uh[2] : PCF(137);
  else td = n >= 3 ? 0.5f * (r.uh[1] + r.uh[2]) : PCF(137);
  o.td = td;
  if (major && (n % 5) == 0) r.df_y = fmaf(r.df_x, PCF(140), PCF(141) * td);
  float yv = r.df_y;
  if (!(major && n == 0)) {
    const float dT = (stage == 1 || stage == 2) ? 0.005f : 0.01f;
    float rate = yv - r.rl_prev;
    yv = r.rl_prev + fminf(fmaxf(rate, dT * PCF(143)), dT * PCF(142));
  }
  o.rl_out = yv;
  o.deltaz_RP = satf(yv, PCF(145), PCF(144));
  // СУ PID (altitude loop) -- only integrated when the configuration can close it.  Its output is the
  // pitch reference, which the СС PID differentiates with a gain of Kd*N ~ 390, so the whole
  // altitude-error chain is float64 (a float32 altitude quantises it at ~1e-5 rad).
  float use_cs = 0.f;
  double cs_pre = 0.0, cs_d = 0.0, e_h = 0.0, vzh_d = 0.0;
  if (GEN) {
    use_cs = (r.flags & FL_USE_CTRL) ? 1.f : 0.f;
    e_h = r.href - h_d;  // h_zh - h
    cs_d = (e_h * c.mp.PID_CS[2] - csf) * c.mp.PID_CS[3];
    cs_pre = e_h * c.mp.PID_CS[0] + csi + cs_d;
    vzh_d = fmin(fmax(cs_pre, Pc(4)), Pc(6));
    o.vartheta_zh = (float)vzh_d;
  } else {
    o.vartheta_zh = 0.f;
  }
  // pitch error in float64
  double vref_d = (GEN && use_cs >= PCF(146)) ? vzh_d : r.vartheta;
  double dv_d = vref_d - th_fold;
  o.dv = dv_d;
  float dv = (float)dv_d;
  // СС PID
  float ss_d = (dv * mp.PID_SS[2] - ssf) * mp.PID_SS[3];
  float ss_pre = fmaf(dv, mp.PID_SS[0], ssi) + ss_d;
  o.U_com_PID = satf(ss_pre, PCF(5), PCF(7));
  if (mp.use_RL >= PCF(148)) o.U_com = PCF(147) > fabsf(o.U_com_PID) ? 0.f : o.U_com_PID;
  else o.U_com = mp.use_PID_SS >= PCF(9) ? o.U_com_PID : r.deltaz;
  float ax = (Fx * cs - sn * Fy) * mp.inv_m0;
  float ay = fmaf(fmaf(Fy, cs, Fx * sn), mp.inv_m0, -mp.g);
  float dze = mp.use_RP >= PCF(149) ? o.deltaz_RP : o.U_com;
  float Cm = fmaf(PCF(217) * dCm * Ka, dze * PCF(150), mz);
  float wzd = Cm * (rV2 * mp.half_Sc_over_Iz);
  // clamping anti-windup (СС)
  float dz = ss_pre - satf(ss_pre, PCF(5), PCF(7));
  float ss_i = mp.PID_SS[1] * dv;
  o.and_ss = (ss_pre * PCF(291) != dz) && (sgnf(dz) == sgnf(ss_i));
  if (major) memout_ss = (r.flags & FL_MEM_SS) != 0;
  if (memout_ss) ss_i = PCF(10);
  f_h = Vy; f_Vx = ax; f_Vy = ay; f_wz = wzd; f_ssi = ss_i; f_ssf = ss_d;
  f_itse = dv_d * dv_d * t_d;
  if (GEN) {
    double dzc = cs_pre - vzh_d;
    double cs_i = e_h * c.mp.PID_CS[1];
    o.and_cs = (cs_pre * Pc(292) != dzc) && (((dzc > 0.0) - (dzc < 0.0)) == ((cs_i > 0.0) - (cs_i < 0.0)));
    if (major) memout_cs = (r.flags & FL_MEM_CS) != 0;
    if (memout_cs) cs_i = Pc(11);
    f_csi = cs_i; f_csf = cs_d;
  } else {
    o.and_cs = false; f_csi = 0.0; f_csf = 0.0;
  }
}

// model_simple_step in the mixed formulation.  On return r holds the post-update state, `o` the
// stage-4 pass and s4 the stage-4 (predictor) state values.
template <bool GEN>
__device__ __forceinline__ void model_step32(const float* __restrict__ sP, const float* __restrict__ sR, const MP32& mp,
                                             const DevCfg& c, RegsMx& r, PassMx& o, Stage4Mx& s4, bool want_x) {
  const int n = r.tick;
  const double t0 = (double)n * kH;
  const float hh = (float)kH, hhalf = 0.5f * (float)kH;
  // float32 copies of the accumulated state for the stage evaluations
  const float y_h = (float)r.h, y_Vx = (float)r.Vx, y_Vy = (float)r.Vy, y_wz = (float)r.wz, y_ssi = (float)r.ssi,
              y_ssf = (float)r.ssf;
  float X_h = y_h, X_Vx = y_Vx, X_Vy = y_Vy, X_wz = y_wz, X_ssi = y_ssi, X_ssf = y_ssf;
  double X_th = r.th, Xd_h = r.h, X_csi = r.csi, X_csf = r.csf;
  float a_h = 0, a_Vx = 0, a_Vy = 0, a_wz = 0, a_ssi = 0, a_ssf = 0, a_th = 0, a_x = 0;
  double a_dvi = 0, a_itse = 0, a_csi = 0, a_csf = 0;
  bool memout_ss = false, memout_cs = false;
  float u_n = 0.f;
#pragma unroll 1
  for (int s = 0; s < 4; s++) {
    const double t_d = s == 0 ? t0 : (s == 3 ? (double)(n + 1) * kH : t0 + 0.5 * kH);
    float f_h, f_Vx, f_Vy, f_wz, f_ssi, f_ssf;
    double f_itse, f_csi, f_csf;
    pass32<GEN>(sP, sR, mp, c, s, n, X_th, t_d, X_h, Xd_h, X_Vx, X_Vy, X_wz, X_ssi, X_ssf, X_csi, X_csf, r, memout_ss,
                memout_cs, o, f_h, f_Vx, f_Vy, f_wz, f_ssi, f_ssf, f_csi, f_csf, f_itse);
    if (s == 0) {
      // update(): discrete filter, rate-limiter memory, Memory blocks, Derivative history, delay push
      if ((n % 5) == 0) r.df_x = fmaf(PCF(138), r.df_x, PCF(139) * o.td);
      r.rl_prev = o.rl_out;
      r.flags = (r.flags & ~(FL_MEM_SS | FL_MEM_CS)) | (o.and_ss ? FL_MEM_SS : 0) | (o.and_cs ? FL_MEM_CS : 0);
      const double dvdt_major = n >= 1 ? (o.dv - r.d1_u) * 100.0 : 0.0;
      r.d1_u = o.dv;
      r.d2_u = (float)dvdt_major;
      u_n = o.U_com;
    }
    const float w = (s == 0 || s == 3) ? 1.f : 2.f;
    a_h = fmaf(w, f_h, a_h); a_Vx = fmaf(w, f_Vx, a_Vx); a_Vy = fmaf(w, f_Vy, a_Vy); a_wz = fmaf(w, f_wz, a_wz);
    a_ssi = fmaf(w, f_ssi, a_ssi); a_ssf = fmaf(w, f_ssf, a_ssf); a_th = fmaf(w, X_wz, a_th);
    if (GEN) { a_csi = fma((double)w, f_csi, a_csi); a_csf = fma((double)w, f_csf, a_csf); }
    if (want_x) a_x = fmaf(w, X_Vx, a_x);
    a_dvi = fma((double)w, o.dv, a_dvi);
    a_itse = fma((double)w, f_itse, a_itse);
    if (s < 3) {
      const float cf = (s == 2) ? hh : hhalf;
      if (s == 2) {  // integral / position signals at stage 4 = y + h*f2
        s4.dvi = fma((double)hh, o.dv, r.dvi);
        s4.itse = fma((double)hh, f_itse, r.itse);
        s4.x = want_x ? (float)r.x + hh * X_Vx : 0.f;
      }
      X_th = fma((double)cf, (double)X_wz, r.th);  // theta' = wz (stage value)
      X_h = fmaf(cf, f_h, y_h); X_Vx = fmaf(cf, f_Vx, y_Vx); X_Vy = fmaf(cf, f_Vy, y_Vy); X_wz = fmaf(cf, f_wz, y_wz);
      X_ssi = fmaf(cf, f_ssi, y_ssi); X_ssf = fmaf(cf, f_ssf, y_ssf);
      if (GEN) {
        X_csi = fma((double)cf, f_csi, r.csi); X_csf = fma((double)cf, f_csf, r.csf);
        Xd_h = fma((double)cf, (double)f_h, r.h);
      }
    }
  }
  s4.h = X_h; s4.Vx = X_Vx; s4.Vy = X_Vy; s4.wz = X_wz;
  const float h6 = (float)(kH / 6.0);
  const double h6d = kH / 6.0;
  r.h += (double)(h6 * a_h); r.Vx += (double)(h6 * a_Vx); r.Vy += (double)(h6 * a_Vy); r.wz += (double)(h6 * a_wz);
  r.ssi += (double)(h6 * a_ssi); r.ssf += (double)(h6 * a_ssf); r.th += (double)(h6 * a_th);
  if (GEN) { r.csi = fma(h6d, a_csi, r.csi); r.csf = fma(h6d, a_csf, r.csf); }
  if (want_x) r.x += (double)(h6 * a_x);
  r.dvi = fma(h6d, a_dvi, r.dvi);
  r.itse = fma(h6d, a_itse, r.itse);
  r.uh[0] = r.uh[1]; r.uh[1] = r.uh[2]; r.uh[2] = r.uh[3]; r.uh[3] = u_n;
  r.tick = n + 1;
}

// model_simple_initialize + Model.initialize (core/model.py:238-244)
__device__ __forceinline__ void model_init32(const double s0[6], RegsMx& r) {
  r.x = s0[0]; r.h = s0[1]; r.th = s0[4]; r.Vx = s0[2]; r.Vy = s0[3]; r.wz = s0[5];
  r.csi = Pc(2); r.csf = Pc(0); r.ssi = Pc(3); r.ssf = Pc(1);
  r.dvi = Pc(293); r.itse = Pc(297);
  r.df_x = PCF(8); r.df_y = 0.f; r.rl_prev = 0.f;
  r.uh[0] = r.uh[1] = r.uh[2] = r.uh[3] = 0.f;
  r.d1_u = 0.0; r.d2_u = 0.f;
  r.tick = 0;
  r.flags &= ~(FL_MEM_SS | FL_MEM_CS);
  r.sig_upid = 0.f; r.sig_vzh = 0.f;
  r.deltaz = 0.f;
}

}  // namespace b747
