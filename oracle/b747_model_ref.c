/*
 * b747_model_ref.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * float64 restatement of the Simulink-Coder ERT model in the reference's
 * core/model_simple_win64.dll, block by block in the DLL's evaluation order
 * (addresses are RVAs in that DLL).  Pinned against the
 * DLL's own machine code by tests/test_oracle_vs_dll.py.
 */
#include <math.h>
#include <string.h>

#include "../include/b747_params.h"
#include "b747_oracle.h"

static const double PB[B747_NP] = B747_P_INIT; /* model_simple_P, dll@0x24640 */
static const uint32_t MAXIDX[8] = B747_MAXIDX_INIT;

#define H_STEP 0.01 /* fixed step (stepSize0, dll@0x25f10) */

void b747o_model_defaults(b747o_model *m) {
  static const double pid_cs[4] = B747_DEF_PID_CS, pid_ss[4] = B747_DEF_PID_SS, s0[6] = B747_DEF_STATE0;
  memset(m, 0, sizeof *m);
  memcpy(m->PID_CS, pid_cs, sizeof pid_cs);
  memcpy(m->PID_SS, pid_ss, sizeof pid_ss);
  memcpy(m->state0, s0, sizeof s0);
  m->h_zh = B747_DEF_H_ZH; m->use_RP = B747_DEF_USE_RP; m->use_PID_SS = B747_DEF_USE_PID_SS;
  m->use_PID_CS = B747_DEF_USE_PID_CS; m->deltaz = B747_DEF_DELTAZ; m->vartheta = B747_DEF_VARTHETA;
  m->P = B747_DEF_P; m->Iz = B747_DEF_IZ; m->S = B747_DEF_S; m->c_ = B747_DEF_C; m->g = B747_DEF_G;
  m->m0 = B747_DEF_M0; m->use_RL = B747_DEF_USE_RL;
}

/* model_simple_initialize, dll@0x12a0-0x16c2 */
void b747o_model_initialize(b747o_model *m) {
  /* all signals, B, DW := 0 */
  memset(m->state, 0, sizeof m->state);
  m->sim_time = m->vartheta_zh = m->U_com_PID = m->CXa = m->CYa = m->mz = m->K_alpha = m->dCm_ddeltaz = 0;
  m->U_com = m->deltaz_RP = m->dvartheta = m->dvartheta_int = m->dvartheta_dt = m->dvartheta_dt_dt = 0;
  m->TAE = m->ITAE = m->TSE = m->ITSE = m->AE = m->IAE = m->SE = m->ISE = m->alpha = m->V = m->Mach = 0;
  memset(m->dX, 0, sizeof m->dX);
  m->t = 0; m->tick = 0; m->tid2 = 0; m->first = 1;
  const double *s0 = m->state0;
  double *X = m->X;
  X[0] = s0[0]; X[1] = s0[1];
  X[2] = cos(s0[4] / 2.0); X[3] = 0.0; X[4] = 0.0; X[5] = sin(s0[4] / 2.0);
  X[6] = s0[2]; X[7] = s0[3]; X[8] = s0[5];
  X[9] = PB[2]; X[10] = PB[0]; X[11] = PB[3]; X[12] = PB[1];
  for (int i = 0; i < 5; i++) X[13 + i] = PB[293 + i];
  m->ic_t = -INFINITY;
  m->df_x = PB[8]; m->df_y = 0;
  m->rl_prev = 0; m->rl_t = INFINITY; m->rl_out = 0;
  m->d1.tA = m->d1.tB = m->d2.tA = m->d2.tB = INFINITY;
  m->d1.uA = m->d1.uB = m->d2.uA = m->d2.uB = 0;
  memset(m->ring_u, 0, sizeof m->ring_u);
  memset(m->ring_t, 0, sizeof m->ring_t);
  m->ring_u[0] = PB[137]; m->ring_t[0] = 0.0;
  m->tail = m->head = m->last = 0; m->size = B747O_RING;
  m->mem_ss = m->mem_cs = m->and_ss = m->and_cs = m->memout_ss = m->memout_cs = 0;
  memset(m->sumA, 0, sizeof m->sumA);
  m->td = 0;
}

/* plook_binx-style prelookup used by look2_binlx (dll@0x1000): index + fraction,
 * linear extrapolation on both ends. */
static uint32_t prelookup(double u, const double *bp, uint32_t maxIndex, double *frac) {
  uint32_t iLeft;
  if (u <= bp[0]) {
    iLeft = 0;
    *frac = (u - bp[0]) / (bp[1] - bp[0]);
  } else if (u < bp[maxIndex]) {
    uint32_t bpIdx = maxIndex >> 1, iRght = maxIndex;
    iLeft = 0;
    while (iRght - iLeft > 1) {
      if (u < bp[bpIdx]) iRght = bpIdx; else iLeft = bpIdx;
      bpIdx = (iRght + iLeft) >> 1;
    }
    *frac = (u - bp[iLeft]) / (bp[iLeft + 1] - bp[iLeft]);
  } else {
    iLeft = maxIndex - 1;
    *frac = (u - bp[maxIndex - 1]) / (bp[maxIndex] - bp[maxIndex - 1]);
  }
  return iLeft;
}

static double look2(double u0, double u1, const double *bp0, const double *bp1, const double *tab,
                    const uint32_t *maxIndex, uint32_t stride) {
  double f0, f1;
  uint32_t i0 = prelookup(u0, bp0, maxIndex[0], &f0);
  uint32_t i1 = prelookup(u1, bp1, maxIndex[1], &f1);
  uint32_t o = i1 * stride + i0;
  double yL = (tab[o + 1] - tab[o]) * f0 + tab[o];
  o += stride;
  double yR = (tab[o + 1] - tab[o]) * f0 + tab[o];
  return (yR - yL) * f1 + yL;
}

static double look1(double u, const double *bp, const double *tab, uint32_t maxIndex) {
  double f;
  uint32_t i = prelookup(u, bp, maxIndex, &f);
  return (tab[i + 1] - tab[i]) * f + tab[i];
}

/* rt_TDelayInterpolate, dll@0x29e0 (continuous transport delay, linear interpolation) */
static double tdelay(b747o_model *m, double tMinusDelay, double tStart, double initOutput) {
  const double *tBuf = m->ring_t, *uBuf = m->ring_u;
  int bufSz = m->size, oldestIdx = m->tail, newIdx = m->head;
  if (newIdx == 0 && oldestIdx == 0 && tMinusDelay > tStart) return initOutput;
  if (tMinusDelay <= tStart) return initOutput;
  double t1, t2, u1, u2;
  if (tMinusDelay <= tBuf[oldestIdx]) {
    int tempIdx = oldestIdx + 1;
    if (oldestIdx == bufSz - 1) tempIdx = 0;
    t1 = tBuf[oldestIdx]; t2 = tBuf[tempIdx]; u1 = uBuf[oldestIdx]; u2 = uBuf[tempIdx];
  } else {
    int i = m->last;
    if (tBuf[i] < tMinusDelay) {
      while (tBuf[i] < tMinusDelay) {
        if (i == newIdx) break;
        i = (i < bufSz - 1) ? i + 1 : 0;
      }
    } else {
      while (tBuf[i] >= tMinusDelay) i = (i > 0) ? i - 1 : bufSz - 1;
      i = (i < bufSz - 1) ? i + 1 : 0;
    }
    m->last = i;
    if (i == 0) { t1 = tBuf[bufSz - 1]; u1 = uBuf[bufSz - 1]; } else { t1 = tBuf[i - 1]; u1 = uBuf[i - 1]; }
    t2 = tBuf[i]; u2 = uBuf[i];
  }
  if (t2 == t1) return (tMinusDelay >= t2) ? u2 : u1;
  double f1 = (t2 - tMinusDelay) / (t2 - t1);
  double f2 = 1.0 - f1;
  return f1 * u1 + f2 * u2;
}

static double deriv_out(const b747o_deriv *d, double u, double t) {
  if (d->tA >= t && d->tB >= t) return 0.0;
  double lt = d->tA, lu = d->uA;
  if (d->tA < d->tB) {
    if (d->tB < t) { lt = d->tB; lu = d->uB; }
  } else if (d->tA >= t) {
    lt = d->tB; lu = d->uB;
  }
  return (u - lu) / (t - lt);
}

static void deriv_upd(b747o_deriv *d, double u, double t) {
  if (d->tA == INFINITY) { d->tA = t; d->uA = u; }
  else if (d->tB == INFINITY) { d->tB = t; d->uB = u; }
  else if (d->tA < d->tB) { d->tA = t; d->uA = u; }
  else { d->tB = t; d->uB = u; }
}

static double sat(double u, double lo, double hi) { return u > hi ? hi : (u >= lo ? u : lo); }
static int sgn8(double x) { return x < 0.0 ? -1 : (x > 0.0 ? 1 : 0); }

static double rt_atan2(double u0, double u1) { /* rt_atan2d_snf */
  if (isnan(u0) || isnan(u1)) return NAN;
  if (isinf(u0) && isinf(u1)) return atan2(u0 > 0 ? 1.0 : -1.0, u1 > 0 ? 1.0 : -1.0);
  if (u1 == 0.0) return u0 > 0.0 ? M_PI / 2.0 : (u0 < 0.0 ? -(M_PI / 2.0) : 0.0);
  return atan2(u0, u1);
}

static double rt_pow(double u0, double u1) { /* rt_powd_snf, dll@0x3530 */
  if (isnan(u0) || isnan(u1)) return NAN;
  double a0 = fabs(u0), a1 = fabs(u1);
  if (isinf(u1)) {
    if (a0 == 1.0) return 1.0;
    if (a0 > 1.0) return u1 > 0.0 ? INFINITY : 0.0;
    return u1 > 0.0 ? 0.0 : INFINITY;
  }
  if (a1 == 0.0) return 1.0;
  if (a1 == 1.0) return u1 > 0.0 ? u0 : 1.0 / u0;
  if (u1 == 2.0) return u0 * u0;
  if (u1 == 0.5 && u0 >= 0.0) return sqrt(u0);
  if (u0 < 0.0 && u1 > floor(u1)) return NAN;
  return pow(u0, u1);
}

/* One pass over the block diagram (model_simple_step body, dll@0x176c-0x2711). */
static void outputs(b747o_model *m, int major) {
  const double *X = m->X;
  const double t = m->t;
  /* quaternion normalisation and pitch angle */
  double n = sqrt(X[2] * X[2] + X[3] * X[3] + X[4] * X[4] + X[5] * X[5]);
  double q3 = X[5] / n, q0 = X[2] / n, q1 = X[3] / n, q2 = X[4] / n;
  double th = asin((q1 * q2 + q3 * q0) * 2.0);
  /* IC block: emits state0 at the first time point */
  if (m->ic_t == -INFINITY || m->ic_t == t) {
    m->ic_t = t;
    memcpy(m->state, m->state0, sizeof m->state);
  } else {
    m->state[0] = X[0]; m->state[1] = X[1]; m->state[2] = X[6]; m->state[3] = X[7];
    m->state[4] = th; m->state[5] = X[8];
  }
  double sn = sin(th), cs = cos(th);
  double Vx = X[6], Vy = X[7];
  double ub = cs * Vx + sn * Vy;
  double wb = cs * Vy - sn * Vx;
  /* MATLAB-style scaled 2-norm of (ub, wb) */
  double scale = 3.3121686421112381E-170, y, a;
  a = fabs(ub);
  if (a > scale) { y = 1.0; scale = a; } else { double r = a / scale; y = r * r; }
  a = fabs(wb);
  if (a > scale) { double r = scale / a; y = y * r * r + 1.0; scale = a; } else { double r = a / scale; y += r * r; }
  double V = scale * sqrt(y);
  m->V = V;
  double alpha = -rt_atan2(wb, ub);
  m->alpha = alpha;
  /* ISA atmosphere */
  double h = X[1];
  double hs = h > PB[17] ? PB[17] : (h >= PB[18] ? h : PB[18]);
  double T = PB[16] - hs * PB[19];
  double asnd = sqrt(T * PB[20]);
  double ad = alpha * PB[21];
  double Mach = V / asnd;
  m->Mach = Mach;
  if (major) { m->sumA[1] = m->aero_err[1] + PB[51]; m->sumA[0] = m->aero_err[0] + PB[51]; }
  double CYa = look2(Mach, ad, PB + 42, PB + 46, PB + 22, MAXIDX + 0, 4) * m->sumA[1];
  double CXa = look2(Mach, CYa, PB + 108, PB + 112, PB + 52, MAXIDX + 2, 4) * m->sumA[0];
  m->CYa = CYa; m->CXa = CXa;
  double Tr = T * PB[127];
  double pw = (Tr < 0.0 && PB[128] > floor(PB[128])) ? -rt_pow(-Tr, PB[128]) : rt_pow(Tr, PB[128]);
  double dh = PB[130] - h;
  double xs = dh > PB[131] ? PB[131] : (dh >= PB[132] ? dh : PB[132]);
  double rho = pw / Tr * PB[129] * exp(xs * PB[133] * (1.0 / T));
  double rV2 = rho * (V * V);
  double qS = rV2 * PB[134] * m->S;
  double sa = sin(alpha), ca = cos(alpha);
  double mD = PB[126] * CXa * qS;
  double Lf = qS * CYa;
  double Fx = mD * ca + sa * Lf + m->P;
  double Fy = ca * Lf - mD * sa + 0.0;
  /* actuator: transport delay -> discrete filter (Ts 0.05) -> rate limiter -> saturation */
  double td = tdelay(m, t - PB[136], 0.0, PB[137]);
  m->td = td;
  if (major && m->tid2 == 0) m->df_y = m->df_x * PB[140] + PB[141] * td;
  double yv = m->df_y;
  if (m->rl_t != INFINITY) {
    double dT = t - m->rl_t, rate = yv - m->rl_prev;
    if (rate > dT * PB[142]) yv = dT * PB[142] + m->rl_prev;
    else if (dT * PB[143] > rate) yv = dT * PB[143] + m->rl_prev;
  }
  m->rl_out = yv;
  m->deltaz_RP = sat(yv, PB[145], PB[144]);
  /* СУ PID: altitude error -> pitch reference */
  double e_h = m->h_zh - h;
  double cs_d = (e_h * m->PID_CS[2] - X[10]) * m->PID_CS[3];
  double cs_pre = e_h * m->PID_CS[0] + X[9] + cs_d;
  m->vartheta_zh = sat(cs_pre, PB[4], PB[6]);
  double vref = m->use_PID_CS >= PB[146] ? m->vartheta_zh : m->vartheta;
  double dvt = vref - th;
  m->dvartheta = dvt;
  /* СС PID: pitch error -> elevator command */
  double ss_d = (dvt * m->PID_SS[2] - X[12]) * m->PID_SS[3];
  double ss_pre = dvt * m->PID_SS[0] + X[11] + ss_d;
  m->U_com_PID = sat(ss_pre, PB[5], PB[7]);
  if (m->use_RL >= PB[148]) m->U_com = PB[147] > fabs(0.0 - m->U_com_PID) ? 0.0 : m->U_com_PID;
  else m->U_com = m->use_PID_SS >= PB[9] ? m->U_com_PID : m->deltaz;
  if (major) { m->sumA[3] = m->aero_err[3] + PB[216]; m->sumA[4] = m->aero_err[4] + PB[216]; }
  m->dCm_ddeltaz = look2(h, Mach, PB + 201, PB + 206, PB + 151, MAXIDX + 4, 5) * m->sumA[3];
  m->K_alpha = look1(ad, PB + 225, PB + 218, 6) * m->sumA[4];
  if (major) m->sumA[2] = m->aero_err[2] + PB[216];
  m->mz = look2(Mach, ad, PB + 276, PB + 280, PB + 232, MAXIDX + 6, 4) * m->sumA[2];
  double ax = (Fx * cs - sn * Fy) / m->m0;
  double ay = (Fy * cs + Fx * sn) / m->m0 - m->g;
  double dze = m->use_RP >= PB[149] ? m->deltaz_RP : m->U_com;
  double Cm = PB[217] * m->dCm_ddeltaz * m->K_alpha * (dze * PB[150]) + m->mz;
  double wzd = Cm * (rV2 * PB[135] * m->S * m->c_) / m->Iz;
  double wz = X[8];
  double dq0 = (-wz) * q3 * 0.5, dq1 = q2 * wz * 0.5, dq2 = (-wz) * q1 * 0.5, dq3 = q0 * wz * 0.5;
  /* clamping anti-windup (СС) */
  double dz = ss_pre > PB[7] ? ss_pre - PB[7] : (ss_pre >= PB[5] ? 0.0 : ss_pre - PB[5]);
  double ss_i = m->PID_SS[1] * dvt;
  m->and_ss = (ss_pre * PB[291] != dz) && (sgn8(dz) == sgn8(ss_i));
  if (major) m->memout_ss = m->mem_ss;
  if (m->memout_ss) ss_i = PB[10];
  m->sim_time = t;
  m->dvartheta_dt = deriv_out(&m->d1, dvt, t);
  m->dvartheta_dt_dt = deriv_out(&m->d2, m->dvartheta_dt, t);
  m->SE = dvt * dvt; m->TSE = m->SE * t; m->AE = fabs(dvt); m->TAE = m->AE * t;
  dz = cs_pre > PB[6] ? cs_pre - PB[6] : (cs_pre >= PB[4] ? 0.0 : cs_pre - PB[4]);
  double cs_i = e_h * m->PID_CS[1];
  m->and_cs = (cs_pre * PB[292] != dz) && (sgn8(dz) == sgn8(cs_i));
  if (major) m->memout_cs = m->mem_cs;
  if (m->memout_cs) cs_i = PB[11];
  m->dvartheta_int = X[13]; m->ITAE = X[14]; m->IAE = X[15]; m->ISE = X[16]; m->ITSE = X[17];
  double *dX = m->dX;
  dX[0] = Vx; dX[1] = Vy; dX[2] = dq0; dX[3] = dq1; dX[4] = dq2; dX[5] = dq3;
  dX[6] = ax; dX[7] = ay; dX[8] = wzd; dX[9] = cs_i; dX[10] = cs_d; dX[11] = ss_i; dX[12] = ss_d;
  dX[13] = dvt; dX[14] = m->TAE; dX[15] = m->AE; dX[16] = m->SE; dX[17] = m->TSE;
}

/* model_simple_update part of the step, dll@0x271a-0x2908 (major only) */
static void update(b747o_model *m) {
  const double t = m->t;
  /* transport-delay ring push of (t, U_com) */
  m->head = (m->head < m->size - 1) ? m->head + 1 : 0;
  if (m->head == m->tail) m->tail = (m->tail < m->size - 1) ? m->tail + 1 : 0;
  m->ring_t[m->head] = t;
  m->ring_u[m->head] = m->U_com;
  if (m->tid2 == 0) m->df_x = PB[138] * m->df_x + PB[139] * m->td;
  m->rl_prev = m->rl_out; m->rl_t = t;
  m->mem_ss = m->and_ss;
  deriv_upd(&m->d1, m->dvartheta, t);
  deriv_upd(&m->d2, m->dvartheta_dt, t);
  m->mem_cs = m->and_cs;
}

/* model_simple_step, dll@0x16d0, with rt_ertODEUpdateContinuousStates (ode4), dll@0x2c60 */
void b747o_model_step(b747o_model *m) {
  double tnew = (double)(m->tick + 1) * H_STEP;
  outputs(m, 1);
  update(m);
  if (m->first) { m->first = 0; outputs(m, 0); }
  double y[18], f0[18], f1[18], f2[18];
  const double h = H_STEP, temp = 0.5 * h;
  double t0 = m->t;
  memcpy(y, m->X, sizeof y);
  memcpy(f0, m->dX, sizeof f0);
  for (int i = 0; i < 18; i++) m->X[i] = y[i] + temp * f0[i];
  m->t = t0 + temp;
  outputs(m, 0);
  memcpy(f1, m->dX, sizeof f1);
  for (int i = 0; i < 18; i++) m->X[i] = y[i] + temp * f1[i];
  outputs(m, 0);
  memcpy(f2, m->dX, sizeof f2);
  for (int i = 0; i < 18; i++) m->X[i] = y[i] + h * f2[i];
  m->t = tnew;
  outputs(m, 0);
  const double h6 = h / 6.0;
  for (int i = 0; i < 18; i++) m->X[i] = y[i] + h6 * (f0[i] + 2.0 * f1[i] + 2.0 * f2[i] + m->dX[i]);
  m->tick++;
  m->t = tnew;
  m->tid2 = (uint8_t)((m->tid2 + 1) % 5);
}

static void mdl_init_cb(void *c) { b747o_model_initialize((b747o_model *)c); }
static void mdl_step_cb(void *c) { b747o_model_step((b747o_model *)c); }

void b747o_iface_from_model(b747o_iface *f, b747o_model *m) {
  f->ctx = m; f->initialize = mdl_init_cb; f->step = mdl_step_cb;
  f->state0 = m->state0; f->h_zh = &m->h_zh; f->use_RP = &m->use_RP; f->use_PID_SS = &m->use_PID_SS;
  f->use_PID_CS = &m->use_PID_CS; f->PID_SS = m->PID_SS; f->PID_CS = m->PID_CS; f->deltaz = &m->deltaz;
  f->vartheta = &m->vartheta; f->P = &m->P; f->aero_err = m->aero_err;
  f->state = m->state; f->sim_time = &m->sim_time; f->vartheta_zh = &m->vartheta_zh; f->U_com_PID = &m->U_com_PID;
  f->CXa = &m->CXa; f->CYa = &m->CYa; f->mz = &m->mz; f->K_alpha = &m->K_alpha; f->dCm_ddeltaz = &m->dCm_ddeltaz;
  f->U_com = &m->U_com; f->deltaz_RP = &m->deltaz_RP; f->dvartheta = &m->dvartheta;
  f->dvartheta_int = &m->dvartheta_int; f->dvartheta_dt = &m->dvartheta_dt;
  f->dvartheta_dt_dt = &m->dvartheta_dt_dt; f->ITSE = &m->ITSE;
}
