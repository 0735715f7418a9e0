/*
 * b747_env_ref.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * float64 restatement of the reference's Python layers above the model library,
 * written against the backend-neutral b747o_iface so the SAME code is checked on
 * the reference DLL and then reused over the C restatement for large batches:
 *   Model.initialize/step        core/model.py:238-250
 *   Controller.reset             core/controller.py:134-201
 *   Controller.step              core/controller.py:231-264
 *   Controller properties        core/controller.py:267-344
 *   ControllerEnv obs/reward/... env/ctrl_env.py:109-278
 *   calc_exp_k                   tools/general.py:32-33 (constants passed in cfg.rew)
 * Random draws use a counter-based Philox4x32-10 stream keyed by (seed, env, episode)
 * instead of CPython's Mersenne Twister (stream parity with MT is a stated non-goal,
 * SURVEY.md 8d); the distributions and draw order are the reference's.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "b747_oracle.h"

#define DEG (M_PI / 180.0)

/* ------------------------------ Philox ------------------------------ */
void b747o_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double b747o_uniform53(uint64_t seed, uint64_t env_id, uint64_t episode_idx, uint32_t draw) {
  uint32_t ctr[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)episode_idx, draw >> 1};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, w[4];
  b747o_philox4x32(ctr, key, w);
  uint32_t a = w[2 * (draw & 1)] >> 5, b = w[2 * (draw & 1) + 1] >> 6;
  return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}

static double uni(const b747o_env_cfg *c, uint64_t env, uint64_t ep, uint32_t draw, double a, double b) {
  return a + (b - a) * b747o_uniform53(c->seed, env, ep, draw); /* random.uniform */
}

/* ------------------------------ helpers ------------------------------ */
int b747o_obs_dim(int obs_type) {
  switch (obs_type) {
    case B747_OBS_PID_LIKE: return 3;
    case B747_OBS_SPEED_MODE: return 5;
    case B747_OBS_PID_AERO: return 8;
    case B747_OBS_PID_SPEED_AERO: return 10;
    case B747_OBS_MODEL_STATE: return 7;
  }
  return -1;
}

/* Controller.is_done: model.time >= tk with time == fl(tick*0.01) (dll@0x172f-0x1747) */
int64_t b747o_done_tick(double tk) {
  if (!(tk == tk)) return INT64_MAX;
  if (tk <= 0) return 0;
  if (isinf(tk)) return INT64_MAX;
  int64_t n = (int64_t)floor(tk / 0.01) - 2;
  if (n < 0) n = 0;
  while (!((double)n * 0.01 >= tk)) n++;
  return n;
}

static double nan_to_num(double x) { /* np.nan_to_num, core/model.py:167-168 */
  if (x != x) return 0.0;
  if (isinf(x)) return x > 0 ? DBL_MAX : -DBL_MAX;
  return x;
}

void b747o_env_init(b747o_env *e, const b747o_env_cfg *cfg, const b747o_iface *mdl, uint64_t env_id) {
  memset(e, 0, sizeof *e);
  e->cfg = *cfg; e->mdl = *mdl; e->env_id = env_id; e->obs_dim = b747o_obs_dim(cfg->obs_type);
  /* Controller._init_model, core/controller.py:128-131 */
  int use_ctrl = cfg->ctrl_type == B747_CTRL_SEMI_MANUAL || cfg->ctrl_type == B747_CTRL_FULL_AUTO;
  int manual = cfg->ctrl_type == B747_CTRL_MANUAL || cfg->ctrl_type == B747_CTRL_SEMI_MANUAL;
  *e->mdl.use_RP = 1.0; *e->mdl.use_PID_CS = use_ctrl; *e->mdl.use_PID_SS = !manual;
  e->ep.use_ctrl = use_ctrl;
}

/* Controller.reset random part, core/controller.py:144-193 */
void b747o_env_draw_episode(const b747o_env_cfg *c, uint64_t env, uint64_t epi, b747o_episode *ep) {
  memset(ep, 0, sizeof *ep);
  ep->use_ctrl = c->ctrl_type == B747_CTRL_SEMI_MANUAL || c->ctrl_type == B747_CTRL_FULL_AUTO;
  ep->h_ref = 11000.0;
  double h0 = uni(c, env, epi, 0, 1000, 11000);
  double Vx = uni(c, env, epi, 1, 100, 265);
  double Vy = uni(c, env, epi, 2, -20, 20);
  double wz0 = uni(c, env, epi, 3, -0.001, 0.001);
  if (c->reset_ref_mode == B747_RESET_CONST) {
    double v = uni(c, env, epi, 4, -c->vartheta_max, -1 * M_PI / 180);
    v *= (b747o_uniform53(c->seed, env, epi, 5) < 0.5) ? 1.0 : -1.0;
    ep->vref_const = v;
  } else if (c->reset_ref_mode == B747_RESET_OSCILLATING) {
    double A1 = uni(c, env, epi, 4, 0, c->vartheta_max);
    double A2 = uni(c, env, epi, 5, 0, c->vartheta_max - A1);
    double A3 = uni(c, env, epi, 6, 0, c->vartheta_max - A1 - A2);
    ep->osc_A[0] = A1; ep->osc_A[1] = A2; ep->osc_A[2] = A3;
    for (int i = 0; i < 3; i++) ep->osc_f[i] = uni(c, env, epi, 7 + i, 0.01, 0.5);
    ep->oscillating = 1;
  } else if (c->reset_ref_mode == B747_RESET_HYBRID) {
    int use_ctrl = b747o_uniform53(c->seed, env, epi, 4) < 0.5;
    ep->use_ctrl = use_ctrl;
    if (use_ctrl) ep->h_ref = h0 + uni(c, env, epi, 5, -1000, 1000);
    else ep->vref_const = uni(c, env, epi, 5, -c->vartheta_max, c->vartheta_max);
  }
  ep->state0[0] = 0; ep->state0[1] = h0; ep->state0[2] = Vx; ep->state0[3] = Vy; ep->state0[4] = 0; ep->state0[5] = wz0;
  if (c->disturbance_mode == B747_DIST_AERO) {
    static const double mean[5] = {-0.1, 0.1, -0.1, -0.1, 0.1};
    for (int i = 0; i < 5; i++) {
      if (c->has_fixed_aero_err) { ep->aero_err[i] = c->fixed_aero_err[i]; continue; }
      double u1 = 1.0 - b747o_uniform53(c->seed, env, epi, 10 + 2 * i);
      double u2 = b747o_uniform53(c->seed, env, epi, 11 + 2 * i);
      double z = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
      ep->aero_err[i] = mean[i] + 0.5 * z; /* np.random.normal(mean, 0.5) */
    }
  }
}

static double vartheta_func(const b747o_episode *ep, double t) {
  if (!ep->oscillating) return ep->vref_const;
  /* A1*sin(2*pi*f1*t)+A2*sin(2*pi*f2*t)+A3*sin(2*pi*f3*t), core/controller.py:165 */
  return ep->osc_A[0] * sin(2 * M_PI * ep->osc_f[0] * t) + ep->osc_A[1] * sin(2 * M_PI * ep->osc_f[1] * t) +
         ep->osc_A[2] * sin(2 * M_PI * ep->osc_f[2] * t);
}

/* Controller.vartheta_ref, core/controller.py:267-270 */
static double vartheta_ref(const b747o_env *e) {
  return *e->mdl.use_PID_CS ? *e->mdl.vartheta_zh : *e->mdl.vartheta;
}

/* ControllerEnv._get_obs, env/ctrl_env.py:200-247 */
static void get_obs(const b747o_env *e, double *obs) {
  static const double mx_pid[3] = {60 * M_PI, M_PI, M_PI};
  const b747o_iface *m = &e->mdl;
  int n = 0;
  double mx[10];
  if (e->cfg.obs_type == B747_OBS_MODEL_STATE) {
    obs[0] = vartheta_ref(e);
    for (int i = 0; i < 6; i++) obs[1 + i] = nan_to_num(m->state[i]);
    const double mxs[7] = {10 * M_PI / 180, 12000, 15000, 500, 100, M_PI, M_PI};
    memcpy(mx, mxs, sizeof mxs);
    n = 7;
  } else {
    obs[0] = *m->dvartheta_int; obs[1] = *m->dvartheta; obs[2] = *m->dvartheta_dt;
    memcpy(mx, mx_pid, sizeof mx_pid);
    n = 3;
    if (e->cfg.obs_type == B747_OBS_SPEED_MODE || e->cfg.obs_type == B747_OBS_PID_SPEED_AERO) {
      obs[n] = nan_to_num(m->state[2]); mx[n++] = 500;
      obs[n] = nan_to_num(m->state[3]); mx[n++] = 100;
    }
    if (e->cfg.obs_type == B747_OBS_PID_AERO || e->cfg.obs_type == B747_OBS_PID_SPEED_AERO) {
      obs[n] = *m->CXa; mx[n++] = 0.5;
      obs[n] = *m->CYa; mx[n++] = 2;
      obs[n] = *m->mz; mx[n++] = 0.6;
      obs[n] = *m->dCm_ddeltaz; mx[n++] = 0.05;
      obs[n] = *m->K_alpha; mx[n++] = 1.;
    }
  }
  if (e->cfg.norm_obs)
    for (int i = 0; i < n; i++) obs[i] /= mx[i];
}

/* Controller.quality, core/controller.py:334-336 */
static double quality(const b747o_env *e) {
  double vr = vartheta_ref(e);
  return exp(-60 * 0.1 * *e->mdl.ITSE / (e->cfg.tk * (vr * vr)));
}

/* ControllerEnv.get_reward, env/ctrl_env.py:109-192 */
static double get_reward(b747o_env *e) {
  const b747o_iface *m = &e->mdl;
  const double *k = e->cfg.rew;
  double vr = vartheta_ref(e);
  double vf = vr != 0.0 ? vr : e->cfg.vartheta_max;
  double dv = *m->dvartheta, time = *m->sim_time;
  switch (e->cfg.rew_type) {
    case B747_REW_CLASSIC: {
      double k1 = k[0], k2 = k[1], k3 = k[2], k0 = k[3], kITSE = k[4], kf = k[5], kt = k[6], ko = k[7];
      double r1 = 0.50 * exp(-k0 * (k1 * fabs(dv) + k2 * 1 * fabs(*m->dvartheta_dt) + k3 * fabs(*m->dvartheta_dt_dt)) / fabs(vf));
      double r2 = (vr * dv < 0) ? 0.20 * exp(-ko * fabs(dv / vf)) : 0.20;
      double r3 = (fabs(dv / vf) > 0.05) ? 0.20 * exp(-kt * time) : 0.20;
      double r4 = 0.1 * exp(-kITSE * *m->ITSE / (vf * vf));
      double rf = 0;
      if (e->cfg.ctrl_mode == B747_MODE_DIRECT)
        rf = -kf * fabs(dv / (2 * vf)) * (fabs(*m->deltaz - *m->U_com_PID)) / (34 * M_PI / 180);
      return r1 + r2 + r3 + r4 + rf;
    }
    case B747_REW_PID_LIKE:
      return exp(-k[0] * fabs(*m->U_com - *m->U_com_PID) / (34 * M_PI / 180));
    case B747_REW_QUALITY:
    case B747_REW_MINIMAL:
      return quality(e);
    case B747_REW_TF_REFERENCE: {
      double overshoot = fabs(dv / vf) * 100;
      if (overshoot > 5) e->tf_tp = time;
      return exp(-k[2] * fabs(overshoot - k[0]) * fabs(k[1] - e->tf_tp));
    }
  }
  return 0.0;
}

/* Controller._post_step with use_storage, core/controller.py:209-228: names, order and unit conversions as there
 * (`deltaz_real*180/pi`, `vartheta_ref*180/pi` multiply first; state_dict['vartheta'] is `v *= 180/pi`). */
static void post_step(b747o_env *e, double action) {
  if (!e->rec || e->rec_n >= e->rec_cap) return;
  const b747o_iface *m = &e->mdl;
  double *r = e->rec + (size_t)e->rec_n * B747O_NREC;
  r[B747O_REC_t] = *m->sim_time;
  r[B747O_REC_U_com] = *m->U_com;
  r[B747O_REC_U_PID] = *m->U_com_PID;
  r[B747O_REC_deltaz] = *m->deltaz_RP * 180 / M_PI;
  r[B747O_REC_hzh] = *m->h_zh;
  r[B747O_REC_vartheta_ref] = vartheta_ref(e) * 180 / M_PI;
  r[B747O_REC_U_RL] = action;
  r[B747O_REC_x] = nan_to_num(m->state[0]); r[B747O_REC_y] = nan_to_num(m->state[1]);
  r[B747O_REC_Vx] = nan_to_num(m->state[2]); r[B747O_REC_Vy] = nan_to_num(m->state[3]);
  r[B747O_REC_vartheta] = nan_to_num(m->state[4]) * (180 / M_PI);
  r[B747O_REC_wz] = nan_to_num(m->state[5]);
  e->rec_n++;
}
void b747o_env_set_recorder(b747o_env *e, double *buf, int32_t capacity_steps) {
  e->rec = buf; e->rec_cap = capacity_steps; e->rec_n = 0;
}
int32_t b747o_env_recorded(const b747o_env *e) { return e->rec_n; }

void b747o_env_reset_to(b747o_env *e, const b747o_episode *ep, double *obs) {
  b747o_iface *m = &e->mdl;
  e->rec_n = 0;
  if (ep != &e->ep) e->ep = *ep;
  if (e->cfg.reset_ref_mode != B747_RESET_HYBRID)
    e->ep.use_ctrl = e->cfg.ctrl_type == B747_CTRL_SEMI_MANUAL || e->cfg.ctrl_type == B747_CTRL_FULL_AUTO;
  if (e->cfg.reset_ref_mode == B747_RESET_HYBRID) {
    /* Controller._init_model re-creates the Model: every tunable back to its default */
    static const double pid_cs[4] = {0.0069214, 0.00057832, 0.0083279, 1.8385};
    static const double pid_ss[4] = {-5.9151, -1.2404, -6.6927, 58.0826};
    memcpy(m->PID_CS, pid_cs, sizeof pid_cs); memcpy(m->PID_SS, pid_ss, sizeof pid_ss);
    *m->h_zh = 11000.0; *m->P = 275000.0; memset(m->aero_err, 0, 5 * sizeof(double));
    *m->use_RP = 1.0; *m->use_PID_CS = ep->use_ctrl;
    *m->use_PID_SS = !(e->cfg.ctrl_type == B747_CTRL_MANUAL || e->cfg.ctrl_type == B747_CTRL_SEMI_MANUAL);
  }
  memcpy(m->state0, e->ep.state0, 6 * sizeof(double)); /* Model.set_initial */
  if (e->cfg.disturbance_mode == B747_DIST_AERO) memcpy(m->aero_err, e->ep.aero_err, 5 * sizeof(double));
  /* Model.initialize, core/model.py:238-244 */
  m->initialize(m->ctx);
  *m->deltaz = 0; *m->vartheta = 0;
  e->step_count = 0; e->ep_return = 0;
  get_obs(e, obs);
}

void b747o_env_reset(b747o_env *e, double *obs) {
  b747o_episode ep;
  if (e->cfg.reset_ref_mode == B747_RESET_NONE) { /* no random reset: same state0 / reference again */
    b747o_env_reset_to(e, &e->ep, obs);
    return;
  }
  b747o_env_draw_episode(&e->cfg, e->env_id, e->episode_idx, &ep);
  e->episode_idx++;
  b747o_env_reset_to(e, &ep, obs);
}

int b747o_env_step(b747o_env *e, double action, double *obs, double *reward) {
  b747o_iface *m = &e->mdl;
  const b747o_env_cfg *c = &e->cfg;
  if (c->norm_act) action *= c->action_max; /* env/ctrl_env.py:262-264 */
  /* Controller.step, core/controller.py:233-251 */
  if (!e->ep.use_ctrl) *m->vartheta = vartheta_func(&e->ep, *m->sim_time);
  else *m->h_zh = e->ep.h_ref;
  if (!*m->use_PID_SS) {
    const double lim = 17 * M_PI / 180;
    double dz;
    switch (c->ctrl_mode) {
      case B747_MODE_ADD_PROC: dz = (1 + action) * *m->U_com_PID; dz = dz < -lim ? -lim : (dz > lim ? lim : dz); break;
      case B747_MODE_ADD_DIRECT: dz = action + *m->U_com_PID; dz = dz < -lim ? -lim : (dz > lim ? lim : dz); break;
      case B747_MODE_ANG_VEL: dz = *m->deltaz + action * c->sample_time; dz = dz < -lim ? -lim : (dz > lim ? lim : dz); break;
      default: dz = action; break;
    }
    *m->deltaz = dz;
  }
  for (int k = 0; k < c->substeps; k++) { /* K-loop, core/controller.py:255-264 */
    m->step(m->ctx);
    post_step(e, action);
  }
  get_obs(e, obs);
  double r = get_reward(e);
  *reward = r;
  e->ep_return += r;
  e->step_count++;
  /* ControllerEnv.is_done, env/ctrl_env.py:255-257 (is_nan_err is dead: nan_to_num) */
  int done = *m->sim_time >= c->tk; /* Controller.is_done, core/controller.py:316-319 */
  if (c->use_limiter && (fabs(nan_to_num(m->state[4])) > 5 * M_PI / 180 + c->vartheta_max || *m->deltaz > c->action_max))
    done = 1;
  return done;
}
