// b747_scalar_legacy.cu -- the boundary of the reference's legacy dynamics library (core/model_win64.dll) over the CUDA
// path: lib/model.so.  Same dynamics as model_simple (include/b747_scalar_legacy.h says how that was established), so
// the same one-environment float64 handle steps it; this unit only maps the legacy symbol surface onto it:
// 3-D state layout, deltaz_com / deltaz_real / deltaz_ref names, aerodynamic coefficients tapped before the aero-error
// gains, aero_err[4] not connected, no use_RP switch, I[3] instead of Iz.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b747.h"
#include "../../include/b747_params.h"
#include "../../include/b747_scalar_legacy.h"

extern "C" {
double state[6], sim_time, vartheta_zh, deltaz_ref, deltaz_com, deltaz_real, CXa, CYa, mz, K_alpha, dCm_ddeltaz, dvartheta,
    dvartheta_int, dvartheta_dt, dvartheta_dt_dt, TAE, ITAE, TSE, ITSE, AE, IAE, SE, ISE;
// .data defaults of model_win64.dll
double state0[6] = {0.0, 11000.0, 0.0, 259.1667, 0.0, 0.0};
double h_zh = 5000.0, use_PID_SS = 0.0, use_PID_CS = 0.0, use_RL = 0.0;
double PID_SS[4] = B747_DEF_PID_SS, PID_CS[4] = B747_DEF_PID_CS;
double deltaz = 0.0, vartheta = 0.0, P = B747_DEF_P;
double aero_err[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
double I[3] = {24700000.0, 44900000.0, 67300000.0};
double S = B747_DEF_S, c_ = B747_DEF_C, g = B747_DEF_G, m0 = B747_DEF_M0;
}

namespace {
b747_handle* g_h = nullptr;
const double kP[B747_NP] = B747_P_INIT;  // model_simple_P == the legacy model_P values for the shared blocks

[[noreturn]] void die(const char* what, int rc) {
  fprintf(stderr, "model (b747 CUDA, legacy boundary): %s failed (%d): %s\n", what, rc, b747_last_error());
  abort();
}
#define CK(call)                 \
  do {                           \
    int rc_ = (call);            \
    if (rc_) die(#call, rc_);    \
  } while (0)

void ensure_handle() {
  if (g_h) return;
  b747_cfg c;
  memset(&c, 0, sizeof c);
  c.abi_version = B747_ABI_VERSION;
  const char* dev = getenv("B747_DEVICE");
  c.device = dev ? atoi(dev) : 0;
  c.dtype = B747_F64; c.n_envs = 1;
  c.obs_type = B747_OBS_PID_LIKE; c.rew_type = B747_REW_CLASSIC; c.ctrl_type = B747_CTRL_MANUAL;
  c.ctrl_mode = B747_MODE_DIRECT; c.reset_ref_mode = B747_RESET_NONE; c.disturbance_mode = B747_DIST_NONE;
  c.substeps = 1; c.env_layer = 0; c.export_signals = 1; c.done_tick = INT64_MAX;
  c.tk = 1e300; c.action_max = 1; c.vartheta_max = 1; c.sample_time = 0.01;
  CK(b747_create(&c, &g_h));
}

void set1(const char* name, double v) { CK(b747_set_field(g_h, b747_field_index(name), &v)); }
double get1(const char* name) {
  double v;
  CK(b747_get_field(g_h, b747_field_index(name), &v));
  return v;
}

void push_params() {
  const double one = 1.0;
  CK(b747_set_param(g_h, "PID_SS", PID_SS, 4)); CK(b747_set_param(g_h, "PID_CS", PID_CS, 4));
  CK(b747_set_param(g_h, "P", &P, 1)); CK(b747_set_param(g_h, "Iz", &I[2], 1)); CK(b747_set_param(g_h, "S", &S, 1));
  CK(b747_set_param(g_h, "c_", &c_, 1)); CK(b747_set_param(g_h, "g", &g, 1)); CK(b747_set_param(g_h, "m0", &m0, 1));
  CK(b747_set_param(g_h, "use_RP", &one, 1));  // the legacy diagram has no use_RP switch: actuator always in the loop
  CK(b747_set_param(g_h, "use_RL", &use_RL, 1)); CK(b747_set_param(g_h, "use_PID_SS", &use_PID_SS, 1));
  set1("deltaz", deltaz); set1("vartheta", vartheta); set1("h_zh", h_zh);
  static const char* an[4] = {"aerr0", "aerr1", "aerr2", "aerr3"};
  for (int k = 0; k < 4; k++) set1(an[k], aero_err[k]);
  set1("aerr4", 0.0);  // the K_alpha error input is not connected in the legacy diagram
  double fl = get1("flags");
  int f = ((int)fl & ~4) | (use_PID_CS >= 1.0 ? 4 : 0);
  set1("flags", (double)f);
}

void pull_signals() {
  state[0] = get1("sig_state_x"); state[1] = get1("sig_state_y"); state[2] = 0.0;
  state[3] = get1("sig_state_Vx"); state[4] = get1("sig_state_Vy"); state[5] = 0.0;
  sim_time = get1("sig_sim_time"); vartheta_zh = get1("sig_vartheta_zh");
  deltaz_ref = get1("sig_U_com_PID"); deltaz_com = get1("sig_U_com"); deltaz_real = get1("sig_deltaz_RP");
  CXa = get1("sig_CXa_tab"); CYa = get1("sig_CYa_tab"); mz = get1("sig_mz_tab");
  dCm_ddeltaz = get1("sig_dCm_tab") * kP[217];  // tapped after the per-degree -> per-radian gain (Gain2)
  K_alpha = get1("sig_K_alpha");
#define G(v) v = get1("sig_" #v)
  G(dvartheta); G(dvartheta_int); G(dvartheta_dt); G(dvartheta_dt_dt); G(TAE); G(ITAE); G(TSE); G(ITSE); G(AE); G(IAE);
  G(SE); G(ISE);
#undef G
}
}  // namespace

extern "C" void model_initialize(void) {
  ensure_handle();
  set1("state0_x", state0[0]); set1("state0_y", state0[1]); set1("state0_Vx", state0[3]); set1("state0_Vy", state0[4]);
  set1("state0_vartheta", 0.0); set1("state0_wz", 0.0);
  CK(b747_model_initialize(g_h));
  CK(b747_synchronize(g_h));
  pull_signals();
  // model.py's initialize() zeroes deltaz / vartheta on the Python side for model_simple; the DLL itself leaves the
  // tunables alone, and so does this boundary
}

extern "C" void model_step(void) {
  ensure_handle();
  push_params();
  CK(b747_model_step(g_h, 1));
  CK(b747_synchronize(g_h));
  pull_signals();
}

extern "C" void model_terminate(void) {
  if (g_h) b747_synchronize(g_h);
}
