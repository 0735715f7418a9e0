// b747_scalar.cu -- the reference's own plugin boundary, re-provided over the CUDA path.
//
// core/model.py:104-164 loads `core/model_simple.so` (Linux) / `model_simple_win64.dll` (Windows) with
// ctypes, calls three `void f(void)` entry points and reads/writes ~45 `double` globals with `in_dll`.
// This translation unit exports exactly those symbols, so an unchanged Model-style wrapper binds to
// model_simple.so built from this repo.  Behind them sits a private one-environment float64 handle of
// the batched engine: `model_simple_step()` uploads the tunable globals, launches the same sm_100a
// kernel the batched path uses (n_envs = 1) and downloads the stage-4 signals into the exported
// globals.  There is no CPU implementation: without a CUDA device the entry points abort loudly.
//
// Like the DLL, the library's state is process-global and non-reentrant; the reference gets private
// instances by copying the library file per Model (core/model.py:99-110), which works here too
// because the library is self-contained.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b747.h"
#include "../../include/b747_params.h"
#include "../../include/b747_scalar.h"

extern "C" {
// ---- signals (core/model.py:129-151; alpha, V, Mach are exported but unbound by Python) ----
double state[6], sim_time, vartheta_zh, U_com_PID, CXa, CYa, mz, K_alpha, dCm_ddeltaz, U_com, deltaz_RP, dvartheta,
    dvartheta_int, dvartheta_dt, dvartheta_dt_dt, TAE, ITAE, TSE, ITSE, AE, IAE, SE, ISE, alpha, V, Mach;
// ---- parameters (core/model.py:154-164) with the DLL's .data defaults ----
double state0[6] = B747_DEF_STATE0;
double h_zh = B747_DEF_H_ZH, use_RP = B747_DEF_USE_RP, use_PID_SS = B747_DEF_USE_PID_SS, use_PID_CS = B747_DEF_USE_PID_CS;
double PID_SS[4] = B747_DEF_PID_SS, PID_CS[4] = B747_DEF_PID_CS;
double deltaz = B747_DEF_DELTAZ, vartheta = B747_DEF_VARTHETA, P = B747_DEF_P;
double aero_err[5] = B747_DEF_AERO_ERR;
double Iz = B747_DEF_IZ, S = B747_DEF_S, c_ = B747_DEF_C, g = B747_DEF_G, m0 = B747_DEF_M0, use_RL = B747_DEF_USE_RL;
}

namespace {
b747_handle* g_h = nullptr;

[[noreturn]] void die(const char* what, int rc) {
  fprintf(stderr, "model_simple (b747 CUDA): %s failed (%d): %s\n", what, rc, b747_last_error());
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
  CK(b747_set_param(g_h, "PID_SS", PID_SS, 4)); CK(b747_set_param(g_h, "PID_CS", PID_CS, 4));
  CK(b747_set_param(g_h, "P", &P, 1)); CK(b747_set_param(g_h, "Iz", &Iz, 1)); CK(b747_set_param(g_h, "S", &S, 1));
  CK(b747_set_param(g_h, "c_", &c_, 1)); CK(b747_set_param(g_h, "g", &g, 1)); CK(b747_set_param(g_h, "m0", &m0, 1));
  CK(b747_set_param(g_h, "use_RP", &use_RP, 1)); CK(b747_set_param(g_h, "use_RL", &use_RL, 1));
  CK(b747_set_param(g_h, "use_PID_SS", &use_PID_SS, 1));
  set1("deltaz", deltaz); set1("vartheta", vartheta); set1("h_zh", h_zh);
  static const char* an[5] = {"aerr0", "aerr1", "aerr2", "aerr3", "aerr4"};
  for (int k = 0; k < 5; k++) set1(an[k], aero_err[k]);
  // the kernel reads `use_PID_CS >= 1` from the per-env flag word (bit 2), keeping the Memory bits
  double fl = get1("flags");
  int f = ((int)fl & ~4) | (use_PID_CS >= 1.0 ? 4 : 0);
  set1("flags", (double)f);
}

void pull_signals() {
  static const char* sn[6] = {"sig_state_x", "sig_state_y", "sig_state_Vx", "sig_state_Vy", "sig_state_vartheta", "sig_state_wz"};
  for (int k = 0; k < 6; k++) state[k] = get1(sn[k]);
#define G(v) v = get1("sig_" #v)
  G(sim_time); G(vartheta_zh); G(U_com_PID); G(CXa); G(CYa); G(mz); G(K_alpha); G(dCm_ddeltaz); G(U_com); G(deltaz_RP);
  G(dvartheta); G(dvartheta_int); G(dvartheta_dt); G(dvartheta_dt_dt); G(TAE); G(ITAE); G(TSE); G(ITSE); G(AE); G(IAE);
  G(SE); G(ISE); G(alpha); G(V); G(Mach);
#undef G
}
}  // namespace

extern "C" void model_simple_initialize(void) {
  ensure_handle();
  static const char* s0n[6] = {"state0_x", "state0_y", "state0_Vx", "state0_Vy", "state0_vartheta", "state0_wz"};
  for (int k = 0; k < 6; k++) set1(s0n[k], state0[k]);
  CK(b747_model_initialize(g_h));
  CK(b747_synchronize(g_h));
  pull_signals();  // all zero, like the DLL after initialize (dll@0x13e6-0x14d3)
}

extern "C" void model_simple_step(void) {
  ensure_handle();
  push_params();
  CK(b747_model_step(g_h, 1));
  CK(b747_synchronize(g_h));
  pull_signals();
}

// The DLL's terminate is a bare `ret` (dll@0x29d0): state survives it.  Only drain the stream here.
extern "C" void model_simple_terminate(void) {
  if (g_h) b747_synchronize(g_h);
}
