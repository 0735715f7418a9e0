/*
 * b747_oracle_api.c -- TEST INFRASTRUCTURE (oracle), not product code.
 * Small ctypes-friendly helpers around b747o_model (allocation + named field access).
 */
#include <stdlib.h>
#include <string.h>

#include "b747_oracle.h"

b747o_model *b747o_model_new(void) {
  b747o_model *m = malloc(sizeof *m);
  b747o_model_defaults(m);
  b747o_model_initialize(m);
  return m;
}

void b747o_model_free(b747o_model *m) { free(m); }

void b747o_model_step_n(b747o_model *m, long n) {
  for (long i = 0; i < n; i++) b747o_model_step(m);
}

#define F(name) if (!strcmp(n, #name)) return (double *)&m->name
double *b747o_model_ptr(b747o_model *m, const char *n) {
  F(state0); F(h_zh); F(use_RP); F(use_PID_SS); F(use_PID_CS); F(PID_SS); F(PID_CS); F(deltaz); F(vartheta);
  F(P); F(aero_err); F(Iz); F(S); F(c_); F(g); F(m0); F(use_RL);
  F(state); F(sim_time); F(vartheta_zh); F(U_com_PID); F(CXa); F(CYa); F(mz); F(K_alpha); F(dCm_ddeltaz);
  F(U_com); F(deltaz_RP); F(dvartheta); F(dvartheta_int); F(dvartheta_dt); F(dvartheta_dt_dt);
  F(TAE); F(ITAE); F(TSE); F(ITSE); F(AE); F(IAE); F(SE); F(ISE); F(alpha); F(V); F(Mach);
  F(X); F(dX); F(t); F(df_x); F(df_y); F(rl_prev); F(rl_t); F(td);
  return NULL;
}
#undef F

uint32_t b747o_model_tick(const b747o_model *m) { return m->tick; }
