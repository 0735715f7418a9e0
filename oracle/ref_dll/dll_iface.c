/*
 * dll_iface.c -- TEST INFRASTRUCTURE (oracle), not product code.
 * Presents one mapped instance of the reference DLL (pe_host.c) through the
 * backend-neutral b747o_iface, and provides the CPU-baseline rollout drivers that
 * run the oracle's env layer (b747_env_ref.c) over the DLL's own machine code.
 */
#include <stdlib.h>
#include <string.h>

#include "../b747_oracle.h"

typedef struct b747ref_inst b747ref_inst;
b747ref_inst *b747ref_open(void);
void b747ref_close(b747ref_inst *in);
void *b747ref_sym(b747ref_inst *in, const char *name);
void b747ref_call(void *fn);

typedef struct dll_ctx { b747ref_inst *inst; void *f_init, *f_step; } dll_ctx;

static void dll_init_cb(void *c) { b747ref_call(((dll_ctx *)c)->f_init); }
static void dll_step_cb(void *c) { b747ref_call(((dll_ctx *)c)->f_step); }

#define SYM(field, name) f->field = (double *)b747ref_sym(in, name)
int b747ref_iface(b747ref_inst *in, b747o_iface *f) {
  dll_ctx *c = calloc(1, sizeof *c);
  c->inst = in;
  c->f_init = b747ref_sym(in, "model_simple_initialize");
  c->f_step = b747ref_sym(in, "model_simple_step");
  if (!c->f_init || !c->f_step) { free(c); return -1; }
  f->ctx = c; f->initialize = dll_init_cb; f->step = dll_step_cb;
  SYM(state0, "state0"); SYM(h_zh, "h_zh"); SYM(use_RP, "use_RP"); SYM(use_PID_SS, "use_PID_SS");
  SYM(use_PID_CS, "use_PID_CS"); SYM(PID_SS, "PID_SS"); SYM(PID_CS, "PID_CS"); SYM(deltaz, "deltaz");
  SYM(vartheta, "vartheta"); SYM(P, "P"); SYM(aero_err, "aero_err");
  SYM(state, "state"); SYM(sim_time, "sim_time"); SYM(vartheta_zh, "vartheta_zh"); SYM(U_com_PID, "U_com_PID");
  SYM(CXa, "CXa"); SYM(CYa, "CYa"); SYM(mz, "mz"); SYM(K_alpha, "K_alpha"); SYM(dCm_ddeltaz, "dCm_ddeltaz");
  SYM(U_com, "U_com"); SYM(deltaz_RP, "deltaz_RP"); SYM(dvartheta, "dvartheta"); SYM(dvartheta_int, "dvartheta_int");
  SYM(dvartheta_dt, "dvartheta_dt"); SYM(dvartheta_dt_dt, "dvartheta_dt_dt"); SYM(ITSE, "ITSE");
  return 0;
}
#undef SYM

/* One env (ControllerEnv equivalent) over a private DLL instance. */
typedef struct b747ref_env { b747ref_inst *inst; b747o_env env; } b747ref_env;

b747ref_env *b747ref_env_create(const b747o_env_cfg *cfg, uint64_t env_id) {
  b747ref_env *r = calloc(1, sizeof *r);
  r->inst = b747ref_open();
  if (!r->inst) { free(r); return NULL; }
  b747o_iface f;
  memset(&f, 0, sizeof f);
  if (b747ref_iface(r->inst, &f)) { b747ref_close(r->inst); free(r); return NULL; }
  b747o_env_init(&r->env, cfg, &f, env_id);
  return r;
}

void b747ref_env_destroy(b747ref_env *r) {
  if (!r) return;
  free(r->env.mdl.ctx);
  b747ref_close(r->inst);
  free(r);
}

b747o_env *b747ref_env_get(b747ref_env *r) { return &r->env; }

/* Roll `n_steps` env steps with auto-reset; actions[n_steps]; outputs optional (may be NULL).
 * Returns the number of env steps executed.  This is the loop bench.py times as the
 * reference CPU arm (DLL machine code + C env layer, one instance per worker). */
long b747ref_env_rollout(b747ref_env *r, long n_steps, const double *actions, double *obs_out, double *rew_out,
                         uint8_t *done_out, int auto_reset) {
  b747o_env *e = &r->env;
  double obs[16], rew;
  int od = e->obs_dim;
  for (long k = 0; k < n_steps; k++) {
    int d = b747o_env_step(e, actions[k], obs, &rew);
    if (obs_out) memcpy(obs_out + k * od, obs, sizeof(double) * od);
    if (rew_out) rew_out[k] = rew;
    if (done_out) done_out[k] = (uint8_t)d;
    if (d && auto_reset) b747o_env_reset(e, obs);
  }
  return n_steps;
}
