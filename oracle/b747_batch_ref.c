/*
 * b747_batch_ref.c -- TEST INFRASTRUCTURE (oracle), not product code.
 * N independent (restatement model + env layer) pairs stepped in a loop: the CPU
 * checker the GPU parity tests compare against, and bench.py's "port" cpu_baseline.
 * Environments are independent, so large batches are stepped by a few pthreads (the 4096-env x 1000-step
 * parity case of BASELINE configs[1] then takes seconds on the host cores).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "b747_oracle.h"

struct b747o_batch {
  int64_t n;
  b747o_model *models;
  b747o_env *envs;
};

b747o_batch *b747o_batch_create(const b747o_env_cfg *cfg, int64_t n, uint64_t env_id_offset) {
  b747o_batch *b = calloc(1, sizeof *b);
  b->n = n;
  b->models = malloc(sizeof(b747o_model) * (size_t)n);
  b->envs = malloc(sizeof(b747o_env) * (size_t)n);
  for (int64_t i = 0; i < n; i++) {
    b747o_iface f;
    b747o_model_defaults(&b->models[i]);
    b747o_model_initialize(&b->models[i]);
    b747o_iface_from_model(&f, &b->models[i]);
    b747o_env_init(&b->envs[i], cfg, &f, env_id_offset + (uint64_t)i);
  }
  return b;
}

void b747o_batch_destroy(b747o_batch *b) {
  if (!b) return;
  free(b->models); free(b->envs); free(b);
}

void b747o_batch_reset(b747o_batch *b, double *obs) {
  for (int64_t i = 0; i < b->n; i++) b747o_env_reset(&b->envs[i], obs + i * b->envs[i].obs_dim);
}

void b747o_batch_reset_to(b747o_batch *b, const b747o_episode *eps, double *obs) {
  for (int64_t i = 0; i < b->n; i++) b747o_env_reset_to(&b->envs[i], &eps[i], obs + i * b->envs[i].obs_dim);
}

typedef struct {
  b747o_batch *b;
  const double *actions;
  double *obs, *rew, *terminal_obs;
  uint8_t *done;
  int auto_reset;
  int64_t lo, hi;
} step_job;

static void *step_range(void *arg) {
  step_job *j = arg;
  for (int64_t i = j->lo; i < j->hi; i++) {
    b747o_env *e = &j->b->envs[i];
    int od = e->obs_dim;
    int d = b747o_env_step(e, j->actions[i], j->obs + i * od, j->rew + i);
    j->done[i] = (uint8_t)d;
    if (j->terminal_obs) memcpy(j->terminal_obs + i * od, j->obs + i * od, sizeof(double) * od);
    if (d && j->auto_reset) b747o_env_reset(e, j->obs + i * od);
  }
  return NULL;
}

void b747o_batch_step(b747o_batch *b, const double *actions, double *obs, double *rew, uint8_t *done,
                      double *terminal_obs, int auto_reset) {
  enum { MAXT = 32 };
  long nt = b->n >= 256 ? sysconf(_SC_NPROCESSORS_ONLN) : 1;
  if (nt > MAXT) nt = MAXT;
  if (nt < 1) nt = 1;
  step_job jobs[MAXT];
  pthread_t th[MAXT];
  int started[MAXT] = {0};
  for (long t = 0; t < nt; t++) {
    step_job j = {b, actions, obs, rew, terminal_obs, done, auto_reset, b->n * t / nt, b->n * (t + 1) / nt};
    jobs[t] = j;
    if (t > 0 && pthread_create(&th[t], NULL, step_range, &jobs[t]) == 0) started[t] = 1;
  }
  step_range(&jobs[0]);
  for (long t = 1; t < nt; t++) {
    if (started[t]) pthread_join(th[t], NULL);
    else step_range(&jobs[t]);  /* thread creation failed: run the slice here */
  }
}

b747o_env *b747o_batch_env(b747o_batch *b, int64_t i) { return &b->envs[i]; }
b747o_model *b747o_batch_model(b747o_batch *b, int64_t i) { return &b->models[i]; }

/* named field `name`[idx] of every model of the batch (per-step state / signal parity checks) */
double *b747o_model_ptr(b747o_model *m, const char *n);
int b747o_batch_gather(b747o_batch *b, const char *name, int idx, double *out) {
  const double *p0 = b747o_model_ptr(&b->models[0], name);
  if (!p0) return -1;
  const size_t off = (size_t)((const char *)(p0 + idx) - (const char *)&b->models[0]);
  for (int64_t i = 0; i < b->n; i++) out[i] = *(const double *)((const char *)&b->models[i] + off);
  return 0;
}
void b747o_batch_ticks(b747o_batch *b, int64_t *out) {
  for (int64_t i = 0; i < b->n; i++) out[i] = (int64_t)b->models[i].tick;
}
