/*
 * b747_batch_ref.c -- TEST INFRASTRUCTURE (oracle), not product code.
 * N independent (restatement model + env layer) pairs stepped in a loop: the CPU
 * checker the GPU parity tests compare against, and bench.py's "port" cpu_baseline.
 */
#include <stdlib.h>
#include <string.h>

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

void b747o_batch_step(b747o_batch *b, const double *actions, double *obs, double *rew, uint8_t *done,
                      double *terminal_obs, int auto_reset) {
  for (int64_t i = 0; i < b->n; i++) {
    b747o_env *e = &b->envs[i];
    int od = e->obs_dim;
    int d = b747o_env_step(e, actions[i], obs + i * od, rew + i);
    done[i] = (uint8_t)d;
    if (terminal_obs) memcpy(terminal_obs + i * od, obs + i * od, sizeof(double) * od);
    if (d && auto_reset) b747o_env_reset(e, obs + i * od);
  }
}

b747o_env *b747o_batch_env(b747o_batch *b, int64_t i) { return &b->envs[i]; }
b747o_model *b747o_batch_model(b747o_batch *b, int64_t i) { return &b->models[i]; }
