/*
 * b747_oracle.h -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Plain-C float64 restatement of the reference's hot path:
 *   - the Simulink-Coder model inside core/model_simple_win64.dll
 *     (model_simple_initialize dll@0x12a0, model_simple_step dll@0x16d0,
 *      ode4 dll@0x2c60, look2_binlx dll@0x1000, rt_TDelayInterpolate dll@0x29e0;
 *      block-by-block description: SURVEY.md Appendix B), and
 *   - the Python layers above it: core/model.py:238-250 (Model.initialize/step),
 *     core/controller.py:134-264,267-344 (Controller.reset/step/properties),
 *     env/ctrl_env.py:109-278 (ControllerEnv obs / reward / done / step / reset).
 *
 * Parity is PINNED: tests/test_oracle_vs_dll.py runs this restatement against the
 * DLL's own machine code (oracle/_ref/libb747_ref.so) and against the committed
 * golden vectors in tests/golden/ that were generated from the DLL.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 */
#ifndef B747_ORACLE_H
#define B747_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B747O_RING 1024

typedef struct b747o_deriv { double tA, uA, tB, uB; } b747o_deriv;

/* One model instance == the DLL's process-global state (tunables + signals + rtM/B/DW/X). */
typedef struct b747o_model {
  /* ---- tunable globals (core/model.py:154-164 + unbound exports) ---- */
  double state0[6], h_zh, use_RP, use_PID_SS, use_PID_CS, PID_SS[4], PID_CS[4];
  double deltaz, vartheta, P, aero_err[5], Iz, S, c_, g, m0, use_RL;
  /* ---- exported signals (core/model.py:129-151 + alpha, V, Mach) ---- */
  double state[6], sim_time, vartheta_zh, U_com_PID, CXa, CYa, mz, K_alpha, dCm_ddeltaz;
  double U_com, deltaz_RP, dvartheta, dvartheta_int, dvartheta_dt, dvartheta_dt_dt;
  double TAE, ITAE, TSE, ITSE, AE, IAE, SE, ISE, alpha, V, Mach;
  /* ---- continuous states and their derivatives ---- */
  double X[18], dX[18];
  /* ---- solver info ---- */
  double t;
  uint32_t tick;
  uint8_t tid2, first;
  /* ---- DWork ---- */
  double ic_t;               /* IC block FirstOutputTime */
  double df_x, df_y;         /* discrete state-space state and held output */
  double rl_prev, rl_t, rl_out;
  b747o_deriv d1, d2;
  double ring_u[B747O_RING], ring_t[B747O_RING];
  int32_t tail, head, last, size;
  uint8_t mem_ss, mem_cs, and_ss, and_cs, memout_ss, memout_cs;
  double sumA[5];            /* aero_err + 1 sums, held between major steps */
  double td;                 /* transport-delay output of the last pass */
} b747o_model;

void b747o_model_defaults(b747o_model *m);   /* .data defaults of the DLL */
void b747o_model_initialize(b747o_model *m); /* model_simple_initialize */
void b747o_model_step(b747o_model *m);       /* model_simple_step */

/* ------------------------------------------------------------------ */
/* Backend-neutral view of "a loaded model library": the three entry points
 * and pointers to the named double globals.  The DLL and the restatement
 * both fit, so ONE env-layer implementation is checked on the DLL and then
 * reused with the restatement for large batches. */
typedef struct b747o_iface {
  void *ctx;
  void (*initialize)(void *ctx);
  void (*step)(void *ctx);
  /* params */
  double *state0, *h_zh, *use_RP, *use_PID_SS, *use_PID_CS, *PID_SS, *PID_CS, *deltaz, *vartheta, *P, *aero_err;
  /* signals */
  double *state, *sim_time, *vartheta_zh, *U_com_PID, *CXa, *CYa, *mz, *K_alpha, *dCm_ddeltaz, *U_com,
      *deltaz_RP, *dvartheta, *dvartheta_int, *dvartheta_dt, *dvartheta_dt_dt, *ITSE;
} b747o_iface;

void b747o_iface_from_model(b747o_iface *f, b747o_model *m);

/* Enums: values equal the reference's Enum values
 * (core/controller.py:14-36, env/ctrl_env.py:16-30). */
enum { B747_CTRL_FULL_AUTO = 0, B747_CTRL_AUTO = 1, B747_CTRL_SEMI_MANUAL = 2, B747_CTRL_MANUAL = 3 };
enum { B747_MODE_DIRECT = 0, B747_MODE_ADD_PROC = 1, B747_MODE_ANG_VEL = 2, B747_MODE_ADD_DIRECT = 3 };
enum { B747_MODE_NONE = -1 }; /* ctrl_mode=None: action law as DIRECT (core/controller.py:241), no rf reward term (env/ctrl_env.py:141) */
enum { B747_RESET_NONE = -1, B747_RESET_CONST = 0, B747_RESET_OSCILLATING = 1, B747_RESET_HYBRID = 2 };
enum { B747_DIST_NONE = -1, B747_DIST_AERO = 0 };
enum { B747_OBS_PID_LIKE = 0, B747_OBS_SPEED_MODE = 1, B747_OBS_PID_AERO = 2, B747_OBS_PID_SPEED_AERO = 3, B747_OBS_MODEL_STATE = 4 };
enum { B747_REW_CLASSIC = 0, B747_REW_PID_LIKE = 1, B747_REW_QUALITY = 2, B747_REW_MINIMAL = 3, B747_REW_TF_REFERENCE = 4 };

/* Environment configuration: the ControllerEnv / Controller constructor arguments. */
typedef struct b747o_env_cfg {
  int32_t obs_type, rew_type, ctrl_type, ctrl_mode, reset_ref_mode, disturbance_mode;
  int32_t norm_obs, norm_act, use_limiter;
  int32_t substeps;      /* K = round(sample_time/dt), core/controller.py:261 */
  int64_t done_tick;     /* smallest tick with fl(tick*0.01) >= tk */
  double tk, action_max, vartheta_max, sample_time;
  double rew[8];         /* CLASSIC: k1,k2,k3(normalised),k0,kITSE,kf,kt,ko; others see b747_env_ref.c */
  double fixed_aero_err[5];
  int32_t has_fixed_aero_err;
  uint64_t seed;
} b747o_env_cfg;

/* Per-episode reference / initial condition (what Controller.reset decides). */
typedef struct b747o_episode {
  double state0[6];
  int32_t use_ctrl;       /* СУ PID in the loop (SEMI_MANUAL / FULL_AUTO) */
  double vref_const;      /* CONST / HYBRID: constant pitch reference */
  double osc_A[3], osc_f[3];
  int32_t oscillating;
  double h_ref;           /* altitude reference when use_ctrl */
  double aero_err[5];
} b747o_episode;

typedef struct b747o_env {
  b747o_env_cfg cfg;
  b747o_iface mdl;
  b747o_episode ep;
  uint64_t env_id, episode_idx;
  int64_t step_count;      /* env steps in this episode */
  double ep_return;
  double tf_tp;            /* TF_REFERENCE closure state (env/ctrl_env.py:178) */
  int32_t obs_dim;
  /* Controller(use_storage=True): what Controller._post_step records after every model step
   * (core/controller.py:209-228), rec[step][B747O_NREC]; reset clears it (core/controller.py:195-199). */
  double *rec;
  int32_t rec_cap, rec_n;
} b747o_env;
enum { B747O_REC_t = 0, B747O_REC_U_com, B747O_REC_U_PID, B747O_REC_deltaz, B747O_REC_hzh, B747O_REC_vartheta_ref,
       B747O_REC_U_RL, B747O_REC_x, B747O_REC_y, B747O_REC_Vx, B747O_REC_Vy, B747O_REC_vartheta, B747O_REC_wz, B747O_NREC };
void b747o_env_set_recorder(b747o_env *e, double *buf, int32_t capacity_steps);
int32_t b747o_env_recorded(const b747o_env *e);

int b747o_obs_dim(int obs_type);
int64_t b747o_done_tick(double tk);
void b747o_env_init(b747o_env *e, const b747o_env_cfg *cfg, const b747o_iface *mdl, uint64_t env_id);
/* Controller.reset with random draws (counter-based Philox keyed by seed/env/episode). */
void b747o_env_draw_episode(const b747o_env_cfg *cfg, uint64_t env_id, uint64_t episode_idx, b747o_episode *ep);
/* Controller.reset(state0) + ControllerEnv.reset: applies `ep`, initialises the model, writes obs. */
void b747o_env_reset_to(b747o_env *e, const b747o_episode *ep, double *obs);
void b747o_env_reset(b747o_env *e, double *obs);
/* ControllerEnv.step: action is the raw (possibly normalised) action; returns done. */
int b747o_env_step(b747o_env *e, double action, double *obs, double *reward);

/* Philox4x32-10 (shared definition with the CUDA path; checked by KATs in tests). */
void b747o_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double b747o_uniform53(uint64_t seed, uint64_t env_id, uint64_t episode_idx, uint32_t draw);

/* Batched drivers over the restatement (used by the GPU parity tests and the CPU baseline). */
typedef struct b747o_batch b747o_batch;
b747o_batch *b747o_batch_create(const b747o_env_cfg *cfg, int64_t n_envs, uint64_t env_id_offset);
void b747o_batch_destroy(b747o_batch *b);
void b747o_batch_reset(b747o_batch *b, double *obs);
void b747o_batch_reset_to(b747o_batch *b, const b747o_episode *eps, double *obs);
/* auto_reset: on done, obs <- reset obs and terminal_obs <- last obs (SB3 VecEnv contract). */
void b747o_batch_step(b747o_batch *b, const double *actions, double *obs, double *rew, uint8_t *done,
                      double *terminal_obs, int auto_reset);
b747o_env *b747o_batch_env(b747o_batch *b, int64_t i);
b747o_model *b747o_batch_model(b747o_batch *b, int64_t i);

#ifdef __cplusplus
}
#endif
#endif
