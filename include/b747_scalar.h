/*
 * b747_scalar.h -- the reference's scalar plugin boundary, as exported by model_simple.so.
 *
 * These are exactly the symbols core/model.py binds with ctypes:
 *   functions  core/model.py:124-126   {model}_initialize / {model}_step / {model}_terminate, void(void)
 *   signals    core/model.py:129-151   read-only doubles, hold the values of the 4th RK stage after a step
 *   parameters core/model.py:154-164   read/write doubles
 * plus the exports the DLL has but Python leaves unbound (Iz S c_ g m0 use_RL alpha V Mach).
 * Lifecycle (same as the DLL): write parameters -> model_simple_initialize() (re-reads state0, zeroes
 * time and signals) -> { write deltaz / vartheta / h_zh ; model_simple_step() }* -> read signals.
 * Process-global, non-reentrant; copy the library file to get an independent instance
 * (core/model.py:99-110).  Compute runs on CUDA device $B747_DEVICE (default 0); no CPU path.
 */
#ifndef B747_SCALAR_H
#define B747_SCALAR_H
#ifdef __cplusplus
extern "C" {
#endif

void model_simple_initialize(void); /* replaces dll@0x12a0 */
void model_simple_step(void);       /* replaces dll@0x16d0 (+ ode4 dll@0x2c60) */
void model_simple_terminate(void);  /* replaces dll@0x29d0 (no-op) */

/* signals */
extern double state[6], sim_time, vartheta_zh, U_com_PID, CXa, CYa, mz, K_alpha, dCm_ddeltaz, U_com, deltaz_RP,
    dvartheta, dvartheta_int, dvartheta_dt, dvartheta_dt_dt, TAE, ITAE, TSE, ITSE, AE, IAE, SE, ISE, alpha, V, Mach;
/* parameters */
extern double state0[6], h_zh, use_RP, use_PID_SS, use_PID_CS, PID_SS[4], PID_CS[4], deltaz, vartheta, P, aero_err[5];
extern double Iz, S, c_, g, m0, use_RL;

#ifdef __cplusplus
}
#endif
#endif
