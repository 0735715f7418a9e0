/*
 * b747_scalar_legacy.h -- the boundary of the reference's LEGACY dynamics library, core/model_win64.dll (March 2022,
 * MinGW build of an earlier revision of the same Simulink diagram), as exported by b747_rl_ctrl_b200/lib/model.so.
 *
 * What that DLL is (established by executing it next to core/model_simple_win64.dll, tests/test_legacy_model.py,
 * oracle/legacy.py): the SAME dynamics -- identical trajectories to ~1e-12 over 2000 steps with the same inputs, the
 * remaining difference being its MinGW x87 libm -- behind a different symbol surface:
 *   functions   model_initialize / model_step / model_terminate          (model_win64.dll exports 39-41)
 *   state[6], state0[6]   laid out as a 3-D position / velocity vector [x, y, z, Vx, Vy, Vz] (z = Vz = 0 always); the
 *                         initial pitch and pitch rate are not settable (0)
 *   deltaz_com / deltaz_real / deltaz_ref   = model_simple's U_com / deltaz_RP / U_com_PID
 *   CXa, CYa, mz, dCm_ddeltaz   tapped BEFORE the (1 + aero_err) gains; dCm_ddeltaz after the per-degree -> per-radian gain
 *   aero_err[4] (the K_alpha error) is not connected; there is no use_RP switch (the actuator is always in the loop)
 *   I[3] = (Ixx, Iyy, Izz): the pitch dynamics use Izz = 67.3e6, the value model_simple exports as Iz
 *   .data defaults: state0 = (0, 11000, 0, 259.1667, 0, 0), h_zh = 5000, vartheta = 0
 * Not provided: the Simulink-Coder internals the DLL also exports (model_P, model_X, model_DW, model_M, look1_binlx,
 * look2_binlx, rt_* helpers, model_GetCAPIStaticMap) -- no Python in the reference binds them.
 *
 * Lifecycle and threading as b747_scalar.h; compute runs on CUDA device $B747_DEVICE, no CPU path.
 */
#ifndef B747_SCALAR_LEGACY_H
#define B747_SCALAR_LEGACY_H
#ifdef __cplusplus
extern "C" {
#endif

void model_initialize(void); /* replaces model_win64.dll `model_initialize` */
void model_step(void);       /* replaces model_win64.dll `model_step` @0xe40 */
void model_terminate(void);

/* signals */
extern double state[6], sim_time, vartheta_zh, deltaz_ref, deltaz_com, deltaz_real, CXa, CYa, mz, K_alpha, dCm_ddeltaz,
    dvartheta, dvartheta_int, dvartheta_dt, dvartheta_dt_dt, TAE, ITAE, TSE, ITSE, AE, IAE, SE, ISE;
/* parameters */
extern double state0[6], h_zh, use_PID_SS, use_PID_CS, use_RL, PID_SS[4], PID_CS[4], deltaz, vartheta, P, aero_err[5];
extern double I[3], S, c_, g, m0;

#ifdef __cplusplus
}
#endif
#endif
