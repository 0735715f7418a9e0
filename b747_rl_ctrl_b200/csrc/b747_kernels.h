// b747_kernels.h -- host-visible launchers of the env-step kernels (internal to libb747_b200.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "b747_common.cuh"

namespace b747 {

// Device pointers of a float64 handle.  slots: [NSLOT_F64 + 6][n_pad] doubles (the last six rows are
// the DLL's `state0` parameter); sig: [NSIG][n_pad] stage-4 signal export or nullptr.
struct StateF64 {
  double* slots;
  int* tick;
  int* flags;
  uint32_t* ep_idx;
  double* sig;
  double* stats;     // [4] episodes, sum return, sum length, sum return^2
  double* last_ret;  // [n_pad] return of the most recently finished episode
  int* last_len;     // [n_pad] its length in env steps
  TraceState trace;  // recorder / step-response tracker (optional)
};

// Device pointers of an f32 (throughput) handle; layout in b747_kernels_f32.cu.
struct StateF32 {
  double2* D = nullptr;
  float4* F = nullptr;
  float* sig = nullptr;
  double* stats = nullptr;
  double* last_ret = nullptr;
  int* last_len = nullptr;
  TraceState trace;  // recorder / step-response tracker (optional; routes the launch to the general kernel)
  float4* tables = nullptr;  // merged-axis look-up tables (b747_tables.h), ft::CELLS float4
  unsigned int* tile_ctr = nullptr;  // 64 tile counters of the persistent step kernel (one per launch in flight)
};

int f32_alloc(const DevCfg& c, StateF32& s, bool export_signals, cudaStream_t stream);
void f32_free(StateF32& s);
bool f32_is_lean(const DevCfg& c);
bool f32_needs_cs(const DevCfg& c);
void f32_warm_launch();
// out4 != nullptr: packed outputs (one float4 = obs[3] + reward per env, done flags as one bit per env in done_bits);
// obs / rew / done / term_obs are then unused.  prefetch_actions: `actions` is host-mapped memory.
void launch_env_step32(const DevCfg& c, const StateF32& st, const float* actions, float* obs, float* rew, uint8_t* done,
                       float* term_obs, cudaStream_t s, float4* out4 = nullptr, uint32_t* done_bits = nullptr,
                       bool prefetch_actions = false);
void launch_reset32(const DevCfg& c, const StateF32& st, const uint8_t* mask, const b747_episode* eps, float* obs,
                    cudaStream_t s);
void launch_defaults32(const DevCfg& c, const StateF32& st, cudaStream_t s);
int f32_field_io(const DevCfg& c, StateF32& s, int kind, int row, const char* name, double* out, const double* in,
                 cudaStream_t stream);

void launch_env_step64(const DevCfg& c, const StateF64& st, const double* actions, double* obs, double* rew,
                       uint8_t* done, double* term_obs, cudaStream_t s);
void launch_model_step64(const DevCfg& c, const StateF64& st, int n_steps, cudaStream_t s);
void launch_reset64(const DevCfg& c, const StateF64& st, const uint8_t* mask, const b747_episode* eps, double* obs,
                    cudaStream_t s);
void launch_defaults64(const DevCfg& c, const StateF64& st, cudaStream_t s);
void launch_model_init64(const DevCfg& c, const StateF64& st, cudaStream_t s);
// calc_stepinfo's final arithmetic on the tracker (snapshot != 0: the last finished episode): out[env][5] =
// overshoot %, rise time, settling time, static error, Controller.quality (NaN where the reference returns None)
void launch_transfer_metrics(const DevCfg& c, const TraceState& tr, int which, int snapshot, double* out_dev,
                             cudaStream_t s);

}  // namespace b747
