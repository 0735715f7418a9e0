// b747_tables.h -- aerodynamic look-up tables of the throughput (f32) path, re-gridded on MERGED axes.
//
// The DLL interpolates four 2-D tables and one 1-D table with look2_binlx / look1 (dll@0x1000,
// inline @0x2083): CYa(Mach|P42.., alpha|P46..), CXa(Mach|P108.., CYa|P112..),
// dCm(h|P201.., Mach|P206..), mz(Mach|P276.., alpha|P280..), K_alpha(alpha|P225..).  Mach is searched on
// three different breakpoint sets and alpha on three others -- seven interval searches per diagram pass.
//
// A bilinear interpolant stays bilinear on every sub-cell of its grid, and look2_binlx extrapolates the
// end cells linearly (the end cell's own polynomial), so re-sampling each table on the UNION of the
// breakpoint sets of an operand is exact: one Mach axis, one alpha axis, the altitude axis and the CYa axis
// serve all five tables, with one interval and one offset per operand.  Each merged axis is also EXTENDED by
// one sentinel interval per open end (Mach up to 2, alpha -40..75 deg, h -4..24 km, CYa -4..4) whose cells
// carry the end cell's polynomial: inside the sentinels an operand never "leaves the table", so the kernel's
// interval cache (below) stays valid while the DLL would be extrapolating.  Beyond the sentinels the end
// cell is used as is (linear extrapolation, exactly look2_binlx) and merely never caches.
// The tables are re-sampled here in float64 from model_simple_P and rounded once to float32.
//
// Cells are stored as polynomial coefficients in the OFFSETS d = u - bp[i] from the interval's lower
// breakpoint (not in the 0..1 fractions):  y = (c0 + c1 d0) + (c2 + c3 d0) d1  -- three FMAs, no fraction
// multiply, no yR - yL subtraction.  The kernel keeps {bp, width} of the current interval of every axis and
// the coefficients of the current cells IN REGISTERS; a pass validates the cache with one unsigned compare
// per axis (as_uint(d) < as_uint(width)  <=>  0 <= d < width) and touches shared memory only when an
// operand crossed a breakpoint (Mach, alpha, h move ~1e-4 of an interval per pass).
//
// Layout (float4 units, staged into shared memory by every block of k_env_step32):
//   axis record, one per interval i:      {bp[i], bp[i+1]-bp[i], k0, k1}; (k0, k1) = K_alpha(alpha) = k0 + k1 d
//                                         on the alpha axis, 0 elsewhere.  The alpha axis is in radians.
//   2-D cell (d0 = Mach offset for every table, d1 = the other operand's offset): CXa and dCm cells are stored as
//   {c3, c1, c2, c0}: one packed FFMA2 (c3, c1) dM + (c2, c0) gives both inner terms, one FFMA finishes the table
//   The CXa cells carry the gain P126 of the drag path and the dCm cells the gains P217 P150 of the elevator-moment path
//   (constants of the diagram between the look-up and its only consumer), so the kernel multiplies neither.
//   CYa and mz share both axes; their cells are interleaved component-wise, {cy0, mz0, cy1, mz1} {cy2, mz2, cy3, mz3},
//   so that two 128-bit loads fill four aligned register pairs for packed FFMA2 evaluation of both tables at once.
//   coarse index map per axis: LUT_N bytes, bucket b of the flight-envelope range -> interval holding the bucket's lower
//                                         edge; the first search of a launch starts there (the kernel then walks 0..2 steps)
// Host-only code (no CUDA types); the device side reads it through the offsets below.
#pragma once
#include <math.h>
#include <stddef.h>

#include <algorithm>
#include <vector>

#include "../../include/b747_params.h"

namespace b747 {
namespace ft {

constexpr int NM = 17, NA = 23, NH = 6, NC = 15;  // intervals per merged + extended axis (checked by build())
constexpr int AXM = 0, AXA = AXM + NM, AXH = AXA + NA, AXC = AXH + NH;  // float4 offsets
constexpr int T_HM = AXC + NC;           // dCm cells  [iM][iH]      d0 = Mach offset, d1 = h offset
constexpr int T_MC = T_HM + NM * NH;     // CXa cells  [iC][iM]      d0 = Mach offset, d1 = CYa offset
constexpr int T_MA = T_MC + NC * NM;     // CYa,mz cells [iA][iM][2] d0 = Mach offset, d1 = alpha offset (rad)
constexpr int LUT = T_MA + 2 * NA * NM;   // 4 coarse index maps (Mach, alpha, h, CYa), LUT_N bytes each
constexpr int LUT_N = 64;
constexpr int CELLS = LUT + 4 * LUT_N / 16;  // 1216 float4 = 19456 bytes
// sentinel breakpoints (alpha in degrees, like the DLL's tables)
constexpr double EXT_M_HI = 2.0, EXT_A_LO = -40.0, EXT_A_HI = 75.0, EXT_H_LO = -4000.0, EXT_H_HI = 24000.0, EXT_C_LO = -4.0,
                 EXT_C_HI = 4.0;
// ranges the coarse index maps resolve (flight envelope; operands outside start from the end bucket and walk)
constexpr double LUT_M_LO = 0.0, LUT_M_HI = 1.0, LUT_A_LO = -10.0, LUT_A_HI = 35.0, LUT_H_LO = -4000.0, LUT_H_HI = 24000.0,
                 LUT_C_LO = -0.4, LUT_C_HI = 1.6;

// ---- float64 restatement of the DLL's interpolation (test oracle for the re-gridding) ----
inline void prelook(double u, const double* bp, int maxIndex, int& idx, double& frac) {
  if (u <= bp[0]) { idx = 0; frac = (u - bp[0]) / (bp[1] - bp[0]); return; }
  if (u < bp[maxIndex]) {
    int lo = 0, hi = maxIndex;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (u < bp[mid]) hi = mid; else lo = mid; }
    idx = lo; frac = (u - bp[lo]) / (bp[lo + 1] - bp[lo]); return;
  }
  idx = maxIndex - 1; frac = (u - bp[maxIndex - 1]) / (bp[maxIndex] - bp[maxIndex - 1]);
}
inline double look2(double u0, double u1, const double* bp0, int n0, const double* bp1, int n1, const double* tab) {
  int i0, i1; double f0, f1;
  prelook(u0, bp0, n0 - 1, i0, f0);
  prelook(u1, bp1, n1 - 1, i1, f1);
  const double* p = tab + i1 * n0 + i0;
  const double yL = p[0] + f0 * (p[1] - p[0]);
  const double yR = p[n0] + f0 * (p[n0 + 1] - p[n0]);
  return yL + f1 * (yR - yL);
}
inline double look1(double u, const double* bp, int n, const double* tab) {
  int i; double f;
  prelook(u, bp, n - 1, i, f);
  return tab[i] + f * (tab[i + 1] - tab[i]);
}

struct Orig {
  double P[B747_NP];
  Orig() { const double p[B747_NP] = B747_P_INIT; for (int i = 0; i < B747_NP; i++) P[i] = p[i]; }
  double CYa(double M, double a) const { return look2(M, a, P + 42, 4, P + 46, 5, P + 22); }
  double CXa(double M, double cy) const { return look2(M, cy, P + 108, 4, P + 112, 14, P + 52); }
  double dCm(double h, double M) const { return look2(h, M, P + 201, 5, P + 206, 10, P + 151); }
  double mz(double M, double a) const { return look2(M, a, P + 276, 4, P + 280, 11, P + 232); }
  double Ka(double a) const { return look1(a, P + 225, 7, P + 218); }
};

// gains folded into the CXa / dCm cells (model_simple_P indices as in SURVEY.md Appendix A)
inline double gain_cxa(const double* P) { return P[126]; }
inline double gain_dcm(const double* P) { return P[217] * P[150]; }

struct Fast {
  std::vector<double> bM, bA, bH, bC;  // merged + extended breakpoints (alpha in degrees)
  std::vector<float> v;                // CELLS * 4 floats
  bool ok = false;
};

inline std::vector<double> merged(std::initializer_list<std::pair<const double*, int>> sets, double ext_lo, double ext_hi) {
  std::vector<double> r;
  for (auto& s : sets) r.insert(r.end(), s.first, s.first + s.second);
  std::sort(r.begin(), r.end());
  r.erase(std::unique(r.begin(), r.end()), r.end());
  if (ext_lo < r.front()) r.insert(r.begin(), ext_lo);
  if (ext_hi > r.back()) r.push_back(ext_hi);
  return r;
}

inline Fast build() {
  const Orig O;
  const double* P = O.P;
  Fast F;
  // Mach is a speed ratio: the merged axis starts at its first breakpoint 0, no lower sentinel
  F.bM = merged({{P + 42, 4}, {P + 108, 4}, {P + 206, 10}, {P + 276, 4}}, 0.0, EXT_M_HI);
  F.bA = merged({{P + 46, 5}, {P + 225, 7}, {P + 280, 11}}, EXT_A_LO, EXT_A_HI);
  F.bH = merged({{P + 201, 5}}, EXT_H_LO, EXT_H_HI);
  F.bC = merged({{P + 112, 14}}, EXT_C_LO, EXT_C_HI);
  if ((int)F.bM.size() != NM + 1 || (int)F.bA.size() != NA + 1 || (int)F.bH.size() != NH + 1 || (int)F.bC.size() != NC + 1)
    return F;  // ok == false: the parameter set does not fit the compiled layout
  F.v.assign((size_t)CELLS * 4, 0.f);
  // `unit`: breakpoints are stored divided by it (the alpha axis is kept in radians: the DLL converts alpha to
  // degrees with the gain P21 before the look-up, the kernel skips that multiply)
  const double rad = P[21];
  auto axis = [&](int off, const std::vector<double>& b, double unit) {
    const int n = (int)b.size() - 1;
    for (int i = 0; i < n; i++) {
      float* q = &F.v[(size_t)(off + i) * 4];
      q[0] = (float)(b[i] / unit);
      q[1] = (float)((b[i + 1] - b[i]) / unit);
    }
  };
  axis(AXM, F.bM, 1.0); axis(AXA, F.bA, rad); axis(AXH, F.bH, 1.0); axis(AXC, F.bC, 1.0);
  for (int i = 0; i < NA; i++) {  // K_alpha rides in the alpha axis record
    float* q = &F.v[(size_t)(AXA + i) * 4];
    const double k0 = O.Ka(F.bA[i]), k1 = O.Ka(F.bA[i + 1]);
    q[2] = (float)k0; q[3] = (float)((k1 - k0) / ((F.bA[i + 1] - F.bA[i]) / rad));
  }
  // corner values -> coefficients in the offsets; w0, w1 = interval widths in the kernel's units
  auto cell = [&](float* q, double t00, double t10, double t01, double t11, double w0, double w1) {
    q[0] = (float)t00; q[1] = (float)((t10 - t00) / w0); q[2] = (float)((t01 - t00) / w1);
    q[3] = (float)(((t11 - t01) - (t10 - t00)) / (w0 * w1));
  };
  auto wd = [](const std::vector<double>& b, int i, double unit) { return (b[i + 1] - b[i]) / unit; };
  auto cell_inner = [&](float* q, double gain, double t00, double t10, double t01, double t11, double w0, double w1) {
    float c[4];
    cell(c, gain * t00, gain * t10, gain * t01, gain * t11, w0, w1);
    q[0] = c[3]; q[1] = c[1]; q[2] = c[2]; q[3] = c[0];
  };
  for (int iM = 0; iM < NM; iM++)
    for (int iH = 0; iH < NH; iH++)
      cell_inner(&F.v[(size_t)(T_HM + iM * NH + iH) * 4], gain_dcm(P), O.dCm(F.bH[iH], F.bM[iM]), O.dCm(F.bH[iH], F.bM[iM + 1]),
           O.dCm(F.bH[iH + 1], F.bM[iM]), O.dCm(F.bH[iH + 1], F.bM[iM + 1]), wd(F.bM, iM, 1.0), wd(F.bH, iH, 1.0));
  for (int iC = 0; iC < NC; iC++)
    for (int iM = 0; iM < NM; iM++)
      cell_inner(&F.v[(size_t)(T_MC + iC * NM + iM) * 4], gain_cxa(P), O.CXa(F.bM[iM], F.bC[iC]), O.CXa(F.bM[iM + 1], F.bC[iC]),
           O.CXa(F.bM[iM], F.bC[iC + 1]), O.CXa(F.bM[iM + 1], F.bC[iC + 1]), wd(F.bM, iM, 1.0), wd(F.bC, iC, 1.0));
  for (int iA = 0; iA < NA; iA++)
    for (int iM = 0; iM < NM; iM++) {
      float* q = &F.v[(size_t)(T_MA + 2 * (iA * NM + iM)) * 4];
      float cy[4], mz[4];
      cell(cy, O.CYa(F.bM[iM], F.bA[iA]), O.CYa(F.bM[iM + 1], F.bA[iA]), O.CYa(F.bM[iM], F.bA[iA + 1]),
           O.CYa(F.bM[iM + 1], F.bA[iA + 1]), wd(F.bM, iM, 1.0), wd(F.bA, iA, rad));
      cell(mz, O.mz(F.bM[iM], F.bA[iA]), O.mz(F.bM[iM + 1], F.bA[iA]), O.mz(F.bM[iM], F.bA[iA + 1]),
           O.mz(F.bM[iM + 1], F.bA[iA + 1]), wd(F.bM, iM, 1.0), wd(F.bA, iA, rad));
      for (int k = 0; k < 4; k++) { q[2 * k] = cy[k]; q[2 * k + 1] = mz[k]; }
    }
  auto lut = [&](int k, const std::vector<double>& b, double lo, double hi) {
    unsigned char* L = reinterpret_cast<unsigned char*>(&F.v[(size_t)LUT * 4]) + k * LUT_N;
    const int n = (int)b.size() - 1;
    for (int j = 0; j < LUT_N; j++) {
      const double u = lo + (hi - lo) * j / LUT_N;
      int i = 0;
      while (i < n - 1 && u >= b[i + 1]) i++;
      L[j] = (unsigned char)i;
    }
  };
  lut(0, F.bM, LUT_M_LO, LUT_M_HI); lut(1, F.bA, LUT_A_LO, LUT_A_HI); lut(2, F.bH, LUT_H_LO, LUT_H_HI);
  lut(3, F.bC, LUT_C_LO, LUT_C_HI);
  F.ok = true;
  return F;
}

// ---- evaluation of the fast layout on the host (float64 arithmetic on the float32 entries): what the
// kernel computes, minus its float32 rounding.  Used by b747_selftest_tables(). ----
// interval search on the float32 records, as the kernel's axis_find does: largest i with bp[i] <= u
inline int find(const Fast& F, int off, int n, double u) {
  int i = 0;
  for (int j = 1; j < n; j++)
    if (u >= (double)F.v[(size_t)(off + j) * 4]) i = j;
  return i;
}
inline double bil(const float* c, double d0, double d1) {
  return ((double)c[0] + (double)c[1] * d0) + ((double)c[2] + (double)c[3] * d0) * d1;
}
inline double bil_inner(const float* q, double d0, double d1) {  // {c3, c1, c2, c0} cells
  return ((double)q[3] + (double)q[1] * d0) + ((double)q[2] + (double)q[0] * d0) * d1;
}
struct FastEval {
  const Fast& F;
  void eval(double M, double a, double h, double out[5]) const {  // a in degrees; CYa, CXa, dCm, mz, Ka
    const Orig O;
    const double ar = a / O.P[21];
    const int iM = find(F, AXM, NM, M), iA = find(F, AXA, NA, ar), iH = find(F, AXH, NH, h);
    auto off = [&](int ax, int i, double u) { return u - (double)F.v[(size_t)(ax + i) * 4]; };
    const double dM = off(AXM, iM, M), dA = off(AXA, iA, ar), dH = off(AXH, iH, h);
    const float* q = &F.v[(size_t)(T_MA + 2 * (iA * NM + iM)) * 4];
    const float cy[4] = {q[0], q[2], q[4], q[6]}, mz[4] = {q[1], q[3], q[5], q[7]};
    out[0] = bil(cy, dM, dA);
    out[3] = bil(mz, dM, dA);
    const int iC = find(F, AXC, NC, out[0]);
    out[1] = bil_inner(&F.v[(size_t)(T_MC + iC * NM + iM) * 4], dM, off(AXC, iC, out[0])) / gain_cxa(O.P);
    out[2] = bil_inner(&F.v[(size_t)(T_HM + iM * NH + iH) * 4], dM, dH) / gain_dcm(O.P);
    const float* k = &F.v[(size_t)(AXA + iA) * 4];
    out[4] = (double)k[2] + dA * (double)k[3];
  }
};

}  // namespace ft
}  // namespace b747
