// b747_tables.h -- aerodynamic look-up tables of the throughput (f32) path, re-gridded on MERGED axes.
//
// The DLL interpolates four 2-D tables and one 1-D table with look2_binlx / look1 (dll@0x1000,
// inline @0x2083): CYa(Mach|P42.., alpha|P46..), CXa(Mach|P108.., CYa|P112..),
// dCm(h|P201.., Mach|P206..), mz(Mach|P276.., alpha|P280..), K_alpha(alpha|P225..).  Mach is searched on
// three different breakpoint sets and alpha on three others -- seven interval searches per diagram pass.
//
// A bilinear interpolant stays bilinear on every sub-cell of its grid, and look2_binlx extrapolates the
// end cells linearly, so re-sampling each table on the UNION of the breakpoint sets of an operand is
// exact: one Mach axis (16 intervals), one alpha axis (21), the altitude axis (4) and the CYa axis (13)
// serve all five tables, with one interval search and one fraction per operand.  The tables are
// re-sampled here in float64 from model_simple_P and rounded once to float32.
//
// Layout (float4 units, staged into shared memory by every block of k_env_step32):
//   axis record, one per interval i:      {lo, hi, bp[i], 1/(bp[i+1]-bp[i])}; lo of the first / hi of the
//                                         last interval are NaN (never "left": end cells extrapolate)
//   K_alpha record per alpha interval:    {t[i], t[i+1]-t[i], 0, 0}
//   2-D cell (i0 along axis 0, i1 axis 1): {t00, t10-t00, t01, t11-t01}
//   CYa and mz share both axes; their cells are interleaved (2 float4 per cell).
// Host-only code (no CUDA types); the device side reads it through the offsets below.
#pragma once
#include <math.h>
#include <stddef.h>

#include <algorithm>
#include <vector>

#include "../../include/b747_params.h"

namespace b747 {
namespace ft {

constexpr int NM = 16, NA = 21, NH = 4, NC = 13;  // intervals per merged axis (checked by build())
constexpr int AXM = 0, AXA = AXM + NM, AXH = AXA + NA, AXC = AXH + NH, KA = AXC + NC;  // float4 offsets
constexpr int T_HM = KA + NA;            // dCm cells  [iM][iH]
constexpr int T_MC = T_HM + NM * NH;     // CXa cells  [iC][iM]
constexpr int T_MA = T_MC + NC * NM;     // CYa,mz cells [iA][iM][2]
constexpr int CELLS = T_MA + 2 * NA * NM;  // 1019 float4 = 16304 bytes

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

struct Fast {
  std::vector<double> bM, bA, bH, bC;  // merged breakpoints
  std::vector<float> v;                // CELLS * 4 floats
  bool ok = false;
};

inline std::vector<double> merged(std::initializer_list<std::pair<const double*, int>> sets) {
  std::vector<double> r;
  for (auto& s : sets) r.insert(r.end(), s.first, s.first + s.second);
  std::sort(r.begin(), r.end());
  r.erase(std::unique(r.begin(), r.end()), r.end());
  return r;
}

inline Fast build() {
  const Orig O;
  const double* P = O.P;
  Fast F;
  F.bM = merged({{P + 42, 4}, {P + 108, 4}, {P + 206, 10}, {P + 276, 4}});
  F.bA = merged({{P + 46, 5}, {P + 225, 7}, {P + 280, 11}});
  F.bH.assign(P + 201, P + 206);
  F.bC.assign(P + 112, P + 126);
  if ((int)F.bM.size() != NM + 1 || (int)F.bA.size() != NA + 1 || (int)F.bH.size() != NH + 1 || (int)F.bC.size() != NC + 1)
    return F;  // ok == false: the parameter set does not fit the compiled layout
  F.v.assign((size_t)CELLS * 4, 0.f);
  const float qnan = nanf("");
  // `unit`: breakpoints are stored divided by it (the alpha axis is kept in radians: the DLL converts alpha to
  // degrees with the gain P21 before the look-up, the kernel skips that multiply)
  auto axis = [&](int off, const std::vector<double>& b, double unit) {
    const int n = (int)b.size() - 1;
    for (int i = 0; i < n; i++) {
      float* q = &F.v[(size_t)(off + i) * 4];
      q[0] = i == 0 ? qnan : (float)(b[i] / unit);
      q[1] = i == n - 1 ? qnan : (float)(b[i + 1] / unit);
      q[2] = (float)(b[i] / unit);
      q[3] = (float)(unit / (b[i + 1] - b[i]));
    }
  };
  axis(AXM, F.bM, 1.0); axis(AXA, F.bA, P[21]); axis(AXH, F.bH, 1.0); axis(AXC, F.bC, 1.0);
  for (int i = 0; i < NA; i++) {
    float* q = &F.v[(size_t)(KA + i) * 4];
    q[0] = (float)O.Ka(F.bA[i]); q[1] = (float)(O.Ka(F.bA[i + 1]) - O.Ka(F.bA[i]));
  }
  auto cell = [&](float* q, double t00, double t10, double t01, double t11) {
    q[0] = (float)t00; q[1] = (float)(t10 - t00); q[2] = (float)t01; q[3] = (float)(t11 - t01);
  };
  for (int iM = 0; iM < NM; iM++)
    for (int iH = 0; iH < NH; iH++)
      cell(&F.v[(size_t)(T_HM + iM * NH + iH) * 4], O.dCm(F.bH[iH], F.bM[iM]), O.dCm(F.bH[iH + 1], F.bM[iM]),
           O.dCm(F.bH[iH], F.bM[iM + 1]), O.dCm(F.bH[iH + 1], F.bM[iM + 1]));
  for (int iC = 0; iC < NC; iC++)
    for (int iM = 0; iM < NM; iM++)
      cell(&F.v[(size_t)(T_MC + iC * NM + iM) * 4], O.CXa(F.bM[iM], F.bC[iC]), O.CXa(F.bM[iM + 1], F.bC[iC]),
           O.CXa(F.bM[iM], F.bC[iC + 1]), O.CXa(F.bM[iM + 1], F.bC[iC + 1]));
  for (int iA = 0; iA < NA; iA++)
    for (int iM = 0; iM < NM; iM++) {
      float* q = &F.v[(size_t)(T_MA + 2 * (iA * NM + iM)) * 4];
      cell(q, O.CYa(F.bM[iM], F.bA[iA]), O.CYa(F.bM[iM + 1], F.bA[iA]), O.CYa(F.bM[iM], F.bA[iA + 1]),
           O.CYa(F.bM[iM + 1], F.bA[iA + 1]));
      cell(q + 4, O.mz(F.bM[iM], F.bA[iA]), O.mz(F.bM[iM + 1], F.bA[iA]), O.mz(F.bM[iM], F.bA[iA + 1]),
           O.mz(F.bM[iM + 1], F.bA[iA + 1]));
    }
  F.ok = true;
  return F;
}

// ---- evaluation of the fast layout on the host (float64 arithmetic on the float32 entries): what the
// kernel computes, minus its float32 rounding.  Used by b747_selftest_tables(). ----
inline int find(const std::vector<double>& b, double u) {
  const int n = (int)b.size() - 1;
  int i = 0;
  while (i < n - 1 && u >= b[i + 1]) i++;
  return i;
}
inline double frac(const Fast& F, int off, int i, double u) {
  const float* q = &F.v[(size_t)(off + i) * 4];
  return (u - (double)q[2]) * (double)q[3];
}
inline double bil(const float* c, double f0, double f1) {
  const double yL = c[0] + f0 * c[1], yR = c[2] + f0 * c[3];
  return yL + f1 * (yR - yL);
}
struct FastEval {
  const Fast& F;
  void eval(double M, double a, double h, double out[5]) const {  // a in degrees; CYa, CXa, dCm, mz, Ka
    const Orig O;
    const int iM = find(F.bM, M), iA = find(F.bA, a), iH = find(F.bH, h);
    const double fM = frac(F, AXM, iM, M), fA = frac(F, AXA, iA, a / O.P[21]), fH = frac(F, AXH, iH, h);
    const float* q = &F.v[(size_t)(T_MA + 2 * (iA * NM + iM)) * 4];
    out[0] = bil(q, fM, fA);
    out[3] = bil(q + 4, fM, fA);
    const int iC = find(F.bC, out[0]);
    out[1] = bil(&F.v[(size_t)(T_MC + iC * NM + iM) * 4], fM, frac(F, AXC, iC, out[0]));
    out[2] = bil(&F.v[(size_t)(T_HM + iM * NH + iH) * 4], fH, fM);
    const float* k = &F.v[(size_t)(KA + iA) * 4];
    out[4] = k[0] + fA * k[1];
  }
};

}  // namespace ft
}  // namespace b747
