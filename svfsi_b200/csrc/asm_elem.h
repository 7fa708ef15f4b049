// asm_elem.h -- the per-element part of CONSTRUCT_FLUID shared by the assembly kernels (asm_kernels.cu,
// asm_visit.cu): GNN for TET4 (S/NN.f:1515-1561) and the four Gauss points of FLUID3D_M / FLUID3D_C
// (S/FLUID.f:192-560, 813-1084) reduced to per-element sums.
#pragma once
#include <cuda_runtime.h>
#include <float.h>

#include "kernels.h"

namespace svfsi {

static constexpr int NE = 128;       // elements per CTA
static constexpr int NEP = NE + 1;   // padded field stride in shared memory (bank spread)


// S/UTIL.f:879-903 ISZERO(x) with one argument
__device__ __forceinline__ bool iszero1(double x) {
  const double a = fabs(x);
  const double nrm = a > DBL_EPSILON ? a : DBL_EPSILON;
  return a / nrm < 10.0 * DBL_EPSILON;
}

// GNN for TET4, S/NN.f:1515-1561: Jacobian, inverse, metric ks, Nx.  FAST: one reciprocal of Jac
// instead of nine divisions (1 ulp differences; parity tolerance is 1e-12).
template <bool FAST = false>
__device__ __forceinline__ void gnn_tet4(const double xl[4][3], double Nx[4][3], double &Jac,
                                         double ks[3][3]) {
  double X[3][3], XI[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) X[r][c] = xl[c][r] - xl[3][r];
  Jac = X[0][0] * X[1][1] * X[2][2] + X[0][1] * X[1][2] * X[2][0] + X[0][2] * X[1][0] * X[2][1] -
        X[0][0] * X[1][2] * X[2][1] - X[0][1] * X[1][0] * X[2][2] - X[0][2] * X[1][1] * X[2][0];
  XI[0][0] = (X[1][1] * X[2][2] - X[1][2] * X[2][1]);
  XI[0][1] = (X[2][1] * X[0][2] - X[2][2] * X[0][1]);
  XI[0][2] = (X[0][1] * X[1][2] - X[0][2] * X[1][1]);
  XI[1][0] = (X[1][2] * X[2][0] - X[1][0] * X[2][2]);
  XI[1][1] = (X[2][2] * X[0][0] - X[2][0] * X[0][2]);
  XI[1][2] = (X[0][2] * X[1][0] - X[0][0] * X[1][2]);
  XI[2][0] = (X[1][0] * X[2][1] - X[1][1] * X[2][0]);
  XI[2][1] = (X[2][0] * X[0][1] - X[2][1] * X[0][0]);
  XI[2][2] = (X[0][0] * X[1][1] - X[0][1] * X[1][0]);
  const double rJ = FAST ? 1.0 / Jac : 0.0;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) XI[r][c] = FAST ? XI[r][c] * rJ : XI[r][c] / Jac;
  ks[0][0] = XI[0][0] * XI[0][0] + XI[1][0] * XI[1][0] + XI[2][0] * XI[2][0];
  ks[0][1] = XI[0][1] * XI[0][0] + XI[1][1] * XI[1][0] + XI[2][1] * XI[2][0];
  ks[0][2] = XI[0][2] * XI[0][0] + XI[1][2] * XI[1][0] + XI[2][2] * XI[2][0];
  ks[1][1] = XI[0][1] * XI[0][1] + XI[1][1] * XI[1][1] + XI[2][1] * XI[2][1];
  ks[1][2] = XI[0][1] * XI[0][2] + XI[1][1] * XI[1][2] + XI[2][1] * XI[2][2];
  ks[2][2] = XI[0][2] * XI[0][2] + XI[1][2] * XI[1][2] + XI[2][2] * XI[2][2];
  ks[1][0] = ks[0][1];
  ks[2][0] = ks[0][2];
  ks[2][1] = ks[1][2];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    Nx[0][c] = XI[0][c];
    Nx[1][c] = XI[1][c];
    Nx[2][c] = XI[2][c];
    Nx[3][c] = -XI[0][c] - XI[1][c] - XI[2][c];
  }
}

// ---------------------------------------------------------------------------
// Phase 1 for one element: gather, GNN, the four Gauss points reduced to the compact record.
// rec[f * NEP] receives field f (the caller's shared-memory slot).
// Gauss-point sums of one element, before they are folded into the 80-double record
struct ElemAcc {
  double Nx[4][3];
  double A[4][4], c2[4], r2[4], sTC, sTM;
  double sRM[3][3], sNrV[4][3], lR4[4];
  double w, wl, wr;
};

template <bool FAST = false>
__device__ __forceinline__ void fluid_elem_compute(const FluidPar &par, int e,
                                                   const int *__restrict__ ien,
                                                   const double *__restrict__ x,
                                                   const double *__restrict__ Ag,
                                                   const double *__restrict__ Yg,
                                                   const double *__restrict__ Bf, ElemAcc &acc,
                                                   int *nodeOut, int *__restrict__ badJac) {
    const double gs = (5.0 + 3.0 * sqrt(5.0)) / 20.0, gt = (5.0 - sqrt(5.0)) / 20.0;
    int nd[4];
    {
      const int4 v = __ldg((const int4 *)ien + e);
      nd[0] = v.x; nd[1] = v.y; nd[2] = v.z; nd[3] = v.w;
    }
    double xl[4][3], al[4][3], yl[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      nodeOut[a] = nd[a];
      const double *xp = x + (size_t)nd[a] * 3;
      xl[a][0] = __ldg(xp); xl[a][1] = __ldg(xp + 1); xl[a][2] = __ldg(xp + 2);
      const double2 *ap = (const double2 *)(Ag + (size_t)nd[a] * 4);
      const double2 a01 = __ldg(ap), a23 = __ldg(ap + 1);
      al[a][0] = a01.x; al[a][1] = a01.y; al[a][2] = a23.x;
      const double2 *yp = (const double2 *)(Yg + (size_t)nd[a] * 4);
      const double2 y01 = __ldg(yp), y23 = __ldg(yp + 1);
      yl[a][0] = y01.x; yl[a][1] = y01.y; yl[a][2] = y23.x; yl[a][3] = y23.y;
      if (Bf) {  // ud uses al - bfl (S/FLUID.f:236-238)
        const double *bp = Bf + (size_t)nd[a] * 3;
        al[a][0] = al[a][0] - __ldg(bp);
        al[a][1] = al[a][1] - __ldg(bp + 1);
        al[a][2] = al[a][2] - __ldg(bp + 2);
      }
    }
    double Nx[4][3], Jac, ks[3][3];
    gnn_tet4<FAST>(xl, Nx, Jac, ks);
    if (iszero1(Jac)) atomicAdd(badJac, 1);

    const double rho = par.rho, mu = par.mu;
    const double T1c = par.af * par.gam * par.dt;
    const double amd = par.am / T1c;
    const double w = (1.0 / 24.0) * Jac;
    const double wl = w * T1c;
    const double wr = w * rho;

    // element constants: velocity gradient ux(j,i) = d u_i / d x_j, pressure gradient
    double ux[3][3], px[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      px[j] = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++) ux[j][i] = 0.0;
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        px[j] = px[j] + Nx[a][j] * yl[a][3];
#pragma unroll
        for (int i = 0; i < 3; i++) ux[j][i] = ux[j][i] + Nx[a][j] * yl[a][i];
      }
    }
    const double divU = ux[0][0] + ux[1][1] + ux[2][2];
    double es[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) es[j][i] = ux[j][i] + ux[i][j];

    double tq = 1.0 / par.dt;
    const double kT = 4.0 * (tq * tq);
    double kS = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) kS = kS + ks[j][i] * ks[j][i];
    tq = mu / rho;
    kS = 36.0 * kS * (tq * tq);
    const double trks = ks[0][0] + ks[1][1] + ks[2][2];
    // FAST: tauM = rsqrt(.)/rho, tauC = rho sqrt(.)/tr(ks), tauB = rho rsqrt(.) -- two rsqrt per
    // Gauss point instead of two sqrt + three divisions (S/FLUID.f:376-412 up to 1-2 ulp)
    const double irho = FAST ? 1.0 / rho : 0.0, rho_itrks = FAST ? rho / trks : 0.0;

    // accumulators over the Gauss points
    double A[4][4], c2[4], r2[4], sTC = 0.0, sTM = 0.0;
    double sRM[3][3], sNrV[4][3], lR4[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      c2[a] = 0.0; r2[a] = 0.0; lR4[a] = 0.0;
#pragma unroll
      for (int b = 0; b < 4; b++) A[a][b] = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++) sNrV[a][i] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) sRM[j][i] = 0.0;

#pragma unroll 1
    for (int g = 0; g < 4; g++) {
      // N(:,g): S/NN.f:268-275 (xi) and :654-658 (N4 = 1 - xi1 - xi2 - xi3 as computed)
      double Ng[4];
      Ng[0] = (g == 0) ? gs : gt;
      Ng[1] = (g == 1) ? gs : gt;
      Ng[2] = (g == 2) ? gs : gt;
      Ng[3] = 1.0 - Ng[0] - Ng[1] - Ng[2];
      double ud[3], u[3], p = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++) { ud[i] = -par.f[i]; u[i] = 0.0; }
#pragma unroll
      for (int a = 0; a < 4; a++) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
          ud[i] = ud[i] + Ng[a] * al[a][i];
          u[i] = u[i] + Ng[a] * yl[a][i];
        }
        p = p + Ng[a] * yl[a][3];
      }
      double kU = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) kU = kU + u[j] * u[i] * ks[j][i];
      const double kSum = kT + kU + kS;
      const double rsK = FAST ? rsqrt(kSum) : 0.0;
      const double tauM = FAST ? rsK * irho : 1.0 / (rho * sqrt(kSum));
      double rV[3], up[3], ua[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        rV[i] = ud[i] + u[0] * ux[0][i] + u[1] * ux[1][i] + u[2] * ux[2][i];
        up[i] = -tauM * (rho * rV[i] + px[i]);
      }
      const double tauC = FAST ? (kSum * rsK) * rho_itrks : 1.0 / (tauM * trks);
      double tauB = 0.0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) tauB = tauB + up[j] * up[i] * ks[j][i];
      if (iszero1(tauB)) tauB = DBL_EPSILON;
      tauB = FAST ? rho * rsqrt(tauB) : rho / sqrt(tauB);
#pragma unroll
      for (int i = 0; i < 3; i++) ua[i] = u[i] + up[i];
      const double pa = p - tauC * divU;
#pragma unroll
      for (int i = 0; i < 3; i++)
        rV[i] = tauB * (up[0] * ux[0][i] + up[1] * ux[1][i] + up[2] * ux[2][i]);
      // rM(j,i), S/FLUID.f:427-439, summed over g
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          double v = mu * es[j][i] - rho * up[i] * ua[j] + rV[i] * up[j];
          if (i == j) v = v - pa;
          sRM[j][i] = sRM[j][i] + v;
        }
#pragma unroll
      for (int i = 0; i < 3; i++)
        rV[i] = ud[i] + ua[0] * ux[0][i] + ua[1] * ux[1][i] + ua[2] * ux[2][i];

      double uNx[4], upNx[4], uaNx[4];
#pragma unroll
      for (int a = 0; a < 4; a++) {
        uNx[a] = u[0] * Nx[a][0] + u[1] * Nx[a][1] + u[2] * Nx[a][2];
        upNx[a] = up[0] * Nx[a][0] + up[1] * Nx[a][1] + up[2] * Nx[a][2];
        uaNx[a] = uNx[a] + upNx[a];
#pragma unroll
        for (int i = 0; i < 3; i++) sNrV[a][i] = sNrV[a][i] + Ng[a] * rV[i];
        // continuity residual, S/FLUID.f:1046-1049
        lR4[a] = lR4[a] + (Ng[a] * divU - upNx[a]);
        c2[a] = c2[a] + tauM * uaNx[a];
        r2[a] = r2[a] + tauM * (uNx[a] + amd * Ng[a]);
      }
      sTC = sTC + tauC;
      sTM = sTM + tauM;
      // diagonal-term scalar of the momentum tangent (S/FLUID.f:497-498 plus the
      // -rho tauM uaNx_a updu(i,i,b) part with updu(i,i,b) = -rho uNx_b)
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int a = 0; a < 4; a++)
          A[a][b] = A[a][b] + (rho * amd * Ng[b] * (Ng[a] + rho * tauM * uaNx[a]) +
                               rho * Ng[a] * (uNx[b] + upNx[b]) + tauB * upNx[a] * upNx[b] +
                               rho * tauM * uaNx[a] * (rho * uNx[b]));
    }

    // hand the sums to the caller
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
      for (int i = 0; i < 3; i++) { acc.Nx[a][i] = Nx[a][i]; acc.sNrV[a][i] = sNrV[a][i]; }
#pragma unroll
      for (int b = 0; b < 4; b++) acc.A[a][b] = A[a][b];
      acc.c2[a] = c2[a]; acc.r2[a] = r2[a]; acc.lR4[a] = lR4[a];
    }
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) acc.sRM[j][i] = sRM[j][i];
    acc.sTC = sTC; acc.sTM = sTM; acc.w = w; acc.wl = wl; acc.wr = wr;
}

// ---------------------------------------------------------------------------
// The same reduction with the algebra regrouped (results differ from fluid_elem_compute by re-association
// only, ~1e-16 relative):
//  * A_ab = sum_g X_a Y_b + Z_a upNx_b with X_a = N_a + rho tauM uaNx_a, Y_b = rho (amd N_b + uNx_b),
//    Z_a = rho N_a + tauB upNx_a: the four-term expression of S/FLUID.f:497-498 plus the tauM updu part is this
//    rank-two form exactly (the two rho N_a uNx_b terms cancel): 2 FMA per (a, b, g) instead of ~6 operations;
//  * the viscous part mu es(j,i) of rM is the same at the four Gauss points: 4 mu es is added once;
//  * w_i = sum_j up_j ux(j,i) serves both tauB-scaled rV and the second rV (= rV + w);
//  * u.ks.u and up.ks.up through one matrix-vector product each.
// UNR = unroll factor of the Gauss-point loop (1, 2 or 4).
template <int UNR>
__device__ __forceinline__ void fluid_elem_compute2(const FluidPar &par, int e,
                                                    const int *__restrict__ ien,
                                                    const double *__restrict__ x,
                                                    const double *__restrict__ Ag,
                                                    const double *__restrict__ Yg,
                                                    const double *__restrict__ Bf, ElemAcc &acc,
                                                    int *nodeOut, int *__restrict__ badJac) {
  const double gs = (5.0 + 3.0 * sqrt(5.0)) / 20.0, gt = (5.0 - sqrt(5.0)) / 20.0;
  int nd[4];
  {
    const int4 v = __ldg((const int4 *)ien + e);
    nd[0] = v.x; nd[1] = v.y; nd[2] = v.z; nd[3] = v.w;
  }
  double xl[4][3], al[4][3], yl[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    nodeOut[a] = nd[a];
    const double *xp = x + (size_t)nd[a] * 3;
    xl[a][0] = __ldg(xp); xl[a][1] = __ldg(xp + 1); xl[a][2] = __ldg(xp + 2);
    const double2 *ap = (const double2 *)(Ag + (size_t)nd[a] * 4);
    const double2 a01 = __ldg(ap), a23 = __ldg(ap + 1);
    al[a][0] = a01.x; al[a][1] = a01.y; al[a][2] = a23.x;
    const double2 *yp = (const double2 *)(Yg + (size_t)nd[a] * 4);
    const double2 y01 = __ldg(yp), y23 = __ldg(yp + 1);
    yl[a][0] = y01.x; yl[a][1] = y01.y; yl[a][2] = y23.x; yl[a][3] = y23.y;
    if (Bf) {
      const double *bp = Bf + (size_t)nd[a] * 3;
      al[a][0] = al[a][0] - __ldg(bp);
      al[a][1] = al[a][1] - __ldg(bp + 1);
      al[a][2] = al[a][2] - __ldg(bp + 2);
    }
  }
  double Nx[4][3], Jac, ks[3][3];
  gnn_tet4<true>(xl, Nx, Jac, ks);
  if (iszero1(Jac)) atomicAdd(badJac, 1);

  const double rho = par.rho, mu = par.mu;
  const double T1c = par.af * par.gam * par.dt;
  const double amd = par.am / T1c;
  const double w = (1.0 / 24.0) * Jac;

  double ux[3][3], px[3];
#pragma unroll
  for (int j = 0; j < 3; j++) {
    px[j] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) ux[j][i] = 0.0;
  }
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      px[j] = fma(Nx[a][j], yl[a][3], px[j]);
#pragma unroll
      for (int i = 0; i < 3; i++) ux[j][i] = fma(Nx[a][j], yl[a][i], ux[j][i]);
    }
  const double divU = ux[0][0] + ux[1][1] + ux[2][2];
  double tq = 1.0 / par.dt;
  const double kT = 4.0 * (tq * tq);
  double kS = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) kS = fma(ks[j][i], ks[j][i], kS);
  tq = mu / rho;
  kS = 36.0 * kS * (tq * tq);
  const double trks = ks[0][0] + ks[1][1] + ks[2][2];
  const double irho = 1.0 / rho, rho_itrks = rho / trks;

  double A[4][4], c2[4], r2[4], sTC = 0.0, sTM = 0.0, spa = 0.0;
  double sRM[3][3], sNrV[4][3], lR4[4];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    c2[a] = 0.0; r2[a] = 0.0; lR4[a] = 0.0;
#pragma unroll
    for (int b = 0; b < 4; b++) A[a][b] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) sNrV[a][i] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) sRM[j][i] = 0.0;

#pragma unroll UNR
  for (int g = 0; g < 4; g++) {
    double Ng[4];
    Ng[0] = (g == 0) ? gs : gt;
    Ng[1] = (g == 1) ? gs : gt;
    Ng[2] = (g == 2) ? gs : gt;
    Ng[3] = 1.0 - Ng[0] - Ng[1] - Ng[2];
    double ud[3], u[3], p = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) { ud[i] = -par.f[i]; u[i] = 0.0; }
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        ud[i] = fma(Ng[a], al[a][i], ud[i]);
        u[i] = fma(Ng[a], yl[a][i], u[i]);
      }
      p = fma(Ng[a], yl[a][3], p);
    }
    double kv[3];
#pragma unroll
    for (int i = 0; i < 3; i++) kv[i] = fma(ks[2][i], u[2], fma(ks[1][i], u[1], ks[0][i] * u[0]));
    const double kU = fma(u[2], kv[2], fma(u[1], kv[1], u[0] * kv[0]));
    const double kSum = kT + kU + kS;
    const double rsK = rsqrt(kSum);
    const double tauM = rsK * irho;
    double rV[3], up[3], ua[3], wv[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      rV[i] = fma(u[2], ux[2][i], fma(u[1], ux[1][i], fma(u[0], ux[0][i], ud[i])));
      up[i] = -tauM * fma(rho, rV[i], px[i]);
    }
    const double tauC = (kSum * rsK) * rho_itrks;
#pragma unroll
    for (int i = 0; i < 3; i++) kv[i] = fma(ks[2][i], up[2], fma(ks[1][i], up[1], ks[0][i] * up[0]));
    double tauB = fma(up[2], kv[2], fma(up[1], kv[1], up[0] * kv[0]));
    if (iszero1(tauB)) tauB = DBL_EPSILON;
    tauB = rho * rsqrt(tauB);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      ua[i] = u[i] + up[i];
      wv[i] = fma(up[2], ux[2][i], fma(up[1], ux[1][i], up[0] * ux[0][i]));
    }
    spa += p - tauC * divU;
    // rM(j,i) without its constant viscous part and without the pressure on the diagonal (added after the loop)
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const double q = rho * up[i], rb = tauB * wv[i];
#pragma unroll
      for (int j = 0; j < 3; j++) sRM[j][i] = fma(rb, up[j], fma(-q, ua[j], sRM[j][i]));
    }
#pragma unroll
    for (int i = 0; i < 3; i++) rV[i] = rV[i] + wv[i];       // ud + ua . grad u
    double uNx[4], upNx[4], X[4], Z[4], Y[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      uNx[a] = fma(u[2], Nx[a][2], fma(u[1], Nx[a][1], u[0] * Nx[a][0]));
      upNx[a] = fma(up[2], Nx[a][2], fma(up[1], Nx[a][1], up[0] * Nx[a][0]));
      const double t = tauM * (uNx[a] + upNx[a]);
      const double y = fma(amd, Ng[a], uNx[a]);
      c2[a] += t;
      r2[a] = fma(tauM, y, r2[a]);
      X[a] = fma(rho, t, Ng[a]);
      Y[a] = rho * y;
      Z[a] = fma(tauB, upNx[a], rho * Ng[a]);
#pragma unroll
      for (int i = 0; i < 3; i++) sNrV[a][i] = fma(Ng[a], rV[i], sNrV[a][i]);
      lR4[a] += fma(Ng[a], divU, -upNx[a]);
    }
    sTC += tauC;
    sTM += tauM;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) A[a][b] = fma(Z[a], upNx[b], fma(X[a], Y[b], A[a][b]));
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {
#pragma unroll
    for (int i = 0; i < 3; i++) { acc.Nx[a][i] = Nx[a][i]; acc.sNrV[a][i] = sNrV[a][i]; }
#pragma unroll
    for (int b = 0; b < 4; b++) acc.A[a][b] = A[a][b];
    acc.c2[a] = c2[a]; acc.r2[a] = r2[a]; acc.lR4[a] = lR4[a];
  }
  const double mu4 = 4.0 * mu;
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double v = fma(mu4, ux[j][i] + ux[i][j], sRM[j][i]);    // 4 mu es(j,i)
      if (i == j) v = v - spa;
      acc.sRM[j][i] = v;
    }
  acc.sTC = sTC; acc.sTM = sTM; acc.w = w; acc.wl = w * T1c; acc.wr = w * rho;
}

}  // namespace svfsi
