/*
 * oracle/ora_elem.c -- TEST INFRASTRUCTURE ONLY (see ora.h header).
 *
 * CPU restatement of the svFSI element loop for 3-D TET4 / VMS / constant
 * viscosity / no mesh motion: TET4 tables, GNN, FLUID3D_M, FLUID3D_C, HEATS3D,
 * LHSA, DOASSEM, CONSTRUCT_FLUID, CONSTRUCT_HEATS.  Statement order follows
 * the Fortran so that rounding is the same when compiled with
 * -ffp-contract=off.  Pinned to the reference's source text (see ora.h): bit-identical
 * to CONSTRUCT_FLUID / CONSTRUCT_HEATS / LHSA executed from /root/reference.
 */
#include "ora.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <time.h>

/* ------------------------------------------------------------------ */
/* S/UTIL.f:56   eps = EPSILON(eps)                                     */
static const double ora_eps = DBL_EPSILON;

/* S/UTIL.f:879-903  ISZERO(ia) with the optional argument absent */
int ora_iszero(double ia) {
  double a = fabs(ia), b = 0.0, nrm;
  if (fabs(b) > fabs(a)) { double tmp = a; a = b; b = tmp; }
  nrm = (a > ora_eps) ? a : ora_eps;
  return ((a - b) / nrm < 10.0 * ora_eps) ? 1 : 0;
}

/* S/NN.f:268-275 (GETGIP, TET4) and S/NN.f:654-671 (GETGNN, TET4).
 * N[g][a], Nxi[a][i]. */
void ora_tet4_tables(double w[4], double N[4][4], double Nxi[4][3]) {
  double s, t, xi[4][3];
  int g, a, i;
  for (g = 0; g < 4; g++) w[g] = 1.0 / 24.0;
  s = (5.0 + 3.0 * sqrt(5.0)) / 20.0;
  t = (5.0 - sqrt(5.0)) / 20.0;
  xi[0][0] = s; xi[0][1] = t; xi[0][2] = t;
  xi[1][0] = t; xi[1][1] = s; xi[1][2] = t;
  xi[2][0] = t; xi[2][1] = t; xi[2][2] = s;
  xi[3][0] = t; xi[3][1] = t; xi[3][2] = t;
  for (g = 0; g < 4; g++) {
    N[g][0] = xi[g][0];
    N[g][1] = xi[g][1];
    N[g][2] = xi[g][2];
    N[g][3] = 1.0 - xi[g][0] - xi[g][1] - xi[g][2];
  }
  for (a = 0; a < 4; a++)
    for (i = 0; i < 3; i++) Nxi[a][i] = 0.0;
  Nxi[0][0] = 1.0; Nxi[1][1] = 1.0; Nxi[2][2] = 1.0;
  Nxi[3][0] = -1.0; Nxi[3][1] = -1.0; Nxi[3][2] = -1.0;
}

/* S/NN.f:1515-1561  GNN, insd = 3, eNoN = 4.
 * xl[a][i] = x(i,a); Nxi[a][i] = Nxi(i,a); Nx[a][i] = Nx(i,a); ks[i][j]. */
void ora_gnn3(const double Nxi[4][3], const double xl[4][3], double Nx[4][3],
              double *Jac, double ks[3][3]) {
  double xXi[3][3], xiX[3][3]; /* [row][col], 0-based of xXi(r,c) */
  int a, r, c;
  for (r = 0; r < 3; r++)
    for (c = 0; c < 3; c++) xXi[r][c] = 0.0;
  for (a = 0; a < 4; a++) {
    for (r = 0; r < 3; r++) xXi[r][0] = xXi[r][0] + xl[a][r] * Nxi[a][0];
    for (r = 0; r < 3; r++) xXi[r][1] = xXi[r][1] + xl[a][r] * Nxi[a][1];
    for (r = 0; r < 3; r++) xXi[r][2] = xXi[r][2] + xl[a][r] * Nxi[a][2];
  }
#define X(i, j) xXi[(i)-1][(j)-1]
#define XI(i, j) xiX[(i)-1][(j)-1]
#define KS(i, j) ks[(i)-1][(j)-1]
  *Jac = X(1,1)*X(2,2)*X(3,3) + X(1,2)*X(2,3)*X(3,1) + X(1,3)*X(2,1)*X(3,2)
       - X(1,1)*X(2,3)*X(3,2) - X(1,2)*X(2,1)*X(3,3) - X(1,3)*X(2,2)*X(3,1);

  XI(1,1) = (X(2,2)*X(3,3) - X(2,3)*X(3,2)) / *Jac;
  XI(1,2) = (X(3,2)*X(1,3) - X(3,3)*X(1,2)) / *Jac;
  XI(1,3) = (X(1,2)*X(2,3) - X(1,3)*X(2,2)) / *Jac;
  XI(2,1) = (X(2,3)*X(3,1) - X(2,1)*X(3,3)) / *Jac;
  XI(2,2) = (X(3,3)*X(1,1) - X(3,1)*X(1,3)) / *Jac;
  XI(2,3) = (X(1,3)*X(2,1) - X(1,1)*X(2,3)) / *Jac;
  XI(3,1) = (X(2,1)*X(3,2) - X(2,2)*X(3,1)) / *Jac;
  XI(3,2) = (X(3,1)*X(1,2) - X(3,2)*X(1,1)) / *Jac;
  XI(3,3) = (X(1,1)*X(2,2) - X(1,2)*X(2,1)) / *Jac;

  KS(1,1) = XI(1,1)*XI(1,1) + XI(2,1)*XI(2,1) + XI(3,1)*XI(3,1);
  KS(1,2) = XI(1,2)*XI(1,1) + XI(2,2)*XI(2,1) + XI(3,2)*XI(3,1);
  KS(1,3) = XI(1,3)*XI(1,1) + XI(2,3)*XI(2,1) + XI(3,3)*XI(3,1);
  KS(2,2) = XI(1,2)*XI(1,2) + XI(2,2)*XI(2,2) + XI(3,2)*XI(3,2);
  KS(2,3) = XI(1,2)*XI(1,3) + XI(2,2)*XI(2,3) + XI(3,2)*XI(3,3);
  KS(3,3) = XI(1,3)*XI(1,3) + XI(2,3)*XI(2,3) + XI(3,3)*XI(3,3);
  KS(2,1) = KS(1,2);
  KS(3,1) = KS(1,3);
  KS(3,2) = KS(2,3);

  for (a = 0; a < 4; a++) {
    for (c = 0; c < 3; c++) Nx[a][c] = 0.0;
    Nx[a][0] = Nx[a][0] + Nxi[a][0]*XI(1,1) + Nxi[a][1]*XI(2,1) + Nxi[a][2]*XI(3,1);
    Nx[a][1] = Nx[a][1] + Nxi[a][0]*XI(1,2) + Nxi[a][1]*XI(2,2) + Nxi[a][2]*XI(3,2);
    Nx[a][2] = Nx[a][2] + Nxi[a][0]*XI(1,3) + Nxi[a][1]*XI(2,3) + Nxi[a][2]*XI(3,3);
  }
#undef X
#undef XI
#undef KS
}

/* 1-based accessors over the C arrays used below (Fortran (i,a) -> [a-1][i-1]) */
#define NWX(i, a) Nwx[(a)-1][(i)-1]
#define NQX(i, a) Nqx[(a)-1][(i)-1]
#define NWXX(i, a) Nwxx[(a)-1][(i)-1]
#define NW(a) Nw[(a)-1]
#define NQ(a) Nq[(a)-1]
#define AL(i, a) al[(a)-1][(i)-1]
#define YL(i, a) yl[(a)-1][(i)-1]
#define BFL(i, a) bfl[(a)-1][(i)-1]
#define LR(i, a) lR[(a)-1][(i)-1]
#define LK(k, a, b) lK[(b)-1][(a)-1][(k)-1]
#define KXI(i, j) Kxi[(i)-1][(j)-1]
#define UX(i, j) ux[(i)-1][(j)-1]
#define UXX(i, j, k) uxx[(i)-1][(j)-1][(k)-1]
#define ES(i, j) es[(i)-1][(j)-1]
#define ES_X(i, j, k) es_x[(i)-1][(j)-1][(k)-1]
#define ESNX(i, a) esNx[(a)-1][(i)-1]
#define UPDU(i, j, a) updu[(a)-1][(j)-1][(i)-1]
#define RM(i, j) rM[(i)-1][(j)-1]

/* Shared front part of FLUID3D_M (S/FLUID.f:229-365) and FLUID3D_C
 * (S/FLUID.f:847-982): interpolation, strain rate, viscosity. */
typedef struct {
  double ud[3], u[3], ux[3][3], uxx[3][3][3], divU, d2u2[3], p, px[3];
  double es[3][3], es_x[3][3][3], esNx[4][3], mu_x[3], gam, mu, mu_s, mu_g;
} ora_gp_t;

static void ora_fluid_front(const ora_fluid_par_t *par, const double Nw[4],
                            const double Nq[4], const double Nwx[4][3],
                            const double Nqx[4][3], const double Nwxx[4][6],
                            const double al[4][4], const double yl[4][4],
                            const double bfl[4][3], int with_p, ora_gp_t *q) {
  double (*ux)[3] = q->ux;
  double (*uxx)[3][3] = q->uxx;
  double (*es)[3] = q->es;
  double (*es_x)[3][3] = q->es_x;
  double (*esNx)[3] = q->esNx;
  double *ud = q->ud, *u = q->u, *px = q->px, *d2u2 = q->d2u2, *mu_x = q->mu_x;
  int a, k;

  ud[0] = -par->f[0]; ud[1] = -par->f[1]; ud[2] = -par->f[2];
  u[0] = u[1] = u[2] = 0.0;
  memset(q->ux, 0, sizeof(q->ux));
  memset(q->uxx, 0, sizeof(q->uxx));
  for (a = 1; a <= 4; a++) {
    ud[0] = ud[0] + NW(a)*(AL(1,a) - BFL(1,a));
    ud[1] = ud[1] + NW(a)*(AL(2,a) - BFL(2,a));
    ud[2] = ud[2] + NW(a)*(AL(3,a) - BFL(3,a));

    u[0] = u[0] + NW(a)*YL(1,a);
    u[1] = u[1] + NW(a)*YL(2,a);
    u[2] = u[2] + NW(a)*YL(3,a);

    UX(1,1) = UX(1,1) + NWX(1,a)*YL(1,a);
    UX(2,1) = UX(2,1) + NWX(2,a)*YL(1,a);
    UX(3,1) = UX(3,1) + NWX(3,a)*YL(1,a);
    UX(1,2) = UX(1,2) + NWX(1,a)*YL(2,a);
    UX(2,2) = UX(2,2) + NWX(2,a)*YL(2,a);
    UX(3,2) = UX(3,2) + NWX(3,a)*YL(2,a);
    UX(1,3) = UX(1,3) + NWX(1,a)*YL(3,a);
    UX(2,3) = UX(2,3) + NWX(2,a)*YL(3,a);
    UX(3,3) = UX(3,3) + NWX(3,a)*YL(3,a);

    for (k = 1; k <= 3; k++) {
      UXX(1,k,1) = UXX(1,k,1) + NWXX(1,a)*YL(k,a);
      UXX(2,k,2) = UXX(2,k,2) + NWXX(2,a)*YL(k,a);
      UXX(3,k,3) = UXX(3,k,3) + NWXX(3,a)*YL(k,a);
      UXX(2,k,1) = UXX(2,k,1) + NWXX(4,a)*YL(k,a);
      UXX(3,k,2) = UXX(3,k,2) + NWXX(5,a)*YL(k,a);
      UXX(1,k,3) = UXX(1,k,3) + NWXX(6,a)*YL(k,a);
    }
  }
  q->divU = UX(1,1) + UX(2,2) + UX(3,3);

  for (k = 1; k <= 3; k++) {
    UXX(1,k,2) = UXX(2,k,1);
    UXX(2,k,3) = UXX(3,k,2);
    UXX(3,k,1) = UXX(1,k,3);
  }
  d2u2[0] = UXX(1,1,1) + UXX(2,1,2) + UXX(3,1,3);
  d2u2[1] = UXX(1,2,1) + UXX(2,2,2) + UXX(3,2,3);
  d2u2[2] = UXX(1,3,1) + UXX(2,3,2) + UXX(3,3,3);

  /* pressure and its gradient; FLUID3D_C does not form p (S/FLUID.f:925-930) */
  q->p = 0.0;
  px[0] = px[1] = px[2] = 0.0;
  for (a = 1; a <= 4; a++) {
    if (with_p) q->p = q->p + NQ(a)*YL(4,a);
    px[0] = px[0] + NQX(1,a)*YL(4,a);
    px[1] = px[1] + NQX(2,a)*YL(4,a);
    px[2] = px[2] + NQX(3,a)*YL(4,a);
  }
  /* mvMsh = .FALSE. (fluid only): S/FLUID.f:303-309 skipped */

  ES(1,1) = UX(1,1) + UX(1,1);
  ES(2,2) = UX(2,2) + UX(2,2);
  ES(3,3) = UX(3,3) + UX(3,3);
  ES(2,1) = UX(2,1) + UX(1,2);
  ES(3,2) = UX(3,2) + UX(2,3);
  ES(1,3) = UX(1,3) + UX(3,1);
  ES(1,2) = ES(2,1);
  ES(2,3) = ES(3,2);
  ES(3,1) = ES(1,3);

  for (a = 1; a <= 4; a++) {
    ESNX(1,a) = ES(1,1)*NWX(1,a) + ES(2,1)*NWX(2,a) + ES(3,1)*NWX(3,a);
    ESNX(2,a) = ES(1,2)*NWX(1,a) + ES(2,2)*NWX(2,a) + ES(3,2)*NWX(3,a);
    ESNX(3,a) = ES(1,3)*NWX(1,a) + ES(2,3)*NWX(2,a) + ES(3,3)*NWX(3,a);
  }

  for (k = 1; k <= 3; k++) {
    ES_X(1,1,k) = UXX(1,1,k) + UXX(1,1,k);
    ES_X(2,2,k) = UXX(2,2,k) + UXX(2,2,k);
    ES_X(3,3,k) = UXX(3,3,k) + UXX(3,3,k);
    ES_X(2,1,k) = UXX(2,1,k) + UXX(1,2,k);
    ES_X(3,2,k) = UXX(3,2,k) + UXX(2,3,k);
    ES_X(1,3,k) = UXX(1,3,k) + UXX(3,1,k);
    ES_X(1,2,k) = ES_X(2,1,k);
    ES_X(2,3,k) = ES_X(3,2,k);
    ES_X(3,1,k) = ES_X(1,3,k);
  }
  for (k = 1; k <= 3; k++) {
    mu_x[k-1] = (ES_X(1,1,k)*ES(1,1) + ES_X(2,2,k)*ES(2,2)
              +  ES_X(3,3,k)*ES(3,3))*0.5
              +  ES_X(2,1,k)*ES(2,1) + ES_X(3,2,k)*ES(3,2)
              +  ES_X(1,3,k)*ES(1,3);
  }

  q->gam = ES(1,1)*ES(1,1) + ES(2,1)*ES(2,1) + ES(3,1)*ES(3,1)
         + ES(1,2)*ES(1,2) + ES(2,2)*ES(2,2) + ES(3,2)*ES(3,2)
         + ES(1,3)*ES(1,3) + ES(2,3)*ES(2,3) + ES(3,3)*ES(3,3);
  q->gam = sqrt(0.5*q->gam);

  /* GETVISCOSITY, viscType_Const: S/FLUID.f:1700-1703 */
  q->mu = par->mu;
  q->mu_s = q->mu;
  q->mu_g = 0.0;
  if (ora_iszero(q->gam)) q->mu_g = 0.0;
  else q->mu_g = q->mu_g / q->gam;
  mu_x[0] = q->mu_g*mu_x[0]; mu_x[1] = q->mu_g*mu_x[1]; mu_x[2] = q->mu_g*mu_x[2];
}

/* S/FLUID.f:367-402 (also :991-1018): tauM, rV, rS, up */
static void ora_fluid_tau(const ora_fluid_par_t *par, const double Kxi[3][3],
                          const ora_gp_t *q, double *tauM, double rV[3],
                          double up[3]) {
  const double ctM = 1.0, ctC = 36.0;
  const double *u = q->u, *ud = q->ud, *px = q->px, *mu_x = q->mu_x, *d2u2 = q->d2u2;
  const double (*ux)[3] = q->ux;
  const double (*es)[3] = q->es;
  double rho = par->rho, mu = q->mu, kT, kU, kS, rS[3], tq;

  tq = ctM / par->dt;
  kT = 4.0*(tq*tq);                                  /* 4*(ctM/dt)**2 */

  kU = u[0]*u[0]*KXI(1,1) + u[1]*u[0]*KXI(2,1) + u[2]*u[0]*KXI(3,1)
     + u[0]*u[1]*KXI(1,2) + u[1]*u[1]*KXI(2,2) + u[2]*u[1]*KXI(3,2)
     + u[0]*u[2]*KXI(1,3) + u[1]*u[2]*KXI(2,3) + u[2]*u[2]*KXI(3,3);

  kS = KXI(1,1)*KXI(1,1) + KXI(2,1)*KXI(2,1) + KXI(3,1)*KXI(3,1)
     + KXI(1,2)*KXI(1,2) + KXI(2,2)*KXI(2,2) + KXI(3,2)*KXI(3,2)
     + KXI(1,3)*KXI(1,3) + KXI(2,3)*KXI(2,3) + KXI(3,3)*KXI(3,3);
  tq = mu / rho;
  kS = ctC * kS * (tq*tq);                           /* (mu/rho)**2 */

  *tauM = 1.0 / (rho * sqrt(kT + kU + kS));

  rV[0] = ud[0] + u[0]*UX(1,1) + u[1]*UX(2,1) + u[2]*UX(3,1);
  rV[1] = ud[1] + u[0]*UX(1,2) + u[1]*UX(2,2) + u[2]*UX(3,2);
  rV[2] = ud[2] + u[0]*UX(1,3) + u[1]*UX(2,3) + u[2]*UX(3,3);

  rS[0] = mu_x[0]*ES(1,1) + mu_x[1]*ES(2,1) + mu_x[2]*ES(3,1) + mu*d2u2[0];
  rS[1] = mu_x[0]*ES(1,2) + mu_x[1]*ES(2,2) + mu_x[2]*ES(3,2) + mu*d2u2[1];
  rS[2] = mu_x[0]*ES(1,3) + mu_x[1]*ES(2,3) + mu_x[2]*ES(3,3) + mu*d2u2[2];

  up[0] = -*tauM*(rho*rV[0] + px[0] - rS[0]);
  up[1] = -*tauM*(rho*rV[1] + px[1] - rS[1]);
  up[2] = -*tauM*(rho*rV[2] + px[2] - rS[2]);
}

/* S/FLUID.f:465-478 (also :1020-1036): T1 and updu for one node a.
 * updu[a][j][i] holds updu(i,j,a). */
static void ora_fluid_updu(const ora_gp_t *q, double rho, const double Nwx[4][3],
                           const double Nwxx[4][6], int a, double uNx_a,
                           double updu[4][3][3]) {
  const double *mu_x = q->mu_x, *d2u2 = q->d2u2;
  const double (*esNx)[3] = q->esNx;
  double mu = q->mu, mu_g = q->mu_g, T1;
  T1 = -rho*uNx_a + mu*(NWXX(1,a) + NWXX(2,a) + NWXX(3,a))
     + mu_x[0]*NWX(1,a) + mu_x[1]*NWX(2,a) + mu_x[2]*NWX(3,a);

  UPDU(1,1,a) = mu_x[0]*NWX(1,a) + d2u2[0]*mu_g*ESNX(1,a) + T1;
  UPDU(2,1,a) = mu_x[1]*NWX(1,a) + d2u2[0]*mu_g*ESNX(2,a);
  UPDU(3,1,a) = mu_x[2]*NWX(1,a) + d2u2[0]*mu_g*ESNX(3,a);

  UPDU(1,2,a) = mu_x[0]*NWX(2,a) + d2u2[1]*mu_g*ESNX(1,a);
  UPDU(2,2,a) = mu_x[1]*NWX(2,a) + d2u2[1]*mu_g*ESNX(2,a) + T1;
  UPDU(3,2,a) = mu_x[2]*NWX(2,a) + d2u2[1]*mu_g*ESNX(3,a);

  UPDU(1,3,a) = mu_x[0]*NWX(3,a) + d2u2[2]*mu_g*ESNX(1,a);
  UPDU(2,3,a) = mu_x[1]*NWX(3,a) + d2u2[2]*mu_g*ESNX(2,a);
  UPDU(3,3,a) = mu_x[2]*NWX(3,a) + d2u2[2]*mu_g*ESNX(3,a) + T1;
}

/* S/FLUID.f:192-560  FLUID3D_M with vmsFlag = .TRUE., eNoNw = eNoNq = 4 */
static void ora_fluid3d_m(const ora_fluid_par_t *par, double w,
                          const double Kxi[3][3], const double Nw[4],
                          const double Nq[4], const double Nwx[4][3],
                          const double Nqx[4][3], const double Nwxx[4][6],
                          const double al[4][4], const double yl[4][4],
                          const double bfl[4][3], double lR[4][4],
                          double lK[4][4][16]) {
  ora_gp_t q;
  double rho = par->rho, T1, T2, amd, wl, wr, tauM, tauC, tauB, pa, mu, mu_g;
  double up[3], ua[3], rV[3], rM[3][3], updu[4][3][3];
  double uNx[4], upNx[4], uaNx[4], NxNx;
  const double *u, *ud;
  const double (*ux)[3];
  const double (*es)[3];
  const double (*esNx)[3];
  int a, b;

  T1  = par->af * par->gam * par->dt;
  amd = par->am / T1;
  wl  = w*T1;
  wr  = w*rho;

  ora_fluid_front(par, Nw, Nq, Nwx, Nqx, Nwxx, al, yl, bfl, 1, &q);
  u = q.u; ud = q.ud; ux = q.ux; es = q.es; esNx = q.esNx;
  mu = q.mu; mu_g = q.mu_g;
  ora_fluid_tau(par, Kxi, &q, &tauM, rV, up);

  /* vmsFlag branch, S/FLUID.f:404-418 */
  tauC = 1.0 / (tauM * (KXI(1,1) + KXI(2,2) + KXI(3,3)));
  tauB = up[0]*up[0]*KXI(1,1) + up[1]*up[0]*KXI(2,1)
       + up[2]*up[0]*KXI(3,1) + up[0]*up[1]*KXI(1,2)
       + up[1]*up[1]*KXI(2,2) + up[2]*up[1]*KXI(3,2)
       + up[0]*up[2]*KXI(1,3) + up[1]*up[2]*KXI(2,3)
       + up[2]*up[2]*KXI(3,3);
  if (ora_iszero(tauB)) tauB = ora_eps;
  tauB = rho / sqrt(tauB);

  ua[0] = u[0] + up[0];
  ua[1] = u[1] + up[1];
  ua[2] = u[2] + up[2];
  pa    = q.p - tauC*q.divU;

  rV[0] = tauB*(up[0]*UX(1,1) + up[1]*UX(2,1) + up[2]*UX(3,1));
  rV[1] = tauB*(up[0]*UX(1,2) + up[1]*UX(2,2) + up[2]*UX(3,2));
  rV[2] = tauB*(up[0]*UX(1,3) + up[1]*UX(2,3) + up[2]*UX(3,3));

  RM(1,1) = mu*ES(1,1) - rho*up[0]*ua[0] + rV[0]*up[0] - pa;
  RM(2,1) = mu*ES(2,1) - rho*up[0]*ua[1] + rV[0]*up[1];
  RM(3,1) = mu*ES(3,1) - rho*up[0]*ua[2] + rV[0]*up[2];

  RM(1,2) = mu*ES(1,2) - rho*up[1]*ua[0] + rV[1]*up[0];
  RM(2,2) = mu*ES(2,2) - rho*up[1]*ua[1] + rV[1]*up[1] - pa;
  RM(3,2) = mu*ES(3,2) - rho*up[1]*ua[2] + rV[1]*up[2];

  RM(1,3) = mu*ES(1,3) - rho*up[2]*ua[0] + rV[2]*up[0];
  RM(2,3) = mu*ES(2,3) - rho*up[2]*ua[1] + rV[2]*up[1];
  RM(3,3) = mu*ES(3,3) - rho*up[2]*ua[2] + rV[2]*up[2] - pa;

  rV[0] = ud[0] + ua[0]*UX(1,1) + ua[1]*UX(2,1) + ua[2]*UX(3,1);
  rV[1] = ud[1] + ua[0]*UX(1,2) + ua[1]*UX(2,2) + ua[2]*UX(3,2);
  rV[2] = ud[2] + ua[0]*UX(1,3) + ua[1]*UX(2,3) + ua[2]*UX(3,3);

  /* local residue, S/FLUID.f:446-479 */
  for (a = 1; a <= 4; a++) {
    LR(1,a) = LR(1,a) + wr*NW(a)*rV[0] + w*(NWX(1,a)*RM(1,1)
            + NWX(2,a)*RM(2,1) + NWX(3,a)*RM(3,1));
    LR(2,a) = LR(2,a) + wr*NW(a)*rV[1] + w*(NWX(1,a)*RM(1,2)
            + NWX(2,a)*RM(2,2) + NWX(3,a)*RM(3,2));
    LR(3,a) = LR(3,a) + wr*NW(a)*rV[2] + w*(NWX(1,a)*RM(1,3)
            + NWX(2,a)*RM(2,3) + NWX(3,a)*RM(3,3));

    uNx[a-1]  = u[0]*NWX(1,a)  + u[1]*NWX(2,a)  + u[2]*NWX(3,a);
    upNx[a-1] = up[0]*NWX(1,a) + up[1]*NWX(2,a) + up[2]*NWX(3,a);
    uaNx[a-1] = uNx[a-1] + upNx[a-1];

    ora_fluid_updu(&q, rho, Nwx, Nwxx, a, uNx[a-1], updu);
  }

  /* tangent, S/FLUID.f:482-545 */
  for (b = 1; b <= 4; b++) {
    for (a = 1; a <= 4; a++) {
      RM(1,1) = NWX(1,a)*NWX(1,b);
      RM(2,1) = NWX(2,a)*NWX(1,b);
      RM(3,1) = NWX(3,a)*NWX(1,b);
      RM(1,2) = NWX(1,a)*NWX(2,b);
      RM(2,2) = NWX(2,a)*NWX(2,b);
      RM(3,2) = NWX(3,a)*NWX(2,b);
      RM(1,3) = NWX(1,a)*NWX(3,b);
      RM(2,3) = NWX(2,a)*NWX(3,b);
      RM(3,3) = NWX(3,a)*NWX(3,b);

      NxNx = NWX(1,a)*NWX(1,b) + NWX(2,a)*NWX(2,b) + NWX(3,a)*NWX(3,b);

      T1 = mu*NxNx + rho*amd*NW(b)*(NW(a) + rho*tauM*uaNx[a-1])
         + rho*NW(a)*(uNx[b-1]+upNx[b-1]) + tauB*upNx[a-1]*upNx[b-1];

      T2 = (mu + tauC)*RM(1,1) + ESNX(1,a)*mu_g*ESNX(1,b)
         - rho*tauM*uaNx[a-1]*UPDU(1,1,b);
      LK(1,a,b)  = LK(1,a,b)  + wl*(T2 + T1);

      T2 = mu*RM(2,1) + tauC*RM(1,2) + ESNX(1,a)*mu_g*ESNX(2,b)
         - rho*tauM*uaNx[a-1]*UPDU(2,1,b);
      LK(2,a,b)  = LK(2,a,b)  + wl*(T2);

      T2 = mu*RM(3,1) + tauC*RM(1,3) + ESNX(1,a)*mu_g*ESNX(3,b)
         - rho*tauM*uaNx[a-1]*UPDU(3,1,b);
      LK(3,a,b)  = LK(3,a,b)  + wl*(T2);

      T2 = mu*RM(1,2) + tauC*RM(2,1) + ESNX(2,a)*mu_g*ESNX(1,b)
         - rho*tauM*uaNx[a-1]*UPDU(1,2,b);
      LK(5,a,b)  = LK(5,a,b)  + wl*(T2);

      T2 = (mu + tauC)*RM(2,2) + ESNX(2,a)*mu_g*ESNX(2,b)
         - rho*tauM*uaNx[a-1]*UPDU(2,2,b);
      LK(6,a,b)  = LK(6,a,b)  + wl*(T2 + T1);

      T2 = mu*RM(3,2) + tauC*RM(2,3) + ESNX(2,a)*mu_g*ESNX(3,b)
         - rho*tauM*uaNx[a-1]*UPDU(3,2,b);
      LK(7,a,b)  = LK(7,a,b)  + wl*(T2);

      T2 = mu*RM(1,3) + tauC*RM(3,1) + ESNX(3,a)*mu_g*ESNX(1,b)
         - rho*tauM*uaNx[a-1]*UPDU(1,3,b);
      LK(9,a,b)  = LK(9,a,b)  + wl*(T2);

      T2 = mu*RM(2,3) + tauC*RM(3,2) + ESNX(3,a)*mu_g*ESNX(2,b)
         - rho*tauM*uaNx[a-1]*UPDU(2,3,b);
      LK(10,a,b) = LK(10,a,b) + wl*(T2);

      T2 = (mu + tauC)*RM(3,3) + ESNX(3,a)*mu_g*ESNX(3,b)
         - rho*tauM*uaNx[a-1]*UPDU(3,3,b);
      LK(11,a,b) = LK(11,a,b) + wl*(T2 + T1);
    }
  }

  /* S/FLUID.f:547-557 */
  for (b = 1; b <= 4; b++) {
    for (a = 1; a <= 4; a++) {
      T1 = rho*tauM*uaNx[a-1];
      LK(4,a,b)  = LK(4,a,b)  - wl*(NWX(1,a)*NQ(b) - NQX(1,b)*T1);
      LK(8,a,b)  = LK(8,a,b)  - wl*(NWX(2,a)*NQ(b) - NQX(2,b)*T1);
      LK(12,a,b) = LK(12,a,b) - wl*(NWX(3,a)*NQ(b) - NQX(3,b)*T1);
    }
  }
}

/* S/FLUID.f:813-1084  FLUID3D_C with vmsFlag = .TRUE. */
static void ora_fluid3d_c(const ora_fluid_par_t *par, double w,
                          const double Kxi[3][3], const double Nw[4],
                          const double Nq[4], const double Nwx[4][3],
                          const double Nqx[4][3], const double Nwxx[4][6],
                          const double al[4][4], const double yl[4][4],
                          const double bfl[4][3], double lR[4][4],
                          double lK[4][4][16]) {
  ora_gp_t q;
  double rho = par->rho, T1, T2, amd, wl, tauM, uNx, upNx, NxNx;
  double up[3], rV[3], updu[4][3][3];
  int a, b;

  T1  = par->af * par->gam * par->dt;
  amd = par->am / T1;
  wl  = w*T1;

  ora_fluid_front(par, Nw, Nq, Nwx, Nqx, Nwxx, al, yl, bfl, 0, &q);
  ora_fluid_tau(par, Kxi, &q, &tauM, rV, up);
  for (a = 1; a <= 4; a++) {
    uNx = q.u[0]*NWX(1,a) + q.u[1]*NWX(2,a) + q.u[2]*NWX(3,a);
    ora_fluid_updu(&q, rho, Nwx, Nwxx, a, uNx, updu);
  }

  /* S/FLUID.f:1045-1049 */
  for (a = 1; a <= 4; a++) {
    upNx    = up[0]*NQX(1,a) + up[1]*NQX(2,a) + up[2]*NQX(3,a);
    LR(4,a) = LR(4,a) + w*(NQ(a)*q.divU - upNx);
  }

  /* S/FLUID.f:1052-1070 */
  for (b = 1; b <= 4; b++) {
    T1 = rho*amd*NW(b);
    for (a = 1; a <= 4; a++) {
      T2 = NQX(1,a)*(UPDU(1,1,b) - T1) + NQX(2,a)*UPDU(1,2,b)
         + NQX(3,a)*UPDU(1,3,b);
      LK(13,a,b) = LK(13,a,b) + wl*(NQ(a)*NWX(1,b) - tauM*T2);

      T2 = NQX(1,a)*UPDU(2,1,b) + NQX(2,a)*(UPDU(2,2,b) - T1)
         + NQX(3,a)*UPDU(2,3,b);
      LK(14,a,b) = LK(14,a,b) + wl*(NQ(a)*NWX(2,b) - tauM*T2);

      T2 = NQX(1,a)*UPDU(3,1,b) + NQX(2,a)*UPDU(3,2,b)
         + NQX(3,a)*(UPDU(3,3,b) - T1);
      LK(15,a,b) = LK(15,a,b) + wl*(NQ(a)*NWX(3,b) - tauM*T2);
    }
  }

  /* S/FLUID.f:1072-1081 */
  for (b = 1; b <= 4; b++) {
    for (a = 1; a <= 4; a++) {
      NxNx = NQX(1,a)*NQX(1,b) + NQX(2,a)*NQX(2,b) + NQX(3,a)*NQX(3,b);
      LK(16,a,b) = LK(16,a,b) + wl*tauM*NxNx;
    }
  }
}

/* One element of CONSTRUCT_FLUID, S/FLUID.f:84-168: lR, lK zeroed, shape data
 * once (lShpF), Gauss loop 1 (momentum) then Gauss loop 2 (continuity).
 * lR[a][i] = lR(i,a); lK[b][a][k] = lK(k,a,b).  *jac_flag is set when
 * ISZERO(Jac) (S/FLUID.f:115). */
void ora_fluid_element(const ora_fluid_par_t *par, const double xl[4][3],
                       const double al[4][4], const double yl[4][4],
                       const double bfl[4][3], double lR[4][4],
                       double lK[4][4][16], int *jac_flag) {
  double w4[4], N[4][4], Nxi[4][3], Nwx[4][3], Nqx[4][3], Nwxx[4][6];
  double Jac, ksix[3][3], w;
  int g;

  ora_tet4_tables(w4, N, Nxi);
  memset(lR, 0, sizeof(double)*16);
  memset(lK, 0, sizeof(double)*256);
  /* Nwxx: GNNxx solves K X = 0 for TET4 (fs%Nxx == 0, S/FS.f:193-195) => 0 */
  memset(Nwxx, 0, sizeof(Nwxx));

  for (g = 0; g < 4; g++) {
    if (g == 0) {
      ora_gnn3(Nxi, xl, Nqx, &Jac, ksix);
      if (ora_iszero(Jac) && jac_flag) *jac_flag = 1;
      ora_gnn3(Nxi, xl, Nwx, &Jac, ksix);
      if (ora_iszero(Jac) && jac_flag) *jac_flag = 1;
    }
    w = w4[g] * Jac;
    ora_fluid3d_m(par, w, ksix, N[g], N[g], Nwx, Nqx, Nwxx, al, yl, bfl, lR, lK);
  }
  for (g = 0; g < 4; g++) {
    if (g == 0) {
      ora_gnn3(Nxi, xl, Nwx, &Jac, ksix);
      ora_gnn3(Nxi, xl, Nqx, &Jac, ksix);
    }
    w = w4[g] * Jac;
    ora_fluid3d_c(par, w, ksix, N[g], N[g], Nwx, Nqx, Nwxx, al, yl, bfl, lR, lK);
  }
}

/* S/HEATS.f:116-159 HEATS3D inside the element of CONSTRUCT_HEATS :60-92.
 * lK[a][b] = lK(1,a,b) (dof = 1). */
void ora_heat_element(const ora_heat_par_t *par, const double xl[4][3],
                      const double al[4], const double yl[4], double lR[4],
                      double lK[4][4], int *jac_flag) {
  double w4[4], N[4][4], Nxi[4][3], Nx[4][3], Jac = 0.0, ksix[3][3];
  double nu = par->nu, s = par->s, rho = par->rho, T1, amd, wl, w, Td, Tx[3];
  int g, a, b;

  ora_tet4_tables(w4, N, Nxi);
  for (a = 0; a < 4; a++) {
    lR[a] = 0.0;
    for (b = 0; b < 4; b++) lK[a][b] = 0.0;
  }
  for (g = 0; g < 4; g++) {
    if (g == 0) {
      ora_gnn3(Nxi, xl, Nx, &Jac, ksix);
      if (ora_iszero(Jac) && jac_flag) *jac_flag = 1;
    }
    w = w4[g] * Jac;

    T1  = par->af*par->gam*par->dt;
    amd = par->am * rho/T1;
    wl  = w*T1;

    Td = -s;
    Tx[0] = Tx[1] = Tx[2] = 0.0;
    for (a = 0; a < 4; a++) {
      Td = Td + N[g][a]*al[a];
      Tx[0] = Tx[0] + Nx[a][0]*yl[a];
      Tx[1] = Tx[1] + Nx[a][1]*yl[a];
      Tx[2] = Tx[2] + Nx[a][2]*yl[a];
    }
    Td = Td * rho;

    for (a = 0; a < 4; a++) {
      lR[a] = lR[a] + w*(N[g][a]*Td
            + (Nx[a][0]*Tx[0] + Nx[a][1]*Tx[1] + Nx[a][2]*Tx[2])*nu);
      for (b = 0; b < 4; b++) {
        lK[a][b] = lK[a][b] + wl*(N[g][a]*N[g][b]*amd
                 + nu*(Nx[a][0]*Nx[b][0] + Nx[a][1]*Nx[b][1] + Nx[a][2]*Nx[b][2]));
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/* S/LHSA.f:57-72,174-199,205-241: node-graph CSR with ascending unique column
 * ids (diagonal included), 1-based rowPtr[tnNo+1], colPtr[nnz].  The reference
 * keeps a dense (mnnzeic x tnNo) table with sorted insertion (ADDCOL) and
 * grows it (RESIZ); per-row growable arrays give the identical result.
 * Returns 0, or the 1-based id of an isolated node (S/LHSA.f:176-178). */
int ora_lhsa(int tnNo, int nEl, const int *IEN, int **rowPtr_out,
             int **colPtr_out, int *nnz_out) {
  int **uInd = (int **)calloc((size_t)tnNo, sizeof(int *));
  int *cnt = (int *)calloc((size_t)tnNo, sizeof(int));
  int *cap = (int *)calloc((size_t)tnNo, sizeof(int));
  int e, a, b, i, j, rowN, colN, nnz = 0, *rowPtr, *colPtr, isolated = 0;

  for (e = 0; e < nEl; e++) {
    for (a = 0; a < 4; a++) {
      rowN = IEN[4*e + a] - 1;
      for (b = 0; b < 4; b++) {
        int *row, n;
        colN = IEN[4*e + b];
        /* ADDCOL(rowN, colN), S/LHSA.f:205-241 */
        row = uInd[rowN]; n = cnt[rowN];
        for (i = 0; i < n; i++) if (colN <= row[i]) break;
        if (i < n && row[i] == colN) continue;
        if (n == cap[rowN]) {
          cap[rowN] = cap[rowN] ? cap[rowN] + (cap[rowN]/5 > 5 ? cap[rowN]/5 : 5) : 40;
          row = uInd[rowN] = (int *)realloc(row, sizeof(int)*(size_t)cap[rowN]);
        }
        for (j = n; j > i; j--) row[j] = row[j-1];
        row[i] = colN;
        cnt[rowN] = n + 1;
      }
    }
  }
  for (a = 0; a < tnNo; a++) {
    if (cnt[a] == 0 && !isolated) isolated = a + 1;
    nnz += cnt[a];
  }
  rowPtr = (int *)malloc(sizeof(int)*(size_t)(tnNo + 1));
  colPtr = (int *)malloc(sizeof(int)*(size_t)(nnz > 0 ? nnz : 1));
  j = 1;
  rowPtr[0] = 1;
  for (a = 0; a < tnNo; a++) {
    for (i = 0; i < cnt[a]; i++) colPtr[j - 1 + i] = uInd[a][i];
    j += cnt[a];
    rowPtr[a + 1] = j;
    free(uInd[a]);
  }
  free(uInd); free(cnt); free(cap);
  *rowPtr_out = rowPtr; *colPtr_out = colPtr; *nnz_out = nnz;
  return isolated;
}

void ora_free(void *p) { free(p); }

/* S/LHSA.f:266-298 DOASSEM: binary search per (a,b), Val += lK, R += lR.
 * lK is lK(dof*dof, d, d), lR is lR(dof, d) in Fortran order. */
void ora_doassem(int dof, int d, const int *eqN, const double *lK,
                 const double *lR, const int *rowPtr, const int *colPtr,
                 double *R, double *Val) {
  int a, b, k, ptr, rowN, colN, left, right, dd = dof*dof;
  for (a = 1; a <= d; a++) {
    rowN = eqN[a-1];
    if (rowN == 0) continue;
    for (k = 0; k < dof; k++) R[(size_t)(rowN-1)*dof + k] += lR[(a-1)*dof + k];
    for (b = 1; b <= d; b++) {
      colN = eqN[b-1];
      if (colN == 0) continue;
      left  = rowPtr[rowN-1];
      right = rowPtr[rowN];
      ptr   = (right + left)/2;
      while (colN != colPtr[ptr-1]) {
        if (colN > colPtr[ptr-1]) left = ptr;
        else right = ptr;
        ptr = (right + left)/2;
      }
      for (k = 0; k < dd; k++)
        Val[(size_t)(ptr-1)*dd + k] += lK[((size_t)(b-1)*d + (a-1))*dd + k];
    }
  }
}

/* S/FLUID.f:40-190 CONSTRUCT_FLUID (single fluid domain, TET4).
 * faithful != 0 additionally emulates the reference's per-element costs that
 * do not change the result (ALLOCATE/DEALLOCATE at :101-103,170 and the two
 * GETTHOODFS deep copies at :98,141) -- used only by the CPU-baseline timing.
 * Returns the number of elements with ISZERO(Jac). */
int ora_construct_fluid(const ora_fluid_par_t *par, int nEl, const int *IEN,
                        const double *x, const double *Ag, const double *Yg,
                        const double *Bf, const int *rowPtr, const int *colPtr,
                        double *R, double *Val, int faithful) {
  double xl[4][3], al[4][4], yl[4][4], bfl[4][3], lR[4][4], lK[4][4][16];
  int ptr[4], e, a, i, Ac, nbad = 0;
  for (e = 0; e < nEl; e++) {
    int flag = 0;
    void *t1 = 0, *t2 = 0, *t3 = 0;
    for (a = 0; a < 4; a++) {
      Ac = IEN[4*(size_t)e + a];
      ptr[a] = Ac;
      for (i = 0; i < 3; i++) xl[a][i] = x[3*(size_t)(Ac-1) + i];
      for (i = 0; i < 4; i++) al[a][i] = Ag[4*(size_t)(Ac-1) + i];
      for (i = 0; i < 4; i++) yl[a][i] = Yg[4*(size_t)(Ac-1) + i];
      for (i = 0; i < 3; i++) bfl[a][i] = Bf[3*(size_t)(Ac-1) + i];
    }
    if (faithful) {
      /* fs(1), fs(2): w(4), xi(3,4), N(4,4), Nx(3,4,4), Nxx(6,4,4) each; plus
       * xwl,Nwx,Nwxx,xql,Nqx work arrays */
      t1 = calloc(2*(4 + 12 + 16 + 48 + 96), sizeof(double));
      t2 = calloc(2*(4 + 12 + 16 + 48 + 96), sizeof(double));
      t3 = calloc(12 + 12 + 24 + 12 + 12, sizeof(double));
    }
    ora_fluid_element(par, xl, al, yl, bfl, lR, lK, &flag);
    if (faithful) { free(t1); free(t2); free(t3); }
    nbad += flag;
    ora_doassem(4, 4, ptr, &lK[0][0][0], &lR[0][0], rowPtr, colPtr, R, Val);
  }
  return nbad;
}

/* S/HEATS.f:39-113 CONSTRUCT_HEATS; tDof = dof = 1 so Ag, Yg are [tnNo]. */
int ora_construct_heats(const ora_heat_par_t *par, int nEl, const int *IEN,
                        const double *x, const double *Ag, const double *Yg,
                        const int *rowPtr, const int *colPtr, double *R,
                        double *Val) {
  double xl[4][3], al[4], yl[4], lR[4], lK[4][4], lKf[16];
  int ptr[4], e, a, b, i, Ac, nbad = 0;
  for (e = 0; e < nEl; e++) {
    int flag = 0;
    for (a = 0; a < 4; a++) {
      Ac = IEN[4*(size_t)e + a];
      ptr[a] = Ac;
      for (i = 0; i < 3; i++) xl[a][i] = x[3*(size_t)(Ac-1) + i];
      al[a] = Ag[Ac-1];
      yl[a] = Yg[Ac-1];
    }
    ora_heat_element(par, xl, al, yl, lR, lK, &flag);
    nbad += flag;
    for (a = 0; a < 4; a++)
      for (b = 0; b < 4; b++) lKf[b*4 + a] = lK[a][b]; /* lK(1,a,b) */
    ora_doassem(1, 4, ptr, lKf, lR, rowPtr, colPtr, R, Val);
  }
  return nbad;
}

double ora_wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9*(double)ts.tv_nsec;
}
