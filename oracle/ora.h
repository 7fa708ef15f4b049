/*
 * oracle/ora.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, FP64, int32 1-based ids) of the svFSI / svFSILS
 * fluid Newton-iteration hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (svfsi_b200/csrc) never links, imports or calls it.
 *
 * PARITY PIN: the reference ships no golden vectors and no tests and cannot be
 * COMPILED in this image (no Fortran compiler, no MPI) -- SURVEY.md 8c.  The
 * restatement follows the cited Fortran lines statement by statement and is
 * pinned to the reference's own SOURCE TEXT: oracle/refexec.py executes the
 * reference's routines from /root/reference/Code/Source (one and several MPI
 * tasks), tests/golden/make_ref_golden.py commits their results, and
 * tests/test_reference_golden.py requires this library to reproduce them --
 * it does, bit for bit (element loops, FSILS solvers, time integrator).  Not
 * covered by that pin: a compiled run's compiler and MPI-library rounding
 * (1e-16 effects).  Independent checks on top (tests/test_oracle_*.py):
 * tangent vs finite differences of the residual, SpMV vs SciPy BSR, solver
 * residual checks, 1-rank vs k-rank agreement.
 *
 * All citations are relative to /root/reference/Code/Source
 * (S/ = svFSI/, L/ = svFSILS/).
 */
#ifndef SVFSI_ORACLE_H
#define SVFSI_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums, L/FSILS_STRUCT.h:53-62 --------------------------------- */
#define ORA_LS_TYPE_CG 798
#define ORA_LS_TYPE_GMRES 797
#define ORA_LS_TYPE_NS 796
#define ORA_LS_TYPE_BICGS 795
#define ORA_PRECOND_FSILS 701
#define ORA_PRECOND_RCS 709
#define ORA_BC_TYPE_DIR 0
#define ORA_BC_TYPE_NEU 1
#define ORA_BCOP_TYPE_ADD 0
#define ORA_BCOP_TYPE_PRE 1

/* ---- element level -------------------------------------------------- */
typedef struct {
  double rho, mu, f[3];      /* S/FLUID.f:212-216, GETVISCOSITY Const :1700-1703 */
  double dt, af, am, gam;    /* S/FLUID.f:218-219 */
} ora_fluid_par_t;

typedef struct {
  double nu, s, rho;         /* S/HEATS.f:127-129 */
  double dt, af, am, gam;
} ora_heat_par_t;

void ora_tet4_tables(double w[4], double N[4][4], double Nxi[4][3]);
void ora_gnn3(const double Nxi[4][3], const double xl[4][3], double Nx[4][3],
              double *Jac, double ks[3][3]);
int ora_iszero(double ia);
void ora_fluid_element(const ora_fluid_par_t *par, const double xl[4][3],
                       const double al[4][4], const double yl[4][4],
                       const double bfl[4][3], double lR[4][4],
                       double lK[4][4][16], int *jac_flag);
void ora_heat_element(const ora_heat_par_t *par, const double xl[4][3],
                      const double al[4], const double yl[4], double lR[4],
                      double lK[4][4], int *jac_flag);

int ora_lhsa(int tnNo, int nEl, const int *IEN, int **rowPtr_out,
             int **colPtr_out, int *nnz_out);
void ora_free(void *p);
void ora_doassem(int dof, int d, const int *eqN, const double *lK,
                 const double *lR, const int *rowPtr, const int *colPtr,
                 double *R, double *Val);
int ora_construct_fluid(const ora_fluid_par_t *par, int nEl, const int *IEN,
                        const double *x, const double *Ag, const double *Yg,
                        const double *Bf, const int *rowPtr, const int *colPtr,
                        double *R, double *Val, int faithful);
int ora_construct_heats(const ora_heat_par_t *par, int nEl, const int *IEN,
                        const double *x, const double *Ag, const double *Yg,
                        const int *rowPtr, const int *colPtr, double *R,
                        double *Val);

/* ---- FSILS: a "world" of nTasks simulated ranks in one process ------ */
typedef struct {
  int iP;     /* 1-based neighbour rank, L/FSILS_STRUCT.h:119 */
  int n;
  int *ptr;   /* 1-based reordered local ids */
} ora_cs_t;

typedef struct {
  int foC, incFlag, coupledFlag, sharedFlag;
  int nNo, dof, bGrp;
  int *glob;          /* 1-based reordered ids */
  double nS, res;
  double *val, *valM; /* [nNo][dof] */
} ora_face_t;

typedef struct {
  int gnNo, nNo, nnz, nFaces, mynNo, shnNo, nReq;
  int *colPtr;   /* [nnz] reordered column ids, 1-based */
  int *rowPtr;   /* [nNo][2] (first,last) inclusive, 1-based, by reordered row */
  int *diagPtr;  /* [nNo] */
  int *map;      /* [nNo] svFSI local id -> reordered id, 1-based */
  ora_cs_t *cS;
  ora_face_t *face;
} ora_lhs_t;

typedef struct {
  int nTasks;
  ora_lhs_t *lhs; /* [nTasks] */
} ora_world_t;

typedef struct {
  int suc, mItr, sD, itr;
  double absTol, relTol, iNorm, fNorm, dB, callD;
} ora_subls_t;

typedef struct {
  int LS_type, Resm, Resc;
  ora_subls_t GM, CG, RI;
} ora_ls_t;

ora_world_t *ora_world_create(int nTasks, int gnNo, const int *nNo,
                              const int *nnz, const int *const *gNodes,
                              const int *const *rowPtr,
                              const int *const *colPtr, int nFaces);
void ora_world_free(ora_world_t *w);
/* queries used by the tests */
void ora_world_info(const ora_world_t *w, int rank, int *mynNo, int *shnNo,
                    int *nReq);
void ora_world_map(const ora_world_t *w, int rank, int *map);
void ora_world_rowptr(const ora_world_t *w, int rank, int *rowPtr2,
                      int *colPtr, int *diagPtr);
int ora_world_cs(const ora_world_t *w, int rank, int i, int *iP, int *n,
                 int *ptr);

void ora_bc_create(ora_world_t *w, int faIn, const int *nNo, int dof,
                   int BC_type, const int *const *gNodes,
                   const double *const *Val);
void ora_ls_create(ora_ls_t *ls, int LS_type);

void ora_commuv(const ora_world_t *w, int dof, double *const *R);
void ora_sparmul_vv(const ora_world_t *w, int dof, const double *const *K,
                    const double *const *U, double *const *KU);
void ora_sparmul_vs(const ora_world_t *w, int dof, const double *const *K,
                    const double *const *U, double *const *KU);
void ora_sparmul_sv(const ora_world_t *w, int dof, const double *const *K,
                    const double *const *U, double *const *KU);
void ora_sparmul_ss(const ora_world_t *w, const double *const *K,
                    const double *const *U, double *const *KU);
double ora_dotv(const ora_world_t *w, int dof, const double *const *U,
                const double *const *V);
double ora_normv(const ora_world_t *w, int dof, const double *const *U);

void ora_fsils_solve(ora_world_t *w, ora_ls_t *ls, int dof, double *const *Ri,
                     double *const *Val, int prec, const int *incL,
                     const double *res);

/* single-rank timing helpers for bench.py's cpu_baseline leg */
double ora_wtime(void);

#ifdef __cplusplus
}
#endif
#endif
