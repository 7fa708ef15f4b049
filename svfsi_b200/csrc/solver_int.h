// solver_int.h -- pieces of solver.cu shared with nssolver.cu
#pragma once
#include "core.h"

namespace svfsi {

// Mapped pinned words written by the device.  Every Krylov iteration ends with a publish step that
// stores (sequence number << 1 | stop flag AS OF THAT ITERATION) in flag[seq & 63] -- one word, one
// store, no system-scope fence on the device side (`progress` is no longer written).
// The host enqueues iteration i, then waits for the flag of iteration i-1: it never idles the
// GPU, and -- because the flag is a function of all-reduced scalars only -- every rank reads the
// same value for the same iteration and so enqueues the same number of NCCL operations.
struct HostMirror {
  volatile int progress;
  volatile int flag[64];
};
extern HostMirror *g_hm, *g_hm_dev;
extern int g_seq;
int ensure_mirror();
int publish(const KrylovCtl *ctl);          // enqueue; returns the sequence number
int wait_flag(int seq, int *flag);           // block until `seq` was published; 0 ok

struct Bump {
  char *base;
  size_t off = 0;
  explicit Bump(void *b) : base((char *)b) {}
  double *take(size_t nd) {
    double *p = (double *)(base + off);
    off += ((nd * sizeof(double) + 255) / 256) * 256;
    return p;
  }
};
size_t padded(size_t nd);

// layout of the scalar area used by GMRES
struct GmresScal {
  KrylovCtl *ctl;
  double *hcol, *h, *cc, *ss, *err, *y, *coef, *faceS, *tmp;
  size_t doubles;
};
GmresScal gmres_scal(double *base, int sD, int nFaces);
int ensure_small_n(size_t nd);

bool any_coupled();
int addbcmul(int op, int dof, const double *X, double *Y, double *sS, const int *done, bool dotsDone = false);
bool addbcmul_fork_dots(int dof, const double *X, const int *done);
int bcpre(int nsd, double *sS);
int dot_dev(const double *U, const double *V, size_t nOwned, double *out, const int *done);
double now_s();

// GMRES(lhs, ls, dof, Val, R, X) out of place (L/GMRES.f:51-169); work >= gmres_out_work(...)
size_t gmres_out_work(int sD, size_t n);
int gmres_outofplace(svfsi_subls_t *ls, int dof, const double *Val, const double *R, double *X,
                     double *work, double *scal);
// CGRAD_SCHUR (L/CGRAD.f:51-123); work >= cg_schur_work(n, dof)
size_t cg_schur_work(size_t nNo, int dof);
int cgrad_schur(svfsi_subls_t *ls, int dof, const double *D, const double *G, const double *L,
                double *R, double *work, double *scal);

// bicgs_rcs.cu: BICGSS/BICGSV (L/BICGS.f:50-180) and PRECONDRCS (L/PRECOND.f:150-368)
int bicgs(svfsi_subls_t *ls, int dof, const double *K, double *R);
int precondrcs(int dof, double *Val, double *R, double *W2, double *work);

}  // namespace svfsi
