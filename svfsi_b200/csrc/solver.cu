// solver.cu -- host orchestration of FSILS_SOLVE on the device (L/SOLVE.f:51-143):
// PRECONDDIAG (L/PRECOND.f:50-145), GMRESV/GMRESS (L/GMRES.f:171-431), GMRES
// out-of-place (L/GMRES.f:51-169), CGRADS/CGRADV/CGRAD_SCHUR (L/CGRAD.f),
// NSSOLVER (L/NSSOLVER.f:52-233) with GE (L/GE.f).
//
// The Krylov loops never synchronise with the host per iteration: all scalars
// (Hessenberg column, Givens rotations, residual estimate, CG alpha/beta) live
// in device memory, the convergence test sets a device flag that turns the
// remaining queued kernels of the cycle into no-ops, and the host watches a
// progress word in mapped pinned memory to stay a bounded number of iterations
// ahead of the GPU.
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <chrono>

#include "core.h"
#include "solver_int.h"

namespace svfsi {


HostMirror *g_hm = nullptr;      // host pointer
HostMirror *g_hm_dev = nullptr;  // device alias
int g_seq = 0;

int ensure_mirror() {
  if (g_hm) return 0;
  void *p = nullptr;
  CUDA_TRY(cudaHostAlloc(&p, sizeof(HostMirror), cudaHostAllocMapped));
  g_hm = (HostMirror *)p;
  memset((void *)g_hm, 0, sizeof(HostMirror));
  void *d = nullptr;
  CUDA_TRY(cudaHostGetDevicePointer(&d, p, 0));
  g_hm_dev = (HostMirror *)d;
  return 0;
}

// progress/done publication + small scalar steps
__global__ void publish_kernel(HostMirror *hm, int seq, const KrylovCtl *ctl) {
  hm->flag[seq & 63] = (seq << 1) | (ctl->done ? 1 : 0);
}

int publish(const KrylovCtl *ctl) {
  Ctx &c = ctx();
  g_seq++;
  publish_kernel<<<1, 1, 0, c.stream>>>(g_hm_dev, g_seq, ctl);
  count_launch();
  return g_seq;
}

int wait_flag(int seq, int *flag) {
  Ctx &c = ctx();
  unsigned long spins = 0;
  // slot word = (sequence number << 1) | stop flag, written with ONE store (no fence on the device side)
  while ((g_hm->flag[seq & 63] >> 1) != seq) {
    if (c.h_status && c.h_status[0]) return comm_check();   // an in-kernel peer wait timed out
    if ((++spins & 0xFFFFF) == 0) {  // every ~1M polls make sure the stream is still alive
      cudaError_t e = cudaStreamQuery(c.stream);
      if (e != cudaSuccess && e != cudaErrorNotReady)
        return fail(SVFSI_ERR_CUDA, std::string("Krylov loop: ") + cudaGetErrorString(e));
      if (e == cudaSuccess && (g_hm->flag[seq & 63] >> 1) != seq)
        return fail(SVFSI_ERR_CUDA, "Krylov loop: stream drained without publishing progress");
    }
  }
  *flag = g_hm->flag[seq & 63] & 1;
  return 0;
}

__global__ void ctl_reset_kernel(KrylovCtl *ctl, int clear_all) {
  ctl->done = 0;
  ctl->ilast = 0;
  if (clear_all) {
    ctl->suc = 0;
    ctl->itr = 0;
  }
}

// eps / iNorm from the squared norm in *ss (L/GMRES.f:296-300)
__global__ void gmres_init_kernel(KrylovCtl *ctl, const double *ss, double absTol, double relTol) {
  const double nrm = sqrt(*ss);
  ctl->iNorm = nrm;
  ctl->fNorm = nrm;
  const double e = relTol * nrm;
  ctl->eps = absTol > e ? absTol : e;
}

// err(1) = ||u1|| (L/GMRES.f:324)
__global__ void gmres_err0_kernel(const double *ss, double *err) { err[0] = sqrt(*ss); }

// ------------------------------------------------------------------ CG scalars
// scal[0] = err, scal[1] = errO, scal[2] = alpha, scal[3] = errO/err, scal[4] = err/errO
__global__ void cg_init_kernel(KrylovCtl *ctl, const double *ss, double absTol, double relTol) {
  const double nrm = sqrt(*ss);
  ctl->iNorm = nrm;
  const double e0 = relTol * nrm;
  double eps = absTol > e0 ? absTol : e0;
  eps = eps * eps;
  ctl->eps = eps;
  const double err = nrm * nrm;
  ctl->scal[0] = err;
  ctl->scal[1] = err;
  ctl->done = 0;
  ctl->suc = 0;
  ctl->ilast = 0;
  if (err < eps) {  // loop exits at i = 1 with suc (L/CGRAD.f:151-154)
    ctl->suc = 1;
    ctl->done = 1;
  }
}
// alpha = errO / <P, KP>   (L/CGRAD.f:159)
__global__ void cg_alpha_kernel(KrylovCtl *ctl, const double *pkp) {
  if (ctl->done) return;
  ctl->scal[1] = ctl->scal[0];  // errO = err
  ctl->scal[2] = ctl->scal[1] / *pkp;
}
// err = ||R||^2, convergence test of the NEXT loop trip (L/CGRAD.f:151-154,164-165)
__global__ void cg_err_kernel(KrylovCtl *ctl, const double *rr, int mItr, volatile int *pubFlag,
                              volatile int *pubProgress, int seq) {
  if (!ctl->done) {
  double e = sqrt(*rr);
  e = e * e;
  ctl->scal[0] = e;
  ctl->scal[3] = ctl->scal[1] / e;
  ctl->scal[4] = e / ctl->scal[1];
  ctl->ilast += 1;  // completed iterations
  // the reference still updates P before leaving the loop at the next trip's test;
  // P is dead after that, so stopping here gives the same X, R, err, itr
  // (if this was trip mItr the loop ends without the test: suc stays .FALSE., L/CGRAD.f:150-154)
  if (ctl->ilast < mItr && e < ctl->eps) {
    ctl->suc = 1;
    ctl->done = 1;
  }
  }
  *pubFlag = (seq << 1) | (ctl->done ? 1 : 0);
}

// X += alpha P ; R -= alpha KP ; partial sums of R.R over the owned range
__global__ void __launch_bounds__(256) cg_update_kernel(const KrylovCtl *ctl, double *__restrict__ X,
                                                        double *__restrict__ R,
                                                        const double *__restrict__ P,
                                                        const double *__restrict__ KP, size_t n,
                                                        size_t nOwned, double *__restrict__ partial) {
  if (ctl->done) return;
  __shared__ double smem[8];
  const double alpha = ctl->scal[2];
  double acc = 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    X[e] = X[e] + alpha * P[e];
    const double r = R[e] + (-alpha) * KP[e];
    R[e] = r;
    if (e < nOwned) acc = fma(r, r, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += smem[w];
    partial[blockIdx.x] = t;
  }
}
// P = (P + (errO/err) R) * (err/errO)   (L/CGRAD.f:166-167: OMPSUM then OMPMUL)
__global__ void __launch_bounds__(256) cg_pupdate_kernel(const KrylovCtl *ctl, double *__restrict__ P,
                                                         const double *__restrict__ R, size_t n) {
  if (ctl->done) return;
  const double s1 = ctl->scal[3], s2 = ctl->scal[4];
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x)
    P[e] = (P[e] + s1 * R[e]) * s2;
}

size_t padded(size_t nd) { return ((nd * sizeof(double) + 255) / 256) * 256; }

bool any_coupled() {
  for (const Face &f : ctx().face)
    if (f.created && f.coupled) return true;
  return false;
}

// ADDBCMUL (L/ADDBCMUL.f:53-114).  sS = device scratch scalars (one per face).
// The dot half of ADDBCMUL(ADD) for the faces that live on this rank alone, on the SIDE stream: it only needs
// X = u(i), so it runs under the SpMV that produces u(i+1); addbcmul(..., dotsDone = true) then joins and only
// launches the updates.  Returns true if anything was forked.
bool addbcmul_fork_dots(int dof, const double *X, const int *done) {
  Ctx &c = ctx();
  bool any = false;
  for (size_t fi = 0; fi < c.face.size(); fi++) {
    Face &f = c.face[fi];
    if (!f.created || !f.coupled || f.shared || f.nNo == 0) continue;
    if (!any) {
      cudaEventRecord(c.evFork, c.stream);          // u(i) is complete
      cudaStreamWaitEvent(c.stream2, c.evFork, 0);
      any = true;
    }
    launch_face_dotp(c.stream2, f.nNo, f.dof, dof, f.d_glob, f.d_valM, X, c.nNo,
                     c.d_partial + c.partialDoubles - 64 * (fi + 1), done);
  }
  if (any) cudaEventRecord(c.evJoin, c.stream2);
  return any;
}

int addbcmul(int op, int dof, const double *X, double *Y, double *sS, const int *done, bool dotsDone) {
  Ctx &c = ctx();
  if (dotsDone) cudaStreamWaitEvent(c.stream, c.evJoin, 0);
  for (size_t fi = 0; fi < c.face.size(); fi++) {
    Face &f = c.face[fi];
    if (!f.created || !f.coupled) continue;
    const double coef = (op == 0) ? f.res : -f.res / (1.0 + (f.res * f.nS));
    double *S = sS + fi;
    if (f.shared) {
      launch_face_dot(c.stream, f.nNo, f.dof, dof, f.d_glob, f.d_valM, X, c.mynNo, S, done);
      if (int rc = allreduce_dev(S, 1)) return rc;
    } else {
      // the face lives on one rank (L/ADDBCMUL.f:80-92, not sharedFlag): nothing to do elsewhere, one launch here
      if (f.nNo > 0) {   // scratch for the per-CTA partials: behind the multi-dot partials (never live together)
        double *part = c.d_partial + c.partialDoubles - 64 * (fi + 1);
        if (!(dotsDone && op == 0))
          launch_face_dotp(c.stream, f.nNo, f.dof, dof, f.d_glob, f.d_valM, X, c.nNo, part, done);
        launch_face_axpyp(c.stream, f.nNo, f.dof, dof, f.d_glob, f.d_valM, coef, part, S, Y, done);
      }
      continue;
    }
    launch_face_axpy(c.stream, f.nNo, f.dof, dof, f.d_glob, f.d_valM, coef, S, Y, done);
  }
  return 0;
}

// BCPRE (L/GMRES.f:393-429, L/NSSOLVER.f:307-341): nS = ||valM||^2
int bcpre(int nsd, double *sS) {
  Ctx &c = ctx();
  bool any = false;
  for (size_t fi = 0; fi < c.face.size(); fi++) {
    Face &f = c.face[fi];
    if (!f.created || !f.coupled) continue;
    any = true;
    launch_face_norm2(c.stream, f.nNo, f.dof, nsd, f.d_glob, f.d_valM,
                      f.shared ? c.mynNo : c.nNo, sS + fi);
    if (f.shared)
      if (int rc = allreduce_dev(sS + fi, 1)) return rc;
  }
  if (!any) return 0;
  CUDA_TRY(cudaMemcpyAsync(c.h_small, sS, sizeof(double) * c.face.size(), cudaMemcpyDeviceToHost,
                           c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  for (size_t fi = 0; fi < c.face.size(); fi++) {
    Face &f = c.face[fi];
    if (!f.created || !f.coupled) continue;
    double v = c.h_small[fi];
    if (f.shared) {  // NORMV(...)**2: sqrt then square (L/GMRES.f:413-414)
      v = sqrt(v);
      v = v * v;
    }
    f.nS = v;
  }
  return 0;
}

// squared 2-norm / dot over owned entries -> *out (device), all-reduced
int dot_dev(const double *U, const double *V, size_t nOwned, double *out, const int *done) {
  // one kernel: partial sums + block sums (+ peer all-reduce)
  return multidot_column(U, 0, const_cast<double *>(V), nOwned, 1, out, nullptr, done, false);
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

GmresScal gmres_scal(double *base, int sD, int nFaces) {
  GmresScal g;
  size_t o = 0;
  g.ctl = (KrylovCtl *)base;
  o += 32;
  g.hcol = base + o; o += sD + 2;
  g.h = base + o; o += (size_t)(sD + 1) * sD;
  g.cc = base + o; o += sD;
  g.ss = base + o; o += sD;
  g.err = base + o; o += sD + 1;
  g.y = base + o; o += sD;
  g.coef = base + o; o += sD + 1;
  g.faceS = base + o; o += nFaces + 1;
  g.tmp = base + o; o += 8;
  g.doubles = o;
  return g;
}

static size_t g_smallCap = 1 << 17;
static double *g_dW = nullptr, *g_dRcs = nullptr;   // W of PRECONDDIAG; Wr, Wc, W1 of PRECONDRCS
static size_t g_wCap = 0, g_rcsCap = 0;
void nssolver_free_static();   // nssolver.cu

// gpu_finalize_: everything this file and nssolver.cu keep between calls belongs to the device that
// is going away
void solver_free_static() {
  if (g_dW) cudaFree(g_dW);
  if (g_dRcs) cudaFree(g_dRcs);
  g_dW = g_dRcs = nullptr;
  g_wCap = g_rcsCap = 0;
  g_smallCap = 1 << 17;
  if (g_hm) cudaFreeHost((void *)g_hm);
  g_hm = g_hm_dev = nullptr;
  g_seq = 0;
  nssolver_free_static();
}

int ensure_small_n(size_t nd) {
  Ctx &c = ctx();
  if (int rc = ensure_small()) return rc;
  size_t &cap = g_smallCap;
  if (nd <= cap) return 0;
  cudaFree(c.d_small);
  cudaFreeHost(c.h_small);
  c.d_small = nullptr;
  c.h_small = nullptr;
  cap = nd + 1024;
  CUDA_TRY(cudaMalloc(&c.d_small, cap * sizeof(double)));
  CUDA_TRY(cudaMemset(c.d_small, 0, cap * sizeof(double)));
  CUDA_TRY(cudaMallocHost(&c.h_small, cap * sizeof(double)));
  return 0;
}

// One Arnoldi cycle (L/GMRES.f:327-367 and the shared core of :106-152): u[0]
// holds the normalised residual, err[0] its norm.  Enqueues up to sD columns and
// stops enqueueing once the device reports convergence.
int arnoldi_cycle(int kind, int dof, const double *Val, double *u, size_t stride, size_t n,
                  size_t nOwned, int sD, GmresScal &g, bool pre, double *unCondU) {
  Ctx &c = ctx();
  const int *done = &g.ctl->done;
  int seqPrev = -1;
  for (int i = 1; i <= sD; i++) {
    double *ui = u + (size_t)i * stride, *um = u + (size_t)(i - 1) * stride;
    // the SpMV's halo RECEIVE rides on the column kernel below (its first CTAs wait for the neighbours
    // and add what arrived) unless something reads the whole of u(i+1) before: the BCOP_TYPE_PRE step
    const bool preStep = (kind == 0) && pre && any_coupled();
    bool pend = false;
    const bool forked = (kind == 0) && addbcmul_fork_dots(dof, um, done);   // valM . u(i) under the SpMV
    if (int rc = sparmul(kind, dof, Val, um, ui, done, preStep ? nullptr : &pend)) return rc;
    if (kind == 0) {
      // rank-one face term: reads u(i), adds to u(i+1) on the face nodes (sums commute with the receive)
      if (int rc = addbcmul(0, dof, um, ui, g.faceS, done, forked)) return rc;
      if (preStep) {
        launch_vecop(c.stream, VOP_COPY, unCondU, ui, nullptr, n, nullptr, 0.0, done);
        if (int rc = addbcmul(1, dof, unCondU, ui, g.faceS, done)) return rc;
      }
    }
    const int seq = ++g_seq;
    {
      // [halo receive] + multi-dot + block sums + all-reduce + Givens column + stop test + flag
      // publication: one kernel
      ColArgs col{g.ctl, i, sD, g.h, g.cc, g.ss, g.err, g.coef, &g_hm_dev->flag[seq & 63],
                  &g_hm_dev->progress, seq};
      if (int rc = multidot_column(u, stride, ui, nOwned, i + 1, g.hcol, &col, done, pend)) return rc;
    }
    {
      ProfScope ps(PROF_AXPY);
      // the column kernel may have set `done`; the update of u(i+1) must still run
      // for the column that converged?  No: after EXIT the reference never uses
      // u(:,:,i+1) (L/GMRES.f:363-381 sums j = 1..i), so skipping is exact.
      launch_multi_axpy_scale(c.stream, u, stride, ui, n, i, g.coef, &g.ctl->inv, done);
    }
    if (seqPrev >= 0) {  // flag of the PREVIOUS iteration: deterministic across ranks
      int flag = 0;
      if (int rc = wait_flag(seqPrev, &flag)) return rc;
      if (flag) break;
    }
    seqPrev = seq;
  }
  return 0;
}

// GMRESV / GMRESS in place (L/GMRES.f:273-431, :171-271)
int gmres_inplace(svfsi_subls_t *ls, int dof, const double *Val, double *R, bool scalar) {
  Ctx &c = ctx();
  const int sD = ls->sD;
  const size_t n = (size_t)c.nNo * dof, nOwned = (size_t)c.mynNo * dof;
  const size_t stride = padded(n) / sizeof(double);
  const int kind = scalar ? 3 : 0;
  if (int rc = ensure_mirror()) return rc;
  GmresScal g = gmres_scal(nullptr, sD, (int)c.face.size());
  if (int rc = ensure_small_n(g.doubles)) return rc;
  g = gmres_scal(c.d_small, sD, (int)c.face.size());
  // workspace: u(sD+1), X
  if (int rc = ensure_ws((size_t)(sD + 2) * stride * sizeof(double) + 4096)) return rc;
  Bump b((char *)c.d_ws);
  double *u = b.take((size_t)(sD + 1) * stride);
  double *X = b.take(n);
  const int *nodone = nullptr;

  const double t0 = now_s();
  ls->suc = 0;
  ctl_reset_kernel<<<1, 1, 0, c.stream>>>(g.ctl, 1);
  count_launch();
  if (int rc = dot_dev(R, R, nOwned, g.tmp, nodone)) return rc;
  gmres_init_kernel<<<1, 1, 0, c.stream>>>(g.ctl, g.tmp, ls->absTol, ls->relTol);
  count_launch();
  launch_vecop(c.stream, VOP_ZERO, X, nullptr, nullptr, n, nullptr, 0.0, nodone);
  if (!scalar)
    if (int rc = bcpre(dof - 1, g.faceS)) return rc;
  KrylovCtl hc;
  CUDA_TRY(cudaMemcpyAsync(&hc, g.ctl, sizeof(hc), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  ls->iNorm = hc.iNorm;
  ls->fNorm = hc.iNorm;
  ls->itr = 0;
  if (ls->iNorm <= ls->absTol) {
    ls->callD = DBL_EPSILON;
    ls->dB = 0.0;
    return 0;
  }
  int itrHost = 0;
  for (int l = 1; l <= ls->mItr; l++) {
    ls->dB = ls->fNorm;
    itrHost++;
    ctl_reset_kernel<<<1, 1, 0, c.stream>>>(g.ctl, 0);
    count_launch();
    if (l == 1) {
      // X = 0 in the first cycle: K X (and the rank-one face term) is exactly zero, so u = R - K X = R bit for
      // bit.  The reference multiplies by the zero vector (L/GMRES.f:313-318) and counts it in ls%itr; the count
      // is kept (itrHost above), the 3.5 GB pass over Val is not made.
      launch_vecop(c.stream, VOP_COPY, u, R, nullptr, n, nullptr, 0.0, nodone);
    } else {
      if (int rc = sparmul(kind, dof, Val, X, u, nodone)) return rc;
      if (!scalar)
        if (int rc = addbcmul(0, dof, X, u, g.faceS, nodone)) return rc;
      launch_vecop(c.stream, VOP_SUB_FROM, u, R, nullptr, n, nullptr, 0.0, nodone);
    }
    if (int rc = dot_dev(u, u, nOwned, g.tmp, nodone)) return rc;
    gmres_err0_kernel<<<1, 1, 0, c.stream>>>(g.tmp, g.err);
    count_launch();
    launch_vecop(c.stream, VOP_DIV_DEV, u, nullptr, nullptr, n, g.err, 0.0, nodone);
    if (int rc = arnoldi_cycle(kind, dof, Val, u, stride, n, nOwned, sD, g, false, nullptr)) return rc;
    {
      ProfScope ps(PROF_SMALL);
      launch_gmres_backsub(c.stream, g.ctl, sD, g.h, g.err, g.y);
    }
    {
      ProfScope ps(PROF_AXPY);
      launch_multi_axpy_acc(c.stream, u, stride, X, n, &g.ctl->ilast, sD, g.y);
    }
    CUDA_TRY(cudaMemcpyAsync(&hc, g.ctl, sizeof(hc), cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    ls->fNorm = hc.fNorm;
    if (hc.suc) {
      ls->suc = 1;
      break;
    }
  }
  ls->itr = itrHost + hc.itr;
  launch_vecop(c.stream, VOP_COPY, R, X, nullptr, n, nullptr, 0.0, nodone);
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  ls->callD = now_s() - t0;
  ls->dB = 10.0 * log(ls->fNorm / ls->dB);
  return 0;
}

// CGRADS / CGRADV (L/CGRAD.f:125-242)
int cgrad(svfsi_subls_t *ls, int dof, const double *K, double *R) {
  Ctx &c = ctx();
  const size_t n = (size_t)c.nNo * dof, nOwned = (size_t)c.mynNo * dof;
  const int kind = dof == 1 ? 3 : 0;
  if (int rc = ensure_mirror()) return rc;
  if (int rc = ensure_small_n(4096)) return rc;
  KrylovCtl *ctl = (KrylovCtl *)c.d_small;
  double *sc = c.d_small + 32;  // [0] = <P,KP> / R.R scratch
  if (int rc = ensure_ws(3 * padded(n) + 4096)) return rc;
  Bump b((char *)c.d_ws);
  double *P = b.take(n), *KP = b.take(n), *X = b.take(n);
  const int *done = &ctl->done;
  const int nblk = multidot_nblk();

  const double t0 = now_s();
  if (int rc = dot_dev(R, R, nOwned, sc, nullptr)) return rc;
  cg_init_kernel<<<1, 1, 0, c.stream>>>(ctl, sc, ls->absTol, ls->relTol);
  count_launch();
  launch_vecop(c.stream, VOP_COPY, P, R, nullptr, n, nullptr, 0.0, nullptr);
  launch_vecop(c.stream, VOP_ZERO, X, nullptr, nullptr, n, nullptr, 0.0, nullptr);
  int seqPrev = publish(ctl);
  for (int i = 1; i <= ls->mItr; i++) {
    // the halo receive of K P rides on the <P, K P> kernel (its first CTAs wait for the neighbours)
    bool pend = false;
    if (int rc = sparmul(kind, dof, K, P, KP, done, &pend)) return rc;
    if (int rc = multidot_column(P, 0, KP, nOwned, 1, sc, nullptr, done, pend)) return rc;
    cg_alpha_kernel<<<1, 1, 0, c.stream>>>(ctl, sc);
    {
      ProfScope ps(PROF_AXPY);
      cg_update_kernel<<<nblk, 256, 0, c.stream>>>(ctl, X, R, P, KP, n, nOwned, c.d_partial);
    }
    if (int rc = reduce_allreduce(c.d_partial, 1, sc + 1, done)) return rc;
    const int seq = ++g_seq;
    cg_err_kernel<<<1, 1, 0, c.stream>>>(ctl, sc + 1, ls->mItr, &g_hm_dev->flag[seq & 63],
                                         &g_hm_dev->progress, seq);
    {
      ProfScope ps(PROF_AXPY);
      cg_pupdate_kernel<<<148 * 8, 256, 0, c.stream>>>(ctl, P, R, n);
    }
    count_launch(4);
    int flag = 0;
    if (int rc = wait_flag(seqPrev, &flag)) return rc;
    if (flag) break;
    seqPrev = seq;
  }
  launch_vecop(c.stream, VOP_COPY, R, X, nullptr, n, nullptr, 0.0, nullptr);
  KrylovCtl hc;
  CUDA_TRY(cudaMemcpyAsync(&hc, ctl, sizeof(hc), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  ls->suc = hc.suc;
  ls->iNorm = hc.iNorm;
  ls->itr = hc.ilast;
  const double err = hc.scal[0], errO = hc.scal[1];
  ls->fNorm = sqrt(err);
  ls->callD = now_s() - t0;
  if (errO < DBL_EPSILON) ls->dB = 0.0;
  else ls->dB = 5.0 * log(err / errO);
  return 0;
}

// GMRES(lhs, ls, dof, Val, R, X), out of place (L/GMRES.f:51-169): the inner solver of NSSOLVER.
// ls%itr and ls%callD ACCUMULATE over calls; the Sherman-Morrison-like BCOP_TYPE_PRE correction is
// applied whenever a face is coupled (:88-91, :112-115).
size_t gmres_out_work(int sD, size_t n) { return ((size_t)(sD + 2)) * (padded(n) / sizeof(double)) + 64; }

int gmres_outofplace(svfsi_subls_t *ls, int dof, const double *Val, const double *R, double *X,
                     double *work, double *scal) {
  Ctx &c = ctx();
  const int sD = ls->sD;
  const size_t n = (size_t)c.nNo * dof, nOwned = (size_t)c.mynNo * dof;
  const size_t stride = padded(n) / sizeof(double);
  if (int rc = ensure_mirror()) return rc;
  GmresScal g = gmres_scal(scal, sD, (int)c.face.size());
  double *u = work;
  double *unCondU = work + (size_t)(sD + 1) * stride;
  const bool coupled = any_coupled();
  const double t0 = now_s();
  ls->suc = 0;
  ctl_reset_kernel<<<1, 1, 0, c.stream>>>(g.ctl, 1);
  count_launch();
  launch_vecop(c.stream, VOP_ZERO, X, nullptr, nullptr, n, nullptr, 0.0, nullptr);
  KrylovCtl hc;
  memset(&hc, 0, sizeof(hc));
  int itrHost = 0;
  for (int l = 1; l <= ls->mItr; l++) {
    ctl_reset_kernel<<<1, 1, 0, c.stream>>>(g.ctl, 0);
    count_launch();
    if (l == 1) {
      launch_vecop(c.stream, VOP_COPY, u, R, nullptr, n, nullptr, 0.0, nullptr);
    } else {
      itrHost++;
      if (int rc = sparmul(0, dof, Val, X, u, nullptr)) return rc;
      if (int rc = addbcmul(0, dof, X, u, g.faceS, nullptr)) return rc;
      launch_vecop(c.stream, VOP_SUB_FROM, u, R, nullptr, n, nullptr, 0.0, nullptr);
    }
    if (coupled) {
      launch_vecop(c.stream, VOP_COPY, unCondU, u, nullptr, n, nullptr, 0.0, nullptr);
      if (int rc = addbcmul(1, dof, unCondU, u, g.faceS, nullptr)) return rc;
    }
    if (int rc = dot_dev(u, u, nOwned, g.tmp, nullptr)) return rc;
    gmres_err0_kernel<<<1, 1, 0, c.stream>>>(g.tmp, g.err);
    count_launch();
    if (l == 1) {
      // eps = err(1); early return when already below absTol (:95-104)
      gmres_init_kernel<<<1, 1, 0, c.stream>>>(g.ctl, g.tmp, ls->absTol, ls->relTol);
      count_launch();
      CUDA_TRY(cudaMemcpyAsync(&hc, g.ctl, sizeof(hc), cudaMemcpyDeviceToHost, c.stream));
      CUDA_TRY(cudaStreamSynchronize(c.stream));
      if (hc.iNorm <= ls->absTol) {
        ls->callD = DBL_EPSILON;
        ls->dB = 0.0;
        return 0;
      }
      ls->iNorm = hc.iNorm;
      ls->fNorm = hc.iNorm;
    }
    ls->dB = ls->fNorm;
    launch_vecop(c.stream, VOP_DIV_DEV, u, nullptr, nullptr, n, g.err, 0.0, nullptr);
    if (int rc = arnoldi_cycle(0, dof, Val, u, stride, n, nOwned, sD, g, true, unCondU)) return rc;
    {
      ProfScope ps(PROF_SMALL);
      launch_gmres_backsub(c.stream, g.ctl, sD, g.h, g.err, g.y);
    }
    {
      ProfScope ps(PROF_AXPY);
      launch_multi_axpy_acc(c.stream, u, stride, X, n, &g.ctl->ilast, sD, g.y);
    }
    CUDA_TRY(cudaMemcpyAsync(&hc, g.ctl, sizeof(hc), cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    ls->fNorm = hc.fNorm;
    if (hc.suc) {
      ls->suc = 1;
      break;
    }
  }
  ls->itr += itrHost + hc.itr;
  ls->callD = now_s() - t0 + ls->callD;
  ls->dB = 10.0 * log(ls->fNorm / ls->dB);
  return 0;
}

// CGRAD_SCHUR (L/CGRAD.f:51-123): CG on the operator  L p - D (G p)  (NSSOLVER passes D = Gt = -G^T).
size_t cg_schur_work(size_t nNo, int dof) {
  return 4 * (padded(nNo) / sizeof(double)) + 2 * (padded(nNo * dof) / sizeof(double)) + 64;
}

int cgrad_schur(svfsi_subls_t *ls, int dof, const double *D, const double *G, const double *L,
                double *R, double *work, double *scal) {
  Ctx &c = ctx();
  const size_t n = (size_t)c.nNo, nOwned = (size_t)c.mynNo, nv = n * dof;
  if (int rc = ensure_mirror()) return rc;
  KrylovCtl *ctl = (KrylovCtl *)scal;
  double *sc = scal + 32;
  double *faceS = scal + 48;
  Bump b(work);
  double *X = b.take(n), *P = b.take(n), *SP = b.take(n), *DGP = b.take(n);
  double *GP = b.take(nv), *unCondU = b.take(nv);
  const int *done = &ctl->done;
  const int nblk = multidot_nblk();
  const bool coupled = any_coupled();

  const double t0 = now_s();
  if (int rc = dot_dev(R, R, nOwned, sc, nullptr)) return rc;
  cg_init_kernel<<<1, 1, 0, c.stream>>>(ctl, sc, ls->absTol, ls->relTol);
  count_launch();
  launch_vecop(c.stream, VOP_COPY, P, R, nullptr, n, nullptr, 0.0, nullptr);
  launch_vecop(c.stream, VOP_ZERO, X, nullptr, nullptr, n, nullptr, 0.0, nullptr);
  int seqPrev = publish(ctl);
  for (int i = 1; i <= ls->mItr; i++) {
    if (int rc = sparmul(2, dof, G, P, GP, done)) return rc;
    if (coupled) {
      launch_vecop(c.stream, VOP_COPY, unCondU, GP, nullptr, nv, nullptr, 0.0, done);
      if (int rc = addbcmul(1, dof, unCondU, GP, faceS, done)) return rc;
    }
    if (int rc = sparmul(1, dof, D, GP, DGP, done)) return rc;
    if (int rc = sparmul(3, 1, L, P, SP, done)) return rc;
    launch_vecop(c.stream, VOP_AXPY, SP, DGP, nullptr, n, nullptr, -1.0, done);  // SP = SP - DGP
    if (int rc = dot_dev(P, SP, nOwned, sc, done)) return rc;
    cg_alpha_kernel<<<1, 1, 0, c.stream>>>(ctl, sc);
    {
      ProfScope ps(PROF_AXPY);
      cg_update_kernel<<<nblk, 256, 0, c.stream>>>(ctl, X, R, P, SP, n, nOwned, c.d_partial);
    }
    if (int rc = reduce_allreduce(c.d_partial, 1, sc + 1, done)) return rc;
    const int seq = ++g_seq;
    cg_err_kernel<<<1, 1, 0, c.stream>>>(ctl, sc + 1, ls->mItr, &g_hm_dev->flag[seq & 63],
                                         &g_hm_dev->progress, seq);
    {
      ProfScope ps(PROF_AXPY);
      cg_pupdate_kernel<<<148 * 8, 256, 0, c.stream>>>(ctl, P, R, n);
    }
    count_launch(4);
    int flag = 0;
    if (int rc = wait_flag(seqPrev, &flag)) return rc;
    if (flag) break;
    seqPrev = seq;
  }
  launch_vecop(c.stream, VOP_COPY, R, X, nullptr, n, nullptr, 0.0, nullptr);
  KrylovCtl hc;
  CUDA_TRY(cudaMemcpyAsync(&hc, ctl, sizeof(hc), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  ls->suc = hc.suc;
  ls->iNorm = hc.iNorm;
  const double err = hc.scal[0], errO = hc.scal[1];
  ls->fNorm = sqrt(err);
  ls->callD = now_s() - t0 + ls->callD;
  ls->itr = ls->itr + hc.ilast;
  if (errO < DBL_EPSILON) ls->dB = 0.0;
  else ls->dB = 5.0 * log(err / errO);
  return 0;
}

// PRECONDDIAG (L/PRECOND.f:50-145); W receives the scaling (Wc of L/SOLVE.f:98-100)
int preconddiag(int dof, double *Val, double *R, double *W, bool deferValScale) {
  Ctx &c = ctx();
  ProfScope ps(PROF_PRECOND);
  launch_diag_extract(c.stream, c.nNo, dof, c.d_diag, Val, W);
  if (int rc = halo_sum(W, dof, nullptr)) return rc;
  launch_w_finalize(c.stream, (size_t)c.nNo * dof, W);
  for (Face &f : c.face) {
    if (!f.created || !f.inc) continue;
    if (f.bGrp == SVFSI_BC_TYPE_DIR)
      launch_w_dirichlet(c.stream, f.nNo, f.dof, dof, f.d_glob, f.d_val, W);
  }
  if (deferValScale) c.valScaleW = W;   // K <- W K W rides on the first SpMV of the solve (core.cu sparmul)
  else launch_scale_val(c.stream, c.nnz, dof, c.d_rowOf, c.d_col, W, Val);
  launch_vecop(c.stream, VOP_MUL, R, W, nullptr, (size_t)c.nNo * dof, nullptr, 0.0, nullptr);
  for (Face &f : c.face) {
    if (!f.created || !f.coupled) continue;
    launch_face_valM(c.stream, f.nNo, f.dof, dof, f.d_glob, f.d_val, W, f.d_valM);
  }
  return 0;
}

int nssolver_dev(svfsi_ls_t *ls, int dof, const double *Val, double *R);  // nssolver.cu

void ls_defaults(svfsi_ls_t *ls, int LS_type) {
  // FSILS_LS_CREATE, L/LS.f:69-95
  memset(ls, 0, sizeof(*ls));
  ls->LS_type = LS_type;
  switch (LS_type) {
    case SVFSI_LS_TYPE_NS:
      ls->RI.relTol = 0.4; ls->GM.relTol = 1.e-2; ls->CG.relTol = 0.2;
      ls->RI.mItr = 10; ls->GM.mItr = 2; ls->CG.mItr = 500;
      ls->GM.sD = 100; ls->RI.sD = 100;
      break;
    case SVFSI_LS_TYPE_GMRES:
      ls->RI.relTol = 0.1; ls->RI.mItr = 4; ls->RI.sD = 250;
      break;
    case SVFSI_LS_TYPE_CG:
      ls->RI.relTol = 1.e-2; ls->RI.mItr = 1000;
      break;
    case SVFSI_LS_TYPE_BICGS:
      ls->RI.relTol = 1.e-2; ls->RI.mItr = 500;
      break;
    default: break;
  }
  ls->RI.absTol = 1.e-10; ls->GM.absTol = 1.e-10; ls->CG.absTol = 1.e-10;
}

int fsils_solve_dev(svfsi_ls_t *ls, int dof, int prec, const int32_t *incL, const double *res) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "FSILS_SOLVE before FSILS_LHS_CREATE");
  if (c.dof != dof || !c.d_R || !c.d_Val)
    return fail(SVFSI_ERR_STATE, "no device-resident system of this dof");
  // face flags, L/SOLVE.f:69-91
  bool anyNeu = false;
  for (size_t fi = 0; fi < c.face.size(); fi++) {
    Face &f = c.face[fi];
    f.inc = true;
    if (incL && incL[fi] == 0) f.inc = false;
    if (f.bGrp == SVFSI_BC_TYPE_NEU) anyNeu = true;
  }
  if (!res && anyNeu) return fail(SVFSI_ERR_ARG, "FSILS: res is required for Neu surfaces");
  for (size_t fi = 0; fi < c.face.size(); fi++) {
    Face &f = c.face[fi];
    f.coupled = false;
    if (!f.inc) continue;
    if (f.bGrp == SVFSI_BC_TYPE_NEU && res[fi] != 0.0) {
      f.res = res[fi];
      f.coupled = true;
    }
  }
  // "Krylov space dimension": the fused multi-dot / column kernels hold one Arnoldi column (sD + 1
  // inner products) in kArMax-sized shared-memory arrays and partial-sum rows
  if (ls->LS_type == SVFSI_LS_TYPE_GMRES || ls->LS_type == SVFSI_LS_TYPE_NS) {
    const int sDk = (ls->LS_type == SVFSI_LS_TYPE_NS) ? ls->GM.sD : ls->RI.sD;
    if (sDk < 1 || sDk + 1 > kArMax)
        return fail(SVFSI_ERR_ARG, "FSILS: Krylov space dimension must be 1.." + std::to_string(kArMax - 1));
  }
  if (dof < 1 || dof > 4) return fail(SVFSI_ERR_ARG, "FSILS: dof must be 1..4");
  if (prec != SVFSI_PRECOND_FSILS && prec != SVFSI_PRECOND_RCS)
    return fail(SVFSI_ERR_UNSUPPORTED,
                "FSILS: this linear solver and preconditioner combination is not supported");
  if (int rc = ensure_small()) return rc;

  ProfScope ps(PROF_SOLVE);
  // the solvers may grow (reallocate) the workspace, so W lives in its own buffer
  double *&d_W = g_dW;
  size_t &wCap = g_wCap;
  const size_t wNeed = (size_t)c.nNo * dof * sizeof(double);
  if (wCap < wNeed) {
    if (d_W) cudaFree(d_W);
    CUDA_TRY(cudaMalloc(&d_W, wNeed));
    wCap = wNeed;
  }
  c.valScaleW = nullptr;
  if (prec == SVFSI_PRECOND_FSILS) {
    // in-place GMRES on one rank with 4 x 4 blocks (the benchmark's path): the scaling of Val is fused into the
    // first product of the Arnoldi loop
    static int fuseScale = -1;
    if (fuseScale < 0) {
      const char *e = getenv("SVFSI_FUSE_SCALE");
      fuseScale = e ? atoi(e) : 1;
    }
    const bool fusedComm = c.nranks > 1 && !c.nbr.empty() && c.p2p.on && c.p2p.fuse;
    const bool defer = fuseScale && ls->LS_type == SVFSI_LS_TYPE_GMRES && dof == 4 && (c.nranks == 1 || fusedComm);
    if (int rc = preconddiag(dof, c.d_Val, c.d_R, d_W, defer)) return rc;
  } else {
    double *&d_rcs = g_dRcs;
    size_t &rcsCap = g_rcsCap;
    const size_t need = 3 * padded((size_t)c.nNo * dof);
    if (rcsCap < need) {
      if (d_rcs) cudaFree(d_rcs);
      CUDA_TRY(cudaMalloc(&d_rcs, need));
      rcsCap = need;
    }
    if (int rc = precondrcs(dof, c.d_Val, c.d_R, d_W, d_rcs)) return rc;
  }

  int rc = 0;
  switch (ls->LS_type) {
    case SVFSI_LS_TYPE_NS: rc = nssolver_dev(ls, dof, c.d_Val, c.d_R); break;
    case SVFSI_LS_TYPE_GMRES: rc = gmres_inplace(&ls->RI, dof, c.d_Val, c.d_R, dof == 1); break;
    case SVFSI_LS_TYPE_CG: rc = cgrad(&ls->RI, dof, c.d_Val, c.d_R); break;
    case SVFSI_LS_TYPE_BICGS: rc = bicgs(&ls->RI, dof, c.d_Val, c.d_R); break;
    default: rc = fail(SVFSI_ERR_UNSUPPORTED, "FSILS: LS_type not implemented on the device");
  }
  if (rc) return rc;
  if (int rc2 = flush_val_scale()) return rc2;   // a solve that returned before its first product
  launch_vecop(c.stream, VOP_MUL, c.d_R, d_W, nullptr, (size_t)c.nNo * dof, nullptr, 0.0, nullptr);
  return 0;
}

}  // namespace svfsi
