// bicgs_rcs.cu -- the two remaining arms of FSILS_SOLVE's dispatch table (L/SOLVE.f:98-131)
// on the device: BICGSS / BICGSV (L/BICGS.f:50-180) and the row/column-scaling preconditioner
// PRECONDRCS (L/PRECOND.f:150-368).
//
// BiCGStab keeps every scalar (alpha, omega, rho, beta, err) in the device control block; one
// iteration is 2 SpMVs, 3 fused reductions (each one pass producing one or two inner products)
// and 3 fused vector updates, with no host synchronisation (the host reads the stop flag of the
// previous iteration from mapped pinned memory, like the GMRES and CG drivers of solver.cu).
// Workspace vectors are laid out so that a two-column multi-dot yields <S,T>,<T,T> and
// <R,R>,<Rh,R> in one pass each.
#include <float.h>
#include <math.h>

#include "core.h"
#include "solver_int.h"

namespace svfsi {

// scal[0] = err, [1] = errO, [2] = alpha, [3] = omega, [4] = rho, [5] = beta
__global__ void bicgs_init_kernel(KrylovCtl *ctl, const double *ss, double absTol, double relTol) {
  const double err = sqrt(*ss);
  ctl->iNorm = err;
  const double e0 = relTol * err;
  ctl->eps = absTol > e0 ? absTol : e0;
  ctl->scal[0] = err;
  ctl->scal[1] = err;
  ctl->scal[4] = err * err;  // rho (L/BICGS.f:78)
  ctl->done = 0;
  ctl->suc = 0;
  ctl->ilast = 0;
  if (err < ctl->eps) {  // first trip of the loop exits with suc (:85-88)
    ctl->suc = 1;
    ctl->done = 1;
  }
}
// alpha = rho / <Rh, V>   (:90)
__global__ void bicgs_alpha_kernel(KrylovCtl *ctl, const double *rhv) {
  if (ctl->done) return;
  ctl->scal[2] = ctl->scal[4] / *rhv;
}
// omega = NORMV(T); omega = <T,S>/(omega*omega)   (:93-94); d[0] = <S,T>, d[1] = <T,T>
__global__ void bicgs_omega_kernel(KrylovCtl *ctl, const double *d) {
  if (ctl->done) return;
  const double nt = sqrt(d[1]);
  ctl->scal[3] = d[0] / (nt * nt);
}
// errO = err; err = NORMV(R); rhoO = rho; rho = <R,Rh>; beta = rho*alpha/(rhoO*omega)  (:97-101)
// and the test of the NEXT loop trip (:85-88); d[0] = <R,R>, d[1] = <Rh,R>
__global__ void bicgs_err_kernel(KrylovCtl *ctl, const double *d, int mItr, volatile int *pubFlag,
                                 volatile int *pubProgress, int seq) {
  if (!ctl->done) {
    ctl->scal[1] = ctl->scal[0];
    const double err = sqrt(d[0]);
    ctl->scal[0] = err;
    const double rhoO = ctl->scal[4];
    const double rho = d[1];
    ctl->scal[4] = rho;
    ctl->scal[5] = rho * ctl->scal[2] / (rhoO * ctl->scal[3]);
    ctl->ilast += 1;
    // the reference updates P once more before the test; P is dead after the exit
    if (ctl->ilast < mItr && err < ctl->eps) {
      ctl->suc = 1;
      ctl->done = 1;
    }
  }
  *pubFlag = (seq << 1) | (ctl->done ? 1 : 0);
}
// S = R - alpha V
__global__ void __launch_bounds__(256) bicgs_s_kernel(const KrylovCtl *ctl, double *__restrict__ S,
                                                      const double *__restrict__ R,
                                                      const double *__restrict__ V, size_t n) {
  if (ctl->done) return;
  const double alpha = ctl->scal[2];
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x)
    S[e] = R[e] - alpha * V[e];
}
// X = X + alpha P + omega S ; R = S - omega T
__global__ void __launch_bounds__(256) bicgs_xr_kernel(const KrylovCtl *ctl, double *__restrict__ X,
                                                       double *__restrict__ R,
                                                       const double *__restrict__ P,
                                                       const double *__restrict__ S,
                                                       const double *__restrict__ T, size_t n) {
  if (ctl->done) return;
  const double alpha = ctl->scal[2], omega = ctl->scal[3];
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    const double s = S[e];
    X[e] = (X[e] + alpha * P[e]) + omega * s;
    R[e] = s - omega * T[e];
  }
}
// P = R + beta (P - omega V)
__global__ void __launch_bounds__(256) bicgs_p_kernel(const KrylovCtl *ctl, double *__restrict__ P,
                                                      const double *__restrict__ R,
                                                      const double *__restrict__ V, size_t n) {
  if (ctl->done) return;
  const double omega = ctl->scal[3], beta = ctl->scal[5];
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x)
    P[e] = R[e] + beta * (P[e] - omega * V[e]);
}

// BICGSS / BICGSV (L/BICGS.f:50-180)
int bicgs(svfsi_subls_t *ls, int dof, const double *K, double *Rio) {
  Ctx &c = ctx();
  const size_t n = (size_t)c.nNo * dof, nOwned = (size_t)c.mynNo * dof;
  const size_t stride = padded(n) / sizeof(double);
  const int kind = dof == 1 ? 3 : 0;
  if (int rc = ensure_mirror()) return rc;
  if (int rc = ensure_small_n(4096)) return rc;
  KrylovCtl *ctl = (KrylovCtl *)c.d_small;
  double *sc = c.d_small + 32;
  if (int rc = ensure_ws(7 * stride * sizeof(double) + 4096)) return rc;
  // [R, Rh] and [S, T] are adjacent pairs: two-column multi-dots
  double *w = c.d_ws;
  double *R = w, *Rh = w + stride, *S = w + 2 * stride, *T = w + 3 * stride, *P = w + 4 * stride,
         *V = w + 5 * stride, *X = w + 6 * stride;
  const int *done = &ctl->done;
  const int grid = 148 * 8;

  const double t0 = now_s();
  launch_vecop(c.stream, VOP_COPY, R, Rio, nullptr, n, nullptr, 0.0, nullptr);
  if (int rc = dot_dev(R, R, nOwned, sc, nullptr)) return rc;
  bicgs_init_kernel<<<1, 1, 0, c.stream>>>(ctl, sc, ls->absTol, ls->relTol);
  count_launch();
  launch_vecop(c.stream, VOP_ZERO, X, nullptr, nullptr, n, nullptr, 0.0, nullptr);
  launch_vecop(c.stream, VOP_COPY, P, R, nullptr, n, nullptr, 0.0, nullptr);
  launch_vecop(c.stream, VOP_COPY, Rh, R, nullptr, n, nullptr, 0.0, nullptr);
  int seqPrev = publish(ctl);
  for (int i = 1; i <= ls->mItr; i++) {
    if (int rc = sparmul(kind, dof, K, P, V, done)) return rc;
    if (int rc = dot_dev(Rh, V, nOwned, sc, done)) return rc;
    bicgs_alpha_kernel<<<1, 1, 0, c.stream>>>(ctl, sc);
    {
      ProfScope ps(PROF_AXPY);
      bicgs_s_kernel<<<grid, 256, 0, c.stream>>>(ctl, S, R, V, n);
    }
    if (int rc = sparmul(kind, dof, K, S, T, done)) return rc;
    {
      ProfScope ps(PROF_DOT);
      launch_multidot(c.stream, S, stride, T, nOwned, 2, c.d_partial, done);
    }
    if (int rc = reduce_allreduce(c.d_partial, 2, sc + 2, done)) return rc;
    bicgs_omega_kernel<<<1, 1, 0, c.stream>>>(ctl, sc + 2);
    {
      ProfScope ps(PROF_AXPY);
      bicgs_xr_kernel<<<grid, 256, 0, c.stream>>>(ctl, X, R, P, S, T, n);
    }
    {
      ProfScope ps(PROF_DOT);
      launch_multidot(c.stream, R, stride, R, nOwned, 2, c.d_partial, done);
    }
    if (int rc = reduce_allreduce(c.d_partial, 2, sc + 4, done)) return rc;
    const int seq = ++g_seq;
    bicgs_err_kernel<<<1, 1, 0, c.stream>>>(ctl, sc + 4, ls->mItr, &g_hm_dev->flag[seq & 63],
                                            &g_hm_dev->progress, seq);
    {
      ProfScope ps(PROF_AXPY);
      bicgs_p_kernel<<<grid, 256, 0, c.stream>>>(ctl, P, R, V, n);
    }
    count_launch(7);
    int flag = 0;
    if (int rc = wait_flag(seqPrev, &flag)) return rc;
    if (flag) break;
    seqPrev = seq;
  }
  launch_vecop(c.stream, VOP_COPY, Rio, X, nullptr, n, nullptr, 0.0, nullptr);
  KrylovCtl hc;
  CUDA_TRY(cudaMemcpyAsync(&hc, ctl, sizeof(hc), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  ls->suc = hc.suc;
  ls->iNorm = hc.iNorm;
  ls->itr = hc.ilast;
  const double err = hc.scal[0], errO = hc.scal[1];
  ls->fNorm = err;
  ls->callD = now_s() - t0;
  if (errO < DBL_EPSILON) ls->dB = 0.0;
  else ls->dB = 10.0 * log(err / errO);
  return 0;
}

// ---------------------------------------------------------------------------
// PRECONDRCS kernels
__global__ void rcs_fill_kernel(size_t n, double v, double *__restrict__ a, double *__restrict__ b,
                                double *__restrict__ c) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  a[t] = v;
  if (b) b[t] = v;
  if (c) c[t] = v;
}
// the 0/1 renormalisation of the halo-summed Dirichlet mask (L/PRECOND.f:195-197)
__global__ void rcs_mask_kernel(size_t n, double *__restrict__ W) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double v = W[t] - 0.5;
  v = v / fabs(v);
  W[t] = (v + fabs(v)) * 0.5;
}
// unit diagonal on the killed rows (:203-237)
__global__ void rcs_diag_kernel(int nNo, int dof, const int *__restrict__ diag,
                                const double *__restrict__ W, double *__restrict__ Val) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nNo * dof) return;
  const int a = t / dof, i = t - a * dof;
  double *p = Val + (size_t)diag[a] * dof * dof + i * dof + i;
  *p = W[t] * (*p - 1.0) + 1.0;
}
// row / column max norms (:255-318).  |v| >= 0, so the IEEE bit pattern orders like the value:
// atomicMax on the 64-bit pattern is an exact, order-independent (deterministic) max.
__global__ void __launch_bounds__(256) rcs_max_kernel(int nnz, int dof, const int *__restrict__ rowOf,
                                                      const int *__restrict__ col,
                                                      const double *__restrict__ Val,
                                                      unsigned long long *__restrict__ Wr,
                                                      unsigned long long *__restrict__ Wc) {
  const int dd = dof * dof;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nnz * dd) return;
  const int p = (int)(t / dd), r = (int)(t - (size_t)p * dd);
  const int i = r / dof, k = r - i * dof;
  const unsigned long long v = (unsigned long long)__double_as_longlong(fabs(Val[t]));
  // one lane per row of the block would do for Wr; entries of a block row sit in adjacent
  // lanes, so fold them with shuffles when dof == 4 (4 consecutive lanes = one block row)
  atomicMax(Wr + (size_t)rowOf[p] * dof + i, v);
  atomicMax(Wc + (size_t)col[p] * dof + k, v);
}
// out[0] = max |1 - Wr|, out[1] = max |1 - Wc| over this rank (as bit patterns)
__global__ void __launch_bounds__(256) rcs_dev_kernel(size_t n, const double *__restrict__ Wr,
                                                      const double *__restrict__ Wc,
                                                      unsigned long long *__restrict__ out) {
  double mr = 0.0, mc = 0.0;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    mr = fmax(mr, fabs(1.0 - Wr[e]));
    mc = fmax(mc, fabs(1.0 - Wc[e]));
  }
  for (int o = 16; o > 0; o >>= 1) {
    mr = fmax(mr, __shfl_xor_sync(0xffffffffu, mr, o));
    mc = fmax(mc, __shfl_xor_sync(0xffffffffu, mc, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, (unsigned long long)__double_as_longlong(mr));
    atomicMax(out + 1, (unsigned long long)__double_as_longlong(mc));
  }
}
// Wr = 1/sqrt(Wr), Wc = 1/sqrt(Wc), W1 *= Wr, W2 *= Wc  (:326-333)
__global__ void rcs_invsqrt_kernel(size_t n, double *__restrict__ Wr, double *__restrict__ Wc,
                                   double *__restrict__ W1, double *__restrict__ W2) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double r = 1.0 / sqrt(Wr[t]), cc = 1.0 / sqrt(Wc[t]);
  Wr[t] = r;
  Wc[t] = cc;
  W1[t] = W1[t] * r;
  W2[t] = W2[t] * cc;
}

// PRECONDRCS (L/PRECOND.f:150-368).  W2 receives the solution scaling (Wc of L/SOLVE.f:136);
// work = 3 vectors of nNo*dof doubles.  One host synchronisation per sweep (<= 10 sweeps).
int precondrcs(int dof, double *Val, double *R, double *W2, double *work) {
  Ctx &c = ctx();
  ProfScope ps(PROF_PRECOND);
  const size_t n = (size_t)c.nNo * dof;
  const size_t stride = padded(n) / sizeof(double);
  double *Wr = work, *Wc = work + stride, *W1 = work + 2 * stride;
  const unsigned nb = (unsigned)((n + 255) / 256);
  if (int rc = ensure_small()) return rc;

  rcs_fill_kernel<<<nb, 256, 0, c.stream>>>(n, 1.0, Wr, W1, W2);
  count_launch();
  for (Face &f : c.face) {
    if (!f.created || !f.inc) continue;
    if (f.bGrp == SVFSI_BC_TYPE_DIR)
      launch_w_dirichlet(c.stream, f.nNo, f.dof, dof, f.d_glob, f.d_val, Wr);
  }
  if (int rc = halo_sum(Wr, dof, nullptr)) return rc;
  rcs_mask_kernel<<<nb, 256, 0, c.stream>>>(n, Wr);
  // PREMUL(Wr), R = Wr*R, POSMUL(Wr): the fused kernel multiplies (Val*W_row)*W_col per entry
  launch_scale_val2(c.stream, c.nnz, dof, c.d_rowOf, c.d_col, Wr, Wr, Val);
  launch_vecop(c.stream, VOP_MUL, R, Wr, nullptr, n, nullptr, 0.0, nullptr);
  rcs_diag_kernel<<<(unsigned)((c.nNo * dof + 255) / 256), 256, 0, c.stream>>>(c.nNo, dof, c.d_diag,
                                                                             Wr, Val);
  count_launch(2);

  const int maxiter = 10;
  const double tol = 2.0;
  bool flag = true;
  int iter = 0;
  unsigned long long *dmax = (unsigned long long *)(c.d_small + 64);
  double *dflag = c.d_small + 66;
  while (flag) {
    iter++;
    if (iter >= maxiter) flag = false;
    CUDA_TRY(cudaMemsetAsync(Wr, 0, n * sizeof(double), c.stream));
    CUDA_TRY(cudaMemsetAsync(Wc, 0, n * sizeof(double), c.stream));
    CUDA_TRY(cudaMemsetAsync(dmax, 0, 2 * sizeof(unsigned long long), c.stream));
    const size_t tot = (size_t)c.nnz * dof * dof;
    rcs_max_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c.stream>>>(
        c.nnz, dof, c.d_rowOf, c.d_col, Val, (unsigned long long *)Wr, (unsigned long long *)Wc);
    count_launch();
    if (int rc = halo_sum(Wr, dof, nullptr)) return rc;   // sums, as the reference does (:320-321)
    if (int rc = halo_sum(Wc, dof, nullptr)) return rc;
    rcs_dev_kernel<<<148 * 4, 256, 0, c.stream>>>(n, Wr, Wc, dmax);
    count_launch();
    CUDA_TRY(cudaMemcpyAsync(c.h_small, dmax, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    const bool conv = (c.h_small[0] < tol) && (c.h_small[1] < tol);
    if (conv) flag = false;
    rcs_invsqrt_kernel<<<nb, 256, 0, c.stream>>>(n, Wr, Wc, W1, W2);
    count_launch();
    launch_scale_val2(c.stream, c.nnz, dof, c.d_rowOf, c.d_col, Wr, Wc, Val);
    if (c.nranks > 1) {  // MPI_ALLGATHER(flag) + ANY (:338-342)
      c.h_small[0] = flag ? 1.0 : 0.0;
      CUDA_TRY(cudaMemcpyAsync(dflag, c.h_small, sizeof(double), cudaMemcpyHostToDevice, c.stream));
      if (int rc = allreduce_dev(dflag, 1)) return rc;
      CUDA_TRY(cudaMemcpyAsync(c.h_small, dflag, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      CUDA_TRY(cudaStreamSynchronize(c.stream));
      flag = c.h_small[0] > 0.0;
    }
  }
  launch_vecop(c.stream, VOP_MUL, R, W1, nullptr, n, nullptr, 0.0, nullptr);
  return 0;
}

}  // namespace svfsi
