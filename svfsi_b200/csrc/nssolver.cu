// nssolver.cu -- NSSOLVER ("BIPN", svFSI's default linear solver for fluid) on the device:
// L/NSSOLVER.f:52-233 with DEPART :237-305, BCPRE :307-341 and the dense GE solve (L/GE.f:51-151).
//
// Everything that touches nNo- or nnz-sized data runs in CUDA kernels (block split, the four
// SpMV shapes, the inner GMRES / Schur-CG, the Gram matrix by fused multi-dots, the basis
// recombinations by fused multi-axpys); the <= 20 x 20 normal-equation solve is host code, once
// per outer iteration, exactly as the reference does it.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "solver_int.h"

namespace svfsi {

namespace {

// L/GE.f:51-151: Gauss elimination with diagonal pre-scaling and partial pivoting.
// A(nV,N) column-major, B(N) in/out.  Returns true on success.
bool ge(int nV, int N, const double *A, double *B) {
  auto AA = [&](int i, int j) { return A[(size_t)(j - 1) * nV + (i - 1)]; };
  if (N <= 0) return false;
  std::vector<double> W(N);
  for (int i = 1; i <= N; i++) {
    if (fabs(AA(i, i)) < 2.2250738585072014e-308) {  // TINY(A)
      for (int j = 0; j < N; j++) B[j] = 0.0;
      return false;
    }
    W[i - 1] = 1.0 / sqrt(fabs(AA(i, i)));
  }
  std::vector<double> Cm((size_t)N * (N + 1));
  auto C = [&](int i, int j) -> double & { return Cm[(size_t)(j - 1) * N + (i - 1)]; };
  for (int i = 1; i <= N; i++) {
    for (int j = 1; j <= N; j++) C(i, j) = W[i - 1] * W[j - 1] * AA(i, j);
    C(i, N + 1) = W[i - 1] * B[i - 1];
  }
  const double eps = 2.220446049250313e-16;
  if (N == 1) {
    B[0] = C(1, 2) / C(1, 1);
    B[0] = B[0] * W[0];
    return true;
  } else if (N == 2) {
    const double pivot = C(1, 1) * C(2, 2) - C(2, 1) * C(1, 2);
    if (fabs(pivot) < eps) {
      B[0] = B[1] = 0.0;
      return false;
    }
    B[0] = (C(1, 3) * C(2, 2) - C(2, 3) * C(1, 2)) / pivot;
    B[1] = (C(2, 3) * C(1, 1) - C(1, 3) * C(2, 1)) / pivot;
    B[0] = W[0] * B[0];
    B[1] = W[1] * B[1];
    return true;
  }
  for (int m = 1; m <= N - 1; m++) {
    int ipv = m;
    double pivot = fabs(C(m, m));
    for (int i = m + 1; i <= N; i++)
      if (fabs(C(i, m)) > pivot) { ipv = i; pivot = fabs(C(i, m)); }
    if (fabs(pivot) < eps) {
      for (int j = 0; j < N; j++) B[j] = 0.0;
      return false;
    }
    if (ipv != m)
      for (int j = m; j <= N + 1; j++) std::swap(C(m, j), C(ipv, j));
    for (int i = m + 1; i <= N; i++) {
      const double saveEl = C(i, m) / C(m, m);
      C(i, m) = 0.0;
      for (int j = m + 1; j <= N + 1; j++) C(i, j) = C(i, j) - saveEl * C(m, j);
    }
  }
  for (int j = N; j >= 1; j--) {
    for (int i = j + 1; i <= N; i++) C(j, N + 1) = C(j, N + 1) - C(j, i) * C(i, N + 1);
    C(j, N + 1) = C(j, N + 1) / C(j, j);
  }
  for (int i = 1; i <= N; i++) B[i - 1] = W[i - 1] * C(i, N + 1);
  return true;
}

// tpos[p] = block position of the transposed entry (col(p), row(p)); -1 if absent
__global__ void tpos_kernel(int nnz, const int *__restrict__ rowOf, const int *__restrict__ col,
                            const int *__restrict__ rowPtr, int *__restrict__ tpos) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const int i = rowOf[p], k = col[p];
  int l = -1;
  for (int j = rowPtr[k]; j < rowPtr[k + 1]; j++)
    if (col[j] == i) { l = j; break; }
  tpos[p] = l;
}

// out[j] = a[j] + b[j]
__global__ void add_small_kernel(int n, const double *a, const double *b, double *out) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = a[j] + b[j];
}

int *g_tpos = nullptr;
int g_tpos_gen = -1;

}  // namespace

void nssolver_free_static() {
  if (g_tpos) cudaFree(g_tpos);
  g_tpos = nullptr;
  g_tpos_gen = -1;
}

int nssolver_dev(svfsi_ls_t *ls, int dof, const double *Val, double *Ri) {
  Ctx &c = ctx();
  const int nsd = dof - 1;
  if (nsd != 2 && nsd != 3) return fail(SVFSI_ERR_ARG, "FSILS: Not defined nsd for DEPART");
  const int mItr = ls->RI.mItr;
  const int nB = 2 * mItr;
  const size_t n = (size_t)c.nNo, nOwned = (size_t)c.mynNo;
  const size_t sV = padded(n * nsd) / sizeof(double), sS = padded(n) / sizeof(double);
  const size_t nnz = (size_t)c.nnz;
  if (int rc = ensure_mirror()) return rc;

  // ---- scalar area: [inner GMRES / CG block | NS block]
  GmresScal gs = gmres_scal(nullptr, ls->GM.sD, (int)c.face.size());
  const size_t innerScal = gs.doubles > 4096 ? gs.doubles : 4096;
  const size_t nsScal = (size_t)4 * (nB + 2) * (nB + 2) + 64;
  if (int rc = ensure_small_n(innerScal + nsScal)) return rc;
  double *scalInner = c.d_small;
  double *scalNS = c.d_small + innerScal;
  double *dV = scalNS;                         // NCDOTV parts  [nB+1]
  double *dS = scalNS + (nB + 2);              // NCDOTS parts
  double *dT = scalNS + 2 * (nB + 2);          // tmp(c) of one k: sums, all-reduced
  double *dCoef = scalNS + 3 * (nB + 2);       // xB uploaded for the recombinations
  double *faceS = scalNS + 4 * (nB + 2);

  // ---- workspace
  const size_t wInner = std::max(gmres_out_work(ls->GM.sD, n * nsd), cg_schur_work(n, nsd));
  const size_t total = 2 * sV + 2 * sS                 // Rm, Rmi, Rc, Rci
                       + (size_t)mItr * (sV + sS)      // U, P
                       + (size_t)nB * (sV + sS)        // MU, MP
                       + padded(nnz * nsd * nsd) / 8 + 3 * (padded(nnz * nsd) / 8) + padded(nnz) / 8
                       + wInner + 1024;
  if (int rc = ensure_ws(total * sizeof(double))) return rc;
  Bump b(c.d_ws);
  double *Rm = b.take(n * nsd), *Rmi = b.take(n * nsd), *Rc = b.take(n), *Rci = b.take(n);
  double *U = b.take((size_t)mItr * sV), *P = b.take((size_t)mItr * sS);
  double *MU = b.take((size_t)nB * sV), *MP = b.take((size_t)nB * sS);
  double *mK = b.take(nnz * nsd * nsd), *mG = b.take(nnz * nsd), *mD = b.take(nnz * nsd);
  double *Gt = b.take(nnz * nsd), *mL = b.take(nnz);
  double *wIn = b.take(wInner);

  // transposed-position map: the pattern is fixed, build once per lhs (the reference searches
  // every call, L/NSSOLVER.f:292-302)
  if (g_tpos_gen != c.lhsGen) {
    if (g_tpos) cudaFree(g_tpos);
    CUDA_TRY(cudaMalloc(&g_tpos, sizeof(int) * nnz));
    tpos_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, c.stream>>>(c.nnz, c.d_rowOf, c.d_col,
                                                                     c.d_rowPtr, g_tpos);
    count_launch();
    g_tpos_gen = c.lhsGen;
  }

  std::vector<double> tmp((size_t)nB * nB + nB, 0.0), A((size_t)nB * nB, 0.0), B(nB, 0.0),
      xB(nB, 0.0), oldxB(nB, 0.0);
  auto AM = [&](int i, int j) -> double & { return A[(size_t)(j - 1) * nB + (i - 1)]; };

  launch_split_mc(c.stream, c.nNo, dof, Ri, Rmi, Rci);
  launch_vecop(c.stream, VOP_COPY, Rm, Rmi, nullptr, n * nsd, nullptr, 0.0, nullptr);
  launch_vecop(c.stream, VOP_COPY, Rc, Rci, nullptr, n, nullptr, 0.0, nullptr);
  if (int rc = dot_dev(Rm, Rm, nOwned * nsd, dV, nullptr)) return rc;
  if (int rc = dot_dev(Rc, Rc, nOwned, dV + 1, nullptr)) return rc;
  CUDA_TRY(cudaMemcpyAsync(c.h_small, dV, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  double eps;
  {
    double nm = sqrt(c.h_small[0]), nc = sqrt(c.h_small[1]);  // NORMV / NORMS
    eps = sqrt(nm * nm + nc * nc);
  }
  ls->RI.iNorm = eps;
  ls->RI.fNorm = eps * eps;
  ls->CG.callD = 0.0;
  ls->GM.callD = 0.0;
  ls->CG.itr = 0;
  ls->GM.itr = 0;
  const double t0 = now_s();
  ls->RI.suc = 0;
  eps = std::max(ls->RI.absTol, ls->RI.relTol * eps);

  // DEPART
  launch_depart(c.stream, c.nnz, nsd, Val, mK, mG, mD, mL);
  CUDA_TRY(cudaMemsetAsync(Gt, 0, sizeof(double) * nnz * nsd, c.stream));
  launch_gt(c.stream, c.nnz, nsd, g_tpos, mG, Gt);
  if (int rc = bcpre(nsd, faceS)) return rc;

  int i, iB = 0, iBB = 0;
  for (i = 1; i <= mItr; i++) {
    iB = 2 * i - 1;
    iBB = 2 * i;
    ls->RI.dB = ls->RI.fNorm;
    double *Ui = U + (size_t)(i - 1) * sV, *Pi = P + (size_t)(i - 1) * sS;
    double *MU1 = MU + (size_t)(iB - 1) * sV, *MU2 = MU + (size_t)(iBB - 1) * sV;
    double *MP1 = MP + (size_t)(iB - 1) * sS, *MP2 = MP + (size_t)(iBB - 1) * sS;

    // U = K^-1 Rm
    if (int rc = gmres_outofplace(&ls->GM, nsd, mK, Rm, Ui, wIn, scalInner)) return rc;
    // P = Rc - D U
    if (int rc = sparmul(1, nsd, mD, Ui, Pi, nullptr)) return rc;
    launch_vecop(c.stream, VOP_SUB_FROM, Pi, Rc, nullptr, n, nullptr, 0.0, nullptr);
    // P = [L + G^T G]^-1 P
    if (int rc = cgrad_schur(&ls->CG, nsd, Gt, mG, mL, Pi, wIn, scalInner)) return rc;
    // MU1 = G P ; MU2 = Rm - G P
    if (int rc = sparmul(2, nsd, mG, Pi, MU1, nullptr)) return rc;
    launch_vecop(c.stream, VOP_SUB, MU2, Rm, MU1, n * nsd, nullptr, 0.0, nullptr);
    // U = K^-1 [Rm - G P]
    if (int rc = gmres_outofplace(&ls->GM, nsd, mK, MU2, Ui, wIn, scalInner)) return rc;
    // MU2 = K U (+ coupled BC) ; MP1 = L P ; MP2 = D U
    if (int rc = sparmul(0, nsd, mK, Ui, MU2, nullptr)) return rc;
    if (int rc = addbcmul(0, nsd, Ui, MU2, faceS, nullptr)) return rc;
    if (int rc = sparmul(3, 1, mL, Pi, MP1, nullptr)) return rc;
    if (int rc = sparmul(1, nsd, mD, Ui, MP2, nullptr)) return rc;

    // Gram matrix rows iB, iBB (L/NSSOLVER.f:136-151): NCDOTV + NCDOTS, one all-reduce
    int cnt = 0;
    for (int k = iB; k <= iBB; k++) {
      const double *MUk = MU + (size_t)(k - 1) * sV, *MPk = MP + (size_t)(k - 1) * sS;
      {
        ProfScope ps(PROF_DOT);
        launch_multidot(c.stream, MU, sV, MUk, nOwned * nsd, k, c.d_partial, nullptr);
        launch_reduce_partials(c.stream, c.d_partial, k, dV, nullptr);
        launch_multidot(c.stream, Rmi, 0, MUk, nOwned * nsd, 1, c.d_partial, nullptr);
        launch_reduce_partials(c.stream, c.d_partial, 1, dV + k, nullptr);
        launch_multidot(c.stream, MP, sS, MPk, nOwned, k, c.d_partial, nullptr);
        launch_reduce_partials(c.stream, c.d_partial, k, dS, nullptr);
        launch_multidot(c.stream, Rci, 0, MPk, nOwned, 1, c.d_partial, nullptr);
        launch_reduce_partials(c.stream, c.d_partial, 1, dS + k, nullptr);
        add_small_kernel<<<1, 64, 0, c.stream>>>(k + 1, dV, dS, dT);
        count_launch();
      }
      if (int rc = allreduce_dev(dT, (size_t)k + 1)) return rc;
      CUDA_TRY(cudaMemcpyAsync(c.h_small, dT, sizeof(double) * (k + 1), cudaMemcpyDeviceToHost,
                               c.stream));
      CUDA_TRY(cudaStreamSynchronize(c.stream));
      for (int j = 0; j <= k; j++) tmp[cnt++] = c.h_small[j];
    }
    cnt = 0;
    for (int k = iB; k <= iBB; k++) {
      for (int j = 1; j <= k; j++) {
        AM(j, k) = tmp[cnt];
        AM(k, j) = tmp[cnt];
        cnt++;
      }
      B[k - 1] = tmp[cnt++];
    }
    xB = B;
    if (ge(nB, iBB, A.data(), xB.data())) {
      oldxB = xB;
    } else {
      if (c.rank == 0) fprintf(stderr, " FSILS: Singular matrix detected\n");
      xB = oldxB;
      if (i > 1) {
        iB -= 2;
        iBB -= 2;
      }
      break;
    }
    double sum = 0.0;
    for (int j = 0; j < iBB; j++) sum += xB[j] * B[j];
    ls->RI.fNorm = ls->RI.iNorm * ls->RI.iNorm - sum;
    if (ls->RI.fNorm < eps * eps) {
      ls->RI.suc = 1;
      break;
    }
    // Rm = Rmi - sum_j xB(j) MU_j ; Rc = Rci - sum_j xB(j) MP_j
    CUDA_TRY(cudaMemcpyAsync(dCoef, xB.data(), sizeof(double) * iBB, cudaMemcpyHostToDevice, c.stream));
    launch_vecop(c.stream, VOP_COPY, Rm, Rmi, nullptr, n * nsd, nullptr, 0.0, nullptr);
    launch_vecop(c.stream, VOP_COPY, Rc, Rci, nullptr, n, nullptr, 0.0, nullptr);
    launch_multi_axpy_scale(c.stream, MU, sV, Rm, n * nsd, iBB, dCoef, nullptr, nullptr);
    launch_multi_axpy_scale(c.stream, MP, sS, Rc, n, iBB, dCoef, nullptr, nullptr);
    CUDA_TRY(cudaStreamSynchronize(c.stream));  // xB (host vector) is re-used next iteration
  }
  if (i > mItr) {
    ls->RI.itr = mItr;
  } else {
    ls->RI.itr = i;
    CUDA_TRY(cudaMemcpyAsync(dCoef, xB.data(), sizeof(double) * nB, cudaMemcpyHostToDevice, c.stream));
    launch_vecop(c.stream, VOP_COPY, Rc, Rci, nullptr, n, nullptr, 0.0, nullptr);
    launch_multi_axpy_scale(c.stream, MP, sS, Rc, n, iBB, dCoef, nullptr, nullptr);
  }
  if (int rc = dot_dev(Rc, Rc, nOwned, dV, nullptr)) return rc;
  CUDA_TRY(cudaMemcpyAsync(c.h_small, dV, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  {
    double nc = sqrt(c.h_small[0]);
    ls->Resc = (int)lround(100.0 * (nc * nc) / ls->RI.fNorm);
    ls->Resm = 100 - ls->Resc;
  }
  // Rmi = xB(2) U_1 + sum_{i>=2} xB(2i) U_i ; Rci = xB(1) P_1 + sum xB(2i-1) P_i
  {
    std::vector<double> cu(mItr, 0.0), cp(mItr, 0.0);
    for (int k = 1; k <= ls->RI.itr; k++) {
      cu[k - 1] = -xB[2 * k - 1];
      cp[k - 1] = -xB[2 * k - 2];
    }
    CUDA_TRY(cudaMemcpyAsync(dCoef, cu.data(), sizeof(double) * mItr, cudaMemcpyHostToDevice, c.stream));
    CUDA_TRY(cudaMemcpyAsync(dCoef + mItr, cp.data(), sizeof(double) * mItr, cudaMemcpyHostToDevice,
                             c.stream));
    launch_vecop(c.stream, VOP_ZERO, Rmi, nullptr, nullptr, n * nsd, nullptr, 0.0, nullptr);
    launch_vecop(c.stream, VOP_ZERO, Rci, nullptr, nullptr, n, nullptr, 0.0, nullptr);
    launch_multi_axpy_scale(c.stream, U, sV, Rmi, n * nsd, ls->RI.itr, dCoef, nullptr, nullptr);
    launch_multi_axpy_scale(c.stream, P, sS, Rci, n, ls->RI.itr, dCoef + mItr, nullptr, nullptr);
    CUDA_TRY(cudaStreamSynchronize(c.stream));
  }
  ls->RI.callD = now_s() - t0;
  ls->RI.dB = 5.0 * log(ls->RI.fNorm / ls->RI.dB);
  if (ls->Resc < 0 || ls->Resm < 0) {
    ls->Resc = 0;
    ls->Resm = 0;
    ls->RI.dB = 0;
    ls->RI.fNorm = 0.0;
    if (c.rank == 0)
      fprintf(stderr, "Warning: unexpected behavior in FSILS (likely due to the ill-conditioned LHS matrix)\n");
  }
  ls->RI.fNorm = sqrt(ls->RI.fNorm);
  launch_join_mc(c.stream, c.nNo, dof, Rmi, Rci, Ri);
  return 0;
}

}  // namespace svfsi
