// asm_kernels.cu -- element loops of svFSI on sm_100a: CONSTRUCT_FLUID
// (S/FLUID.f:40-190 with FLUID3D_M :192-560 and FLUID3D_C :813-1084) and
// CONSTRUCT_HEATS (S/HEATS.f:39-159) for linear tetrahedra, fused with the
// DOASSEM scatter (S/LHSA.f:266-298) into the dof x dof block-CSR matrix.
//
// B200 design (not a transcription of the Fortran loop):
//  * For TET4 the shape-function gradients, metric tensor, velocity gradient and
//    every product of them are element constants; only N_a(g)-weighted
//    quantities change between the four Gauss points (S/NN.f:268-275,654-671).
//    The kernel therefore reduces the four Gauss points to ~40 per-element
//    scalars (phase 1, one thread per element) and expands them into the sixteen
//    4x4 tangent blocks only at scatter time (phase 2).  Identically-zero terms
//    of the reference (second derivatives, viscosity gradient)
//    are dropped.  ~2 kflop/element instead of ~15 kflop: the kernel is bound by
//    the 2 KB/element scatter, not by FP64 issue.
//  * Phase 2 is warp-cooperative: 8 consecutive lanes own one 128-byte tangent
//    block (16 bytes per lane), so every warp-level store/atomic instruction
//    covers four whole 128-byte lines of Val.
//  * Two scatter variants: FP64 atomics (RED.ADD.F64, any element order) and a
//    plain read-modify-write used with the greedy element colouring (no two
//    elements of a colour share a node: deterministic).
#include <cuda_runtime.h>
#include <float.h>
#include <stdlib.h>

#include "asm_elem.h"
#include "ctx.h"
#include "kernels.h"

namespace svfsi {

// The per-element compact record: 80 doubles (640 B).
//   node record a (8 doubles at a*8): Nx(1..3,a), C2_a, sum tauC, wl, sum tauM, R2_a
//       C2_a = rho sum_g tauM uaNx_a ; R2_a = rho sum_g tauM (uNx_a + amd N_a)
//       (the three element-wide scalars are replicated in every node record so that one
//        64-byte node record holds everything a tangent block needs from that node)
//   (D,E)(a,b) at 32 + (a*4+b)*2 : D = 4 mu Nx_a.Nx_b + A_ab (diagonal term of the momentum
//       tangent), E = sum tauM * Nx_a.Nx_b (dC/dP entry)
//   lR(i,a) at 64 + a*4 + i
// In shared memory (scatter variants) field f of slot s lives at f*NEP + s; in global memory
// (gather variant) at e*80 + f.
enum {
  N_NX = 0, N_C2 = 3, N_STC = 4, N_WL = 5, N_STM = 6, N_R2 = 7,
  F_DE = 32, F_LR = 64, F_COUNT = 80
};

struct Tet4Tab {
  double N[4][4];  // N[g][a]
  double sN[4];    // sum_g N[g][a]
};

__device__ __forceinline__ void tet4_tab(Tet4Tab &t) {
  // S/NN.f:268-275 (GETGIP) and :654-658 (GETGNN); N4 = 1 - xi1 - xi2 - xi3 as computed
  const double s = (5.0 + 3.0 * sqrt(5.0)) / 20.0;
  const double q = (5.0 - sqrt(5.0)) / 20.0;
#pragma unroll
  for (int g = 0; g < 4; g++) {
    const double x0 = (g == 0) ? s : q, x1 = (g == 1) ? s : q, x2 = (g == 2) ? s : q;
    t.N[g][0] = x0;
    t.N[g][1] = x1;
    t.N[g][2] = x2;
    t.N[g][3] = 1.0 - x0 - x1 - x2;
  }
#pragma unroll
  for (int a = 0; a < 4; a++) t.sN[a] = t.N[0][a] + t.N[1][a] + t.N[2][a] + t.N[3][a];
}

__device__ __forceinline__ void red_add(double *p, double v) { atomicAdd(p, v); }

// Fold the sums into the compact record.  PH < 0: all 80 fields, field f at rec[f*NEP];
// PH = 0/1: only fields [40 PH, 40 PH + 40), field f at rec[(f - 40 PH)*NEP] (two-pass staging
// through half the shared memory).  f is a compile-time constant at every store.
template <int PH>
__device__ __forceinline__ void rec_put(double *rec, int f, double v) {
  if (PH < 0) rec[f * NEP] = v;
  else if (f / 40 == PH) rec[(f - 40 * PH) * NEP] = v;
}
template <int PH>
__device__ __forceinline__ void fluid_elem_store(const FluidPar &par, const ElemAcc &acc,
                                                 double *rec) {
  const double rho = par.rho, mu = par.mu;
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
      for (int i = 0; i < 3; i++) rec_put<PH>(rec, a * 8 + N_NX + i, acc.Nx[a][i]);
      rec_put<PH>(rec, a * 8 + N_C2, rho * acc.c2[a]);
      rec_put<PH>(rec, a * 8 + N_STC, acc.sTC);
      rec_put<PH>(rec, a * 8 + N_WL, acc.wl);
      rec_put<PH>(rec, a * 8 + N_STM, acc.sTM);
      rec_put<PH>(rec, a * 8 + N_R2, rho * acc.r2[a]);
#pragma unroll
      for (int b = 0; b < 4; b++) {
        const double nn = acc.Nx[a][0] * acc.Nx[b][0] + acc.Nx[a][1] * acc.Nx[b][1] +
                          acc.Nx[a][2] * acc.Nx[b][2];
        rec_put<PH>(rec, F_DE + (a * 4 + b) * 2, 4.0 * (mu * nn) + acc.A[a][b]);
        rec_put<PH>(rec, F_DE + (a * 4 + b) * 2 + 1, acc.sTM * nn);
      }
#pragma unroll
      for (int i = 0; i < 3; i++)
        rec_put<PH>(rec, F_LR + a * 4 + i,
                    acc.wr * acc.sNrV[a][i] +
                        acc.w * (acc.Nx[a][0] * acc.sRM[0][i] + acc.Nx[a][1] * acc.sRM[1][i] +
                                 acc.Nx[a][2] * acc.sRM[2][i]));
      rec_put<PH>(rec, F_LR + a * 4 + 3, acc.w * acc.lR4[a]);
    }
}

__device__ __forceinline__ void fluid_elem_record(const FluidPar &par, int e,
                                                  const int *__restrict__ ien,
                                                  const double *__restrict__ x,
                                                  const double *__restrict__ Ag,
                                                  const double *__restrict__ Yg,
                                                  const double *__restrict__ Bf, double *rec,
                                                  int *nodeOut, int *__restrict__ badJac) {
  ElemAcc acc;
  fluid_elem_compute(par, e, ien, x, Ag, Yg, Bf, acc, nodeOut, badJac);
  fluid_elem_store<-1>(par, acc, rec);
}

// The pair of tangent entries (row i = q>>1, columns 2(q&1), 2(q&1)+1) of block (a,b) that lane q
// of an 8-lane group owns, expanded from the compact record (field f at rec[f*STRIDE]).
// S/FLUID.f:482-557 (momentum rows) and :1052-1081 (continuity row), Gauss sums pre-reduced.
// Branch-free: every lane issues the same six loads (two of them 128-bit in the global layout)
// and selects, so a warp never serialises over the four (row kind, column pair) cases.
template <int STRIDE>
__device__ __forceinline__ double2 ld2(const double *p) {
  if (STRIDE == 1) return __ldg((const double2 *)p);
  return make_double2(p[0], p[STRIDE]);
}
__device__ __forceinline__ double sN_of(int a) {
  // sum_g N(a,g) with N from S/NN.f:268-275,654-658 (mathematically 1; kept as computed)
  const double s = (5.0 + 3.0 * sqrt(5.0)) / 20.0, t = (5.0 - sqrt(5.0)) / 20.0;
  const double s012 = ((s + t) + t) + t, s1 = ((t + s) + t) + t, s2 = ((t + t) + s) + t;
  const double n30 = 1.0 - s - t - t, n31 = 1.0 - t - s - t, n32 = 1.0 - t - t - s, n33 = 1.0 - t - t - t;
  const double s3 = ((n30 + n31) + n32) + n33;
  return a == 0 ? s012 : (a == 1 ? s1 : (a == 2 ? s2 : s3));
}
template <int STRIDE>
__device__ __forceinline__ void tangent_pair(const double *__restrict__ rec, int a, int b, int q,
                                             double mu4, double &v0, double &v1) {
  const int i = q >> 1, j0 = (q & 1) * 2;
  const bool row3 = (i == 3), col2 = (j0 == 2);
  const double *ra = rec + (size_t)(a * 8) * STRIDE, *rb = rec + (size_t)(b * 8) * STRIDE;
  const double2 A = ld2<STRIDE>(ra + (size_t)j0 * STRIDE);   // (Nx_j0, Nx_j0+1) of a; j0=2: (Nx_2, C2_a)
  const double2 B = ld2<STRIDE>(rb + (size_t)j0 * STRIDE);
  const double2 S = ld2<STRIDE>(ra + (size_t)N_STC * STRIDE);  // (sum tauC, wl)
  const double ai = ra[(size_t)i * STRIDE];                   // Nx_i of a (row 3: unused)
  const double bi = rb[(size_t)(row3 ? N_R2 : i) * STRIDE];    // Nx_i of b, or R2_b on the continuity row
  const double de = rec[(size_t)(F_DE + (a * 4 + b) * 2 + (row3 ? 1 : 0)) * STRIDE];
  const double wl = S.y, sTC = S.x;
  const double sNa = sN_of(a), sNb = sN_of(b);
  // momentum rows, velocity columns
  const double m0 = (mu4 * (A.x * bi) + sTC * (ai * B.x) + ((i == j0) ? de : 0.0)) * wl;
  const double m1 = (mu4 * (A.y * bi) + sTC * (ai * B.y) + ((i == 1) ? de : 0.0)) * wl;
  const double mp = -wl * (ai * sNb - bi * A.y);               // pressure column (A.y = C2_a)
  // continuity row
  const double c0 = wl * (sNa * B.x + A.x * bi);               // bi = R2_b
  const double c1 = wl * (sNa * B.y + A.y * bi);
  const double cp = wl * de;                                   // de = sum tauM Nx_a.Nx_b
  v0 = row3 ? c0 : m0;
  v1 = row3 ? (col2 ? cp : c1) : (col2 ? mp : m1);
}

// ---------------------------------------------------------------------------
// Scatter variants (atomic / coloured): phases 1 and 2 in one CTA through shared memory.
template <bool ATOMIC>
__global__ void __launch_bounds__(NE) fluid_asm_kernel(FluidPar par, int n, int e0,
                                                       const int *__restrict__ elems,
                                                       const int *__restrict__ ien,
                                                       const int *__restrict__ edest,
                                                       const double *__restrict__ x,
                                                       const double *__restrict__ Ag,
                                                       const double *__restrict__ Yg,
                                                       const double *__restrict__ Bf,
                                                       double *__restrict__ R,
                                                       double *__restrict__ Val,
                                                       int *__restrict__ badJac) {
  extern __shared__ double sm[];      // [F_COUNT][NEP]
  __shared__ int sEl[NE];             // element id per slot (-1 = none)
  __shared__ int sNode[NE * 4];

  const int slot = threadIdx.x;
  const int idx = blockIdx.x * NE + slot;
  int e = -1;
  if (idx < n) e = elems ? elems[e0 + idx] : e0 + idx;
  sEl[slot] = e;
  if (e >= 0) fluid_elem_record(par, e, ien, x, Ag, Yg, Bf, sm + slot, sNode + slot * 4, badJac);
  __syncthreads();

  // ---------------- phase 2a: residual scatter, R(:,Ac) += lR(:,a) ----------------
  for (int it = threadIdx.x; it < NE * 16; it += NE) {
    const int s = it >> 4, r = it & 15;
    if (sEl[s] < 0) continue;
    const double v = sm[(F_LR + r) * NEP + s];
    double *dst = R + (size_t)sNode[s * 4 + (r >> 2)] * 4 + (r & 3);
    if (ATOMIC) red_add(dst, v);
    else *dst += v;
  }

  // ---------------- phase 2b: tangent scatter, 8 lanes per 4x4 block ----------------
  const double mu4 = 4.0 * par.mu;
  for (int it = threadIdx.x; it < NE * 128; it += NE) {
    const int s = it >> 7;
    const int el = sEl[s];
    if (el < 0) continue;
    const int blk = (it >> 3) & 15, q = it & 7;
    double v0, v1;
    tangent_pair<NEP>(sm + s, blk >> 2, blk & 3, q, mu4, v0, v1);
    const int p = __ldg(edest + (size_t)el * 16 + blk);
    double *dst = Val + (size_t)p * 16 + q * 2;
    if (ATOMIC) {
      red_add(dst, v0);
      red_add(dst + 1, v1);
    } else {
      double2 cur = *(double2 *)dst;
      cur.x += v0;
      cur.y += v1;
      *(double2 *)dst = cur;
    }
  }
}

// ---------------------------------------------------------------------------
// Gather variant ("owner computes", deterministic, no atomics, no colouring):
//   kernel A: compact records of all elements -> elemP[e][64] (coalesced through smem)
//   kernel B: 8 lanes per Val block sum the contributions of the elements around that
//             (row,col) edge in ascending element order -- the reference's DOASSEM
//             accumulation order (S/LHSA.f:275-295) -- and write the block ONCE.
//   kernel C: same for the residual, one thread per (node, component).
__global__ void __launch_bounds__(NE) fluid_record_kernel(FluidPar par, int n,
                                                          const int *__restrict__ ien,
                                                          const double *__restrict__ x,
                                                          const double *__restrict__ Ag,
                                                          const double *__restrict__ Yg,
                                                          const double *__restrict__ Bf,
                                                          double *__restrict__ elemP,
                                                          int *__restrict__ badJac) {
  extern __shared__ double sm[];  // [F_COUNT][NEP]
  const int slot = threadIdx.x;
  const int e = blockIdx.x * NE + slot;
  int nodes[4];
  if (e < n) fluid_elem_record(par, e, ien, x, Ag, Yg, Bf, sm + slot, nodes, badJac);
  __syncthreads();
  const int nHere = min(NE, n - blockIdx.x * NE);
  double *out = elemP + (size_t)blockIdx.x * NE * F_COUNT;
  for (int t = threadIdx.x; t < nHere * F_COUNT; t += NE) {
    const int s = t / F_COUNT, f = t - s * F_COUNT;
    out[t] = sm[f * NEP + s];
  }
}

__global__ void __launch_bounds__(256) fluid_gather_val_kernel(int nnz, double mu4,
                                                               const int *__restrict__ blkOrder,
                                                               const int *__restrict__ adjPtr,
                                                               const int *__restrict__ adj,
                                                               const double *__restrict__ elemP,
                                                               double *__restrict__ Val) {
  const int lane = threadIdx.x & 31, q = lane & 7;
  const unsigned gmask = 0xFFu << (lane & 24);
  const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 3);
  if (g >= nnz) return;
  const int p = blkOrder ? __ldg(blkOrder + g) : g;
  const int s = __ldg(adjPtr + p), e = __ldg(adjPtr + p + 1);
  double acc0 = 0.0, acc1 = 0.0;
  for (int base = s; base < e; base += 8) {
    const int mine = base + q;
    const int cq = (mine < e) ? __ldg(adj + mine) : 0;
    const int cnt = min(8, e - base);
    for (int k = 0; k < cnt; k++) {
      const int pk = __shfl_sync(gmask, cq, k, 8);
      double v0, v1;
      tangent_pair<1>(elemP + (size_t)(pk >> 4) * F_COUNT, (pk >> 2) & 3, pk & 3, q, mu4, v0, v1);
      acc0 += v0;
      acc1 += v1;
    }
  }
  __stcs((double2 *)(Val + (size_t)p * 16) + q, make_double2(acc0, acc1));
}

// Record prefetch for kernel B.  The gather is latency bound (1560 SM cycles per warp step of four
// contributions at 36 resident warps: ncu, profiles/r01_ncu_asm_summary.md): every step waits for its six
// operand loads, 44% of which miss L2 (the records were just streamed to DRAM by kernel A).  A group
// knows its next eight contributions as soon as it has read its slice of the list, so lane q
// prefetches the three 64..128-byte pieces contribution q will read (node record a, node record b,
// the (D,E) pair) before the group walks the eight contributions in order: eight records in
// flight per group instead of one, no registers held.  PF: 1 = prefetch.global.L1 (CCTL.E.PF1),
// 2 = prefetch.global.L2 (CCTL.E.PF2).  (Loads into registers that are never read do not survive ptxas.)
template <int PF>
__device__ __forceinline__ void prefetch_contribution(const double *__restrict__ elemP, int pk, bool valid) {
  if (PF == 0 || !valid) return;
  const double *rec = elemP + (size_t)(pk >> 4) * F_COUNT;
  const double *ra = rec + ((pk >> 2) & 3) * 8, *rb = rec + (pk & 3) * 8;
  const double *de = rec + F_DE + (pk & 15) * 2;
  if (PF == 1) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(ra));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(rb));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(de));
  } else {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ra));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(rb));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(de));
  }
}

// Kernel B with tuning knobs (measured in profiles/r01_asm_variants.md): U2 = two contributions in
// flight per lane, THREADS per CTA, MINB = minimum resident CTAs (register cap)
template <bool U2, int THREADS, int MINB, int PF = 0>
__global__ void __launch_bounds__(THREADS, MINB) fluid_gather_val_t_kernel(
    int nnz, double mu4, const int *__restrict__ blkOrder, const int *__restrict__ adjPtr,
    const int *__restrict__ adj, const double *__restrict__ elemP, double *__restrict__ Val) {
  const int lane = threadIdx.x & 31, q = lane & 7;
  const unsigned gmask = 0xFFu << (lane & 24);
  const int g = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 3);
  if (g >= nnz) return;
  const int p = blkOrder ? __ldg(blkOrder + g) : g;
  const int s = __ldg(adjPtr + p), e = __ldg(adjPtr + p + 1);
  double acc0 = 0.0, acc1 = 0.0;
  for (int base = s; base < e; base += 8) {
    const int mine = base + q;
    const int cq = (mine < e) ? __ldg(adj + mine) : 0;
    prefetch_contribution<PF>(elemP, cq, mine < e);
    const int cnt = min(8, e - base);
    int k = 0;
    if (U2) {
      for (; k + 1 < cnt; k += 2) {
        const int pa = __shfl_sync(gmask, cq, k, 8), pb = __shfl_sync(gmask, cq, k + 1, 8);
        double a0, a1, b0, b1;
        tangent_pair<1>(elemP + (size_t)(pa >> 4) * F_COUNT, (pa >> 2) & 3, pa & 3, q, mu4, a0, a1);
        tangent_pair<1>(elemP + (size_t)(pb >> 4) * F_COUNT, (pb >> 2) & 3, pb & 3, q, mu4, b0, b1);
        acc0 += a0; acc1 += a1;
        acc0 += b0; acc1 += b1;
      }
    }
    for (; k < cnt; k++) {
      const int pk = __shfl_sync(gmask, cq, k, 8);
      double v0, v1;
      tangent_pair<1>(elemP + (size_t)(pk >> 4) * F_COUNT, (pk >> 2) & 3, pk & 3, q, mu4, v0, v1);
      acc0 += v0;
      acc1 += v1;
    }
  }
  __stcs((double2 *)(Val + (size_t)p * 16) + q, make_double2(acc0, acc1));
}

// Kernel A, second version: register budget capped at 128 (four CTAs of 128 threads per SM instead
// of two) and the record staged through shared memory in two halves of 40 fields, so that the
// staging buffer (41 KB per CTA) no longer limits residency.
__global__ void __launch_bounds__(NE, 4) fluid_record2_kernel(FluidPar par, int n,
                                                              const int *__restrict__ ien,
                                                              const double *__restrict__ x,
                                                              const double *__restrict__ Ag,
                                                              const double *__restrict__ Yg,
                                                              const double *__restrict__ Bf,
                                                              double *__restrict__ elemP,
                                                              int *__restrict__ badJac) {
  extern __shared__ double sm[];  // [40][NEP]
  const int slot = threadIdx.x;
  const int e = blockIdx.x * NE + slot;
  int nodes[4];
  ElemAcc acc;
  if (e < n) fluid_elem_compute(par, e, ien, x, Ag, Yg, Bf, acc, nodes, badJac);
  const int nHere = min(NE, n - blockIdx.x * NE);
  double *out = elemP + (size_t)blockIdx.x * NE * F_COUNT;
  if (e < n) fluid_elem_store<0>(par, acc, sm + slot);
  __syncthreads();
  for (int t = threadIdx.x; t < nHere * 40; t += NE) {
    const int s = t / 40, f = t - s * 40;
    out[s * F_COUNT + f] = sm[f * NEP + s];
  }
  __syncthreads();
  if (e < n) fluid_elem_store<1>(par, acc, sm + slot);
  __syncthreads();
  for (int t = threadIdx.x; t < nHere * 40; t += NE) {
    const int s = t / 40, f = t - s * 40;
    out[s * F_COUNT + 40 + f] = sm[f * NEP + s];
  }
}

// Kernel A, third version (default): same one-thread-per-element reduction with
//  * the FAST arithmetic of fluid_elem_compute (one reciprocal of Jac, two rsqrt per Gauss point instead
//    of two sqrt + three divisions: ~600 fewer instructions per element),
//  * the record staged as 40 (field 2k, field 2k+1) PAIRS: sm2[k][slot] double2 -> 40 conflict-free
//    STS.128 per thread instead of 80 STS.64, and a copy-out of 40 LDS.128 + STG.128 per thread with
//    incremental (slot, pair) indices (the 80-iteration loop with a division by 80 was ~20% of the
//    kernel's instructions),
//  * the ten distinct Nx_a.Nx_b products computed once.
template <int STRIDE = NEP>
__device__ __forceinline__ void fluid_elem_store_pairs(const FluidPar &par, const ElemAcc &acc,
                                                       double2 *rec2) {
  constexpr int NEP = STRIDE;   // (shadows the file-scope stride: slots per field in the staging buffer)
  const double rho = par.rho, mu = par.mu;
  double nn[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = a; b < 4; b++) {
      nn[a][b] = acc.Nx[a][0] * acc.Nx[b][0] + acc.Nx[a][1] * acc.Nx[b][1] + acc.Nx[a][2] * acc.Nx[b][2];
      nn[b][a] = nn[a][b];
    }
#pragma unroll
  for (int a = 0; a < 4; a++) {
    rec2[(a * 4 + 0) * NEP] = make_double2(acc.Nx[a][0], acc.Nx[a][1]);
    rec2[(a * 4 + 1) * NEP] = make_double2(acc.Nx[a][2], rho * acc.c2[a]);
    rec2[(a * 4 + 2) * NEP] = make_double2(acc.sTC, acc.wl);
    rec2[(a * 4 + 3) * NEP] = make_double2(acc.sTM, rho * acc.r2[a]);
#pragma unroll
    for (int b = 0; b < 4; b++)
      rec2[(F_DE / 2 + a * 4 + b) * NEP] =
          make_double2(4.0 * (mu * nn[a][b]) + acc.A[a][b], acc.sTM * nn[a][b]);
    double lr[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
      lr[i] = acc.wr * acc.sNrV[a][i] +
              acc.w * (acc.Nx[a][0] * acc.sRM[0][i] + acc.Nx[a][1] * acc.sRM[1][i] +
                       acc.Nx[a][2] * acc.sRM[2][i]);
    rec2[(F_LR / 2 + a * 2 + 0) * NEP] = make_double2(lr[0], lr[1]);
    rec2[(F_LR / 2 + a * 2 + 1) * NEP] = make_double2(lr[2], acc.w * acc.lR4[a]);
  }
}
__global__ void __launch_bounds__(NE) fluid_record3_kernel(FluidPar par, int n,
                                                           const int *__restrict__ ien,
                                                           const double *__restrict__ x,
                                                           const double *__restrict__ Ag,
                                                           const double *__restrict__ Yg,
                                                           const double *__restrict__ Bf,
                                                           double *__restrict__ elemP,
                                                           int *__restrict__ badJac) {
  extern __shared__ double2 sm2[];  // [F_COUNT / 2][NEP]
  constexpr int NP = F_COUNT / 2;   // 40 pairs per element
  const int slot = threadIdx.x;
  const int e = blockIdx.x * NE + slot;
  int nodes[4];
  if (e < n) {
    ElemAcc acc;
    fluid_elem_compute<true>(par, e, ien, x, Ag, Yg, Bf, acc, nodes, badJac);
    fluid_elem_store_pairs(par, acc, sm2 + slot);
  }
  __syncthreads();
  const int nHere = min(NE, n - blockIdx.x * NE);
  double2 *out = (double2 *)(elemP + (size_t)blockIdx.x * NE * F_COUNT);
  // thread t copies pairs t, t + NE, ...: (slot, pair) advance by (NE / NP, NE % NP) with carry
  int s = threadIdx.x / NP, k = threadIdx.x - s * NP;
#pragma unroll 4
  for (int t = threadIdx.x; t < nHere * NP; t += NE) {
    __stcg(out + t, sm2[k * NEP + s]);
    s += NE / NP;
    k += NE % NP;
    if (k >= NP) { k -= NP; s++; }
  }
}

// Kernel A, fourth version: the regrouped Gauss-point algebra of fluid_elem_compute2 (asm_elem.h), same record
// and staging as v3.  UNR = unroll factor of the Gauss-point loop.
template <int UNR>
__global__ void __launch_bounds__(NE) fluid_record6_kernel(FluidPar par, int n,
                                                           const int *__restrict__ ien,
                                                           const double *__restrict__ x,
                                                           const double *__restrict__ Ag,
                                                           const double *__restrict__ Yg,
                                                           const double *__restrict__ Bf,
                                                           double *__restrict__ elemP,
                                                           int *__restrict__ badJac) {
  extern __shared__ double2 sm2[];  // [F_COUNT / 2][NEP]
  constexpr int NP = F_COUNT / 2;
  const int slot = threadIdx.x;
  const int e = blockIdx.x * NE + slot;
  int nodes[4];
  if (e < n) {
    ElemAcc acc;
    fluid_elem_compute2<UNR>(par, e, ien, x, Ag, Yg, Bf, acc, nodes, badJac);
    fluid_elem_store_pairs(par, acc, sm2 + slot);
  }
  __syncthreads();
  const int nHere = min(NE, n - blockIdx.x * NE);
  double2 *out = (double2 *)(elemP + (size_t)blockIdx.x * NE * F_COUNT);
  int s = threadIdx.x / NP, k = threadIdx.x - s * NP;
#pragma unroll 4
  for (int t = threadIdx.x; t < nHere * NP; t += NE) {
    __stcg(out + t, sm2[k * NEP + s]);
    s += NE / NP;
    k += NE % NP;
    if (k >= NP) { k -= NP; s++; }
  }
}

// Kernel B, second version: one CTA walks a whole chunk of GCH = 512 consecutive (length-sorted)
// blocks in 16 rounds of 32 groups, so that the ~34 block rows of a chunk -- whose element records
// overlap 16-fold -- are gathered through ONE SM's L1; two contributions are in flight per lane.
static constexpr int GCH = 512;
__global__ void __launch_bounds__(256, 5) fluid_gather_val2_kernel(int nnz, double mu4,
                                                                   const int *__restrict__ blkOrder,
                                                                   const int *__restrict__ adjPtr,
                                                                   const int *__restrict__ adj,
                                                                   const double *__restrict__ elemP,
                                                                   double *__restrict__ Val) {
  const int lane = threadIdx.x & 31, q = lane & 7;
  const unsigned gmask = 0xFFu << (lane & 24);
  const int grp = threadIdx.x >> 3;
  const int base0 = blockIdx.x * GCH;
#pragma unroll 1
  for (int it = 0; it < GCH / 32; it++) {
    const int g = base0 + it * 32 + grp;
    if (g >= nnz) break;   // whole 8-lane groups leave together
    const int p = blkOrder ? __ldg(blkOrder + g) : g;
    const int s = __ldg(adjPtr + p), e = __ldg(adjPtr + p + 1);
    double acc0 = 0.0, acc1 = 0.0;
    for (int base = s; base < e; base += 8) {
      const int mine = base + q;
      const int cq = (mine < e) ? __ldg(adj + mine) : 0;
      const int cnt = min(8, e - base);
      int k = 0;
      for (; k + 1 < cnt; k += 2) {
        const int pa = __shfl_sync(gmask, cq, k, 8), pb = __shfl_sync(gmask, cq, k + 1, 8);
        double a0, a1, b0, b1;
        tangent_pair<1>(elemP + (size_t)(pa >> 4) * F_COUNT, (pa >> 2) & 3, pa & 3, q, mu4, a0, a1);
        tangent_pair<1>(elemP + (size_t)(pb >> 4) * F_COUNT, (pb >> 2) & 3, pb & 3, q, mu4, b0, b1);
        acc0 += a0; acc1 += a1;      // ascending element order, as the reference's element loop
        acc0 += b0; acc1 += b1;
      }
      if (k < cnt) {
        const int pa = __shfl_sync(gmask, cq, k, 8);
        double a0, a1;
        tangent_pair<1>(elemP + (size_t)(pa >> 4) * F_COUNT, (pa >> 2) & 3, pa & 3, q, mu4, a0, a1);
        acc0 += a0; acc1 += a1;
      }
    }
    __stcs((double2 *)(Val + (size_t)p * 16) + q, make_double2(acc0, acc1));
  }
}

__global__ void __launch_bounds__(256) fluid_gather_r_kernel(int nNo, const int *__restrict__ adjPtr,
                                                             const int *__restrict__ adj,
                                                             const double *__restrict__ elemP,
                                                             double *__restrict__ R) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nNo * 4) return;
  const int node = t >> 2, i = t & 3;
  double acc = 0.0;
  for (int k = __ldg(adjPtr + node); k < __ldg(adjPtr + node + 1); k++) {
    const int pk = __ldg(adj + k);
    acc += __ldg(elemP + (size_t)(pk >> 2) * F_COUNT + F_LR + (pk & 3) * 4 + i);
  }
  R[t] = acc;
}

// ---------------------------------------------------------------------------
// Kernel B, lean version.  Same ownership as fluid_gather_val_t_kernel (8 lanes per block, lane q owns
// row i = q>>1, columns j0 = 2(q&1), j0+1; contributions in ascending element order), but the four
// (row kind, column pair) cases of tangent_pair are folded into LANE-CONSTANT coefficients,
//     v0 = wl ( P0 (A.x bi) + u  B.x  + T0 de ),      u  = sTC ai   (momentum rows) | 1 (continuity row)
//     v1 = wl ( P1 (A.y bi) + u1 By   + T1 de ),      u1 = ai, By = -1 on the pressure column
// so the inner loop carries 6 selects instead of 16 + 13 predicate computations, and every operand
// address is one IMAD.WIDE from a 32-bit record index with the lane's field offset folded into the
// base pointer.  ~40 instructions per warp step of four contributions instead of ~75 (SASS).
// sum_g N_a(g) (sN_of: 1 to within an ulp) is taken as exactly 1 here: a 1e-16 relative change.
template <int THREADS, int MINB, int PF = 0>
__global__ void __launch_bounds__(THREADS, MINB) fluid_gather_lean_kernel(
    int nnz, double mu4, const int *__restrict__ blkOrder, const int *__restrict__ adjPtr,
    const int *__restrict__ adj, const double *__restrict__ elemP, double *__restrict__ Val) {
  const int lane = threadIdx.x & 31, q = lane & 7;
  const unsigned gmask = 0xFFu << (lane & 24);
  const int g = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 3);
  if (g >= nnz) return;
  const int p = blkOrder ? __ldg(blkOrder + g) : g;
  const int s = __ldg(adjPtr + p), e = __ldg(adjPtr + p + 1);
  const int i = q >> 1, j0 = (q & 1) * 2;
  const bool row3 = (i == 3), col2 = (j0 == 2), pcol = col2 && !row3;
  const double P0 = row3 ? 1.0 : mu4;
  const double P1 = row3 ? (col2 ? 0.0 : 1.0) : (col2 ? 1.0 : mu4);
  const double T0 = (!row3 && i == j0) ? 1.0 : 0.0;
  const double T1 = row3 ? (col2 ? 1.0 : 0.0) : ((!col2 && i == 1) ? 1.0 : 0.0);
  const double cBy = row3 ? 0.0 : -1.0;      // By on the last column pair: -sum_g N_b (momentum) | unused
  // lane-constant field offsets inside a node record / the element record (32-bit index arithmetic,
  // one IMAD.WIDE.U32 per address)
  const unsigned oA = (unsigned)j0;                            // (Nx_j0, Nx_j0+1 | C2) of a node record
  const unsigned oI = row3 ? (unsigned)N_R2 : (unsigned)i;     // Nx_i, or R2 on the continuity row
  const unsigned oD = (unsigned)F_DE + (row3 ? 1u : 0u);       // D_ab | E_ab
  double acc0 = 0.0, acc1 = 0.0;
  for (int base = s; base < e; base += 8) {
    const int mine = base + q;
    const int cq = (mine < e) ? __ldg(adj + mine) : 0;
    prefetch_contribution<PF>(elemP, cq, mine < e);
    const int cnt = min(8, e - base);
#pragma unroll 1
    for (int k = 0; k < cnt; k++) {
      const unsigned pk = (unsigned)__shfl_sync(gmask, cq, k, 8);
      const unsigned eb = (pk >> 4) * (unsigned)F_COUNT;       // record index (32 bit: nEl * 80 < 2^32)
      const unsigned ia = eb + ((pk << 1) & 24u), ib = eb + ((pk << 3) & 24u);
      const double2 A = __ldg((const double2 *)(elemP + (size_t)(ia + oA)));
      const double2 B = __ldg((const double2 *)(elemP + (size_t)(ib + oA)));
      const double2 S = __ldg((const double2 *)(elemP + (size_t)(ia + (unsigned)N_STC)));
      const double ai = __ldg(elemP + (size_t)(ia + oI));
      const double bi = __ldg(elemP + (size_t)(ib + oI));
      const double de = __ldg(elemP + (size_t)(ia + ((pk << 1) & 6u) + oD));
      const double t = S.x * ai;
      const double u = row3 ? 1.0 : t;
      const double u1 = pcol ? ai : u;
      const double By = col2 ? cBy : B.y;
      double s0 = u * B.x;
      double s1 = u1 * By;
      s0 = fma(P0, A.x * bi, s0);
      s1 = fma(P1, A.y * bi, s1);
      s0 = fma(T0, de, s0);
      s1 = fma(T1, de, s1);
      acc0 += S.y * s0;
      acc1 += S.y * s1;
    }
  }
  __stcs((double2 *)(Val + (size_t)p * 16) + q, make_double2(acc0, acc1));
}

// ---------------------------------------------------------------------------
// Kernel B, wide-load version.  Measured (profiles/r01_asm_variants.md): neither 24% fewer instructions
// (lean kernel) nor record prefetch moves kernel B, and its time equals the L1TEX wavefront-queue model
// of B300_MICROARCH.md ("1.0 cycle per LDG + 2.07 per additional 128-byte line of the same LDG"): the
// four 8-lane groups of a warp work on four different elements, so each of the SIX operand loads of a
// step touches four lines = 43 cycles per step = 6.1 ms.  The lever is therefore the NUMBER of load
// instructions, not bytes or instructions.  sm_100 has 256-bit loads (LDG.E.ENL2.256): one of them
// fetches the first half of a node record (Nx_0..2, C2), which holds the lane's column pair AND its
// row component; the second half (sum tauC, wl, sum tauM, R2) comes from node a on the momentum rows
// and from node b on the continuity row (the element-wide scalars are replicated in every node record,
// and the continuity row needs R2 of b).  Four loads per step instead of six; operands are picked
// with lane-constant selects.  Arithmetic as in the lean kernel.
struct __align__(32) dbl4 { double x, y, z, w; };
__device__ __forceinline__ dbl4 ldg256(const double *p) {
  dbl4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fluid_gather_wide_kernel(
    int nnz, double mu4, const int *__restrict__ blkOrder, const int *__restrict__ adjPtr,
    const int *__restrict__ adj, const double *__restrict__ elemP, double *__restrict__ Val) {
  const int lane = threadIdx.x & 31, q = lane & 7;
  const unsigned gmask = 0xFFu << (lane & 24);
  const int g = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 3);
  if (g >= nnz) return;
  const int p = blkOrder ? __ldg(blkOrder + g) : g;
  const int s = __ldg(adjPtr + p), e = __ldg(adjPtr + p + 1);
  const int i = q >> 1, j0 = (q & 1) * 2;
  const bool row3 = (i == 3), col2 = (j0 == 2), pcol = col2 && !row3, i0 = (i == 0), i1 = (i == 1);
  const double P0 = row3 ? 1.0 : mu4;
  const double P1 = row3 ? (col2 ? 0.0 : 1.0) : (col2 ? 1.0 : mu4);
  const double T0 = (!row3 && i == j0) ? 1.0 : 0.0;
  const double T1 = row3 ? (col2 ? 1.0 : 0.0) : ((!col2 && i == 1) ? 1.0 : 0.0);
  const double cBy = row3 ? 0.0 : -1.0;
  const unsigned oD = (unsigned)F_DE + (row3 ? 1u : 0u);
  double acc0 = 0.0, acc1 = 0.0;
  for (int base = s; base < e; base += 8) {
    const int mine = base + q;
    const int cq = (mine < e) ? __ldg(adj + mine) : 0;
    const int cnt = min(8, e - base);
#pragma unroll 1
    for (int k = 0; k < cnt; k++) {
      const unsigned pk = (unsigned)__shfl_sync(gmask, cq, k, 8);
      const unsigned eb = (pk >> 4) * (unsigned)F_COUNT;       // record index (32 bit: nEl * 80 < 2^32)
      const unsigned ia = eb + ((pk << 1) & 24u), ib = eb + ((pk << 3) & 24u);
      const dbl4 La = ldg256(elemP + (size_t)ia);                          // Nx_0..2, C2 of a
      const dbl4 Lb = ldg256(elemP + (size_t)ib);                          // Nx_0..2 (, C2) of b
      const dbl4 Ls = ldg256(elemP + (size_t)((row3 ? ib : ia) + 4u));     // sum tauC, wl, -, R2
      const double de = __ldg(elemP + (size_t)(ia + ((pk << 1) & 6u) + oD));
      const double Ax = col2 ? La.z : La.x, Ay = col2 ? La.w : La.y;
      const double Bx = col2 ? Lb.z : Lb.x, By = col2 ? cBy : Lb.y;
      const double ai = i0 ? La.x : (i1 ? La.y : La.z);
      const double bi = row3 ? Ls.w : (i0 ? Lb.x : (i1 ? Lb.y : Lb.z));
      const double t = Ls.x * ai;
      const double u = row3 ? 1.0 : t;
      const double u1 = pcol ? ai : u;
      double s0 = u * Bx;
      double s1 = u1 * By;
      s0 = fma(P0, Ax * bi, s0);
      s1 = fma(P1, Ay * bi, s1);
      s0 = fma(T0, de, s0);
      s1 = fma(T1, de, s1);
      acc0 += Ls.y * s0;
      acc1 += Ls.y * s1;
    }
  }
  __stcs((double2 *)(Val + (size_t)p * 16) + q, make_double2(acc0, acc1));
}

// ---------------------------------------------------------------------------
// Kernel B, quad version: FOUR lanes per block, lane r owns row r of the 4x4 block (four entries).
// ncu's per-instruction stall samples of the 8-lane kernel (profiles/r01_asm_variants.md) put 30% of the
// wait on the three dependent loads that start a group (processing order -> list bounds -> list) and
// 37% on the first use of each contribution's operands: a group is a chain of ~9 serialised memory
// round trips, and only warps-per-SM x groups-per-warp of them run at once.  With four lanes per
// block a warp carries EIGHT chains instead of four for about the same registers; the 256-bit loads
// of the wide kernel give a lane the whole (Nx_0..2, C2) of both nodes, which is exactly what a row of
// the block needs.  Per step: 4 loads and ~55 instructions for eight contributions (8-lane kernel:
// 6 loads, ~75 instructions for four).  Contributions still arrive in ascending element order.
template <int THREADS, int MINB, bool DESC = false>
__global__ void __launch_bounds__(THREADS, MINB) fluid_gather_quad_kernel(
    int nnz, double mu4, const int *__restrict__ blkOrder, const int *__restrict__ adjPtr,
    const int *__restrict__ adj, const double *__restrict__ elemP, double *__restrict__ Val,
    const int4 *__restrict__ desc) {
  const int lane = threadIdx.x & 31, r = lane & 3;
  const unsigned gmask = 0xFu << (lane & 28);
  const int g = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 2);
  if (g >= nnz) return;   // whole 4-lane groups leave together
  int p, s, e;
  if (DESC) {   // one 16-byte load: (block, list begin, list end)
    const int4 d = __ldg(desc + g);
    p = d.x; s = d.y; e = d.z;
  } else {
    p = blkOrder ? __ldg(blkOrder + g) : g;
    s = __ldg(adjPtr + p); e = __ldg(adjPtr + p + 1);
  }
  const bool row3 = (r == 3), i0 = (r == 0), i1 = (r == 1), i2 = (r == 2);
  const double P = row3 ? 1.0 : mu4;       // coefficient of Nx_a(j) * bi, j < 3
  const double P3 = row3 ? 0.0 : 1.0;      // last column: C2_a * Nx_b(i) on a momentum row
  const double c3 = row3 ? 0.0 : -1.0;     // last column: -Nx_a(i) (sum_g N_b = 1) on a momentum row
  const unsigned oD = (unsigned)F_DE + (row3 ? 1u : 0u);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int cq = (s + r < e) ? __ldg(adj + s + r) : 0;
  for (int base = s; base < e; base += 4) {
    // the next slice of the list is in flight while this one is processed
    const int nxt = base + 4 + r;
    const int cqn = (nxt < e) ? __ldg(adj + nxt) : 0;
    const int cnt = min(4, e - base);
#pragma unroll 1
    for (int k = 0; k < cnt; k++) {
      const unsigned pk = (unsigned)__shfl_sync(gmask, cq, k, 4);
      const unsigned eb = (pk >> 4) * (unsigned)F_COUNT;       // record index (32 bit: nEl * 80 < 2^32)
      const unsigned ia = eb + ((pk << 1) & 24u), ib = eb + ((pk << 3) & 24u);
      const dbl4 La = ldg256(elemP + (size_t)ia);                          // Nx_0..2, C2 of a
      const dbl4 Lb = ldg256(elemP + (size_t)ib);                          // Nx_0..2 of b
      const dbl4 Ls = ldg256(elemP + (size_t)((row3 ? ib : ia) + 4u));     // sum tauC, wl, -, R2
      const double de = __ldg(elemP + (size_t)(ia + ((pk << 1) & 6u) + oD));   // D_ab | E_ab
      const double ai = i0 ? La.x : (i1 ? La.y : La.z);        // Nx_i of a (unused on the continuity row)
      const double bi = row3 ? Ls.w : (i0 ? Lb.x : (i1 ? Lb.y : Lb.z));   // Nx_i of b | R2_b
      const double u = row3 ? 1.0 : Ls.x * ai;
      double s0 = fma(u, Lb.x, i0 ? de : 0.0);
      double s1 = fma(u, Lb.y, i1 ? de : 0.0);
      double s2 = fma(u, Lb.z, i2 ? de : 0.0);
      double s3 = fma(ai, c3, row3 ? de : 0.0);
      s0 = fma(P, La.x * bi, s0);
      s1 = fma(P, La.y * bi, s1);
      s2 = fma(P, La.z * bi, s2);
      s3 = fma(P3, La.w * bi, s3);
      a0 += Ls.y * s0;
      a1 += Ls.y * s1;
      a2 += Ls.y * s2;
      a3 += Ls.y * s3;
    }
    cq = cqn;
  }
  double2 *out = (double2 *)(Val + (size_t)p * 16 + r * 4);
  __stcs(out, make_double2(a0, a1));
  __stcs(out + 1, make_double2(a2, a3));
}

// ---------------------------------------------------------------------------
// Kernels B + C, pair-owner version (default).  The tangent blocks (a,b) and (b,a) of one element
// are built from the SAME operands with the roles of the two nodes exchanged (S/FLUID.f:482-557,
// :1052-1081): Nx_a, Nx_b, C2, R2, the element-wide (sum tauC, wl), and only the (D,E) pair differs.
// So 8 lanes own the PAIR of blocks {(r,c), (c,r)}, c > r: they walk the (element, a, b) list of (r,c)
// once -- which is also the list of (c,r) with a and b swapped -- load 7 operands instead of 2 x 6,
// evaluate both blocks and write each ONCE.  Diagonal blocks (r,r) are single; their list is the list
// of elements around node r in ascending order, i.e. exactly what the residual gather walks, so
// lanes 0..3 of a diagonal group also sum lR(:,a) and write R(:,r): kernel C disappears.
// Every block still receives its contributions in ascending element order starting from 0.0
// (deterministic; the accumulation order of the reference's element loop, S/LHSA.f:275-295).
__device__ __forceinline__ void tangent_eval(double2 A, double2 B, double2 S, double ai, double bi,
                                             double de, double sNa, double sNb, int i, int j0,
                                             bool row3, bool col2, double mu4, double &v0, double &v1) {
  const double wl = S.y, sTC = S.x;
  // momentum rows, velocity columns (ai = Nx_i of a, bi = Nx_i of b)
  const double m0 = (mu4 * (A.x * bi) + sTC * (ai * B.x) + ((i == j0) ? de : 0.0)) * wl;
  const double m1 = (mu4 * (A.y * bi) + sTC * (ai * B.y) + ((i == 1) ? de : 0.0)) * wl;
  const double mp = -wl * (ai * sNb - bi * A.y);               // pressure column (A.y = C2_a)
  // continuity row (bi = R2_b)
  const double c0 = wl * (sNa * B.x + A.x * bi);
  const double c1 = wl * (sNa * B.y + A.y * bi);
  const double cp = wl * de;                                   // de = sum tauM Nx_a.Nx_b
  v0 = row3 ? c0 : m0;
  v1 = row3 ? (col2 ? cp : c1) : (col2 ? mp : m1);
}
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fluid_gather_pairs_kernel(
    int nPair, double mu4, const int *__restrict__ pairList, const int *__restrict__ pairT,
    const int *__restrict__ rowOf, const int *__restrict__ adjPtr, const int *__restrict__ adj,
    const double *__restrict__ elemP, double *__restrict__ Val, double *__restrict__ R) {
  const int lane = threadIdx.x & 31, q = lane & 7;
  const unsigned gmask = 0xFFu << (lane & 24);
  const int g = (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 3);
  if (g >= nPair) return;   // whole 8-lane groups leave together
  const int p = __ldg(pairList + g), pt = __ldg(pairT + g);
  const bool diag = (pt == p), both = (pt >= 0) && !diag;
  const int s = __ldg(adjPtr + p), e = __ldg(adjPtr + p + 1);
  const int i = q >> 1, j0 = (q & 1) * 2;
  const bool row3 = (i == 3), col2 = (j0 == 2);
  // lane-constant offsets inside a node record: (Nx_j0, Nx_j0+1 | C2); Nx_i on a momentum row, R2
  // on the continuity row (where Nx_i of the ROW node is not used)
  const int oI = row3 ? N_R2 : i, oDE = row3 ? 1 : 0;
  double acc0 = 0.0, acc1 = 0.0, t0 = 0.0, t1 = 0.0, r = 0.0;
  for (int base = s; base < e; base += 8) {
    const int mine = base + q;
    const int cq = (mine < e) ? __ldg(adj + mine) : 0;
    const int cnt = min(8, e - base);
    for (int k = 0; k < cnt; k++) {
      const int pk = __shfl_sync(gmask, cq, k, 8);
      const int a = (pk >> 2) & 3, b = pk & 3;
      const double *rec = elemP + (size_t)(pk >> 4) * F_COUNT;
      const double *ra = rec + a * 8, *rb = rec + b * 8;
      const double2 A = __ldg((const double2 *)(ra + j0));
      const double2 S = __ldg((const double2 *)(ra + N_STC));
      const double ai = __ldg(ra + oI);
      const double de = __ldg(rec + F_DE + (a * 4 + b) * 2 + oDE);
      const double sNa = sN_of(a), sNb = sN_of(b);
      double v0, v1;
      if (diag) {   // group-uniform: a == b
        tangent_eval(A, A, S, ai, ai, de, sNa, sNa, i, j0, row3, col2, mu4, v0, v1);
        if (q < 4) r += __ldg(rec + F_LR + a * 4 + q);
      } else {
        const double2 B = __ldg((const double2 *)(rb + j0));
        const double bi = __ldg(rb + oI);
        tangent_eval(A, B, S, ai, bi, de, sNa, sNb, i, j0, row3, col2, mu4, v0, v1);
        if (both) {
          const double de2 = __ldg(rec + F_DE + (b * 4 + a) * 2 + oDE);
          double w0, w1;
          tangent_eval(B, A, S, bi, ai, de2, sNb, sNa, i, j0, row3, col2, mu4, w0, w1);
          t0 += w0;
          t1 += w1;
        }
      }
      acc0 += v0;
      acc1 += v1;
    }
  }
  __stcs((double2 *)(Val + (size_t)p * 16) + q, make_double2(acc0, acc1));
  if (both) __stcs((double2 *)(Val + (size_t)pt * 16) + q, make_double2(t0, t1));
  if (diag && q < 4) R[(size_t)__ldg(rowOf + p) * 4 + q] = r;
}

// ---------------------------------------------------------------------------
// Kernels B + C, row-owner version: ONE WARP PER BLOCK ROW.  The warp walks the (element, a) pairs
// around its node in ascending element order; at every visit lane group b = lane>>3 expands the
// tangent block (a, b) (two entries per lane, q = lane&7) and adds it to the row's block
// (row, ien[e][b]) held in shared memory; lanes 0..3 carry the residual.  The finished row leaves
// as ONE coalesced streaming store of nblk*128 bytes.
//  * b and q are lane constants, so the four cases of tangent_pair (momentum / continuity row,
//    velocity / pressure column pair) become lane-constant COEFFICIENTS of one expression
//        v = wl * ( P (A bi) + Q (x1 B) + T de )
//    (coefficients 0 / 1 select exactly: 1*x and x+0 do not round), and all lane-dependent record
//    offsets are loop invariant.  ~55 instructions and ~9 L1 wavefronts of operands per visit of four
//    contributions; the block-owner kernel needs ~85 and ~25 (every 8-lane group reads its own two
//    node records).
//  * the slot of block (row, ien[e][b]) inside the row comes from nodeSlots (four 8-bit slots per
//    visit, built once per pattern), read coalesced next to the visit list.
//  * the four blocks of a visit are distinct (four distinct nodes of a tet), so lanes never collide
//    inside a visit; successive visits may hit the same block from another lane group: __syncwarp.
//  * every block still receives its contributions in ascending element order starting from 0.0:
//    deterministic, the accumulation order of the reference's element loop.
struct VisitOps {
  double2 A, B, S;
  double ai, bi, de;
  int slot;
};
template <int U, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) fluid_gather_rows_kernel(
    int nNo, int maxRow, double mu4, const int *__restrict__ rowPtr,
    const int *__restrict__ nodeAdjPtr, const int *__restrict__ nodeAdj,
    const int *__restrict__ nodeSlots, const double *__restrict__ elemP, double *__restrict__ Val,
    double *__restrict__ R) {
  extern __shared__ double2 smrow[];   // [WARPS][maxRow * 8]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * WARPS + warp;
  if (row >= nNo) return;
  double2 *acc = smrow + (size_t)warp * maxRow * 8 + (lane & 7);
  const int b = lane >> 3, q = lane & 7;
  const int i = q >> 1, j0 = (q & 1) * 2;
  const bool row3 = (i == 3), col2 = (j0 == 2), momc0 = !row3 && !col2, byConst = col2 && !row3;
  // lane-constant coefficients (S/FLUID.f:482-557 momentum rows, :1052-1081 continuity row)
  const double P0 = row3 ? 1.0 : mu4;
  const double P1 = row3 ? (col2 ? 0.0 : 1.0) : (col2 ? 1.0 : mu4);
  const double Q1c = (row3 && col2) ? 0.0 : 1.0;
  const double T0 = (!row3 && i == j0) ? 1.0 : 0.0;
  const double T1 = row3 ? (col2 ? 1.0 : 0.0) : ((!col2 && i == 1) ? 1.0 : 0.0);
  const double nsNb = -sN_of(b);
  // lane-constant record offsets: A, ai, de relative to node record a; B, bi to the element record
  const int oA = j0, oAi = row3 ? 0 : i, oDE = F_DE + b * 2 + (row3 ? 1 : 0);
  const int oB = b * 8 + j0, oBi = b * 8 + (row3 ? N_R2 : i);
  const int shift = b * 8;
  const int rp0 = __ldg(rowPtr + row), nblk = __ldg(rowPtr + row + 1) - rp0;
  for (int t = lane; t < nblk * 8; t += 32) smrow[(size_t)warp * maxRow * 8 + t] = make_double2(0.0, 0.0);
  __syncwarp();
  const int s = __ldg(nodeAdjPtr + row), e = __ldg(nodeAdjPtr + row + 1);
  double r = 0.0;

  auto load = [&](int pk, int sl, VisitOps &o, double &lr) {
    const int a = pk & 3;
    const double *rec = elemP + (size_t)(pk >> 2) * F_COUNT, *ra = rec + a * 8;
    o.slot = (sl >> shift) & 0xff;
    o.A = __ldg((const double2 *)(ra + oA));
    o.B = __ldg((const double2 *)(rec + oB));
    o.S = __ldg((const double2 *)(ra + N_STC));
    o.ai = __ldg(ra + oAi);
    o.bi = __ldg(rec + oBi);
    o.de = __ldg(ra + oDE);
    lr = (lane < 4) ? __ldg(rec + F_LR + a * 4 + lane) : 0.0;
  };
  auto apply = [&](int pk, const VisitOps &o, double lr) {
    const double x1 = row3 ? sN_of(pk & 3) : o.ai;
    const double Q0 = row3 ? 1.0 : o.S.x;
    const double Q1 = momc0 ? o.S.x : Q1c;
    const double By = byConst ? nsNb : o.B.y;
    const double v0 = fma(T0, o.de, fma(Q0, x1 * o.B.x, P0 * (o.A.x * o.bi)));
    const double v1 = fma(T1, o.de, fma(Q1, x1 * By, P1 * (o.A.y * o.bi)));
    double2 c = acc[o.slot * 8];
    c.x += v0 * o.S.y;
    c.y += v1 * o.S.y;
    acc[o.slot * 8] = c;
    r += lr;
    __syncwarp();
  };

  for (int base = s; base < e; base += 32) {
    const bool have = base + lane < e;
    const int mine = have ? __ldg(nodeAdj + base + lane) : 0;
    const int mysl = have ? __ldg(nodeSlots + base + lane) : 0;
    const int cnt = min(32, e - base);
    int k = 0;
    if (U > 1) {
      // U visits in flight: all loads issued before the first accumulation (still applied in
      // ascending visit order)
      for (; k + U <= cnt; k += U) {
        int pk[U];
        VisitOps o[U];
        double lr[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          pk[u] = __shfl_sync(0xffffffffu, mine, k + u);
          load(pk[u], __shfl_sync(0xffffffffu, mysl, k + u), o[u], lr[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) apply(pk[u], o[u], lr[u]);
      }
    }
    for (; k < cnt; k++) {
      const int pk = __shfl_sync(0xffffffffu, mine, k);
      const int sl = __shfl_sync(0xffffffffu, mysl, k);
      VisitOps o;
      double lr;
      load(pk, sl, o, lr);
      apply(pk, o, lr);
    }
  }
  double2 *out = (double2 *)(Val + (size_t)rp0 * 16);
  const double2 *fin = smrow + (size_t)warp * maxRow * 8;
  for (int t = lane; t < nblk * 8; t += 32) __stcs(out + t, fin[t]);
  if (lane < 4) R[(size_t)row * 4 + lane] = r;
}

// nodeSlots[k] for visit k = nodeAdj entry (element e, local node a) of row r: the positions inside
// row r of the four blocks (r, ien[e][b]), 8 bits each (rows longer than 255 blocks do not use
// the row-owner kernel)
__global__ void build_node_slots_kernel(int nNo, const int *__restrict__ rowPtr,
                                        const int *__restrict__ nodeAdjPtr,
                                        const int *__restrict__ nodeAdj,
                                        const int *__restrict__ edest, int *__restrict__ slots) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nNo) return;
  const int rp0 = rowPtr[row];
  for (int k = nodeAdjPtr[row]; k < nodeAdjPtr[row + 1]; k++) {
    const int pk = nodeAdj[k];
    int v = 0;
    for (int b = 0; b < 4; b++)
      v |= ((edest[(size_t)(pk >> 2) * 16 + (pk & 3) * 4 + b] - rp0) & 0xff) << (8 * b);
    slots[k] = v;
  }
}
void launch_build_node_slots(cudaStream_t st, int nNo, const int *rowPtr, const int *nodeAdjPtr,
                             const int *nodeAdj, const int *edest, int *slots) {
  if (nNo <= 0) return;
  build_node_slots_kernel<<<(nNo + 127) / 128, 128, 0, st>>>(nNo, rowPtr, nodeAdjPtr, nodeAdj, edest,
                                                           slots);
}

static void fluid_attr_once() {
  static bool attr = false;
  if (attr) return;
  const int smem = (int)((size_t)F_COUNT * NEP * sizeof(double));
  cudaFuncSetAttribute(fluid_asm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(fluid_asm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(fluid_record_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(fluid_record2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)((size_t)40 * NEP * sizeof(double)));
  cudaFuncSetAttribute(fluid_record3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(fluid_record6_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(fluid_record6_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(fluid_record6_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  attr = true;
}

void launch_fluid_asm(cudaStream_t st, const FluidPar &par, int n, int e0, const int *elems,
                      const int *ien, const int *edest, const double *x, const double *Ag,
                      const double *Yg, const double *Bf, double *R, double *Val, int atomic,
                      int *badJac) {
  if (n <= 0) return;
  count_launch();
  const size_t smem = (size_t)F_COUNT * NEP * sizeof(double);
  fluid_attr_once();
  const int blocks = (n + NE - 1) / NE;
  if (atomic)
    fluid_asm_kernel<true><<<blocks, NE, smem, st>>>(par, n, e0, elems, ien, edest, x, Ag, Yg, Bf,
                                                     R, Val, badJac);
  else
    fluid_asm_kernel<false><<<blocks, NE, smem, st>>>(par, n, e0, elems, ien, edest, x, Ag, Yg, Bf,
                                                      R, Val, badJac);
}

// bit 0: record kernel v2 (128 registers, two-pass staging); bit 1: chunked gather kernel.
// bits 2..4: knobs of the templated tangent-gather kernel (1 = two contributions in flight,
// 2 = 128-thread CTAs, 4 = 32-register cap).  bit 5 (32): row-owner kernel for B + C (one warp per block
// row, accumulation in shared memory); bit 6 (64): two visits in flight in it.
// bit 7 (128): record kernel v3.  bits 10..13: pair-owner kernel for B + C (see the dispatch below).
// bits 14..19: lean / prefetching / wide-load / quad block-owner kernels (see the dispatch below).
// bit 23 (8388608): record kernel v4 (regrouped Gauss-point algebra, asm_elem.h fluid_elem_compute2), bits 24 / 25:
// its Gauss-point loop unrolled by 2 / 4.
// Default = 42733696 = 790656 + 8388608 + 33554432: record kernel v4 with the Gauss-point loop fully unrolled +
// quad gather (four lanes per block, 64-register cap, block descriptors), the fastest measured combination
// (profiles/r02_asm_experiments.md: 2.20 + 4.03 + 0.33 ms at 10M tets; round 1 default 790656: 2.59 + 3.99 +
// 0.33; the 8-lane kernel takes 6.15, the row-owner and pair-owner kernels 6.9 and 7.2; 64-element CTAs with a
// 170 / 128-register cap for more resident warps spill and take 3.5 / 4.7 ms);
// SVFSI_ASM_TUNE overrides (kernel-variant timings in profiles/).
int asm_tune() {
  static int t = -1;
  if (t < 0) {
    const char *e = getenv("SVFSI_ASM_TUNE");
    t = e ? atoi(e) : 42733696;
  }
  return t;
}

void launch_fluid_gather_parts(cudaStream_t st, int parts, const FluidPar &par, int nEl, int nNo,
                               int nnz, const int *ien, const double *x, const double *Ag,
                               const double *Yg, const double *Bf, double *elemP,
                               const int *blkOrder, const int *blkAdjPtr, const int *blkAdj,
                               const int *nodeAdjPtr, const int *nodeAdj, double *R, double *Val,
                               int *badJac, int tune, const int *rowPtr, const int *nodeSlots,
                               int maxRow, PairLists pairs) {
  if (nEl <= 0) return;
  fluid_attr_once();
  // bit 5: row-owner kernel (B and C in one launch); needs the row's blocks in shared memory
  const bool rows = (tune & 32) && rowPtr && nodeSlots && maxRow > 0 && maxRow <= 64;
  if (parts & 1) {
    count_launch();
    // bit 23 (8388608): records v4 = regrouped algebra; bits 24 / 25: Gauss-point loop unrolled by 2 / 4
    if (tune & (1 << 23)) {
      const size_t smem = (size_t)F_COUNT * NEP * sizeof(double);
      if (tune & (1 << 25))
        fluid_record6_kernel<4><<<(nEl + NE - 1) / NE, NE, smem, st>>>(par, nEl, ien, x, Ag, Yg, Bf, elemP, badJac);
      else if (tune & (1 << 24))
        fluid_record6_kernel<2><<<(nEl + NE - 1) / NE, NE, smem, st>>>(par, nEl, ien, x, Ag, Yg, Bf, elemP, badJac);
      else
        fluid_record6_kernel<1><<<(nEl + NE - 1) / NE, NE, smem, st>>>(par, nEl, ien, x, Ag, Yg, Bf, elemP, badJac);
    } else if (tune & 128) {
      const size_t smem = (size_t)F_COUNT * NEP * sizeof(double);
      fluid_record3_kernel<<<(nEl + NE - 1) / NE, NE, smem, st>>>(par, nEl, ien, x, Ag, Yg, Bf,
                                                                  elemP, badJac);
    } else if (tune & 1) {
      const size_t smem = (size_t)40 * NEP * sizeof(double);
      fluid_record2_kernel<<<(nEl + NE - 1) / NE, NE, smem, st>>>(par, nEl, ien, x, Ag, Yg, Bf,
                                                                  elemP, badJac);
    } else {
      const size_t smem = (size_t)F_COUNT * NEP * sizeof(double);
      fluid_record_kernel<<<(nEl + NE - 1) / NE, NE, smem, st>>>(par, nEl, ien, x, Ag, Yg, Bf,
                                                                 elemP, badJac);
    }
  }
  // bit 10 (1024): pair-owner kernel (B and C in one launch); bit 11 (2048): 256-thread CTAs
  if ((tune & 1024) && pairs.list && pairs.n > 0 && (parts & 6)) {
    count_launch();
    const size_t lanes = (size_t)pairs.n * 8;
#define GP(T, MB)                                                                              \
  fluid_gather_pairs_kernel<T, MB><<<(unsigned)((lanes + T - 1) / T), T, 0, st>>>(                 \
      pairs.n, 4.0 * par.mu, pairs.list, pairs.tpos, pairs.rowOf, blkAdjPtr, blkAdj, elemP, Val, R)
    // bit 11 (2048): 256-thread CTAs; bit 12 (4096) / bit 13 (8192): register cap for 1280 / 1024
    // resident threads per SM (48 / 64 registers; uncapped: 77)
    if (tune & 4096) {
      if (tune & 2048) GP(256, 5); else GP(128, 10);
    } else if (tune & 8192) {
      if (tune & 2048) GP(256, 4); else GP(128, 8);
    } else {
      if (tune & 2048) GP(256, 1); else GP(128, 1);
    }
#undef GP
    return;
  }
  if (rows && (parts & 6)) {
    // parts 2 and 4 are one kernel here (timed under either bit)
    count_launch();
    // bit 6 (64): two visits in flight; bit 9 (512): four; bit 8 (256): 8 warps per CTA instead of 4
#define GR(U, W)                                                                                 \
  fluid_gather_rows_kernel<U, W><<<(nNo + W - 1) / W, W * 32, (size_t)W * maxRow * 128, st>>>(     \
      nNo, maxRow, 4.0 * par.mu, rowPtr, nodeAdjPtr, nodeAdj, nodeSlots, elemP, Val, R)
    const int u = (tune & 512) ? 4 : ((tune & 64) ? 2 : 1);
    if (tune & 256) {
      if (u == 4) GR(4, 8); else if (u == 2) GR(2, 8); else GR(1, 8);
    } else {
      if (u == 4) GR(4, 4); else if (u == 2) GR(2, 4); else GR(1, 4);
    }
#undef GR
    return;
  }
  // bit 18 (262144): quad kernel (four lanes per block); bit 11 (2048): 256-thread CTAs; bit 13 (8192) /
  // bit 12 (4096): 48 / 64-register cap
  if ((parts & 2) && (tune & 262144) && (double)nEl * F_COUNT < 4.0e9) {
    count_launch();
    const size_t lanes = (size_t)nnz * 4;
    static int rowMajor = -1;   // experiment: plain row-major block order (no length sorting)
    if (rowMajor < 0) { const char *e = getenv("SVFSI_ASM_ROWMAJOR"); rowMajor = e ? atoi(e) : 0; }
    if (rowMajor) { blkOrder = nullptr; tune &= ~524288; }
#define GQ(T, MB)                                                                              \
  do {                                                                                         \
    if ((tune & 524288) && pairs.desc)                                                         \
      fluid_gather_quad_kernel<T, MB, true><<<(unsigned)((lanes + T - 1) / T), T, 0, st>>>(      \
          nnz, 4.0 * par.mu, blkOrder, blkAdjPtr, blkAdj, elemP, Val, pairs.desc);             \
    else                                                                                       \
      fluid_gather_quad_kernel<T, MB, false><<<(unsigned)((lanes + T - 1) / T), T, 0, st>>>(     \
          nnz, 4.0 * par.mu, blkOrder, blkAdjPtr, blkAdj, elemP, Val, nullptr);                \
  } while (0)
    // bit 19 (524288): block descriptors (one load instead of two dependent ones at group start)
    if (tune & 8192) {
      if (tune & 2048) GQ(256, 5); else GQ(128, 10);
    } else if (tune & 4096) {
      if (tune & 2048) GQ(256, 4); else GQ(128, 8);
    } else {
      if (tune & 2048) GQ(256, 1); else GQ(128, 1);
    }
#undef GQ
    parts &= ~2;
  }
  // bit 17 (131072): wide-load block-owner kernel; bit 11 (2048): 256-thread CTAs; bit 13 (8192) / bit 12
  // (4096): 48 / 64-register cap
  if ((parts & 2) && (tune & 131072) && (double)nEl * F_COUNT < 4.0e9) {
    count_launch();
    const size_t lanes = (size_t)nnz * 8;
#define GW(T, MB)                                                                              \
  fluid_gather_wide_kernel<T, MB><<<(unsigned)((lanes + T - 1) / T), T, 0, st>>>(                \
      nnz, 4.0 * par.mu, blkOrder, blkAdjPtr, blkAdj, elemP, Val)
    if (tune & 8192) {
      if (tune & 2048) GW(256, 5); else GW(128, 10);
    } else if (tune & 4096) {
      if (tune & 2048) GW(256, 4); else GW(128, 8);
    } else {
      if (tune & 2048) GW(256, 1); else GW(128, 1);
    }
#undef GW
    parts &= ~2;
  }
  // bit 14 (16384): lean block-owner kernel; bit 11 (2048): 256-thread CTAs; bit 12 (4096) / bit 13
  // (8192): 32 / 48-register cap
  if ((parts & 2) && (tune & 16384) && (double)nEl * F_COUNT < 4.0e9) {
    count_launch();
    const size_t lanes = (size_t)nnz * 8;
#define GL(T, MB, PF)                                                                          \
  fluid_gather_lean_kernel<T, MB, PF><<<(unsigned)((lanes + T - 1) / T), T, 0, st>>>(            \
      nnz, 4.0 * par.mu, blkOrder, blkAdjPtr, blkAdj, elemP, Val)
    const int pf = (tune >> 15) & 3;   // bits 15, 16: record prefetch (1 = L1, 2 = L2)
    if (pf == 1) { if (tune & 8192) GL(128, 10, 1); else GL(128, 1, 1); }
    else if (pf == 2) { if (tune & 8192) GL(128, 10, 2); else GL(128, 1, 2); }
    else if (tune & 4096) {
      if (tune & 2048) GL(256, 8, 0); else GL(128, 16, 0);
    } else if (tune & 8192) {
      if (tune & 2048) GL(256, 5, 0); else GL(128, 10, 0);
    } else {
      if (tune & 2048) GL(256, 1, 0); else GL(128, 1, 0);
    }
#undef GL
    parts &= ~2;
  }
  if (parts & 2) {
    count_launch();
    const size_t lanes = (size_t)nnz * 8;
#define GV(U2, T, MB)                                                                          \
  fluid_gather_val_t_kernel<U2, T, MB><<<(unsigned)((lanes + T - 1) / T), T, 0, st>>>(           \
      nnz, 4.0 * par.mu, blkOrder, blkAdjPtr, blkAdj, elemP, Val)
    const int knob = (tune >> 2) & 7;   // bits 2..4: 1 = U2, 2 = 128 threads, 4 = 32-register cap
    const int pf = (tune >> 15) & 3;    // bits 15, 16: record prefetch (1 = L1, 2 = L2)
#define GVP(PF)                                                                                  \
  fluid_gather_val_t_kernel<false, 128, 1, PF><<<(unsigned)((lanes + 127) / 128), 128, 0, st>>>(   \
      nnz, 4.0 * par.mu, blkOrder, blkAdjPtr, blkAdj, elemP, Val)
    if (pf == 1) GVP(1);
    else if (pf == 2) GVP(2);
    else if (tune & 2)
      fluid_gather_val2_kernel<<<(unsigned)((nnz + GCH - 1) / GCH), 256, 0, st>>>(
          nnz, 4.0 * par.mu, blkOrder, blkAdjPtr, blkAdj, elemP, Val);
    else if (knob == 1) GV(true, 256, 1);
    else if (knob == 2) GV(false, 128, 1);
    else if (knob == 3) GV(true, 128, 1);
    else if (knob == 4) GV(false, 256, 8);
    else if (knob == 5) GV(true, 256, 8);
    else if (knob == 6) GV(false, 128, 16);
    else if (knob == 7) GV(true, 128, 16);
    else
      fluid_gather_val_kernel<<<(unsigned)(((size_t)nnz * 8 + 255) / 256), 256, 0, st>>>(
          nnz, 4.0 * par.mu, blkOrder, blkAdjPtr, blkAdj, elemP, Val);
  }
  if (parts & 4) {
    count_launch();
    fluid_gather_r_kernel<<<(nNo * 4 + 255) / 256, 256, 0, st>>>(nNo, nodeAdjPtr, nodeAdj, elemP,
                                                                 R);
  }
}

void launch_fluid_gather(cudaStream_t st, const FluidPar &par, int nEl, int nNo, int nnz,
                         const int *ien, const double *x, const double *Ag, const double *Yg,
                         const double *Bf, double *elemP, const int *blkOrder,
                         const int *blkAdjPtr, const int *blkAdj, const int *nodeAdjPtr,
                         const int *nodeAdj, double *R, double *Val, int *badJac,
                         const int *rowPtr, const int *nodeSlots, int maxRow, PairLists pairs) {
  launch_fluid_gather_parts(st, 7, par, nEl, nNo, nnz, ien, x, Ag, Yg, Bf, elemP, blkOrder,
                            blkAdjPtr, blkAdj, nodeAdjPtr, nodeAdj, R, Val, badJac, asm_tune(),
                            rowPtr, nodeSlots, maxRow, pairs);
}

// ---------------------------------------------------------------------------
// CONSTRUCT_HEATS / HEATS3D, S/HEATS.f:60-108,116-159 (dof = 1): one thread per
// element; lK(a,b) = wl (amd sum_g N_a N_b + 4 nu Nx_a.Nx_b)
template <bool ATOMIC>
__global__ void __launch_bounds__(128) heat_asm_kernel(HeatPar par, int n, int e0,
                                                       const int *__restrict__ elems,
                                                       const int *__restrict__ ien,
                                                       const int *__restrict__ edest,
                                                       const double *__restrict__ x,
                                                       const double *__restrict__ Ag,
                                                       const double *__restrict__ Yg,
                                                       double *__restrict__ R,
                                                       double *__restrict__ Val,
                                                       int *__restrict__ badJac) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int e = elems ? elems[e0 + idx] : e0 + idx;
  Tet4Tab tab;
  tet4_tab(tab);
  int nd[4];
  {
    const int4 v = __ldg((const int4 *)ien + e);
    nd[0] = v.x; nd[1] = v.y; nd[2] = v.z; nd[3] = v.w;
  }
  double xl[4][3], al[4], yl[4];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const double *xp = x + (size_t)nd[a] * 3;
    xl[a][0] = __ldg(xp); xl[a][1] = __ldg(xp + 1); xl[a][2] = __ldg(xp + 2);
    al[a] = __ldg(Ag + nd[a]);
    yl[a] = __ldg(Yg + nd[a]);
  }
  double Nx[4][3], Jac, ks[3][3];
  gnn_tet4(xl, Nx, Jac, ks);
  if (iszero1(Jac)) atomicAdd(badJac, 1);
  const double T1 = par.af * par.gam * par.dt;
  const double amd = par.am * par.rho / T1;
  const double w = (1.0 / 24.0) * Jac;
  const double wl = w * T1;
  double Tx[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int i = 0; i < 3; i++) Tx[i] = Tx[i] + Nx[a][i] * yl[a];
  double lR[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int g = 0; g < 4; g++) {
    double Td = -par.s;
#pragma unroll
    for (int a = 0; a < 4; a++) Td = Td + tab.N[g][a] * al[a];
    Td = Td * par.rho;
#pragma unroll
    for (int a = 0; a < 4; a++)
      lR[a] = lR[a] + w * (tab.N[g][a] * Td +
                           (Nx[a][0] * Tx[0] + Nx[a][1] * Tx[1] + Nx[a][2] * Tx[2]) * par.nu);
  }
  const int4 *dp = (const int4 *)(edest + (size_t)e * 16);
#pragma unroll
  for (int a = 0; a < 4; a++) {
    if (ATOMIC) red_add(R + nd[a], lR[a]);
    else R[nd[a]] += lR[a];
    const int4 d4 = __ldg(dp + a);
    const int dst[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
    for (int b = 0; b < 4; b++) {
      double v = 0.0;
#pragma unroll
      for (int g = 0; g < 4; g++)
        v = v + wl * (tab.N[g][a] * tab.N[g][b] * amd +
                      par.nu * (Nx[a][0] * Nx[b][0] + Nx[a][1] * Nx[b][1] + Nx[a][2] * Nx[b][2]));
      if (ATOMIC) red_add(Val + dst[b], v);
      else Val[dst[b]] += v;
    }
  }
}

// Gather variant for heatS (dof = 1), same owner-computes scheme as the fluid one: kernel A writes
// a 20-double record per element (lK(4,4) then lR(4)), kernel B gives every scalar Val entry to
// one thread that sums the (element, a, b) triples around its edge in ascending element order,
// kernel C does the same per node for R.  Deterministic, no atomics, no zero fill.
static constexpr int HREC = 20;
__global__ void __launch_bounds__(128) heat_record_kernel(HeatPar par, int n,
                                                          const int *__restrict__ ien,
                                                          const double *__restrict__ x,
                                                          const double *__restrict__ Ag,
                                                          const double *__restrict__ Yg,
                                                          double *__restrict__ rec,
                                                          int *__restrict__ badJac) {
  __shared__ double sm[HREC][129];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) {
    Tet4Tab tab;
    tet4_tab(tab);
    int nd[4];
    {
      const int4 v = __ldg((const int4 *)ien + e);
      nd[0] = v.x; nd[1] = v.y; nd[2] = v.z; nd[3] = v.w;
    }
    double xl[4][3], al[4], yl[4];
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double *xp = x + (size_t)nd[a] * 3;
      xl[a][0] = __ldg(xp); xl[a][1] = __ldg(xp + 1); xl[a][2] = __ldg(xp + 2);
      al[a] = __ldg(Ag + nd[a]);
      yl[a] = __ldg(Yg + nd[a]);
    }
    double Nx[4][3], Jac, ks[3][3];
    gnn_tet4(xl, Nx, Jac, ks);
    if (iszero1(Jac)) atomicAdd(badJac, 1);
    const double T1 = par.af * par.gam * par.dt;
    const double amd = par.am * par.rho / T1;
    const double w = (1.0 / 24.0) * Jac;
    const double wl = w * T1;
    double Tx[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int i = 0; i < 3; i++) Tx[i] = Tx[i] + Nx[a][i] * yl[a];
    double lR[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int g = 0; g < 4; g++) {
      double Td = -par.s;
#pragma unroll
      for (int a = 0; a < 4; a++) Td = Td + tab.N[g][a] * al[a];
      Td = Td * par.rho;
#pragma unroll
      for (int a = 0; a < 4; a++)
        lR[a] = lR[a] + w * (tab.N[g][a] * Td +
                             (Nx[a][0] * Tx[0] + Nx[a][1] * Tx[1] + Nx[a][2] * Tx[2]) * par.nu);
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
      for (int b = 0; b < 4; b++) {
        double v = 0.0;
#pragma unroll
        for (int g = 0; g < 4; g++)
          v = v + wl * (tab.N[g][a] * tab.N[g][b] * amd +
                        par.nu * (Nx[a][0] * Nx[b][0] + Nx[a][1] * Nx[b][1] + Nx[a][2] * Nx[b][2]));
        sm[a * 4 + b][threadIdx.x] = v;
      }
      sm[16 + a][threadIdx.x] = lR[a];
    }
  }
  __syncthreads();
  const int nHere = min(128, n - (int)blockIdx.x * 128);
  double *out = rec + (size_t)blockIdx.x * 128 * HREC;
  for (int t = threadIdx.x; t < nHere * HREC; t += 128) {
    const int s = t / HREC, f = t - s * HREC;
    out[t] = sm[f][s];
  }
}
__global__ void __launch_bounds__(256) heat_gather_val_kernel(int nnz, const int *__restrict__ adjPtr,
                                                              const int *__restrict__ adj,
                                                              const double *__restrict__ rec,
                                                              double *__restrict__ Val) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  double acc = 0.0;
  for (int k = __ldg(adjPtr + p); k < __ldg(adjPtr + p + 1); k++) {
    const int pk = __ldg(adj + k);
    acc += __ldg(rec + (size_t)(pk >> 4) * HREC + (pk & 15));
  }
  Val[p] = acc;
}
__global__ void __launch_bounds__(256) heat_gather_r_kernel(int nNo, const int *__restrict__ adjPtr,
                                                            const int *__restrict__ adj,
                                                            const double *__restrict__ rec,
                                                            double *__restrict__ R) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nNo) return;
  double acc = 0.0;
  for (int k = __ldg(adjPtr + t); k < __ldg(adjPtr + t + 1); k++) {
    const int pk = __ldg(adj + k);
    acc += __ldg(rec + (size_t)(pk >> 2) * HREC + 16 + (pk & 3));
  }
  R[t] = acc;
}
void launch_heat_gather(cudaStream_t st, const HeatPar &par, int nEl, int nNo, int nnz,
                        const int *ien, const double *x, const double *Ag, const double *Yg,
                        double *rec, const int *blkAdjPtr, const int *blkAdj, const int *nodeAdjPtr,
                        const int *nodeAdj, double *R, double *Val, int *badJac) {
  if (nEl <= 0) return;
  count_launch(3);
  heat_record_kernel<<<(nEl + 127) / 128, 128, 0, st>>>(par, nEl, ien, x, Ag, Yg, rec, badJac);
  heat_gather_val_kernel<<<(nnz + 255) / 256, 256, 0, st>>>(nnz, blkAdjPtr, blkAdj, rec, Val);
  heat_gather_r_kernel<<<(nNo + 255) / 256, 256, 0, st>>>(nNo, nodeAdjPtr, nodeAdj, rec, R);
}

void launch_heat_asm(cudaStream_t st, const HeatPar &par, int n, int e0, const int *elems,
                     const int *ien, const int *edest, const double *x, const double *Ag,
                     const double *Yg, double *R, double *Val, int atomic, int *badJac) {
  if (n <= 0) return;
  count_launch();
  const int blocks = (n + 127) / 128;
  if (atomic)
    heat_asm_kernel<true><<<blocks, 128, 0, st>>>(par, n, e0, elems, ien, edest, x, Ag, Yg, R, Val,
                                                  badJac);
  else
    heat_asm_kernel<false><<<blocks, 128, 0, st>>>(par, n, e0, elems, ien, edest, x, Ag, Yg, R,
                                                   Val, badJac);
}

// ---------------------------------------------------------------------------
// per-element scatter map: one binary search per (a,b) done ONCE at setup instead
// of per assembly (S/LHSA.f:282-292).  Rows keep svFSI's ascending ORIGINAL column
// order, which is not ascending in reordered ids, so the search runs on a key
// array `colKey` (original ids) -- here col holds reordered ids and the rows are
// short (<= ~30), so a linear scan is used.
__global__ void build_edest_kernel(int nEl, const int *__restrict__ ien,
                                   const int *__restrict__ rowPtr, const int *__restrict__ col,
                                   int *__restrict__ edest, int *__restrict__ missing) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nEl * 16) return;
  const int e = (int)(t >> 4), a = (int)((t >> 2) & 3), b = (int)(t & 3);
  const int row = ien[(size_t)e * 4 + a], c = ien[(size_t)e * 4 + b];
  int p = -1;
  for (int j = rowPtr[row]; j < rowPtr[row + 1]; j++)
    if (col[j] == c) { p = j; break; }
  if (p < 0) atomicAdd(missing, 1);
  edest[t] = p;
}
void launch_build_edest(cudaStream_t st, int nEl, const int *ien, const int *rowPtr,
                        const int *col, int *edest, int *missing) {
  if (nEl <= 0) return;
  count_launch();
  size_t tot = (size_t)nEl * 16;
  build_edest_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(nEl, ien, rowPtr, col, edest,
                                                                    missing);
}

}  // namespace svfsi
