// asm_gather5.cu -- default kernels of the owner-computes (gather) fluid assembly, second generation
// (CONSTRUCT_FLUID + FLUID3D_M/C + DOASSEM, S/FLUID.f:40-190, S/LHSA.f:266-298).
//
// What bounds the first-generation block-owner gather (asm_kernels.cu fluid_gather_quad_kernel): the L1TEX
// pipe spends ~1 cycle per load instruction plus ~2.07 per ADDITIONAL 128-byte line the instruction touches
// (B300_MICROARCH.md).  Eight 4-lane groups per warp read eight different element records, so each of the four
// loads of a contribution touches eight lines: 62 cycles per warp step of eight contributions, 4.4 ms at 10M
// tets -- the measured 4.0 ms (profiles/r01_ncu_asm_summary.md: L1 wavefronts 81%, everything else idle).
// Two changes cut the lines per load instruction:
//  * record v5, 64 doubles = FOUR 128-byte lines (was 80 doubles), arranged so that a contribution (a, b)
//    needs exactly one 32-byte load from each of three lines:
//      [ 0..15] NX : per node n: Nx_n(1..3), R2_n          R2 = rho sum_g tauM (uNx + amd N)
//      [16..31] D  : D_ab = 4 mu Nx_a.Nx_b + A_ab          (E_ab = sum tauM * Nx_a.Nx_b is recomputed)
//      [32..47] LR : lR(i, a)
//      [48..63] SC : per node a: C2_a, sum tauC, wl, sum tauM     C2 = rho sum_g tauM uaNx
//  * PAIRED processing order: the blocks (r,c) and (c,r) have the same element list (the elements around the
//    edge r-c, ascending) with a and b exchanged, so when they sit in ADJACENT 4-lane groups of a warp the two
//    groups read the SAME lines in every load instruction -- four distinct records per warp instruction
//    instead of eight.  Diagonal blocks (list = the elements around the node) come after all pairs; their
//    groups also sum lR(:,a) and write R (the separate residual-gather kernel is gone) and skip the second
//    node load (a == b).
// Every block still receives its contributions in ascending element order starting from 0.0: deterministic,
// bitwise repeatable, the accumulation order of the reference's element loop.
#include <cuda_runtime.h>
#include <float.h>

#include "asm_elem.h"
#include "ctx.h"
#include "kernels.h"

namespace svfsi {

static constexpr int RECQ = 64;     // doubles per record

struct __align__(32) dq4 { double x, y, z, w; };
__device__ __forceinline__ dq4 ldq_nc(const double *p) {
  dq4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------
// Kernel A: one thread per element, the four Gauss points reduced to the record (fluid_elem_compute,
// asm_elem.h), staged through shared memory as 32 double2 pairs so that the record leaves coalesced.
static constexpr int NPQ = RECQ / 2;
__device__ __forceinline__ void fluid_elem_store5(const FluidPar &par, const ElemAcc &acc, double2 *rec2) {
  const double rho = par.rho, mu = par.mu;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    rec2[(a * 2 + 0) * NEP] = make_double2(acc.Nx[a][0], acc.Nx[a][1]);
    rec2[(a * 2 + 1) * NEP] = make_double2(acc.Nx[a][2], rho * acc.r2[a]);
  }
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
    rec2[(8 + a * 2 + 0) * NEP] = make_double2(4.0 * (mu * nn[a][0]) + acc.A[a][0], 4.0 * (mu * nn[a][1]) + acc.A[a][1]);
    rec2[(8 + a * 2 + 1) * NEP] = make_double2(4.0 * (mu * nn[a][2]) + acc.A[a][2], 4.0 * (mu * nn[a][3]) + acc.A[a][3]);
    double lr[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
      lr[i] = acc.wr * acc.sNrV[a][i] +
              acc.w * (acc.Nx[a][0] * acc.sRM[0][i] + acc.Nx[a][1] * acc.sRM[1][i] + acc.Nx[a][2] * acc.sRM[2][i]);
    rec2[(16 + a * 2 + 0) * NEP] = make_double2(lr[0], lr[1]);
    rec2[(16 + a * 2 + 1) * NEP] = make_double2(lr[2], acc.w * acc.lR4[a]);
    rec2[(24 + a * 2 + 0) * NEP] = make_double2(rho * acc.c2[a], acc.sTC);
    rec2[(24 + a * 2 + 1) * NEP] = make_double2(acc.wl, acc.sTM);
  }
}

__global__ void __launch_bounds__(NE) fluid_record5_kernel(FluidPar par, int e0, int e1,
                                                           const int *__restrict__ ien,
                                                           const double *__restrict__ x,
                                                           const double *__restrict__ Ag,
                                                           const double *__restrict__ Yg,
                                                           const double *__restrict__ Bf,
                                                           double *__restrict__ recs, unsigned ringMask,
                                                           int *__restrict__ badJac) {
  extern __shared__ double2 smq[];  // [NPQ][NEP]
  const int slot = threadIdx.x;
  const int eb = e0 + blockIdx.x * NE;
  const int e = eb + slot;
  int nodes[4];
  if (e < e1) {
    ElemAcc acc;
    fluid_elem_compute<true>(par, e, ien, x, Ag, Yg, Bf, acc, nodes, badJac);
    fluid_elem_store5(par, acc, smq + slot);
  }
  __syncthreads();
  const int nHere = min(NE, e1 - eb);
  // thread t copies pairs t, t + NE, ...: NE = 4 * NPQ, so (slot, pair) advance by (4, 0)
  const int k = threadIdx.x & (NPQ - 1);
  for (int s = threadIdx.x / NPQ; s < nHere; s += NE / NPQ) {
    // the record of element eb + s lives at ring slot (eb + s) & ringMask (ringMask = ~0u: no ring)
    double2 *out = (double2 *)(recs + (size_t)((unsigned)(eb + s) & ringMask) * RECQ);
    __stcg(out + k, smq[k * NEP + s]);
  }
}

// ---------------------------------------------------------------------------
// Kernels B + C: four lanes per block, lane r = row r of the 4x4 block, transposed blocks in adjacent groups.
// desc[g] = (block, list begin, list end, row + 1 for a diagonal block | 0).
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fluid_gather_quad5_kernel(
    int g0, int g1, double mu4, const int4 *__restrict__ desc, const int *__restrict__ adj,
    const double *__restrict__ recs, unsigned ringMask, double *__restrict__ Val, double *__restrict__ R) {
  const int lane = threadIdx.x & 31, r = lane & 3;
  const unsigned gmask = 0xFu << (lane & 28);
  const int g = g0 + (int)((blockIdx.x * (unsigned)THREADS + threadIdx.x) >> 2);
  if (g >= g1) return;   // whole 4-lane groups leave together
  const int4 d = __ldg(desc + g);
  const int p = d.x, s = d.y, e = d.z;
  const bool isDiag = d.w != 0;
  const bool row3 = (r == 3);
  // lane-constant coefficients (S/FLUID.f:482-557 momentum rows, :1052-1081 continuity row)
  const double P = row3 ? 1.0 : mu4;       // coefficient of Nx_a(j) * bi, j < 3
  const double P3 = row3 ? 0.0 : 1.0;      // last column: C2_a * Nx_b(i) on a momentum row
  const double c3 = row3 ? 0.0 : -1.0;     // last column: -Nx_a(i) (sum_g N_b = 1) on a momentum row
  const double k0 = (r == 0) ? 1.0 : 0.0, k1 = (r == 1) ? 1.0 : 0.0, k2 = (r == 2) ? 1.0 : 0.0,
               k3 = row3 ? 1.0 : 0.0;      // which entry of the lane's row takes the D / E term
  const bool i0 = (r == 0), i1 = (r == 1);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, lr = 0.0;
  int cq = (s + r < e) ? __ldg(adj + s + r) : 0;
  for (int base = s; base < e; base += 4) {
    // the next slice of the list is in flight while this one is processed
    const int nxt = base + 4 + r;
    const int cqn = (nxt < e) ? __ldg(adj + nxt) : 0;
    const int cnt = min(4, e - base);
#pragma unroll 1
    for (int k = 0; k < cnt; k++) {
      const unsigned pk = (unsigned)__shfl_sync(gmask, cq, k, 4);
      const double *rec = recs + (size_t)((pk >> 4) & ringMask) * RECQ;
      const unsigned oa = (pk & 12u), ob = (pk & 3u) << 2;       // 4 a, 4 b
      const dq4 A = ldq_nc(rec + oa);                             // Nx_a, R2_a
      const dq4 S = ldq_nc(rec + 48 + oa);                        // C2_a, sum tauC, wl, sum tauM
      const double D = __ldg(rec + 16 + oa + (pk & 3u));          // D_ab
      dq4 B = A;
      if (!isDiag) B = ldq_nc(rec + ob);                          // Nx_b, R2_b
      else lr += __ldg(rec + 32 + oa + r);                        // lR(r, a)
      // Nx_i of a / of b (R2_b on the continuity row): selects, not loads -- the kernel is bound by the L1
      // pipe (every extra load instruction costs as much as one of the four above), not by issue slots
      const double ai = i0 ? A.x : (i1 ? A.y : A.z);
      const double bi = row3 ? B.w : (i0 ? B.x : (i1 ? B.y : B.z));
      const double nn = fma(A.z, B.z, fma(A.y, B.y, A.x * B.x));
      const double de = row3 ? S.w * nn : D;                      // E_ab | D_ab
      const double u = row3 ? 1.0 : S.y * ai;
      double s0 = fma(u, B.x, k0 * de);
      double s1 = fma(u, B.y, k1 * de);
      double s2 = fma(u, B.z, k2 * de);
      double s3 = fma(ai, c3, k3 * de);
      const double pb = P * bi;
      s0 = fma(pb, A.x, s0);
      s1 = fma(pb, A.y, s1);
      s2 = fma(pb, A.z, s2);
      s3 = fma(P3 * bi, S.x, s3);
      a0 = fma(S.z, s0, a0);
      a1 = fma(S.z, s1, a1);
      a2 = fma(S.z, s2, a2);
      a3 = fma(S.z, s3, a3);
    }
    cq = cqn;
  }
  double2 *out = (double2 *)(Val + (size_t)p * 16 + r * 4);
  __stcs(out, make_double2(a0, a1));
  __stcs(out + 1, make_double2(a2, a3));
  if (isDiag) R[(size_t)(d.w - 1) * 4 + r] = lr;
}

static void g5_attr_once() {
  static bool done = false;
  if (done) return;
  cudaFuncSetAttribute(fluid_record5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)((size_t)NPQ * NEP * sizeof(double2)));
  done = true;
}

// parts: 1 = records of elements [e0, e1), 2 = descriptor entries [g0, g1)
void launch_fluid_gather5(cudaStream_t st, int parts, const FluidPar &par, int e0, int e1, int g0, int g1,
                          const int *ien, const double *x, const double *Ag, const double *Yg, const double *Bf,
                          double *recs, unsigned ringMask, const int4 *desc, const int *adj, double *R,
                          double *Val, int *badJac, int knob) {
  g5_attr_once();
  if ((parts & 1) && e1 > e0) {
    count_launch();
    const size_t smem = (size_t)NPQ * NEP * sizeof(double2);
    fluid_record5_kernel<<<(e1 - e0 + NE - 1) / NE, NE, smem, st>>>(par, e0, e1, ien, x, Ag, Yg, Bf, recs, ringMask,
                                                                    badJac);
  }
  if ((parts & 2) && g1 > g0) {
    count_launch();
    const size_t lanes = (size_t)(g1 - g0) * 4;
    if (knob & 1)
      fluid_gather_quad5_kernel<256, 4><<<(unsigned)((lanes + 255) / 256), 256, 0, st>>>(
          g0, g1, 4.0 * par.mu, desc, adj, recs, ringMask, Val, R);
    else
      fluid_gather_quad5_kernel<128, 8><<<(unsigned)((lanes + 127) / 128), 128, 0, st>>>(
          g0, g1, 4.0 * par.mu, desc, adj, recs, ringMask, Val, R);
  }
}

}  // namespace svfsi
