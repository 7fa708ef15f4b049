// kernels.h -- launchers of the hand-written sm_100a kernels (la_kernels.cu,
// asm_kernels.cu).  All launch on the given stream, return nothing, and bump the
// context's launch counter.  `done` (may be NULL) is a device flag: when it is
// non-zero the kernel returns immediately (used to run the Krylov loop ahead of
// the host without a sync per iteration).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svfsi {

struct FluidPar {
  double rho, mu, f[3], dt, af, am, gam;
};
struct HeatPar {
  double nu, s, rho, dt, af, am, gam;
};

// Control block of one GMRES / CG solve, resident on the device.
struct KrylovCtl {
  int done;     // stop flag: later kernels of the current cycle become no-ops
  int suc;      // ls%suc
  int ilast;    // number of Krylov vectors built in the current cycle
  int itr;      // ls%itr (SpMV count) accumulated on the device
  double eps;   // absolute stopping threshold
  double inv;   // 1/h(i+1,i) of the current column
  double fNorm, iNorm, dBref;
  double scal[8];  // scratch scalars (alpha, err, errO, ...)
};

// ---------------- block-CSR SpMV (L/SPARMUL.f) ----------------
// rows [r0,r1) of KU = K*U.  kind: 0 VV (dof x dof blocks), 1 VS (1 x dof), 2 SV (dof x 1),
// 3 SS.  dof in {1,2,3,4}.
void launch_spmv(cudaStream_t st, int kind, int dof, int r0, int r1, const int *rowPtr,
                 const int *col, const double *K, const double *U, double *KU, const int *done);
// rows [r0,r1) and [r2,r3) in ONE launch (the two boundary slabs of the reordered numbering)
void launch_spmv2(cudaStream_t st, int kind, int dof, int r0, int r1, int r2, int r3,
                  const int *rowPtr, const int *col, const double *K, const double *U, double *KU,
                  const int *done);

// K <- (W_row K) W_col and KU = K U in one pass over Val (dof = 4, single rank): the first product of a solve
void launch_spmv_vv4_scale(cudaStream_t st, int nNo, const int *rowPtr, const int *col, double *K,
                           const double *W, const double *U, double *KU);

// ---------------- halo (L/INCOMMU.f) ----------------
void launch_pack(cudaStream_t st, int dof, int nShared, const int *packIdx, const double *R,
                 double *sbuf, const int *done);
void launch_unpack_add(cudaStream_t st, int dof, int nUniq, const int *uniqNode,
                       const int *uniqPtr, const int *uniqSlot, const double *rbuf, double *R,
                       const int *done);

// peer-memory versions (NVLink stores into the neighbour's IPC-mapped receive buffer + flag)
struct P2PDev {
  char **peer;          // [nranks] arena bases
  size_t offMail, offHalo, offMailLL;   // offMailLL: flag-in-data mailboxes of the fused column kernel
  int haloCap, rank, nranks;
  int *errDev;              // sticky "a flag wait timed out" word in device memory (Ctx::d_flag + 2)
  volatile int *errHost;    // the same in mapped pinned host memory (Ctx::h_status[0])
  long long timeoutNs;      // bound of every in-kernel flag wait
};
// SpMV with the halo send fused into the kernel (la_kernels.cu: spmv_*_fused_kernel)
struct SpmvFuse {
  int shnNo, mynNo, nNo;   // row ranges of the reordered numbering (L/LHS.f:134-165)
  int nBnd, bndCtas;       // boundary rows = shnNo + (nNo - mynNo); CTAs that hold them
  const int *sendPtr;      // [nBnd+1] CSR over boundary rows -> destinations
  const int *sendRank;     // peer rank of each destination
  const int *sendOff;      // node offset inside that peer's receive slot
  const int *nbrRank;
  int nNbr;
  P2PDev pd;
  int seq;
  unsigned int *counter;
  unsigned long long *trace;   // SVFSI_TRACE_FILE: 8 time stamps of this column (NULL: off)
};
// scaleW != NULL (kind 0, dof 4 only): K <- (W_row K) W_col on the way (K is written)
void launch_spmv_fused(cudaStream_t st, int kind, int dof, SpmvFuse f, const int *rowPtr,
                       const int *col, const double *K, const double *U, double *KU,
                       const int *done, const double *scaleW = nullptr);
void launch_halo_send(cudaStream_t st, const P2PDev &pd, int dof, int nShared, int nNbr,
                      const int *packIdx, const int *slotNbr, const int *nbrRank, const int *nbrOff,
                      const int *nbrPeerOff, const double *R, int seq, unsigned int *counter);
void launch_halo_recv_add(cudaStream_t st, const P2PDev &pd, int dof, int nNbr, const int *nbrRank,
                          int nUniq, const int *uniqNode, const int *uniqPtr, const int *uniqSlot,
                          double *R, int seq);
// out[j] = sum over ranks (rank order) of (partial ? sum_b partial[j*nblk+b] : out[j]), j < k <= kArMax
// arguments of the GMRES column step (Givens / Hessenberg / stop test, L/GMRES.f:342-366) when it
// is fused into the tail of a reduction kernel; ctl == NULL: no column step
struct ColArgs {
  KrylovCtl *ctl;
  int i, sD;
  double *h, *c, *s, *err, *coef;
  volatile int *pubFlag, *pubProgress;
  int seq;
};
void launch_p2p_allreduce(cudaStream_t st, const P2PDev &pd, const double *partial, int nblk, int k,
                          double *out, int seq, const ColArgs *col = nullptr);

// ---- one kernel per Gram-Schmidt column: [halo receive] + multi-dot + block sums + [all-reduce] + [column]
// The halo sum that FSILS_SPARMUL* ends with (L/SPARMUL.f:130 -> L/INCOMMU.f:91-151) only touches the
// rows shared with other ranks; when the SpMV's result goes straight into inner products
// (L/GMRES.f:328-339) the first CTAs of the multi-dot grid wait for the neighbours' flags, add what
// arrived to THEIR rows and then take those rows' part of the inner products, while every other CTA
// streams the interior rows at once.  The last CTA to finish sums the per-CTA partials, exchanges them
// with the peers (stores into their mailboxes, rank-ordered sum: bit-identical on all ranks), and runs
// the Givens / Hessenberg column step.  Replaces halo_recv_add + multidot + reduce/all-reduce(+column):
// three launches and two rank synchronisations per column instead of five and two.
struct HaloRecv {
  int on;                 // 0: the vector is already halo-consistent
  int nNbr, dof, seq;     // neighbours to wait for, dofs per node of w, halo sequence number
  int shnNo, mynNo, nUniq;
  const int *nbrRank, *uniqNode, *uniqPtr, *uniqSlot;
};
struct DotTail {
  int nranks;             // > 1: all-reduce over the peers' mailboxes (pd), sequence number arSeq
  P2PDev pd;
  int arSeq;
  double *out;            // [k] the reduced inner products
  unsigned int *counter;  // CTA ticket (zero between launches)
  ColArgs col;            // col.ctl == NULL: no column step
  unsigned long long *trace;   // SVFSI_TRACE_FILE: 8 time stamps of this column (NULL: off)
};
void launch_multidot_fused(cudaStream_t st, const double *U, size_t stride, double *w, size_t nOwned,
                           int k, double *partial, const int *done, const HaloRecv &hr,
                           const DotTail &tail);
// single rank: out[j] = sum_b partial[j*nblk+b] and the column step, one kernel (k <= kArMax)
void launch_reduce_column(cudaStream_t st, const double *partial, int nblk, int k, double *out,
                          const ColArgs &col);

// ---------------- vectors ----------------
// partial[j*nblk + b] = sum over block b of U_j . w, j < k; U_j = U + j*stride. n doubles.
int multidot_nblk();
void launch_multidot(cudaStream_t st, const double *U, size_t stride, const double *w, size_t n,
                     int k, double *partial, const int *done);
// out[j] = sum_b partial[j*nblk+b]
void launch_reduce_partials(cudaStream_t st, const double *partial, int k, double *out,
                            const int *done);
// w = (w - sum_{j<k} coef[j]*U_j) * (*scale)   (coef, scale on device; scale may be NULL = 1)
void launch_multi_axpy_scale(cudaStream_t st, const double *U, size_t stride, double *w, size_t n,
                             int k, const double *coef, const double *scale, const int *done);
// X += sum_{j<*kdev} y[j]*U_j   (k read from the device)
void launch_multi_axpy_acc(cudaStream_t st, const double *U, size_t stride, double *X, size_t n,
                           const int *kdev, int kmax, const double *y);
// elementwise helpers: op codes
enum VecOp {
  VOP_COPY = 0,       // a = b
  VOP_SUB_FROM = 1,   // a = b - a
  VOP_SCALE_DEV = 2,  // a = a * (*s)            (s on device)
  VOP_DIV_DEV = 3,    // a = a / (*s)
  VOP_AXPY_DEV = 4,   // a = a + (*s) * b
  VOP_AXMY_DEV = 5,   // a = a - (*s) * b
  VOP_MUL = 6,        // a = a * b               (elementwise)
  VOP_ZERO = 7,
  VOP_AXPY = 8,       // a = a + sh * b          (host scalar)
  VOP_SCALE = 9,      // a = sh * a
  VOP_XPBY_DEV = 10,  // a = b + (*s) * a        (CG direction update)
  VOP_SUB = 11        // a = b - c
};
void launch_vecop(cudaStream_t st, int op, double *a, const double *b, const double *c, size_t n,
                  const double *sdev, double shost, const int *done);
// interleave / de-interleave momentum and continuity parts (NSSOLVER Rm/Rc split)
void launch_split_mc(cudaStream_t st, int nNo, int dof, const double *R, double *Rm, double *Rc);
void launch_join_mc(cudaStream_t st, int nNo, int dof, const double *Rm, const double *Rc,
                    double *R);

// ---------------- permutation between svFSI and FSILS layouts ----------------
// dst[perm[a]][m] = src[a][m]  (scatter) ; dst[a][m] = src[perm[a]][m] (gather)
void launch_perm_scatter(cudaStream_t st, int n, int m, const int *perm, const double *src,
                         double *dst);
void launch_perm_gather(cudaStream_t st, int n, int m, const int *perm, const double *src,
                        double *dst);

// ---------------- Jacobi preconditioner (L/PRECOND.f:50-145) ----------------
void launch_diag_extract(cudaStream_t st, int nNo, int dof, const int *diag, const double *Val,
                         double *W);
void launch_w_finalize(cudaStream_t st, size_t n, double *W);
void launch_w_dirichlet(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob,
                        const double *val, double *W);
void launch_scale_val(cudaStream_t st, int nnz, int dof, const int *rowOf, const int *col,
                      const double *W, double *Val);
// Val = (Wrow_row * Val) * Wcol_col: PREMUL(Wrow) then POSMUL(Wcol) (L/PRECOND.f:372-489) fused
void launch_scale_val2(cudaStream_t st, int nnz, int dof, const int *rowOf, const int *col,
                       const double *Wrow, const double *Wcol, double *Val);
void launch_face_valM(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob,
                      const double *val, const double *W, double *valM);
// ADDBCMUL (L/ADDBCMUL.f): S = sum_a sum_i valM(i,a) X(i,glob(a)) over nodes with glob < ownedLimit
void launch_face_dot(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob,
                     const double *valM, const double *X, int ownedLimit, double *S,
                     const int *done);
// Y(i,glob(a)) += valM(i,a) * coef * (*S)
void launch_face_axpy(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob,
                      const double *valM, double coef, const double *S, double *Y,
                      const int *done);
// a face that lives on one rank (no all-reduce between the dot and the update): both spread over 32 CTAs;
// partial = 32 doubles of scratch
void launch_face_dotp(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob, const double *valM,
                      const double *X, int ownedLimit, double *partial, const int *done);
void launch_face_axpyp(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob, const double *valM,
                       double coef, const double *partial, double *S, double *Y, const int *done);
// nS = sum valM^2 over nodes with glob < ownedLimit
void launch_face_norm2(cudaStream_t st, int nFaceNo, int fdof, int nsd, const int *glob,
                       const double *valM, int ownedLimit, double *S);

// ---------------- Krylov scalar kernels ----------------
void launch_gmres_column(cudaStream_t st, KrylovCtl *ctl, int i, int sD, const double *hcol,
                         double *h, double *c, double *s, double *err, double *coef,
                         volatile int *pubFlag, volatile int *pubProgress, int seq);
void launch_gmres_backsub(cudaStream_t st, KrylovCtl *ctl, int sD, const double *h,
                          const double *err, double *y);
// DEPART (L/NSSOLVER.f:237-305)
void launch_depart(cudaStream_t st, int nnz, int nsd, const double *Val, double *mK, double *mG,
                   double *mD, double *mL);
void launch_gt(cudaStream_t st, int nnz, int nsd, const int *tpos, const double *mG, double *Gt);

// ---------------- element loops (S/FLUID.f, S/HEATS.f) ----------------
// variant: SVFSI_ASM_ATOMIC / COLORED.  elems == NULL -> elements [e0, e0+n); else elems[e0..e0+n)
void launch_fluid_asm(cudaStream_t st, const FluidPar &par, int n, int e0, const int *elems,
                      const int *ien, const int *edest, const double *x, const double *Ag,
                      const double *Yg, const double *Bf, double *R, double *Val, int atomic,
                      int *badJac);
void launch_heat_asm(cudaStream_t st, const HeatPar &par, int n, int e0, const int *elems,
                     const int *ien, const int *edest, const double *x, const double *Ag,
                     const double *Yg, double *R, double *Val, int atomic, int *badJac);
// pair-owner gather (asm_kernels.cu fluid_gather_pairs_kernel): the blocks (r,c) with c >= r in the
// length-sorted processing order, and for each the position of the transposed block (c,r)
// (== itself on the diagonal, -1 if the pattern has no transposed entry)
struct PairLists {
  const int *list = nullptr, *tpos = nullptr, *rowOf = nullptr;
  int n = 0;
  // block descriptors in processing order for the quad gather kernel: (block, list begin, list end, 0)
  // in ONE 16-byte load instead of the dependent pair blkOrder[g] -> adjPtr[p], adjPtr[p+1]
  const int4 *desc = nullptr;
};
int build_block_desc(cudaStream_t st, int nnz, const int *blkOrder, const int *adjPtr, int4 **desc,
                     const int *rowOf = nullptr, const int *col = nullptr);
int build_pair_lists(cudaStream_t st, int nnz, const int *blkOrder, const int *rowOf, const int *col,
                     const int *rowPtr, int **pairList, int **pairT, int *nPair);
// gather variant: element records + owner-computes accumulation (deterministic, no atomics)
void launch_fluid_gather(cudaStream_t st, const FluidPar &par, int nEl, int nNo, int nnz,
                         const int *ien, const double *x, const double *Ag, const double *Yg,
                         const double *Bf, double *elemP, const int *blkOrder,
                         const int *blkAdjPtr, const int *blkAdj, const int *nodeAdjPtr,
                         const int *nodeAdj, double *R, double *Val, int *badJac,
                         const int *rowPtr = nullptr, const int *nodeSlots = nullptr, int maxRow = 0,
                         PairLists pairs = PairLists());
// heatS gather variant; rec >= 20 doubles per element (the fluid record buffer is reused)
void launch_heat_gather(cudaStream_t st, const HeatPar &par, int nEl, int nNo, int nnz,
                        const int *ien, const double *x, const double *Ag, const double *Yg,
                        double *rec, const int *blkAdjPtr, const int *blkAdj, const int *nodeAdjPtr,
                        const int *nodeAdj, double *R, double *Val, int *badJac);
// the three kernels separately (parts: 1 records, 2 tangent gather, 4 residual gather) with an
// explicit kernel-variant mask (see asm_tune()); used by gpu_time_kernel_
void launch_fluid_gather_parts(cudaStream_t st, int parts, const FluidPar &par, int nEl, int nNo,
                               int nnz, const int *ien, const double *x, const double *Ag,
                               const double *Yg, const double *Bf, double *elemP,
                               const int *blkOrder, const int *blkAdjPtr, const int *blkAdj,
                               const int *nodeAdjPtr, const int *nodeAdj, double *R, double *Val,
                               int *badJac, int tune, const int *rowPtr = nullptr,
                               const int *nodeSlots = nullptr, int maxRow = 0,
                               PairLists pairs = PairLists());
// row-owner gather: positions of the four blocks of every (node, element) visit inside the row
void launch_build_node_slots(cudaStream_t st, int nNo, const int *rowPtr, const int *nodeAdjPtr,
                             const int *nodeAdj, const int *edest, int *slots);
int asm_tune();
// second-generation gather assembly (asm_gather5.cu): 512-byte records v5 + four lanes per block with the
// transposed blocks (r,c) / (c,r) in adjacent groups.  parts: 1 = records of elements [e0, e1), 2 = descriptor
// entries [g0, g1) of `desc` (build_paired_desc); ringMask: the record of element e lives at slot
// e & ringMask (~0u: one slot per element)
void launch_fluid_gather5(cudaStream_t st, int parts, const FluidPar &par, int e0, int e1, int g0, int g1,
                          const int *ien, const double *x, const double *Ag, const double *Yg, const double *Bf,
                          double *recs, unsigned ringMask, const int4 *desc, const int *adj, double *R,
                          double *Val, int *badJac, int knob);
// descriptors (block, list begin, list end, row + 1 | 0) in the paired processing order: all (r,c) / (c,r)
// pairs first (adjacent entries), then the diagonal blocks and any block without a transposed partner
int build_paired_desc(cudaStream_t st, int nPair, const int *pairList, const int *pairT, const int *adjPtr,
                      const int *rowOf, int nnz, int4 **desc, int *nOffEntries);
// adjacency lists for the gather variant (built once at gpu_mesh_create_)
int build_gather_adjacency(cudaStream_t st, int nEl, int nNo, int nnz, const int *ien,
                           const int *edest, int **blkAdjPtr, int **blkAdj, int **nodeAdjPtr,
                           int **nodeAdj, int **blkOrder);
// edest[e][a*4+b] = device block index of (row ien[e][a], col ien[e][b])
// (a,b) pairs that are not in the pattern get -1 and are counted into *missing
void launch_build_edest(cudaStream_t st, int nEl, const int *ien, const int *rowPtr,
                        const int *col, int *edest, int *missing);

int spmv_fused_quad_enabled();   // the fused SpMV + halo-send kernel runs 4 lanes per row
int set_spmv_quad(int on);   // SPARMULVV dof=4 kernel variant on the unfused path (la_kernels.cu)
int set_spmv_small(int mode); // small-shape SpMV kernel family (la_kernels.cu, SVFSI_SPMV_SMALL); returns the previous mode

void count_launch(int n = 1);

}  // namespace svfsi
