// la_kernels.cu -- hand-written sm_100a kernels of the FSILS linear-algebra core:
// block-CSR SpMV (L/SPARMUL.f), halo pack/unpack (L/INCOMMU.f), fused
// multi-dot / multi-axpy for classical Gram-Schmidt (L/GMRES.f:337-349,
// L/DOT.f, L/OMPLA.f), Jacobi scaling (L/PRECOND.f:50-145, 372-489), the
// coupled-BC rank-1 update (L/ADDBCMUL.f) and the scalar Hessenberg/Givens step.
//
// All of these are HBM-bandwidth bound (no tensor cores: nothing here is a dense
// contraction).  Design rules: 128-bit loads, every 32-byte sector fully used,
// grids sized as multiples of the 148 SMs, reductions by warp shuffles, and no
// host synchronisation inside the Krylov loop (kernels test a device flag).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "ctx.h"
#include "kernels.h"

namespace svfsi {

static constexpr int kSMs = 148;

#define DONE_GUARD(done) \
  if ((done) != nullptr && *(volatile const int *)(done) != 0) return;

// Programmatic dependent launch (sm_90+): the kernels of a Gram-Schmidt column are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, trigger their successor as soon as all their CTAs have
// started (PDL_TRIGGER at the top), and wait for their predecessor's completion + memory flush (PDL_WAIT) before
// the first access to anything it wrote -- the successor's launch latency and prologue overlap the
// predecessor's last wave (~3 us per kernel boundary, three to four boundaries per column).  Without the launch
// attribute both instructions are no-ops.  SVFSI_PDL=0 turns the attribute off.
#define PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;")
#define PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SVFSI_PDL");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ double2 ldg_stream2(const double2 *p) {
  // streaming 128-bit load: matrix values are read exactly once per SpMV
  return __ldcs(p);
}

// flag words of the peer-memory exchanges (see the halo kernels below)
__device__ __forceinline__ void st_flag_sys(volatile int *p, int v) {
  __threadfence_system();
  *p = v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Spin on a word in local memory that the peer writes -- BOUNDED: a dead or late rank must not
// deadlock every GPU of the box (the reference's MPI_WAIT would block too, but an MPI job is killed
// as a whole; a spinning kernel cannot be).  After pd.timeoutNs the waiter raises the sticky error
// words (device copy: later kernels of the stream see it on their first poll and leave at once;
// mapped host copy: the host returns SVFSI_ERR_COMM at its next synchronisation point) and goes on
// with whatever the buffers hold.
__device__ __forceinline__ void wait_flag_sys(volatile int *p, int v, const P2PDev &pd) {
  const int e0 = *(volatile int *)pd.errDev;   // issued together with the first poll
  if (*p != v && e0 == 0) {
    const unsigned long long t0 = global_ns();
    unsigned it = 0;
    while (*p != v) {
      if ((++it & 255u) == 0) {
        if (*(volatile int *)pd.errDev != 0) break;
        if (global_ns() - t0 > (unsigned long long)pd.timeoutNs) {
          *(volatile int *)pd.errDev = 1;
          *pd.errHost = 1;
          break;
        }
      }
    }
  }
  __threadfence_system();
}
// flags live at the start of the arena: int[4][64]: 0 = halo slot0, 1 = halo slot1, 2 = ar slot0, 3 = ar slot1
__device__ __forceinline__ volatile int *flag_ptr(char *base, int which, int src) {
  return (volatile int *)base + which * 64 + src;
}

// ---------------------------------------------------------------------------
// SPARMULVV, dof = 4 (L/SPARMUL.f:98-113).  8 lanes own one block row: lane q
// loads the q-th 16-byte piece of every 128-byte block (one LDG.128 per lane,
// a warp instruction covers four whole 128-byte lines), multiplies by the
// matching half of U(:,col) and the two halves are combined by one shuffle.
// Column ids of 8 consecutive blocks are fetched by one coalesced load and
// broadcast by shuffles, so the U gather does not wait on a dependent load per
// block.
// SCALE: K <- (W_row K) W_col on the way (PRECONDDIAG's scaling fused into the first product of a solve; the
// same two multiplications per entry as scale_val4_kernel), W = the scaling vector
template <bool SCALE = false>
__device__ __forceinline__ double spmv_vv4_row(int row, int q, int h, unsigned gmask,
                                               const int *__restrict__ rowPtr,
                                               const int *__restrict__ col,
                                               const double2 *K,
                                               const double2 *__restrict__ U,
                                               const double *__restrict__ W = nullptr) {
  const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
  const double wr = SCALE ? __ldg(W + (size_t)row * 4 + (q >> 1)) : 1.0;
  double acc = 0.0;
  for (int base = s; base < e; base += 8) {
    const int mine = base + q;
    const int cq = (mine < e) ? __ldg(col + mine) : 0;
    const int cnt = min(8, e - base);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int c = __shfl_sync(gmask, cq, k, 8);
      if (k < cnt) {
        double2 kv = ldg_stream2(K + (size_t)(base + k) * 8 + q);
        const double2 uv = __ldg(U + (size_t)c * 2 + h);
        if (SCALE) {
          const double2 wc = __ldg((const double2 *)W + (size_t)c * 2 + h);
          kv.x = (kv.x * wr) * wc.x;
          kv.y = (kv.y * wr) * wc.y;
          __stcs(const_cast<double2 *>(K) + (size_t)(base + k) * 8 + q, kv);
        }
        acc = fma(kv.x, uv.x, acc);
        acc = fma(kv.y, uv.y, acc);
      }
    }
  }
  acc += __shfl_xor_sync(gmask, acc, 1, 8);
  return acc;
}

__global__ void __launch_bounds__(256) spmv_vv4_kernel(int r0, int r1, int r2, int r3,
                                                        const int *__restrict__ rowPtr,
                                                        const int *__restrict__ col,
                                                        const double2 *__restrict__ K,
                                                        const double2 *__restrict__ U,
                                                        double *__restrict__ KU,
                                                        const int *done) {
  DONE_GUARD(done);
  const int lane = threadIdx.x & 31;
  const int q = lane & 7;
  const int h = q & 1;
  const unsigned gmask = 0xFFu << (lane & 24);
  // rows [r0,r1) followed by rows [r2, ...): the two boundary slabs go out in one launch
  int row = r0 + (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 3);
  if (row >= r1) row += r2 - r1;
  if (row >= r3) return;  // whole 8-lane groups leave together
  const double acc = spmv_vv4_row(row, q, h, gmask, rowPtr, col, K, U);
  if (h == 0) KU[(size_t)row * 4 + (q >> 1)] = acc;
}

// ---------------------------------------------------------------------------
// SPARMULVV dof = 4, quad version: FOUR lanes per block row; lane r holds row r of every 4x4 block
// (one 256-bit streaming load, LDG.E.EF.ENL2.256) and the whole U(:,col) (one 256-bit load, the
// same address for the four lanes), so KU(r,row) = sum_j sum_m K(r,m,j) U(m,col_j) accumulates in
// ONE register in exactly the j, m order of L/SPARMUL.f:104-111 -- no partial sums, no shuffle
// reduction -- and a warp keeps eight rows (eight dependent chains rowPtr -> col -> U) in flight
// instead of four.  The next four column ids are fetched while the current four are used.
struct __align__(32) ldbl4 { double x, y, z, w; };
__device__ __forceinline__ ldbl4 ld256_stream(const double *p) {
  ldbl4 r;
  asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ ldbl4 ld256_nc(const double *p) {
  ldbl4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double spmv_vv4_quad_row(int row, int r, unsigned gmask,
                                                    const int *__restrict__ rowPtr,
                                                    const int *__restrict__ col,
                                                    const double *__restrict__ K,
                                                    const double *__restrict__ U) {
  const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
  double acc = 0.0;
  int cq = (s + r < e) ? __ldg(col + s + r) : 0;
  for (int base = s; base < e; base += 4) {
    const int nxt = base + 4 + r;
    const int cqn = (nxt < e) ? __ldg(col + nxt) : 0;
    const int cnt = min(4, e - base);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (k < cnt) {   // group-uniform
        const int c = __shfl_sync(gmask, cq, k, 4);
        const ldbl4 kv = ld256_stream(K + (size_t)(base + k) * 16 + r * 4);
        const ldbl4 uv = ld256_nc(U + (size_t)c * 4);
        acc = fma(kv.x, uv.x, acc);
        acc = fma(kv.y, uv.y, acc);
        acc = fma(kv.z, uv.z, acc);
        acc = fma(kv.w, uv.w, acc);
      }
    }
    cq = cqn;
  }
  return acc;
}
__global__ void __launch_bounds__(256) spmv_vv4_quad_kernel(int r0, int r1, int r2, int r3,
                                                             const int *__restrict__ rowPtr,
                                                             const int *__restrict__ col,
                                                             const double *__restrict__ K,
                                                             const double *__restrict__ U,
                                                             double *__restrict__ KU,
                                                             const int *done) {
  PDL_TRIGGER();
  PDL_WAIT();
  DONE_GUARD(done);
  const int lane = threadIdx.x & 31, r = lane & 3;
  const unsigned gmask = 0xFu << (lane & 28);
  int row = r0 + (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 2);
  if (row >= r1) row += r2 - r1;
  if (row >= r3) return;  // whole 4-lane groups leave together
  KU[(size_t)row * 4 + r] = spmv_vv4_quad_row(row, r, gmask, rowPtr, col, K, U);
}

// PRECONDDIAG's K <- W K W (L/PRECOND.f:372-489) fused into the FIRST product of the solve: the block row is
// read once, scaled with the same two multiplications per entry as scale_val4_kernel ((K * W_row) * W_col),
// written back, and multiplied with U in the same pass -- one 3.3 GB read of Val less per Newton iteration.
// Bit-identical to the scaling pass followed by spmv_vv4_quad_kernel.
__global__ void __launch_bounds__(256) spmv_vv4_quad_scale_kernel(int nNo, const int *__restrict__ rowPtr,
                                                                   const int *__restrict__ col,
                                                                   double *__restrict__ K,
                                                                   const double *__restrict__ W,
                                                                   const double *__restrict__ U,
                                                                   double *__restrict__ KU) {
  const int lane = threadIdx.x & 31, r = lane & 3;
  const unsigned gmask = 0xFu << (lane & 28);
  const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 2);
  if (row >= nNo) return;  // whole 4-lane groups leave together
  const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
  const double wr = __ldg(W + (size_t)row * 4 + r);
  double acc = 0.0;
  int cq = (s + r < e) ? __ldg(col + s + r) : 0;
  for (int base = s; base < e; base += 4) {
    const int nxt = base + 4 + r;
    const int cqn = (nxt < e) ? __ldg(col + nxt) : 0;
    const int cnt = min(4, e - base);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (k < cnt) {   // group-uniform
        const int c = __shfl_sync(gmask, cq, k, 4);
        double *kp = K + (size_t)(base + k) * 16 + r * 4;
        ldbl4 kv = ld256_stream(kp);
        const ldbl4 wc = ld256_nc(W + (size_t)c * 4);
        const ldbl4 uv = ld256_nc(U + (size_t)c * 4);
        kv.x = (kv.x * wr) * wc.x;
        kv.y = (kv.y * wr) * wc.y;
        kv.z = (kv.z * wr) * wc.z;
        kv.w = (kv.w * wr) * wc.w;
        asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(kp), "d"(kv.x), "d"(kv.y), "d"(kv.z),
                     "d"(kv.w) : "memory");
        acc = fma(kv.x, uv.x, acc);
        acc = fma(kv.y, uv.y, acc);
        acc = fma(kv.z, uv.z, acc);
        acc = fma(kv.w, uv.w, acc);
      }
    }
    cq = cqn;
  }
  KU[(size_t)row * 4 + r] = acc;
}
void launch_spmv_vv4_scale(cudaStream_t st, int nNo, const int *rowPtr, const int *col, double *K,
                           const double *W, const double *U, double *KU) {
  if (nNo <= 0) return;
  count_launch();
  const int blocks = (int)(((size_t)nNo * 4 + 255) / 256);
  spmv_vv4_quad_scale_kernel<<<blocks, 256, 0, st>>>(nNo, rowPtr, col, K, W, U, KU);
}

// ---------------------------------------------------------------------------
// SpMV + halo send in ONE kernel (peer-memory path, FSILS_SPARMUL* = product followed by
// FSILS_COMMUV, L/SPARMUL.f:130).  The rows shared with other ranks -- the two slabs at the ends
// of the reordered numbering -- are given to the FIRST CTAs of the grid; each of their results is
// stored locally AND straight into the receive buffer of every rank sharing that node (NVLink
// stores into the peer's IPC-mapped arena).  The last boundary CTA to finish raises the
// neighbours' flags, so the transfer is in flight while the remaining CTAs stream the interior
// rows.  A set `done` flag skips the arithmetic but never the flag protocol (the receiving
// kernel of the same exchange waits on it).
__device__ __forceinline__ bool fuse_map_row(const SpmvFuse &f, int rowsPerCta, int grp, int &row,
                                             int &bidx) {
  if (f.trace && blockIdx.x == 0 && threadIdx.x == 0) f.trace[5] = global_ns();   // SpMV kernel start
  if ((int)blockIdx.x < f.bndCtas) {
    bidx = blockIdx.x * rowsPerCta + grp;
    if (bidx >= f.nBnd) return false;
    row = bidx < f.shnNo ? bidx : f.mynNo + (bidx - f.shnNo);
    return true;
  }
  bidx = -1;
  row = f.shnNo + ((int)blockIdx.x - f.bndCtas) * rowsPerCta + grp;
  return row < f.mynNo;
}
__device__ __forceinline__ void fuse_send(const SpmvFuse &f, int bidx, int rd, int comp, double v) {
  const int slot = f.seq & 1;
  for (int k = __ldg(f.sendPtr + bidx); k < __ldg(f.sendPtr + bidx + 1); k++) {
    double *dst = (double *)(f.pd.peer[__ldg(f.sendRank + k)] + f.pd.offHalo) +
                  (size_t)slot * f.pd.haloCap + (size_t)__ldg(f.sendOff + k) * rd + comp;
    *dst = v;
  }
}
__device__ __forceinline__ void fuse_publish(const SpmvFuse &f) {
  if ((int)blockIdx.x >= f.bndCtas) return;   // block-uniform
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(f.counter, 1u) == (unsigned)f.bndCtas - 1u);
  __syncthreads();
  if (last) {
    if ((int)threadIdx.x < f.nNbr)
      st_flag_sys(flag_ptr(f.pd.peer[f.nbrRank[threadIdx.x]], f.seq & 1, f.pd.rank), f.seq);
    if (threadIdx.x == 0) *f.counter = 0;
    if (threadIdx.x == 0 && f.trace) f.trace[6] = global_ns();   // halo of this SpMV published
  }
}

template <bool SCALE>
__global__ void __launch_bounds__(256) spmv_vv4_fused_kernel(SpmvFuse f,
                                                              const int *__restrict__ rowPtr,
                                                              const int *__restrict__ col,
                                                              const double2 *K,
                                                              const double2 *__restrict__ U,
                                                              double *__restrict__ KU,
                                                              const int *done, const double *__restrict__ W) {
  PDL_TRIGGER();
  PDL_WAIT();
  const bool skip = (done != nullptr && *(volatile const int *)done != 0);
  const int lane = threadIdx.x & 31, q = lane & 7, h = q & 1;
  const unsigned gmask = 0xFFu << (lane & 24);
  int row, bidx;
  const bool have = fuse_map_row(f, 32, threadIdx.x >> 3, row, bidx);
  if (have && !skip) {
    const double acc = spmv_vv4_row<SCALE>(row, q, h, gmask, rowPtr, col, K, U, W);
    if (h == 0) {
      KU[(size_t)row * 4 + (q >> 1)] = acc;
      if (bidx >= 0) fuse_send(f, bidx, 4, q >> 1, acc);
    }
  }
  fuse_publish(f);
}
// the same with four lanes per block row (64 rows per CTA): every lane owns one component of its row
// and sends it itself
__global__ void __launch_bounds__(256) spmv_vv4_quad_fused_kernel(SpmvFuse f,
                                                                   const int *__restrict__ rowPtr,
                                                                   const int *__restrict__ col,
                                                                   const double *__restrict__ K,
                                                                   const double *__restrict__ U,
                                                                   double *__restrict__ KU,
                                                                   const int *done) {
  const bool skip = (done != nullptr && *(volatile const int *)done != 0);
  const int lane = threadIdx.x & 31, r = lane & 3;
  const unsigned gmask = 0xFu << (lane & 28);
  int row, bidx;
  const bool have = fuse_map_row(f, 64, threadIdx.x >> 2, row, bidx);
  if (have && !skip) {
    const double acc = spmv_vv4_quad_row(row, r, gmask, rowPtr, col, K, U);
    KU[(size_t)row * 4 + r] = acc;
    if (bidx >= 0) fuse_send(f, bidx, 4, r, acc);
  }
  fuse_publish(f);
}

// Generic shapes (VV dof<=3, VS, SV, SS): 4 lanes per row, each lane takes blocks
// q, q+4, ... of the row; partial results combined by shuffles.  BR x BC is the
// block shape: VV d: (d,d); VS d: (1,d); SV d: (d,1); SS: (1,1).
template <int BR, int BC>
__device__ __forceinline__ void spmv_generic_row(int row, int q, unsigned gmask,
                                                 const int *__restrict__ rowPtr,
                                                 const int *__restrict__ col,
                                                 const double *__restrict__ K,
                                                 const double *__restrict__ U, double (&acc)[BR]) {
  const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
#pragma unroll
  for (int l = 0; l < BR; l++) acc[l] = 0.0;
  for (int j = s + q; j < e; j += 4) {
    const int c = __ldg(col + j);
    const double *k = K + (size_t)j * (BR * BC);
    double u[BC];
#pragma unroll
    for (int m = 0; m < BC; m++) u[m] = __ldg(U + (size_t)c * BC + m);
#pragma unroll
    for (int l = 0; l < BR; l++)
#pragma unroll
      for (int m = 0; m < BC; m++) acc[l] = fma(__ldcs(k + l * BC + m), u[m], acc[l]);
  }
#pragma unroll
  for (int l = 0; l < BR; l++) {
    acc[l] += __shfl_xor_sync(gmask, acc[l], 1, 4);
    acc[l] += __shfl_xor_sync(gmask, acc[l], 2, 4);
  }
}

template <int BR, int BC>
__global__ void __launch_bounds__(256) spmv_generic_kernel(int r0, int r1, int r2, int r3,
                                                            const int *__restrict__ rowPtr,
                                                            const int *__restrict__ col,
                                                            const double *__restrict__ K,
                                                            const double *__restrict__ U,
                                                            double *__restrict__ KU,
                                                            const int *done) {
  DONE_GUARD(done);
  const int lane = threadIdx.x & 31;
  const int q = lane & 3;
  const unsigned gmask = 0xFu << (lane & 28);
  int row = r0 + (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 2);
  if (row >= r1) row += r2 - r1;
  if (row >= r3) return;
  double acc[BR];
  spmv_generic_row<BR, BC>(row, q, gmask, rowPtr, col, K, U, acc);
  if (q == 0) {
#pragma unroll
    for (int l = 0; l < BR; l++) KU[(size_t)row * BR + l] = acc[l];
  }
}

template <int BR, int BC>
__global__ void __launch_bounds__(256) spmv_generic_fused_kernel(SpmvFuse f,
                                                                  const int *__restrict__ rowPtr,
                                                                  const int *__restrict__ col,
                                                                  const double *__restrict__ K,
                                                                  const double *__restrict__ U,
                                                                  double *__restrict__ KU,
                                                                  const int *done) {
  const bool skip = (done != nullptr && *(volatile const int *)done != 0);
  const int lane = threadIdx.x & 31, q = lane & 3;
  const unsigned gmask = 0xFu << (lane & 28);
  int row, bidx;
  const bool have = fuse_map_row(f, 64, threadIdx.x >> 2, row, bidx);
  if (have && !skip) {
    double acc[BR];
    spmv_generic_row<BR, BC>(row, q, gmask, rowPtr, col, K, U, acc);
    if (q == 0) {
#pragma unroll
      for (int l = 0; l < BR; l++) {
        KU[(size_t)row * BR + l] = acc[l];
        if (bidx >= 0) fuse_send(f, bidx, BR, l, acc[l]);
      }
    }
  }
  fuse_publish(f);
}

// ---------------------------------------------------------------------------
// "Flat row" SpMV for the small block shapes of NSSOLVER / CG (3x3, 1x3, 3x1, 1x1, 2x2, 1x2, 2x1):
// T lanes own one block row and read its values as ONE contiguous run of doubles -- lane t takes
// flat entries t, t+T, ... of the row, so a warp instruction covers whole 128-byte lines (the
// lane-per-block kernel above touches a different line per lane and is bound by L1 wavefronts,
// 8.7 per 72-byte block).  T is a multiple of BR*BC, hence every lane keeps a fixed (l, m) entry
// position and needs one U component; the T/BR partial sums of an output row meet in shared
// memory and are added in a fixed order (deterministic).  32/T rows per warp.
template <int BR, int BC, int T>
struct FlatCfg {
  static constexpr int BB = BR * BC;       // doubles per block
  static constexpr int BPI = T / BB;       // blocks per iteration
  static constexpr int RPW = 32 / T;       // rows per warp
  static constexpr int RPC = RPW * 8;      // rows per 256-thread CTA
  static_assert(T % BB == 0 && T <= 32, "T must be a multiple of the block size");
};
template <int BR, int BC, int T>
__device__ __forceinline__ void spmv_flat_row(bool laneOk, bool have, int row, int sub, int t,
                                              const int *__restrict__ rowPtr,
                                              const int *__restrict__ col,
                                              const double *__restrict__ K,
                                              const double *__restrict__ U, double *smw,
                                              double &out) {
  using C = FlatCfg<BR, BC, T>;
  const int lb = t / C::BB, e = t - lb * C::BB, m = e % BC;
  double acc = 0.0;
  if (have) {
    const int s = __ldg(rowPtr + row), end = __ldg(rowPtr + row + 1);
    for (int j = s + lb; j < end; j += C::BPI) {
      const int c = __ldg(col + j);
      acc = fma(__ldcs(K + (size_t)j * C::BB + e), __ldg(U + (size_t)c * BC + m), acc);
    }
  }
  if (laneOk) smw[sub * T + t] = acc;   // lanes beyond RPW*T own no slot
  __syncwarp();
  out = 0.0;
  if (have && t < BR) {        // lane t adds the partials of output row l = t: blocks, then m
#pragma unroll
    for (int b = 0; b < C::BPI; b++)
#pragma unroll
      for (int mm = 0; mm < BC; mm++) out += smw[sub * T + b * C::BB + t * BC + mm];
  }
  __syncwarp();
}

template <int BR, int BC, int T>
__global__ void __launch_bounds__(256) spmv_flat_kernel(int r0, int r1, int r2, int r3,
                                                        const int *__restrict__ rowPtr,
                                                        const int *__restrict__ col,
                                                        const double *__restrict__ K,
                                                        const double *__restrict__ U,
                                                        double *__restrict__ KU, const int *done) {
  DONE_GUARD(done);
  using C = FlatCfg<BR, BC, T>;
  __shared__ double sm[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int sub = lane / T, t = lane - sub * T;
  int row = r0 + ((int)blockIdx.x * 8 + w) * C::RPW + sub;
  if (row >= r1) row += r2 - r1;
  const bool have = (sub < C::RPW) && (row < r3);
  double out;
  spmv_flat_row<BR, BC, T>(sub < C::RPW, have, row, sub, t, rowPtr, col, K, U, sm[w], out);
  if (have && t < BR) KU[(size_t)row * BR + t] = out;
}

template <int BR, int BC, int T>
__global__ void __launch_bounds__(256) spmv_flat_fused_kernel(SpmvFuse f,
                                                              const int *__restrict__ rowPtr,
                                                              const int *__restrict__ col,
                                                              const double *__restrict__ K,
                                                              const double *__restrict__ U,
                                                              double *__restrict__ KU,
                                                              const int *done) {
  using C = FlatCfg<BR, BC, T>;
  __shared__ double sm[8][32];
  const bool skip = (done != nullptr && *(volatile const int *)done != 0);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int sub = lane / T, t = lane - sub * T;
  int row = 0, bidx = -1;
  bool have = false;
  if (sub < C::RPW) have = fuse_map_row(f, C::RPC, w * C::RPW + sub, row, bidx);
  have = have && !skip;
  double out;
  spmv_flat_row<BR, BC, T>(sub < C::RPW, have, row, sub, t, rowPtr, col, K, U, sm[w], out);
  if (have && t < BR) {
    KU[(size_t)row * BR + t] = out;
    if (bidx >= 0) fuse_send(f, bidx, BR, t, out);
  }
  fuse_publish(f);
}

template <int BR, int BC, int T>
static void launch_flat(cudaStream_t st, int r0, int r1, int r2, int r3, const int *rowPtr,
                        const int *col, const double *K, const double *U, double *KU,
                        const int *done) {
  const int rows = (r1 - r0) + (r3 - r2);
  const int rpc = FlatCfg<BR, BC, T>::RPC;
  spmv_flat_kernel<BR, BC, T><<<(rows + rpc - 1) / rpc, 256, 0, st>>>(r0, r1, r2, r3, rowPtr, col, K,
                                                                      U, KU, done);
}
template <int BR, int BC, int T>
static void launch_flat_fused(cudaStream_t st, SpmvFuse f, const int *rowPtr, const int *col,
                              const double *K, const double *U, double *KU, const int *done) {
  const int rpc = FlatCfg<BR, BC, T>::RPC;
  f.bndCtas = (f.nBnd + rpc - 1) / rpc;
  const int inner = f.mynNo - f.shnNo;
  const int blocks = f.bndCtas + (inner + rpc - 1) / rpc;
  if (blocks <= 0) return;
  spmv_flat_fused_kernel<BR, BC, T><<<blocks, 256, 0, st>>>(f, rowPtr, col, K, U, KU, done);
}
// shape -> kernel.  MEASURED SLOWER than the lane-per-block kernels on B200 (10M tets: NSSOLVER step
// 12.8 vs 9.5 ms of SpMV, heat CG 7.1 vs 2.8 ms; profiles/r01_spmv_shapes.md): with ~15 blocks per row
// the per-row overhead (row pointers, shared-memory meeting point, serial tail) outweighs the
// coalescing gain.  Kept behind SVFSI_SPMV_FLAT=1 as a measured negative result; default off.
static bool use_flat() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SVFSI_SPMV_FLAT");
    v = (e && atoi(e) == 1) ? 1 : 0;
  }
  return v != 0;
}
#define FLAT_DISPATCH(CALL)                                              \
  do {                                                                   \
    if (kind == 3 || dof == 1) { CALL(1, 1, 16); return; }               \
    if (kind == 0 && dof == 3) { CALL(3, 3, 27); return; }               \
    if (kind == 0 && dof == 2) { CALL(2, 2, 16); return; }               \
    if (kind == 1 && dof == 3) { CALL(1, 3, 15); return; }               \
    if (kind == 2 && dof == 3) { CALL(3, 1, 15); return; }               \
    if (kind == 1 && dof == 2) { CALL(1, 2, 16); return; }               \
    if (kind == 2 && dof == 2) { CALL(2, 1, 16); return; }               \
  } while (0)

// ---------------------------------------------------------------------------
// "Row-chunk stream" SpMV for the small block shapes of NSSOLVER / CG (K 3x3, G 3x1, D 1x3, L 1x1, heat 1x1;
// L/SPARMUL.f:135-297).  The lane-per-block kernel above reads a different 128-byte line per lane (blocks are
// 72 / 24 / 8 bytes apart) and is bound by L1 wavefronts, not HBM (0.53-0.66 of the measured peak,
// profiles/r01_spmv_shapes.md).  Here a CTA owns ROWS consecutive block rows -- whose values and column ids are
// ONE contiguous run -- copies that run to shared memory with fully coalesced 128-bit loads (every byte of every
// line used, tens of independent loads in flight per thread), then LPR = 256 / ROWS lanes per row walk the
// blocks out of shared memory and gather U through L1 / L2.  Partial sums of a row meet by shuffles in a fixed
// order (deterministic).  The fused variant (halo send, boundary rows first) is the same kernel.
struct StreamMap {
  int a0, a1, b0, b1, c0, c1;   // three contiguous row ranges (the last two may be empty)
  int n0, n1;                   // CTAs of the first two ranges
};
template <int BR, int BC, int ROWS>
__global__ void __launch_bounds__(256) spmv_stream_kernel(StreamMap mp, int fused, SpmvFuse f, int capB,
                                                          const int *__restrict__ rowPtr,
                                                          const int *__restrict__ col,
                                                          const double *__restrict__ K,
                                                          const double *__restrict__ U,
                                                          double *__restrict__ KU, const int *done) {
  constexpr int BB = BR * BC, LPR = 256 / ROWS;
  extern __shared__ double2 sm2[];
  double *sK = (double *)sm2;                       // [capB * BB + 2]
  int *sCol = (int *)(sK + (size_t)capB * BB + 2);  // [capB]
  int *sRp = sCol + capB;                           // [ROWS + 1]
  const bool skip = (done != nullptr && *(volatile const int *)done != 0);
  if (skip && !fused) return;
  int rs, re;
  {
    const int b = blockIdx.x;
    if (b < mp.n0) { rs = mp.a0 + b * ROWS; re = min(rs + ROWS, mp.a1); }
    else if (b < mp.n0 + mp.n1) { rs = mp.b0 + (b - mp.n0) * ROWS; re = min(rs + ROWS, mp.b1); }
    else { rs = mp.c0 + (b - mp.n0 - mp.n1) * ROWS; re = min(rs + ROWS, mp.c1); }
  }
  const int nr = re - rs;
  if (!skip && nr > 0) {
    for (int t = threadIdx.x; t <= nr; t += 256) sRp[t] = __ldg(rowPtr + rs + t);
    __syncthreads();
    const int b0 = sRp[0], nb = sRp[nr] - b0;
    // values: K[b0*BB .. (b0+nb)*BB) mirrored in shared memory with the same 16-byte parity
    const size_t g0 = (size_t)b0 * BB;
    const int pad = (int)(g0 & 1), nd = nb * BB;
    const int head = pad ? 1 : 0;
    if (head && threadIdx.x == 0 && nd > 0) sK[pad] = __ldcs(K + g0);
    const int np = (nd - head) >> 1;
    const double2 *Kv = (const double2 *)(K + g0 + head);
    double2 *sv = (double2 *)(sK + pad + head);
    for (int t = threadIdx.x; t < np; t += 256) sv[t] = __ldcs(Kv + t);
    if (threadIdx.x == 0 && head + 2 * np < nd) sK[pad + nd - 1] = __ldcs(K + g0 + nd - 1);
    for (int t = threadIdx.x; t < nb; t += 256) sCol[t] = __ldg(col + b0 + t);
    __syncthreads();
    const int rl = threadIdx.x / LPR, q = threadIdx.x % LPR;
    double acc[BR];
#pragma unroll
    for (int l = 0; l < BR; l++) acc[l] = 0.0;
    if (rl < nr) {
      const int js = sRp[rl] - b0, je = sRp[rl + 1] - b0;
      for (int j = js + q; j < je; j += LPR) {
        const int c = sCol[j];
        const double *k = sK + pad + (size_t)j * BB;
        double u[BC];
#pragma unroll
        for (int m = 0; m < BC; m++) u[m] = __ldg(U + (size_t)c * BC + m);
#pragma unroll
        for (int l = 0; l < BR; l++)
#pragma unroll
          for (int m = 0; m < BC; m++) acc[l] = fma(k[l * BC + m], u[m], acc[l]);
      }
    }
#pragma unroll
    for (int l = 0; l < BR; l++)
#pragma unroll
      for (int o = 1; o < LPR; o <<= 1) acc[l] += __shfl_xor_sync(0xffffffffu, acc[l], o);
    if (rl < nr && q == 0) {
      const int row = rs + rl;
      int bidx = -1;
      if (fused && (int)blockIdx.x < mp.n0 + mp.n1) bidx = row < f.shnNo ? row : f.shnNo + (row - f.mynNo);
#pragma unroll
      for (int l = 0; l < BR; l++) {
        KU[(size_t)row * BR + l] = acc[l];
        if (bidx >= 0) fuse_send(f, bidx, BR, l, acc[l]);
      }
    }
  }
  if (fused) fuse_publish(f);
}

static int spmv_stream_mode() {
  static int v = -1;
  if (v < 0) {
    // MEASURED (profiles/r02_spmv_stream.md, 10M tets): NSSOLVER step 25.1 ms with it vs 24.2 ms without
    // (SpMV mix at 0.585 vs 0.641 of the measured peak), heat CG 6.45 vs 6.47 ms: the staging pass costs
    // what the coalescing gains.  Off by default; SVFSI_SPMV_STREAM=1 selects it (parity-tested both ways).
    const char *e = getenv("SVFSI_SPMV_STREAM");
    v = e ? atoi(e) : 0;
  }
  return v;
}
// returns false if the shape / row lengths do not fit (caller falls back to the lane-per-block kernel)
template <int BR, int BC, int ROWS>
static bool launch_stream(cudaStream_t st, StreamMap mp, int fused, SpmvFuse f, const int *rowPtr,
                          const int *col, const double *K, const double *U, double *KU, const int *done) {
  const int maxRow = ctx().maxRowLen;
  if (maxRow <= 0) return false;
  const int capB = ROWS * maxRow;
  const size_t smem = sizeof(double) * ((size_t)capB * BR * BC + 2) + sizeof(int) * ((size_t)capB + ROWS + 1);
  if (smem > 100 * 1024) return false;
  static size_t attrSet = 0;
  if (smem > attrSet) {
    cudaFuncSetAttribute(spmv_stream_kernel<BR, BC, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attrSet = smem;
  }
  mp.n0 = (mp.a1 - mp.a0 + ROWS - 1) / ROWS;
  mp.n1 = (mp.b1 - mp.b0 + ROWS - 1) / ROWS;
  const int n2 = (mp.c1 - mp.c0 + ROWS - 1) / ROWS;
  f.bndCtas = mp.n0 + mp.n1;
  const int blocks = mp.n0 + mp.n1 + n2;
  if (blocks <= 0) return true;
  spmv_stream_kernel<BR, BC, ROWS><<<blocks, 256, smem, st>>>(mp, fused, f, capB, rowPtr, col, K, U, KU, done);
  return true;
}
static bool stream_dispatch(cudaStream_t st, int kind, int dof, StreamMap mp, int fused, const SpmvFuse &f,
                            const int *rowPtr, const int *col, const double *K, const double *U, double *KU,
                            const int *done) {
  if (!spmv_stream_mode()) return false;
#define STR(BR, BC, ROWS) return launch_stream<BR, BC, ROWS>(st, mp, fused, f, rowPtr, col, K, U, KU, done)
  if (kind == 3 || dof == 1) STR(1, 1, 128);
  if (kind == 0) { if (dof == 2) STR(2, 2, 64); if (dof == 3) STR(3, 3, 32); return false; }
  if (kind == 1) { if (dof == 2) STR(1, 2, 64); if (dof == 3) STR(1, 3, 64); STR(1, 4, 64); }
  if (dof == 2) STR(2, 1, 64);
  if (dof == 3) STR(3, 1, 64);
  STR(4, 1, 64);
#undef STR
}

// ---------------------------------------------------------------------------
// Small block shapes (NSSOLVER: K 3x3, G 3x1, D 1x3, L 1x1; heat CG: 1x1; L/SPARMUL.f:135-297), third attempt.
// What the two measured negative results above have in common with the lane-per-block kernel is the length of
// the DEPENDENT chain a thread walks: row pointers -> (column id -> values and U) once per block it owns, i.e.
// 9-12 serialised memory round trips per thread with 12-76 bytes in flight each.  The dof = 4 quad kernel got to
// the HBM roofline by keeping whole rows in flight; the two families below do the same for the small shapes:
// every load a lane needs for the whole row is ISSUED BEFORE the first one is used (fixed unroll, predicated
// on the row length; longer rows take another trip), so a thread's chain is row pointers -> column ids +
// values -> U: three round trips whatever the row length.
//  * hoist<L, UNR>: lane-per-block as before (L lanes per row, lane q owns blocks q, q+L, ...), UNR blocks per
//    lane per trip.  With L = 4 the summation order -- per lane ascending j, then the xor tree -- is the
//    generic kernel's: bit-identical results.
//  * run<T, STEPS>: the T lanes of a row read the row's values as ONE contiguous run of doubles (lane t takes
//    entries t, t+T, ...: every warp-level load covers whole 128-byte lines, where lane-per-block touches a
//    different line per lane), entry d belongs to block d / (BR BC), component (l, m) = ((d % BB) / BC, d % BC);
//    per-lane partial sums of output row l meet by an xor tree (fixed order: deterministic).
// volatile asm: the compiler keeps these loads in program order -- all of a row's loads are issued before the
// first use (left to itself it sinks every load next to its FMA to save registers, which serialises them)
__device__ __forceinline__ double ldv_cs(const double *p) {
  double v;
  asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ldv_nc(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void ldv_cs2(const double *p, double &x, double &y) {   // 16-byte aligned
  asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "l"(p));
}
__device__ __forceinline__ void ldv_nc2(const double *p, double &x, double &y) {   // 16-byte aligned
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "l"(p));
}
__device__ __forceinline__ void ldv_nc_i2(const int *p, int &x, int &y) {            // 8-byte aligned
  asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "l"(p));
}
// matrix values: streaming (evict-first) or plain read-only loads
template <bool NC> __device__ __forceinline__ double ldk(const double *p) { return NC ? ldv_nc(p) : ldv_cs(p); }
template <bool NC> __device__ __forceinline__ void ldk2(const double *p, double &x, double &y) {
  if (NC) ldv_nc2(p, x, y); else ldv_cs2(p, x, y);
}
__device__ __forceinline__ int ldv_nc_i(const int *p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
template <int T>
__device__ __forceinline__ unsigned lane_group_mask(int lane) {
  return T >= 32 ? 0xffffffffu : (((1u << (T & 31)) - 1u) << (lane & ~(T - 1)));
}
// WIDE (blocks of 9 or 3 doubles, 8-byte aligned): the values of a block are fetched with 128-bit loads from the
// 16-byte aligned window that contains it -- block j starts on a 16-byte boundary iff j is even, and a lane's blocks
// j = base + q + k L all have the parity of s + q when L is even -- 5 / 2 loads per block instead of 9 / 3 and as many
// fewer L1 tag look-ups (the lane-per-block kernels are bound by those: every load touches ~24 lines per warp).  The
// window of an odd block starts one double early; an even block's last load is 64 bits wide (nothing is read past
// the end of K).
template <int BB, bool NC>
__device__ __forceinline__ void ld_block_wide(const double *p, int odd, double (&v)[BB]) {
  static_assert(BB == 9 || BB == 3, "wide loads: 3x3, 3x1, 1x3 blocks");
  const double *a = p - odd;
  double w[BB + 1];
#pragma unroll
  for (int i = 0; i + 2 < BB + 1; i += 2) ldk2<NC>(a + i, w[i], w[i + 1]);
  if (odd) ldk2<NC>(a + BB - 1, w[BB - 1], w[BB]);
  else { w[BB - 1] = ldk<NC>(a + BB - 1); w[BB] = 0.0; }
#pragma unroll
  for (int i = 0; i < BB; i++) v[i] = odd ? w[i + 1] : w[i];
}
// OPT bits: 1 = wide loads of the values, 2 = wide loads of U(:,col) for BC = 3 (U(1:3,c) starts on a 16-byte
// boundary iff c is even: one 128-bit + one 64-bit load in the order the parity dictates, nothing over-read),
// 4 = plain read-only loads of the values instead of streaming (evict-first) ones
template <int BR, int BC, int L, int UNR, int OPT = 0>
__device__ __forceinline__ void spmv_hoist_row(int row, int q, unsigned gmask,
                                               const int *__restrict__ rowPtr,
                                               const int *__restrict__ col,
                                               const double *__restrict__ K,
                                               const double *__restrict__ U, double (&acc)[BR]) {
  constexpr int BB = BR * BC;
  const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
#pragma unroll
  for (int l = 0; l < BR; l++) acc[l] = 0.0;
  for (int base = s; base < e; base += L * UNR) {   // one trip for rows of up to L*UNR blocks
    int c[UNR];
    double kv[UNR][BB], u[UNR][BC];
    // a lane without a block in this trip reads the row's first block (a line the group touches anyway) and
    // drops the values
#pragma unroll
    for (int k = 0; k < UNR; k++) {
      const int j = base + q + k * L;
      c[k] = ldv_nc_i(col + (j < e ? j : s));
    }
#pragma unroll
    for (int k = 0; k < UNR; k++) {
      const int j = base + q + k * L;
      const int jj = j < e ? j : s;
      const double *kp = K + (size_t)jj * BB;
      if constexpr ((OPT & 1) && (BB == 9 || BB == 3)) {
        ld_block_wide<BB, (OPT & 4) != 0>(kp, jj & 1, kv[k]);
      } else {
#pragma unroll
        for (int i = 0; i < BB; i++) kv[k][i] = ldk<(OPT & 4) != 0>(kp + i);
      }
    }
#pragma unroll
    for (int k = 0; k < UNR; k++) {
      const double *up = U + (size_t)c[k] * BC;
      if constexpr ((OPT & 2) && BC == 3) {
        if (c[k] & 1) { u[k][0] = ldv_nc(up); ldv_nc2(up + 1, u[k][1], u[k][2]); }
        else { ldv_nc2(up, u[k][0], u[k][1]); u[k][2] = ldv_nc(up + 2); }
      } else {
#pragma unroll
        for (int m = 0; m < BC; m++) u[k][m] = ldv_nc(up + m);
      }
    }
#pragma unroll
    for (int k = 0; k < UNR; k++) {
      const bool ok = base + q + k * L < e;
#pragma unroll
      for (int l = 0; l < BR; l++)
#pragma unroll
        for (int m = 0; m < BC; m++) acc[l] = fma(ok ? kv[k][l * BC + m] : 0.0, u[k][m], acc[l]);
    }
  }
#pragma unroll
  for (int l = 0; l < BR; l++)
#pragma unroll
    for (int o = 1; o < L; o <<= 1) acc[l] += __shfl_xor_sync(gmask, acc[l], o, L);
}
// 1x1 blocks, two entries per load: lane q takes the entry PAIRS q, q+L, ... of the row's window -- the row's
// entries extended down to an even index, so that every pair is one 128-bit load of values and one 64-bit load of
// column ids (half the L1 wavefronts of the one-entry-per-lane kernels, which is what bounds them).  The entry
// before an odd row start and the one after an odd row end are read and dropped (they exist: the value and
// column-id arrays are 16- and 8-byte aligned and at least nnz + (nnz & 1) entries long).
template <int L, int UNR, bool NC>
__device__ __forceinline__ void spmv_pair_row(int row, int q, unsigned gmask, const int *__restrict__ rowPtr,
                                              const int *__restrict__ col, const double *__restrict__ K,
                                              const double *__restrict__ U, double (&acc)[1]) {
  const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
  const int s2 = s & ~1;
  double a = 0.0;
  for (int base = s2; base < e; base += 2 * L * UNR) {
    int c0[UNR], c1[UNR];
    double k0[UNR], k1[UNR], u0[UNR], u1[UNR];
#pragma unroll
    for (int k = 0; k < UNR; k++) {
      const int j = base + 2 * (q + k * L);
      ldv_nc_i2(col + (j < e ? j : s2), c0[k], c1[k]);
    }
#pragma unroll
    for (int k = 0; k < UNR; k++) {
      const int j = base + 2 * (q + k * L);
      ldk2<NC>(K + (j < e ? j : s2), k0[k], k1[k]);
    }
#pragma unroll
    for (int k = 0; k < UNR; k++) {
      u0[k] = ldv_nc(U + c0[k]);
      u1[k] = ldv_nc(U + c1[k]);
    }
#pragma unroll
    for (int k = 0; k < UNR; k++) {
      const int j = base + 2 * (q + k * L);
      a = fma((j >= s && j < e) ? k0[k] : 0.0, u0[k], a);
      a = fma((j + 1 >= s && j + 1 < e) ? k1[k] : 0.0, u1[k], a);
    }
  }
#pragma unroll
  for (int o = 1; o < L; o <<= 1) a += __shfl_xor_sync(gmask, a, o, L);
  acc[0] = a;
}
template <int BR, int BC, int T, int STEPS>
__device__ __forceinline__ void spmv_run_row(int row, int t, unsigned gmask,
                                             const int *__restrict__ rowPtr,
                                             const int *__restrict__ col,
                                             const double *__restrict__ K,
                                             const double *__restrict__ U, double (&acc)[BR]) {
  constexpr unsigned BB = BR * BC;
  const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
  const unsigned n = (unsigned)(e - s) * BB;   // doubles in this row's run
  const double *kb = K + (size_t)s * BB;
  const int *cb = col + s;
#pragma unroll
  for (int l = 0; l < BR; l++) acc[l] = 0.0;
  for (unsigned d0 = 0; d0 < n; d0 += T * STEPS) {   // one trip for rows of up to T*STEPS doubles
    int c[STEPS];
    double kv[STEPS], uv[STEPS];
    // a lane past the end of the run reads the row's first entry and drops it
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      const unsigned d = d0 + (unsigned)(k * T + t);
      c[k] = ldv_nc_i(cb + (d < n ? d / BB : 0u));
    }
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      const unsigned d = d0 + (unsigned)(k * T + t);
      kv[k] = ldv_cs(kb + (d < n ? d : 0u));
    }
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      const unsigned d = d0 + (unsigned)(k * T + t);
      const unsigned m = (d % BB) % (unsigned)BC;
      uv[k] = ldv_nc(U + (size_t)c[k] * BC + m);
    }
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      const unsigned d = d0 + (unsigned)(k * T + t);
      const int l = (int)((d % BB) / (unsigned)BC);
      const double kk = (d < n) ? kv[k] : 0.0;
#pragma unroll
      for (int ll = 0; ll < BR; ll++) {
        const double v = fma(kk, uv[k], acc[ll]);
        acc[ll] = (BR == 1 || l == ll) ? v : acc[ll];
      }
    }
  }
#pragma unroll
  for (int l = 0; l < BR; l++)
#pragma unroll
    for (int o = 1; o < T; o <<= 1) acc[l] += __shfl_xor_sync(gmask, acc[l], o, T);
}
//  * run-async<T, STEPS> (FAM = 2): the run kernel with every global read made an ASYNCHRONOUS copy into shared
//    memory (cp.async / LDGSTS: no destination register, so neither the compiler nor ptxas can serialise the loads
//    to save registers -- the SASS of the two families above shows them sinking every load next to its FMA).  A
//    row's chain is: row pointers -> [column ids and values in flight together] -> [U in flight] -> sums out of
//    shared memory.  Each lane reads back only the value / U slots it copied itself; column ids cross lanes and
//    are fenced by a group-wide __syncwarp.
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int BR, int BC, int T, int STEPS>
struct RunAsyncCfg {
  static constexpr int BB = BR * BC;
  static constexpr int SPAN = T * STEPS;                 // doubles of a row handled per trip
  static constexpr int CB = (SPAN + BB - 1) / BB + 1;    // blocks a trip can touch (a trip may start mid-block)
  static constexpr int RPC = 256 / T;                    // rows per CTA
  static constexpr size_t SMEM = (size_t)RPC * (2 * SPAN * sizeof(double) + CB * sizeof(int));
};
template <int BR, int BC, int T, int STEPS>
__device__ __forceinline__ void spmv_run_async_row(int row, int t, unsigned gmask, const int *__restrict__ rowPtr,
                                                   const int *__restrict__ col, const double *__restrict__ K,
                                                   const double *__restrict__ U, double *sK, double *sU, int *sC,
                                                   double (&acc)[BR]) {
  using C = RunAsyncCfg<BR, BC, T, STEPS>;
  constexpr unsigned BB = C::BB;
  const int s = __ldg(rowPtr + row), e = __ldg(rowPtr + row + 1);
  const unsigned n = (unsigned)(e - s) * BB;
  const double *kb = K + (size_t)s * BB;
  const int *cb = col + s;
#pragma unroll
  for (int l = 0; l < BR; l++) acc[l] = 0.0;
  for (unsigned d0 = 0; d0 < n; d0 += C::SPAN) {
    const unsigned b0 = d0 / BB;                               // first block of this trip
    const unsigned nbt = min((unsigned)C::CB, (unsigned)(e - s) - b0);
    for (unsigned i = t; i < nbt; i += T) cp_async4(sC + i, cb + b0 + i);
    cp_async_commit();
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      const unsigned d = d0 + (unsigned)(k * T + t);
      if (d < n) cp_async8(sK + k * T + t, kb + d);
    }
    cp_async_commit();
    cp_async_wait<1>();          // my column-id copies have landed ...
    __syncwarp(gmask);           // ... and so have those of the other lanes of the row
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      const unsigned d = d0 + (unsigned)(k * T + t);
      if (d < n) {
        const int c = sC[d / BB - b0];
        cp_async8(sU + k * T + t, U + (size_t)c * BC + (d % BB) % (unsigned)BC);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
#pragma unroll
    for (int k = 0; k < STEPS; k++) {
      const unsigned d = d0 + (unsigned)(k * T + t);
      if (d < n) {
        const int l = (int)((d % BB) / (unsigned)BC);
        const double kk = sK[k * T + t], uu = sU[k * T + t];
#pragma unroll
        for (int ll = 0; ll < BR; ll++) {
          const double v = fma(kk, uu, acc[ll]);
          acc[ll] = (BR == 1 || l == ll) ? v : acc[ll];
        }
      }
    }
    __syncwarp(gmask);           // the next trip overwrites the column ids
  }
#pragma unroll
  for (int l = 0; l < BR; l++)
#pragma unroll
    for (int o = 1; o < T; o <<= 1) acc[l] += __shfl_xor_sync(gmask, acc[l], o, T);
}
template <int BR, int BC, int T, int STEPS>
__global__ void __launch_bounds__(256) spmv_run_async_kernel(int fused, int r0, int r1, int r2, int r3, SpmvFuse f,
                                                             const int *__restrict__ rowPtr,
                                                             const int *__restrict__ col,
                                                             const double *__restrict__ K,
                                                             const double *__restrict__ U,
                                                             double *__restrict__ KU, const int *done) {
  using C = RunAsyncCfg<BR, BC, T, STEPS>;
  extern __shared__ double2 sm2[];
  const bool skip = (done != nullptr && *(volatile const int *)done != 0);
  if (skip && !fused) return;
  const int lane = threadIdx.x & 31, t = lane & (T - 1), grp = threadIdx.x / T;
  const unsigned gmask = lane_group_mask<T>(lane);
  double *sK = (double *)sm2 + (size_t)grp * (2 * C::SPAN);
  double *sU = sK + C::SPAN;
  int *sC = (int *)((double *)sm2 + (size_t)C::RPC * 2 * C::SPAN) + grp * C::CB;
  int row = 0, bidx = -1;
  bool have;
  if (fused) {
    have = fuse_map_row(f, C::RPC, grp, row, bidx);
  } else {
    row = r0 + (int)blockIdx.x * C::RPC + grp;
    if (row >= r1) row += r2 - r1;
    have = row < r3;
  }
  if (have && !skip) {
    double acc[BR];
    spmv_run_async_row<BR, BC, T, STEPS>(row, t, gmask, rowPtr, col, K, U, sK, sU, sC, acc);
#pragma unroll
    for (int l = 0; l < BR; l++)
      if (t == l % T) {
        KU[(size_t)row * BR + l] = acc[l];
        if (bidx >= 0) fuse_send(f, bidx, BR, l, acc[l]);
      }
  }
  if (fused) fuse_publish(f);
}
template <int BR, int BC, int T, int STEPS>
static void launch_run_async(cudaStream_t st, int fused, int r0, int r1, int r2, int r3, SpmvFuse f,
                             const int *rowPtr, const int *col, const double *K, const double *U, double *KU,
                             const int *done) {
  using C = RunAsyncCfg<BR, BC, T, STEPS>;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(spmv_run_async_kernel<BR, BC, T, STEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)C::SMEM);
    attr = true;
  }
  int blocks;
  if (fused) {
    f.bndCtas = (f.nBnd + C::RPC - 1) / C::RPC;
    blocks = f.bndCtas + (f.mynNo - f.shnNo + C::RPC - 1) / C::RPC;
  } else {
    blocks = ((r1 - r0) + (r3 - r2) + C::RPC - 1) / C::RPC;
  }
  if (blocks <= 0) return;
  spmv_run_async_kernel<BR, BC, T, STEPS><<<blocks, 256, C::SMEM, st>>>(fused, r0, r1, r2, r3, f, rowPtr, col, K, U,
                                                                       KU, done);
}
// FAM = 0: hoist (LPR lanes per row, P = blocks per lane and trip); 3: hoist with wide value loads; 10 + OPT: hoist
// with the OPT bits of spmv_hoist_row; 20 / 21: entry pairs (1x1 only; 21 = plain instead of streaming loads);
// 1: run (LPR lanes per row, P = steps); 2 (run-async) has its own kernel
template <int FAM, int BR, int BC, int LPR, int P>
__device__ __forceinline__ void spmv_small_row(int row, int t, unsigned gmask, const int *__restrict__ rowPtr,
                                               const int *__restrict__ col, const double *__restrict__ K,
                                               const double *__restrict__ U, double (&acc)[BR]) {
  if constexpr (FAM == 1) spmv_run_row<BR, BC, LPR, P>(row, t, gmask, rowPtr, col, K, U, acc);
  else if constexpr (FAM >= 20 && BR * BC == 1) spmv_pair_row<LPR, P, (FAM & 1) != 0>(row, t, gmask, rowPtr, col, K, U, acc);
  else if constexpr (FAM >= 10) spmv_hoist_row<BR, BC, LPR, P, FAM - 10>(row, t, gmask, rowPtr, col, K, U, acc);
  else spmv_hoist_row<BR, BC, LPR, P, FAM == 3 ? 1 : 0>(row, t, gmask, rowPtr, col, K, U, acc);
}
template <int FAM, int BR, int BC, int LPR, int P>
__global__ void __launch_bounds__(256) spmv_small_kernel(int r0, int r1, int r2, int r3,
                                                          const int *__restrict__ rowPtr,
                                                          const int *__restrict__ col,
                                                          const double *__restrict__ K,
                                                          const double *__restrict__ U,
                                                          double *__restrict__ KU, const int *done) {
  DONE_GUARD(done);
  const int lane = threadIdx.x & 31, t = lane & (LPR - 1);
  const unsigned gmask = lane_group_mask<LPR>(lane);
  int row = r0 + (int)((blockIdx.x * 256u + threadIdx.x) / (unsigned)LPR);
  if (row >= r1) row += r2 - r1;
  if (row >= r3) return;   // whole LPR-lane groups leave together
  double acc[BR];
  spmv_small_row<FAM, BR, BC, LPR, P>(row, t, gmask, rowPtr, col, K, U, acc);
#pragma unroll
  for (int l = 0; l < BR; l++)
    if (t == l % LPR) KU[(size_t)row * BR + l] = acc[l];   // every lane holds the full sums
}
template <int FAM, int BR, int BC, int LPR, int P>
__global__ void __launch_bounds__(256) spmv_small_fused_kernel(SpmvFuse f, const int *__restrict__ rowPtr,
                                                                const int *__restrict__ col,
                                                                const double *__restrict__ K,
                                                                const double *__restrict__ U,
                                                                double *__restrict__ KU, const int *done) {
  const bool skip = (done != nullptr && *(volatile const int *)done != 0);
  const int lane = threadIdx.x & 31, t = lane & (LPR - 1);
  const unsigned gmask = lane_group_mask<LPR>(lane);
  int row, bidx;
  const bool have = fuse_map_row(f, 256 / LPR, threadIdx.x / LPR, row, bidx);
  if (have && !skip) {
    double acc[BR];
    spmv_small_row<FAM, BR, BC, LPR, P>(row, t, gmask, rowPtr, col, K, U, acc);
#pragma unroll
    for (int l = 0; l < BR; l++)
      if (t == l % LPR) {
        KU[(size_t)row * BR + l] = acc[l];
        if (bidx >= 0) fuse_send(f, bidx, BR, l, acc[l]);
      }
  }
  fuse_publish(f);
}
template <int FAM, int BR, int BC, int LPR, int P>
static void launch_small(cudaStream_t st, int r0, int r1, int r2, int r3, const int *rowPtr, const int *col,
                         const double *K, const double *U, double *KU, const int *done) {
  if constexpr (FAM == 3 || FAM >= 10) {   // 128-bit loads need 16-byte aligned arrays (column ids: 8)
    if ((((uintptr_t)K | (uintptr_t)U) & 15) || ((uintptr_t)col & 7)) {
      launch_small<0, BR, BC, 4, 4>(st, r0, r1, r2, r3, rowPtr, col, K, U, KU, done);
      return;
    }
  }
  if constexpr (FAM == 2) {
    SpmvFuse none;
    memset(&none, 0, sizeof(none));
    launch_run_async<BR, BC, LPR, P>(st, 0, r0, r1, r2, r3, none, rowPtr, col, K, U, KU, done);
  } else {
    const int rows = (r1 - r0) + (r3 - r2);
    const int blocks = (int)(((size_t)rows * LPR + 255) / 256);
    spmv_small_kernel<FAM, BR, BC, LPR, P><<<blocks, 256, 0, st>>>(r0, r1, r2, r3, rowPtr, col, K, U, KU, done);
  }
}
template <int FAM, int BR, int BC, int LPR, int P>
static void launch_small_fused(cudaStream_t st, SpmvFuse f, const int *rowPtr, const int *col, const double *K,
                               const double *U, double *KU, const int *done) {
  if constexpr (FAM == 3 || FAM >= 10) {
    if ((((uintptr_t)K | (uintptr_t)U) & 15) || ((uintptr_t)col & 7)) {
      launch_small_fused<0, BR, BC, 4, 4>(st, f, rowPtr, col, K, U, KU, done);
      return;
    }
  }
  if constexpr (FAM == 2) {
    launch_run_async<BR, BC, LPR, P>(st, 1, 0, 0, 0, 0, f, rowPtr, col, K, U, KU, done);
  } else {
    constexpr int rpc = 256 / LPR;
    f.bndCtas = (f.nBnd + rpc - 1) / rpc;
    const int inner = f.mynNo - f.shnNo;
    const int blocks = f.bndCtas + (inner + rpc - 1) / rpc;
    if (blocks <= 0) return;
    spmv_small_fused_kernel<FAM, BR, BC, LPR, P><<<blocks, 256, 0, st>>>(f, rowPtr, col, K, U, KU, done);
  }
}
// SVFSI_SPMV_SMALL = 0: lane-per-block kernels (round 1); 1..8: the configurations below (applied to every
// shape); unset / -1: the per-shape choice measured on a B200 at 10M tets (profiles/r02_spmv_small.md).
static int g_spmv_small = -2;
static int spmv_small_mode() {
  if (g_spmv_small == -2) {
    const char *e = getenv("SVFSI_SPMV_SMALL");
    g_spmv_small = e ? atoi(e) : -1;
  }
  return g_spmv_small;
}
int set_spmv_small(int mode) {
  const int prev = spmv_small_mode();
  g_spmv_small = mode;
  return prev;
}
// per-shape default (mode -1): index = shape class (0: 1x1, 1: 3x3, 2: 3x1, 3: 1x3, 4: everything else)
// Measured at 10.03M tets on a B200 (profiles/r02_spmv_small.md; lane-per-block = 0.56 / 0.58 / 0.60 / 0.53):
//   1x1: hoisted 4 lanes x 4 entries 0.78;  3x3: wide value loads, 8 lanes x 2 blocks, plain loads 0.89;
//   3x1: the same 0.76;  1x3: hoisted 4 x 4 0.71;  the other shapes: hoisted 4 x 4 (not measured at scale)
static const int kSmallDefault[5] = {5, 9, 9, 5, 5};
// CALL(FAM, BR, BC, LPR, P): FAM 0 hoist (P blocks per lane), 1 run (P steps), 2 run-async (P steps).
// Modes: 1, 2 = run; 3, 4, 6 = run-async; 5 = hoist (4 lanes x 4 blocks); 7, 8 = hoist with wide loads (4 x 4,
// 8 x 2; 1x1: plain hoist 8 x 2, 2 x 8).  Returns from the enclosing function after a launch.
#define SMALL_SHAPE(BR, BC, T1, S1, T2, S2, T6, S6)                                               \
  do {                                                                                            \
    if (md == 1) { CALL_(1, BR, BC, T1, S1); return; }                                            \
    if (md == 2) { CALL_(1, BR, BC, T2, S2); return; }                                            \
    if (md == 3) { CALL_(2, BR, BC, T1, S1); return; }                                            \
    if (md == 4) { CALL_(2, BR, BC, T2, S2); return; }                                            \
    if (md == 5) { CALL_(0, BR, BC, 4, 4); return; }                                              \
    if (md == 6) { CALL_(2, BR, BC, T6, S6); return; }                                            \
    if (BR * BC == 1) {                                                                           \
      if (md == 7) { CALL_(0, 1, 1, 8, 2); return; }                                              \
      if (md == 8) { CALL_(0, 1, 1, 2, 8); return; }                                              \
      if (md == 9) { CALL_(14, 1, 1, 4, 4); return; }        /* plain loads */                    \
      if (md == 10) { CALL_(20, 1, 1, 4, 3); return; }       /* entry pairs */                    \
      if (md == 11) { CALL_(21, 1, 1, 4, 3); return; }       /* entry pairs, plain loads */       \
      if (md == 12) { CALL_(20, 1, 1, 8, 2); return; }                                            \
      CALL_(20, 1, 1, 2, 5); return;                                                              \
    }                                                                                             \
    if (md == 7) { CALL_(3, BR, BC, 4, 4); return; }                                              \
    if (md == 8) { CALL_(3, BR, BC, 8, 2); return; }                                              \
    if (md == 9) { CALL_(15, BR, BC, 8, 2); return; }        /* wide values, plain loads */       \
    if (md == 10) { CALL_(13, BR, BC, 8, 2); return; }       /* wide values + wide U */           \
    if (md == 11) { CALL_(17, BR, BC, 8, 2); return; }       /* wide values + wide U, plain */    \
    if (md == 12) { CALL_(13, BR, BC, 4, 4); return; }                                            \
    CALL_(12, BR, BC, 4, 4); return;                         /* 13: wide U only */                \
  } while (0)
#define SMALL_DISPATCH()                                                                          \
  do {                                                                                            \
    const bool ss = (kind == 3 || dof == 1);                                                      \
    const int cls = ss ? 0 : (dof == 3 ? (kind == 0 ? 1 : (kind == 2 ? 2 : 3)) : 4);              \
    int md = spmv_small_mode();                                                                   \
    if (md < 0) md = kSmallDefault[cls];                                                          \
    if (md >= 1 && md <= 13) {                                                                    \
      if (cls == 0) SMALL_SHAPE(1, 1, 4, 4, 8, 2, 16, 1);                                         \
      if (cls == 1) SMALL_SHAPE(3, 3, 16, 9, 32, 5, 8, 17);                                       \
      if (cls == 2) SMALL_SHAPE(3, 1, 8, 6, 16, 3, 4, 12);                                        \
      if (cls == 3) SMALL_SHAPE(1, 3, 8, 6, 16, 3, 4, 12);                                        \
      /* the 2-D shapes and dof = 4 VS / SV: hoisted lane-per-block */                            \
      if (kind == 0 && dof == 2) { CALL_(0, 2, 2, 4, 4); return; }                                \
      if (kind == 1 && dof == 2) { CALL_(0, 1, 2, 4, 4); return; }                                \
      if (kind == 2 && dof == 2) { CALL_(0, 2, 1, 4, 4); return; }                                \
      if (kind == 1 && dof == 4) { CALL_(0, 1, 4, 4, 4); return; }                                \
      if (kind == 2 && dof == 4) { CALL_(0, 4, 1, 4, 4); return; }                                \
    }                                                                                             \
  } while (0)

template <int BR, int BC>
static void launch_generic(cudaStream_t st, int r0, int r1, int r2, int r3, const int *rowPtr,
                           const int *col, const double *K, const double *U, double *KU,
                           const int *done) {
  const int rows = (r1 - r0) + (r3 - r2);
  const int blocks = (rows * 4 + 255) / 256;
  spmv_generic_kernel<BR, BC><<<blocks, 256, 0, st>>>(r0, r1, r2, r3, rowPtr, col, K, U, KU, done);
}

// SVFSI_SPMV_QUAD=1|0 selects the four-lanes-per-row SPARMULVV kernel on the unfused path (default 1;
// measured in profiles/r01_spmv_quad.md).  The fused SpMV + halo-send kernel of the multi-GPU path
// still uses 8 lanes per row.
static int g_spmv_quad = -1;
static bool spmv_quad() {
  if (g_spmv_quad < 0) {
    const char *e = getenv("SVFSI_SPMV_QUAD");
    g_spmv_quad = e ? (atoi(e) != 0) : 1;   // default: quad (0.589 vs 0.648 ms at 10M tets)
  }
  return g_spmv_quad != 0;
}
// SVFSI_SPMV_FUSED_QUAD=1|0: the same choice for the fused SpMV + halo-send kernel of the multi-GPU path
static bool spmv_fused_quad() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("SVFSI_SPMV_FUSED_QUAD");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
}
int spmv_fused_quad_enabled() { return spmv_fused_quad() ? 1 : 0; }
// kernel-variant timings (gpu_time_kernel_): returns the previous setting
int set_spmv_quad(int on) {
  const int prev = spmv_quad() ? 1 : 0;
  g_spmv_quad = on ? 1 : 0;
  return prev;
}

void launch_spmv2(cudaStream_t st, int kind, int dof, int r0, int r1, int r2, int r3,
                  const int *rowPtr, const int *col, const double *K, const double *U, double *KU,
                  const int *done) {
  if (r1 < r0) r1 = r0;
  if (r3 < r2) r3 = r2;
  const int rows = (r1 - r0) + (r3 - r2);
  if (rows <= 0) return;
  count_launch();
  if (kind == 0 && dof == 4) {
    if (spmv_quad()) {
      const int blocks = (int)(((size_t)rows * 4 + 255) / 256);
      launch_pdl(spmv_vv4_quad_kernel, dim3(blocks), dim3(256), 0, st, r0, r1, r2, r3, rowPtr, col, K, U, KU, done);
      return;
    }
    const int blocks = (int)(((size_t)rows * 8 + 255) / 256);
    spmv_vv4_kernel<<<blocks, 256, 0, st>>>(r0, r1, r2, r3, rowPtr, col, (const double2 *)K,
                                            (const double2 *)U, KU, done);
    return;
  }
  if (use_flat()) {
#define FL(BR, BC, T) launch_flat<BR, BC, T>(st, r0, r1, r2, r3, rowPtr, col, K, U, KU, done)
    FLAT_DISPATCH(FL);
#undef FL
  }
  {
    StreamMap mp{r0, r1, r2, r3, 0, 0, 0, 0};
    SpmvFuse none;
    memset(&none, 0, sizeof(none));
    if (stream_dispatch(st, kind, dof, mp, 0, none, rowPtr, col, K, U, KU, done)) return;
  }
#define CALL_(FAM, BR, BC, LPR, P) launch_small<FAM, BR, BC, LPR, P>(st, r0, r1, r2, r3, rowPtr, col, K, U, KU, done)
  SMALL_DISPATCH();
#undef CALL_
#define GEN(BR, BC) launch_generic<BR, BC>(st, r0, r1, r2, r3, rowPtr, col, K, U, KU, done)
  if (kind == 3 || dof == 1) {
    GEN(1, 1);
  } else if (kind == 0) {
    if (dof == 2) GEN(2, 2); else GEN(3, 3);
  } else if (kind == 1) {
    if (dof == 2) GEN(1, 2); else if (dof == 3) GEN(1, 3); else GEN(1, 4);
  } else {
    if (dof == 2) GEN(2, 1); else if (dof == 3) GEN(3, 1); else GEN(4, 1);
  }
#undef GEN
}

template <int BR, int BC>
static void launch_generic_fused(cudaStream_t st, const SpmvFuse &f, int blocks, const int *rowPtr,
                                 const int *col, const double *K, const double *U, double *KU,
                                 const int *done) {
  spmv_generic_fused_kernel<BR, BC><<<blocks, 256, 0, st>>>(f, rowPtr, col, K, U, KU, done);
}
// all rows of this rank in one launch, boundary rows first, halo send fused (see above).
// f.bndCtas is filled in here (it depends on the rows per CTA of the kernel shape).
void launch_spmv_fused(cudaStream_t st, int kind, int dof, SpmvFuse f, const int *rowPtr,
                       const int *col, const double *K, const double *U, double *KU,
                       const int *done, const double *scaleW) {
  count_launch();
  const bool vv4 = (kind == 0 && dof == 4);
  if (!vv4 && use_flat()) {
#define FLF(BR, BC, T) launch_flat_fused<BR, BC, T>(st, f, rowPtr, col, K, U, KU, done)
    FLAT_DISPATCH(FLF);
#undef FLF
  }
  if (!vv4) {
    StreamMap mp{0, f.shnNo, f.mynNo, f.nNo, f.shnNo, f.mynNo, 0, 0};
    if (stream_dispatch(st, kind, dof, mp, 1, f, rowPtr, col, K, U, KU, done)) return;
  }
  if (!vv4) {
#define CALL_(FAM, BR, BC, LPR, P) launch_small_fused<FAM, BR, BC, LPR, P>(st, f, rowPtr, col, K, U, KU, done)
    SMALL_DISPATCH();
#undef CALL_
  }
  const bool quad = vv4 && spmv_fused_quad() && !scaleW;
  const int rpc = (vv4 && !quad) ? 32 : 64;
  f.bndCtas = (f.nBnd + rpc - 1) / rpc;
  const int inner = f.mynNo - f.shnNo;
  const int blocks = f.bndCtas + (inner + rpc - 1) / rpc;
  if (blocks <= 0) return;
  if (quad) {
    spmv_vv4_quad_fused_kernel<<<blocks, 256, 0, st>>>(f, rowPtr, col, K, U, KU, done);
    return;
  }
  if (vv4) {
    if (scaleW)   // first product of a solve: carries PRECONDDIAG's K <- W K W
      launch_pdl(spmv_vv4_fused_kernel<true>, dim3(blocks), dim3(256), 0, st, f, rowPtr, col, (const double2 *)K,
                 (const double2 *)U, KU, done, scaleW);
    else
      launch_pdl(spmv_vv4_fused_kernel<false>, dim3(blocks), dim3(256), 0, st, f, rowPtr, col, (const double2 *)K,
                 (const double2 *)U, KU, done, (const double *)nullptr);
    return;
  }
#define GENF(BR, BC) launch_generic_fused<BR, BC>(st, f, blocks, rowPtr, col, K, U, KU, done)
  if (kind == 3 || dof == 1) {
    GENF(1, 1);
  } else if (kind == 0) {
    if (dof == 2) GENF(2, 2); else GENF(3, 3);
  } else if (kind == 1) {
    if (dof == 2) GENF(1, 2); else if (dof == 3) GENF(1, 3); else GENF(1, 4);
  } else {
    if (dof == 2) GENF(2, 1); else if (dof == 3) GENF(3, 1); else GENF(4, 1);
  }
#undef GENF
}

void launch_spmv(cudaStream_t st, int kind, int dof, int r0, int r1, const int *rowPtr,
                 const int *col, const double *K, const double *U, double *KU, const int *done) {
  launch_spmv2(st, kind, dof, r0, r1, r1, r1, rowPtr, col, K, U, KU, done);
}

// ---------------------------------------------------------------------------
// halo sum, L/INCOMMU.f:56-151
__global__ void pack_kernel(int dof, int nShared, const int *__restrict__ packIdx,
                            const double *__restrict__ R, double *__restrict__ sbuf,
                            const int *done) {
  DONE_GUARD(done);
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nShared * dof) return;
  int s = t / dof, d = t - s * dof;
  sbuf[t] = R[(size_t)packIdx[s] * dof + d];
}

// each unique shared node adds its neighbours' contributions in ascending
// neighbour-rank order (L/INCOMMU.f:91-96) -- deterministic
__global__ void unpack_add_kernel(int dof, int nUniq, const int *__restrict__ uniqNode,
                                  const int *__restrict__ uniqPtr,
                                  const int *__restrict__ uniqSlot,
                                  const double *__restrict__ rbuf, double *__restrict__ R,
                                  const int *done) {
  DONE_GUARD(done);
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nUniq * dof) return;
  int u = t / dof, d = t - u * dof;
  size_t at = (size_t)uniqNode[u] * dof + d;
  double v = R[at];
  for (int k = uniqPtr[u]; k < uniqPtr[u + 1]; k++) v = v + rbuf[(size_t)uniqSlot[k] * dof + d];
  R[at] = v;
}

void launch_pack(cudaStream_t st, int dof, int nShared, const int *packIdx, const double *R,
                 double *sbuf, const int *done) {
  if (nShared <= 0) return;
  count_launch();
  int n = nShared * dof;
  pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(dof, nShared, packIdx, R, sbuf, done);
}
void launch_unpack_add(cudaStream_t st, int dof, int nUniq, const int *uniqNode,
                       const int *uniqPtr, const int *uniqSlot, const double *rbuf, double *R,
                       const int *done) {
  if (nUniq <= 0) return;
  count_launch();
  int n = nUniq * dof;
  unpack_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(dof, nUniq, uniqNode, uniqPtr, uniqSlot,
                                                     rbuf, R, done);
}

// ---------------------------------------------------------------------------
// Peer-memory halo sum and all-reduce: the FSILS_COMMUV / MPI_ALLREDUCE steps written as kernels
// that store straight into the peers' IPC-mapped arenas over NVLink (no NCCL launch, no
// host).  Flags carry a sequence number; two alternating slots make back-to-back exchanges safe
// (a rank can be at most one exchange ahead of a neighbour: it needs the neighbour's previous
// message to get there).
__global__ void __launch_bounds__(256) halo_send_kernel(P2PDev pd, int dof, int nShared, int nNbr,
                                                        const int *__restrict__ packIdx,
                                                        const int *__restrict__ slotNbr,
                                                        const int *__restrict__ nbrRank,
                                                        const int *__restrict__ nbrOff,
                                                        const int *__restrict__ nbrPeerOff,
                                                        const double *__restrict__ R, int seq,
                                                        unsigned int *counter) {
  const int slot = seq & 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nShared * dof) {
    const int s = t / dof, d = t - s * dof;
    const int i = slotNbr[s];
    double *dst = (double *)(pd.peer[nbrRank[i]] + pd.offHalo) + (size_t)slot * pd.haloCap +
                  (size_t)(nbrPeerOff[i] + (s - nbrOff[i])) * dof + d;
    *dst = R[(size_t)packIdx[s] * dof + d];
  }
  // last CTA to finish publishes the flags
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    if (threadIdx.x < nNbr)
      st_flag_sys(flag_ptr(pd.peer[nbrRank[threadIdx.x]], slot, pd.rank), seq);
    if (threadIdx.x == 0) *counter = 0;
  }
}

__global__ void __launch_bounds__(256) halo_recv_add_kernel(P2PDev pd, int dof, int nNbr,
                                                            const int *__restrict__ nbrRank, int nUniq,
                                                            const int *__restrict__ uniqNode,
                                                            const int *__restrict__ uniqPtr,
                                                            const int *__restrict__ uniqSlot,
                                                            double *__restrict__ R, int seq) {
  const int slot = seq & 1;
  if (threadIdx.x < nNbr) wait_flag_sys(flag_ptr(pd.peer[pd.rank], slot, nbrRank[threadIdx.x]), seq, pd);
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nUniq * dof) return;
  const int u = t / dof, d = t - u * dof;
  const double *rbuf = (const double *)(pd.peer[pd.rank] + pd.offHalo) + (size_t)slot * pd.haloCap;
  const size_t at = (size_t)uniqNode[u] * dof + d;
  double v = R[at];
  for (int k = uniqPtr[u]; k < uniqPtr[u + 1]; k++) v = v + __ldcv(rbuf + (size_t)uniqSlot[k] * dof + d);
  R[at] = v;
}

void launch_halo_send(cudaStream_t st, const P2PDev &pd, int dof, int nShared, int nNbr,
                      const int *packIdx, const int *slotNbr, const int *nbrRank, const int *nbrOff,
                      const int *nbrPeerOff, const double *R, int seq, unsigned int *counter) {
  count_launch();
  const int n = nShared * dof;
  halo_send_kernel<<<(n + 255) / 256, 256, 0, st>>>(pd, dof, nShared, nNbr, packIdx, slotNbr, nbrRank,
                                                    nbrOff, nbrPeerOff, R, seq, counter);
}
void launch_halo_recv_add(cudaStream_t st, const P2PDev &pd, int dof, int nNbr, const int *nbrRank,
                          int nUniq, const int *uniqNode, const int *uniqPtr, const int *uniqSlot,
                          double *R, int seq) {
  count_launch();
  const int n = nUniq * dof;
  halo_recv_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(pd, dof, nNbr, nbrRank, nUniq, uniqNode, uniqPtr,
                                                        uniqSlot, R, seq);
}

// One column of the Arnoldi/Givens recurrence, L/GMRES.f:342-366 (scalar part), as the tail of
// the reduction kernels below.  sh[0..i] = <u_j, u_{i+1}> (already summed over blocks and ranks)
// in shared memory.  The rotations are a sequential recurrence: thread 0 runs it on shared memory
// (c, s staged there by the whole block) instead of chasing global-memory latencies.
__device__ void column_step_block(const ColArgs &a, double *sh, double *sc, double *ss, double *sv) {
  const int i = a.i;
  const bool active = (a.ctl->done == 0);
  for (int j = threadIdx.x; j < i; j += blockDim.x) {
    sv[j] = sh[j];
    if (j < i - 1) { sc[j] = a.c[j]; ss[j] = a.s[j]; }
  }
  __syncthreads();
  if (threadIdx.x == 0 && active) {
    double hh = sh[i];
    for (int j = 0; j < i; j++) hh = hh - sh[j] * sh[j];
    hh = sqrt(fabs(hh));
    a.ctl->inv = 1.0 / hh;
    sh[i] = hh;
    for (int j = 0; j < i - 1; j++) {
      const double tmp = sc[j] * sh[j] + ss[j] * sh[j + 1];
      sh[j + 1] = -ss[j] * sh[j] + sc[j] * sh[j + 1];
      sh[j] = tmp;
    }
    const double tmp = sqrt(sh[i - 1] * sh[i - 1] + sh[i] * sh[i]);
    const double ci = sh[i - 1] / tmp, si = sh[i] / tmp;
    a.c[i - 1] = ci;
    a.s[i - 1] = si;
    sh[i - 1] = tmp;
    sh[i] = 0.0;
    const double e0 = a.err[i - 1];
    a.err[i] = -si * e0;
    a.err[i - 1] = ci * e0;
    a.ctl->ilast = i;
    a.ctl->itr += 1;
    if (fabs(-si * e0) < a.ctl->eps) {
      a.ctl->suc = 1;
      a.ctl->done = 1;
    }
  }
  __syncthreads();
  if (active) {
    double *hc = a.h + (size_t)(i - 1) * (a.sD + 1);  // h(:,i), 0-based rows
    for (int j = threadIdx.x; j <= i; j += blockDim.x) {
      hc[j] = sh[j];
      if (j < i) a.coef[j] = sv[j];
    }
  }
  // publish the stop flag AS OF THIS COLUMN to the host (mapped pinned memory), see solver_int.h
  if (threadIdx.x == 0 && a.pubFlag) *a.pubFlag = (a.seq << 1) | (a.ctl->done ? 1 : 0);   // ONE word: no fence
}

// single rank: block sums of the multi-dot partials + the column step in ONE kernel
__global__ void __launch_bounds__(512) reduce_column_kernel(const double *__restrict__ partial, int nblk,
                                                            int k, double *__restrict__ out, ColArgs col) {
  __shared__ double sh[kArMax], sc[kArMax], ss[kArMax], sv[kArMax];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (col.ctl->done == 0) {
    for (int j = wid; j < k; j += nw) {
      double v = 0.0;
      for (int b = lane; b < nblk; b += 32) v += partial[(size_t)j * nblk + b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) { sh[j] = v; out[j] = v; }
    }
  }
  __syncthreads();
  column_step_block(col, sh, sc, ss, sv);
}

// reduce the multi-dot partials (optional) and all-reduce k <= kArMax scalars in ONE kernel:
// every rank stores its k values into every peer's mailbox, flags, waits for all peers, and sums
// in rank order -- so all ranks obtain bit-identical results (the Krylov control flow depends on it).
__global__ void __launch_bounds__(512) p2p_allreduce_kernel(P2PDev pd, const double *__restrict__ partial,
                                                            int nblk, int k, double *__restrict__ out,
                                                            int seq, ColArgs col) {
  __shared__ double mine[kArMax];
  __shared__ double sc[kArMax], ss[kArMax], sv[kArMax];
  const int slot = seq & 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (partial) {
    for (int j = wid; j < k; j += nw) {
      double v = 0.0;
      for (int b = lane; b < nblk; b += 32) v += partial[(size_t)j * nblk + b];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) mine[j] = v;
    }
  } else {
    for (int j = threadIdx.x; j < k; j += blockDim.x) mine[j] = out[j];
  }
  __syncthreads();
  for (int p = 0; p < pd.nranks; p++) {
    double *mb = (double *)(pd.peer[p] + pd.offMail) + ((size_t)slot * pd.nranks + pd.rank) * kArMax;
    for (int j = threadIdx.x; j < k; j += blockDim.x) mb[j] = mine[j];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < pd.nranks) st_flag_sys(flag_ptr(pd.peer[threadIdx.x], 2 + slot, pd.rank), seq);
  if (threadIdx.x < pd.nranks) wait_flag_sys(flag_ptr(pd.peer[pd.rank], 2 + slot, threadIdx.x), seq, pd);
  __syncthreads();
  const double *mb = (const double *)(pd.peer[pd.rank] + pd.offMail) + (size_t)slot * pd.nranks * kArMax;
  __syncthreads();   // everyone is done reading mine[] as the send buffer
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    double v = 0.0;
    for (int r = 0; r < pd.nranks; r++) v += __ldcv(mb + (size_t)r * kArMax + j);
    out[j] = v;
    mine[j] = v;
  }
  if (col.ctl) {     // the Givens column step of GMRES rides on the same kernel
    __syncthreads();
    column_step_block(col, mine, sc, ss, sv);
  }
}

void launch_p2p_allreduce(cudaStream_t st, const P2PDev &pd, const double *partial, int nblk, int k,
                          double *out, int seq, const ColArgs *col) {
  count_launch();
  ColArgs none;
  memset(&none, 0, sizeof(none));
  p2p_allreduce_kernel<<<1, 512, 0, st>>>(pd, partial, nblk, k, out, seq, col ? *col : none);
}
void launch_reduce_column(cudaStream_t st, const double *partial, int nblk, int k, double *out,
                          const ColArgs &col) {
  count_launch();
  reduce_column_kernel<<<1, 512, 0, st>>>(partial, nblk, k, out, col);
}

// ---------------------------------------------------------------------------
// fused multi-dot: one pass over w and k basis vectors (replaces the i+1 separate
// FSILS_NCDOTV passes of L/GMRES.f:337-339).  Deterministic two-stage reduction.
static constexpr int kDotBlocks = kSMs * 4;
static constexpr int kDotThreads = 256;
int multidot_nblk() { return kDotBlocks; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int JT>
__device__ __forceinline__ void multidot_tile(const double *__restrict__ U, size_t stride,
                                              const double *__restrict__ w, size_t n, int j0,
                                              double *__restrict__ partial, double *smem) {
  double acc[JT];
#pragma unroll
  for (int jj = 0; jj < JT; jj++) acc[jj] = 0.0;
  const size_t n2 = n >> 1;
  const double2 *w2 = (const double2 *)w;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2;
       e += (size_t)gridDim.x * blockDim.x) {
    const double2 wv = __ldg(w2 + e);
#pragma unroll
    for (int jj = 0; jj < JT; jj++) {
      const double2 uv = __ldcs((const double2 *)(U + (size_t)(j0 + jj) * stride) + e);
      acc[jj] = fma(uv.x, wv.x, acc[jj]);
      acc[jj] = fma(uv.y, wv.y, acc[jj]);
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const double wv = w[n - 1];
#pragma unroll
    for (int jj = 0; jj < JT; jj++) acc[jj] = fma(U[(size_t)(j0 + jj) * stride + n - 1], wv, acc[jj]);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int jj = 0; jj < JT; jj++) {
    double v = warp_sum(acc[jj]);
    if (lane == 0) smem[wid * JT + jj] = v;
  }
  __syncthreads();
  if (threadIdx.x < JT) {
    double v = 0.0;
    for (int wq = 0; wq < kDotThreads / 32; wq++) v += smem[wq * JT + threadIdx.x];
    partial[(size_t)(j0 + threadIdx.x) * gridDim.x + blockIdx.x] = v;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kDotThreads) multidot_kernel(const double *__restrict__ U, size_t stride,
                                                               const double *__restrict__ w, size_t n,
                                                               int k, double *__restrict__ partial,
                                                               const int *done) {
  DONE_GUARD(done);
  __shared__ double smem[(kDotThreads / 32) * 8];
  int j0 = 0;
  while (k - j0 >= 8) { multidot_tile<8>(U, stride, w, n, j0, partial, smem); j0 += 8; }
  if (k - j0 >= 4) { multidot_tile<4>(U, stride, w, n, j0, partial, smem); j0 += 4; }
  if (k - j0 >= 2) { multidot_tile<2>(U, stride, w, n, j0, partial, smem); j0 += 2; }
  if (k - j0 >= 1) { multidot_tile<1>(U, stride, w, n, j0, partial, smem); j0 += 1; }
}

__global__ void reduce_partials_kernel(const double *__restrict__ partial, int nblk,
                                       double *__restrict__ out, const int *done) {
  DONE_GUARD(done);
  __shared__ double smem[8];
  const int j = blockIdx.x;
  double v = 0.0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) v += partial[(size_t)j * nblk + b];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int wq = 0; wq < (int)(blockDim.x >> 5); wq++) t += smem[wq];
    out[j] = t;
  }
}

// ---------------------------------------------------------------------------
// The fused column kernel (see kernels.h: HaloRecv / DotTail).
//   CTAs [0, nRecv): wait for the neighbours' halo flags, add the received contributions to the shared rows of
//     w (ascending neighbour order per node, L/INCOMMU.f:91-96), then take the inner products over THEIR chunk
//     of the rows shared with lower ranks [0, shnNo) -- rows >= mynNo are not owned and do not enter the dots.
//   CTAs [nRecv, grid): inner products over the interior rows [shnNo, mynNo), grid-stride.
//   last CTA (ticket): block sums, peer all-reduce, Givens column.
template <int JT>
__device__ __forceinline__ void mdf_tile(const double *U, size_t stride, const double *w,
                                         size_t lo, size_t hi, unsigned crank, unsigned ccount, int j0,
                                         double *__restrict__ partial, double *smem) {
  double acc[JT];
#pragma unroll
  for (int jj = 0; jj < JT; jj++) acc[jj] = 0.0;
  if (hi > lo) {
    const size_t lo2 = lo + (lo & 1);                 // vector body on 16-byte aligned pairs
    const size_t np = hi > lo2 ? (hi - lo2) >> 1 : 0;
    const double2 *w2 = (const double2 *)(w + lo2);
    // narrow tiles carry few loads per trip: unroll so that >= 8 loads are in flight per thread
    constexpr int UNR = JT >= 4 ? 1 : (JT >= 2 ? 2 : 4);
    const size_t step = (size_t)ccount * blockDim.x;
    size_t e = (size_t)crank * blockDim.x + threadIdx.x;
    for (; e + (UNR - 1) * step < np; e += UNR * step) {
      double2 wv[UNR], uv[UNR][JT];
#pragma unroll
      for (int r = 0; r < UNR; r++) {
        wv[r] = w2[e + r * step];   // plain load: this CTA may have just written these rows (halo receive)
#pragma unroll
        for (int jj = 0; jj < JT; jj++)
          uv[r][jj] = __ldcs((const double2 *)(U + (size_t)(j0 + jj) * stride + lo2) + e + r * step);
      }
#pragma unroll
      for (int r = 0; r < UNR; r++)
#pragma unroll
        for (int jj = 0; jj < JT; jj++) {
          acc[jj] = fma(uv[r][jj].x, wv[r].x, acc[jj]);
          acc[jj] = fma(uv[r][jj].y, wv[r].y, acc[jj]);
        }
    }
    for (; e < np; e += step) {
      const double2 wv = w2[e];
#pragma unroll
      for (int jj = 0; jj < JT; jj++) {
        const double2 uv = __ldcs((const double2 *)(U + (size_t)(j0 + jj) * stride + lo2) + e);
        acc[jj] = fma(uv.x, wv.x, acc[jj]);
        acc[jj] = fma(uv.y, wv.y, acc[jj]);
      }
    }
    if (crank == 0 && threadIdx.x == 0) {             // unpaired first / last element of the range
      if (lo & 1) {
        const double wv = w[lo];
#pragma unroll
        for (int jj = 0; jj < JT; jj++) acc[jj] = fma(U[(size_t)(j0 + jj) * stride + lo], wv, acc[jj]);
      }
      if (lo2 + 2 * np < hi) {
        const double wv = w[hi - 1];
#pragma unroll
        for (int jj = 0; jj < JT; jj++) acc[jj] = fma(U[(size_t)(j0 + jj) * stride + hi - 1], wv, acc[jj]);
      }
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int jj = 0; jj < JT; jj++) {
    double v = warp_sum(acc[jj]);
    if (lane == 0) smem[wid * JT + jj] = v;
  }
  __syncthreads();
  if (threadIdx.x < JT) {
    double v = 0.0;
    for (int wq = 0; wq < kDotThreads / 32; wq++) v += smem[wq * JT + threadIdx.x];
    partial[(size_t)(j0 + threadIdx.x) * gridDim.x + blockIdx.x] = v;
  }
  __syncthreads();
}

// (U is not __restrict__: in the Arnoldi loop U_{k-1} IS w, and the receive CTAs write w before reading it)
template <int MINB>
__global__ void __launch_bounds__(kDotThreads, MINB) multidot_fused_kernel(const double *U, size_t stride,
                                                                     double *w, size_t nOwned, int k,
                                                                     double *partial, const int *done,
                                                                     HaloRecv hr, DotTail tail, int nRecv) {
  __shared__ double smem[(kDotThreads / 32) * 8];
  __shared__ double sh[kArMax], sc[kArMax], ss[kArMax], sv[kArMax];
  __shared__ bool last;
  PDL_TRIGGER();
  PDL_WAIT();
  const bool skip = (done != nullptr && *(volatile const int *)done != 0);
  size_t lo, hi;
  unsigned crank, ccount;
  if (tail.trace && blockIdx.x == 0 && threadIdx.x == 0) tail.trace[0] = global_ns();   // kernel start
  if ((int)blockIdx.x < nRecv) {
    // ---- halo receive for this CTA's share of the shared rows (the flag protocol runs even when `skip`)
    const int slot = hr.seq & 1;
    const P2PDev &pd = tail.pd;
    if ((int)threadIdx.x < hr.nNbr)
      wait_flag_sys(flag_ptr(pd.peer[pd.rank], slot, hr.nbrRank[threadIdx.x]), hr.seq, pd);
    __syncthreads();
    if (tail.trace && blockIdx.x == 0 && threadIdx.x == 0) tail.trace[1] = global_ns();   // neighbours' halos in
    const double *rbuf = (const double *)(pd.peer[pd.rank] + pd.offHalo) + (size_t)slot * pd.haloCap;
    const int dof = hr.dof;
    // rows [0, shnNo) are the first shnNo unique shared nodes, in order: contiguous chunks of them per CTA
    const size_t L = (size_t)hr.shnNo * dof;
    size_t chunk = (L + nRecv - 1) / nRecv;
    chunk += chunk & 1;
    lo = (size_t)blockIdx.x * chunk;
    hi = lo + chunk;
    if (lo > L) lo = L;
    if (hi > L) hi = L;
    if (!skip) {
      for (size_t t = lo + threadIdx.x; t < hi; t += blockDim.x) {
        const int u = (int)(t / dof), d = (int)(t - (size_t)u * dof);
        double v = w[t];
        for (int q = hr.uniqPtr[u]; q < hr.uniqPtr[u + 1]; q++) v = v + __ldcv(rbuf + (size_t)hr.uniqSlot[q] * dof + d);
        w[t] = v;
      }
      // rows shared with higher ranks [mynNo, nNo): not owned, no inner product -- spread over the receive CTAs
      const size_t H = (size_t)(hr.nUniq - hr.shnNo) * dof;
      for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < H; t += (size_t)nRecv * blockDim.x) {
        const int u = hr.shnNo + (int)(t / dof), d = (int)(t % dof);
        const size_t at = (size_t)hr.uniqNode[u] * dof + d;
        double v = w[at];
        for (int q = hr.uniqPtr[u]; q < hr.uniqPtr[u + 1]; q++) v = v + __ldcv(rbuf + (size_t)hr.uniqSlot[q] * dof + d);
        w[at] = v;
      }
    }
    __syncthreads();
    if (hi > nOwned) hi = nOwned;
    if (lo > hi) lo = hi;
    crank = 0; ccount = 1;
  } else {
    lo = hr.on ? (size_t)hr.shnNo * hr.dof : 0;
    if (lo > nOwned) lo = nOwned;
    hi = nOwned;
    crank = blockIdx.x - nRecv; ccount = gridDim.x - nRecv;
  }
  if (!skip) {
    // k vectors in ceil(k / 8) passes of (nearly) equal width: 9 = 5 + 4, not 8 + 1 (a one-vector pass runs at
    // half the bandwidth of a wide one)
    const int nt = (k + 7) / 8;
    int j0 = 0;
    for (int t = 0; t < nt; t++) {
      const int jt = (k - j0 + (nt - t) - 1) / (nt - t);
      switch (jt) {
        case 8: mdf_tile<8>(U, stride, w, lo, hi, crank, ccount, j0, partial, smem); break;
        case 7: mdf_tile<7>(U, stride, w, lo, hi, crank, ccount, j0, partial, smem); break;
        case 6: mdf_tile<6>(U, stride, w, lo, hi, crank, ccount, j0, partial, smem); break;
        case 5: mdf_tile<5>(U, stride, w, lo, hi, crank, ccount, j0, partial, smem); break;
        case 4: mdf_tile<4>(U, stride, w, lo, hi, crank, ccount, j0, partial, smem); break;
        case 3: mdf_tile<3>(U, stride, w, lo, hi, crank, ccount, j0, partial, smem); break;
        case 2: mdf_tile<2>(U, stride, w, lo, hi, crank, ccount, j0, partial, smem); break;
        default: mdf_tile<1>(U, stride, w, lo, hi, crank, ccount, j0, partial, smem); break;
      }
      j0 += jt;
    }
  }
  // ---- ticket: the last CTA to arrive owns the tail
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(tail.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *tail.counter = 0;
  __threadfence();
  if (tail.trace && threadIdx.x == 0) tail.trace[2] = global_ns();   // every CTA of this rank has finished
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nblk = gridDim.x;
  // block sums: the loads of a row first (nblk <= 444: at most 14 per lane, all in flight), then the adds in a
  // fixed order
  for (int j = wid; j < k; j += nw) {
    double pv[14];
#pragma unroll
    for (int q = 0; q < 14; q++) {
      const int b = lane + 32 * q;
      pv[q] = (!skip && b < nblk) ? __ldcg(partial + (size_t)j * nblk + b) : 0.0;
    }
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 14; q++) v += pv[q];
    v = warp_sum(v);
    if (lane == 0) sh[j] = v;
  }
  __syncthreads();
  if (tail.nranks > 1) {
    // All-reduce with the flag IN the data (the "LL" protocol of NCCL): every value travels as one 16-byte store
    // (low word, seq, high word, seq) into the mailbox of every peer; a receiver spins on the two sequence words
    // of each value.  No system-scope fence, no separate flag store, no second round trip: the latency is ONE
    // NVLink store.  8-byte halves carry their own tag, so the protocol does not rely on 16-byte atomicity.
    // Values are summed in rank order: bit-identical on all ranks.
    const P2PDev &pd = tail.pd;
    const int slot = tail.arSeq & 1;
    const unsigned tag = (unsigned)tail.arSeq;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(sh[j]);
      const uint4 pkt = make_uint4((unsigned)(bits & 0xffffffffull), tag, (unsigned)(bits >> 32), tag);
      for (int p = 0; p < pd.nranks; p++) {
        uint4 *mb = (uint4 *)(pd.peer[p] + pd.offMailLL) + ((size_t)slot * pd.nranks + pd.rank) * kArMax;
        asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(mb + j), "r"(pkt.x), "r"(pkt.y),
                     "r"(pkt.z), "r"(pkt.w) : "memory");
      }
    }
    if (tail.trace && threadIdx.x == 0) tail.trace[7] = global_ns();   // own contribution on its way to the peers
    const uint4 *mbIn = (const uint4 *)(pd.peer[pd.rank] + pd.offMailLL) + (size_t)slot * pd.nranks * kArMax;
    const int e0 = *(volatile int *)pd.errDev;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      double v = 0.0;
      for (int r = 0; r < pd.nranks; r++) {
        const uint4 *src = mbIn + (size_t)r * kArMax + j;
        uint4 pkt;
        unsigned long long t0 = 0;
        unsigned it = 0;
        for (;;) {
          asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(pkt.x), "=r"(pkt.y), "=r"(pkt.z), "=r"(pkt.w) : "l"(src) : "memory");
          if ((pkt.y == tag && pkt.w == tag) || e0) break;
          if ((++it & 255u) == 0) {     // bounded like every peer wait (wait_flag_sys)
            if (t0 == 0) t0 = global_ns();
            if (*(volatile int *)pd.errDev != 0) break;
            if (global_ns() - t0 > (unsigned long long)pd.timeoutNs) {
              *(volatile int *)pd.errDev = 1;
              *pd.errHost = 1;
              break;
            }
          }
        }
        v += __longlong_as_double((long long)(((unsigned long long)pkt.z << 32) | pkt.x));
      }
      sh[j] = v;
    }
    __syncthreads();
    if (tail.trace && threadIdx.x == 0) tail.trace[3] = global_ns();   // every peer's contribution has arrived
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) tail.out[j] = sh[j];
  if (tail.col.ctl) {
    __syncthreads();
    column_step_block(tail.col, sh, sc, ss, sv);
  }
  if (tail.trace && threadIdx.x == 0) tail.trace[4] = global_ns();   // end of the column kernel
}

void launch_multidot_fused(cudaStream_t st, const double *U, size_t stride, double *w, size_t nOwned,
                           int k, double *partial, const int *done, const HaloRecv &hr,
                           const DotTail &tail) {
  count_launch();
  int nRecv = 0;
  if (hr.on) {
    const size_t ent = (size_t)hr.nUniq * hr.dof;
    nRecv = (int)((ent + 2047) / 2048);           // ~8 entries per thread
    if (nRecv < 1) nRecv = 1;
    if (nRecv > 32) nRecv = 32;
  }
  // 88 registers: two CTAs per SM without spills (measured: 7.0 TB/s at k = 34; capped to 80 registers for three
  // CTAs per SM it spills and drops to 5.1 TB/s, profiles/r02_dot_fused.md); ONE wave of 2 x 148 CTAs
  static int minb = -1;
  if (minb < 0) {
    const char *e = getenv("SVFSI_DOT_MINB");
    minb = e ? atoi(e) : 2;
  }
  if (minb == 3)
    multidot_fused_kernel<3><<<kSMs * 3, kDotThreads, 0, st>>>(U, stride, w, nOwned, k, partial, done, hr,
                                                               tail, nRecv);
  else
    launch_pdl(multidot_fused_kernel<2>, dim3(kSMs * 2), dim3(kDotThreads), 0, st, U, stride, w, nOwned, k, partial,
               done, hr, tail, nRecv);
}

void launch_multidot(cudaStream_t st, const double *U, size_t stride, const double *w, size_t n,
                     int k, double *partial, const int *done) {
  if (k <= 0) return;
  count_launch();
  multidot_kernel<<<kDotBlocks, kDotThreads, 0, st>>>(U, stride, w, n, k, partial, done);
}
void launch_reduce_partials(cudaStream_t st, const double *partial, int k, double *out,
                            const int *done) {
  if (k <= 0) return;
  count_launch();
  reduce_partials_kernel<<<k, 256, 0, st>>>(partial, kDotBlocks, out, done);
}

// fused multi-axpy + scale: w = (w - sum_j coef[j] U_j) * scale, the i OMPSUMV
// passes and the OMPMULV of L/GMRES.f:342-349 in one pass, same j order.
__global__ void __launch_bounds__(256) multi_axpy_scale_kernel(const double *__restrict__ U, size_t stride,
                                                               double *__restrict__ w, size_t n, int k,
                                                               const double *__restrict__ coef,
                                                               const double *__restrict__ scale,
                                                               const int *done) {
  PDL_TRIGGER();
  PDL_WAIT();
  DONE_GUARD(done);
  extern __shared__ double sc[];
  for (int j = threadIdx.x; j < k; j += blockDim.x) sc[j] = coef[j];
  __syncthreads();
  const double s = scale ? *scale : 1.0;
  const size_t n2 = n >> 1;
  double2 *w2 = (double2 *)w;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2;
       e += (size_t)gridDim.x * blockDim.x) {
    double2 v = w2[e];
#pragma unroll 4
    for (int j = 0; j < k; j++) {
      const double2 uv = __ldcs((const double2 *)(U + (size_t)j * stride) + e);
      v.x = fma(-sc[j], uv.x, v.x);
      v.y = fma(-sc[j], uv.y, v.y);
    }
    v.x *= s;
    v.y *= s;
    w2[e] = v;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double v = w[n - 1];
    for (int j = 0; j < k; j++) v = fma(-sc[j], U[(size_t)j * stride + n - 1], v);
    w[n - 1] = v * s;
  }
}

void launch_multi_axpy_scale(cudaStream_t st, const double *U, size_t stride, double *w, size_t n,
                             int k, const double *coef, const double *scale, const int *done) {
  count_launch();
  launch_pdl(multi_axpy_scale_kernel, dim3(kSMs * 8), dim3(256), (size_t)(k > 0 ? k : 1) * sizeof(double), st, U,
             stride, w, n, k, coef, scale, done);
}

// X += sum_{j<k} y[j] U_j with k read from the device (number of Krylov vectors
// actually built), L/GMRES.f:377-380
__global__ void __launch_bounds__(256) multi_axpy_acc_kernel(const double *__restrict__ U, size_t stride,
                                                             double *__restrict__ X, size_t n,
                                                             const int *__restrict__ kdev, int kmax,
                                                             const double *__restrict__ y) {
  extern __shared__ double sc[];
  const int k = min(*kdev, kmax);
  for (int j = threadIdx.x; j < k; j += blockDim.x) sc[j] = y[j];
  __syncthreads();
  const size_t n2 = n >> 1;
  double2 *x2 = (double2 *)X;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2;
       e += (size_t)gridDim.x * blockDim.x) {
    double2 v = x2[e];
#pragma unroll 4
    for (int j = 0; j < k; j++) {
      const double2 uv = __ldcs((const double2 *)(U + (size_t)j * stride) + e);
      v.x = fma(sc[j], uv.x, v.x);
      v.y = fma(sc[j], uv.y, v.y);
    }
    x2[e] = v;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double v = X[n - 1];
    for (int j = 0; j < k; j++) v = fma(sc[j], U[(size_t)j * stride + n - 1], v);
    X[n - 1] = v;
  }
}

void launch_multi_axpy_acc(cudaStream_t st, const double *U, size_t stride, double *X, size_t n,
                           const int *kdev, int kmax, const double *y) {
  count_launch();
  multi_axpy_acc_kernel<<<kSMs * 8, 256, (size_t)(kmax > 0 ? kmax : 1) * sizeof(double), st>>>(
      U, stride, X, n, kdev, kmax, y);
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vecop_kernel(int op, double *__restrict__ a, const double *__restrict__ b,
                                                    const double *__restrict__ c, size_t n,
                                                    const double *__restrict__ sdev, double sh,
                                                    const int *done) {
  DONE_GUARD(done);
  const double s = sdev ? *sdev : sh;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    double v;
    switch (op) {
      case VOP_COPY: v = b[e]; break;
      case VOP_SUB_FROM: v = b[e] - a[e]; break;
      case VOP_SCALE_DEV: case VOP_SCALE: v = a[e] * s; break;
      case VOP_DIV_DEV: v = a[e] / s; break;
      case VOP_AXPY_DEV: case VOP_AXPY: v = a[e] + s * b[e]; break;
      case VOP_AXMY_DEV: v = a[e] - s * b[e]; break;
      case VOP_MUL: v = a[e] * b[e]; break;
      case VOP_ZERO: v = 0.0; break;
      case VOP_XPBY_DEV: v = b[e] + s * a[e]; break;
      case VOP_SUB: v = b[e] - c[e]; break;
      default: v = a[e];
    }
    a[e] = v;
  }
}

void launch_vecop(cudaStream_t st, int op, double *a, const double *b, const double *c, size_t n,
                  const double *sdev, double shost, const int *done) {
  if (n == 0) return;
  count_launch();
  size_t want = (n + 255) / 256;
  int blocks = (int)(want < (size_t)kSMs * 8 ? want : (size_t)kSMs * 8);
  vecop_kernel<<<blocks, 256, 0, st>>>(op, a, b, c, n, sdev, shost, done);
}

__global__ void split_mc_kernel(int nNo, int dof, const double *__restrict__ R,
                                double *__restrict__ Rm, double *__restrict__ Rc) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nNo) return;
  const int nsd = dof - 1;
  for (int d = 0; d < nsd; d++) Rm[(size_t)a * nsd + d] = R[(size_t)a * dof + d];
  Rc[a] = R[(size_t)a * dof + nsd];
}
__global__ void join_mc_kernel(int nNo, int dof, const double *__restrict__ Rm,
                               const double *__restrict__ Rc, double *__restrict__ R) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nNo) return;
  const int nsd = dof - 1;
  for (int d = 0; d < nsd; d++) R[(size_t)a * dof + d] = Rm[(size_t)a * nsd + d];
  R[(size_t)a * dof + nsd] = Rc[a];
}
void launch_split_mc(cudaStream_t st, int nNo, int dof, const double *R, double *Rm, double *Rc) {
  count_launch();
  split_mc_kernel<<<(nNo + 255) / 256, 256, 0, st>>>(nNo, dof, R, Rm, Rc);
}
void launch_join_mc(cudaStream_t st, int nNo, int dof, const double *Rm, const double *Rc,
                    double *R) {
  count_launch();
  join_mc_kernel<<<(nNo + 255) / 256, 256, 0, st>>>(nNo, dof, Rm, Rc, R);
}

// ---------------------------------------------------------------------------
__global__ void perm_scatter_kernel(int n, int m, const int *__restrict__ perm,
                                    const double *__restrict__ src, double *__restrict__ dst) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * m) return;
  int a = (int)(t / m), d = (int)(t - (size_t)a * m);
  dst[(size_t)perm[a] * m + d] = src[t];
}
__global__ void perm_gather_kernel(int n, int m, const int *__restrict__ perm,
                                   const double *__restrict__ src, double *__restrict__ dst) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * m) return;
  int a = (int)(t / m), d = (int)(t - (size_t)a * m);
  dst[t] = src[(size_t)perm[a] * m + d];
}
void launch_perm_scatter(cudaStream_t st, int n, int m, const int *perm, const double *src,
                         double *dst) {
  if (n <= 0) return;
  count_launch();
  size_t tot = (size_t)n * m;
  perm_scatter_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, m, perm, src, dst);
}
void launch_perm_gather(cudaStream_t st, int n, int m, const int *perm, const double *src,
                        double *dst) {
  if (n <= 0) return;
  count_launch();
  size_t tot = (size_t)n * m;
  perm_gather_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n, m, perm, src, dst);
}

// ---------------------------------------------------------------------------
// PRECONDDIAG pieces, L/PRECOND.f:66-142
__global__ void diag_extract_kernel(int nNo, int dof, const int *__restrict__ diag,
                                    const double *__restrict__ Val, double *__restrict__ W) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nNo * dof) return;
  int a = t / dof, i = t - a * dof;
  W[t] = Val[(size_t)diag[a] * dof * dof + i * dof + i];
}
__global__ void w_finalize_kernel(size_t n, double *__restrict__ W) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double w = W[t];
  if (w == 0.0) w = 1.0;
  W[t] = 1.0 / sqrt(fabs(w));
}
__global__ void w_dirichlet_kernel(int nFaceNo, int fdof, int dof, const int *__restrict__ glob,
                                   const double *__restrict__ val, double *__restrict__ W) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nFaceNo) return;
  const int m = fdof < dof ? fdof : dof;
  for (int i = 0; i < m; i++) W[(size_t)glob[a] * dof + i] *= val[(size_t)a * fdof + i];
}
// K = (W_row K) W_col fused (PREMUL then POSMUL, same multiplication order per entry)
__global__ void __launch_bounds__(256) scale_val4_kernel(int nnz, const int *__restrict__ rowOf,
                                                          const int *__restrict__ col,
                                                          const double *__restrict__ Wrow,
                                                          const double *__restrict__ Wcol,
                                                          double2 *__restrict__ Val) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nnz * 8) return;
  const int p = (int)(t >> 3), q = (int)(t & 7);
  const int i = q >> 1, k0 = (q & 1) * 2;
  const double wr = __ldg(Wrow + (size_t)__ldg(rowOf + p) * 4 + i);
  const double2 wc = __ldg((const double2 *)(Wcol + (size_t)__ldg(col + p) * 4 + k0));
  double2 v = Val[t];
  v.x = (v.x * wr) * wc.x;
  v.y = (v.y * wr) * wc.y;
  Val[t] = v;
}
__global__ void scale_val_generic_kernel(int nnz, int dof, const int *__restrict__ rowOf,
                                         const int *__restrict__ col,
                                         const double *__restrict__ Wrow,
                                         const double *__restrict__ Wcol, double *__restrict__ Val) {
  const int dd = dof * dof;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nnz * dd) return;
  const int p = (int)(t / dd), r = (int)(t - (size_t)p * dd);
  const int i = r / dof, k = r - i * dof;
  Val[t] = (Val[t] * Wrow[(size_t)rowOf[p] * dof + i]) * Wcol[(size_t)col[p] * dof + k];
}
__global__ void face_valM_kernel(int nFaceNo, int fdof, int dof, const int *__restrict__ glob,
                                 const double *__restrict__ val, const double *__restrict__ W,
                                 double *__restrict__ valM) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nFaceNo) return;
  const int m = fdof < dof ? fdof : dof;
  for (int i = 0; i < m; i++)
    valM[(size_t)a * fdof + i] = val[(size_t)a * fdof + i] * W[(size_t)glob[a] * dof + i];
}

void launch_diag_extract(cudaStream_t st, int nNo, int dof, const int *diag, const double *Val,
                         double *W) {
  count_launch();
  diag_extract_kernel<<<(nNo * dof + 255) / 256, 256, 0, st>>>(nNo, dof, diag, Val, W);
}
void launch_w_finalize(cudaStream_t st, size_t n, double *W) {
  count_launch();
  w_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, W);
}
void launch_w_dirichlet(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob,
                        const double *val, double *W) {
  if (nFaceNo <= 0) return;
  count_launch();
  w_dirichlet_kernel<<<(nFaceNo + 255) / 256, 256, 0, st>>>(nFaceNo, fdof, dof, glob, val, W);
}
void launch_scale_val2(cudaStream_t st, int nnz, int dof, const int *rowOf, const int *col,
                       const double *Wrow, const double *Wcol, double *Val) {
  count_launch();
  if (dof == 4) {
    size_t tot = (size_t)nnz * 8;
    scale_val4_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(nnz, rowOf, col, Wrow, Wcol,
                                                                     (double2 *)Val);
  } else {
    size_t tot = (size_t)nnz * dof * dof;
    scale_val_generic_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(nnz, dof, rowOf, col,
                                                                            Wrow, Wcol, Val);
  }
}
void launch_scale_val(cudaStream_t st, int nnz, int dof, const int *rowOf, const int *col,
                      const double *W, double *Val) {
  launch_scale_val2(st, nnz, dof, rowOf, col, W, W, Val);
}
void launch_face_valM(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob,
                      const double *val, const double *W, double *valM) {
  if (nFaceNo <= 0) return;
  count_launch();
  face_valM_kernel<<<(nFaceNo + 255) / 256, 256, 0, st>>>(nFaceNo, fdof, dof, glob, val, W, valM);
}

// ---------------------------------------------------------------------------
// ADDBCMUL, L/ADDBCMUL.f:53-114 -- a face has O(10^3..10^4) nodes: one block
__global__ void __launch_bounds__(1024) face_dot_kernel(int nFaceNo, int fdof, int dof,
                                                        const int *__restrict__ glob,
                                                        const double *__restrict__ valM,
                                                        const double *__restrict__ X, int ownedLimit,
                                                        int square, double *__restrict__ S,
                                                        const int *done) {
  DONE_GUARD(done);
  __shared__ double smem[32];
  const int m = fdof < dof ? fdof : dof;
  double v = 0.0;
  for (int a = threadIdx.x; a < nFaceNo; a += blockDim.x) {
    const int Ac = glob[a];
    if (Ac >= ownedLimit) continue;
    for (int i = 0; i < m; i++) {
      const double vm = valM[(size_t)a * fdof + i];
      v += square ? vm * vm : vm * X[(size_t)Ac * dof + i];
    }
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int wq = 0; wq < (int)(blockDim.x >> 5); wq++) t += smem[wq];
    *S = t;
  }
}
// ADDBCMUL for a face that lives on one rank (L/ADDBCMUL.f:80-92): S = valM . X, then Y += coef S valM.  It sits
// on the critical path of every Gram-Schmidt column of the rank that holds the face (the other ranks wait for it
// in the all-reduce).  A single CTA is bound by ONE SM's L1 pipe -- the face nodes are ~13 KB apart in X, so every
// lane of every load touches its own line: 23 us for 4225 nodes (ncu, profiles/r02_launches_10M.md) -- so the
// face is spread over kFaceCtas CTAs: partial dots (fixed order inside a CTA), then every CTA of the update
// kernel adds the kFaceCtas partials in the same fixed order and updates its share of the nodes.
static constexpr int kFaceCtas = 32;
__global__ void __launch_bounds__(256) face_dotp_kernel(int nFaceNo, int fdof, int dof,
                                                        const int *__restrict__ glob,
                                                        const double *__restrict__ valM,
                                                        const double *__restrict__ X, int ownedLimit,
                                                        double *__restrict__ partial, const int *done) {
  DONE_GUARD(done);
  __shared__ double smem[8];
  const int m = fdof < dof ? fdof : dof;
  const int chunk = (nFaceNo + gridDim.x - 1) / gridDim.x;
  const int a0 = blockIdx.x * chunk, a1 = min(a0 + chunk, nFaceNo);
  double v = 0.0;
  for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x) {
    const int Ac = __ldg(glob + a);
    if (Ac >= ownedLimit) continue;
    for (int i = 0; i < m; i++) v += __ldg(valM + (size_t)a * fdof + i) * X[(size_t)Ac * dof + i];
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int wq = 0; wq < 8; wq++) t += smem[wq];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) face_axpyp_kernel(int nFaceNo, int fdof, int dof,
                                                         const int *__restrict__ glob,
                                                         const double *__restrict__ valM, double coef,
                                                         const double *__restrict__ partial, double *__restrict__ S,
                                                         double *__restrict__ Y, const int *done) {
  PDL_TRIGGER();
  PDL_WAIT();
  DONE_GUARD(done);
  __shared__ double tot;
  if (threadIdx.x < 32) {
    double t = ((int)threadIdx.x < (int)gridDim.x) ? partial[threadIdx.x] : 0.0;
    t = warp_sum(t);          // the same fixed tree in every CTA
    if (threadIdx.x == 0) {
      tot = t;
      if (blockIdx.x == 0) *S = t;
    }
  }
  __syncthreads();
  const int m = fdof < dof ? fdof : dof;
  const double s = coef * tot;
  const int chunk = (nFaceNo + gridDim.x - 1) / gridDim.x;
  const int a0 = blockIdx.x * chunk, a1 = min(a0 + chunk, nFaceNo);
  for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x) {
    const size_t at = (size_t)__ldg(glob + a) * dof;
    for (int i = 0; i < m; i++) Y[at + i] += __ldg(valM + (size_t)a * fdof + i) * s;
  }
}
// partial: kFaceCtas doubles of scratch.  The two halves are separate launchers so that the dot (which only
// needs X = u(i)) can run on a side stream under the SpMV that produces Y = u(i+1).
void launch_face_dotp(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob, const double *valM,
                      const double *X, int ownedLimit, double *partial, const int *done) {
  count_launch();
  face_dotp_kernel<<<kFaceCtas, 256, 0, st>>>(nFaceNo, fdof, dof, glob, valM, X, ownedLimit, partial, done);
}
void launch_face_axpyp(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob, const double *valM,
                       double coef, const double *partial, double *S, double *Y, const int *done) {
  count_launch();
  launch_pdl(face_axpyp_kernel, dim3(kFaceCtas), dim3(256), 0, st, nFaceNo, fdof, dof, glob, valM, coef, partial, S,
             Y, done);
}

__global__ void face_axpy_kernel(int nFaceNo, int fdof, int dof, const int *__restrict__ glob,
                                 const double *__restrict__ valM, double coef,
                                 const double *__restrict__ S, double *__restrict__ Y,
                                 const int *done) {
  DONE_GUARD(done);
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= nFaceNo) return;
  const int m = fdof < dof ? fdof : dof;
  const double s = coef * (*S);
  for (int i = 0; i < m; i++) Y[(size_t)glob[a] * dof + i] += valM[(size_t)a * fdof + i] * s;
}
void launch_face_dot(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob,
                     const double *valM, const double *X, int ownedLimit, double *S,
                     const int *done) {
  count_launch();
  face_dot_kernel<<<1, 1024, 0, st>>>(nFaceNo, fdof, dof, glob, valM, X, ownedLimit, 0, S, done);
}
void launch_face_norm2(cudaStream_t st, int nFaceNo, int fdof, int nsd, const int *glob,
                       const double *valM, int ownedLimit, double *S) {
  count_launch();
  face_dot_kernel<<<1, 1024, 0, st>>>(nFaceNo, fdof, nsd, glob, valM, nullptr, ownedLimit, 1, S,
                                      nullptr);
}
void launch_face_axpy(cudaStream_t st, int nFaceNo, int fdof, int dof, const int *glob,
                      const double *valM, double coef, const double *S, double *Y,
                      const int *done) {
  if (nFaceNo <= 0) return;
  count_launch();
  face_axpy_kernel<<<(nFaceNo + 255) / 256, 256, 0, st>>>(nFaceNo, fdof, dof, glob, valM, coef, S,
                                                          Y, done);
}

// ---------------------------------------------------------------------------
// One column of the Arnoldi/Givens recurrence, L/GMRES.f:342-366 (scalar part).
// hcol[0..i] = <u_j, u_{i+1}> (already all-reduced).  Column-major h(sD+1, sD).
__global__ void gmres_column_kernel(KrylovCtl *ctl, int i, int sD, const double *__restrict__ hcol,
                                    double *__restrict__ h, double *__restrict__ c,
                                    double *__restrict__ s, double *__restrict__ err,
                                    double *__restrict__ coef, volatile int *pubFlag,
                                    volatile int *pubProgress, int seq) {
  if (threadIdx.x != 0) return;
  if (!ctl->done) {
  double *hc = h + (size_t)(i - 1) * (sD + 1);  // h(:,i), 0-based rows
  double hh = hcol[i];
  for (int j = 0; j < i; j++) {
    const double v = hcol[j];
    hc[j] = v;
    coef[j] = v;
    hh = hh - v * v;
  }
  hh = sqrt(fabs(hh));
  ctl->inv = 1.0 / hh;
  hc[i] = hh;
  for (int j = 0; j < i - 1; j++) {
    const double tmp = c[j] * hc[j] + s[j] * hc[j + 1];
    hc[j + 1] = -s[j] * hc[j] + c[j] * hc[j + 1];
    hc[j] = tmp;
  }
  const double tmp = sqrt(hc[i - 1] * hc[i - 1] + hc[i] * hc[i]);
  c[i - 1] = hc[i - 1] / tmp;
  s[i - 1] = hc[i] / tmp;
  hc[i - 1] = tmp;
  hc[i] = 0.0;
  err[i] = -s[i - 1] * err[i - 1];
  err[i - 1] = c[i - 1] * err[i - 1];
  ctl->ilast = i;
  ctl->itr += 1;
  if (fabs(err[i]) < ctl->eps) {
    ctl->suc = 1;
    ctl->done = 1;
  }
  }
  // publish the stop flag AS OF THIS COLUMN to the host (mapped pinned memory), see solver_int.h
  if (pubFlag) *pubFlag = (seq << 1) | (ctl->done ? 1 : 0);
}
void launch_gmres_column(cudaStream_t st, KrylovCtl *ctl, int i, int sD, const double *hcol,
                         double *h, double *c, double *s, double *err, double *coef,
                         volatile int *pubFlag, volatile int *pubProgress, int seq) {
  count_launch();
  gmres_column_kernel<<<1, 32, 0, st>>>(ctl, i, sD, hcol, h, c, s, err, coef, pubFlag, pubProgress,
                                        seq);
}

// back substitution, L/GMRES.f:370-376; fNorm = |err(i+1)| (:382)
__global__ void gmres_backsub_kernel(KrylovCtl *ctl, int sD, const double *__restrict__ h,
                                     const double *__restrict__ err, double *__restrict__ y) {
  if (threadIdx.x != 0) return;
  const int i = ctl->ilast;
  for (int j = 0; j < i; j++) y[j] = err[j];
  for (int j = i - 1; j >= 0; j--) {
    for (int k = j + 1; k < i; k++) y[j] = y[j] - h[(size_t)k * (sD + 1) + j] * y[k];
    y[j] = y[j] / h[(size_t)j * (sD + 1) + j];
  }
  ctl->fNorm = fabs(err[i]);
}
void launch_gmres_backsub(cudaStream_t st, KrylovCtl *ctl, int sD, const double *h,
                          const double *err, double *y) {
  count_launch();
  gmres_backsub_kernel<<<1, 32, 0, st>>>(ctl, sD, h, err, y);
}

// ---------------------------------------------------------------------------
// DEPART, L/NSSOLVER.f:237-290: split dof x dof blocks into K (nsd x nsd), G (nsd x 1),
// D (1 x nsd), L (1 x 1)
__global__ void depart_kernel(int nnz, int nsd, const double *__restrict__ Val,
                              double *__restrict__ mK, double *__restrict__ mG,
                              double *__restrict__ mD, double *__restrict__ mL) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const int dof = nsd + 1;
  const double *t = Val + (size_t)p * dof * dof;
  for (int i = 0; i < nsd; i++) {
    for (int j = 0; j < nsd; j++) mK[(size_t)p * nsd * nsd + i * nsd + j] = t[i * dof + j];
    mG[(size_t)p * nsd + i] = t[i * dof + nsd];
    mD[(size_t)p * nsd + i] = t[nsd * dof + i];
  }
  mL[p] = t[nsd * dof + nsd];
}
// Gt(:, tpos[p]) = -mG(:, p): tpos = position of the transposed block (L/NSSOLVER.f:292-302)
__global__ void gt_kernel(int nnz, int nsd, const int *__restrict__ tpos,
                          const double *__restrict__ mG, double *__restrict__ Gt) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const int l = tpos[p];
  if (l < 0) return;
  for (int d = 0; d < nsd; d++) Gt[(size_t)l * nsd + d] = -mG[(size_t)p * nsd + d];
}
void launch_depart(cudaStream_t st, int nnz, int nsd, const double *Val, double *mK, double *mG,
                   double *mD, double *mL) {
  count_launch();
  depart_kernel<<<(nnz + 255) / 256, 256, 0, st>>>(nnz, nsd, Val, mK, mG, mD, mL);
}
void launch_gt(cudaStream_t st, int nnz, int nsd, const int *tpos, const double *mG, double *Gt) {
  count_launch();
  gt_kernel<<<(nnz + 255) / 256, 256, 0, st>>>(nnz, nsd, tpos, mG, Gt);
}

}  // namespace svfsi
