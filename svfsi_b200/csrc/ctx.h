// ctx.h -- process-wide device context of the svFSI hot path on one B200.
//
// Everything on the device lives in the FSILS (reordered) node numbering of
// L/LHS.f:134-211: rows [0,shnNo) are shared with lower ranks, [shnNo,mynNo)
// interior, [mynNo,nNo) shared with higher ranks.  The block-CSR rows are stored
// in that order with each row's blocks kept in svFSI's original (ascending
// original column id) order, so the SpMV sums in the reference's order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/svfsi_b200.h"

namespace svfsi {

// timer slots of gpu_prof_get_
enum ProfSlot {
  PROF_ASM = 0,      // element loop kernels (fluid / heat)
  PROF_SPMV = 1,     // block-CSR SpMV kernels (all shapes)
  PROF_HALO = 2,     // pack + NCCL + unpack
  PROF_DOT = 3,      // fused multi-dot / dot / norm kernels
  PROF_AXPY = 4,     // fused multi-axpy / scale kernels
  PROF_PRECOND = 5,  // Jacobi scaling
  PROF_SMALL = 6,    // Hessenberg / Givens / scalar kernels
  PROF_ALLREDUCE = 7,
  PROF_SOLVE = 8,    // whole FSILS_SOLVE on the device
  PROF_NSLOTS = SVFSI_NTIMERS
};

struct Face {
  bool created = false;
  bool coupled = false, shared = false, inc = true;
  int nNo = 0, dof = 0, bGrp = SVFSI_BC_TYPE_DIR;
  double nS = 0.0, res = 0.0;
  int *d_glob = nullptr;     // [nNo] 0-based reordered node ids
  double *d_val = nullptr;   // [nNo][dof]
  double *d_valM = nullptr;  // [nNo][dof]
  std::vector<int> glob;     // host copy, 0-based reordered
};

struct Neighbor {
  int iP = 0;  // 0-based peer rank
  int n = 0;
  int off = 0;            // offset (in nodes) of this neighbour's slab in the pack buffers
  std::vector<int> ptr;   // 0-based reordered node ids (host copy)
};

// Peer-memory communication arena (one cudaMalloc per rank, IPC-mapped by every peer).
//   [flags: 4 x 64 ints][all-reduce mailbox: 2 slots x nranks x kArMax doubles]
//   [flag-in-data mailbox: 2 slots x nranks x kArMax x 16 bytes][halo receive buffer: 2 slots x haloCap doubles]
// Identical layout on every rank so a peer address is base(peer) + the local offset.
struct P2P {
  bool on = false;
  char *arena = nullptr;
  size_t bytes = 0;
  std::vector<char *> peer;      // [nranks] mapped base of every rank's arena (own = arena)
  size_t offMail = 0, offHalo = 0, offMailLL = 0;
  int haloCap = 0;               // doubles per halo slot
  std::vector<int> peerOff;      // [nbr] offset (nodes) of MY slab inside neighbour i's receive buffer
  int arSeq = 0, haloSeq = 0;    // advance identically on every rank
  char **d_peer = nullptr;       // device copy of peer[]
  int *d_nbrRank = nullptr, *d_nbrOff = nullptr, *d_nbrPeerOff = nullptr, *d_nbrN = nullptr;
  int *d_slotNbr = nullptr;      // [nShared] neighbour index of every pack slot
  // fused SpMV + send: destinations of every boundary row (CSR over boundary-row index)
  bool fuse = false;
  int nBnd = 0;
  int *d_sendPtr = nullptr, *d_sendRank = nullptr, *d_sendOff = nullptr;
  unsigned int *d_counter = nullptr;
};
static constexpr int kArMax = 512;

struct EventPair {
  cudaEvent_t a, b;
  int slot;
};

struct Ctx {
  bool inited = false;
  int device = 0, rank = 0, nranks = 1;
  cudaStream_t stream = nullptr;
  // side stream + events: small kernels that only need the INPUT of the running SpMV (face dots of ADDBCMUL)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  void *nccl = nullptr;  // ncclComm_t
  svfsi_allgather_i32_fn host_allgather = nullptr;
  void *host_allgather_ctx = nullptr;
  std::string err;
  int64_t launches = 0;

  // ---- lhs (FSILS_lhsType) ----
  bool lhs = false;
  int lhsGen = 0;            // bumped by every FSILS_LHS_CREATE (invalidates cached maps)
  int gnNo = 0, nNo = 0, nnz = 0, nFaces = 0, mynNo = 0, shnNo = 0;
  std::vector<int> map;       // [nNo] svFSI local id (0-based) -> reordered id (0-based)
  std::vector<int> rowPtrDev; // host copy of device rowPtr (0-based, nNo+1)
  int maxRowLen = 0;          // longest block row (sizes the row-owner assembly's shared memory)
  std::vector<Neighbor> nbr;
  std::vector<Face> face;
  int *d_perm = nullptr;     // [nNo] = map
  int *d_rowPtr = nullptr;   // [nNo+1] device-layout block offsets
  int *d_col = nullptr;      // [nnz] reordered column ids, device layout
  int *d_diag = nullptr;     // [nNo] block index of the diagonal
  int *d_vperm = nullptr;    // [nnz] svFSI block position -> device block position
  int *d_rowOf = nullptr;    // [nnz] row of each block (device layout)
  // halo
  int nShared = 0;           // total entries in pack buffers (sum of nbr.n)
  int *d_packIdx = nullptr;  // [nShared] node id per pack slot
  // SVFSI_TRACE_FILE=<path>: device time stamps (globaltimer ns) of the column kernels, 8 per column, written
  // to <path>.<rank> at gpu_finalize_ (tools/trace_columns.py reads them)
  unsigned long long *d_trace = nullptr;
  int traceCol = 0;
  bool uniqOrdered = false;  // uniqNode = [0..shnNo) then nodes >= mynNo (what the fused receive assumes)
  int pendRecvDof = 0;       // dof of the vector whose halo receive a sparmul left pending
  unsigned int *d_ticket = nullptr;  // CTA ticket of the fused multi-dot kernel
  int nUniq = 0;             // unique shared nodes
  int *d_uniqNode = nullptr; // [nUniq]
  int *d_uniqPtr = nullptr;  // [nUniq+1] -> d_uniqSlot
  int *d_uniqSlot = nullptr; // pack-slot ids in ascending neighbour order
  double *d_sbuf = nullptr, *d_rbuf = nullptr;  // [nShared*4]
  P2P p2p;

  // ---- mesh ----
  bool mesh = false;
  int nEl = 0;
  int *d_ien = nullptr;      // [nEl][4] 0-based reordered node ids
  int *d_edest = nullptr;    // [nEl][16] device block index of (a,b)
  double *d_x = nullptr;     // [nNo][3]
  int ncolors = 0;
  std::vector<int> colorOff; // [ncolors+1]
  int *d_colorElems = nullptr;  // [nEl] element ids grouped by colour
  // gather variant adjacency
  int *d_blkAdjPtr = nullptr;   // [nnz+1]
  int *d_blkAdj = nullptr;      // [16*nEl] (e<<4 | a<<2 | b)
  int *d_nodeAdjPtr = nullptr;  // [nNo+1]
  int *d_nodeAdj = nullptr;     // [4*nEl] (e<<2 | a)
  int *d_nodeSlots = nullptr;   // [4*nEl] row-local slots of the visit's four blocks (8 bits each)
  double *d_elemP = nullptr;    // [nEl][64] per-element compact records (gather variant)
  int *d_blkOrder = nullptr;    // [nnz padded] block processing order (length-sorted chunks)
  int *d_pairList = nullptr;    // [nPair] blocks (r,c), c >= r, in processing order (pair-owner gather)
  int *d_pairT = nullptr;       // [nPair] position of the transposed block (c,r)
  int nPair = 0;
  int4 *d_blkDesc = nullptr;    // [nnz] (block, list begin, list end, 0) in processing order
  int4 *d_pairDesc = nullptr;   // [nnz] paired order: (r,c) / (c,r) adjacent, then diagonal blocks (row + 1 in .w)
  int nPairDescOff = 0;         // entries of d_pairDesc that belong to off-diagonal pairs

  // ---- system ----
  int dof = 0;                // dof of the resident R/Val
  double *d_R = nullptr;      // [nNo][4]
  double *d_Val = nullptr;    // [nnz][16]
  double *d_Ag = nullptr, *d_Yg = nullptr, *d_Bf = nullptr;  // [nNo][4],[nNo][4],[nNo][3]
  double *d_stage = nullptr;  // staging for H2D/D2H permutes
  size_t stageBytes = 0;
  // device int words: [0] bad-Jacobian count (monotone since gpu_init_, never cleared so that a hit is
  // not lost before the host has seen it), [1] scratch of the setup checks, [2] sticky "peer flag wait
  // timed out"
  // PRECONDDIAG's scaling of Val left to the first FSILS_SPARMULVV of the solve (fused kernel); W of that scaling
  const double *valScaleW = nullptr;
  int *d_flag = nullptr;
  // mapped pinned host words the device (or an async copy) writes and the host reads at its next
  // synchronisation point without another copy: [0] comm time-out, [1] copy of d_flag[0]
  volatile int *h_status = nullptr;
  int *h_status_dev = nullptr;   // device alias of h_status
  int jacSeen = 0;               // value of the bad-Jacobian count already reported
  double commTimeoutS = 120.0;   // bound of every in-kernel peer-flag wait (gpu_set_comm_timeout_)

  // ---- solver workspace (grown on demand) ----
  double *d_ws = nullptr;
  size_t wsBytes = 0;
  double *d_small = nullptr;  // small scalars: dots, Hessenberg, control block
  double *h_small = nullptr;  // pinned mirror
  double *d_partial = nullptr;
  size_t partialDoubles = 0;

  // ---- profiling ----
  bool prof = false;
  std::vector<EventPair> evPool;
  size_t evUsed = 0;
  double profMs[PROF_NSLOTS] = {0};
  int64_t profN[PROF_NSLOTS] = {0};
  double profSpmvBytes = 0.0;   // algorithmic bytes of the SpMVs issued while profiling
  int64_t profSpmvOps = 0;
};

Ctx &ctx();
int fail(int code, const std::string &msg);

#define CUDA_TRY(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      return svfsi::fail(SVFSI_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

}  // namespace svfsi
