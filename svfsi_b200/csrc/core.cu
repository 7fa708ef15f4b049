// core.cu -- context, NCCL plumbing (dlopen'd so a 1-GPU run never needs it), the
// FSILS halo sum (L/INCOMMU.f) and all-reduce (L/BCAST.f) on the library stream,
// layout permutations and event-based profiling.
#include "core.h"

#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace svfsi {

Ctx &ctx() {
  static Ctx c;
  return c;
}

int fail(int code, const std::string &msg) {
  ctx().err = msg;
  fprintf(stderr, "svfsi_b200: error %d: %s\n", code, msg.c_str());
  return code;
}

void count_launch(int n) { ctx().launches += n; }

// ------------------------------------------------------------------ NCCL
namespace {
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

#define NCCL_TRY(expr)                                                                    \
  do {                                                                                    \
    ncclResult_t _r = (expr);                                                             \
    if (_r != ncclSuccess)                                                                \
      return fail(SVFSI_ERR_COMM, std::string(#expr) + ": " +                             \
                                      (g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?")); \
  } while (0)
}  // namespace

int nccl_load() {
  if (g_nccl.h) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return fail(SVFSI_ERR_COMM, std::string("dlopen(libnccl.so.2): ") + dlerror());
#define SYM(field, name)                                                    \
  *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);                         \
  if (!g_nccl.field) return fail(SVFSI_ERR_COMM, std::string("dlsym ") + name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return 0;
}

int nccl_unique_id(void *uid128) {
  if (int rc = nccl_load()) return rc;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(uid128, &id, 128);
  return 0;
}

int nccl_init_rank(const void *uid128, int nranks, int rank) {
  if (int rc = nccl_load()) return rc;
  ncclUniqueId id;
  memcpy(&id, uid128, 128);
  ncclComm_t comm;
  NCCL_TRY(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx().nccl = (void *)comm;
  return 0;
}

void nccl_destroy() {
  if (ctx().nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx().nccl);
  ctx().nccl = nullptr;
}

int nccl_allreduce_sum(double *buf, size_t n, cudaStream_t st) {
  NCCL_TRY(g_nccl.AllReduce(buf, buf, n, ncclFloat64, ncclSum, (ncclComm_t)ctx().nccl, st));
  return 0;
}

int nccl_allgather_i32(const int32_t *send, int32_t n, int32_t *recv, cudaStream_t st) {
  NCCL_TRY(g_nccl.AllGather(send, recv, (size_t)n, ncclInt32, (ncclComm_t)ctx().nccl, st));
  return 0;
}

int nccl_sendrecv(const double *sbuf, double *rbuf, const std::vector<Neighbor> &nbr, int dof,
                  cudaStream_t st) {
  ncclComm_t comm = (ncclComm_t)ctx().nccl;
  NCCL_TRY(g_nccl.GroupStart());
  for (const Neighbor &nb : nbr) {
    NCCL_TRY(g_nccl.Send(sbuf + (size_t)nb.off * dof, (size_t)nb.n * dof, ncclFloat64, nb.iP, comm, st));
    NCCL_TRY(g_nccl.Recv(rbuf + (size_t)nb.off * dof, (size_t)nb.n * dof, ncclFloat64, nb.iP, comm, st));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  return 0;
}

// ------------------------------------------------------------------ peer-memory arena
// One cudaMalloc per rank, exported with cudaIpcGetMemHandle and mapped by every peer
// (cudaIpcOpenMemHandle): NVSwitch gives every GPU load/store access to every peer.  The 64-byte
// handles travel through the host all-gather the caller lent us (MPI in the Fortran shim).
template <typename T>
static int upload_vec(T **d, const std::vector<T> &h) {
  if (*d) cudaFree(*d);
  *d = nullptr;
  CUDA_TRY(cudaMalloc((void **)d, sizeof(T) * (h.empty() ? 1 : h.size())));
  if (!h.empty()) CUDA_TRY(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return 0;
}

void p2p_teardown() {
  Ctx &c = ctx();
  P2P &p = c.p2p;
  if (p.on || p.arena) {
    cudaDeviceSynchronize();
    for (int r = 0; r < (int)p.peer.size(); r++)
      if (r != c.rank && p.peer[r]) cudaIpcCloseMemHandle(p.peer[r]);
    if (p.arena) cudaFree(p.arena);
  }
  if (p.d_peer) cudaFree(p.d_peer);
  if (p.d_nbrRank) cudaFree(p.d_nbrRank);
  if (p.d_nbrOff) cudaFree(p.d_nbrOff);
  if (p.d_nbrPeerOff) cudaFree(p.d_nbrPeerOff);
  if (p.d_nbrN) cudaFree(p.d_nbrN);
  if (p.d_slotNbr) cudaFree(p.d_slotNbr);
  if (p.d_counter) cudaFree(p.d_counter);
  if (p.d_sendPtr) cudaFree(p.d_sendPtr);
  if (p.d_sendRank) cudaFree(p.d_sendRank);
  if (p.d_sendOff) cudaFree(p.d_sendOff);
  p = P2P();
}

int p2p_setup() {
  Ctx &c = ctx();
  P2P &p = c.p2p;
  if (c.nranks == 1) return 0;
  const char *env = getenv("SVFSI_COMM");
  int want = !(env && strcmp(env, "nccl") == 0);
  // every rank must take the same decision: all-gather the local capability
  int can = want;
  if (can) {
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < c.nranks) can = 0;  // single-node, one visible device per rank expected
  }
  std::vector<int32_t> all(c.nranks);
  int32_t mine = can;
  if (int rc = host_allgather_i32(&mine, 1, all.data())) return rc;
  for (int v : all)
    if (!v) return 0;  // stay on NCCL

  // layout (identical on all ranks): halo capacity = max over ranks
  int32_t myShared = c.nShared;
  if (int rc = host_allgather_i32(&myShared, 1, all.data())) return rc;
  int maxShared = 0;
  for (int v : all) maxShared = v > maxShared ? v : maxShared;
  p.haloCap = ((maxShared * 4 + 31) / 32) * 32;
  p.offMail = 4096;
  p.offMailLL = p.offMail + sizeof(double) * 2 * (size_t)c.nranks * kArMax;
  p.offHalo = p.offMailLL + 16 * 2 * (size_t)c.nranks * kArMax;
  p.bytes = p.offHalo + sizeof(double) * 2 * (size_t)(p.haloCap > 0 ? p.haloCap : 32);
  CUDA_TRY(cudaMalloc((void **)&p.arena, p.bytes));
  CUDA_TRY(cudaMemset(p.arena, 0, p.bytes));
  CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t hnd;
  CUDA_TRY(cudaIpcGetMemHandle(&hnd, p.arena));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  std::vector<int32_t> hs((size_t)16 * c.nranks);
  if (int rc = host_allgather_i32((const int32_t *)&hnd, 16, hs.data())) return rc;
  p.peer.assign(c.nranks, nullptr);
  int ok = 1;
  for (int r = 0; r < c.nranks; r++) {
    if (r == c.rank) { p.peer[r] = p.arena; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, hs.data() + (size_t)16 * r, 64);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
    p.peer[r] = (char *)ptr;
  }
  mine = ok;
  if (int rc = host_allgather_i32(&mine, 1, all.data())) return rc;
  for (int v : all)
    if (!v) { p2p_teardown(); return 0; }

  // where does MY slab start inside each neighbour's receive buffer?  all-gather the offset tables
  std::vector<int32_t> offRow(c.nranks, -1), offAll((size_t)c.nranks * c.nranks);
  for (const Neighbor &nb : c.nbr) offRow[nb.iP] = nb.off;
  if (int rc = host_allgather_i32(offRow.data(), c.nranks, offAll.data())) return rc;
  std::vector<int> nbrRank, nbrOff, nbrPeerOff, nbrN, slotNbr(c.nShared);
  for (size_t i = 0; i < c.nbr.size(); i++) {
    const Neighbor &nb = c.nbr[i];
    nbrRank.push_back(nb.iP);
    nbrOff.push_back(nb.off);
    nbrN.push_back(nb.n);
    const int po = offAll[(size_t)nb.iP * c.nranks + c.rank];
    if (po < 0) return fail(SVFSI_ERR_COMM, "asymmetric halo schedule");
    nbrPeerOff.push_back(po);
    for (int s = 0; s < nb.n; s++) slotNbr[nb.off + s] = (int)i;
  }
  if (int rc = upload_vec(&p.d_peer, p.peer)) return rc;
  if (int rc = upload_vec(&p.d_nbrRank, nbrRank)) return rc;
  if (int rc = upload_vec(&p.d_nbrOff, nbrOff)) return rc;
  if (int rc = upload_vec(&p.d_nbrPeerOff, nbrPeerOff)) return rc;
  if (int rc = upload_vec(&p.d_nbrN, nbrN)) return rc;
  if (int rc = upload_vec(&p.d_slotNbr, slotNbr)) return rc;
  {
    // destinations of every boundary row, for the SpMV kernels that send from inside the kernel
    p.nBnd = c.shnNo + (c.nNo - c.mynNo);
    std::vector<std::vector<std::pair<int, int>>> dst((size_t)p.nBnd);
    bool ok2 = true;
    for (size_t i = 0; i < c.nbr.size() && ok2; i++) {
      const Neighbor &nb = c.nbr[i];
      for (int t = 0; t < nb.n; t++) {
        const int row = nb.ptr[t];
        int b;
        if (row < c.shnNo) b = row;
        else if (row >= c.mynNo && row < c.nNo) b = c.shnNo + (row - c.mynNo);
        else { ok2 = false; break; }
        dst[b].push_back({nb.iP, nbrPeerOff[i] + t});
      }
    }
    const char *ef = getenv("SVFSI_SPMV_FUSE");
    p.fuse = ok2 && !(ef && atoi(ef) == 0);
    if (p.fuse) {
      std::vector<int> sp((size_t)p.nBnd + 1, 0), sr, so;
      for (int b = 0; b < p.nBnd; b++) {
        for (auto &d : dst[b]) { sr.push_back(d.first); so.push_back(d.second); }
        sp[b + 1] = (int)sr.size();
      }
      if (int rc = upload_vec(&p.d_sendPtr, sp)) return rc;
      if (int rc = upload_vec(&p.d_sendRank, sr)) return rc;
      if (int rc = upload_vec(&p.d_sendOff, so)) return rc;
    }
  }
  CUDA_TRY(cudaMalloc((void **)&p.d_counter, sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(p.d_counter, 0, sizeof(unsigned int)));
  p.arSeq = 0;
  p.haloSeq = 0;
  p.on = true;
  // nobody may start writing into a peer before that peer has zeroed its arena: one barrier
  if (int rc = host_allgather_i32(&mine, 1, all.data())) return rc;
  return 0;
}

static P2PDev p2p_dev() {
  Ctx &c = ctx();
  P2PDev pd;
  pd.peer = c.p2p.d_peer;
  pd.offMail = c.p2p.offMail;
  pd.offHalo = c.p2p.offHalo;
  pd.offMailLL = c.p2p.offMailLL;
  pd.haloCap = c.p2p.haloCap;
  pd.rank = c.rank;
  pd.nranks = c.nranks;
  pd.errDev = c.d_flag + 2;
  pd.errHost = c.h_status_dev;
  pd.timeoutNs = (long long)(c.commTimeoutS * 1e9);
  return pd;
}

// ------------------------------------------------------------------ sticky error words
int status_words_init() {
  Ctx &c = ctx();
  if (!c.d_flag) {
    CUDA_TRY(cudaMalloc(&c.d_flag, sizeof(int) * 16));
    CUDA_TRY(cudaMemset(c.d_flag, 0, sizeof(int) * 16));
  }
  if (!c.h_status) {
    void *p = nullptr, *d = nullptr;
    CUDA_TRY(cudaHostAlloc(&p, sizeof(int) * 16, cudaHostAllocMapped));
    memset(p, 0, sizeof(int) * 16);
    CUDA_TRY(cudaHostGetDevicePointer(&d, p, 0));
    c.h_status = (volatile int *)p;
    c.h_status_dev = (int *)d;
  }
  c.jacSeen = 0;
  const char *e = getenv("SVFSI_COMM_TIMEOUT_S");
  if (e && atof(e) > 0.0) c.commTimeoutS = atof(e);
  return 0;
}
void status_words_free() {
  Ctx &c = ctx();
  if (c.d_flag) cudaFree(c.d_flag);
  c.d_flag = nullptr;
  if (c.h_status) cudaFreeHost((void *)c.h_status);
  c.h_status = nullptr;
  c.h_status_dev = nullptr;
}
int comm_check() {
  Ctx &c = ctx();
  if (c.h_status && c.h_status[0])
    return fail(SVFSI_ERR_COMM, "a peer-flag wait timed out inside a halo / all-reduce kernel (a rank is "
                                "dead or more than the comm time-out late); results since then are invalid");
  return 0;
}
void jac_publish() {
  Ctx &c = ctx();
  if (c.d_flag && c.h_status)
    cudaMemcpyAsync((void *)(c.h_status + 1), c.d_flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream);
}
int jac_check() {
  Ctx &c = ctx();
  if (!c.h_status) return 0;
  const int bad = c.h_status[1];
  if (bad != c.jacSeen) {
    c.jacSeen = bad;
    return fail(SVFSI_ERR_JAC, "Jac < 0 @ element (ISZERO(Jac), S/FLUID.f:115)");
  }
  return 0;
}

// ------------------------------------------------------------------ collectives
static int halo_send(const double *R, int dof, const int *done) {
  Ctx &c = ctx();
  if (c.p2p.on) {
    c.p2p.haloSeq++;
    launch_halo_send(c.stream, p2p_dev(), dof, c.nShared, (int)c.nbr.size(), c.d_packIdx,
                     c.p2p.d_slotNbr, c.p2p.d_nbrRank, c.p2p.d_nbrOff, c.p2p.d_nbrPeerOff, R,
                     c.p2p.haloSeq, c.p2p.d_counter);
    return 0;
  }
  launch_pack(c.stream, dof, c.nShared, c.d_packIdx, R, c.d_sbuf, done);
  return nccl_sendrecv(c.d_sbuf, c.d_rbuf, c.nbr, dof, c.stream);
}
static int halo_recv(double *R, int dof, const int *done) {
  Ctx &c = ctx();
  if (c.p2p.on) {
    launch_halo_recv_add(c.stream, p2p_dev(), dof, (int)c.nbr.size(), c.p2p.d_nbrRank, c.nUniq,
                         c.d_uniqNode, c.d_uniqPtr, c.d_uniqSlot, R, c.p2p.haloSeq);
    return 0;
  }
  launch_unpack_add(c.stream, dof, c.nUniq, c.d_uniqNode, c.d_uniqPtr, c.d_uniqSlot, c.d_rbuf, R,
                    done);
  return 0;
}

int halo_sum(double *R, int dof, const int *done) {
  Ctx &c = ctx();
  if (c.nranks == 1 || c.nbr.empty()) return 0;
  ProfScope ps(PROF_HALO);
  if (int rc = halo_send(R, dof, done)) return rc;
  return halo_recv(R, dof, done);
}

int allreduce_dev(double *buf, size_t n) {
  Ctx &c = ctx();
  if (c.nranks == 1) return 0;
  ProfScope ps(PROF_ALLREDUCE);
  if (c.p2p.on && n <= (size_t)kArMax) {
    c.p2p.arSeq++;
    launch_p2p_allreduce(c.stream, p2p_dev(), nullptr, 0, (int)n, buf, c.p2p.arSeq);
    return 0;
  }
  return nccl_allreduce_sum(buf, n, c.stream);
}

// out[j] = all-reduce( sum_b partial[j*nblk + b] ), j < k: the tail of every fused multi-dot
int reduce_allreduce(const double *partial, int k, double *out, const int *done) {
  Ctx &c = ctx();
  if (c.nranks > 1 && c.p2p.on && k <= kArMax) {
    ProfScope ps(PROF_ALLREDUCE);
    c.p2p.arSeq++;
    launch_p2p_allreduce(c.stream, p2p_dev(), partial, multidot_nblk(), k, out, c.p2p.arSeq);
    return 0;
  }
  {
    ProfScope ps(PROF_DOT);
    launch_reduce_partials(c.stream, partial, k, out, done);
  }
  if (c.nranks == 1) return 0;
  ProfScope ps(PROF_ALLREDUCE);
  return nccl_allreduce_sum(out, (size_t)k, c.stream);
}

// the same followed by the GMRES column step; fused into one kernel on the single-rank and the
// peer-memory paths, separate kernels on the NCCL fallback
int reduce_allreduce_column(const double *partial, int k, double *out, const ColArgs &col,
                            const int *done) {
  Ctx &c = ctx();
  if (k <= kArMax) {
    if (c.nranks > 1 && c.p2p.on) {
      ProfScope ps(PROF_ALLREDUCE);
      c.p2p.arSeq++;
      launch_p2p_allreduce(c.stream, p2p_dev(), partial, multidot_nblk(), k, out, c.p2p.arSeq, &col);
      return 0;
    }
    if (c.nranks == 1) {
      ProfScope ps(PROF_SMALL);
      launch_reduce_column(c.stream, partial, multidot_nblk(), k, out, col);
      return 0;
    }
  }
  if (int rc = reduce_allreduce(partial, k, out, done)) return rc;
  ProfScope ps(PROF_SMALL);
  launch_gmres_column(c.stream, col.ctl, col.i, col.sD, out, col.h, col.c, col.s, col.err, col.coef,
                      col.pubFlag, col.pubProgress, col.seq);
  return 0;
}

int host_allgather_i32(const int32_t *send, int32_t n, int32_t *recv) {
  Ctx &c = ctx();
  if (c.nranks == 1) {
    memcpy(recv, send, sizeof(int32_t) * (size_t)n);
    return 0;
  }
  if (c.host_allgather) {
    if (c.host_allgather(c.host_allgather_ctx, send, n, recv) != 0)
      return fail(SVFSI_ERR_COMM, "host all-gather callback failed");
    return 0;
  }
  if (!c.nccl) return fail(SVFSI_ERR_COMM, "no host all-gather registered and no NCCL communicator");
  int32_t *ds = nullptr, *dr = nullptr;
  CUDA_TRY(cudaMalloc(&ds, sizeof(int32_t) * (size_t)(n > 0 ? n : 1)));
  CUDA_TRY(cudaMalloc(&dr, sizeof(int32_t) * (size_t)(n > 0 ? n : 1) * c.nranks));
  CUDA_TRY(cudaMemcpyAsync(ds, send, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, c.stream));
  if (int rc = nccl_allgather_i32(ds, n, dr, c.stream)) return rc;
  CUDA_TRY(cudaMemcpyAsync(recv, dr, sizeof(int32_t) * (size_t)n * c.nranks, cudaMemcpyDeviceToHost,
                           c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  cudaFree(ds);
  cudaFree(dr);
  return 0;
}

int row_dof(int kind, int dof) { return (kind == 0 || kind == 2) ? dof : 1; }
int col_dof(int kind, int dof) { return (kind == 0 || kind == 1) ? dof : 1; }

static constexpr int kTraceCols = 4096;
static unsigned long long *trace_slot(bool advance) {
  Ctx &c = ctx();
  static int on = -1;
  if (on < 0) on = getenv("SVFSI_TRACE_FILE") ? 1 : 0;
  if (!on) return nullptr;
  if (!c.d_trace) {
    if (cudaMalloc((void **)&c.d_trace, sizeof(unsigned long long) * 8 * kTraceCols) != cudaSuccess) return nullptr;
    cudaMemset(c.d_trace, 0, sizeof(unsigned long long) * 8 * kTraceCols);
  }
  unsigned long long *p = c.d_trace + (size_t)(c.traceCol % kTraceCols) * 8;
  if (advance) c.traceCol++;
  return p;
}
void trace_dump() {
  Ctx &c = ctx();
  const char *path = getenv("SVFSI_TRACE_FILE");
  if (!c.d_trace || !path) return;
  std::vector<unsigned long long> h((size_t)8 * kTraceCols);
  cudaMemcpy(h.data(), c.d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  const std::string fn = std::string(path) + "." + std::to_string(c.rank);
  if (FILE *fh = fopen(fn.c_str(), "w")) {
    const int n = c.traceCol < kTraceCols ? c.traceCol : kTraceCols;
    for (int i = 0; i < n; i++) {
      for (int j = 0; j < 8; j++) fprintf(fh, "%llu ", h[(size_t)i * 8 + j]);
      fprintf(fh, "\n");
    }
    fclose(fh);
  }
  cudaFree(c.d_trace);
  c.d_trace = nullptr;
  c.traceCol = 0;
}

int multidot_column(const double *U, size_t stride, double *w, size_t nOwned, int k, double *out,
                    const ColArgs *col, const int *done, bool recvPending) {
  Ctx &c = ctx();
  const bool fusedOk = k <= kArMax && (c.nranks == 1 || c.p2p.on) && (!recvPending || c.uniqOrdered);
  static int useFused = -1;
  if (useFused < 0) {
    const char *e = getenv("SVFSI_DOT_FUSED");
    useFused = e ? (atoi(e) != 0) : 1;
  }
  if (fusedOk && useFused) {
    if (!c.d_ticket) {
      CUDA_TRY(cudaMalloc((void **)&c.d_ticket, sizeof(unsigned int)));
      CUDA_TRY(cudaMemsetAsync(c.d_ticket, 0, sizeof(unsigned int), c.stream));
    }
    HaloRecv hr;
    memset(&hr, 0, sizeof(hr));
    DotTail tail;
    memset(&tail, 0, sizeof(tail));
    if (recvPending) {
      hr.on = 1;
      hr.nNbr = (int)c.nbr.size(); hr.dof = c.pendRecvDof; hr.seq = c.p2p.haloSeq;
      hr.shnNo = c.shnNo; hr.mynNo = c.mynNo; hr.nUniq = c.nUniq;
      hr.nbrRank = c.p2p.d_nbrRank; hr.uniqNode = c.d_uniqNode; hr.uniqPtr = c.d_uniqPtr;
      hr.uniqSlot = c.d_uniqSlot;
    }
    tail.nranks = c.nranks;
    if (c.nranks > 1) {
      tail.pd = p2p_dev();
      tail.arSeq = ++c.p2p.arSeq;
    }
    tail.out = out;
    tail.counter = c.d_ticket;
    if (col) tail.col = *col;
    tail.trace = col ? trace_slot(true) : nullptr;   // Gram-Schmidt columns only
    ProfScope ps(PROF_DOT);
    launch_multidot_fused(c.stream, U, stride, w, nOwned, k, c.d_partial, done, hr, tail);
    return 0;
  }
  if (recvPending) {
    ProfScope ps(PROF_HALO);
    if (int rc = halo_recv(w, c.pendRecvDof, done)) return rc;
  }
  {
    ProfScope ps(PROF_DOT);
    launch_multidot(c.stream, U, stride, w, nOwned, k, c.d_partial, done);
  }
  if (col) return reduce_allreduce_column(c.d_partial, k, out, *col, done);
  return reduce_allreduce(c.d_partial, k, out, done);
}

int sparmul(int kind, int dof, const double *K, const double *U, double *KU, const int *done,
            bool *deferRecv) {
  Ctx &c = ctx();
  if (kind == 3) dof = 1;
  const int rd = row_dof(kind, dof);
  if (c.prof) {  // nnz*(8 BR BC + 4) + nNo*(8 + 8 BR + 8 BC), this rank
    const int cd = col_dof(kind, dof);
    c.profSpmvBytes += (double)c.nnz * (8.0 * rd * cd + 4.0) + (double)c.nNo * (8.0 + 8.0 * rd + 8.0 * cd);
    c.profSpmvOps += 1;
  }
  if (c.nranks > 1 && !c.nbr.empty() && c.p2p.on && c.p2p.fuse) {
    // ONE kernel: boundary rows in the first CTAs, their results stored straight into the
    // neighbours' receive buffers, flags raised by the last boundary CTA, interior rows behind
    {
      const bool carriesScale = (c.valScaleW && K == c.d_Val && kind == 0 && dof == 4);
      if (carriesScale && c.prof) {   // timed under "precond": the SpMV roofline is about the plain kernel
        c.profSpmvBytes -= (double)c.nnz * (8.0 * 16 + 4.0) + (double)c.nNo * (8.0 + 64.0);
        c.profSpmvOps -= 1;
      }
      ProfScope ps(carriesScale ? PROF_PRECOND : PROF_SPMV);
      c.p2p.haloSeq++;
      SpmvFuse f;
      f.shnNo = c.shnNo; f.mynNo = c.mynNo; f.nNo = c.nNo;
      f.nBnd = c.p2p.nBnd; f.bndCtas = 0;
      f.sendPtr = c.p2p.d_sendPtr; f.sendRank = c.p2p.d_sendRank; f.sendOff = c.p2p.d_sendOff;
      f.nbrRank = c.p2p.d_nbrRank; f.nNbr = (int)c.nbr.size();
      f.pd = p2p_dev(); f.seq = c.p2p.haloSeq; f.counter = c.p2p.d_counter;
      f.trace = deferRecv ? trace_slot(false) : nullptr;   // the SpMV of the column the next column kernel closes
      const double *scaleW = (c.valScaleW && K == c.d_Val && kind == 0 && dof == 4) ? c.valScaleW : nullptr;
      launch_spmv_fused(c.stream, kind, dof, f, c.d_rowPtr, c.d_col, K, U, KU, done, scaleW);
      if (scaleW) c.valScaleW = nullptr;
    }
    if (deferRecv && c.uniqOrdered) {   // the next multidot_column on KU receives
      *deferRecv = true;
      c.pendRecvDof = rd;
      return 0;
    }
    ProfScope ps(PROF_HALO);
    return halo_recv(KU, rd, done);
  }
  if (c.nranks > 1 && !c.nbr.empty()) {
    // rows shared with other ranks first (the two contiguous slabs at the ends of the reordered
    // numbering, L/LHS.f:134-165), push them to the neighbours, then the interior rows while the
    // halo is in flight, then add what arrived (L/SPARMUL.f:130 + L/INCOMMU.f:91-96)
    {
      ProfScope ps(PROF_SPMV);
      launch_spmv2(c.stream, kind, dof, 0, c.shnNo, c.mynNo, c.nNo, c.d_rowPtr, c.d_col, K, U, KU,
                   done);
    }
    {
      ProfScope ps(PROF_HALO);
      if (int rc = halo_send(KU, rd, done)) return rc;
    }
    {
      ProfScope ps(PROF_SPMV);
      launch_spmv(c.stream, kind, dof, c.shnNo, c.mynNo, c.d_rowPtr, c.d_col, K, U, KU, done);
    }
    ProfScope ps(PROF_HALO);
    return halo_recv(KU, rd, done);
  }
  if (c.valScaleW && K == c.d_Val && kind == 0 && dof == 4) {
    // the Jacobi scaling PRECONDDIAG left pending rides on this first product (same arithmetic, one read of
    // Val less); `done` cannot be set before the first column of a solve.  Timed under "precond": the SpMV
    // roofline is about the plain kernel.
    if (c.prof) {
      c.profSpmvBytes -= (double)c.nnz * (8.0 * 16 + 4.0) + (double)c.nNo * (8.0 + 64.0);
      c.profSpmvOps -= 1;
    }
    ProfScope ps(PROF_PRECOND);
    launch_spmv_vv4_scale(c.stream, c.nNo, c.d_rowPtr, c.d_col, c.d_Val, c.valScaleW, U, KU);
    c.valScaleW = nullptr;
    return 0;
  }
  {
    ProfScope ps(PROF_SPMV);
    launch_spmv(c.stream, kind, dof, 0, c.nNo, c.d_rowPtr, c.d_col, K, U, KU, done);
  }
  return 0;
}

// apply a scaling of Val that is still pending (any consumer of Val other than the fused first product)
int flush_val_scale() {
  Ctx &c = ctx();
  if (!c.valScaleW) return 0;
  ProfScope ps(PROF_PRECOND);
  launch_scale_val(c.stream, c.nnz, 4, c.d_rowOf, c.d_col, c.valScaleW, c.d_Val);
  c.valScaleW = nullptr;
  return 0;
}

// ------------------------------------------------------------------ memory
int ensure_ws(size_t bytes) {
  Ctx &c = ctx();
  if (c.wsBytes >= bytes) return 0;
  if (c.d_ws) cudaFree(c.d_ws);
  c.d_ws = nullptr;
  c.wsBytes = 0;
  CUDA_TRY(cudaMalloc(&c.d_ws, bytes));
  c.wsBytes = bytes;
  return 0;
}

int ensure_stage(size_t bytes) {
  Ctx &c = ctx();
  if (c.stageBytes >= bytes) return 0;
  if (c.d_stage) cudaFree(c.d_stage);
  c.d_stage = nullptr;
  c.stageBytes = 0;
  CUDA_TRY(cudaMalloc(&c.d_stage, bytes));
  c.stageBytes = bytes;
  return 0;
}

static constexpr size_t kSmallDoubles = 1 << 17;  // 1 MB of scalars: Hessenberg up to sD ~ 340

int ensure_small() {
  Ctx &c = ctx();
  if (c.d_small) return 0;
  CUDA_TRY(cudaMalloc(&c.d_small, kSmallDoubles * sizeof(double)));
  CUDA_TRY(cudaMemset(c.d_small, 0, kSmallDoubles * sizeof(double)));
  CUDA_TRY(cudaMallocHost(&c.h_small, kSmallDoubles * sizeof(double)));
  c.partialDoubles = (size_t)multidot_nblk() * 512;
  CUDA_TRY(cudaMalloc(&c.d_partial, c.partialDoubles * sizeof(double)));
  return 0;
}

// ------------------------------------------------------------------ layouts
int upload_nodal(const double *host, int m, double *dev) {
  Ctx &c = ctx();
  const size_t bytes = sizeof(double) * (size_t)c.nNo * m;
  if (c.nranks == 1) {  // map is the identity (L/LHS.f:89-111)
    CUDA_TRY(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c.stream));
    return 0;
  }
  if (int rc = ensure_stage(bytes)) return rc;
  CUDA_TRY(cudaMemcpyAsync(c.d_stage, host, bytes, cudaMemcpyHostToDevice, c.stream));
  launch_perm_scatter(c.stream, c.nNo, m, c.d_perm, c.d_stage, dev);
  return 0;
}

int download_nodal(const double *dev, int m, double *host) {
  Ctx &c = ctx();
  const size_t bytes = sizeof(double) * (size_t)c.nNo * m;
  if (c.nranks == 1) {
    CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    return 0;
  }
  if (int rc = ensure_stage(bytes)) return rc;
  launch_perm_gather(c.stream, c.nNo, m, c.d_perm, dev, c.d_stage);
  CUDA_TRY(cudaMemcpyAsync(host, c.d_stage, bytes, cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  return 0;
}

int upload_val(const double *host, int dd, double *dev) {
  Ctx &c = ctx();
  const size_t bytes = sizeof(double) * (size_t)c.nnz * dd;
  if (c.nranks == 1) {
    CUDA_TRY(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c.stream));
    return 0;
  }
  if (int rc = ensure_stage(bytes)) return rc;
  CUDA_TRY(cudaMemcpyAsync(c.d_stage, host, bytes, cudaMemcpyHostToDevice, c.stream));
  launch_perm_scatter(c.stream, c.nnz, dd, c.d_vperm, c.d_stage, dev);
  return 0;
}

int download_val(const double *dev, int dd, double *host) {
  Ctx &c = ctx();
  const size_t bytes = sizeof(double) * (size_t)c.nnz * dd;
  if (c.nranks == 1) {
    CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    return 0;
  }
  if (int rc = ensure_stage(bytes)) return rc;
  launch_perm_gather(c.stream, c.nnz, dd, c.d_vperm, dev, c.d_stage);
  CUDA_TRY(cudaMemcpyAsync(host, c.d_stage, bytes, cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  return 0;
}

// ------------------------------------------------------------------ profiling
ProfScope::ProfScope(int s) : slot(s), on(ctx().prof), idx(0) {
  if (!on) return;
  Ctx &c = ctx();
  if (c.evUsed == c.evPool.size()) {
    EventPair ep;
    cudaEventCreate(&ep.a);
    cudaEventCreate(&ep.b);
    ep.slot = slot;
    c.evPool.push_back(ep);
  }
  idx = c.evUsed++;
  c.evPool[idx].slot = slot;
  cudaEventRecord(c.evPool[idx].a, c.stream);
}
ProfScope::~ProfScope() {
  if (!on) return;
  Ctx &c = ctx();
  cudaEventRecord(c.evPool[idx].b, c.stream);
}

void prof_collect() {
  Ctx &c = ctx();
  if (c.evUsed == 0) return;
  cudaStreamSynchronize(c.stream);
  for (size_t i = 0; i < c.evUsed; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c.evPool[i].a, c.evPool[i].b) == cudaSuccess) {
      c.profMs[c.evPool[i].slot] += ms;
      c.profN[c.evPool[i].slot] += 1;
    }
  }
  c.evUsed = 0;
}

}  // namespace svfsi
