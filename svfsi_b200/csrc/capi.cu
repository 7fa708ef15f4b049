// capi.cu -- the extern "C" boundary declared in include/svfsi_b200.h.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <stdexcept>
#include <vector>

#include "core.h"
#include "lhs_plan.h"

using namespace svfsi;

namespace svfsi {
void faces_free_all();   // face.cu
void solver_free_static();   // solver.cu
int solver_bcpre(int nsd, double *sS);
}

namespace {

template <typename T>
int dev_upload(T **dptr, const std::vector<T> &h) {
  if (*dptr) cudaFree(*dptr);
  *dptr = nullptr;
  const size_t bytes = sizeof(T) * std::max<size_t>(h.size(), 1);
  CUDA_TRY(cudaMalloc((void **)dptr, bytes));
  if (!h.empty())
    CUDA_TRY(cudaMemcpy(*dptr, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return 0;
}

template <typename T>
void dev_free(T **p) {
  if (*p) cudaFree(*p);
  *p = nullptr;
}

int need_init() {
  if (!ctx().inited) return fail(SVFSI_ERR_STATE, "gpu_init_ has not been called");
  return 0;
}

int ensure_system(int dof) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "FSILS_LHS_CREATE has not been called");
  // R and Val are sized for the largest block this library handles (nsd + 1 = 4)
  if (dof < 1 || dof > 4) return fail(SVFSI_ERR_ARG, "dof must be 1..4");
  if (!c.d_R) CUDA_TRY(cudaMalloc(&c.d_R, sizeof(double) * (size_t)c.nNo * 4));
  if (!c.d_Val) CUDA_TRY(cudaMalloc(&c.d_Val, sizeof(double) * (size_t)c.nnz * 16));
  c.dof = dof;
  return 0;
}

int ensure_state() {
  Ctx &c = ctx();
  if (!c.d_Ag) CUDA_TRY(cudaMalloc(&c.d_Ag, sizeof(double) * (size_t)c.nNo * 4));
  if (!c.d_Yg) CUDA_TRY(cudaMalloc(&c.d_Yg, sizeof(double) * (size_t)c.nNo * 4));
  if (!c.d_Bf) CUDA_TRY(cudaMalloc(&c.d_Bf, sizeof(double) * (size_t)c.nNo * 3));
  return 0;
}

bool g_haveBf = false;

// greedy element colouring: no two elements of one colour share a node
void color_elements(int nEl, int nNo, const std::vector<int> &ien, std::vector<int> &colorOff,
                    std::vector<int> &colorElems) {
  const int W = 4;  // up to 256 colours
  std::vector<uint64_t> used((size_t)nNo * W, 0);
  std::vector<int> color(nEl);
  int ncol = 0;
  for (int e = 0; e < nEl; e++) {
    uint64_t m[W];
    for (int w = 0; w < W; w++) m[w] = 0;
    for (int a = 0; a < 4; a++) {
      const uint64_t *u = &used[(size_t)ien[(size_t)e * 4 + a] * W];
      for (int w = 0; w < W; w++) m[w] |= u[w];
    }
    int col = -1;
    for (int w = 0; w < W && col < 0; w++)
      if (~m[w]) col = w * 64 + __builtin_ctzll(~m[w]);
    if (col < 0) throw std::runtime_error("element colouring needs more than 256 colours");
    color[e] = col;
    ncol = std::max(ncol, col + 1);
    for (int a = 0; a < 4; a++) used[(size_t)ien[(size_t)e * 4 + a] * W + col / 64] |= 1ull << (col % 64);
  }
  colorOff.assign(ncol + 1, 0);
  for (int e = 0; e < nEl; e++) colorOff[color[e] + 1]++;
  for (int k = 0; k < ncol; k++) colorOff[k + 1] += colorOff[k];
  colorElems.resize(nEl);
  std::vector<int> pos(colorOff.begin(), colorOff.end() - 1);
  for (int e = 0; e < nEl; e++) colorElems[pos[color[e]]++] = e;
}

static PairLists pair_lists(const Ctx &c) {
  PairLists pl;
  pl.list = c.d_pairList; pl.tpos = c.d_pairT; pl.rowOf = c.d_rowOf; pl.n = c.nPair;
  pl.desc = c.d_blkDesc;
  return pl;
}

// the owner-computes (gather) assembly: parts 1 = element records, 2 = tangent gather, 4 = residual gather.
// tune bit 20 (1048576): second generation (asm_gather5.cu: 512-byte records v5, transposed blocks in adjacent
// groups, residual folded into the diagonal groups; bit 21: 256-thread CTAs), otherwise the first-generation
// kernels of asm_kernels.cu selected by the lower bits.
static constexpr int kTuneVisit = 1 << 20;
int fluid_gather_dispatch(int parts, const FluidPar &par, int tune) {
  Ctx &c = ctx();
  const double *bf = g_haveBf ? c.d_Bf : nullptr;
  if (!c.d_elemP) CUDA_TRY(cudaMalloc(&c.d_elemP, sizeof(double) * 80 * (size_t)c.nEl));
  if ((tune & kTuneVisit) && c.d_pairDesc && (double)c.nEl * 64 < 4.0e9) {
    int p2 = (parts & 1) | ((parts & 6) ? 2 : 0);
    // bit 22: length-sorted processing order of the first generation (diagonal blocks flagged) instead of the
    // paired order
    const int4 *desc = (tune & (1 << 22)) ? c.d_blkDesc : c.d_pairDesc;
    launch_fluid_gather5(c.stream, p2, par, 0, c.nEl, 0, c.nnz, c.d_ien, c.d_x, c.d_Ag, c.d_Yg, bf, c.d_elemP,
                         ~0u, desc, c.d_blkAdj, c.d_R, c.d_Val, c.d_flag, (tune >> 21) & 1);
    return 0;
  }
  launch_fluid_gather_parts(c.stream, parts, par, c.nEl, c.nNo, c.nnz, c.d_ien, c.d_x, c.d_Ag, c.d_Yg, bf,
                            c.d_elemP, c.d_blkOrder, c.d_blkAdjPtr, c.d_blkAdj, c.d_nodeAdjPtr, c.d_nodeAdj,
                            c.d_R, c.d_Val, c.d_flag, tune, c.d_rowPtr, c.d_nodeSlots, c.maxRowLen,
                            pair_lists(c));
  return 0;
}

int run_fluid_asm(const FluidPar &par, int variant) {
  Ctx &c = ctx();
  if (!c.mesh) return fail(SVFSI_ERR_STATE, "gpu_mesh_create_ has not been called");
  if (int rc = ensure_system(4)) return rc;
  // LSALLOC: R = 0, Val = 0 (S/LS.f:44-51); the gather variant writes every block and every
  // residual entry exactly once (0 + contributions), so it needs no zero fill
  if (variant != SVFSI_ASM_GATHER) {
    CUDA_TRY(cudaMemsetAsync(c.d_R, 0, sizeof(double) * (size_t)c.nNo * 4, c.stream));
    CUDA_TRY(cudaMemsetAsync(c.d_Val, 0, sizeof(double) * (size_t)c.nnz * 16, c.stream));
  }
  const double *bf = g_haveBf ? c.d_Bf : nullptr;
  {
    ProfScope ps(PROF_ASM);
    if (variant == SVFSI_ASM_ATOMIC) {
      launch_fluid_asm(c.stream, par, c.nEl, 0, nullptr, c.d_ien, c.d_edest, c.d_x, c.d_Ag, c.d_Yg,
                       bf, c.d_R, c.d_Val, 1, c.d_flag);
    } else if (variant == SVFSI_ASM_COLORED) {
      for (int k = 0; k < c.ncolors; k++)
        launch_fluid_asm(c.stream, par, c.colorOff[k + 1] - c.colorOff[k], c.colorOff[k],
                         c.d_colorElems, c.d_ien, c.d_edest, c.d_x, c.d_Ag, c.d_Yg, bf, c.d_R,
                         c.d_Val, 0, c.d_flag);
    } else if (variant == SVFSI_ASM_GATHER) {
      if (int rc = fluid_gather_dispatch(7, par, asm_tune())) return rc;
    } else {
      return fail(SVFSI_ERR_ARG, "unknown assembly variant");
    }
  }
  jac_publish();   // bad-Jacobian count -> mapped host word, read at the next synchronisation point
  return 0;
}

int run_heat_asm(const HeatPar &par, int variant) {
  Ctx &c = ctx();
  if (!c.mesh) return fail(SVFSI_ERR_STATE, "gpu_mesh_create_ has not been called");
  if (int rc = ensure_system(1)) return rc;
  if (variant != SVFSI_ASM_GATHER) {
    CUDA_TRY(cudaMemsetAsync(c.d_R, 0, sizeof(double) * (size_t)c.nNo, c.stream));
    CUDA_TRY(cudaMemsetAsync(c.d_Val, 0, sizeof(double) * (size_t)c.nnz, c.stream));
  }
  {
    ProfScope ps(PROF_ASM);
    if (variant == SVFSI_ASM_ATOMIC) {
      launch_heat_asm(c.stream, par, c.nEl, 0, nullptr, c.d_ien, c.d_edest, c.d_x, c.d_Ag, c.d_Yg,
                      c.d_R, c.d_Val, 1, c.d_flag);
    } else if (variant == SVFSI_ASM_COLORED) {
      for (int k = 0; k < c.ncolors; k++)
        launch_heat_asm(c.stream, par, c.colorOff[k + 1] - c.colorOff[k], c.colorOff[k],
                        c.d_colorElems, c.d_ien, c.d_edest, c.d_x, c.d_Ag, c.d_Yg, c.d_R, c.d_Val, 0,
                        c.d_flag);
    } else if (variant == SVFSI_ASM_GATHER) {
      if (!c.d_elemP) CUDA_TRY(cudaMalloc(&c.d_elemP, sizeof(double) * 80 * (size_t)c.nEl));
      launch_heat_gather(c.stream, par, c.nEl, c.nNo, c.nnz, c.d_ien, c.d_x, c.d_Ag, c.d_Yg,
                         c.d_elemP, c.d_blkAdjPtr, c.d_blkAdj, c.d_nodeAdjPtr, c.d_nodeAdj, c.d_R,
                         c.d_Val, c.d_flag);
    } else {
      return fail(SVFSI_ERR_ARG, "unknown assembly variant");
    }
  }
  jac_publish();
  return 0;
}

// the reference aborts in the element loop ("Jac < 0 @ element", S/FLUID.f:115).  Host-buffer calls
// report it at once; the device-resident calls (gpu_construct_*_dev_) do not synchronise, so the
// count travels to a mapped host word behind the kernels and the NEXT call that synchronises anyway
// (gpu_solve_dev_, gpu_sync_, gpu_get_*_) returns SVFSI_ERR_JAC.  The count is never cleared.
int check_jac() {
  CUDA_TRY(cudaStreamSynchronize(ctx().stream));
  return jac_check();
}

}  // namespace

extern "C" {

int32_t gpu_nccl_unique_id_(void *uid128) { return nccl_unique_id(uid128); }

int32_t gpu_init_(const int32_t *device, const int32_t *rank, const int32_t *nranks,
                  const void *uid128) {
  Ctx &c = ctx();
  if (c.inited) return fail(SVFSI_ERR_STATE, "gpu_init_ called twice");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(SVFSI_ERR_CUDA,
                std::string("no CUDA device: this library has no CPU fallback (") +
                    cudaGetErrorString(e) + ")");
  c.device = *device;
  c.rank = *rank;
  c.nranks = *nranks;
  CUDA_TRY(cudaSetDevice(c.device));
  CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&c.evFork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&c.evJoin, cudaEventDisableTiming));
  if (int rc = status_words_init()) return rc;
  if (c.nranks > 1) {
    if (!uid128) return fail(SVFSI_ERR_ARG, "nranks > 1 needs the NCCL unique id");
    if (int rc = nccl_init_rank(uid128, c.nranks, c.rank)) return rc;
  }
  c.inited = true;
  return 0;
}

int32_t gpu_set_host_allgather_(svfsi_allgather_i32_fn fn, void *user) {
  ctx().host_allgather = fn;
  ctx().host_allgather_ctx = user;
  return 0;
}

int32_t gpu_last_error_(char *buf, const int32_t *len) {
  if (*len <= 0) return 0;
  strncpy(buf, ctx().err.c_str(), (size_t)*len - 1);
  buf[*len - 1] = 0;
  return 0;
}

int32_t gpu_lhs_free_(void) {
  Ctx &c = ctx();
  if (c.stream) cudaStreamSynchronize(c.stream);
  p2p_teardown();
  for (Face &f : c.face) {
    dev_free(&f.d_glob);
    dev_free(&f.d_val);
    dev_free(&f.d_valM);
  }
  c.face.clear();
  c.nbr.clear();
  dev_free(&c.d_perm); dev_free(&c.d_rowPtr); dev_free(&c.d_col); dev_free(&c.d_diag);
  dev_free(&c.d_vperm); dev_free(&c.d_rowOf); dev_free(&c.d_packIdx); dev_free(&c.d_uniqNode);
  dev_free(&c.d_uniqPtr); dev_free(&c.d_uniqSlot); dev_free(&c.d_sbuf); dev_free(&c.d_rbuf);
  dev_free(&c.d_ien); dev_free(&c.d_edest); dev_free(&c.d_x); dev_free(&c.d_colorElems);
  dev_free(&c.d_blkAdjPtr); dev_free(&c.d_blkAdj); dev_free(&c.d_nodeAdjPtr);
  dev_free(&c.d_nodeAdj); dev_free(&c.d_blkOrder); dev_free(&c.d_elemP); dev_free(&c.d_nodeSlots);
  dev_free(&c.d_pairList); dev_free(&c.d_pairT); c.nPair = 0; dev_free(&c.d_blkDesc);
  dev_free(&c.d_pairDesc); c.nPairDescOff = 0;
  dev_free(&c.d_R); dev_free(&c.d_Val); dev_free(&c.d_Ag); dev_free(&c.d_Yg); dev_free(&c.d_Bf);
  gpu_pic_free_();
  faces_free_all();
  c.lhs = false;
  c.mesh = false;
  g_haveBf = false;
  return 0;
}

int32_t gpu_finalize_(void) {
  Ctx &c = ctx();
  if (!c.inited) return 0;
  trace_dump();
  gpu_lhs_free_();
  dev_free(&c.d_ws); c.wsBytes = 0;
  dev_free(&c.d_stage); c.stageBytes = 0;
  dev_free(&c.d_small); dev_free(&c.d_partial); c.partialDoubles = 0;
  dev_free(&c.d_ticket);
  if (c.h_small) cudaFreeHost(c.h_small);
  c.h_small = nullptr;
  solver_free_static();     // W / RCS work vectors, transposed-position map, host mirror
  status_words_free();
  for (EventPair &ep : c.evPool) { cudaEventDestroy(ep.a); cudaEventDestroy(ep.b); }
  c.evPool.clear(); c.evUsed = 0;
  nccl_destroy();
  if (c.stream) cudaStreamDestroy(c.stream);
  c.stream = nullptr;
  if (c.stream2) cudaStreamDestroy(c.stream2);
  c.stream2 = nullptr;
  if (c.evFork) cudaEventDestroy(c.evFork);
  if (c.evJoin) cudaEventDestroy(c.evJoin);
  c.evFork = c.evJoin = nullptr;
  c.inited = false;
  return 0;
}

int32_t svfsi_lhs_plan_(const int32_t *rank, const int32_t *nranks, const int32_t *gnNo,
                        const int32_t *nNo, const int32_t *maxnNo, const int32_t *aNodes,
                        int32_t *map, int32_t *mynNo, int32_t *shnNo, int32_t *nReq,
                        int32_t *cs_iP, int32_t *cs_n, int32_t *cs_ptr, const int32_t *cs_ptr_cap) {
  try {
    LhsPlan p = lhs_plan(*rank, *nranks, *gnNo, *nNo, *maxnNo, aNodes);
    for (int a = 0; a < *nNo; a++) map[a] = p.map[a] + 1;
    *mynNo = p.mynNo;
    *shnNo = p.shnNo;
    *nReq = (int)p.nbr.size();
    int off = 0;
    for (size_t i = 0; i < p.nbr.size(); i++) {
      cs_iP[i] = p.nbr[i].iP + 1;
      cs_n[i] = (int)p.nbr[i].ptr.size();
      if (off + cs_n[i] > *cs_ptr_cap) return fail(SVFSI_ERR_ARG, "cs_ptr capacity too small");
      for (int v : p.nbr[i].ptr) cs_ptr[off++] = v + 1;
    }
  } catch (const std::exception &ex) {
    return fail(SVFSI_ERR_ARG, ex.what());
  }
  return 0;
}

int32_t gpu_lhs_create_(const int32_t *gnNo_, const int32_t *nNo_, const int32_t *nnz_,
                        const int32_t *gNodes, const int32_t *rowPtr, const int32_t *colPtr,
                        const int32_t *nFaces_) {
  if (int rc = need_init()) return rc;
  Ctx &c = ctx();
  if (c.lhs) return fail(SVFSI_ERR_STATE, "FSILS: LHS is not free. You may use FSILS_LHS_FREE");
  const int nNo = *nNo_, nnz = *nnz_;
  c.gnNo = *gnNo_; c.nNo = nNo; c.nnz = nnz; c.nFaces = *nFaces_;
  c.face.assign(c.nFaces, Face());

  // MPI_ALLREDUCE(MAX) + MPI_ALLGATHERV of the padded node lists (L/LHS.f:113-126)
  std::vector<int32_t> cnt(c.nranks);
  if (int rc = host_allgather_i32(&nNo, 1, cnt.data())) return rc;
  const int maxnNo = *std::max_element(cnt.begin(), cnt.end());
  std::vector<int32_t> part(maxnNo, 0), aNodes((size_t)maxnNo * c.nranks);
  std::copy(gNodes, gNodes + nNo, part.begin());
  if (int rc = host_allgather_i32(part.data(), maxnNo, aNodes.data())) return rc;
  LhsPlan plan;
  try {
    plan = lhs_plan(c.rank, c.nranks, c.gnNo, nNo, maxnNo, aNodes.data());
  } catch (const std::exception &ex) {
    return fail(SVFSI_ERR_ARG, ex.what());
  }
  c.map = plan.map;
  c.mynNo = plan.mynNo;
  c.shnNo = plan.shnNo;

  // device block-CSR in reordered row order, each row in svFSI's original column order
  std::vector<int> inv(nNo);
  for (int a = 0; a < nNo; a++) inv[c.map[a]] = a;
  std::vector<int> rp(nNo + 1, 0), col(nnz), vperm(nnz), rowOf(nnz), diag(nNo, -1);
  for (int r = 0; r < nNo; r++) rp[r + 1] = rp[r] + (rowPtr[inv[r] + 1] - rowPtr[inv[r]]);
  for (int r = 0; r < nNo; r++) {
    const int a = inv[r];
    const int s = rowPtr[a] - 1, len = rowPtr[a + 1] - rowPtr[a];
    for (int k = 0; k < len; k++) {
      const int pd = rp[r] + k;
      const int cr = c.map[colPtr[s + k] - 1];
      col[pd] = cr;
      vperm[s + k] = pd;
      rowOf[pd] = r;
      if (cr == r && diag[r] < 0) diag[r] = pd;
    }
    if (diag[r] < 0) return fail(SVFSI_ERR_ARG, "row without a diagonal entry");
  }
  c.rowPtrDev = rp;
  c.maxRowLen = 0;
  for (int r = 0; r < nNo; r++) c.maxRowLen = std::max(c.maxRowLen, rp[r + 1] - rp[r]);
  if (int rc = dev_upload(&c.d_perm, c.map)) return rc;
  if (int rc = dev_upload(&c.d_rowPtr, rp)) return rc;
  if (int rc = dev_upload(&c.d_col, col)) return rc;
  if (int rc = dev_upload(&c.d_diag, diag)) return rc;
  if (int rc = dev_upload(&c.d_vperm, vperm)) return rc;
  if (int rc = dev_upload(&c.d_rowOf, rowOf)) return rc;

  // halo schedule
  c.nbr.clear();
  std::vector<int> packIdx;
  for (const LhsPlan::Nbr &nb : plan.nbr) {
    Neighbor n;
    n.iP = nb.iP;
    n.n = (int)nb.ptr.size();
    n.off = (int)packIdx.size();
    n.ptr = nb.ptr;
    packIdx.insert(packIdx.end(), nb.ptr.begin(), nb.ptr.end());
    c.nbr.push_back(std::move(n));
  }
  c.nShared = (int)packIdx.size();
  // unique shared nodes -> pack slots in ascending neighbour order (slots are
  // already grouped by ascending neighbour rank)
  std::vector<int> cntU(nNo, 0);
  for (int v : packIdx) cntU[v]++;
  std::vector<int> uniqNode, uniqPtr(1, 0), slotOf(nNo, -1);
  for (int a = 0; a < nNo; a++)
    if (cntU[a]) {
      slotOf[a] = (int)uniqNode.size();
      uniqNode.push_back(a);
      uniqPtr.push_back(uniqPtr.back() + cntU[a]);
    }
  std::vector<int> uniqSlot(packIdx.size()), fill(uniqNode.size(), 0);
  for (size_t s = 0; s < packIdx.size(); s++) {
    const int u = slotOf[packIdx[s]];
    uniqSlot[uniqPtr[u] + fill[u]++] = (int)s;
  }
  c.nUniq = (int)uniqNode.size();
  // the fused halo receive (la_kernels.cu multidot_fused_kernel) relies on what L/LHS.f:134-180 guarantees:
  // the shared nodes are exactly rows [0, shnNo) and [mynNo, nNo)
  c.uniqOrdered = (c.nUniq >= c.shnNo);
  for (int u = 0; u < c.nUniq && c.uniqOrdered; u++)
    c.uniqOrdered = (u < c.shnNo) ? (uniqNode[u] == u) : (uniqNode[u] >= c.mynNo);
  if (int rc = dev_upload(&c.d_packIdx, packIdx)) return rc;
  if (int rc = dev_upload(&c.d_uniqNode, uniqNode)) return rc;
  if (int rc = dev_upload(&c.d_uniqPtr, uniqPtr)) return rc;
  if (int rc = dev_upload(&c.d_uniqSlot, uniqSlot)) return rc;
  dev_free(&c.d_sbuf);
  dev_free(&c.d_rbuf);
  CUDA_TRY(cudaMalloc(&c.d_sbuf, sizeof(double) * 4 * std::max(c.nShared, 1)));
  CUDA_TRY(cudaMalloc(&c.d_rbuf, sizeof(double) * 4 * std::max(c.nShared, 1)));
  c.lhs = true;
  c.lhsGen++;
  c.dof = 0;
  if (int rc = p2p_setup()) return rc;
  return 0;
}

int32_t gpu_lhs_info_(int32_t *mynNo, int32_t *shnNo, int32_t *nReq, int32_t *map) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "no lhs");
  *mynNo = c.mynNo;
  *shnNo = c.shnNo;
  *nReq = (int)c.nbr.size();
  if (map)
    for (int a = 0; a < c.nNo; a++) map[a] = c.map[a] + 1;
  return 0;
}

int32_t gpu_lhs_cs_(const int32_t *i, int32_t *iP, int32_t *n, int32_t *ptr) {
  Ctx &c = ctx();
  if (!c.lhs || *i < 1 || *i > (int)c.nbr.size()) return fail(SVFSI_ERR_ARG, "bad cS index");
  const Neighbor &nb = c.nbr[*i - 1];
  *iP = nb.iP + 1;
  *n = nb.n;
  if (ptr)
    for (int k = 0; k < nb.n; k++) ptr[k] = nb.ptr[k] + 1;
  return 0;
}

int32_t gpu_bc_create_(const int32_t *faIn, const int32_t *nNo_, const int32_t *dof_,
                       const int32_t *BC_type, const int32_t *gNodes, const double *val) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "FSILS_BC_CREATE before FSILS_LHS_CREATE");
  if (*faIn > c.nFaces) return fail(SVFSI_ERR_ARG, "FSILS: faIn is exceeding lhs structure maximum number of face");
  if (*faIn <= 0) return fail(SVFSI_ERR_ARG, "FSILS: faIn should be greater than zero");
  Face &f = c.face[*faIn - 1];
  if (f.created) return fail(SVFSI_ERR_STATE, "FSILS: face is not free, you may use FSILS_BC_FREE to free it");
  const int n = *nNo_, dof = *dof_;
  f.nNo = n; f.dof = dof; f.bGrp = *BC_type;
  if (n < 0 || dof < 1) return fail(SVFSI_ERR_ARG, "FSILS_BC_CREATE: bad nNo / dof");
  f.glob.resize(n);
  for (int a = 0; a < n; a++) {
    if (gNodes[a] < 1 || gNodes[a] > c.nNo)
      return fail(SVFSI_ERR_ARG, "FSILS_BC_CREATE: face node id out of range");
    f.glob[a] = c.map[gNodes[a] - 1];
  }
  std::vector<double> v((size_t)n * dof, 0.0);
  if (val) std::copy(val, val + (size_t)n * dof, v.begin());
  // sharedFlag: more than one rank holds nodes of the face (L/BC.f:99-118)
  f.shared = false;
  if (c.nranks > 1) {
    int32_t mine = n != 0 ? 1 : 0;
    std::vector<int32_t> all(c.nranks);
    if (int rc = host_allgather_i32(&mine, 1, all.data())) return rc;
    int tot = 0;
    for (int x : all) tot += x;
    if (tot > 1) {
      f.shared = true;
      const size_t bytes = sizeof(double) * (size_t)c.nNo * dof;
      if (int rc = ensure_stage(2 * bytes)) return rc;
      std::vector<double> full((size_t)c.nNo * dof, 0.0);
      for (int a = 0; a < n; a++)
        for (int d = 0; d < dof; d++) full[(size_t)f.glob[a] * dof + d] = v[(size_t)a * dof + d];
      double *dv = c.d_stage + (size_t)c.nNo * dof;
      CUDA_TRY(cudaMemcpyAsync(dv, full.data(), bytes, cudaMemcpyHostToDevice, c.stream));
      if (int rc = halo_sum(dv, dof, nullptr)) return rc;
      CUDA_TRY(cudaMemcpyAsync(full.data(), dv, bytes, cudaMemcpyDeviceToHost, c.stream));
      CUDA_TRY(cudaStreamSynchronize(c.stream));
      for (int a = 0; a < n; a++)
        for (int d = 0; d < dof; d++) v[(size_t)a * dof + d] = full[(size_t)f.glob[a] * dof + d];
    }
  }
  if (int rc = dev_upload(&f.d_glob, f.glob)) return rc;
  if (int rc = dev_upload(&f.d_val, v)) return rc;
  if (int rc = dev_upload(&f.d_valM, v)) return rc;
  f.created = true;
  return 0;
}

int32_t gpu_bc_free_(const int32_t *faIn) {
  Ctx &c = ctx();
  if (!c.lhs || *faIn < 1 || *faIn > c.nFaces) return fail(SVFSI_ERR_ARG, "bad faIn");
  Face &f = c.face[*faIn - 1];
  dev_free(&f.d_glob);
  dev_free(&f.d_val);
  dev_free(&f.d_valM);
  f = Face();
  return 0;
}

int32_t gpu_mesh_create_(const int32_t *nEl_, const int32_t *eNoN, const int32_t *IEN,
                         const double *x) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "gpu_mesh_create_ before gpu_lhs_create_");
  if (*eNoN != 4) return fail(SVFSI_ERR_UNSUPPORTED, "only TET4 (eNoN = 4) is implemented");
  const int nEl = *nEl_;
  c.nEl = nEl;
  std::vector<int> ien((size_t)nEl * 4);
  for (size_t k = 0; k < ien.size(); k++) {
    const int a = IEN[k] - 1;
    if (a < 0 || a >= c.nNo) return fail(SVFSI_ERR_ARG, "IEN entry out of range");
    ien[k] = c.map[a];
  }
  if (int rc = dev_upload(&c.d_ien, ien)) return rc;
  dev_free(&c.d_x);
  CUDA_TRY(cudaMalloc(&c.d_x, sizeof(double) * (size_t)c.nNo * 3));
  if (int rc = upload_nodal(x, 3, c.d_x)) return rc;
  dev_free(&c.d_edest);
  CUDA_TRY(cudaMalloc(&c.d_edest, sizeof(int) * (size_t)nEl * 16));
  // every (a,b) of every element must exist in the pattern (DOASSEM would fall off the end of its
  // search, S/LHSA.f:282-292): build_edest counts the missing ones into d_flag[1]
  CUDA_TRY(cudaMemsetAsync(c.d_flag + 1, 0, sizeof(int), c.stream));
  launch_build_edest(c.stream, nEl, c.d_ien, c.d_rowPtr, c.d_col, c.d_edest, c.d_flag + 1);
  {
    int missing = 0;
    CUDA_TRY(cudaMemcpyAsync(&missing, c.d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    if (missing)
      return fail(SVFSI_ERR_ARG, "gpu_mesh_create_: " + std::to_string(missing) +
                                     " element node pairs are not in the rowPtr/colPtr pattern given to "
                                     "gpu_lhs_create_");
  }
  std::vector<int> colorElems;
  try {
    color_elements(nEl, c.nNo, ien, c.colorOff, colorElems);
  } catch (const std::exception &ex) {
    return fail(SVFSI_ERR_ARG, ex.what());
  }
  c.ncolors = (int)c.colorOff.size() - 1;
  if (int rc = dev_upload(&c.d_colorElems, colorElems)) return rc;
  dev_free(&c.d_blkAdjPtr); dev_free(&c.d_blkAdj); dev_free(&c.d_nodeAdjPtr);
  dev_free(&c.d_nodeAdj); dev_free(&c.d_blkOrder); dev_free(&c.d_elemP); dev_free(&c.d_nodeSlots);
  dev_free(&c.d_pairList); dev_free(&c.d_pairT); c.nPair = 0; dev_free(&c.d_blkDesc);
  dev_free(&c.d_pairDesc); c.nPairDescOff = 0;
  if (int rc = build_gather_adjacency(c.stream, nEl, c.nNo, c.nnz, c.d_ien, c.d_edest,
                                      &c.d_blkAdjPtr, &c.d_blkAdj, &c.d_nodeAdjPtr, &c.d_nodeAdj,
                                      &c.d_blkOrder))
    return rc;
  dev_free(&c.d_nodeSlots);
  if (c.maxRowLen <= 64) {
    CUDA_TRY(cudaMalloc(&c.d_nodeSlots, sizeof(int) * (size_t)nEl * 4));
    launch_build_node_slots(c.stream, c.nNo, c.d_rowPtr, c.d_nodeAdjPtr, c.d_nodeAdj, c.d_edest,
                            c.d_nodeSlots);
  }
  if (int rc = build_pair_lists(c.stream, c.nnz, c.d_blkOrder, c.d_rowOf, c.d_col, c.d_rowPtr,
                                &c.d_pairList, &c.d_pairT, &c.nPair))
    return rc;
  if (int rc = build_block_desc(c.stream, c.nnz, c.d_blkOrder, c.d_blkAdjPtr, &c.d_blkDesc, c.d_rowOf, c.d_col))
    return rc;
  if (int rc = build_paired_desc(c.stream, c.nPair, c.d_pairList, c.d_pairT, c.d_blkAdjPtr, c.d_rowOf, c.nnz,
                                 &c.d_pairDesc, &c.nPairDescOff))
    return rc;
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  c.mesh = true;
  return 0;
}

int32_t gpu_mesh_ncolors_(int32_t *ncolors) {
  *ncolors = ctx().ncolors;
  return 0;
}

int32_t gpu_state_upload_(const int32_t *tDof, const double *Ag, const double *Yg,
                          const double *Bf) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "no lhs");
  if (int rc = ensure_state()) return rc;
  if (int rc = upload_nodal(Ag, *tDof, c.d_Ag)) return rc;
  if (int rc = upload_nodal(Yg, *tDof, c.d_Yg)) return rc;
  g_haveBf = false;
  if (Bf && *tDof == 4) {
    if (int rc = upload_nodal(Bf, 3, c.d_Bf)) return rc;
    g_haveBf = true;
  }
  return 0;
}

int32_t gpu_construct_fluid_dev_(const double *rho, const double *mu, const double *f,
                                 const double *dt, const double *af, const double *am,
                                 const double *gam, const int32_t *variant) {
  FluidPar par{*rho, *mu, {f[0], f[1], f[2]}, *dt, *af, *am, *gam};
  return run_fluid_asm(par, *variant);
}

int32_t gpu_construct_fluid_(const double *Ag, const double *Yg, const double *Bf,
                             const double *rho, const double *mu, const double *f,
                             const double *dt, const double *af, const double *am,
                             const double *gam, const int32_t *variant) {
  const int32_t four = 4;
  if (int rc = gpu_state_upload_(&four, Ag, Yg, Bf)) return rc;
  if (int rc = gpu_construct_fluid_dev_(rho, mu, f, dt, af, am, gam, variant)) return rc;
  return check_jac();
}

int32_t gpu_construct_heats_dev_(const double *nu, const double *s, const double *rho,
                                 const double *dt, const double *af, const double *am,
                                 const double *gam, const int32_t *variant) {
  HeatPar par{*nu, *s, *rho, *dt, *af, *am, *gam};
  return run_heat_asm(par, *variant);
}

int32_t gpu_construct_heats_(const double *Ag, const double *Yg, const double *nu,
                             const double *s, const double *rho, const double *dt,
                             const double *af, const double *am, const double *gam,
                             const int32_t *variant) {
  const int32_t one = 1;
  if (int rc = gpu_state_upload_(&one, Ag, Yg, nullptr)) return rc;
  if (int rc = gpu_construct_heats_dev_(nu, s, rho, dt, af, am, gam, variant)) return rc;
  return check_jac();
}

int32_t gpu_get_r_(const int32_t *dof, double *R) {
  Ctx &c = ctx();
  if (!c.d_R) return fail(SVFSI_ERR_STATE, "no device residual");
  return download_nodal(c.d_R, *dof, R);
}
int32_t gpu_set_r_(const int32_t *dof, const double *R) {
  if (int rc = ensure_system(*dof)) return rc;
  return upload_nodal(R, *dof, ctx().d_R);
}
int32_t gpu_get_val_(const int32_t *dof, double *Val) {
  Ctx &c = ctx();
  if (!c.d_Val) return fail(SVFSI_ERR_STATE, "no device matrix");
  return download_val(c.d_Val, *dof * *dof, Val);
}
int32_t gpu_set_val_(const int32_t *dof, const double *Val) {
  if (int rc = ensure_system(*dof)) return rc;
  return upload_val(Val, *dof * *dof, ctx().d_Val);
}

int32_t gpu_commu_dev_(const int32_t *dof) {
  Ctx &c = ctx();
  if (!c.d_R) return fail(SVFSI_ERR_STATE, "no device residual");
  return halo_sum(c.d_R, *dof, nullptr);
}

int32_t gpu_commu_(const int32_t *dof, double *R) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "no lhs");
  if (c.nranks == 1) return 0;  // S/ALLFUN.f:514-533 returns early on one rank
  const size_t bytes = sizeof(double) * (size_t)c.nNo * *dof;
  if (int rc = ensure_stage(2 * bytes)) return rc;
  double *dv = c.d_stage + (size_t)c.nNo * *dof;
  CUDA_TRY(cudaMemcpyAsync(c.d_stage, R, bytes, cudaMemcpyHostToDevice, c.stream));
  launch_perm_scatter(c.stream, c.nNo, *dof, c.d_perm, c.d_stage, dv);
  if (int rc = halo_sum(dv, *dof, nullptr)) return rc;
  launch_perm_gather(c.stream, c.nNo, *dof, c.d_perm, dv, c.d_stage);
  CUDA_TRY(cudaMemcpyAsync(R, c.d_stage, bytes, cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  return comm_check();
}

int32_t gpu_ls_create_(svfsi_ls_t *ls, const int32_t *LS_type) {
  switch (*LS_type) {
    case SVFSI_LS_TYPE_NS: case SVFSI_LS_TYPE_GMRES: case SVFSI_LS_TYPE_CG: case SVFSI_LS_TYPE_BICGS:
      ls_defaults(ls, *LS_type);
      return 0;
    default:
      return fail(SVFSI_ERR_ARG, "FSILS: LS_TYPE is not defined");
  }
}

int32_t gpu_solve_dev_(svfsi_ls_t *ls, const int32_t *dof, const int32_t *prec,
                       const int32_t *incL, const double *res) {
  if (int rc = need_init()) return rc;
  if (int rc = fsils_solve_dev(ls, *dof, *prec, incL, res)) return rc;
  CUDA_TRY(cudaStreamSynchronize(ctx().stream));
  if (int rc = comm_check()) return rc;
  return jac_check();   // of the element loop that produced this system (device-resident path)
}

int32_t gpu_solve_(svfsi_ls_t *ls, const int32_t *dof, double *Ri, const double *Val,
                   const int32_t *prec, const int32_t *incL, const double *res) {
  if (int rc = need_init()) return rc;
  if (int rc = gpu_set_r_(dof, Ri)) return rc;
  if (Val)
    if (int rc = gpu_set_val_(dof, Val)) return rc;
  if (int rc = fsils_solve_dev(ls, *dof, *prec, incL, res)) return rc;
  if (int rc = gpu_get_r_(dof, Ri)) return rc;
  return comm_check();
}

int32_t gpu_sparmul_(const int32_t *kind, const int32_t *dof, const double *K, const double *U,
                     double *KU) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "no lhs");
  const int k = *kind, d = (k == 3) ? 1 : *dof;
  const int br = row_dof(k, d), bc = col_dof(k, d);
  const size_t nK = (size_t)c.nnz * br * bc, nU = (size_t)c.nNo * bc, nKU = (size_t)c.nNo * br;
  double *dK = nullptr, *dU = nullptr, *dKU = nullptr;
  CUDA_TRY(cudaMalloc(&dK, nK * sizeof(double)));
  CUDA_TRY(cudaMalloc(&dU, nU * sizeof(double)));
  CUDA_TRY(cudaMalloc(&dKU, nKU * sizeof(double)));
  int rc = upload_val(K, br * bc, dK);
  if (!rc) rc = upload_nodal(U, bc, dU);
  if (!rc) rc = sparmul(k, d, dK, dU, dKU, nullptr);
  if (!rc) rc = download_nodal(dKU, br, KU);
  if (!rc) rc = comm_check();
  cudaFree(dK); cudaFree(dU); cudaFree(dKU);
  return rc;
}

int32_t gpu_dot_(const int32_t *dof, const double *U, const double *V, double *result) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "no lhs");
  if (int rc = ensure_small()) return rc;
  const size_t n = (size_t)c.nNo * *dof;
  double *dU = nullptr, *dV = nullptr;
  CUDA_TRY(cudaMalloc(&dU, n * sizeof(double)));
  CUDA_TRY(cudaMalloc(&dV, n * sizeof(double)));
  int rc = upload_nodal(U, *dof, dU);
  if (!rc) rc = upload_nodal(V, *dof, dV);
  if (!rc) {
    launch_multidot(c.stream, dU, 0, dV, (size_t)c.mynNo * *dof, 1, c.d_partial, nullptr);
    rc = reduce_allreduce(c.d_partial, 1, c.d_small + 64, nullptr);
  }
  if (!rc) {
    cudaMemcpyAsync(result, c.d_small + 64, sizeof(double), cudaMemcpyDeviceToHost, c.stream);
    cudaStreamSynchronize(c.stream);
    rc = comm_check();
  }
  cudaFree(dU); cudaFree(dV);
  return rc;
}

int32_t gpu_time_kernel_(const int32_t *what, const int32_t *dof, const int32_t *k,
                         const int32_t *reps, const int32_t *variant, double *ms_total) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "no lhs");
  if (int rc = ensure_small()) return rc;
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  const int d = *dof;
  const size_t n = (size_t)c.nNo * d;
  const size_t stride = ((n * 8 + 255) / 256) * 256 / 8;
  int rc = 0;
  if (*what == 0) {
    if (!c.d_Val) return fail(SVFSI_ERR_STATE, "no device matrix");
    if ((rc = ensure_ws(2 * stride * sizeof(double)))) return rc;
    double *U = c.d_ws, *KU = c.d_ws + stride;
    launch_vecop(c.stream, VOP_ZERO, U, nullptr, nullptr, 2 * stride, nullptr, 0.0, nullptr);
    // variant: 0 = the 8-lanes-per-row SPARMULVV kernel, 1 = the 4-lanes-per-row one, < 0 = as configured
    const int prevQuad = (*variant >= 0) ? set_spmv_quad(*variant) : -1;
    launch_spmv(c.stream, d == 1 ? 3 : 0, d, 0, c.nNo, c.d_rowPtr, c.d_col, c.d_Val, U, KU, nullptr);
    CUDA_TRY(cudaEventRecord(a, c.stream));
    for (int r = 0; r < *reps; r++)
      launch_spmv(c.stream, d == 1 ? 3 : 0, d, 0, c.nNo, c.d_rowPtr, c.d_col, c.d_Val, U, KU, nullptr);
    CUDA_TRY(cudaEventRecord(b, c.stream));
    if (prevQuad >= 0) set_spmv_quad(prevQuad);
  } else if (*what == 6) {
    // small-shape SpMV: k = kind (0 VV, 1 VS, 2 SV, 3 SS), variant = SVFSI_SPMV_SMALL mode (0 = lane-per-block,
    // 1..8 = the run / run-async / hoisted configurations); K = the resident Val buffer (any contents), one rank's rows only
    if (!c.d_Val) { if ((rc = ensure_system(4))) return rc; }
    if (*k < 0 || *k > 3 || d < 1 || d > 4) return fail(SVFSI_ERR_ARG, "gpu_time_kernel_(6): kind / dof");
    if ((rc = ensure_ws(2 * stride * sizeof(double)))) return rc;
    double *U = c.d_ws, *KU = c.d_ws + stride;
    launch_vecop(c.stream, VOP_ZERO, U, nullptr, nullptr, 2 * stride, nullptr, 0.0, nullptr);
    const int prev = set_spmv_small(*variant);
    launch_spmv(c.stream, *k, d, 0, c.nNo, c.d_rowPtr, c.d_col, c.d_Val, U, KU, nullptr);
    CUDA_TRY(cudaEventRecord(a, c.stream));
    for (int r = 0; r < *reps; r++)
      launch_spmv(c.stream, *k, d, 0, c.nNo, c.d_rowPtr, c.d_col, c.d_Val, U, KU, nullptr);
    CUDA_TRY(cudaEventRecord(b, c.stream));
    set_spmv_small(prev);
  } else if (*what == 3 || *what == 4) {
    const int kk = *k;
    if ((rc = ensure_ws((size_t)(kk + 1) * stride * sizeof(double)))) return rc;
    double *U = c.d_ws, *w = c.d_ws + (size_t)kk * stride;
    launch_vecop(c.stream, VOP_ZERO, U, nullptr, nullptr, (size_t)(kk + 1) * stride, nullptr, 0.0, nullptr);
    CUDA_TRY(cudaEventRecord(a, c.stream));
    for (int r = 0; r < *reps; r++) {
      if (*what == 3 && *variant == 1) {   // the fused column kernel (multi-dot + block sums in one launch)
        if ((rc = multidot_column(U, stride, w, n, kk, c.d_small + 64, nullptr, nullptr, false))) return rc;
      } else if (*what == 3) launch_multidot(c.stream, U, stride, w, n, kk, c.d_partial, nullptr);
      else launch_multi_axpy_scale(c.stream, U, stride, w, n, kk, c.d_small + 64, nullptr, nullptr);
    }
    CUDA_TRY(cudaEventRecord(b, c.stream));
  } else if (*what == 5) {
    // one kernel of the gather assembly: k = part mask (1 records, 2 tangent gather, 4 residual
    // gather), variant = kernel-variant mask (asm_tune()); uses the resident Ag / Yg state
    if (!c.mesh || !c.d_Ag || !c.d_Yg || !c.d_elemP || !c.d_R || !c.d_Val)
      return fail(SVFSI_ERR_STATE, "gpu_time_kernel_(5): run the gather assembly once first");
    FluidPar par;
    par.rho = 1.06; par.mu = 0.04; par.f[0] = par.f[1] = par.f[2] = 0.0;
    par.dt = 5e-3; par.af = 1.0 / 1.2; par.am = 0.5 * 2.8 / 1.2; par.gam = 0.5 + par.am - par.af;
    CUDA_TRY(cudaEventRecord(a, c.stream));
    for (int r = 0; r < *reps; r++)
      if ((rc = fluid_gather_dispatch(*k, par, *variant))) return rc;
    CUDA_TRY(cudaEventRecord(b, c.stream));
  } else {
    return fail(SVFSI_ERR_ARG, "gpu_time_kernel_: unknown kernel id");
  }
  CUDA_TRY(cudaEventSynchronize(b));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
  *ms_total = ms;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return 0;
}

int32_t gpu_prof_enable_(const int32_t *on) {
  ctx().prof = (*on != 0);
  return 0;
}
int32_t gpu_prof_reset_(void) {
  Ctx &c = ctx();
  prof_collect();
  for (int i = 0; i < PROF_NSLOTS; i++) { c.profMs[i] = 0.0; c.profN[i] = 0; }
  c.profSpmvBytes = 0.0;
  c.profSpmvOps = 0;
  return 0;
}
int32_t gpu_prof_get_(double *ms, int64_t *launches) {
  Ctx &c = ctx();
  prof_collect();
  for (int i = 0; i < PROF_NSLOTS; i++) { ms[i] = c.profMs[i]; launches[i] = c.profN[i]; }
  return 0;
}
int32_t gpu_prof_spmv_(double *bytes, int64_t *ops) {
  *bytes = ctx().profSpmvBytes;
  *ops = ctx().profSpmvOps;
  return 0;
}
int32_t gpu_launch_count_(int64_t *n) {
  *n = ctx().launches;
  return 0;
}
int32_t gpu_set_spmv_small_(const int32_t *mode) {
  if (*mode < -1 || *mode > 13) return fail(SVFSI_ERR_ARG, "gpu_set_spmv_small_: mode must be -1..13");
  set_spmv_small(*mode);
  return 0;
}
int32_t gpu_spmv_variant_(int32_t *variant) {
  const int prev = set_spmv_quad(1);   // read ...
  set_spmv_quad(prev);                 // ... and put back
  *variant = prev | (spmv_fused_quad_enabled() ? 2 : 0);
  return 0;
}
int32_t gpu_comm_mode_(int32_t *mode) {
  Ctx &c = ctx();
  if (c.nranks == 1) *mode = 0;
  else if (!c.p2p.on) *mode = 1;
  else *mode = c.p2p.fuse ? 3 : 2;
  return 0;
}
int32_t gpu_get_stream_(void **stream) {
  *stream = (void *)ctx().stream;
  return 0;
}
int32_t gpu_sync_(void) {
  if (int rc = need_init()) return rc;
  CUDA_TRY(cudaStreamSynchronize(ctx().stream));
  if (int rc = comm_check()) return rc;
  return jac_check();
}
int32_t gpu_set_comm_timeout_(const double *seconds) {
  if (!(*seconds > 0.0)) return fail(SVFSI_ERR_ARG, "gpu_set_comm_timeout_: seconds must be > 0");
  ctx().commTimeoutS = *seconds;
  return 0;
}

}  // extern "C"
