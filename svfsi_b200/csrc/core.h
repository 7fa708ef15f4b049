// core.h -- internal host-side services: NCCL (dlopen'd), halo sum, all-reduce,
// workspace, profiling events.
#pragma once
#include "ctx.h"
#include "kernels.h"

namespace svfsi {

// ---- NCCL via dlopen (only touched when nranks > 1) ----
int nccl_load();                                   // 0 ok
int nccl_unique_id(void *uid128);
int nccl_init_rank(const void *uid128, int nranks, int rank);
void nccl_destroy();
int nccl_allreduce_sum(double *buf, size_t n, cudaStream_t st);      // in place
int nccl_allgather_i32(const int32_t *send, int32_t n, int32_t *recv, cudaStream_t st);  // device bufs
int nccl_sendrecv(const double *sbuf, double *rbuf, const std::vector<Neighbor> &nbr, int dof,
                  cudaStream_t st);

// ---- collectives on the library stream ----
// FSILS_COMMUV/S on a device vector in FSILS order (L/INCOMMU.f:56-151)
int halo_sum(double *R, int dof, const int *done);
// MPI_ALLREDUCE(SUM) of n doubles living on the device (L/BCAST.f, L/DOT.f, L/NORM.f)
int allreduce_dev(double *buf, size_t n);
// out[j] = all-reduce(sum_b partial[j*nblk+b]), j < k (one kernel on the peer-memory path)
int reduce_allreduce(const double *partial, int k, double *out, const int *done);
int reduce_allreduce_column(const double *partial, int k, double *out, const ColArgs &col,
                            const int *done);
// One Gram-Schmidt column: out[j] = all-reduce(<U_j, w> over owned rows), j < k, optionally preceded by
// the halo receive that a sparmul(..., &deferred) left pending on w and followed by the GMRES column step
// (col != NULL) -- ONE kernel on the single-rank and peer-memory paths (launch_multidot_fused), the
// separate kernels otherwise.
int multidot_column(const double *U, size_t stride, double *w, size_t nOwned, int k, double *out,
                    const ColArgs *col, const int *done, bool recvPending);
// peer-memory arena (IPC) set up after the halo schedule is known; falls back to NCCL
int p2p_setup();
void p2p_teardown();
// host-side all-gather of int32 for the setup phase
int host_allgather_i32(const int32_t *send, int32_t n, int32_t *recv);

// ---- full FSILS_SPARMUL*: kernel + halo sum ----
// deferRecv != NULL: on the fused peer-memory path the final halo RECEIVE is left to the caller's next
// multidot_column(..., recvPending = *deferRecv) on KU (set to true when it was deferred)
int sparmul(int kind, int dof, const double *K, const double *U, double *KU, const int *done,
            bool *deferRecv = nullptr);
void trace_dump();                 // SVFSI_TRACE_FILE: write the column time stamps (gpu_finalize_)
int flush_val_scale();            // apply PRECONDDIAG's scaling of Val if it is still pending
int row_dof(int kind, int dof);  // dof of KU
int col_dof(int kind, int dof);  // dof of U

// ---- sticky error words (ctx.h: d_flag / h_status) ----
int status_words_init();   // allocate d_flag + the mapped host words (gpu_init_)
void status_words_free();
// after a host synchronisation: SVFSI_ERR_COMM if an in-kernel peer-flag wait timed out (sticky until
// gpu_finalize_), SVFSI_ERR_JAC if an element loop met ISZERO(Jac) since the last report
int comm_check();
int jac_check();           // host already synchronised with the copy queued by jac_publish()
void jac_publish();        // queue d_flag[0] -> h_status[1] behind the element kernels (no sync)

// ---- workspace ----
int ensure_ws(size_t bytes);          // d_ws >= bytes (contents not preserved)
int ensure_stage(size_t bytes);
int ensure_small();

// ---- profiling (CUDA events on the library stream) ----
struct ProfScope {
  int slot;
  bool on;
  size_t idx;
  explicit ProfScope(int slot);
  ~ProfScope();
};
void prof_collect();  // sync + fold event pairs into ctx.profMs

// ---- solver (solver.cu) ----
int fsils_solve_dev(svfsi_ls_t *ls, int dof, int prec, const int32_t *incL, const double *res);
void ls_defaults(svfsi_ls_t *ls, int LS_type);

// upload / download with permutation between svFSI and FSILS layouts
int upload_nodal(const double *host, int m, double *dev);   // dev[perm[a]][m] = host[a][m]
int download_nodal(const double *dev, int m, double *host);
int upload_val(const double *host, int dd, double *dev);    // by block positions (vperm)
int download_val(const double *dev, int dd, double *host);

}  // namespace svfsi
