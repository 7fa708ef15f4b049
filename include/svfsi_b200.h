/*
 * svfsi_b200.h -- C-ABI of the B200-native svFSI fluid Newton-iteration hot path.
 *
 * Drop-in boundary.  Every entry point is `extern "C"`, uses the
 * gfortran external-name convention of the reference's own optional back-end
 * seam (lower case + trailing underscore, every argument BY REFERENCE, arrays
 * as raw pointers to Fortran-owned, column-major, 1-based-id host memory --
 * Code/Source/svFSI/trilinos_linear_solver.h:185-219) and returns an int32
 * error code (0 = ok) instead of the reference's PRINT + STOP.  No torch / C++
 * types cross the boundary.  State is one process-wide context (one MPI rank
 * <-> one GPU, like the file-scope statics of trilinos_linear_solver.cpp:46-99).
 *
 * Citations are relative to /root/reference/Code/Source (S/ = svFSI/, L/ =
 * svFSILS/).  INTEGER(KIND=LSIP) = int32_t, REAL(KIND=LSRP) = double
 * (L/FSILS_TYPEDEF.h:55-57).
 */
#ifndef SVFSI_B200_H
#define SVFSI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums, L/FSILS_STRUCT.h:53-62 ------------------------------------ */
#define SVFSI_LS_TYPE_CG 798
#define SVFSI_LS_TYPE_GMRES 797
#define SVFSI_LS_TYPE_NS 796
#define SVFSI_LS_TYPE_BICGS 795
#define SVFSI_PRECOND_FSILS 701
#define SVFSI_PRECOND_RCS 709
#define SVFSI_BC_TYPE_DIR 0
#define SVFSI_BC_TYPE_NEU 1

/* assembly variants (north_star: atomic and deterministic graph-coloured) */
#define SVFSI_ASM_ATOMIC 0
#define SVFSI_ASM_COLORED 1
#define SVFSI_ASM_GATHER 2

/* error codes */
#define SVFSI_OK 0
#define SVFSI_ERR_CUDA 1      /* CUDA runtime / no device: the product path has no CPU fallback */
#define SVFSI_ERR_STATE 2     /* call order (e.g. solve before lhs_create) */
#define SVFSI_ERR_ARG 3       /* bad argument (faIn out of range, unknown LS_type, ...) */
#define SVFSI_ERR_JAC 4       /* ISZERO(Jac) at some element, S/FLUID.f:115; from the host-buffer
                               * gpu_construct_*_ at once, from the device-resident gpu_construct_*_dev_
                               * at the next call that synchronises (gpu_solve_dev_, gpu_sync_) */
#define SVFSI_ERR_COMM 5      /* NCCL / host collective */
#define SVFSI_ERR_UNSUPPORTED 6

/* FSILS_subLsType / FSILS_lsType (L/FSILS_STRUCT.h:159-204) flattened to
 * BIND(C)-compatible PODs: LOGICAL -> int32.  IN: mItr, sD, absTol, relTol,
 * LS_type.  OUT: suc, itr, iNorm, fNorm, dB, callD, Resm, Resc. */
typedef struct {
  int32_t suc, mItr, sD, itr;
  double absTol, relTol, iNorm, fNorm, dB, callD;
} svfsi_subls_t;

typedef struct {
  int32_t LS_type, Resm, Resc, reserved;
  svfsi_subls_t GM, CG, RI;
} svfsi_ls_t;

/* Host-side collective the caller lends to the library for the SETUP
 * collectives of FSILS_LHS_CREATE / FSILS_BC_CREATE (MPI_ALLREDUCE /
 * MPI_ALLGATHERV at L/LHS.f:113,125,216 and L/BC.f:102): every rank
 * contributes n int32, recv gets nranks*n in rank order.  The Fortran shim
 * wraps MPI_ALLGATHER; bench.py / tests wrap torch.distributed (nccl / gloo).
 * If none is registered and nranks > 1 the library uses its NCCL communicator. */
typedef int (*svfsi_allgather_i32_fn)(void *ctx, const int32_t *send, int32_t n,
                                      int32_t *recv);

/* ---- life cycle ---------------------------------------------------------- */
/* rank 0 obtains the 128-byte NCCL unique id; the caller broadcasts it (MPI_BCAST
 * in the shim) and passes it to gpu_init_ on every rank.  Replaces the role of
 * FSILS_COMMU_CREATE (L/COMMU.f:50-86): rank, size, communicator. */
int32_t gpu_nccl_unique_id_(void *uid128);
int32_t gpu_init_(const int32_t *device, const int32_t *rank, const int32_t *nranks,
                  const void *uid128 /* may be NULL when nranks == 1 */);
int32_t gpu_set_host_allgather_(svfsi_allgather_i32_fn fn, void *ctx);
int32_t gpu_finalize_(void);
/* last error text (NUL terminated, truncated to *len) */
int32_t gpu_last_error_(char *buf, const int32_t *len);

/* ---- FSILS_LHS_CREATE, L/LHS.f:51-293 ------------------------------------ */
/* gNodes = ltg(nNo) global ids, rowPtr(nNo+1), colPtr(nnz): svFSI's CSR of
 * S/LHSA.f:38-262, all 1-based. */
int32_t gpu_lhs_create_(const int32_t *gnNo, const int32_t *nNo, const int32_t *nnz,
                        const int32_t *gNodes, const int32_t *rowPtr,
                        const int32_t *colPtr, const int32_t *nFaces);
/* FSILS_LHS_FREE, L/LHS.f:295-328 */
int32_t gpu_lhs_free_(void);
/* introspection used by the parity tests: lhs%mynNo, shnNo, nReq, map(nNo) */
int32_t gpu_lhs_info_(int32_t *mynNo, int32_t *shnNo, int32_t *nReq, int32_t *map);
/* cS(i)%iP, %n, %ptr(:) for i = 1..nReq (ptr may be NULL to query n) */
int32_t gpu_lhs_cs_(const int32_t *i, int32_t *iP, int32_t *n, int32_t *ptr);

/* Host-only (no CUDA call): the reordering / halo-schedule part of
 * FSILS_LHS_CREATE for one rank given every rank's gNodes list; used by the CPU
 * tests and by gpu_lhs_create_ itself.  aNodes = [nranks][maxnNo] zero padded
 * (exactly the MPI_ALLGATHERV buffer of L/LHS.f:125).  Outputs: map(nNo),
 * mynNo, shnNo, nReq, and the neighbour table flattened as
 * cs_iP[nReq], cs_n[nReq], cs_ptr[sum n] (caller sizes cs_* by nranks / nranks*nNo...
 * use svfsi_lhs_plan_size_ first). */
int32_t svfsi_lhs_plan_(const int32_t *rank, const int32_t *nranks, const int32_t *gnNo,
                        const int32_t *nNo, const int32_t *maxnNo, const int32_t *aNodes,
                        int32_t *map, int32_t *mynNo, int32_t *shnNo, int32_t *nReq,
                        int32_t *cs_iP, int32_t *cs_n, int32_t *cs_ptr,
                        const int32_t *cs_ptr_cap);

/* ---- FSILS_BC_CREATE, L/BC.f:50-121 / FSILS_BC_FREE :195 ------------------ */
/* val(dof,nNo) may be NULL (-> zeros, :91-97).  gNodes = svFSI local ids. */
int32_t gpu_bc_create_(const int32_t *faIn, const int32_t *nNo, const int32_t *dof,
                       const int32_t *BC_type, const int32_t *gNodes, const double *val);
int32_t gpu_bc_free_(const int32_t *faIn);

/* ---- mesh data of the element loop (msh%IEN, x; S/MOD.f:977-1005) -------- */
/* IEN(eNoN,nEl) 1-based svFSI local node ids, x(nsd,tnNo); eNoN = 4 (TET4),
 * nsd = 3.  Builds the per-element scatter map (replaces the per-entry binary
 * search of DOASSEM, S/LHSA.f:282-292) and the element colouring. */
int32_t gpu_mesh_create_(const int32_t *nEl, const int32_t *eNoN, const int32_t *IEN,
                         const double *x);
int32_t gpu_mesh_ncolors_(int32_t *ncolors);

/* ---- GLOBALEQASSEM -> CONSTRUCT_FLUID, S/EQASSEM.f:46-47, S/FLUID.f:40-190 - */
/* LSALLOC's R = 0, Val = 0 (S/LS.f:44-51) is folded in.  Ag, Yg = (tDof=4,tnNo),
 * Bf = (3,tnNo) or NULL (zeros); prop: rho, mu (viscType_Const), f(3); dt and the
 * generalised-alpha af, am, gam of eq(cEq).  Results stay on the device
 * (R(4,tnNo), Val(16,nnz)). */
int32_t gpu_construct_fluid_(const double *Ag, const double *Yg, const double *Bf,
                             const double *rho, const double *mu, const double *f,
                             const double *dt, const double *af, const double *am,
                             const double *gam, const int32_t *variant);
/* CONSTRUCT_HEATS, S/HEATS.f:39-113: Ag, Yg = (1,tnNo); nu, s, rho. */
int32_t gpu_construct_heats_(const double *Ag, const double *Yg, const double *nu,
                             const double *s, const double *rho, const double *dt,
                             const double *af, const double *am, const double *gam,
                             const int32_t *variant);
/* Same element loops on state already resident on the device (uploaded once by
 * gpu_state_upload_): the kernel-only leg of bench.py. */
int32_t gpu_state_upload_(const int32_t *tDof, const double *Ag, const double *Yg,
                          const double *Bf);
int32_t gpu_construct_fluid_dev_(const double *rho, const double *mu, const double *f,
                                 const double *dt, const double *af, const double *am,
                                 const double *gam, const int32_t *variant);
int32_t gpu_construct_heats_dev_(const double *nu, const double *s, const double *rho,
                                 const double *dt, const double *af, const double *am,
                                 const double *gam, const int32_t *variant);

/* device <-> host access to the assembled system in svFSI's layout (R(dof,tnNo),
 * Val(dof*dof,nnz) in the rowPtr/colPtr order handed to gpu_lhs_create_):
 * needed while SETBCNEU (S/MAIN.f:147) still scatters face terms on the host. */
int32_t gpu_get_r_(const int32_t *dof, double *R);
int32_t gpu_set_r_(const int32_t *dof, const double *R);
int32_t gpu_get_val_(const int32_t *dof, double *Val);
int32_t gpu_set_val_(const int32_t *dof, const double *Val);

/* ---- COMMU(R), S/ALLFUN.f:514-533 -> FSILS_COMMUV/S, L/INCOMMU.f:56-151 --- */
/* host array in svFSI node order, summed over the ranks sharing each node */
int32_t gpu_commu_(const int32_t *dof, double *R);
/* same on the device-resident residual (S/MAIN.f:161) */
int32_t gpu_commu_dev_(const int32_t *dof);

/* ---- LSSOLVE -> FSILS_SOLVE, S/LS.f:98-99, L/SOLVE.f:51-143 -------------- */
/* Ri(dof,nNo): in RHS, out solution (svFSI node order).  Val(dof*dof,nnz) host
 * matrix, or NULL to use the device-resident one left by gpu_construct_*_.
 * incL(nFaces) / res(nFaces) may be NULL (OPTIONAL in the reference). */
int32_t gpu_solve_(svfsi_ls_t *ls, const int32_t *dof, double *Ri, const double *Val,
                   const int32_t *prec, const int32_t *incL, const double *res);
/* device-resident R / Val in, solution left on the device in place of R;
 * gpu_get_r_ fetches it. */
int32_t gpu_solve_dev_(svfsi_ls_t *ls, const int32_t *dof, const int32_t *prec,
                       const int32_t *incL, const double *res);
/* FSILS_LS_CREATE defaults, L/LS.f:50-119 */
int32_t gpu_ls_create_(svfsi_ls_t *ls, const int32_t *LS_type);

/* ---- building blocks exported for parity tests and roofline measurement -- */
/* FSILS_SPARMULVV / VS / SV / SS (L/SPARMUL.f:51-297) incl. the halo sum.
 * K, U, KU are host arrays in the library's (reordered) FSILS layout when
 * `reordered` != 0, else svFSI layout. kind: 0 = VV, 1 = VS, 2 = SV, 3 = SS. */
int32_t gpu_sparmul_(const int32_t *kind, const int32_t *dof, const double *K,
                     const double *U, double *KU);
/* FSILS_DOTV / NORMV over owned nodes + allreduce (L/DOT.f:56, L/NORM.f:55) */
int32_t gpu_dot_(const int32_t *dof, const double *U, const double *V, double *result);

/* timing: device time (ms, CUDA events on the library stream) of `reps`
 * back-to-back launches of one kernel on the resident system.
 * what: 0 = SPARMULVV kernel (dof), 1 = fluid assembly, 2 = heat assembly,
 * 3 = fused multi-dot (k vectors), 4 = fused multi-axpy (k vectors),
 * 5 = one kernel of the gather assembly (k = part mask, variant = kernel-variant mask),
 * 6 = small-shape SpMV (k = kind 0 VV / 1 VS / 2 SV / 3 SS, variant = kernel family, see below). */
int32_t gpu_time_kernel_(const int32_t *what, const int32_t *dof, const int32_t *k,
                         const int32_t *reps, const int32_t *variant, double *ms_total);
/* kernel family of the small-block SpMV shapes (FSILS_SPARMULVV dof <= 3, VS, SV, SS; L/SPARMUL.f:135-297):
 * -1 = per-shape default, 0 = lane-per-block, 1..13 = contiguous-run / asynchronous-run / hoisted configurations
 * (same switch as the environment variable SVFSI_SPMV_SMALL).  Results differ by summation order only. */
int32_t gpu_set_spmv_small_(const int32_t *mode);
/* per-phase device times (ms) and launch counts accumulated since the last
 * reset; see svfsi_b200/csrc/ctx.h for the slot names. */
#define SVFSI_NTIMERS 16
int32_t gpu_prof_enable_(const int32_t *on);
int32_t gpu_prof_reset_(void);
int32_t gpu_prof_get_(double *ms /*[SVFSI_NTIMERS]*/, int64_t *launches /*[SVFSI_NTIMERS]*/);
/* total number of kernels this library launched since gpu_init_ */
/* ---- generalised-alpha time integration on the device --------------------
 * Replaces, for the equation being solved, PICP / PICI / PICC (S/PIC.f:40-297), the strong
 * Dirichlet overwrite SETBCDIR (S/SETBC.f:39-228, std / ustd profiles: the Fortran side still
 * evaluates SETBCDIRL's tmpA/tmpY, :202-232) and the end-of-step copy Ao = An (S/MAIN.f:277-279).
 * Host arrays are (tDof, tnNo) column-major in svFSI node order.  Do may be NULL (no
 * displacement unknowns: fluid, heatS).  Between gpu_pici_ and gpu_picc_ the caller runs
 * gpu_construct_*_dev_, gpu_commu_dev_ and gpu_solve_dev_: no nodal vector crosses PCIe. */
int32_t gpu_pic_init_(const int32_t *tDof, const double *Ao, const double *Yo, const double *Do);
int32_t gpu_pic_free_(void);
int32_t gpu_picp_(const double *gam);                            /* S/PIC.f:82-126   */
int32_t gpu_setbcdir_(const int32_t *faNo, const int32_t *gN, const int32_t *s,
                      const int32_t *lDof, const double *tmpA,
                      const double *tmpY);                       /* S/SETBC.f:118-123; s 1-based */
int32_t gpu_pici_(const double *am, const double *af);           /* S/PIC.f:141-152  */
int32_t gpu_picc_(const double *gam, const double *beta, const double *dt); /* S/PIC.f:203-207 */
int32_t gpu_pic_advance_(void);                                  /* S/MAIN.f:277-279 */
int32_t gpu_pic_get_(const int32_t *which, double *A, double *Y, double *D); /* 0 old, 1 new */

/* ---- face integrals on the device ----------------------------------------
 * gpu_face_create_: one mesh face (faceType: gN, IEN, gE of S/MOD.f) of TRI3 elements on a TET4
 *   mesh; ids 1-based in svFSI's local numbering; call after gpu_mesh_create_.
 * gpu_bassem_neu_fluid_: BASSEMNEUBC + BFLUID + GNNB + DOASSEM (S/EQASSEM.f:90-192,
 *   S/FLUID.f:1279-1336, S/NN.f:1856-1996, S/LHSA.f:266-298) for one Neumann face: adds the
 *   traction residual and the backflow-stabilisation tangent to the device-resident R / Val, reading
 *   the device-resident Yg.  hgN(a) = hg(gN(a)) as SETBCNEUL builds it (S/SETBC.f:292-306).
 * gpu_face_integ_v_: IntegV (S/ALLFUN.f:199-262) of dofs s..s+2 of Yg (which = 0) or of the
 *   time integrator's Yn (which = 1), summed over ranks -- what a resistance BC multiplies by r. */
int32_t gpu_face_create_(const int32_t *iFa, const int32_t *nNo, const int32_t *gN,
                         const int32_t *nEl, const int32_t *eNoN, const int32_t *IEN,
                         const int32_t *gE);
int32_t gpu_face_free_(const int32_t *iFa);
int32_t gpu_bassem_neu_fluid_(const int32_t *iFa, const double *hgN, const double *rho,
                              const double *bfStab, const double *af, const double *gam,
                              const double *dt);
int32_t gpu_face_integ_v_(const int32_t *iFa, const int32_t *which, const int32_t *s, double *flux);

/* algorithmic bytes (nnz*(8 BR BC + 4) + nNo*(8 + 8 BR + 8 BC)) and count of the
 * FSILS_SPARMUL* operations issued since gpu_prof_reset_ while profiling was enabled */
int32_t gpu_prof_spmv_(double *bytes, int64_t *ops);
int32_t gpu_launch_count_(int64_t *n);
/* how the halo sums / all-reduces travel: 0 single rank, 1 NCCL send/recv + all-reduce,
 * 2 peer-memory kernels (CUDA IPC over NVLink), 3 peer-memory with the halo send fused into the
 * SpMV kernel (one launch per FSILS_SPARMUL*; replaces MPI_ISEND/IRECV of L/INCOMMU.f:91-96) */
int32_t gpu_comm_mode_(int32_t *mode);
/* which FSILS_SPARMULVV dof=4 kernel (L/SPARMUL.f:98-113) runs: bit 0 = the unfused path uses 4 lanes per
 * block row with 256-bit loads (default; SVFSI_SPMV_QUAD=0 selects the 8-lane kernel), bit 1 = so does the
 * fused SpMV + halo-send kernel of the multi-GPU path (SVFSI_SPMV_FUSED_QUAD) */
int32_t gpu_spmv_variant_(int32_t *variant);
/* Every in-kernel wait on a peer's flag (the MPI_WAIT of L/INCOMMU.f:98-100 and the MPI_ALLREDUCE of
 * L/DOT.f / L/NORM.f on the peer-memory path) is bounded: after `seconds` (default 120, or
 * SVFSI_COMM_TIMEOUT_S) the kernels give up, the stream drains, and every later call that synchronises
 * returns SVFSI_ERR_COMM until gpu_finalize_ -- a dead or late rank cannot hang the GPUs of the box. */
int32_t gpu_set_comm_timeout_(const double *seconds);
/* cudaStream_t of the library (so a caller can record its own events on it) */
int32_t gpu_get_stream_(void **stream);
int32_t gpu_sync_(void);

#ifdef __cplusplus
}
#endif
#endif /* SVFSI_B200_H */
