/* Plain-C consumer of include/svfsi_b200.h: what a Fortran BIND(C) shim sees.  Checks that the header is valid
 * C (no C++ / torch types), that the flattened FSILS_lsType has the layout the ISO_C_BINDING derived types of
 * INTEGRATION.md assume (a BIND(C) derived type is laid out as the companion C struct), and calls the entry
 * points that need no GPU: gpu_ls_create_ (FSILS_LS_CREATE defaults, L/LS.f:69-95) and svfsi_lhs_plan_
 * (the reordering of FSILS_LHS_CREATE, L/LHS.f:134-211) on a two-rank toy problem.  Prints "ABI OK". */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "svfsi_b200.h"

#define CHECK(c) do { if (!(c)) { printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(void) {
  /* INTEGER(C_INT32_T) x4, REAL(C_DOUBLE) x6 -> 16 + 48 = 64 bytes, no padding */
  CHECK(sizeof(svfsi_subls_t) == 64);
  CHECK(offsetof(svfsi_subls_t, suc) == 0 && offsetof(svfsi_subls_t, mItr) == 4);
  CHECK(offsetof(svfsi_subls_t, sD) == 8 && offsetof(svfsi_subls_t, itr) == 12);
  CHECK(offsetof(svfsi_subls_t, absTol) == 16 && offsetof(svfsi_subls_t, relTol) == 24);
  CHECK(offsetof(svfsi_subls_t, iNorm) == 32 && offsetof(svfsi_subls_t, fNorm) == 40);
  CHECK(offsetof(svfsi_subls_t, dB) == 48 && offsetof(svfsi_subls_t, callD) == 56);
  CHECK(sizeof(svfsi_ls_t) == 16 + 3 * 64);
  CHECK(offsetof(svfsi_ls_t, GM) == 16 && offsetof(svfsi_ls_t, CG) == 80 && offsetof(svfsi_ls_t, RI) == 144);

  svfsi_ls_t ls;
  int32_t t = SVFSI_LS_TYPE_NS;
  CHECK(gpu_ls_create_(&ls, &t) == SVFSI_OK);
  CHECK(ls.LS_type == SVFSI_LS_TYPE_NS && ls.RI.mItr == 10 && ls.GM.mItr == 2 && ls.CG.mItr == 500);
  CHECK(ls.RI.relTol == 0.4 && ls.GM.relTol == 1.e-2 && ls.CG.relTol == 0.2 && ls.GM.sD == 100);
  t = SVFSI_LS_TYPE_GMRES;
  CHECK(gpu_ls_create_(&ls, &t) == SVFSI_OK);
  CHECK(ls.RI.relTol == 0.1 && ls.RI.mItr == 4 && ls.RI.sD == 250 && ls.RI.absTol == 1.e-10);
  t = 123;
  CHECK(gpu_ls_create_(&ls, &t) == SVFSI_ERR_ARG);

  /* two ranks sharing global nodes 3 and 4: rank 0 holds 1..4, rank 1 holds 3..6 */
  int32_t aNodes[2][4] = {{1, 2, 3, 4}, {3, 4, 5, 6}};
  int32_t rank = 0, nranks = 2, gnNo = 6, nNo = 4, maxnNo = 4, map[4], mynNo, shnNo, nReq;
  int32_t iP[2], n[2], ptr[8], cap = 8;
  CHECK(svfsi_lhs_plan_(&rank, &nranks, &gnNo, &nNo, &maxnNo, &aNodes[0][0], map, &mynNo, &shnNo, &nReq, iP, n,
                        ptr, &cap) == SVFSI_OK);
  /* rank 0: nodes shared with the HIGHER rank go to the back (L/LHS.f:139-165): owned = 2 */
  CHECK(mynNo == 2 && shnNo == 0 && nReq == 1 && iP[0] == 2 && n[0] == 2);
  rank = 1;
  CHECK(svfsi_lhs_plan_(&rank, &nranks, &gnNo, &nNo, &maxnNo, &aNodes[0][0], map, &mynNo, &shnNo, &nReq, iP, n,
                        ptr, &cap) == SVFSI_OK);
  /* rank 1: nodes shared with the LOWER rank come first, all four are owned */
  CHECK(mynNo == 4 && shnNo == 2 && nReq == 1 && iP[0] == 1 && n[0] == 2);

  /* every compute entry point refuses to run before gpu_init_ (no CPU fallback) */
  int32_t dof = 4, prec = SVFSI_PRECOND_FSILS;
  CHECK(gpu_solve_dev_(&ls, &dof, &prec, NULL, NULL) != SVFSI_OK);
  CHECK(gpu_sync_() != SVFSI_OK);
  char buf[128];
  int32_t len = 128;
  CHECK(gpu_last_error_(buf, &len) == 0 && strlen(buf) > 0);
  printf("ABI OK\n");
  return 0;
}
