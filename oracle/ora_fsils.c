/*
 * oracle/ora_fsils.c -- TEST INFRASTRUCTURE ONLY (see ora.h header).
 *
 * CPU restatement of the svFSILS linear-solver core: FSILS_LHS_CREATE,
 * FSILS_BC_CREATE, FSILS_COMMUV/S, FSILS_SPARMUL{VV,VS,SV,SS}, DOT/NORM,
 * OMPLA, PRECONDDIAG, ADDBCMUL, GMRES/GMRESS/GMRESV, CGRAD{S,V,_SCHUR},
 * NSSOLVER (+DEPART, BCPRE, GE) and FSILS_SOLVE.
 *
 * MPI is replaced by a "world" of nTasks simulated ranks living in one
 * process: every per-rank array becomes an array of per-rank pointers and
 * every SPMD statement becomes a loop over ranks; collectives (halo sum,
 * allreduce) are done in place between the per-rank arrays.  Control flow in
 * FSILS depends only on all-reduced scalars, so lock-step execution is exact.
 * Pinned to the reference's source text on 1, 2 and 3 MPI tasks (see ora.h).
 */
#include "ora.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define NT (w->nTasks)
#define FOR_RANKS for (int r = 0; r < w->nTasks; r++)

static void *xcalloc(size_t n, size_t s) {
  void *p = calloc(n ? n : 1, s);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}

/* per-rank vectors with `m` doubles per node */
static double **wv_alloc(const ora_world_t *w, int m) {
  double **v = (double **)xcalloc((size_t)NT, sizeof(double *));
  FOR_RANKS v[r] = (double *)xcalloc((size_t)w->lhs[r].nNo*(size_t)m, sizeof(double));
  return v;
}
static double **wnz_alloc(const ora_world_t *w, int m) {
  double **v = (double **)xcalloc((size_t)NT, sizeof(double *));
  FOR_RANKS v[r] = (double *)xcalloc((size_t)w->lhs[r].nnz*(size_t)m, sizeof(double));
  return v;
}
static void wv_free(const ora_world_t *w, double **v) {
  if (!v) return;
  FOR_RANKS free(v[r]);
  free(v);
}

/* ================================================================== */
/* L/LHS.f:51-293  FSILS_LHS_CREATE for every simulated rank           */
ora_world_t *ora_world_create(int nTasks, int gnNo, const int *nNo,
                              const int *nnz, const int *const *gNodes,
                              const int *const *rowPtr,
                              const int *const *colPtr, int nFaces) {
  ora_world_t *w = (ora_world_t *)xcalloc(1, sizeof(*w));
  int maxnNo = 0, tF, i, a, Ac, ai, j, s, e;
  int **ltgNew = NULL;
  w->nTasks = nTasks;
  w->lhs = (ora_lhs_t *)xcalloc((size_t)nTasks, sizeof(ora_lhs_t));

  for (tF = 1; tF <= nTasks; tF++) {
    ora_lhs_t *lhs = &w->lhs[tF-1];
    lhs->gnNo = gnNo; lhs->nNo = nNo[tF-1]; lhs->nnz = nnz[tF-1];
    lhs->nFaces = nFaces;
    lhs->colPtr  = (int *)xcalloc((size_t)lhs->nnz, sizeof(int));
    lhs->rowPtr  = (int *)xcalloc((size_t)lhs->nNo*2, sizeof(int));
    lhs->diagPtr = (int *)xcalloc((size_t)lhs->nNo, sizeof(int));
    lhs->map     = (int *)xcalloc((size_t)lhs->nNo, sizeof(int));
    lhs->face    = (ora_face_t *)xcalloc((size_t)nFaces, sizeof(ora_face_t));
    if (lhs->nNo > maxnNo) maxnNo = lhs->nNo;
  }

  /* sequential case, L/LHS.f:89-111 */
  if (nTasks == 1) {
    ora_lhs_t *lhs = &w->lhs[0];
    for (i = 1; i <= lhs->nnz; i++) lhs->colPtr[i-1] = colPtr[0][i-1];
    for (Ac = 1; Ac <= lhs->nNo; Ac++) {
      s = rowPtr[0][Ac-1];
      e = rowPtr[0][Ac] - 1;
      for (i = s; i <= e; i++) {
        a = colPtr[0][i-1];
        if (Ac == a) { lhs->diagPtr[Ac-1] = i; break; }
      }
      lhs->rowPtr[2*(Ac-1)]   = s;
      lhs->rowPtr[2*(Ac-1)+1] = e;
      lhs->map[Ac-1] = Ac;
    }
    lhs->mynNo = lhs->nNo;
    return w;
  }

  ltgNew = (int **)xcalloc((size_t)nTasks, sizeof(int *));
  /* phase 1 (L/LHS.f:113-211): reorder [shared-with-lower | interior |
   * shared-with-higher] */
  for (tF = 1; tF <= nTasks; tF++) {
    ora_lhs_t *lhs = &w->lhs[tF-1];
    int n = lhs->nNo;
    int *gtlPtr = (int *)xcalloc((size_t)gnNo, sizeof(int));
    int *mine   = (int *)xcalloc((size_t)maxnNo, sizeof(int)); /* aNodes(:,tF) */
    int *ltg    = (int *)xcalloc((size_t)n, sizeof(int));
    memcpy(mine, gNodes[tF-1], sizeof(int)*(size_t)n);
    for (a = 1; a <= n; a++) gtlPtr[gNodes[tF-1][a-1]-1] = a;

    lhs->mynNo = n;
    lhs->shnNo = 0;
    for (i = nTasks; i >= 1; i--) {
      if (i == tF) continue;
      for (a = 1; a <= maxnNo; a++) {
        Ac = (a <= nNo[i-1]) ? gNodes[i-1][a-1] : 0;
        if (Ac == 0) break;
        ai = gtlPtr[Ac-1];
        if (ai != 0) {
          if (mine[ai-1] != 0) {
            if (i < tF) {
              lhs->shnNo = lhs->shnNo + 1;
              ltg[lhs->shnNo-1] = Ac;
            } else {
              ltg[lhs->mynNo-1] = Ac;
              lhs->mynNo = lhs->mynNo - 1;
            }
            mine[ai-1] = 0;
          }
        }
      }
    }
    j = lhs->shnNo + 1;
    for (a = 1; a <= n; a++) {
      Ac = mine[a-1];
      if (Ac != 0) { ltg[j-1] = Ac; j = j + 1; }
    }
    if (j != lhs->mynNo + 1) {
      fprintf(stderr, "FSILS: Unexpected behavior %d %d\n", j, lhs->mynNo);
      abort();
    }
    memset(gtlPtr, 0, sizeof(int)*(size_t)gnNo);
    for (a = 1; a <= n; a++) gtlPtr[ltg[a-1]-1] = a;
    for (a = 1; a <= n; a++) lhs->map[a-1] = gtlPtr[gNodes[tF-1][a-1]-1];

    for (a = 1; a <= n; a++) {
      Ac = lhs->map[a-1];
      lhs->rowPtr[2*(Ac-1)]   = rowPtr[tF-1][a-1];
      lhs->rowPtr[2*(Ac-1)+1] = rowPtr[tF-1][a] - 1;
    }
    for (i = 1; i <= lhs->nnz; i++)
      lhs->colPtr[i-1] = lhs->map[colPtr[tF-1][i-1]-1];
    for (Ac = 1; Ac <= n; Ac++) {
      for (i = lhs->rowPtr[2*(Ac-1)]; i <= lhs->rowPtr[2*(Ac-1)+1]; i++) {
        a = lhs->colPtr[i-1];
        if (Ac == a) { lhs->diagPtr[Ac-1] = i; break; }
      }
    }
    ltgNew[tF-1] = ltg;
    free(gtlPtr); free(mine);
  }

  /* phase 2 (L/LHS.f:213-288): neighbour lists; for a pair (lo,hi) the common
   * list is ordered as in hi's *new* numbering. */
  for (tF = 1; tF <= nTasks; tF++) {
    ora_lhs_t *lhs = &w->lhs[tF-1];
    int n = lhs->nNo, iP;
    int *gtlPtr = (int *)xcalloc((size_t)gnNo, sizeof(int));
    int *disp = (int *)xcalloc((size_t)nTasks, sizeof(int));
    for (a = 1; a <= n; a++) gtlPtr[ltgNew[tF-1][a-1]-1] = a;
    lhs->nReq = 0;
    for (i = 1; i <= nTasks; i++) {
      if (i == tF) continue;
      for (a = 1; a <= nNo[i-1]; a++) {
        Ac = ltgNew[i-1][a-1];
        if (gtlPtr[Ac-1] != 0) disp[i-1]++;
      }
      if (disp[i-1] != 0) lhs->nReq++;
    }
    lhs->cS = (ora_cs_t *)xcalloc((size_t)lhs->nReq, sizeof(ora_cs_t));
    j = 0;
    for (i = 1; i <= nTasks; i++) {
      a = disp[i-1];
      if (a != 0) {
        j++;
        lhs->cS[j-1].iP = i;
        lhs->cS[j-1].n = a;
        lhs->cS[j-1].ptr = (int *)xcalloc((size_t)a, sizeof(int));
      }
    }
    for (i = 1; i <= lhs->nReq; i++) {
      iP = lhs->cS[i-1].iP;
      if (iP < tF) {
        /* MPI_RECV of the list rank iP built by scanning aNodes(:,tF) (my new
         * order) for nodes it also holds -- recomputed here */
        int *held = (int *)xcalloc((size_t)gnNo, sizeof(int));
        for (a = 1; a <= nNo[iP-1]; a++) held[ltgNew[iP-1][a-1]-1] = 1;
        j = 0;
        for (a = 1; a <= n; a++) {
          Ac = ltgNew[tF-1][a-1];
          if (held[Ac-1]) { j++; lhs->cS[i-1].ptr[j-1] = gtlPtr[Ac-1]; }
        }
        free(held);
      } else {
        j = 0;
        for (a = 1; a <= nNo[iP-1]; a++) {
          Ac = ltgNew[iP-1][a-1];
          ai = gtlPtr[Ac-1];
          if (ai != 0) { j++; lhs->cS[i-1].ptr[j-1] = ai; }
        }
      }
    }
    free(gtlPtr); free(disp);
  }
  for (tF = 1; tF <= nTasks; tF++) free(ltgNew[tF-1]);
  free(ltgNew);
  return w;
}

static void ora_face_free(ora_face_t *f) {
  free(f->glob); free(f->val); free(f->valM);
  memset(f, 0, sizeof(*f));
}

void ora_world_free(ora_world_t *w) {
  if (!w) return;
  FOR_RANKS {
    ora_lhs_t *lhs = &w->lhs[r];
    for (int i = 0; i < lhs->nReq; i++) free(lhs->cS[i].ptr);
    for (int i = 0; i < lhs->nFaces; i++) ora_face_free(&lhs->face[i]);
    free(lhs->cS); free(lhs->face); free(lhs->colPtr); free(lhs->rowPtr);
    free(lhs->diagPtr); free(lhs->map);
  }
  free(w->lhs);
  free(w);
}

void ora_world_info(const ora_world_t *w, int rank, int *mynNo, int *shnNo,
                    int *nReq) {
  *mynNo = w->lhs[rank].mynNo; *shnNo = w->lhs[rank].shnNo;
  *nReq = w->lhs[rank].nReq;
}
void ora_world_map(const ora_world_t *w, int rank, int *map) {
  memcpy(map, w->lhs[rank].map, sizeof(int)*(size_t)w->lhs[rank].nNo);
}
void ora_world_rowptr(const ora_world_t *w, int rank, int *rowPtr2, int *colPtr,
                      int *diagPtr) {
  const ora_lhs_t *l = &w->lhs[rank];
  memcpy(rowPtr2, l->rowPtr, sizeof(int)*2*(size_t)l->nNo);
  memcpy(colPtr, l->colPtr, sizeof(int)*(size_t)l->nnz);
  memcpy(diagPtr, l->diagPtr, sizeof(int)*(size_t)l->nNo);
}
int ora_world_cs(const ora_world_t *w, int rank, int i, int *iP, int *n,
                 int *ptr) {
  const ora_cs_t *c = &w->lhs[rank].cS[i];
  *iP = c->iP; *n = c->n;
  if (ptr) memcpy(ptr, c->ptr, sizeof(int)*(size_t)c->n);
  return c->n;
}

/* ================================================================== */
/* L/INCOMMU.f:56-103 FSILS_COMMUV (COMMUS :105-151 is dof = 1).
 * sB is packed on every rank before any rank adds, then each rank adds the
 * neighbours' buffers in ascending neighbour order (:91-96). */
void ora_commuv(const ora_world_t *w, int dof, double *const *R) {
  double ***sB;
  if (NT == 1) return;
  sB = (double ***)xcalloc((size_t)NT, sizeof(double **));
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    sB[r] = (double **)xcalloc((size_t)lhs->nReq, sizeof(double *));
    for (int i = 0; i < lhs->nReq; i++) {
      sB[r][i] = (double *)xcalloc((size_t)lhs->cS[i].n*(size_t)dof, sizeof(double));
      for (int j = 0; j < lhs->cS[i].n; j++) {
        int k = lhs->cS[i].ptr[j];
        for (int d = 0; d < dof; d++)
          sB[r][i][(size_t)j*dof + d] = R[r][(size_t)(k-1)*dof + d];
      }
    }
  }
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    for (int i = 0; i < lhs->nReq; i++) {
      int q = lhs->cS[i].iP - 1, qi = -1;
      const ora_lhs_t *ql = &w->lhs[q];
      for (int t = 0; t < ql->nReq; t++) if (ql->cS[t].iP == r+1) qi = t;
      if (qi < 0 || ql->cS[qi].n != lhs->cS[i].n) {
        fprintf(stderr, "oracle: asymmetric halo schedule\n"); abort();
      }
      for (int j = 0; j < lhs->cS[i].n; j++) {
        int k = lhs->cS[i].ptr[j];
        for (int d = 0; d < dof; d++)
          R[r][(size_t)(k-1)*dof + d] =
              R[r][(size_t)(k-1)*dof + d] + sB[q][qi][(size_t)j*dof + d];
      }
    }
  }
  FOR_RANKS {
    for (int i = 0; i < w->lhs[r].nReq; i++) free(sB[r][i]);
    free(sB[r]);
  }
  free(sB);
}

/* ================================================================== */
/* L/SPARMUL.f:51-133 SPARMULVV; generic-dof loop keeps the left-to-right
 * summation of the unrolled cases (:62-113). */
void ora_sparmul_vv(const ora_world_t *w, int dof, const double *const *K,
                    const double *const *U, double *const *KU) {
  int dd = dof*dof;
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    for (int i = 1; i <= lhs->nNo; i++) {
      double *ku = &KU[r][(size_t)(i-1)*dof];
      for (int l = 0; l < dof; l++) ku[l] = 0.0;
      for (int j = lhs->rowPtr[2*(i-1)]; j <= lhs->rowPtr[2*(i-1)+1]; j++) {
        int col = lhs->colPtr[j-1];
        const double *k = &K[r][(size_t)(j-1)*dd];
        const double *u = &U[r][(size_t)(col-1)*dof];
        for (int l = 0; l < dof; l++) {
          double acc = ku[l];
          for (int m = 0; m < dof; m++) acc = acc + k[l*dof + m]*u[m];
          ku[l] = acc;
        }
      }
    }
  }
  ora_commuv(w, dof, KU);
}

/* L/SPARMUL.f:135-200 SPARMULVS: K(dof,nnz) . U(dof,nNo) -> scalar field */
void ora_sparmul_vs(const ora_world_t *w, int dof, const double *const *K,
                    const double *const *U, double *const *KU) {
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    for (int i = 1; i <= lhs->nNo; i++) {
      double acc = 0.0;
      for (int j = lhs->rowPtr[2*(i-1)]; j <= lhs->rowPtr[2*(i-1)+1]; j++) {
        int col = lhs->colPtr[j-1];
        for (int m = 0; m < dof; m++)
          acc = acc + K[r][(size_t)(j-1)*dof + m]*U[r][(size_t)(col-1)*dof + m];
      }
      KU[r][i-1] = acc;
    }
  }
  ora_commuv(w, 1, KU);
}

/* L/SPARMUL.f:202-271 SPARMULSV: K(dof,nnz) * U(nNo) -> vector field */
void ora_sparmul_sv(const ora_world_t *w, int dof, const double *const *K,
                    const double *const *U, double *const *KU) {
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    for (int i = 1; i <= lhs->nNo; i++) {
      double *ku = &KU[r][(size_t)(i-1)*dof];
      for (int l = 0; l < dof; l++) ku[l] = 0.0;
      for (int j = lhs->rowPtr[2*(i-1)]; j <= lhs->rowPtr[2*(i-1)+1]; j++) {
        int col = lhs->colPtr[j-1];
        for (int l = 0; l < dof; l++)
          ku[l] = ku[l] + K[r][(size_t)(j-1)*dof + l]*U[r][col-1];
      }
    }
  }
  ora_commuv(w, dof, KU);
}

/* L/SPARMUL.f:273-297 SPARMULSS */
void ora_sparmul_ss(const ora_world_t *w, const double *const *K,
                    const double *const *U, double *const *KU) {
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    for (int i = 1; i <= lhs->nNo; i++) {
      double acc = 0.0;
      for (int j = lhs->rowPtr[2*(i-1)]; j <= lhs->rowPtr[2*(i-1)+1]; j++)
        acc = acc + K[r][j-1]*U[r][lhs->colPtr[j-1]-1];
      KU[r][i-1] = acc;
    }
  }
  ora_commuv(w, 1, KU);
}

/* ================================================================== */
/* L/DOT.f:141-191 NCDOTV on one rank (owned nodes 1..mynNo) */
static double ncdot_rank(int dof, int mynNo, const double *U, const double *V) {
  double acc = 0.0;
  for (int i = 0; i < mynNo; i++)
    for (int m = 0; m < dof; m++)
      acc = acc + U[(size_t)i*dof + m]*V[(size_t)i*dof + m];
  return acc;
}
/* L/DOT.f:56-113 DOTV (= NCDOTV + MPI_ALLREDUCE, summed in rank order) */
double ora_dotv(const ora_world_t *w, int dof, const double *const *U,
                const double *const *V) {
  double s = 0.0;
  FOR_RANKS s += ncdot_rank(dof, w->lhs[r].mynNo, U[r], V[r]);
  return s;
}
/* L/NORM.f:55-113 NORMV */
double ora_normv(const ora_world_t *w, int dof, const double *const *U) {
  return sqrt(ora_dotv(w, dof, U, U));
}

/* L/OMPLA.f:67-115 OMPSUMV: U = U + r*V ; :134-182 OMPMULV: U = r*U */
static void ompsum(const ora_world_t *w, int dof, double rr, double *const *U,
                   const double *const *V) {
  FOR_RANKS {
    size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
    for (size_t i = 0; i < n; i++) U[r][i] = U[r][i] + rr*V[r][i];
  }
}
static void ompmul(const ora_world_t *w, int dof, double rr, double *const *U) {
  FOR_RANKS {
    size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
    for (size_t i = 0; i < n; i++) U[r][i] = rr*U[r][i];
  }
}
static void wv_copy(const ora_world_t *w, int dof, double *const *dst,
                    const double *const *src) {
  FOR_RANKS memcpy(dst[r], src[r], sizeof(double)*(size_t)w->lhs[r].nNo*(size_t)dof);
}
static void wv_zero(const ora_world_t *w, int dof, double *const *dst) {
  FOR_RANKS memset(dst[r], 0, sizeof(double)*(size_t)w->lhs[r].nNo*(size_t)dof);
}

/* ================================================================== */
/* L/BC.f:50-121 FSILS_BC_CREATE on every rank. nNo[r], gNodes[r][*] (svFSI
 * local ids, 1-based), Val[r] = val(dof,nNo) or NULL (-> zeros, :91-97). */
void ora_bc_create(ora_world_t *w, int faIn, const int *nNo, int dof,
                   int BC_type, const int *const *gNodes,
                   const double *const *Val) {
  int nHold = 0;
  FOR_RANKS {
    ora_lhs_t *lhs = &w->lhs[r];
    ora_face_t *f;
    if (faIn > lhs->nFaces || faIn <= 0) {
      fprintf(stderr, "FSILS: faIn out of range\n"); abort();
    }
    f = &lhs->face[faIn-1];
    if (f->foC) ora_face_free(f);
    f->foC = 1; /* the reference never sets it (L/BC.f:79-118); harmless */
    f->nNo = nNo[r]; f->dof = dof; f->bGrp = BC_type;
    f->glob = (int *)xcalloc((size_t)f->nNo, sizeof(int));
    f->val  = (double *)xcalloc((size_t)f->nNo*(size_t)dof, sizeof(double));
    f->valM = (double *)xcalloc((size_t)f->nNo*(size_t)dof, sizeof(double));
    for (int a = 0; a < f->nNo; a++) f->glob[a] = lhs->map[gNodes[r][a]-1];
    if (Val && Val[r])
      memcpy(f->val, Val[r], sizeof(double)*(size_t)f->nNo*(size_t)dof);
    if (f->nNo != 0) nHold++;
  }
  if (NT > 1 && nHold > 1) {
    double **v = wv_alloc(w, dof);
    FOR_RANKS {
      ora_face_t *f = &w->lhs[r].face[faIn-1];
      f->sharedFlag = 1;
      for (int a = 0; a < f->nNo; a++)
        for (int d = 0; d < dof; d++)
          v[r][(size_t)(f->glob[a]-1)*dof + d] = f->val[(size_t)a*dof + d];
    }
    ora_commuv(w, dof, v);
    FOR_RANKS {
      ora_face_t *f = &w->lhs[r].face[faIn-1];
      for (int a = 0; a < f->nNo; a++)
        for (int d = 0; d < dof; d++)
          f->val[(size_t)a*dof + d] = v[r][(size_t)(f->glob[a]-1)*dof + d];
    }
    wv_free(w, v);
  }
}

/* L/LS.f:50-119 FSILS_LS_CREATE defaults (:69-95) */
void ora_ls_create(ora_ls_t *ls, int LS_type) {
  memset(ls, 0, sizeof(*ls));
  ls->LS_type = LS_type;
  switch (LS_type) {
  case ORA_LS_TYPE_NS:
    ls->RI.relTol = 0.4; ls->GM.relTol = 1.e-2; ls->CG.relTol = 0.2;
    ls->RI.mItr = 10; ls->GM.mItr = 2; ls->CG.mItr = 500;
    ls->GM.sD = 100; ls->RI.sD = 100;
    break;
  case ORA_LS_TYPE_GMRES:
    ls->RI.relTol = 0.1; ls->RI.mItr = 4; ls->RI.sD = 250;
    break;
  case ORA_LS_TYPE_CG:
    ls->RI.relTol = 1.e-2; ls->RI.mItr = 1000;
    break;
  case ORA_LS_TYPE_BICGS:
    ls->RI.relTol = 1.e-2; ls->RI.mItr = 500;
    break;
  default:
    fprintf(stderr, "FSILS: LS_TYPE is not defined\n"); abort();
  }
  ls->RI.absTol = 1.e-10; ls->GM.absTol = 1.e-10; ls->CG.absTol = 1.e-10;
}

static int any_coupled(const ora_world_t *w) {
  const ora_lhs_t *lhs = &w->lhs[0];
  for (int f = 0; f < lhs->nFaces; f++) if (lhs->face[f].coupledFlag) return 1;
  return 0;
}

/* ================================================================== */
/* L/ADDBCMUL.f:53-114 */
static void addbcmul(const ora_world_t *w, int op_Type, int dof,
                     const double *const *X, double *const *Y) {
  int nFaces = w->lhs[0].nFaces;
  for (int faIn = 0; faIn < nFaces; faIn++) {
    const ora_face_t *f0 = &w->lhs[0].face[faIn];
    if (!f0->coupledFlag) continue;
    if (f0->sharedFlag) {
      double **v = wv_alloc(w, dof);
      double S, coef;
      FOR_RANKS {
        const ora_face_t *f = &w->lhs[r].face[faIn];
        int nsd = f->dof < dof ? f->dof : dof;
        for (int a = 0; a < f->nNo; a++)
          for (int i = 0; i < nsd; i++)
            v[r][(size_t)(f->glob[a]-1)*dof + i] = f->valM[(size_t)a*f->dof + i];
      }
      coef = (op_Type == ORA_BCOP_TYPE_ADD) ? f0->res
           : -f0->res/(1.0 + (f0->res*f0->nS));
      S = coef*ora_dotv(w, dof, (const double *const *)v, X);
      FOR_RANKS {
        const ora_face_t *f = &w->lhs[r].face[faIn];
        int nsd = f->dof < dof ? f->dof : dof;
        for (int a = 0; a < f->nNo; a++) {
          size_t Ac = (size_t)(f->glob[a]-1);
          for (int i = 0; i < nsd; i++)
            Y[r][Ac*dof + i] = Y[r][Ac*dof + i] + v[r][Ac*dof + i]*S;
        }
      }
      wv_free(w, v);
    } else {
      FOR_RANKS {
        const ora_face_t *f = &w->lhs[r].face[faIn];
        int nsd = f->dof < dof ? f->dof : dof;
        double S = 0.0, coef;
        coef = (op_Type == ORA_BCOP_TYPE_ADD) ? f->res
             : -f->res/(1.0 + (f->res*f->nS));
        for (int a = 0; a < f->nNo; a++) {
          size_t Ac = (size_t)(f->glob[a]-1);
          for (int i = 0; i < nsd; i++)
            S = S + f->valM[(size_t)a*f->dof + i]*X[r][Ac*dof + i];
        }
        S = coef*S;
        for (int a = 0; a < f->nNo; a++) {
          size_t Ac = (size_t)(f->glob[a]-1);
          for (int i = 0; i < nsd; i++)
            Y[r][Ac*dof + i] = Y[r][Ac*dof + i] + f->valM[(size_t)a*f->dof + i]*S;
        }
      }
    }
  }
}

/* BCPRE: L/GMRES.f:393-429 and L/NSSOLVER.f:307-341 (nS = ||valM||^2) */
static void bcpre(ora_world_t *w, int nsd) {
  int nFaces = w->lhs[0].nFaces;
  for (int faIn = 0; faIn < nFaces; faIn++) {
    if (!w->lhs[0].face[faIn].coupledFlag) continue;
    if (w->lhs[0].face[faIn].sharedFlag) {
      double **v = wv_alloc(w, nsd);
      double nrm;
      FOR_RANKS {
        const ora_face_t *f = &w->lhs[r].face[faIn];
        for (int a = 0; a < f->nNo; a++)
          for (int i = 0; i < nsd; i++)
            v[r][(size_t)(f->glob[a]-1)*nsd + i] = f->valM[(size_t)a*f->dof + i];
      }
      nrm = ora_normv(w, nsd, (const double *const *)v);
      FOR_RANKS w->lhs[r].face[faIn].nS = nrm*nrm;
      wv_free(w, v);
    } else {
      FOR_RANKS {
        ora_face_t *f = &w->lhs[r].face[faIn];
        f->nS = 0.0;
        for (int a = 0; a < f->nNo; a++)
          for (int i = 0; i < nsd; i++) {
            double t = f->valM[(size_t)a*f->dof + i];
            f->nS = f->nS + t*t;
          }
      }
    }
  }
}

/* ================================================================== */
/* L/PRECOND.f:50-145 PRECONDDIAG with PREMUL :372-428 and POSMUL :430-489 */
static void preconddiag(ora_world_t *w, int dof, double *const *Val,
                        double *const *R, double *const *W) {
  int dd = dof*dof;
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    for (int Ac = 1; Ac <= lhs->nNo; Ac++) {
      int d = lhs->diagPtr[Ac-1];
      for (int i = 1; i <= dof; i++)
        W[r][(size_t)(Ac-1)*dof + i-1] = Val[r][(size_t)(d-1)*dd + (i*dof-dof+i) - 1];
    }
  }
  ora_commuv(w, dof, W);
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    size_t n = (size_t)lhs->nNo*(size_t)dof;
    for (size_t i = 0; i < n; i++) if (W[r][i] == 0.0) W[r][i] = 1.0;
    for (size_t i = 0; i < n; i++) W[r][i] = 1.0/sqrt(fabs(W[r][i]));
    for (int faIn = 0; faIn < lhs->nFaces; faIn++) {
      const ora_face_t *f = &lhs->face[faIn];
      int i;
      if (!f->incFlag) continue;
      i = f->dof < dof ? f->dof : dof;
      if (f->bGrp == ORA_BC_TYPE_DIR) {
        for (int a = 0; a < f->nNo; a++) {
          size_t Ac = (size_t)(f->glob[a]-1);
          for (int k = 0; k < i; k++)
            W[r][Ac*dof + k] = W[r][Ac*dof + k]*f->val[(size_t)a*f->dof + k];
        }
      }
    }
    /* PREMUL: Val = W*Val (row scaling) */
    for (int Ac = 1; Ac <= lhs->nNo; Ac++)
      for (int j = lhs->rowPtr[2*(Ac-1)]; j <= lhs->rowPtr[2*(Ac-1)+1]; j++)
        for (int i = 0; i < dof; i++)
          for (int k = 0; k < dof; k++)
            Val[r][(size_t)(j-1)*dd + i*dof + k] =
                Val[r][(size_t)(j-1)*dd + i*dof + k]*W[r][(size_t)(Ac-1)*dof + i];
    /* R = W*R */
    for (size_t i = 0; i < n; i++) R[r][i] = W[r][i]*R[r][i];
    /* POSMUL: Val = Val*W (column scaling) */
    for (int Ac = 1; Ac <= lhs->nNo; Ac++)
      for (int j = lhs->rowPtr[2*(Ac-1)]; j <= lhs->rowPtr[2*(Ac-1)+1]; j++) {
        int a = lhs->colPtr[j-1];
        for (int i = 0; i < dof; i++)
          for (int k = 0; k < dof; k++)
            Val[r][(size_t)(j-1)*dd + i*dof + k] =
                Val[r][(size_t)(j-1)*dd + i*dof + k]*W[r][(size_t)(a-1)*dof + k];
      }
    for (int faIn = 0; faIn < lhs->nFaces; faIn++) {
      ora_face_t *f = &lhs->face[faIn];
      if (f->coupledFlag) {
        int m = f->dof < dof ? f->dof : dof;
        for (int a = 0; a < f->nNo; a++) {
          size_t Ac = (size_t)(f->glob[a]-1);
          for (int i = 0; i < m; i++)
            f->valM[(size_t)a*f->dof + i] = f->val[(size_t)a*f->dof + i]*W[r][Ac*dof + i];
        }
      }
    }
  }
}

/* ================================================================== */
/* L/PRECOND.f:150-368 PRECONDRCS: Dirichlet rows/columns killed with a 0/1 mask,
 * unit diagonal put back, then iterated max-norm row/column equilibration
 * (at most 10 sweeps, stop when every row/column max is < 3).  W1 scales the RHS,
 * W2 (the Wc of L/SOLVE.f:136) the solution.  Coupled-face valM is NOT updated
 * (the block is commented out in the reference, :352-364). */
static void premul(const ora_lhs_t *lhs, int dof, double *Val, const double *W) {
  int dd = dof*dof;
  for (int Ac = 1; Ac <= lhs->nNo; Ac++)
    for (int j = lhs->rowPtr[2*(Ac-1)]; j <= lhs->rowPtr[2*(Ac-1)+1]; j++)
      for (int i = 0; i < dof; i++)
        for (int k = 0; k < dof; k++)
          Val[(size_t)(j-1)*dd + i*dof + k] =
              Val[(size_t)(j-1)*dd + i*dof + k]*W[(size_t)(Ac-1)*dof + i];
}
static void posmul(const ora_lhs_t *lhs, int dof, double *Val, const double *W) {
  int dd = dof*dof;
  for (int Ac = 1; Ac <= lhs->nNo; Ac++)
    for (int j = lhs->rowPtr[2*(Ac-1)]; j <= lhs->rowPtr[2*(Ac-1)+1]; j++) {
      int a = lhs->colPtr[j-1];
      for (int i = 0; i < dof; i++)
        for (int k = 0; k < dof; k++)
          Val[(size_t)(j-1)*dd + i*dof + k] =
              Val[(size_t)(j-1)*dd + i*dof + k]*W[(size_t)(a-1)*dof + k];
    }
}
static void precondrcs(ora_world_t *w, int dof, double *const *Val,
                       double *const *R, double *const *W1, double *const *W2) {
  int dd = dof*dof, iter = 0, maxiter = 10, flag = 1;
  double tol = 2.0;
  double **Wr = wv_alloc(w, dof), **Wc = wv_alloc(w, dof);

  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    size_t n = (size_t)lhs->nNo*(size_t)dof;
    for (size_t i = 0; i < n; i++) { W1[r][i] = 1.0; W2[r][i] = 1.0; Wr[r][i] = 1.0; }
    for (int faIn = 0; faIn < lhs->nFaces; faIn++) {   /* :181-191 */
      const ora_face_t *f = &lhs->face[faIn];
      int m;
      if (!f->incFlag) continue;
      m = f->dof < dof ? f->dof : dof;
      if (f->bGrp == ORA_BC_TYPE_DIR)
        for (int a = 0; a < f->nNo; a++) {
          size_t Ac = (size_t)(f->glob[a]-1);
          for (int k = 0; k < m; k++)
            Wr[r][Ac*dof + k] = Wr[r][Ac*dof + k]*f->val[(size_t)a*f->dof + k];
        }
    }
  }
  ora_commuv(w, dof, Wr);                               /* :192 */
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    size_t n = (size_t)lhs->nNo*(size_t)dof;
    for (size_t i = 0; i < n; i++) {                    /* :195-197 */
      double v = Wr[r][i] - 0.5;
      v = v/fabs(v);
      Wr[r][i] = (v + fabs(v))*0.5;
    }
    premul(lhs, dof, Val[r], Wr[r]);                    /* :199 */
    for (size_t i = 0; i < n; i++) R[r][i] = Wr[r][i]*R[r][i];
    posmul(lhs, dof, Val[r], Wr[r]);                    /* :201 */
    for (int Ac = 1; Ac <= lhs->nNo; Ac++) {            /* :203-237 */
      int d = lhs->diagPtr[Ac-1];
      for (int i = 1; i <= dof; i++) {
        size_t q = (size_t)(d-1)*dd + (size_t)(i*dof-dof+i) - 1;
        Val[r][q] = Wr[r][(size_t)(Ac-1)*dof + i-1]*(Val[r][q] - 1.0) + 1.0;
      }
    }
  }

  while (flag) {                                        /* :242-344 */
    int conv = 1;
    iter = iter + 1;
    if (iter >= maxiter) flag = 0;
    FOR_RANKS {
      const ora_lhs_t *lhs = &w->lhs[r];
      size_t n = (size_t)lhs->nNo*(size_t)dof;
      for (size_t i = 0; i < n; i++) { Wr[r][i] = 0.0; Wc[r][i] = 0.0; }
      for (int Ac = 1; Ac <= lhs->nNo; Ac++)
        for (int j = lhs->rowPtr[2*(Ac-1)]; j <= lhs->rowPtr[2*(Ac-1)+1]; j++) {
          int a = lhs->colPtr[j-1];
          for (int i = 0; i < dof; i++)
            for (int k = 0; k < dof; k++) {
              double v = fabs(Val[r][(size_t)(j-1)*dd + i*dof + k]);
              if (v > Wr[r][(size_t)(Ac-1)*dof + i]) Wr[r][(size_t)(Ac-1)*dof + i] = v;
              if (v > Wc[r][(size_t)(a-1)*dof + k]) Wc[r][(size_t)(a-1)*dof + k] = v;
            }
        }
    }
    ora_commuv(w, dof, Wr);                             /* :320-321 */
    ora_commuv(w, dof, Wc);
    /* :323-324 -- every rank tests its own vectors and clears ITS flag; the
     * MPI_ALLGATHER + ANY (:338-342) keeps all ranks sweeping while any rank is
     * unconverged (and not at maxiter, which all ranks reach together) */
    FOR_RANKS {
      size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
      double mr = 0.0, mc = 0.0;
      for (size_t i = 0; i < n; i++) {
        double a = fabs(1.0 - Wr[r][i]), b = fabs(1.0 - Wc[r][i]);
        if (a > mr) mr = a;
        if (b > mc) mc = b;
      }
      if (!(mr < tol && mc < tol)) conv = 0;
    }
    if (conv) flag = 0;
    FOR_RANKS {
      const ora_lhs_t *lhs = &w->lhs[r];
      size_t n = (size_t)lhs->nNo*(size_t)dof;
      for (size_t i = 0; i < n; i++) {                  /* :326-327 */
        Wr[r][i] = 1.0/sqrt(Wr[r][i]);
        Wc[r][i] = 1.0/sqrt(Wc[r][i]);
      }
      premul(lhs, dof, Val[r], Wr[r]);                  /* :329-330 */
      posmul(lhs, dof, Val[r], Wc[r]);
      for (size_t i = 0; i < n; i++) {                  /* :332-333 */
        W1[r][i] = W1[r][i]*Wr[r][i];
        W2[r][i] = W2[r][i]*Wc[r][i];
      }
    }
  }
  FOR_RANKS {                                           /* :347 */
    size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
    for (size_t i = 0; i < n; i++) R[r][i] = W1[r][i]*R[r][i];
  }
  wv_free(w, Wr); wv_free(w, Wc);
}

/* ================================================================== */
/* Arnoldi/Givens core shared by GMRES (L/GMRES.f:106-152), GMRESS (:212-263)
 * and GMRESV (:330-381).  u[k] are world vectors, k = 0..sD.  pre != 0 applies
 * the BCOP_TYPE_PRE correction (only GMRES(...,X) does).  Returns i as left by
 * the Fortran DO loop after `IF (i .GT. sD) i = sD`. */
typedef struct {
  int sD; double *h, *c, *s, *err, *y; /* h(sD+1,sD) column-major */
} ora_hess_t;
#define H(i, j) hs->h[((size_t)(j)-1)*(size_t)(hs->sD+1) + (i)-1]

static int arnoldi_cycle(ora_world_t *w, ora_subls_t *ls, int dof,
                         const double *const *Val, double ***u,
                         double **unCondU, ora_hess_t *hs, double eps, int pre,
                         int scalar) {
  int i, j, k;
  double tmp;
  for (i = 1; i <= ls->sD; i++) {
    ls->itr = ls->itr + 1;
    if (scalar)
      ora_sparmul_ss(w, Val, (const double *const *)u[i-1], u[i]);
    else {
      ora_sparmul_vv(w, dof, Val, (const double *const *)u[i-1], u[i]);
      addbcmul(w, ORA_BCOP_TYPE_ADD, dof, (const double *const *)u[i-1], u[i]);
      if (pre && any_coupled(w)) {
        wv_copy(w, dof, unCondU, (const double *const *)u[i]);
        addbcmul(w, ORA_BCOP_TYPE_PRE, dof, (const double *const *)unCondU, u[i]);
      }
    }
    for (j = 1; j <= i+1; j++)  /* NCDOT + BCASTV */
      H(j,i) = ora_dotv(w, dof, (const double *const *)u[j-1],
                        (const double *const *)u[i]);
    for (j = 1; j <= i; j++) {
      ompsum(w, dof, -H(j,i), u[i], (const double *const *)u[j-1]);
      H(i+1,i) = H(i+1,i) - H(j,i)*H(j,i);
    }
    H(i+1,i) = sqrt(fabs(H(i+1,i)));

    ompmul(w, dof, 1.0/H(i+1,i), u[i]);
    for (j = 1; j <= i-1; j++) {
      tmp      =  hs->c[j-1]*H(j,i) + hs->s[j-1]*H(j+1,i);
      H(j+1,i) = -hs->s[j-1]*H(j,i) + hs->c[j-1]*H(j+1,i);
      H(j,i)   =  tmp;
    }
    tmp        = sqrt(H(i,i)*H(i,i) + H(i+1,i)*H(i+1,i));
    hs->c[i-1] = H(i,i)/tmp;
    hs->s[i-1] = H(i+1,i)/tmp;
    H(i,i)     = tmp;
    H(i+1,i)   = 0.0;
    hs->err[i]   = -hs->s[i-1]*hs->err[i-1];
    hs->err[i-1] =  hs->c[i-1]*hs->err[i-1];
    if (fabs(hs->err[i]) < eps) { ls->suc = 1; break; }
  }
  if (i > ls->sD) i = ls->sD;

  for (j = 0; j < i; j++) hs->y[j] = hs->err[j];
  for (j = i; j >= 1; j--) {
    for (k = j+1; k <= i; k++) hs->y[j-1] = hs->y[j-1] - H(j,k)*hs->y[k-1];
    hs->y[j-1] = hs->y[j-1]/H(j,j);
  }
  return i;
}

static void hess_alloc(ora_hess_t *hs, int sD) {
  hs->sD = sD;
  hs->h = (double *)xcalloc((size_t)(sD+1)*(size_t)sD, sizeof(double));
  hs->c = (double *)xcalloc((size_t)sD, sizeof(double));
  hs->s = (double *)xcalloc((size_t)sD, sizeof(double));
  hs->y = (double *)xcalloc((size_t)sD, sizeof(double));
  hs->err = (double *)xcalloc((size_t)sD+1, sizeof(double));
}
static void hess_free(ora_hess_t *hs) {
  free(hs->h); free(hs->c); free(hs->s); free(hs->y); free(hs->err);
}

/* L/GMRES.f:273-431 GMRESV (scalar = 0) and :171-271 GMRESS (scalar = 1) */
static void gmres_inplace(ora_world_t *w, ora_subls_t *ls, int dof,
                          const double *const *Val, double *const *R,
                          int scalar) {
  ora_hess_t hsv, *hs = &hsv;
  double ***u, **X, eps, t0;
  int i, j, l;

  hess_alloc(hs, ls->sD);
  u = (double ***)xcalloc((size_t)ls->sD+1, sizeof(double **));
  for (i = 0; i <= ls->sD; i++) u[i] = wv_alloc(w, dof);
  X = wv_alloc(w, dof);

  t0        = ora_wtime();
  ls->suc   = 0;
  eps       = ora_normv(w, dof, (const double *const *)R);
  ls->iNorm = eps;
  ls->fNorm = eps;
  eps       = (ls->absTol > ls->relTol*eps) ? ls->absTol : ls->relTol*eps;
  ls->itr   = 0;

  if (!scalar) bcpre(w, dof-1);

  if (ls->iNorm <= ls->absTol) {
    ls->callD = DBL_EPSILON;
    ls->dB    = 0.0;
    goto done;   /* R is returned untouched (L/GMRES.f:307-311) */
  }

  for (l = 1; l <= ls->mItr; l++) {
    ls->dB = ls->fNorm;
    ls->itr = ls->itr + 1;
    if (scalar)
      ora_sparmul_ss(w, Val, (const double *const *)X, u[0]);
    else {
      ora_sparmul_vv(w, dof, Val, (const double *const *)X, u[0]);
      addbcmul(w, ORA_BCOP_TYPE_ADD, dof, (const double *const *)X, u[0]);
    }
    FOR_RANKS {
      size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
      for (size_t k = 0; k < n; k++) u[0][r][k] = R[r][k] - u[0][r][k];
    }
    hs->err[0] = ora_normv(w, dof, (const double *const *)u[0]);
    FOR_RANKS {
      size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
      for (size_t k = 0; k < n; k++) u[0][r][k] = u[0][r][k]/hs->err[0];
    }
    i = arnoldi_cycle(w, ls, dof, Val, u, NULL, hs, eps, 0, scalar);
    for (j = 1; j <= i; j++)
      ompsum(w, dof, hs->y[j-1], X, (const double *const *)u[j-1]);
    ls->fNorm = fabs(hs->err[i]);
    if (ls->suc) break;
  }
  wv_copy(w, dof, R, (const double *const *)X);
  ls->callD = ora_wtime() - t0;
  ls->dB    = 10.0*log(ls->fNorm/ls->dB);
done:
  for (i = 0; i <= ls->sD; i++) wv_free(w, u[i]);
  free(u);
  wv_free(w, X);
  hess_free(hs);
}

/* L/GMRES.f:51-169 GMRES(lhs, ls, dof, Val, R, X): out of place, accumulates
 * ls%itr / callD, applies the PRE correction when any face is coupled. */
static void gmres_outofplace(ora_world_t *w, ora_subls_t *ls, int dof,
                             const double *const *Val, const double *const *R,
                             double *const *X) {
  ora_hess_t hsv, *hs = &hsv;
  double ***u, **unCondU, eps, t0;
  int i, j, l;

  hess_alloc(hs, ls->sD);
  u = (double ***)xcalloc((size_t)ls->sD+1, sizeof(double **));
  for (i = 0; i <= ls->sD; i++) u[i] = wv_alloc(w, dof);
  unCondU = wv_alloc(w, dof);

  t0 = ora_wtime();
  ls->suc = 0;
  eps = 0.0;
  wv_zero(w, dof, X);
  for (l = 1; l <= ls->mItr; l++) {
    if (l == 1) {
      wv_copy(w, dof, u[0], R);
    } else {
      ls->itr = ls->itr + 1;
      ora_sparmul_vv(w, dof, Val, (const double *const *)X, u[0]);
      addbcmul(w, ORA_BCOP_TYPE_ADD, dof, (const double *const *)X, u[0]);
      FOR_RANKS {
        size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
        for (size_t k = 0; k < n; k++) u[0][r][k] = R[r][k] - u[0][r][k];
      }
    }
    if (any_coupled(w)) {
      wv_copy(w, dof, unCondU, (const double *const *)u[0]);
      addbcmul(w, ORA_BCOP_TYPE_PRE, dof, (const double *const *)unCondU, u[0]);
    }
    hs->err[0] = ora_normv(w, dof, (const double *const *)u[0]);
    if (l == 1) {
      eps = hs->err[0];
      if (eps <= ls->absTol) {
        ls->callD = DBL_EPSILON;
        ls->dB    = 0.0;
        goto done;
      }
      ls->iNorm = eps;
      ls->fNorm = eps;
      eps = (ls->absTol > ls->relTol*eps) ? ls->absTol : ls->relTol*eps;
    }
    ls->dB = ls->fNorm;
    FOR_RANKS {
      size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
      for (size_t k = 0; k < n; k++) u[0][r][k] = u[0][r][k]/hs->err[0];
    }
    i = arnoldi_cycle(w, ls, dof, Val, u, unCondU, hs, eps, 1, 0);
    for (j = 1; j <= i; j++)
      ompsum(w, dof, hs->y[j-1], X, (const double *const *)u[j-1]);
    ls->fNorm = fabs(hs->err[i]);
    if (ls->suc) break;
  }
  ls->callD = ora_wtime() - t0 + ls->callD;
  ls->dB    = 10.0*log(ls->fNorm/ls->dB);
done:
  for (i = 0; i <= ls->sD; i++) wv_free(w, u[i]);
  free(u);
  wv_free(w, unCondU);
  hess_free(hs);
}

/* ================================================================== */
/* L/CGRAD.f:125-182 CGRADS (dof = 1) and :184-242 CGRADV */
static void cgrad(ora_world_t *w, ora_subls_t *ls, int dof,
                  const double *const *K, double *const *R) {
  double **P = wv_alloc(w, dof), **KP = wv_alloc(w, dof), **X = wv_alloc(w, dof);
  double errO, err, alpha, eps, t0;
  int i;

  t0        = ora_wtime();
  ls->suc   = 0;
  ls->iNorm = ora_normv(w, dof, (const double *const *)R);
  eps       = (ls->absTol > ls->relTol*ls->iNorm) ? ls->absTol : ls->relTol*ls->iNorm;
  eps       = eps*eps;
  errO      = ls->iNorm*ls->iNorm;
  err       = errO;
  wv_copy(w, dof, P, (const double *const *)R);

  for (i = 1; i <= ls->mItr; i++) {
    if (err < eps) { ls->suc = 1; break; }
    errO = err;
    if (dof == 1)
      ora_sparmul_ss(w, K, (const double *const *)P, KP);
    else
      ora_sparmul_vv(w, dof, K, (const double *const *)P, KP);
    alpha = errO/ora_dotv(w, dof, (const double *const *)P, (const double *const *)KP);
    ompsum(w, dof, alpha, X, (const double *const *)P);
    ompsum(w, dof, -alpha, R, (const double *const *)KP);
    err = ora_normv(w, dof, (const double *const *)R);
    err = err*err;
    ompsum(w, dof, errO/err, P, (const double *const *)R);
    ompmul(w, dof, err/errO, P);
  }
  wv_copy(w, dof, R, (const double *const *)X);
  ls->itr   = i - 1;
  ls->fNorm = sqrt(err);
  ls->callD = ora_wtime() - t0;
  if (errO < DBL_EPSILON) ls->dB = 0.0;
  else ls->dB = 5.0*log(err/errO);
  wv_free(w, P); wv_free(w, KP); wv_free(w, X);
}

/* L/BICGS.f:50-111 BICGSS (dof = 1) and :113-180 BICGSV: BiCGStab, statement order kept */
static void bicgs(ora_world_t *w, ora_subls_t *ls, int dof,
                  const double *const *K, double *const *R) {
  double **P = wv_alloc(w, dof), **Rh = wv_alloc(w, dof), **X = wv_alloc(w, dof),
         **V = wv_alloc(w, dof), **S = wv_alloc(w, dof), **T = wv_alloc(w, dof);
  double errO, err, alpha, beta, rho, rhoO, omega, eps, t0;
  int i;
#define CW(x) ((const double *const *)(x))
  t0        = ora_wtime();
  ls->suc   = 0;
  err       = ora_normv(w, dof, CW(R));
  errO      = err;
  ls->iNorm = err;
  eps       = (ls->absTol > ls->relTol*err) ? ls->absTol : ls->relTol*err;
  rho       = err*err;
  wv_zero(w, dof, X);
  wv_copy(w, dof, P, CW(R));
  wv_copy(w, dof, Rh, CW(R));

  for (i = 1; i <= ls->mItr; i++) {
    if (err < eps) { ls->suc = 1; break; }
    if (dof == 1) ora_sparmul_ss(w, K, CW(P), V);
    else ora_sparmul_vv(w, dof, K, CW(P), V);
    alpha = rho/ora_dotv(w, dof, CW(Rh), CW(V));
    FOR_RANKS {                                   /* S = R - alpha*V */
      size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
      for (size_t e = 0; e < n; e++) S[r][e] = R[r][e] - alpha*V[r][e];
    }
    if (dof == 1) ora_sparmul_ss(w, K, CW(S), T);
    else ora_sparmul_vv(w, dof, K, CW(S), T);
    omega = ora_normv(w, dof, CW(T));
    omega = ora_dotv(w, dof, CW(T), CW(S))/(omega*omega);
    FOR_RANKS {
      size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
      for (size_t e = 0; e < n; e++) {
        X[r][e] = X[r][e] + alpha*P[r][e] + omega*S[r][e];
        R[r][e] = S[r][e] - omega*T[r][e];
      }
    }
    errO = err;
    err  = ora_normv(w, dof, CW(R));
    rhoO = rho;
    rho  = ora_dotv(w, dof, CW(R), CW(Rh));
    beta = rho*alpha/(rhoO*omega);
    FOR_RANKS {
      size_t n = (size_t)w->lhs[r].nNo*(size_t)dof;
      for (size_t e = 0; e < n; e++)
        P[r][e] = R[r][e] + beta*(P[r][e] - omega*V[r][e]);
    }
  }
  wv_copy(w, dof, R, CW(X));
  ls->itr   = i - 1;
  ls->fNorm = err;
  ls->callD = ora_wtime() - t0;
  if (errO < DBL_EPSILON) ls->dB = 0.0;
  else ls->dB = 10.0*log(err/errO);
#undef CW
  wv_free(w, P); wv_free(w, Rh); wv_free(w, X); wv_free(w, V); wv_free(w, S); wv_free(w, T);
}

/* ================================================================== */
/* L/CGRAD.f:51-123 CGRAD_SCHUR: operator L.p - D.(G.p) (+ PRE correction) */
static void cgrad_schur(ora_world_t *w, ora_subls_t *ls, int dof,
                        const double *const *D, const double *const *G,
                        const double *const *L, double *const *R) {
  double **X = wv_alloc(w, 1), **P = wv_alloc(w, 1), **SP = wv_alloc(w, 1),
         **DGP = wv_alloc(w, 1), **GP = wv_alloc(w, dof),
         **unCondU = wv_alloc(w, dof);
  double errO, err, alpha, eps, t0;
  int i;

  t0        = ora_wtime();
  ls->suc   = 0;
  ls->iNorm = ora_normv(w, 1, (const double *const *)R);
  eps       = (ls->absTol > ls->relTol*ls->iNorm) ? ls->absTol : ls->relTol*ls->iNorm;
  eps       = eps*eps;
  errO      = ls->iNorm*ls->iNorm;
  err       = errO;
  wv_copy(w, 1, P, (const double *const *)R);

  for (i = 1; i <= ls->mItr; i++) {
    if (err < eps) { ls->suc = 1; break; }
    errO = err;
    ora_sparmul_sv(w, dof, G, (const double *const *)P, GP);
    if (any_coupled(w)) {
      wv_copy(w, dof, unCondU, (const double *const *)GP);
      addbcmul(w, ORA_BCOP_TYPE_PRE, dof, (const double *const *)unCondU, GP);
    }
    ora_sparmul_vs(w, dof, D, (const double *const *)GP, DGP);
    ora_sparmul_ss(w, L, (const double *const *)P, SP);

    ompsum(w, 1, -1.0, SP, (const double *const *)DGP);
    alpha = errO/ora_dotv(w, 1, (const double *const *)P, (const double *const *)SP);
    ompsum(w, 1, alpha, X, (const double *const *)P);
    ompsum(w, 1, -alpha, R, (const double *const *)SP);
    err = ora_normv(w, 1, (const double *const *)R);
    err = err*err;
    ompsum(w, 1, errO/err, P, (const double *const *)R);
    ompmul(w, 1, err/errO, P);
  }
  wv_copy(w, 1, R, (const double *const *)X);
  ls->fNorm = sqrt(err);
  ls->callD = ora_wtime() - t0 + ls->callD;
  ls->itr   = ls->itr + i - 1;
  if (errO < DBL_EPSILON) ls->dB = 0.0;
  else ls->dB = 5.0*log(err/errO);
  wv_free(w, X); wv_free(w, P); wv_free(w, SP); wv_free(w, DGP);
  wv_free(w, GP); wv_free(w, unCondU);
}

/* ================================================================== */
/* L/GE.f:51-151.  A(nV,N) column-major, B(N) in/out.  Returns 1 = success. */
static int ge(int nV, int N, const double *A, double *B) {
#define AA(i, j) A[((size_t)(j)-1)*(size_t)nV + (i)-1]
#define C(i, j) Cm[((size_t)(j)-1)*(size_t)N + (i)-1]
  double *W, *Cm, pivot, saveEl;
  int m, ipv, i, j;
  if (N <= 0) return 0;
  W = (double *)xcalloc((size_t)N, sizeof(double));
  for (i = 1; i <= N; i++) {
    if (fabs(AA(i,i)) < DBL_MIN) {
      for (j = 0; j < N; j++) B[j] = 0.0;
      free(W);
      return 0;
    }
    W[i-1] = 1.0/sqrt(fabs(AA(i,i)));
  }
  Cm = (double *)xcalloc((size_t)N*(size_t)(N+1), sizeof(double));
  for (i = 1; i <= N; i++) {
    for (j = 1; j <= N; j++) C(i,j) = W[i-1]*W[j-1]*AA(i,j);
    C(i,N+1) = W[i-1]*B[i-1];
  }
  if (N == 1) {
    B[0] = C(1,2)/C(1,1);
    B[0] = B[0]*W[0];
    free(W); free(Cm);
    return 1;
  } else if (N == 2) {
    pivot = C(1,1)*C(2,2) - C(2,1)*C(1,2);
    if (fabs(pivot) < DBL_EPSILON) {
      B[0] = B[1] = 0.0;
      free(W); free(Cm);
      return 0;
    }
    B[0] = (C(1,3)*C(2,2) - C(2,3)*C(1,2))/pivot;
    B[1] = (C(2,3)*C(1,1) - C(1,3)*C(2,1))/pivot;
    B[0] = W[0]*B[0];
    B[1] = W[1]*B[1];
    free(W); free(Cm);
    return 1;
  }
  for (m = 1; m <= N-1; m++) {
    ipv = m;
    pivot = fabs(C(m,m));
    for (i = m+1; i <= N; i++) {
      if (fabs(C(i,m)) > pivot) { ipv = i; pivot = fabs(C(i,m)); }
    }
    if (fabs(pivot) < DBL_EPSILON) {
      for (j = 0; j < N; j++) B[j] = 0.0;
      free(W); free(Cm);
      return 0;
    }
    if (ipv != m) {
      for (j = m; j <= N+1; j++) {
        saveEl = C(m,j); C(m,j) = C(ipv,j); C(ipv,j) = saveEl;
      }
    }
    for (i = m+1; i <= N; i++) {
      saveEl = C(i,m)/C(m,m);
      C(i,m) = 0.0;
      for (j = m+1; j <= N+1; j++) C(i,j) = C(i,j) - saveEl*C(m,j);
    }
  }
  for (j = N; j >= 1; j--) {
    for (i = j+1; i <= N; i++) C(j,N+1) = C(j,N+1) - C(j,i)*C(i,N+1);
    C(j,N+1) = C(j,N+1)/C(j,j);
  }
  for (i = 1; i <= N; i++) B[i-1] = W[i-1]*C(i,N+1);
  free(W); free(Cm);
  return 1;
#undef AA
#undef C
}

/* test hook for GE */
int ora_ge(int nV, int N, const double *A, double *B) { return ge(nV, N, A, B); }

/* ================================================================== */
/* L/NSSOLVER.f:52-233 NSSOLVER with DEPART :237-305 (nsd = 3 or 2) */
static void nssolver(ora_world_t *w, ora_ls_t *ls, int dof,
                     const double *const *Val, double *const *Ri) {
  int nsd = dof - 1, dd = dof*dof, nn = nsd*nsd;
  int i, j, k, iB, iBB, nB, c, mItr = ls->RI.mItr, last_i;
  double eps, t0, *tmp, *A, *B, *xB, *oldxB;
  double **Rm = wv_alloc(w, nsd), **Rc = wv_alloc(w, 1), **Rmi = wv_alloc(w, nsd),
         **Rci = wv_alloc(w, 1);
  double ***U, ***P, ***MU, ***MP;
  double **mK = wnz_alloc(w, nn), **mG = wnz_alloc(w, nsd), **mD = wnz_alloc(w, nsd),
         **mL = wnz_alloc(w, 1), **Gt = wnz_alloc(w, nsd);

  iB = mItr;
  nB = 2*iB;
  U  = (double ***)xcalloc((size_t)iB, sizeof(double **));
  P  = (double ***)xcalloc((size_t)iB, sizeof(double **));
  MU = (double ***)xcalloc((size_t)nB, sizeof(double **));
  MP = (double ***)xcalloc((size_t)nB, sizeof(double **));
  for (i = 0; i < iB; i++) { U[i] = wv_alloc(w, nsd); P[i] = wv_alloc(w, 1); }
  for (i = 0; i < nB; i++) { MU[i] = wv_alloc(w, nsd); MP[i] = wv_alloc(w, 1); }
  tmp   = (double *)xcalloc((size_t)nB*nB + nB, sizeof(double));
  A     = (double *)xcalloc((size_t)nB*nB, sizeof(double));
  B     = (double *)xcalloc((size_t)nB, sizeof(double));
  xB    = (double *)xcalloc((size_t)nB, sizeof(double));
  oldxB = (double *)xcalloc((size_t)nB, sizeof(double));
#define AM(i, j) A[((size_t)(j)-1)*(size_t)nB + (i)-1]

  FOR_RANKS {
    for (int a = 0; a < w->lhs[r].nNo; a++) {
      for (int d = 0; d < nsd; d++) Rmi[r][(size_t)a*nsd + d] = Ri[r][(size_t)a*dof + d];
      Rci[r][a] = Ri[r][(size_t)a*dof + dof-1];
    }
  }
  wv_copy(w, nsd, Rm, (const double *const *)Rmi);
  wv_copy(w, 1, Rc, (const double *const *)Rci);
  {
    double nm = ora_normv(w, nsd, (const double *const *)Rm);
    double nc = ora_normv(w, 1, (const double *const *)Rc);
    eps = sqrt(nm*nm + nc*nc);
  }
  ls->RI.iNorm = eps;
  ls->RI.fNorm = eps*eps;
  ls->CG.callD = 0.0;
  ls->GM.callD = 0.0;
  ls->CG.itr   = 0;
  ls->GM.itr   = 0;
  t0           = ora_wtime();
  ls->RI.suc   = 0;
  eps          = (ls->RI.absTol > ls->RI.relTol*eps) ? ls->RI.absTol : ls->RI.relTol*eps;

  /* DEPART */
  FOR_RANKS {
    const ora_lhs_t *lhs = &w->lhs[r];
    for (i = 0; i < lhs->nnz; i++) {
      const double *t = &Val[r][(size_t)i*dd];
      if (nsd == 2) {
        mK[r][(size_t)i*4+0] = t[0]; mK[r][(size_t)i*4+1] = t[1];
        mK[r][(size_t)i*4+2] = t[3]; mK[r][(size_t)i*4+3] = t[4];
        mG[r][(size_t)i*2+0] = t[2]; mG[r][(size_t)i*2+1] = t[5];
        mD[r][(size_t)i*2+0] = t[6]; mD[r][(size_t)i*2+1] = t[7];
        mL[r][i] = t[8];
      } else if (nsd == 3) {
        mK[r][(size_t)i*9+0] = t[0]; mK[r][(size_t)i*9+1] = t[1]; mK[r][(size_t)i*9+2] = t[2];
        mK[r][(size_t)i*9+3] = t[4]; mK[r][(size_t)i*9+4] = t[5]; mK[r][(size_t)i*9+5] = t[6];
        mK[r][(size_t)i*9+6] = t[8]; mK[r][(size_t)i*9+7] = t[9]; mK[r][(size_t)i*9+8] = t[10];
        mG[r][(size_t)i*3+0] = t[3]; mG[r][(size_t)i*3+1] = t[7]; mG[r][(size_t)i*3+2] = t[11];
        mD[r][(size_t)i*3+0] = t[12]; mD[r][(size_t)i*3+1] = t[13]; mD[r][(size_t)i*3+2] = t[14];
        mL[r][i] = t[15];
      } else {
        fprintf(stderr, "FSILS: Not defined nsd for DEPART %d\n", nsd); abort();
      }
    }
    for (i = 1; i <= lhs->nNo; i++) {
      for (j = lhs->rowPtr[2*(i-1)]; j <= lhs->rowPtr[2*(i-1)+1]; j++) {
        k = lhs->colPtr[j-1];
        for (int l = lhs->rowPtr[2*(k-1)]; l <= lhs->rowPtr[2*(k-1)+1]; l++) {
          if (lhs->colPtr[l-1] == i) {
            for (int d = 0; d < nsd; d++)
              Gt[r][(size_t)(l-1)*nsd + d] = -mG[r][(size_t)(j-1)*nsd + d];
            break;
          }
        }
      }
    }
  }
  bcpre(w, nsd);

  iBB = 0;
  for (i = 1; i <= mItr; i++) {
    iB  = 2*i - 1;
    iBB = 2*i;
    ls->RI.dB = ls->RI.fNorm;

    gmres_outofplace(w, &ls->GM, nsd, (const double *const *)mK,
                     (const double *const *)Rm, U[i-1]);
    ora_sparmul_vs(w, nsd, (const double *const *)mD,
                   (const double *const *)U[i-1], P[i-1]);
    FOR_RANKS for (int a = 0; a < w->lhs[r].nNo; a++)
      P[i-1][r][a] = Rc[r][a] - P[i-1][r][a];
    cgrad_schur(w, &ls->CG, nsd, (const double *const *)Gt,
                (const double *const *)mG, (const double *const *)mL, P[i-1]);
    ora_sparmul_sv(w, nsd, (const double *const *)mG,
                   (const double *const *)P[i-1], MU[iB-1]);
    FOR_RANKS {
      size_t n = (size_t)w->lhs[r].nNo*(size_t)nsd;
      for (size_t q = 0; q < n; q++) MU[iBB-1][r][q] = Rm[r][q] - MU[iB-1][r][q];
    }
    gmres_outofplace(w, &ls->GM, nsd, (const double *const *)mK,
                     (const double *const *)MU[iBB-1], U[i-1]);
    ora_sparmul_vv(w, nsd, (const double *const *)mK,
                   (const double *const *)U[i-1], MU[iBB-1]);
    addbcmul(w, ORA_BCOP_TYPE_ADD, nsd, (const double *const *)U[i-1], MU[iBB-1]);
    ora_sparmul_ss(w, (const double *const *)mL, (const double *const *)P[i-1],
                   MP[iB-1]);
    ora_sparmul_vs(w, nsd, (const double *const *)mD,
                   (const double *const *)U[i-1], MP[iBB-1]);

    c = 0;
    for (k = iB; k <= iBB; k++) {
      for (j = 1; j <= k; j++) {
        c = c + 1;
        tmp[c-1] = ora_dotv(w, nsd, (const double *const *)MU[j-1], (const double *const *)MU[k-1])
                 + ora_dotv(w, 1, (const double *const *)MP[j-1], (const double *const *)MP[k-1]);
      }
      c = c + 1;
      tmp[c-1] = ora_dotv(w, nsd, (const double *const *)MU[k-1], (const double *const *)Rmi)
               + ora_dotv(w, 1, (const double *const *)MP[k-1], (const double *const *)Rci);
    }
    c = 0;
    for (k = iB; k <= iBB; k++) {
      for (j = 1; j <= k; j++) {
        c = c + 1;
        AM(j,k) = tmp[c-1];
        AM(k,j) = tmp[c-1];
      }
      c = c + 1;
      B[k-1] = tmp[c-1];
    }

    memcpy(xB, B, sizeof(double)*(size_t)nB);
    if (ge(nB, iBB, A, xB)) {
      memcpy(oldxB, xB, sizeof(double)*(size_t)nB);
    } else {
      fprintf(stderr, "FSILS: Singular matrix detected\n");
      memcpy(xB, oldxB, sizeof(double)*(size_t)nB);
      if (i > 1) { iB = iB - 2; iBB = iBB - 2; }
      break;
    }

    {
      double sum = 0.0;
      for (j = 0; j < iBB; j++) sum += xB[j]*B[j];
      ls->RI.fNorm = ls->RI.iNorm*ls->RI.iNorm - sum;
    }
    if (ls->RI.fNorm < eps*eps) { ls->RI.suc = 1; break; }

    FOR_RANKS {
      size_t n = (size_t)w->lhs[r].nNo;
      for (size_t q = 0; q < n*nsd; q++) Rm[r][q] = Rmi[r][q] - xB[0]*MU[0][r][q];
      for (size_t q = 0; q < n; q++) Rc[r][q] = Rci[r][q] - xB[0]*MP[0][r][q];
      for (j = 2; j <= iBB; j++) {
        for (size_t q = 0; q < n*nsd; q++) Rm[r][q] = Rm[r][q] - xB[j-1]*MU[j-1][r][q];
        for (size_t q = 0; q < n; q++) Rc[r][q] = Rc[r][q] - xB[j-1]*MP[j-1][r][q];
      }
    }
  }
  last_i = i;
  if (last_i > mItr) {
    ls->RI.itr = mItr;
  } else {
    ls->RI.itr = last_i;
    FOR_RANKS {
      size_t n = (size_t)w->lhs[r].nNo;
      for (size_t q = 0; q < n; q++) Rc[r][q] = Rci[r][q] - xB[0]*MP[0][r][q];
      for (j = 2; j <= iBB; j++)
        for (size_t q = 0; q < n; q++) Rc[r][q] = Rc[r][q] - xB[j-1]*MP[j-1][r][q];
    }
  }
  {
    double nc = ora_normv(w, 1, (const double *const *)Rc);
    ls->Resc = (int)lround(100.0*(nc*nc)/ls->RI.fNorm);
    ls->Resm = 100 - ls->Resc;
  }

  FOR_RANKS {
    size_t n = (size_t)w->lhs[r].nNo;
    for (size_t q = 0; q < n*nsd; q++) Rmi[r][q] = xB[1]*U[0][r][q];
    for (size_t q = 0; q < n; q++) Rci[r][q] = xB[0]*P[0][r][q];
    for (i = 2; i <= ls->RI.itr; i++) {
      iB  = 2*i - 1;
      iBB = 2*i;
      for (size_t q = 0; q < n*nsd; q++) Rmi[r][q] = Rmi[r][q] + xB[iBB-1]*U[i-1][r][q];
      for (size_t q = 0; q < n; q++) Rci[r][q] = Rci[r][q] + xB[iB-1]*P[i-1][r][q];
    }
  }

  ls->RI.callD = ora_wtime() - t0;
  ls->RI.dB    = 5.0*log(ls->RI.fNorm/ls->RI.dB);

  if (ls->Resc < 0 || ls->Resm < 0) {
    ls->Resc = 0;
    ls->Resm = 0;
    ls->RI.dB = 0;
    ls->RI.fNorm = 0.0;
    fprintf(stderr, "Warning: unexpected behavior in FSILS (likely due to the"
                    " ill-conditioned LHS matrix)\n");
  }
  ls->RI.fNorm = sqrt(ls->RI.fNorm);

  FOR_RANKS {
    for (int a = 0; a < w->lhs[r].nNo; a++) {
      for (int d = 0; d < nsd; d++) Ri[r][(size_t)a*dof + d] = Rmi[r][(size_t)a*nsd + d];
      Ri[r][(size_t)a*dof + dof-1] = Rci[r][a];
    }
  }
  /* LOGFILE (L/NSSOLVER.f:343-379) writes FSILS_NS.log -- not restated */

  for (i = 0; i < mItr; i++) { wv_free(w, U[i]); wv_free(w, P[i]); }
  for (i = 0; i < 2*mItr; i++) { wv_free(w, MU[i]); wv_free(w, MP[i]); }
  free(U); free(P); free(MU); free(MP);
  free(tmp); free(A); free(B); free(xB); free(oldxB);
  wv_free(w, Rm); wv_free(w, Rc); wv_free(w, Rmi); wv_free(w, Rci);
  wv_free(w, mK); wv_free(w, mG); wv_free(w, mD); wv_free(w, mL); wv_free(w, Gt);
#undef AM
}

/* ================================================================== */
/* L/SOLVE.f:51-143 FSILS_SOLVE.  Ri[r] = Ri(dof,nNo) in svFSI local order
 * (in: RHS, out: solution); Val[r] is Jacobi-scaled in place. */
void ora_fsils_solve(ora_world_t *w, ora_ls_t *ls, int dof, double *const *Ri,
                     double *const *Val, int prec, const int *incL,
                     const double *res) {
  double **R = wv_alloc(w, dof), **Wc = wv_alloc(w, dof);
  int nFaces = w->lhs[0].nFaces;
  if (prec != ORA_PRECOND_FSILS && prec != ORA_PRECOND_RCS) {
    /* L/SOLVE.f:104-107 only prints and carries on with an undefined Wc */
    fprintf(stderr, "FSILS: this preconditioner is not supported\n"); abort();
  }

  FOR_RANKS {
    ora_lhs_t *lhs = &w->lhs[r];
    if (nFaces != 0) {
      int anyNeu = 0;
      for (int f = 0; f < nFaces; f++) {
        lhs->face[f].incFlag = 1;
        if (incL && incL[f] == 0) lhs->face[f].incFlag = 0;
        if (lhs->face[f].bGrp == ORA_BC_TYPE_NEU) anyNeu = 1;
      }
      if (!res && anyNeu) {
        fprintf(stderr, "FSILS: res is required for Neu surfaces\n"); abort();
      }
      for (int f = 0; f < nFaces; f++) {
        lhs->face[f].coupledFlag = 0;
        if (!lhs->face[f].incFlag) continue;
        if (lhs->face[f].bGrp == ORA_BC_TYPE_NEU && res[f] != 0.0) {
          lhs->face[f].res = res[f];
          lhs->face[f].coupledFlag = 1;
        }
      }
    }
    for (int a = 0; a < lhs->nNo; a++)
      for (int d = 0; d < dof; d++)
        R[r][(size_t)(lhs->map[a]-1)*dof + d] = Ri[r][(size_t)a*dof + d];
  }

  if (prec == ORA_PRECOND_FSILS) {
    preconddiag(w, dof, Val, R, Wc);
  } else {
    double **Wr = wv_alloc(w, dof);
    precondrcs(w, dof, Val, R, Wr, Wc);
    wv_free(w, Wr);
  }

  switch (ls->LS_type) {
  case ORA_LS_TYPE_NS:
    nssolver(w, ls, dof, (const double *const *)Val, R);
    break;
  case ORA_LS_TYPE_GMRES:
    gmres_inplace(w, &ls->RI, dof, (const double *const *)Val, R, dof == 1);
    break;
  case ORA_LS_TYPE_CG:
    cgrad(w, &ls->RI, dof, (const double *const *)Val, R);
    break;
  case ORA_LS_TYPE_BICGS:
    bicgs(w, &ls->RI, dof, (const double *const *)Val, R);
    break;
  default:
    fprintf(stderr, "FSILS: LS_type not defined\n"); abort();
  }

  FOR_RANKS {
    ora_lhs_t *lhs = &w->lhs[r];
    size_t n = (size_t)lhs->nNo*(size_t)dof;
    for (size_t i = 0; i < n; i++) R[r][i] = Wc[r][i]*R[r][i];
    for (int a = 0; a < lhs->nNo; a++)
      for (int d = 0; d < dof; d++)
        Ri[r][(size_t)a*dof + d] = R[r][(size_t)(lhs->map[a]-1)*dof + d];
  }
  wv_free(w, R); wv_free(w, Wc);
}
