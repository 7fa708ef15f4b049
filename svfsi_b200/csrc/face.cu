// face.cu -- boundary (face) integrals of the fluid Newton iteration on the device:
//   BASSEMNEUBC + BFLUID + GNNB   Neumann / backflow-stabilised traction on TRI3 faces of TET4
//                                 meshes, S/EQASSEM.f:90-192, S/FLUID.f:1279-1336, S/NN.f:1856-1996
//   IntegV                        flux of a nodal vector through a face (resistance BCs),
//                                 S/ALLFUN.f:199-262
// A face has O(1e3..1e4) elements, so this is latency- not bandwidth-bound work; what matters is that
// it no longer forces R / Val back to the host between the element loop and the linear solve.
// Two kernels, no atomics, bitwise repeatable: (1) one thread per face element evaluates its three
// Gauss points into lR(3,3) and the diagonal tangent scalar lKd(3,3); (2) one thread per face NODE
// adds the contributions of the face elements around it in ascending element order -- the order of
// the reference's element loop -- into its own row of R and Val (every block belongs to one row).
#include <cuda_runtime.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "core.h"

namespace svfsi {

struct MeshFace {
  bool created = false;
  int gen = -1;
  int nNo = 0, nEl = 0;
  int *d_node = nullptr;   // [nEl][3] reordered node ids
  int *d_loc = nullptr;    // [nEl][3] face-local node index (into hgN)
  int *d_par = nullptr;    // [nEl] parent element (0-based)
  int *d_opp = nullptr;    // [nEl] reordered id of the parent's node that is not on the face
  int *d_dest = nullptr;   // [nEl][9] block index of (row node a, col node b)
  int *d_glob = nullptr;   // [nNo] reordered id of every face node
  int *d_adjPtr = nullptr; // [nNo+1] -> d_adj: (e*3 + a) in ascending e
  int *d_adj = nullptr;
  double *d_h = nullptr;   // [nNo] Neumann value per face node
  double *d_lR = nullptr;  // [nEl][9]
  double *d_lK = nullptr;  // [nEl][9]
  double *d_part = nullptr; // [nEl] flux partials
};
static std::vector<MeshFace> g_faces;

static void face_release(MeshFace &f) {
  int **ip[] = {&f.d_node, &f.d_loc, &f.d_par, &f.d_opp, &f.d_dest, &f.d_glob, &f.d_adjPtr, &f.d_adj};
  for (int **p : ip) { if (*p) cudaFree(*p); *p = nullptr; }
  double **dp[] = {&f.d_h, &f.d_lR, &f.d_lK, &f.d_part};
  for (double **p : dp) { if (*p) cudaFree(*p); *p = nullptr; }
  f = MeshFace();
}
void faces_free_all() {
  for (MeshFace &f : g_faces) face_release(f);
  g_faces.clear();
}

// parent's off-face node (GNNB's ptr(eNoNb+1), S/NN.f:1879-1917) and the block of every (a,b)
__global__ void face_setup_kernel(int nEl, const int *__restrict__ node, const int *__restrict__ par,
                                  const int *__restrict__ ien, const int *__restrict__ rowPtr,
                                  const int *__restrict__ col, int *__restrict__ opp,
                                  int *__restrict__ dest, int *__restrict__ bad) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEl) return;
  const int n0 = node[e * 3], n1 = node[e * 3 + 1], n2 = node[e * 3 + 2];
  int o = -1, found = 0;
  for (int b = 0; b < 4; b++) {
    const int v = ien[(size_t)par[e] * 4 + b];
    if (v == n0 || v == n1 || v == n2) found++;
    else o = v;
  }
  if (found != 3 || o < 0) atomicAdd(bad, 1);   // "could not find matching face nodes"
  opp[e] = o;
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) {
      const int row = node[e * 3 + a], c = node[e * 3 + b];
      int p = -1;
      for (int j = rowPtr[row]; j < rowPtr[row + 1]; j++)
        if (col[j] == c) { p = j; break; }
      if (p < 0) atomicAdd(bad, 1);
      dest[e * 9 + a * 3 + b] = p;
    }
}

// GNNB for a TRI3 face of a TET4 element (S/NN.f:1974-1991): n = (x1-x3) x (x2-x3), flipped to
// point away from the parent's off-face node.  |n| = 2 * area.
__device__ __forceinline__ void gnnb_tri3(const double *__restrict__ x, int n0, int n1, int n2, int o,
                                          double n[3]) {
  double a[3], b[3], v[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double x2 = x[(size_t)n2 * 3 + i];
    a[i] = 0.0 + 1.0 * x[(size_t)n0 * 3 + i] + 0.0 * x[(size_t)n1 * 3 + i] + (-1.0) * x2;
    b[i] = 0.0 + 0.0 * x[(size_t)n0 * 3 + i] + 1.0 * x[(size_t)n1 * 3 + i] + (-1.0) * x2;
    v[i] = x[(size_t)n0 * 3 + i] - x[(size_t)o * 3 + i];
  }
  n[0] = a[1] * b[2] - a[2] * b[1];
  n[1] = a[2] * b[0] - a[0] * b[2];
  n[2] = a[0] * b[1] - a[1] * b[0];
  if (n[0] * v[0] + n[1] * v[1] + n[2] * v[2] < 0.0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
}
// TRI3 Gauss rule (S/NN.f:416-422) and shape functions (:1100-1103)
__device__ __forceinline__ void tri3_N(int g, double N[3]) {
  const double s = 2.0 / 3.0, t = 1.0 / 6.0;
  const double x1 = (g == 1) ? s : t, x2 = (g == 2) ? s : t;
  N[0] = x1; N[1] = x2; N[2] = 1.0 - x1 - x2;
}

__global__ void face_bfluid_kernel(int nEl, const int *__restrict__ node, const int *__restrict__ loc,
                                   const int *__restrict__ opp, const double *__restrict__ x,
                                   const double *__restrict__ Yg, const double *__restrict__ h,
                                   double rho, double bfStab, double T1, double *__restrict__ lRo,
                                   double *__restrict__ lKo) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEl) return;
  const int nd[3] = {node[e * 3], node[e * 3 + 1], node[e * 3 + 2]};
  double n[3];
  gnnb_tri3(x, nd[0], nd[1], nd[2], opp[e], n);
  const double Jac = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  const double nV[3] = {n[0] / Jac, n[1] / Jac, n[2] / Jac};
  const double w = (1.0 / 6.0) * Jac;
  const double wl = w * T1;                                   // w*af*gam*dt, S/FLUID.f:1291
  double yl[3][3], hl[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    hl[a] = h[loc[e * 3 + a]];
#pragma unroll
    for (int i = 0; i < 3; i++) yl[a][i] = Yg[(size_t)nd[a] * 4 + i];
  }
  double lR[3][3], lK[3][3];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int i = 0; i < 3; i++) { lR[a][i] = 0.0; lK[a][i] = 0.0; }
#pragma unroll
  for (int g = 0; g < 3; g++) {
    double N[3];
    tri3_N(g, N);
    double hh = 0.0, u[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      hh = hh + N[a] * hl[a];
#pragma unroll
      for (int i = 0; i < 3; i++) u[i] = u[i] + N[a] * yl[a][i];
    }
    double udn = 0.0;
#pragma unroll
    for (int i = 0; i < 3; i++) udn = udn + u[i] * nV[i];
    udn = 0.5 * bfStab * rho * (udn - fabs(udn));
    double hc[3];
#pragma unroll
    for (int i = 0; i < 3; i++) hc[i] = hh * nV[i] + udn * u[i];
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
      for (int i = 0; i < 3; i++) lR[a][i] = lR[a][i] - w * N[a] * hc[i];
#pragma unroll
      for (int b = 0; b < 3; b++) lK[a][b] = lK[a][b] - wl * N[a] * N[b] * udn;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int i = 0; i < 3; i++) {
      lRo[(size_t)e * 9 + a * 3 + i] = lR[a][i];
      lKo[(size_t)e * 9 + a * 3 + i] = lK[a][i];
    }
}

// DOASSEM (S/LHSA.f:266-298) of the face contributions, owner-computes per face node
__global__ void face_scatter_kernel(int nNo, const int *__restrict__ glob, const int *__restrict__ adjPtr,
                                    const int *__restrict__ adj, const int *__restrict__ dest,
                                    const double *__restrict__ lR, const double *__restrict__ lK,
                                    double *__restrict__ R, double *__restrict__ Val) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nNo) return;
  const size_t row = (size_t)glob[t];
  for (int k = adjPtr[t]; k < adjPtr[t + 1]; k++) {
    const int ea = adj[k], e = ea / 3, a = ea - e * 3;
    for (int i = 0; i < 3; i++) R[row * 4 + i] = R[row * 4 + i] + lR[(size_t)e * 9 + a * 3 + i];
    for (int b = 0; b < 3; b++) {
      const double v = lK[(size_t)e * 9 + a * 3 + b];
      double *blk = Val + (size_t)dest[e * 9 + a * 3 + b] * 16;
      blk[0] = blk[0] + v;      // lK(1), lK(6), lK(11): the velocity diagonal of the 4x4 block
      blk[5] = blk[5] + v;
      blk[10] = blk[10] + v;
    }
  }
}

// IntegV (S/ALLFUN.f:199-262): sum_e sum_g w(g) sum_a N(a,g) s(:,Ac).n
__global__ void face_flux_kernel(int nEl, const int *__restrict__ node, const int *__restrict__ opp,
                                 const double *__restrict__ x, const double *__restrict__ S, int ld,
                                 int s0, double *__restrict__ part) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nEl) return;
  const int nd[3] = {node[e * 3], node[e * 3 + 1], node[e * 3 + 2]};
  double n[3];
  gnnb_tri3(x, nd[0], nd[1], nd[2], opp[e], n);
  double acc = 0.0;
  for (int g = 0; g < 3; g++) {
    double N[3];
    tri3_N(g, N);
    double sHat = 0.0;
    for (int a = 0; a < 3; a++)
      for (int i = 0; i < 3; i++) sHat = sHat + N[a] * S[(size_t)nd[a] * ld + s0 + i] * n[i];
    acc = acc + (1.0 / 6.0) * sHat;
  }
  part[e] = acc;
}
// sum of the per-element partials in a FIXED order (deterministic): thread t adds elements t, t + 1024, ...,
// then a shuffle tree.  (One thread adding 8192 partials in ascending order took 305 us per Newton iteration.)
__global__ void __launch_bounds__(1024) face_sum_kernel(int n, const double *__restrict__ part,
                                                        double *__restrict__ out) {
  __shared__ double smem[32];
  double v = 0.0;
  for (int e = threadIdx.x; e < n; e += 1024) v = v + part[e];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = smem[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) *out = t;
  }
}

template <typename T>
static int up(T **d, const std::vector<T> &h) {
  if (*d) cudaFree(*d);
  *d = nullptr;
  CUDA_TRY(cudaMalloc((void **)d, sizeof(T) * std::max<size_t>(h.size(), 1)));
  if (!h.empty()) CUDA_TRY(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return 0;
}
static MeshFace *face_get(int iFa) {
  Ctx &c = ctx();
  if (iFa < 1 || iFa > (int)g_faces.size()) return nullptr;
  MeshFace &f = g_faces[iFa - 1];
  if (!f.created || f.gen != c.lhsGen) return nullptr;
  return &f;
}

const double *pic_state_Yn();   // pic.cu

}  // namespace svfsi

using namespace svfsi;

extern "C" {

int32_t gpu_face_create_(const int32_t *iFa, const int32_t *nNo_, const int32_t *gN,
                         const int32_t *nEl_, const int32_t *eNoN, const int32_t *IEN,
                         const int32_t *gE) {
  Ctx &c = ctx();
  if (!c.lhs || !c.mesh) return fail(SVFSI_ERR_STATE, "gpu_face_create_: needs FSILS_LHS_CREATE and gpu_mesh_create_");
  if (*eNoN != 3) return fail(SVFSI_ERR_UNSUPPORTED, "gpu_face_create_: only TRI3 faces of TET4 meshes");
  if (*iFa < 1 || *iFa > 4096) return fail(SVFSI_ERR_ARG, "gpu_face_create_: bad face id");
  if ((int)g_faces.size() < *iFa) g_faces.resize(*iFa);
  MeshFace &f = g_faces[*iFa - 1];
  face_release(f);
  const int nNo = *nNo_, nEl = *nEl_;
  f.nNo = nNo; f.nEl = nEl;
  std::unordered_map<int, int> loc;
  std::vector<int> glob(nNo);
  for (int a = 0; a < nNo; a++) {
    if (gN[a] < 1 || gN[a] > c.nNo) return fail(SVFSI_ERR_ARG, "gpu_face_create_: node id out of range");
    loc[gN[a]] = a;
    glob[a] = c.map[gN[a] - 1];
  }
  std::vector<int> node((size_t)nEl * 3), lc((size_t)nEl * 3), par(nEl);
  std::vector<std::vector<int>> adj(nNo);
  for (int e = 0; e < nEl; e++) {
    if (gE[e] < 1 || gE[e] > c.nEl) return fail(SVFSI_ERR_ARG, "gpu_face_create_: parent element out of range");
    par[e] = gE[e] - 1;
    for (int a = 0; a < 3; a++) {
      const int Ac = IEN[(size_t)e * 3 + a];
      auto it = loc.find(Ac);
      if (it == loc.end()) return fail(SVFSI_ERR_ARG, "gpu_face_create_: face element node not in gN");
      node[(size_t)e * 3 + a] = c.map[Ac - 1];
      lc[(size_t)e * 3 + a] = it->second;
      adj[it->second].push_back(e * 3 + a);      // ascending e by construction
    }
  }
  std::vector<int> adjPtr(nNo + 1, 0), adjL;
  for (int a = 0; a < nNo; a++) {
    adjL.insert(adjL.end(), adj[a].begin(), adj[a].end());
    adjPtr[a + 1] = (int)adjL.size();
  }
  if (int rc = up(&f.d_node, node)) return rc;
  if (int rc = up(&f.d_loc, lc)) return rc;
  if (int rc = up(&f.d_par, par)) return rc;
  if (int rc = up(&f.d_glob, glob)) return rc;
  if (int rc = up(&f.d_adjPtr, adjPtr)) return rc;
  if (int rc = up(&f.d_adj, adjL)) return rc;
  const size_t ne = std::max<size_t>(nEl, 1);
  CUDA_TRY(cudaMalloc(&f.d_opp, sizeof(int) * ne));
  CUDA_TRY(cudaMalloc(&f.d_dest, sizeof(int) * ne * 9));
  CUDA_TRY(cudaMalloc(&f.d_h, sizeof(double) * std::max<size_t>(nNo, 1)));
  CUDA_TRY(cudaMalloc(&f.d_lR, sizeof(double) * ne * 9));
  CUDA_TRY(cudaMalloc(&f.d_lK, sizeof(double) * ne * 9));
  CUDA_TRY(cudaMalloc(&f.d_part, sizeof(double) * ne));
  if (nEl > 0) {
    CUDA_TRY(cudaMemsetAsync(c.d_flag + 1, 0, sizeof(int), c.stream));
    face_setup_kernel<<<(nEl + 127) / 128, 128, 0, c.stream>>>(nEl, f.d_node, f.d_par, c.d_ien,
                                                               c.d_rowPtr, c.d_col, f.d_opp, f.d_dest,
                                                               c.d_flag + 1);
    count_launch();
    int bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, c.d_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    if (bad) {
      face_release(f);
      return fail(SVFSI_ERR_ARG, "gpu_face_create_: could not find matching face nodes / matrix blocks");
    }
  }
  f.created = true;
  f.gen = c.lhsGen;
  return 0;
}

int32_t gpu_face_free_(const int32_t *iFa) {
  if (*iFa >= 1 && *iFa <= (int)g_faces.size()) face_release(g_faces[*iFa - 1]);
  return 0;
}

int32_t gpu_bassem_neu_fluid_(const int32_t *iFa, const double *hgN, const double *rho,
                              const double *bfStab, const double *af, const double *gam,
                              const double *dt) {
  Ctx &c = ctx();
  MeshFace *f = face_get(*iFa);
  if (!f) return fail(SVFSI_ERR_STATE, "gpu_bassem_neu_fluid_: face not created for this lhs");
  if (!c.d_R || !c.d_Val || c.dof != 4 || !c.d_Yg)
    return fail(SVFSI_ERR_STATE, "gpu_bassem_neu_fluid_: no device-resident fluid system / state");
  if (f->nEl == 0) return 0;
  CUDA_TRY(cudaMemcpyAsync(f->d_h, hgN, sizeof(double) * (size_t)f->nNo, cudaMemcpyHostToDevice, c.stream));
  ProfScope ps(PROF_ASM);
  face_bfluid_kernel<<<(f->nEl + 127) / 128, 128, 0, c.stream>>>(
      f->nEl, f->d_node, f->d_loc, f->d_opp, c.d_x, c.d_Yg, f->d_h, *rho, *bfStab, *af * *gam * *dt,
      f->d_lR, f->d_lK);
  face_scatter_kernel<<<(f->nNo + 127) / 128, 128, 0, c.stream>>>(f->nNo, f->d_glob, f->d_adjPtr, f->d_adj,
                                                                 f->d_dest, f->d_lR, f->d_lK, c.d_R, c.d_Val);
  count_launch(2);
  return 0;
}

int32_t gpu_face_integ_v_(const int32_t *iFa, const int32_t *which, const int32_t *s, double *flux) {
  Ctx &c = ctx();
  MeshFace *f = face_get(*iFa);
  if (!f) return fail(SVFSI_ERR_STATE, "gpu_face_integ_v_: face not created for this lhs");
  const double *S = nullptr;
  if (*which == 0) S = c.d_Yg;
  else if (*which == 1) S = pic_state_Yn();
  if (!S) return fail(SVFSI_ERR_STATE, "gpu_face_integ_v_: the requested state vector is not on the device");
  if (int rc = ensure_small()) return rc;
  double *out = c.d_small + 96;
  CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double), c.stream));
  if (f->nEl > 0) {
    face_flux_kernel<<<(f->nEl + 127) / 128, 128, 0, c.stream>>>(f->nEl, f->d_node, f->d_opp, c.d_x, S, 4,
                                                                 *s - 1, f->d_part);
    face_sum_kernel<<<1, 1024, 0, c.stream>>>(f->nEl, f->d_part, out);
    count_launch(2);
  }
  if (int rc = allreduce_dev(out, 1)) return rc;    // cm%reduce (S/ALLFUN.f:258-259)
  CUDA_TRY(cudaMemcpyAsync(flux, out, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  return 0;
}

}  // extern "C"
