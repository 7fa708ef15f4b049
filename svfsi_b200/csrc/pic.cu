// pic.cu -- the generalised-alpha predictor / initiator / corrector of svFSI (S/PIC.f:40-297) and
// the strong Dirichlet overwrite (S/SETBC.f:39-228) on the device, so that a whole time step of
// Newton iterations runs without moving a nodal vector over PCIe:
//
//   PICP      An = Ao (gam-1)/gam ; Yn = Yo ; Dn = Do                    S/PIC.f:82-126
//   SETBCDIR  An(s:e,Ac) = tmpA(:,a) ; Yn(s:e,Ac) = tmpY(:,a)             S/SETBC.f:118-123
//   PICI      Ag = Ao(1-am) + An am ; Yg = Yo(1-af) + Yn af ; Dg likewise S/PIC.f:141-152
//   PICC      An -= R ; Yn -= R gam dt ; Dn -= R beta dt^2                S/PIC.f:203-207
//   (end of the time step)  Ao = An ; Yo = Yn ; Do = Dn                   S/MAIN.f:277-279
//
// The six state vectors live in the FSILS (reordered) numbering like everything else on the
// device; PICI writes straight into the Ag / Yg buffers the element kernels read.  All of these
// are one-pass streaming kernels (HBM bound, 3-5 vectors of tDof*nNo doubles each).
#include <cuda_runtime.h>

#include <vector>

#include "core.h"

namespace svfsi {

struct PicState {
  int tDof = 0, gen = -1;
  bool haveD = false;
  // Ao = An; Yo = Yn; Do = Dn (S/MAIN.f:277-279) is a pointer swap here; in the reference An/Yn/Dn
  // still EQUAL the new state after those copies, so until the next PICP overwrites the "new"
  // buffers every read of the new state is served from the "old" ones
  bool advanced = false;
  double *Ao = nullptr, *Yo = nullptr, *Do = nullptr, *An = nullptr, *Yn = nullptr, *Dn = nullptr,
         *Dg = nullptr;
};
static PicState g_pic;

static void pic_free() {
  double **ps[] = {&g_pic.Ao, &g_pic.Yo, &g_pic.Do, &g_pic.An, &g_pic.Yn, &g_pic.Dn, &g_pic.Dg};
  for (double **p : ps) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  g_pic.tDof = 0;
  g_pic.gen = -1;
}

__global__ void __launch_bounds__(256) picp_kernel(size_t n, double coef, const double *__restrict__ Ao,
                                                   const double *__restrict__ Yo,
                                                   const double *__restrict__ Do, double *__restrict__ An,
                                                   double *__restrict__ Yn, double *__restrict__ Dn) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    An[e] = Ao[e] * coef;
    Yn[e] = Yo[e];
    if (Dn) Dn[e] = Do[e];
  }
}
__global__ void __launch_bounds__(256) pici_kernel(size_t n, double c1, double c2, double c3, double c4,
                                                   const double *__restrict__ Ao,
                                                   const double *__restrict__ An,
                                                   const double *__restrict__ Yo,
                                                   const double *__restrict__ Yn,
                                                   const double *__restrict__ Do,
                                                   const double *__restrict__ Dn, double *__restrict__ Ag,
                                                   double *__restrict__ Yg, double *__restrict__ Dg) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    Ag[e] = Ao[e] * c1 + An[e] * c2;
    Yg[e] = Yo[e] * c3 + Yn[e] * c4;
    if (Dg) Dg[e] = Do[e] * c3 + Dn[e] * c4;
  }
}
__global__ void __launch_bounds__(256) picc_kernel(size_t n, double c1, double c2,
                                                   const double *__restrict__ R, double *__restrict__ An,
                                                   double *__restrict__ Yn, double *__restrict__ Dn) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (size_t)gridDim.x * blockDim.x) {
    const double r = R[e];
    An[e] = An[e] - r;
    Yn[e] = Yn[e] - r * c1;
    if (Dn) Dn[e] = Dn[e] - r * c2;
  }
}
__global__ void setbcdir_kernel(int faNo, int tDof, int s, int lDof, const int *__restrict__ glob,
                                const double *__restrict__ tmpA, const double *__restrict__ tmpY,
                                double *__restrict__ lA, double *__restrict__ lY) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= faNo * lDof) return;
  const int a = t / lDof, i = t - a * lDof;
  const size_t at = (size_t)glob[a] * tDof + s + i;
  lA[at] = tmpA[t];
  lY[at] = tmpY[t];
}

const double *pic_state_Yn() {
  if (!(g_pic.Yn && g_pic.gen == ctx().lhsGen && g_pic.tDof == 4)) return nullptr;
  return g_pic.advanced ? g_pic.Yo : g_pic.Yn;
}

static int pic_ready() {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "gpu_pic_*: FSILS_LHS_CREATE has not been called");
  if (!g_pic.Ao || g_pic.gen != c.lhsGen)
    return fail(SVFSI_ERR_STATE, "gpu_pic_*: gpu_pic_init_ has not been called for this lhs");
  return 0;
}
static unsigned pic_grid(size_t n) {
  const size_t want = (n + 255) / 256;
  return (unsigned)(want < 148 * 8 ? want : 148 * 8);
}

}  // namespace svfsi

using namespace svfsi;

extern "C" {

int32_t gpu_pic_init_(const int32_t *tDof, const double *Ao, const double *Yo, const double *Do) {
  Ctx &c = ctx();
  if (!c.lhs) return fail(SVFSI_ERR_STATE, "gpu_pic_init_: FSILS_LHS_CREATE has not been called");
  if (*tDof < 1 || *tDof > 4) return fail(SVFSI_ERR_ARG, "gpu_pic_init_: tDof must be 1..4");
  pic_free();
  const size_t bytes = sizeof(double) * (size_t)c.nNo * 4;
  g_pic.tDof = *tDof;
  g_pic.haveD = (Do != nullptr);
  CUDA_TRY(cudaMalloc(&g_pic.Ao, bytes));
  CUDA_TRY(cudaMalloc(&g_pic.Yo, bytes));
  CUDA_TRY(cudaMalloc(&g_pic.An, bytes));
  CUDA_TRY(cudaMalloc(&g_pic.Yn, bytes));
  if (g_pic.haveD) {
    CUDA_TRY(cudaMalloc(&g_pic.Do, bytes));
    CUDA_TRY(cudaMalloc(&g_pic.Dn, bytes));
    CUDA_TRY(cudaMalloc(&g_pic.Dg, bytes));
  }
  if (int rc = upload_nodal(Ao, *tDof, g_pic.Ao)) return rc;
  if (int rc = upload_nodal(Yo, *tDof, g_pic.Yo)) return rc;
  if (g_pic.haveD)
    if (int rc = upload_nodal(Do, *tDof, g_pic.Do)) return rc;
  // An/Yn/Dn start as copies of the old state (S/INITIALIZE.f: An = Ao etc.)
  CUDA_TRY(cudaMemcpyAsync(g_pic.An, g_pic.Ao, bytes, cudaMemcpyDeviceToDevice, c.stream));
  CUDA_TRY(cudaMemcpyAsync(g_pic.Yn, g_pic.Yo, bytes, cudaMemcpyDeviceToDevice, c.stream));
  if (g_pic.haveD)
    CUDA_TRY(cudaMemcpyAsync(g_pic.Dn, g_pic.Do, bytes, cudaMemcpyDeviceToDevice, c.stream));
  g_pic.gen = c.lhsGen;
  g_pic.advanced = false;
  return 0;
}

int32_t gpu_pic_free_(void) {
  pic_free();
  return 0;
}

int32_t gpu_picp_(const double *gam) {
  if (int rc = pic_ready()) return rc;
  Ctx &c = ctx();
  const size_t n = (size_t)c.nNo * g_pic.tDof;
  const double coef = (*gam - 1.0) / *gam;
  g_pic.advanced = false;
  picp_kernel<<<pic_grid(n), 256, 0, c.stream>>>(n, coef, g_pic.Ao, g_pic.Yo, g_pic.Do, g_pic.An,
                                                 g_pic.Yn, g_pic.haveD ? g_pic.Dn : nullptr);
  count_launch();
  return 0;
}

int32_t gpu_setbcdir_(const int32_t *faNo, const int32_t *gN, const int32_t *s, const int32_t *lDof,
                      const double *tmpA, const double *tmpY) {
  if (int rc = pic_ready()) return rc;
  Ctx &c = ctx();
  const int n = *faNo, ld = *lDof, s0 = *s - 1;
  if (n <= 0) return 0;
  if (s0 < 0 || s0 + ld > g_pic.tDof) return fail(SVFSI_ERR_ARG, "gpu_setbcdir_: s/lDof outside tDof");
  std::vector<int> glob(n);
  for (int a = 0; a < n; a++) {
    if (gN[a] < 1 || gN[a] > c.nNo) return fail(SVFSI_ERR_ARG, "gpu_setbcdir_: node id out of range");
    glob[a] = c.map[gN[a] - 1];
  }
  const size_t nv = (size_t)n * ld;
  const size_t bytes = sizeof(double) * 2 * nv + sizeof(int) * (size_t)n + 64;
  if (int rc = ensure_stage(bytes)) return rc;
  double *dA = c.d_stage, *dY = c.d_stage + nv;
  int *dG = (int *)(c.d_stage + 2 * nv);
  CUDA_TRY(cudaMemcpyAsync(dA, tmpA, sizeof(double) * nv, cudaMemcpyHostToDevice, c.stream));
  CUDA_TRY(cudaMemcpyAsync(dY, tmpY, sizeof(double) * nv, cudaMemcpyHostToDevice, c.stream));
  CUDA_TRY(cudaMemcpyAsync(dG, glob.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c.stream));
  setbcdir_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, c.stream>>>(n, g_pic.tDof, s0, ld, dG, dA, dY,
                                                                     g_pic.An, g_pic.Yn);
  count_launch();
  // glob (a host vector) and the caller's arrays may go away right after the call
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  return 0;
}

int32_t gpu_pici_(const double *am, const double *af) {
  if (int rc = pic_ready()) return rc;
  Ctx &c = ctx();
  if (!c.d_Ag || !c.d_Yg) {
    if (!c.d_Ag) CUDA_TRY(cudaMalloc(&c.d_Ag, sizeof(double) * (size_t)c.nNo * 4));
    if (!c.d_Yg) CUDA_TRY(cudaMalloc(&c.d_Yg, sizeof(double) * (size_t)c.nNo * 4));
    if (!c.d_Bf) CUDA_TRY(cudaMalloc(&c.d_Bf, sizeof(double) * (size_t)c.nNo * 3));
  }
  const size_t n = (size_t)c.nNo * g_pic.tDof;
  pici_kernel<<<pic_grid(n), 256, 0, c.stream>>>(n, 1.0 - *am, *am, 1.0 - *af, *af, g_pic.Ao, g_pic.An,
                                                 g_pic.Yo, g_pic.Yn, g_pic.Do, g_pic.Dn, c.d_Ag, c.d_Yg,
                                                 g_pic.haveD ? g_pic.Dg : nullptr);
  count_launch();
  return 0;
}

int32_t gpu_picc_(const double *gam, const double *beta, const double *dt) {
  if (int rc = pic_ready()) return rc;
  Ctx &c = ctx();
  if (!c.d_R || c.dof != g_pic.tDof)
    return fail(SVFSI_ERR_STATE, "gpu_picc_: no device-resident increment of this dof");
  const size_t n = (size_t)c.nNo * g_pic.tDof;
  picc_kernel<<<pic_grid(n), 256, 0, c.stream>>>(n, *gam * *dt, *beta * *dt * *dt, c.d_R, g_pic.An,
                                                 g_pic.Yn, g_pic.haveD ? g_pic.Dn : nullptr);
  count_launch();
  return 0;
}

int32_t gpu_pic_advance_(void) {
  if (int rc = pic_ready()) return rc;
  // Ao = An etc. (S/MAIN.f:277-279): swap the buffers instead of copying
  std::swap(g_pic.Ao, g_pic.An);
  std::swap(g_pic.Yo, g_pic.Yn);
  if (g_pic.haveD) std::swap(g_pic.Do, g_pic.Dn);
  g_pic.advanced = true;
  return 0;
}

int32_t gpu_pic_get_(const int32_t *which, double *A, double *Y, double *D) {
  if (int rc = pic_ready()) return rc;
  // 0: old (Ao, Yo, Do), 1: new (An, Yn, Dn); right after gpu_pic_advance_ both are the same state
  const bool nw = (*which != 0) && !g_pic.advanced;
  if (A) if (int rc = download_nodal(nw ? g_pic.An : g_pic.Ao, g_pic.tDof, A)) return rc;
  if (Y) if (int rc = download_nodal(nw ? g_pic.Yn : g_pic.Yo, g_pic.tDof, Y)) return rc;
  if (D && g_pic.haveD)
    if (int rc = download_nodal(nw ? g_pic.Dn : g_pic.Do, g_pic.tDof, D)) return rc;
  return 0;
}

}  // extern "C"
