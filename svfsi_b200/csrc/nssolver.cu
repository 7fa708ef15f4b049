// nssolver.cu -- NSSOLVER (L/NSSOLVER.f:52-383) on the device.  (placeholder until the
// GMRES/CG paths are validated on hardware)
#include "core.h"
namespace svfsi {
int nssolver_dev(svfsi_ls_t *ls, int dof, const double *Val, double *R) {
  (void)ls; (void)dof; (void)Val; (void)R;
  return fail(SVFSI_ERR_UNSUPPORTED, "NSSOLVER not yet implemented on the device");
}
}  // namespace svfsi
