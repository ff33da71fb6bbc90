#include "context.h"
using namespace dnm;
extern "C" int dnm_evolve(dnm_mat_t, dnm_vec_t, dnm_vec_t, double, double, double, int, int, int *, int *, int *)
{
  set_error("dnm_evolve not built yet");
  return DNM_ERR_UNSUPPORTED;
}
extern "C" int dnm_eigsolve(dnm_mat_t, int, int, double, int, int, uint64_t, int, int *, double *, double *, dnm_vec_t *, int *, int *, int *)
{
  set_error("dnm_eigsolve not built yet");
  return DNM_ERR_UNSUPPORTED;
}
extern "C" int dnm_rdm(dnm_vec_t, const dnm_subspace_t *, int64_t, const int64_t *, double *)
{
  set_error("dnm_rdm not built yet");
  return DNM_ERR_UNSUPPORTED;
}
