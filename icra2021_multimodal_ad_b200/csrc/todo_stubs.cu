// Entry points declared in include/mmad.h that are not implemented yet.
#include "mmad_internal.cuh"
using namespace mmad;
extern "C" {
size_t mmad_train_workspace_bytes(mmad_t, int) { return 0; }
int mmad_train_fwd_bwd(mmad_t, const float*, int, int, long long, const mmad_train_layer_t*, const mmad_train_layer_t*, const float*, float, float, float*, void*, size_t, mmad_allreduce_fn, void*, void*) { set_error("not implemented"); return MMAD_E_UNSUPPORTED; }
int mmad_adam_step(int, float* const*, float* const*, float* const*, float* const*, const long long*, int, float, float, float, float, float, void*) { set_error("not implemented"); return MMAD_E_UNSUPPORTED; }
}
