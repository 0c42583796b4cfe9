// stubs.cu -- entry points declared in include/kameris_b200.h that are not built yet.
// They fail loudly (KAM_E_UNSUPPORTED); nothing falls back to the CPU.
#include "common.cuh"

#define KAM_STUB(name)                                   \
    kam_set_error(#name ": not implemented in this build"); \
    return KAM_E_UNSUPPORTED;

extern "C" {
int kam_dists_host(const void *, int, uint32_t, uint64_t, unsigned, uint32_t *, float *, float *, float *, float *) { KAM_STUB(kam_dists_host) }
}
