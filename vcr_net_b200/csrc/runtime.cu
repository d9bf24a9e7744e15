// Library-wide bookkeeping: launch counter and version string.
#include "common.cuh"

unsigned long long g_vcr_launches = 0;

// number of kernels this library has launched in this process (all threads)
VCR_API long long vcr_launch_count(void) { return (long long)__atomic_load_n(&g_vcr_launches, __ATOMIC_RELAXED); }

VCR_API int vcr_abi_version(void) { return 1; }
