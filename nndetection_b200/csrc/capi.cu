// Error plumbing + version for the C ABI (include/nndet_b200.h).
#include "common.cuh"
#include <stdio.h>
#include <string.h>

namespace {
thread_local char g_err[512] = "";
}

unsigned long long g_nnd_launches = 0;

extern "C" {

unsigned long long nnd_launch_count(void) { return g_nnd_launches; }

int nnd_set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where ? where : "?", cudaGetErrorName(e), cudaGetErrorString(e));
  cudaGetLastError();   // clear the sticky-less error so later calls report their own
  return NND_ERR_CUDA;
}

const char* nnd_last_error(void) { return g_err; }

int nnd_abi_version(void) { return 1; }

const char* nnd_build_arch(void) { return "sm_100a"; }

}  // extern "C"
