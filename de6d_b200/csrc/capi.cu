// C-ABI plumbing shared by every translation unit of libde6d_b200.so: error reporting and library info.
// Error convention: every entry point returns DE6D_OK (0) or a non-zero code and records a message that
// de6d_last_error_string() returns (thread-local).  The reference convention -- fprintf(stderr) + exit(-1)
// inside the native code (sampling_gpu.cu:47-51, iou3d_nms.cpp:31-38) -- is replaced by status codes so the
// host language can raise instead of dying.
#include "common.cuh"
#include <stdio.h>
#include <string.h>
#include <atomic>

static thread_local char g_err[512] = "";

int de6d_set_cuda_error(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
    return DE6D_ERR_CUDA;
}

int de6d_set_error(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

extern "C" const char *de6d_last_error_string(void) { return g_err; }

extern "C" int de6d_version(void) { return 100; }  // 0.1.0

// Number of kernel launches issued through this library since load (per process); bench.py reports it.
static std::atomic<long long> g_launches{0};
void de6d_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long de6d_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char *de6d_build_info(void) {
    return "de6d_b200 0.1.0 sm_100a "
#ifdef __CUDACC_VER_MAJOR__
           "nvcc"
#endif
        ;
}
