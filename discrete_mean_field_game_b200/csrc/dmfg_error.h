// Error plumbing shared by the translation units of libdmfg: a thread-local message and the
// status-code convention of include/dmfg.h (nothing throws across the ABI).
#pragma once
#include <cuda_runtime.h>

namespace dmfg {
// records the message for dmfg_last_error() and returns `code`
int fail(int code, const char* fmt, ...);
int sm_count(int* out);
// number of kernels this process has launched through libdmfg (dmfg_kernel_launches): bumped by DMFG_LAUNCHED
void count_launch();
}  // namespace dmfg

#define DMFG_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return ::dmfg::fail(DMFG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                __FILE__, __LINE__);                                             \
    } while (0)

// after every kernel launch: count it (dmfg_kernel_launches) and surface a launch error
#define DMFG_LAUNCHED()                  \
    do {                                 \
        ::dmfg::count_launch();          \
        DMFG_CUDA(cudaGetLastError());   \
    } while (0)
