// sg_internal.h -- shared by the translation units of libscrooge_b200.so
#pragma once
#include <string>
#include <cuda_runtime.h>

namespace sg {

extern thread_local std::string g_last_error;
int fail(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

}  // namespace sg

#define SG_STR2(x) #x
#define SG_STR(x) SG_STR2(x)
// Return SG_ERR_CUDA (with the message kept for sg_last_error) instead of the reference's exit()
// (src/cuda_util.hpp:3-10).
#define SG_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t sg_e_ = (call);                                                    \
        if (sg_e_ != cudaSuccess) {                                                    \
            cudaGetLastError();                                                        \
            return ::sg::cuda_fail(sg_e_, #call " (" __FILE__ ":" SG_STR(__LINE__) ")"); \
        }                                                                              \
    } while (0)
