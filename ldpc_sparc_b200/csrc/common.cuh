// Shared host-side helpers for libldpcsparc_b200.so (error slot, launch counter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/ldpc_sparc_b200.h"

namespace lsb {

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// number of SMs / max opt-in smem of the current device (cached per device)
int sm_count();
int max_smem_optin();

// Raise -- never lower -- the dynamic shared-memory limit of `kernel` on the current device.  The
// attribute belongs to the (kernel, device) pair, not to a code / design handle: two handles of
// different size sharing one template instantiation must not shrink it under each other.
cudaError_t ensure_dynamic_smem(const void *kernel, size_t bytes);
template <typename K> inline cudaError_t ensure_dynamic_smem(K kernel, size_t bytes)
{
    return ensure_dynamic_smem(reinterpret_cast<const void *>(kernel), bytes);
}

}  // namespace lsb

#define LSB_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e_ = (call);                                                \
        if (e_ != cudaSuccess) {                                                \
            lsb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,         \
                           cudaGetErrorString(e_));                             \
            return LSB_ERR_CUDA;                                                \
        }                                                                       \
    } while (0)

#define LSB_REQUIRE(cond, ...)                                                  \
    do {                                                                        \
        if (!(cond)) {                                                          \
            lsb::set_error(__VA_ARGS__);                                        \
            return LSB_ERR_INVALID;                                             \
        }                                                                       \
    } while (0)
