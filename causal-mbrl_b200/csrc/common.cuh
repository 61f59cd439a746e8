// Shared helpers for libcmrl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cmrl_b200.h"

namespace cmrl {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define CMRL_REQUIRE(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            cmrl::set_error(__VA_ARGS__); \
            return -1;                   \
        }                                \
    } while (0)

// call after every kernel launch: returns cudaError_t (>0) through the C ABI
#define CMRL_CHECK_LAUNCH(name)                                                   \
    do {                                                                          \
        cudaError_t _e = cudaGetLastError();                                      \
        if (_e != cudaSuccess) {                                                  \
            cmrl::set_error("%s: %s", name, cudaGetErrorString(_e));               \
            return (int)_e;                                                       \
        }                                                                         \
        cmrl::count_launch();                                                     \
    } while (0)

#define CMRL_CUDA(call)                                                           \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) {                                                  \
            cmrl::set_error("%s: %s", #call, cudaGetErrorString(_e));              \
            return (int)_e;                                                       \
        }                                                                         \
    } while (0)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// torch.nn.functional.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplusf_(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

__device__ __forceinline__ float act_fwd(float z, int act) {
    if (act == CMRL_ACT_RELU) return fmaxf(z, 0.0f);
    if (act == CMRL_ACT_SILU) return z * sigmoidf_(z);
    return z;
}

__device__ __forceinline__ float act_grad(float z, int act) {
    if (act == CMRL_ACT_RELU) return z > 0.0f ? 1.0f : 0.0f;
    if (act == CMRL_ACT_SILU) {
        float s = sigmoidf_(z);
        return s * (1.0f + z * (1.0f - s));
    }
    return 1.0f;
}

// plain_transition.py:116-117
__device__ __forceinline__ float soft_clamp_logvar(float raw, float mn, float mx, float* lv1_out = nullptr) {
    float lv1 = mx - softplusf_(mx - raw);
    if (lv1_out) *lv1_out = lv1;
    return mn + softplusf_(lv1 - mn);
}

// "Last block out" counter update: after EVERY block of the grid has reached this call, the last one to arrive adds `delta`
// to counter[0].  counter[1] is the arrival ticket (zero between launches).  Blocks read counter[0] when they start, so a
// kernel can consume a device-side counter and advance it without a trailing one-thread launch (CUDA-graph replayable).
__device__ __forceinline__ void last_block_add(int* counter, int delta) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(reinterpret_cast<unsigned*>(counter + 1), 1u) == total - 1) {
            counter[1] = 0;
            counter[0] += delta;
            __threadfence();
        }
    }
}

// cudaFuncSetAttribute applies to the CURRENT device only: one bit per device and call site, so that a process driving
// several GPUs configures the kernels on each of them (a race between host threads only repeats the idempotent call)
struct PerDeviceOnce {
    unsigned long long mask = 0;
    bool pending(int* dev) {
        *dev = 0;
        cudaGetDevice(dev);
        return ((mask >> (*dev & 63)) & 1ull) == 0;
    }
    void done(int dev) { mask |= 1ull << (dev & 63); }
};

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace cmrl
