// Microbenchmark: tcgen05.mma (kind::f16, fp16 x fp16 -> fp32) issue/execute rate vs shared-memory layout and
// cta_group on sm_100a.  Operands are garbage; only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_rate profiles/tools/mma_rate.cu && /tmp/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, a = smem_u32(bar);
    long long spins = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (++spins > (1ll << 26)) __trap();
    }
}

// LAYOUT: 0 = no swizzle (A: LBO 2048 / SBO 128 chunk-major; B: LBO 128 / SBO K/8*128), 6 = SW32 panels, 2 = SW128
template <int CTAS>
__global__ void bench(int N, int layout, int n_mma, int ksteps_per_commit, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x / 32;
    uint32_t rank = 0;
    if (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CTAS == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;  // fp16 1.0
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CTAS == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (threadIdx.x == 0 && rank == 0) {
        const int M = 128 * CTAS;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 64 * 1024);
        const int K = 208;
        uint64_t ad0, bd0;
        uint32_t a_step, b_step;  // start-address advance per K step, in 16 B units
        if (layout == 0) {
            ad0 = make_desc(a_base, 2048, 128, 0);
            bd0 = make_desc(b_base, 128, (K / 8) * 128, 0);
            a_step = 4096 / 16;
            b_step = 256 / 16;
        } else if (layout == 6) {  // SW32: one 32-byte-wide panel per K step, rows 32 B apart, 8-row groups 256 B apart
            ad0 = make_desc(a_base, 16, 256, 6);
            bd0 = make_desc(b_base, 16, 256, 6);
            a_step = (128 * 32) / 16;
            b_step = (uint32_t)((N / CTAS) * 32) / 16;
        } else {  // SW128: 64-element K atoms (8 rows x 128 B = 1024 B), K step = +32 B inside the atom
            ad0 = make_desc(a_base, 16, 1024, 2);
            bd0 = make_desc(b_base, 16, 1024, 2);
            a_step = 32 / 16;
            b_step = 32 / 16;
        }
        long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const int ks = i % 13;
            uint64_t ad = ad0 + (uint64_t)(ks * a_step), bd = bd0 + (uint64_t)(ks * b_step);
            if (layout == 2) {  // wrap inside 4 atoms of 64
                ad = ad0 + (uint64_t)((ks % 4) * 2 + (ks / 4) * (128 * 128 / 16));
                bd = bd0 + (uint64_t)((ks % 4) * 2 + (ks / 4) * ((N / CTAS) * 128 / 16));
            }
            if (CTAS == 1)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
        }
        long long t1 = clock64();
        if (CTAS == 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CTAS == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 0) {
        if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int n_mma = 208;
    for (int ctas : {1, 2})
        for (int layout : {0, 6, 2})
            for (int N : {208, 256, 128}) {
                for (int rep = 0; rep < 2; ++rep) {
                    if (ctas == 1) bench<1><<<1, 128, smem>>>(N, layout, n_mma, 13, out);
                    else {
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                        cfg.attrs = at; cfg.numAttrs = 1;
                        cudaLaunchKernelEx(&cfg, bench<2>, N, layout, n_mma, 13, out);
                    }
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                long long h[2];
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("cta_group::%d layout=%d (0 none, 6 sw32, 2 sw128) M=%d N=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f)\n",
                       ctas, layout, 128 * ctas, N, (double)h[0] / n_mma, (double)h[1] / n_mma, N / 2.0);
            }
    return 0;
}
