// Empirical TMEM accumulator layout + rate for tcgen05.mma cta_group::2 with M=128 (64 rows per CTA) and
// cta_group::1 with M=64.  A[row][k] = row (k==0), B[n][k] = 1 (k==0)  =>  D[row][n] = row for every n.
// Each warp dumps what it sees in its 32 lanes for a few columns.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, a = smem_u32(bar);
    long long spins = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (++spins > (1ll << 26)) __trap();
    }
}
template <int CTAS>
__global__ void bench(int Mtotal, int N, int n_mma, float* dump, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint32_t rank = 0;
    if (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int rows_cta = Mtotal / CTAS;
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
    // A: rows_cta rows, K=16: chunk-major no-swizzle: chunk c (8 k) at c*(rows_cta*16) + row*16
    __half* A = (__half*)smem;
    __half* B = (__half*)(smem + 32 * 1024);  // N/CTAS rows, K=16: (n/8)*(2*128) + (k/8)*128 + (n%8)*16 + (k%8)*2 bytes
    for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) ((uint16_t*)smem)[i] = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < rows_cta; r += blockDim.x) A[r * 8] = __float2half((float)(rank * rows_cta + r));  // k = 0
    for (int n = threadIdx.x; n < N / CTAS; n += blockDim.x) B[(n / 8) * 128 + (n % 8) * 8] = __float2half(1.0f);
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
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(Mtotal >> 4) << 24);
        const uint64_t ad = make_desc(smem_u32(A), rows_cta * 16, 128), bd = make_desc(smem_u32(B), 128, 256);
        long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            if (CTAS == 1)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        }
        if (CTAS == 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        mbar_wait(&bar, 0);
        cyc[0] = clock64() - t0;
    } else {
        if (threadIdx.x == 0) mbar_wait(&bar, 0);
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t q[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]),
                       "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
                     : "r"(taddr + (uint32_t)(N - 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* d = dump + ((rank * 128 + warp * 32 + lane) * 4);
        d[0] = __uint_as_float(r[0]);
        d[1] = __uint_as_float(r[15]);
        d[2] = __uint_as_float(q[0]);
        d[3] = __uint_as_float(q[15]);
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
    float* dump; long long* cyc;
    cudaMalloc(&dump, 256 * 4 * 4); cudaMalloc(&cyc, 8);
    const int smem = 64 * 1024;
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    struct C { int ctas, M, N; } cs[] = {{1, 128, 208}, {1, 64, 208}, {2, 256, 208}, {2, 128, 208}, {2, 128, 104}};
    for (auto& c : cs) {
        for (int n_mma : {1, 201}) {
            cudaMemset(dump, 0xff, 256 * 16);
            if (c.ctas == 1) bench<1><<<1, 128, smem>>>(c.M, c.N, n_mma, dump, cyc);
            else {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, bench<2>, c.M, c.N, n_mma, dump, cyc);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            if (n_mma > 1) { printf("cta_group::%d M=%d N=%d: %.1f cycles per MMA (ideal %.0f)\n", c.ctas, c.M, c.N, (double)h / n_mma, (c.M < 128 ? 128.0 : (double)c.M / c.ctas) * c.N / 256.0 / 1.0 * (c.M / c.ctas >= 128 ? 1 : 1)); continue; }
            float hd[256 * 4];
            cudaMemcpy(hd, dump, sizeof(hd), cudaMemcpyDeviceToHost);
            printf("cta_group::%d M=%d N=%d  TMEM lane -> D row (col 0 | col 15 | col N-16 | col N-1):\n", c.ctas, c.M, c.N);
            for (int cta = 0; cta < c.ctas; ++cta) {
                printf("  CTA %d:", cta);
                for (int l = 0; l < 128; l += 8) printf(" [%d]=%g|%g|%g|%g", l, hd[(cta * 128 + l) * 4], hd[(cta * 128 + l) * 4 + 1], hd[(cta * 128 + l) * 4 + 2], hd[(cta * 128 + l) * 4 + 3]);
                printf("\n");
            }
        }
    }
    return 0;
}
