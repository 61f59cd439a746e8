// Microbenchmark: does tcgen05.ld (epilogue reading accumulator buffer A) slow down tcgen05.mma accumulating into
// buffer B, and what is the tcgen05.ld throughput?  sm_100a, cta_group::1, M=128, N=208, no-swizzle K-major operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_contention profiles/tools/tmem_contention.cu
#include <cstdio>
#include <cstdint>
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
// CTAS = 1: cta_group::1 (M=128); CTAS = 2: cluster of 2, cta_group::2 (M=256, B split by N across the pair)
// mode bit0: run MMAs, bit1: run LDTM loops in `ld_warps` warps, bit2: LDTM loops also do 16 MUFU + 2 STS.128 per chunk
template <int CTAS>
__global__ void bench(int mode, int ld_warps, int n_mma, int n_ld, long long* out, int m_total) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
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
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CTAS == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    const int N = 208;
    if (warp == 16) {  // MMA issuer warp
        if (lane == 0 && (mode & 1) && rank == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(m_total >> 4) << 24);
            const int rows_cta = m_total / CTAS;
            const uint64_t ad0 = make_desc(smem_u32(smem), rows_cta * 16, 128), bd0 = make_desc(smem_u32(smem + 64 * 1024), 128, 26 * 128);
            long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                const int ks = i % 13;
                if (CTAS == 1)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + 256), "l"(ad0 + (uint64_t)(ks * (rows_cta * 32 / 16))), "l"(bd0 + (uint64_t)(ks * 16)), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
                else
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + 256), "l"(ad0 + (uint64_t)(ks * (rows_cta * 32 / 16))), "l"(bd0 + (uint64_t)(ks * 16)), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
            }
            if (CTAS == 1)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
            mbar_wait(&bar, 0);
            out[0] = clock64() - t0;
        }
    } else if (warp < ld_warps && (mode & 2)) {
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t acc = 0;
        long long t0 = clock64();
        for (int i = 0; i < n_ld; ++i) {
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                           "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr + (uint32_t)((i % 13) * 16)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (mode & 4) {
                uint32_t h[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float a = __uint_as_float(r[2 * j]) * 0.5f, b = __uint_as_float(r[2 * j + 1]) * 0.5f, ta, tb;
                    asm volatile("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
                    asm volatile("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
                    asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(fmaf(b, tb, b)), "f"(fmaf(a, ta, a)));
                }
                uint8_t* dst = smem + 128 * 1024 + (i % 8) * 4096 + (warp & 3) * 512 + lane * 16;
                *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(dst + 2048) = make_uint4(h[4], h[5], h[6], h[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc ^= r[j];
            }
        }
        long long t1 = clock64();
        if (lane == 0 && rank == 0) out[1 + warp] = t1 - t0;
        if (acc == 0x12345) out[40] = acc;
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
    cudaMalloc(&out, 64 * 8);
    const int smem = 170 * 1024;
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int n_mma = 520, n_ld = 520;
    struct Cfg { int mode, warps; const char* name; } cfgs[] = {
        {1, 0, "MMA only"}, {2, 4, "LDTM only, 1 warp/SMSP"}, {2, 8, "LDTM only, 2 warps/SMSP"}, {2, 16, "LDTM only, 4 warps/SMSP"},
        {3, 4, "MMA + LDTM 1 warp/SMSP"}, {3, 16, "MMA + LDTM 4 warps/SMSP"},
        {6, 16, "epilogue chunk (LDTM+SiLU+STS) 4 warps/SMSP, no MMA"}, {7, 16, "MMA + epilogue chunk 4 warps/SMSP"},
        {7, 8, "MMA + epilogue chunk 2 warps/SMSP"}, {7, 4, "MMA + epilogue chunk 1 warp/SMSP"}};
    for (int ctas : {1, 2, 3})  // 3 = cta_group::2 with M=128 (64 rows per CTA)
    for (auto& c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(out, 0, 64 * 8);
            if (ctas == 1) bench<1><<<1, 17 * 32, smem>>>(c.mode, c.warps, n_mma, n_ld, out, 128);
            else {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(2); cfg.blockDim = dim3(17 * 32); cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, bench<2>, c.mode, c.warps, n_mma, n_ld, out, ctas == 3 ? 128 : 256);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long h[64];
        cudaMemcpy(h, out, 64 * 8, cudaMemcpyDeviceToHost);
        double ld = 0; int nw = 0;
        for (int w = 0; w < c.warps; ++w) { ld += h[1 + w]; nw++; }
        printf("cta_group::%d M=%d %-52s: MMA %.1f cyc/MMA", ctas == 1 ? 1 : 2, ctas == 1 ? 128 : (ctas == 3 ? 128 : 256), c.name, (double)h[0] / n_mma);
        if (nw) printf(" | chunk %.1f cyc per warp, %.1f per SMSP", ld / nw / n_ld, ld / nw / n_ld / (c.warps / 4));
        printf("\n");
    }
    return 0;
}
