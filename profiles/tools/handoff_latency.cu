// Microbenchmark: latency of the hand-offs a warp-specialised tcgen05 pipeline is built from (sm_100a, cluster of 2).
//   A  local mbarrier ping-pong between two warps, try_wait          (cycles per round trip = 2 hops)
//   B  same with a test_wait spin
//   C  remote ping-pong: CTA1 arrives on CTA0's barrier (mapa), CTA0 answers on CTA1's barrier, try_wait
//   D  fence.proxy.async after a shared store (cycles per fence)
//   E  one cta_group::2 MMA (M=256 N=208 K=16) + commit(multicast) -> wait in the issuing CTA, and in the peer CTA
//   F  13 such MMAs + commit -> wait
//   G  full hand-off loop of the kernels: [32 warps arrive on the leader barrier] -> issuer wakes -> 13 MMAs + commit ->
//      all warps of both CTAs wake -> arrive again ... (cycles per loop = the floor of one layer step of one tile slot)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/tools/handoff_bin profiles/tools/handoff_latency.cu && profiles/tools/handoff_bin  (the binary is git-ignored)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, a = smem_u32(bar);
    long long spins = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (++spins > (1ll << 26)) __trap();
    }
}
__device__ __forceinline__ void mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, a = smem_u32(bar);
    long long spins = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (++spins > (1ll << 26)) __trap();
    }
}
__device__ __forceinline__ void arrive_local(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void arrive_cta(uint64_t* bar, uint32_t cta) {
    asm volatile("{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\nmbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

constexpr int ITERS = 2000;
constexpr int EPI_WARPS = 16;

__global__ void __cluster_dims__(2, 1, 1) bench(long long* out, int spin_mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_init(&bars[3], 1);
        mbar_init(&bars[4], 1);               // commit target (E, F)
        mbar_init(&bars[5], 2 * EPI_WARPS);   // G: A_READY on the leader
        mbar_init(&bars[6], 1);               // G: D_FULL (commit, both CTAs)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;

    // ---- A / B: local ping-pong ------------------------------------------------------------------------------------
    for (int mode = 0; mode < 2; ++mode) {
        uint64_t* ba = &bars[mode * 2];
        uint64_t* bb = &bars[mode * 2 + 1];
        if (rank == 0 && lane == 0 && warp < 2) {
            long long t0 = clock64();
            for (int i = 0; i < ITERS; ++i) {
                if (warp == 0) {
                    arrive_local(ba);
                    if (mode == 0) mbar_try(bb, i & 1); else mbar_test(bb, i & 1);
                } else {
                    if (mode == 0) mbar_try(ba, i & 1); else mbar_test(ba, i & 1);
                    arrive_local(bb);
                }
            }
            if (warp == 0) out[mode] = (clock64() - t0) / ITERS;
        }
        __syncthreads();
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {  // fresh phases for the remote test
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    // ---- C: remote ping-pong: bars[0] lives in CTA0 (CTA1 arrives on it), bars[1] in CTA1 (CTA0 arrives on it) ------
    if (warp == 0 && lane == 0) {
        long long t0 = clock64();
        for (int i = 0; i < ITERS; ++i) {
            if (rank == 1) {
                arrive_cta(&bars[0], 0);
                if (spin_mode) mbar_test(&bars[1], i & 1); else mbar_try(&bars[1], i & 1);
            } else {
                if (spin_mode) mbar_test(&bars[0], i & 1); else mbar_try(&bars[0], i & 1);
                arrive_cta(&bars[1], 1);
            }
        }
        if (rank == 1) out[2] = (clock64() - t0) / ITERS;
    }
    __syncthreads();
    // ---- D: fence.proxy.async after a store --------------------------------------------------------------------------
    if (warp == 1) {
        long long t0 = clock64();
        for (int i = 0; i < ITERS; ++i) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_u32(smem + 150 * 1024 + threadIdx.x * 16)), "r"(i) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        if (lane == 0 && rank == 0) out[3] = (clock64() - t0) / ITERS;
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    // ---- E / F: MMA + commit -> wait -----------------------------------------------------------------------------------
    const int N = 208;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t ad0 = make_desc(smem_u32(smem), 128 * 16, 128), bd0 = make_desc(smem_u32(smem + 64 * 1024), 128, 26 * 128);
    for (int nm = 1; nm <= 13; nm += 12) {
        uint32_t ph = (nm == 1) ? 0 : (ITERS & 1);
        if (warp == 2 && lane == 0) {
            long long t0 = clock64();
            for (int i = 0; i < ITERS; ++i) {
                if (rank == 0) {
                    for (int k = 0; k < nm; ++k) mma2(tmem, ad0 + (uint64_t)(k * 256), bd0 + (uint64_t)(k * 16), idesc, k > 0);
                    commit2(&bars[4]);
                }
                if (spin_mode) mbar_test(&bars[4], ph); else mbar_try(&bars[4], ph);
                ph ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            out[(nm == 1 ? 4 : 6) + rank] = (clock64() - t0) / ITERS;
        }
        __syncthreads();
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    // ---- H: 26 MMAs from ONE warp vs 13 + 13 from TWO warps (different accumulators), each followed by a commit --------
    for (int two = 0; two < 2; ++two) {
        if (threadIdx.x == 0) {
            mbar_init(&bars[0], 1);
            mbar_init(&bars[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (rank == 0 && lane == 0 && (warp == 2 || (two && warp == 3))) {
            const int w = warp - 2;
            long long t0 = clock64();
            for (int i = 0; i < ITERS; ++i) {
                const int reps = two ? 1 : 2;
                for (int r = 0; r < reps; ++r)
                    for (int k = 0; k < 13; ++k) mma2(tmem + (uint32_t)((two ? w : r) * 208), ad0 + (uint64_t)(k * 256), bd0 + (uint64_t)(k * 16), idesc, k > 0);
                commit2(&bars[w]);
                mbar_try(&bars[w], i & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            out[10 + two * 2 + w] = (clock64() - t0) / ITERS;
        }
        __syncthreads();
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    // ---- G: full hand-off loop -----------------------------------------------------------------------------------------
    {
        if (warp < EPI_WARPS) {
            long long t0 = clock64();
            for (int i = 0; i < ITERS; ++i) {
                // "epilogue": nothing to convert; publish exactly like the kernels do
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) arrive_cta(&bars[5], 0);
                if (spin_mode) mbar_test(&bars[6], i & 1); else mbar_try(&bars[6], i & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (threadIdx.x == 0) out[8 + rank] = (clock64() - t0) / ITERS;
        } else if (warp == EPI_WARPS && rank == 0) {
            for (int i = 0; i < ITERS; ++i) {
                if (spin_mode) mbar_test(&bars[5], i & 1); else mbar_try(&bars[5], i & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    for (int k = 0; k < 13; ++k) mma2(tmem, ad0 + (uint64_t)(k * 256), bd0 + (uint64_t)(k * 16), idesc, k > 0);
                    commit2(&bars[6]);
                }
                __syncwarp();
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16 * sizeof(long long));
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int spin = 0; spin < 2; ++spin) {
        for (int i = 0; i < 16; ++i) out[i] = -1;
        bench<<<2, 32 * (EPI_WARPS + 1), 200 * 1024>>>(out, spin);
        cudaError_t e = cudaDeviceSynchronize();
        printf("spin_mode=%s (%s)\n", spin ? "test_wait" : "try_wait", cudaGetErrorString(e));
        printf("  A local ping-pong try_wait : %lld cycles / round trip\n", out[0]);
        printf("  B local ping-pong test_wait: %lld\n", out[1]);
        printf("  C remote ping-pong         : %lld\n", out[2]);
        printf("  D st.shared + fence.proxy.async: %lld per fence\n", out[3]);
        printf("  E 1 MMA + commit -> wait   : leader %lld, peer %lld\n", out[4], out[5]);
        printf("  F 13 MMAs + commit -> wait : leader %lld, peer %lld\n", out[6], out[7]);
        printf("  H 26 MMAs one warp: %lld ; 13+13 MMAs two warps: %lld / %lld per loop\n", out[10], out[12], out[13]);
        printf("  G publish(32 warps) -> 13 MMAs -> commit -> wake: leader %lld, peer %lld per loop\n", out[8], out[9]);
    }
    return 0;
}
