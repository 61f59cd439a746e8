// Which shared-memory word does tcgen05.mma kind::tf32 read for A[m][k] when the A operand is declared MN-major (no swizzle)?
// The A region (2048 floats) holds its own index (exact in tf32), B is K-major with B[n][k] = (n == k), so D[m][n] = the
// index of the word the hardware used as A[m][k=n].  One MMA: M=128, N=16, K=8.
#include <cstdint>
#include <cstdio>
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
constexpr int M = 128, N = 16;
__global__ void probe(float* D, int a_mn, int b_mn, uint32_t lbo, uint32_t sbo, int which) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    float* sA = (float*)smem;                // 2048 floats
    float* sB = (float*)(smem + 16 * 1024);  // 2048 floats
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    // which == 0: probe the A operand (A = index map, B = K-major identity [k/4][n][4]); which == 1: probe B (B = index map
    // over N=16 rows, A = K-major "identity": A[m][k] = (m % 8 == k), so D[m][n] = B_hw[n][m % 8])
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) { sA[i] = which == 0 ? (float)i : 0.f; sB[i] = which == 1 ? (float)i : 0.f; }
    __syncthreads();
    if (which == 0) { for (int k = threadIdx.x; k < 8; k += blockDim.x) sB[(k / 4) * N * 4 + k * 4 + k % 4] = 1.f; }
    else { for (int m = threadIdx.x; m < M; m += blockDim.x) { int k = m % 8; sA[(k / 4) * M * 4 + m * 4 + k % 4] = 1.f; } }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        if (a_mn) idesc |= (1u << 15);
        if (b_mn) idesc |= (1u << 16);
        uint64_t ad = (which == 0) ? make_desc(smem_u32(sA), lbo, sbo) : make_desc(smem_u32(sA), M * 16, 128);
        uint64_t bd = (which == 1) ? make_desc(smem_u32(sB), lbo, sbo) : make_desc(smem_u32(sB), N * 16, 128);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}
int main() {
    static float hD[M * N];
    float* dD;
    cudaMalloc(&dD, sizeof(hD));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    struct C { int which, mn; uint32_t lbo, sbo; } cs[] = {
        {0, 0, 2048, 128}, {0, 1, 128, 256}, {0, 1, 256, 128}, {0, 1, 512, 1024}, {0, 1, 1024, 512}, {0, 1, 2048, 16}, {0, 1, 16, 2048},
        {1, 0, 256, 128}, {1, 1, 128, 256}, {1, 1, 256, 128}, {1, 1, 512, 1024}, {1, 1, 1024, 512}};
    for (auto& c : cs) {
        cudaMemset(dD, 0, sizeof(hD));
        probe<<<1, 128, 32 * 1024>>>(dD, c.which == 0 ? c.mn : 0, c.which == 1 ? c.mn : 0, c.lbo, c.sbo, c.which);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        printf("=== operand %c %s lbo=%u sbo=%u : word index read for (row, k=0..7)\n", c.which ? 'B' : 'A', c.mn ? "MN-major" : "K-major", c.lbo, c.sbo);
        if (c.which == 0) {
            int rows[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 31, 32, 33, 63, 64, 65, 127};
            for (int m : rows) { printf("  m=%3d:", m); for (int k = 0; k < 8; ++k) printf(" %6.0f", hD[m * N + k]); printf("\n"); }
        } else {  // D[m][n] = B_hw[n][m%8]: print for n = 0..15 the eight k values (rows m = 0..7)
            for (int n = 0; n < 16; ++n) { printf("  n=%3d:", n); for (int k = 0; k < 8; ++k) printf(" %6.0f", hD[k * N + n]); printf("\n"); }
        }
    }
    return 0;
}
