// Probe for the fused training kernels (train_tc.cu): what does tcgen05.mma kind::tf32 do on B200 with
//   T1  K-major no-swizzle operands whose fp32 values carry all 24 mantissa bits  -> truncation or rounding to tf32?
//   T2  MN-major no-swizzle operands (the [mn/4][k][4] layout: what a K-major [feature/4][row][4] tile is when the
//       GEMM contracts over ROWS, i.e. the weight-gradient GEMM reads saved activations without a transpose)
//   T3  the A operand in tensor memory (lane = row, one 32-bit column per k)
// One CTA, M=128, N=32, K=16 (two MMAs).  Prints the max abs error against host results under each hypothesis.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
constexpr int M = 128, N = 32, K = 16;
// mode 0: K-major both; 1: MN-major both (desc lbo = 8-row K group stride, sbo = MN group stride); 2: same with lbo/sbo swapped;
// 3: A in tensor memory, B K-major
__global__ void probe(const float* A, const float* B, float* D, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    float* sA = (float*)smem;                // 128 x 16 floats = 8 KB
    float* sB = (float*)(smem + 16 * 1024);  // 32 x 16 floats
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tslot;
    if (mode == 0 || mode == 3) {  // K-major: [k/4][row][4]
        for (int i = threadIdx.x; i < M * K; i += blockDim.x) { int m = i / K, k = i % K; sA[(k / 4) * M * 4 + m * 4 + k % 4] = A[i]; }
        for (int i = threadIdx.x; i < N * K; i += blockDim.x) { int n = i / K, k = i % K; sB[(k / 4) * N * 4 + n * 4 + k % 4] = B[i]; }
    } else {  // MN-major: [mn/4][k][4]
        for (int i = threadIdx.x; i < M * K; i += blockDim.x) { int m = i / K, k = i % K; sA[(m / 4) * K * 4 + k * 4 + m % 4] = A[i]; }
        for (int i = threadIdx.x; i < N * K; i += blockDim.x) { int n = i / K, k = i % K; sB[(n / 4) * K * 4 + k * 4 + n % 4] = B[i]; }
    }
    if (mode == 3) {  // A -> TMEM columns 64.. : lane = row, column = k
        uint32_t r[16];
        const int m = warp * 32 + lane;
        for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(A[m * K + k]);
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 64;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]),
                     "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
                     "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                     : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        if (mode == 1 || mode == 2) idesc |= (1u << 15) | (1u << 16);
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t ad, bd;
            if (mode == 0 || mode == 3) {
                ad = make_desc(smem_u32(sA) + ks * 2 * M * 16, M * 16, 128);
                bd = make_desc(smem_u32(sB) + ks * 2 * N * 16, N * 16, 128);
            } else if (mode == 1) {
                ad = make_desc(smem_u32(sA) + ks * 128, 128, K * 16);
                bd = make_desc(smem_u32(sB) + ks * 128, 128, K * 16);
            } else {
                ad = make_desc(smem_u32(sA) + ks * 128, K * 16, 128);
                bd = make_desc(smem_u32(sB) + ks * 128, K * 16, 128);
            }
            if (mode == 3)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem), "r"(tmem + 64 + ks * 8), "l"(bd), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(ks > 0)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; c += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}
static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float round_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
int main() {
    static float hA[M * K], hB[N * K], hD[M * N];
    srand(1);
    for (auto& v : hA) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : hB) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    const char* names[] = {"K-major smem", "MN-major smem (lbo = K-group stride, sbo = MN-group stride)", "MN-major smem (lbo/sbo swapped)", "A in TMEM, B K-major"};
    for (int mode = 0; mode < 4; ++mode) {
        cudaMemset(dD, 0xff, sizeof(hD));
        probe<<<1, 128, 32 * 1024>>>(dA, dB, dD, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: error %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        double et = 0, er = 0, ef = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double st = 0, sr = 0, sf = 0;
                for (int k = 0; k < K; ++k) {
                    st += (double)trunc_tf32(hA[m * K + k]) * trunc_tf32(hB[n * K + k]);
                    sr += (double)round_tf32(hA[m * K + k]) * round_tf32(hB[n * K + k]);
                    sf += (double)hA[m * K + k] * hB[n * K + k];
                }
                et = fmax(et, fabs(hD[m * N + n] - st)); er = fmax(er, fabs(hD[m * N + n] - sr)); ef = fmax(ef, fabs(hD[m * N + n] - sf));
            }
        printf("mode %d %-62s max|D - trunc| %.3e  max|D - round| %.3e  max|D - fp32| %.3e\n", mode, names[mode], et, er, ef);
    }
    return 0;
}
