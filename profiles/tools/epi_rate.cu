// Microbenchmark: cycles per 16-element epilogue chunk (SiLU via tanh.approx.f32 + pack to fp16x2 + 2x STS.128)
// and per-instruction rates of the conversion used, on one SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t d;
    asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ uint32_t pack_h2_plain(float a, float b) {
    uint32_t d;
    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
template <int MODE>
__global__ void bench(uint32_t* out, long long* cyc, int iters) {
    __shared__ uint4 sbuf[2 * 512];
    float v[16];
    for (int j = 0; j < 16; ++j) v[j] = threadIdx.x * 0.001f + j * 0.1f;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t h[8];
        if (MODE == 0) {  // cvt satfinite only
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = pack_h2(v[2 * j], v[2 * j + 1]);
        } else if (MODE == 1) {  // cvt plain
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = pack_h2_plain(v[2 * j], v[2 * j + 1]);
        } else {  // full chunk: silu + pack (+ STS when MODE==3)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = v[2 * j], b = v[2 * j + 1], ha = 0.5f * a, hb = 0.5f * b, ta, tb;
                asm volatile("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(ha));
                asm volatile("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(hb));
                h[j] = pack_h2(fmaf(ha, ta, ha), fmaf(hb, tb, hb));
            }
        }
        if (MODE == 3) {
            sbuf[threadIdx.x] = make_uint4(h[0], h[1], h[2], h[3]);
            sbuf[512 + threadIdx.x] = make_uint4(h[4], h[5], h[6], h[7]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc ^= h[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += 0.001f;
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc + (MODE == 3 ? sbuf[(threadIdx.x + 1) % 512].x : 0);
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
void run(const char* name, uint32_t* out, long long* cyc) {
    for (int warps : {4, 8, 16}) {
        int iters = 2000;
        bench<MODE><<<1, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-22s warps/SMSP=%d : %.1f cycles per 16-element chunk per warp, %.1f per chunk per SMSP\n", name, warps / 4,
               (double)c / iters, (double)c / iters / (warps / 4));
    }
}
int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    run<0>("8x cvt.satfinite", out, cyc); run<1>("8x cvt plain", out, cyc);
    run<2>("silu chunk", out, cyc); run<3>("silu chunk + 2 STS.128", out, cyc);
    return 0;
}
