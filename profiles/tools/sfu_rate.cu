// Microbenchmark: cycles per warp-level MUFU instruction on one SM sub-partition (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/sfu_rate profiles/tools/sfu_rate.cu && /tmp/sfu_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__device__ __forceinline__ float op(float x) {
    float y;
    if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else if (OP == 4) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    else y = fmaf(x, 1.0001f, 0.5f);
    return y;
}
template <int OP>
__global__ void bench(float* out, long long* cyc, int iters) {
    float v[8];
    for (int j = 0; j < 8; ++j) v[j] = threadIdx.x * 0.001f + j;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = op<OP>(v[j]);
    long long t1 = clock64();
    float s = 0;
    for (int j = 0; j < 8; ++j) s += v[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP>
void run(const char* name, float* out, long long* cyc) {
    for (int warps : {4, 8, 16}) {  // 1, 2, 4 warps per sub-partition
        int iters = 2000;
        bench<OP><<<1, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double per = (double)c / (iters * 8.0 * (warps / 4));
        printf("%-6s warps/SMSP=%d : %.2f cycles per warp-instruction per SMSP\n", name, warps / 4, per);
    }
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    run<0>("tanh", out, cyc); run<1>("ex2", out, cyc); run<2>("rcp", out, cyc); run<3>("rsqrt", out, cyc);
    run<4>("sin", out, cyc); run<5>("ffma", out, cyc);
    return 0;
}
