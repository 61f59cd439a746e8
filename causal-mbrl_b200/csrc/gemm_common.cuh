// Operand description shared by the grouped GEMM kernels (FFMA tier in ens_linear.cu, tcgen05 TF32 tier in gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace cmrl {

struct GemmP {
    const float* A;
    long long a_sp, a_se, a_sm, a_sk;
    const float* Amask;  // optional multiplicative mask on A(m,k)
    long long am_sp, am_se, am_sm, am_sk;
    const float* B;
    long long b_sp, b_se, b_sk, b_sn;
    float* C;   // post-epilogue
    float* C2;  // optional pre-activation copy (fwd)
    long long c_sp, c_se, c_sm;
    const float* bias;  // [P,E,N]
    const float* emul;  // dx: z_prev, same layout as C
    float* bsum;        // dw: db [P,E,N]
    const int* row_offsets;
    int M, N, K, P, E, act, accumulate;
};

enum { MODE_FWD = 0, MODE_DX = 1, MODE_DW = 2 };

// tcgen05 kind::tf32 tier (gemm_tc.cu): passes = 1 (TF32, 1e-3 tier) or 3 (3xTF32 error-compensated, fp32-grade).
// Returns -2 when the shape is not eligible (caller falls back to the FFMA kernel).
template <int MODE>
int launch_gemm_tc(const GemmP& p, int passes, cudaStream_t st, const char* name, float* workspace = nullptr,
                   long long workspace_floats = 0);

}  // namespace cmrl
