// K1 (fp32 tier): grouped per-net GEMMs for EnsembleLinearLayer / ParallelEnsembleLinearLayer
// (cmrl/models/layers.py:60-65,106-111) forward and backward, CUDA-core FFMA with fp32
// accumulation in ascending-k order (row results do not depend on tile placement, so the
// ragged selected-member rollout path is bit-identical to the dense all-member path).
//
// One register-tiled kernel, three operand layouts:
//   fwd : C[b,o] = act(sum_i (x*mask)[b,i] W[i,o] + bias[o])        A k-contig, B n-contig
//   dx  : C[b,i] = (sum_o dz[b,o] W[i,o]) * act'(z_prev[b,i])        A k-contig, B k-contig
//   dw  : C[i,o] = sum_b (x*mask)[b,i] dz[b,o];  db[o] = sum_b dz    A m-contig, B n-contig
// The tensor-core tier (tcgen05) lives in mlp_tc.cu.
#include "gemm_common.cuh"

namespace cmrl {

constexpr int BN = 64;
constexpr int TN = 4;
constexpr int NTHREADS = 256;
constexpr int PAD = 4;

// BK: K depth of one shared-memory tile.  Small grids (one wave of 64-row tiles, the reference batch size) are bound by the
// global-load round trip per K tile, so they run BK = 32 (half as many round trips); large ones BK = 16 with 128-row tiles.
template <int MODE, int BM, int BK>
__global__ void __launch_bounds__(NTHREADS) grouped_sgemm_kernel(const GemmP p) {
    constexpr int TM = BM / 16;
    constexpr bool A_KCONTIG = (MODE != MODE_DW);
    constexpr bool B_KCONTIG = (MODE == MODE_DX);
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    // ---- which net / row range ------------------------------------------------------------
    int pe_p, pe_e, m0, Mloc;
    long long row_base;  // row offset added to m for A and C addressing
    if (p.row_offsets != nullptr) {
        pe_p = blockIdx.z;
        int t = blockIdx.x, e = 0, found = 0;
        int lo = 0, hi = 0;
        for (e = 0; e < p.E; ++e) {
            lo = p.row_offsets[e];
            hi = p.row_offsets[e + 1];
            int nt = (hi - lo + BM - 1) / BM;
            if (t < nt) {
                found = 1;
                break;
            }
            t -= nt;
        }
        if (!found) return;
        pe_e = e;
        m0 = t * BM;
        Mloc = hi - lo;
        row_base = lo;
    } else {
        pe_p = blockIdx.z / p.E;
        pe_e = blockIdx.z % p.E;
        m0 = blockIdx.x * BM;
        Mloc = p.M;
        row_base = 0;
    }
    const int n0 = blockIdx.y * BN;
    const bool ragged = p.row_offsets != nullptr;
    const float* A = p.A + pe_p * p.a_sp + (ragged ? 0 : pe_e * p.a_se) + row_base * p.a_sm;
    const float* Am = p.Amask ? p.Amask + pe_p * p.am_sp + pe_e * p.am_se + (ragged ? row_base * p.am_sm : 0) : nullptr;
    const float* Bp = p.B + pe_p * p.b_sp + pe_e * p.b_se;
    const long long c_off = pe_p * p.c_sp + (ragged ? 0 : pe_e * p.c_se) + row_base * p.c_sm;

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;

    // ---- global->register staging -----------------------------------------------------------
    constexpr int A_PER = BM * BK / NTHREADS;
    constexpr int B_PER = BN * BK / NTHREADS;
    float ra[A_PER], rb[B_PER];

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int j = 0; j < A_PER; ++j) {
            int idx = tid + j * NTHREADS;
            int m, k;
            if (A_KCONTIG) {
                k = idx % BK;
                m = idx / BK;
            } else {
                m = idx % BM;
                k = idx / BM;
            }
            int gm = m0 + m, gk = k0 + k;
            float v = 0.0f;
            if (gm < Mloc && gk < p.K) {
                v = A[gm * p.a_sm + gk * p.a_sk];
                if (Am) v *= Am[gm * p.am_sm + gk * p.am_sk];
            }
            ra[j] = v;
        }
#pragma unroll
        for (int j = 0; j < B_PER; ++j) {
            int idx = tid + j * NTHREADS;
            int n, k;
            if (B_KCONTIG) {
                k = idx % BK;
                n = idx / BK;
            } else {
                n = idx % BN;
                k = idx / BN;
            }
            int gn = n0 + n, gk = k0 + k;
            rb[j] = (gn < p.N && gk < p.K) ? Bp[gk * p.b_sk + gn * p.b_sn] : 0.0f;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int j = 0; j < A_PER; ++j) {
            int idx = tid + j * NTHREADS;
            int m, k;
            if (A_KCONTIG) {
                k = idx % BK;
                m = idx / BK;
            } else {
                m = idx % BM;
                k = idx / BM;
            }
            As[buf][k][m] = ra[j];
        }
#pragma unroll
        for (int j = 0; j < B_PER; ++j) {
            int idx = tid + j * NTHREADS;
            int n, k;
            if (B_KCONTIG) {
                k = idx % BK;
                n = idx / BK;
            } else {
                n = idx % BN;
                k = idx / BN;
            }
            Bs[buf][k][n] = rb[j];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
    float bsum[TN] = {0.f, 0.f, 0.f, 0.f};

    const int nk = (p.K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
                a[i] = v.x;
                a[i + 1] = v.y;
                a[i + 2] = v.z;
                a[i + 3] = v.w;
            }
            {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN]);
                b[0] = v.x;
                b[1] = v.y;
                b[2] = v.z;
                b[3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            if (MODE == MODE_DW) {
#pragma unroll
                for (int j = 0; j < TN; ++j) bsum[j] += b[j];
            }
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue ------------------------------------------------------------------------------
    const long long g = (long long)pe_p * p.E + pe_e;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int gm = m0 + ty * TM + i;
        if (gm >= Mloc) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int gn = n0 + tx * TN + j;
            if (gn >= p.N) continue;
            long long off = c_off + (long long)gm * p.c_sm + gn;
            float v = acc[i][j];
            if (MODE == MODE_FWD) {
                if (p.bias) v += p.bias[g * p.N + gn];
                if (p.C2) p.C2[off] = v;
                p.C[off] = act_fwd(v, p.act);
            } else if (MODE == MODE_DX) {
                if (p.emul) v *= act_grad(p.emul[off], p.act);
                p.C[off] = v;
            } else {
                p.C[off] = p.accumulate ? p.C[off] + v : v;
            }
        }
    }
    if (MODE == MODE_DW) {
        if (p.bsum && blockIdx.x == 0 && ty == 0) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int gn = n0 + tx * TN + j;
                if (gn < p.N) {
                    long long off = g * p.N + gn;
                    p.bsum[off] = p.accumulate ? p.bsum[off] + bsum[j] : bsum[j];
                }
            }
        }
    }
}

template <int MODE>
static int launch_gemm(const GemmP& p, int total_rows, cudaStream_t st, const char* name, int tier = 0,
                       float* workspace = nullptr, long long workspace_floats = 0) {
    const bool ragged = p.row_offsets != nullptr;
    CMRL_REQUIRE(tier == 0 || tier == 1 || tier == 3, "%s: tier must be 0 (fp32 FFMA), 1 (TF32) or 3 (3xTF32)", name);
    if (tier != 0 && !ragged) {
        const int rc = launch_gemm_tc<MODE>(p, tier, st, name, workspace, workspace_floats);
        if (rc != -2) return rc;  // -2: shape not eligible for the tensor-core kernel
    }
    // small row counts -> 64-row tiles for more CTAs; large -> 128-row tiles
    const long long ctas128 = (long long)((total_rows + 127) / 128) * ((p.N + BN - 1) / BN) * (ragged ? p.P : p.P * p.E);
    const bool use64 = ctas128 < 2LL * num_sms();
    const int BMsel = use64 ? 64 : 128;
    dim3 grid;
    grid.x = (total_rows + BMsel - 1) / BMsel + (ragged ? p.E : 0);
    grid.y = (p.N + BN - 1) / BN;
    grid.z = ragged ? p.P : p.P * p.E;
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) return 0;
    CMRL_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "%s: too many nets/columns for one launch", name);
    if (use64)
        grouped_sgemm_kernel<MODE, 64, 32><<<grid, NTHREADS, 0, st>>>(p);
    else
        grouped_sgemm_kernel<MODE, 128, 16><<<grid, NTHREADS, 0, st>>>(p);
    CMRL_CHECK_LAUNCH(name);
    return 0;
}

}  // namespace cmrl

using namespace cmrl;

extern "C" int cmrl_ens_linear_fwd(const float* x, long long x_sp, long long x_se, const float* w, const float* b,
                                   float* y, float* z, int P, int E, int B, int in_size, int out_size,
                                   const float* mask, long long m_sp, long long m_se, long long m_sb,
                                   const int* row_offsets, int act, int tier, void* stream) {
    CMRL_REQUIRE(x && w && y, "ens_linear_fwd: null pointer");
    CMRL_REQUIRE(P >= 1 && E >= 1 && B >= 0 && in_size >= 1 && out_size >= 1, "ens_linear_fwd: bad dims");
    CMRL_REQUIRE(act >= 0 && act <= 2, "ens_linear_fwd: bad activation %d", act);
    if (B == 0) return 0;
    GemmP p{};
    p.A = x;
    p.a_sp = x_sp;
    p.a_se = x_se;
    p.a_sm = in_size;
    p.a_sk = 1;
    p.Amask = mask;
    p.am_sp = m_sp;
    p.am_se = m_se;
    p.am_sm = m_sb;
    p.am_sk = 1;
    p.B = w;
    p.b_sp = (long long)E * in_size * out_size;
    p.b_se = (long long)in_size * out_size;
    p.b_sk = out_size;
    p.b_sn = 1;
    p.C = y;
    p.C2 = z;
    p.c_sp = row_offsets ? (long long)B * out_size : (long long)E * B * out_size;
    p.c_se = (long long)B * out_size;
    p.c_sm = out_size;
    p.bias = b;
    p.row_offsets = row_offsets;
    p.M = B;
    p.N = out_size;
    p.K = in_size;
    p.P = P;
    p.E = E;
    p.act = act;
    return launch_gemm<MODE_FWD>(p, B, (cudaStream_t)stream, "ens_linear_fwd", tier);
}

extern "C" int cmrl_ens_linear_bwd_dx(const float* dz, const float* w, const float* z_prev, float* dx, int P, int E,
                                      int B, int in_size, int out_size, int act_prev, int tier, void* stream) {
    CMRL_REQUIRE(dz && w && dx, "ens_linear_bwd_dx: null pointer");
    CMRL_REQUIRE(P >= 1 && E >= 1 && B >= 0 && in_size >= 1 && out_size >= 1, "ens_linear_bwd_dx: bad dims");
    if (B == 0) return 0;
    GemmP p{};
    p.A = dz;
    p.a_sp = (long long)E * B * out_size;
    p.a_se = (long long)B * out_size;
    p.a_sm = out_size;
    p.a_sk = 1;
    p.B = w;  // B(k=o, n=i) = w[i*out + o]
    p.b_sp = (long long)E * in_size * out_size;
    p.b_se = (long long)in_size * out_size;
    p.b_sk = 1;
    p.b_sn = out_size;
    p.C = dx;
    p.c_sp = (long long)E * B * in_size;
    p.c_se = (long long)B * in_size;
    p.c_sm = in_size;
    p.emul = z_prev;
    p.M = B;
    p.N = in_size;
    p.K = out_size;
    p.P = P;
    p.E = E;
    p.act = z_prev ? act_prev : CMRL_ACT_IDENTITY;
    return launch_gemm<MODE_DX>(p, B, (cudaStream_t)stream, "ens_linear_bwd_dx", tier);
}

extern "C" int cmrl_ens_linear_bwd_dw(const float* x, long long x_sp, long long x_se, const float* mask,
                                      long long m_sp, long long m_se, long long m_sb, const float* dz, float* dw,
                                      float* db, int P, int E, int B, int in_size, int out_size, int accumulate,
                                      int tier, float* workspace, long long workspace_floats, void* stream) {
    CMRL_REQUIRE(x && dz && dw, "ens_linear_bwd_dw: null pointer");
    CMRL_REQUIRE(P >= 1 && E >= 1 && B >= 1 && in_size >= 1 && out_size >= 1, "ens_linear_bwd_dw: bad dims");
    GemmP p{};
    p.A = x;  // A(m=i, k=b) = x[b*in + i]
    p.a_sp = x_sp;
    p.a_se = x_se;
    p.a_sm = 1;
    p.a_sk = in_size;
    p.Amask = mask;
    p.am_sp = m_sp;
    p.am_se = m_se;
    p.am_sm = 1;
    p.am_sk = m_sb;
    p.B = dz;  // B(k=b, n=o)
    p.b_sp = (long long)E * B * out_size;
    p.b_se = (long long)B * out_size;
    p.b_sk = out_size;
    p.b_sn = 1;
    p.C = dw;
    p.c_sp = (long long)E * in_size * out_size;
    p.c_se = (long long)in_size * out_size;
    p.c_sm = out_size;
    p.bsum = db;
    p.M = in_size;
    p.N = out_size;
    p.K = B;
    p.P = P;
    p.E = E;
    p.accumulate = accumulate;
    return launch_gemm<MODE_DW>(p, in_size, (cudaStream_t)stream, "ens_linear_bwd_dw", tier, workspace, workspace_floats);
}
