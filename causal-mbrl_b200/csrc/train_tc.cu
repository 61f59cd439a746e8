// Fused training step on tcgen05 (kind::tf32, 3xTF32 error compensation = fp32-grade): base_dynamics.py:181-186 for one
// ensemble batch in THREE launches instead of ~25:
//
//   cmrl_train_pack_tf32   parameters (fp32 arena) -> per net a blob of K-chunked, K-major, hi/lo split B operands for every
//                          GEMM phase of the step (forward layers with the bias as an extra K row and the 0/1 input mask
//                          folded into W1; transposed weights for the dz chain)
//   cmrl_train_fused       one persistent CTA per (net, 128-row tile): forward through all layers, Gaussian head + NLL
//                          loss + their gradient, then the dz chain back to the first layer.  The activations never leave
//                          the SM as GEMM operands: the A operand of a phase is written by the previous phase's epilogue
//                          (tf32 hi part in shared memory, lo part in TENSOR MEMORY: hi + lo of a 128 x 208 fp32 tile do not
//                          fit shared memory next to a weight ring), weights stream through a TMA ring.  What the weight
//                          gradients need (layer inputs a_l, dz_l) is saved once in the operand's own K-major tile layout.
//   cmrl_train_dw          all weight / bias gradients of all layers in one launch: dW_l = a_{l-1}^T dz_l with the batch as
//                          K.  The saved tiles are K-major in the FEATURE direction, this GEMM contracts over ROWS (tf32
//                          MN-major operands read zeros on this part without a swizzle, profiles/r2_tf32_mn_probe.txt), so
//                          TMA-staged raw tiles are transposed + split in shared memory by 16 warps while the tensor pipe
//                          works on the previous 32-row slice.
//
// tf32 operands are TRUNCATED by the tensor core (profiles/r2_tf32_probe.txt): hi = v & 0xFFFFE000 is exact, lo = v - hi.
#include "tc_common.cuh"

namespace cmrl {
namespace tr {

using namespace tc;

constexpr int TF_ROWS = 128;
constexpr int TF_EPI_WARPS = 16;
constexpr int TF_WARP_W = 16, TF_WARP_MMA = 17;
constexpr int TF_THREADS = 18 * 32;
constexpr int TF_MAX_PH = 16;
constexpr int TF_MAX_LAYERS = 8;  // hidden layers + head
constexpr uint32_t TF_GROUP = TF_ROWS * 16;  // bytes of one 4-wide K group of a 128-row operand tile
constexpr uint32_t SMEM_LIMIT = 232448 - 1024;

enum { PH_FWD = 0, PH_HEAD = 1, PH_DX = 2 };

// Event trace of CTA (0,0) for profiles/tools/train_trace.py: compiled in only with -DCMRL_TRACE (a separate tool build of
// this file; the shipped library carries none of it).  Stamps are clock64 values in [role][1024] slots.
#ifdef CMRL_TRACE
__device__ long long* g_trace_buf = nullptr;
#define TRACE_DECL int trace_n = 0
#define TRACE(role)                                                                                                   \
    do {                                                                                                              \
        if (g_trace_buf != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && trace_n < 1024) \
            g_trace_buf[(role) * 1024 + trace_n++] = clock64();                                                       \
    } while (0)
#else
#define TRACE_DECL
#define TRACE(role)
#endif

// Geometry shared by the three kernels (built on the host by make_geom)
struct Geom {
    int L;            // hidden layers; layer L is the head
    int in0, H, Dn;   // inputs, hidden width, head outputs per net (PLAIN: D, PARALLEL: 1)
    int K0P, HP, NH;  // padded: inputs + bias slot (multiple of 16), hidden + bias slot, 2 * Dn
    int n_ph;
    int ph_kind[TF_MAX_PH], ph_layer[TF_MAX_PH], ph_k[TF_MAX_PH], ph_n[TF_MAX_PH];
    long long ph_off[TF_MAX_PH];  // float offset of the phase's chunks inside a net blob
    long long blob_floats;
    // saved tiles per (net, row tile), floats: x | act'(z_l), a_l (l < L) | dz_l (l <= L)
    long long off_x, off_z[TF_MAX_LAYERS], off_a[TF_MAX_LAYERS], off_dz[TF_MAX_LAYERS + 1], item_floats;
};

static int make_geom(Geom& g, int L, int in0, int H, int Dn) {
    CMRL_REQUIRE(L >= 1 && L + 1 <= TF_MAX_LAYERS && 2 * L + 1 <= TF_MAX_PH, "train_fused: 1..%d hidden layers", TF_MAX_LAYERS - 1);
    g.L = L; g.in0 = in0; g.H = H; g.Dn = Dn;
    g.K0P = (in0 + 1 + 15) / 16 * 16;
    g.HP = (H + 1 + 15) / 16 * 16;
    g.NH = (2 * Dn + 15) / 16 * 16;
    CMRL_REQUIRE(g.K0P <= g.HP, "train_fused: more inputs (%d) than the padded hidden width (%d)", in0, g.HP);
    CMRL_REQUIRE(g.NH <= 32, "train_fused: at most 16 head outputs per net (got %d)", Dn);
    CMRL_REQUIRE(2 * g.HP + g.NH <= 512, "train_fused: hidden width %d does not fit tensor memory (2 x %d + %d columns)", H, g.HP, g.NH);
    int n = 0;
    long long off = 0;
    auto add = [&](int kind, int layer, int k, int nn) {
        g.ph_kind[n] = kind; g.ph_layer[n] = layer; g.ph_k[n] = k; g.ph_n[n] = nn; g.ph_off[n] = off;
        off += (long long)(k / 16) * 32 * nn;  // chunk = 2 parts x 4 groups x nn rows x 4 floats
        ++n;
    };
    for (int l = 0; l < L; ++l) add(PH_FWD, l, l == 0 ? g.K0P : g.HP, g.HP);
    add(PH_HEAD, L, g.HP, g.NH);
    for (int l = L - 1; l >= 0; --l) add(PH_DX, l, l == L - 1 ? g.NH : g.HP, g.HP);  // produces dz_l through W_{l+1}^T
    g.n_ph = n;
    g.blob_floats = off;
    long long o = 0;
    g.off_x = o; o += (long long)g.K0P * TF_ROWS;
    for (int l = 0; l < L; ++l) { g.off_z[l] = o; o += (long long)g.HP * TF_ROWS; g.off_a[l] = o; o += (long long)g.HP * TF_ROWS; }
    for (int l = 0; l < L; ++l) { g.off_dz[l] = o; o += (long long)g.HP * TF_ROWS; }
    g.off_dz[L] = o; o += (long long)g.NH * TF_ROWS;
    g.item_floats = o;
    return 0;
}

// ================================================= pack ====================================================================
struct PackP {
    Geom g;
    const float* w[TF_MAX_LAYERS];  // [(P,)E,in,out]
    const float* b[TF_MAX_LAYERS];  // [(P,)E,1,out]
    const float* mask;              // layer-0 input mask [P,in] / [P,E,in] or NULL
    long long m_sp, m_se;
    int P, E;
    float* out;
};

// one thread per (net, phase, n, 4-wide k group): hi and lo float4
__global__ void train_pack_kernel(const PackP p) {
    const Geom& g = p.g;
    const int net = blockIdx.y, ph = blockIdx.z;
    const int K = g.ph_k[ph], N = g.ph_n[ph], kind = g.ph_kind[ph];
    const int l = kind == PH_DX ? g.ph_layer[ph] + 1 : g.ph_layer[ph];  // the layer whose weights this phase multiplies by
    const int in_l = l == 0 ? g.in0 : g.H, out_l = l == g.L ? 2 * g.Dn : g.H;
    const float* W = p.w[l] + (long long)net * in_l * out_l;
    const float* bias = p.b[l] + (long long)net * out_l;
    const int pp = net / p.E, e = net % p.E;
    const float* mk = (p.mask && l == 0 && kind != PH_DX) ? p.mask + pp * p.m_sp + e * p.m_se : nullptr;
    float* blob = p.out + (long long)net * g.blob_floats + g.ph_off[ph];
    const int groups = K / 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < groups * N; i += gridDim.x * blockDim.x) {
        const int n = i % N, kg = i / N;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = kg * 4 + j;
            float x = 0.f;
            if (kind == PH_DX) {  // B[n = input feature of layer l][k = output feature]: dz_l . W_l^T
                int ko = k;
                if (l == g.L) ko = (k & 1) ? g.Dn + (k >> 1) : (k >> 1);  // head columns are interleaved (mean_d, logvar_d)
                if (n < in_l && k < (l == g.L ? 2 * g.Dn : out_l)) x = W[(long long)n * out_l + ko];
            } else {  // B[n = output feature][k = input feature | bias slot]
                int no = n;
                if (l == g.L) no = (n & 1) ? g.Dn + (n >> 1) : (n >> 1);
                if (n < (l == g.L ? 2 * g.Dn : out_l)) {
                    if (k < in_l) x = W[(long long)k * out_l + no] * (mk ? mk[k] : 1.0f);
                    else if (k == in_l) x = bias[no];
                }
            }
            v[j] = x;
        }
        const int c = kg / 4, gi = kg % 4;
        float* dst = blob + (long long)c * 32 * N + ((long long)gi * N + n) * 4;
        uint32_t h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __float_as_uint(v[j]) & 0xFFFFE000u;
        *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3]));
        *reinterpret_cast<float4*>(dst + 16 * N) = make_float4(v[0] - __uint_as_float(h[0]), v[1] - __uint_as_float(h[1]),
                                                               v[2] - __uint_as_float(h[2]), v[3] - __uint_as_float(h[3]));
    }
}

// ================================================= fused forward + dz chain ================================================
struct TrainP {
    Geom g;
    const float* packed;
    int P, E, B, O, A, layout, act, residual, passes;
    const float *obs, *actn, *target;
    long long obs_se, act_se, tgt_se;  // member strides (floats); rows are dense [B,O] / [B,A] / [B,tgt_ld]
    int tgt_ld;
    const float *min_logvar, *max_logvar;
    int bounds_n;
    float reg, g_scale;
    float* loss;  // [E,B,Dtot] or NULL
    int loss_ld;  // Dtot
    float* scratch;
    int tiles_per_net;
    int stages;
    uint32_t stage_bytes, a_bytes, bar_off;
};

// bars: [0,S) W_FULL, [S,2S) W_EMPTY, 2S A_READY, 2S+1 D_FULL

__device__ __forceinline__ void split_store_chunk(uint32_t a_hi_row, uint32_t a_lo_t, int c, const float* v) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        hi[i] = __float_as_uint(v[i]) & 0xFFFFE000u;
        lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v3_sts128(a_hi_row + (uint32_t)(c * 4 + j) * TF_GROUP, hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
    tmem_st16(a_lo_t + (uint32_t)c * 16, lo);
}
// Saved tile layout: [row quarter q][col / 4][32 rows][4 floats].  Inside a quarter it is the K-major operand layout
// (a 4-column group of 32 rows is 512 contiguous bytes, so a warp's store is one contiguous 512-byte run), and a
// 32-row slice of any range of column groups is ONE contiguous block: the weight-gradient kernel fetches it with a single
// TMA bulk copy (512-byte copies, one per group, cost ~180 cycles each and made that kernel 10x slower).
__device__ __forceinline__ long long tile_off(int q, int lane, int ncg, int cg) { return ((long long)(q * ncg + cg) * 32 + lane) * 4; }
__device__ __forceinline__ void save_chunk(float* tile, int q, int lane, int ncg, int c, const float* v) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(tile + tile_off(q, lane, ncg, c * 4 + j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

__global__ void __launch_bounds__(TF_THREADS, 1) train_fused_kernel(const TrainP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const Geom& g = p.g;
    const int S = p.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
    uint64_t* bar_wfull = bars;
    uint64_t* bar_wempty = bars + S;
    uint64_t* bar_aready = bars + 2 * S;
    uint64_t* bar_dfull = bars + 2 * S + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = p.P * p.E * p.tiles_per_net;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&bar_wfull[s], 1);
            mbar_init(&bar_wempty[s], 1);
        }
        mbar_init(bar_aready, TF_EPI_WARPS);
        mbar_init(bar_dfull, 1);
        fence_barrier_init();
    }
    if (warp == TF_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t col_alo = (uint32_t)g.HP, col_head = (uint32_t)(2 * g.HP);

    if (warp == TF_WARP_W) {
        // ===================== weight producer: one bulk copy per 16-deep K chunk (hi | lo) ================================
        if (lane == 0) {
            uint32_t n = 0;
            TRACE_DECL;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const float* blob = p.packed + (long long)(item / p.tiles_per_net) * g.blob_floats;
                for (int ph = 0; ph < g.n_ph; ++ph) {
                    const uint32_t bytes = 128u * (uint32_t)g.ph_n[ph];
                    const float* src = blob + g.ph_off[ph];
                    const int nch = g.ph_k[ph] / 16;
                    for (int c = 0; c < nch; ++c, ++n) {
                        const uint32_t s = n % (uint32_t)S;
                        if (n >= (uint32_t)S) mbar_wait_sleep(&bar_wempty[s], ((n / (uint32_t)S) - 1) & 1, 32);
                        TRACE(0);
                        mbar_expect_tx(&bar_wfull[s], bytes);
                        tma_bulk_g2s(smem + p.a_bytes + s * p.stage_bytes, src + (long long)c * 32 * g.ph_n[ph], bytes, &bar_wfull[s]);
                    }
                }
            }
        }
    } else if (warp == TF_WARP_MMA) {
        // ===================== MMA issuer =====================================================================================
        uint32_t n = 0, n_a = 0;
        TRACE_DECL;
        const uint32_t a_hi0 = smem_u32(smem);
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            for (int ph = 0; ph < g.n_ph; ++ph) {
                const int N = g.ph_n[ph];
                const uint32_t idesc = umma_idesc_tf32(TF_ROWS, N);
                const uint32_t d_t = tmem + (g.ph_kind[ph] == PH_HEAD ? col_head : 0u);
                const uint32_t part = 64u * (uint32_t)N;  // bytes of one (hi or lo) part of a weight chunk: 4 groups x N x 16
                const uint32_t b_lbo = 16u * (uint32_t)N;
                const int nch = g.ph_k[ph] / 16;
                TRACE(2);
                mbar_wait(bar_aready, n_a & 1);
                ++n_a;
                tc_fence_after();
                TRACE(2);
#pragma unroll 1
                for (int c = 0; c < nch; ++c, ++n) {
                    const uint32_t s = n % (uint32_t)S;
                    mbar_wait(&bar_wfull[s], (n / (uint32_t)S) & 1);
                    if (c == 0 || c == nch - 1) TRACE(2);
                    const uint32_t w_hi = smem_u32(smem + p.a_bytes + s * p.stage_bytes), w_lo = w_hi + part;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t ad = umma_desc(a_hi0 + (uint32_t)(c * 4 + ks * 2) * TF_GROUP, TF_GROUP, 128);
                        const uint64_t bh = umma_desc(w_hi + (uint32_t)ks * 2 * b_lbo, b_lbo, 128);
                        const uint32_t first = (c == 0 && ks == 0) ? 0u : 1u;
                        if (p.passes == 3) {  // small terms first
                            const uint64_t bl = umma_desc(w_lo + (uint32_t)ks * 2 * b_lbo, b_lbo, 128);
                            tf32_mma_ts_uni(d_t, tmem + col_alo + (uint32_t)(c * 16 + ks * 8), bh, idesc, first);
                            tf32_mma_ss_uni(d_t, ad, bl, idesc, 1u);
                            tf32_mma_ss_uni(d_t, ad, bh, idesc, 1u);
                        } else {
                            tf32_mma_ss_uni(d_t, ad, bh, idesc, first);
                        }
                    }
                    umma_commit_uni(&bar_wempty[s]);
                }
                umma_commit_uni(bar_dfull);
                TRACE(2);
            }
        }
    } else {
        // ===================== epilogue warps: TMEM lane quarter q, column set cset =============================================
        const int q = warp & 3, cset = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t a_hi_row = smem_u32(smem) + (uint32_t)row * 16;
        const uint32_t a_lo_t = tmem + lane_off + col_alo;
        const int act = p.act;
        const int Dn = g.Dn;
        uint32_t n_d = 0;
        TRACE_DECL;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int net = item / p.tiles_per_net, tile = item % p.tiles_per_net;
            const int pp = net / p.E, e = net % p.E;
            const int grow = tile * TF_ROWS + row;
            const bool valid = grow < p.B;
            const long long rr = valid ? grow : 0;
            float* sc = p.scratch + (long long)item * g.item_floats;
            // ---- layer-0 operand: [obs | action | 1 | 0...] -------------------------------------------------------------------
            if (cset == 0) {
                const float* po = p.obs + e * p.obs_se + rr * p.O;
                const float* pa = p.actn + e * p.act_se + rr * p.A;
#pragma unroll 1
                for (int c = 0; c < g.K0P / 16; ++c) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int k = c * 16 + i;
                        float x = 0.f;
                        if (k < p.O) x = valid ? po[k] : 0.f;
                        else if (k < g.in0) x = valid ? pa[k - p.O] : 0.f;
                        else if (k == g.in0) x = 1.0f;
                        v[i] = x;
                    }
                    save_chunk(sc + g.off_x, q, lane, g.K0P / 4, c, v);
                    split_store_chunk(a_hi_row, a_lo_t, c, v);
                }
            }
            v5_st_wait();
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready);

            for (int ph = 0; ph < g.n_ph; ++ph) {
                const int kind = g.ph_kind[ph], l = g.ph_layer[ph];
                // prefetch what the head needs before the accumulator is ready
                if (warp == 0) TRACE(1);
                mbar_wait(bar_dfull, n_d & 1);
                ++n_d;
                tc_fence_after();
                if (warp == 0) TRACE(1);
                if (kind == PH_FWD) {
                    float* zt = sc + g.off_z[l];
                    float* at = sc + g.off_a[l];
                    const int nchunks = g.HP / 16;
#pragma unroll 1
                    for (int c = cset; c < nchunks; c += 4) {
                        uint32_t raw[16];
                        float z[16];
                        __syncwarp();
                        v5_ld16_issue(tmem + lane_off + (uint32_t)c * 16, raw);
                        v3_ld_wait();
                        float gz[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {  // activation and its derivative from one exp + one reciprocal
                            const float zi = __uint_as_float(raw[i]);
                            if (act == CMRL_ACT_SILU) {
                                const float sg = __fdividef(1.0f, 1.0f + __expf(-zi));
                                z[i] = zi * sg;
                                gz[i] = sg * (1.0f + zi * (1.0f - sg));
                            } else if (act == CMRL_ACT_RELU) {
                                z[i] = fmaxf(zi, 0.0f);
                                gz[i] = zi > 0.0f ? 1.0f : 0.0f;
                            } else {
                                z[i] = zi;
                                gz[i] = 1.0f;
                            }
                            if (c * 16 + i == g.H) z[i] = 1.0f;  // the next layer's bias input
                        }
                        save_chunk(zt, q, lane, g.HP / 4, c, gz);  // act'(z): what the dz chain multiplies by
                        save_chunk(at, q, lane, g.HP / 4, c, z);
                        split_store_chunk(a_hi_row, a_lo_t, c, z);
                    }
                } else if (kind == PH_DX) {
                    const float* zt = sc + g.off_z[l];
                    float* dt = sc + g.off_dz[l];
                    const int nchunks = g.HP / 16;
#pragma unroll 1
                    for (int c = cset; c < nchunks; c += 4) {
                        uint32_t raw[16];
                        float z[16];
                        __syncwarp();
                        v5_ld16_issue(tmem + lane_off + (uint32_t)c * 16, raw);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 t4 = *reinterpret_cast<const float4*>(zt + tile_off(q, lane, g.HP / 4, c * 4 + j));
                            z[4 * j] = t4.x; z[4 * j + 1] = t4.y; z[4 * j + 2] = t4.z; z[4 * j + 3] = t4.w;
                        }
                        v3_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) z[i] *= __uint_as_float(raw[i]);  // dz = (dz_next . W^T) * act'(z)
                        save_chunk(dt, q, lane, g.HP / 4, c, z);
                        if (l > 0) split_store_chunk(a_hi_row, a_lo_t, c, z);
                    }
                } else if (cset == 0) {
                    // ---- Gaussian head + NLL + their gradient, one output per iteration (columns 2d = mean_d, 2d+1 = logvar_d)
                    float* dt = sc + g.off_dz[g.L];
                    const uint32_t h_t = tmem + lane_off + col_head;
#pragma unroll 1
                    for (int d = 0; d < g.NH / 2; ++d) {
                        float dm = 0.f, dl = 0.f;
                        uint32_t r0, r1;
                        __syncwarp();
                        tmem_ld2(h_t + (uint32_t)(2 * d), r0, r1);
                        v3_ld_wait();
                        if (d < Dn) {
                            const int od = p.layout == CMRL_HEAD_PARALLEL ? pp : d;
                            const int bi = p.bounds_n == 1 ? 0 : od;
                            float m = __uint_as_float(r0);
                            const float r = __uint_as_float(r1);
                            if (p.residual) m += p.obs[e * p.obs_se + rr * p.O + od];
                            const float t = p.target[e * p.tgt_se + rr * p.tgt_ld + od];
                            const float mn = p.min_logvar[bi], mx = p.max_logvar[bi];
                            float lv1;
                            const float lv = soft_clamp_logvar(r, mn, mx, &lv1);
                            const float diff = m - t, iv = expf(-lv);
                            if (valid) {
                                if (p.loss) p.loss[((long long)e * p.B + grow) * p.loss_ld + od] = diff * diff * iv + lv + p.reg;
                                const float s_lo = (lv1 - mn) > 20.0f ? 1.0f : sigmoidf_(lv1 - mn);
                                const float s_hi = (mx - r) > 20.0f ? 1.0f : sigmoidf_(mx - r);
                                dm = p.g_scale * 2.0f * diff * iv;
                                dl = p.g_scale * (1.0f - diff * diff * iv) * s_lo * s_hi;
                            }
                        }
                        const int k = 2 * d;
                        *reinterpret_cast<float2*>(dt + tile_off(q, lane, g.NH / 4, k / 4) + (k & 3)) = make_float2(dm, dl);
                        const uint32_t hm = __float_as_uint(dm) & 0xFFFFE000u, hl = __float_as_uint(dl) & 0xFFFFE000u;
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_hi_row + (uint32_t)(k / 4) * TF_GROUP + (uint32_t)(k & 3) * 4), "r"(hm), "r"(hl) : "memory");
                        tmem_st2(a_lo_t + (uint32_t)k, __float_as_uint(dm - __uint_as_float(hm)), __float_as_uint(dl - __uint_as_float(hl)));
                    }
                }
                if (warp == 0) TRACE(1);
                if (ph + 1 < g.n_ph) {  // the last phase (dz of the first layer) feeds no further GEMM
                    v5_st_wait();
                    tc_fence_before();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_aready);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TF_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ================================================= weight gradients ========================================================
constexpr int DW_THREADS = 18 * 32;
constexpr int DW_NT = 112;          // output columns per CTA
constexpr int DW_SUB = 32;          // rows (= K) per pipeline step
constexpr uint32_t DW_KG = 144;     // bytes between K groups of 4 rows in the transposed operands (128 + 16: bank spread)
constexpr uint32_t DW_MG = 8 * DW_KG;  // bytes between groups of 8 features
constexpr uint32_t DW_RAW_A = 32 * 512, DW_RAW_B = (DW_NT / 4) * 512, DW_RAW = DW_RAW_A + DW_RAW_B;
constexpr uint32_t DW_OP_A = 16 * DW_MG, DW_OP_B = (DW_NT / 8) * DW_MG, DW_OP = 2 * (DW_OP_A + DW_OP_B);
constexpr int DW_RAW_STAGES = 2, DW_OP_STAGES = 2;
constexpr uint32_t DW_BAR_OFF = DW_RAW_STAGES * DW_RAW + DW_OP_STAGES * DW_OP;
constexpr uint32_t DW_SMEM = DW_BAR_OFF + 128;
static_assert(DW_SMEM <= SMEM_LIMIT, "dw kernel shared memory");

struct DwP {
    Geom g;
    const float* scratch;
    int P, E, tiles_per_net;
    int in_l[TF_MAX_LAYERS], out_l[TF_MAX_LAYERS];
    float* dw[TF_MAX_LAYERS];
    float* db[TF_MAX_LAYERS];
    const float* mask;  // layer-0 input mask or NULL
    long long m_sp, m_se;
    int work_off[TF_MAX_LAYERS + 1];  // prefix sums of (m tiles x n tiles) per layer
    int n_tiles[TF_MAX_LAYERS];
    int ksplit, tiles_per_slice;
    float* ws;  // [ksplit][nets][ext_total] partial sums (ksplit > 1)
    long long ext_off[TF_MAX_LAYERS], ext_total;
};

// element (i = input feature or bias row, n = packed output column) of layer l's extended gradient -> its destination
__device__ __forceinline__ void dw_store(const DwP& p, int net, int l, int i, int n, float v) {
    const Geom& g = p.g;
    const int in_l = p.in_l[l], out_l = p.out_l[l];
    if (i > in_l) return;
    int o = n;
    if (l == g.L) {
        if (n >= 2 * g.Dn) return;
        o = (n & 1) ? g.Dn + (n >> 1) : (n >> 1);
    } else if (n >= out_l) {
        return;
    }
    if (i == in_l) {
        p.db[l][(long long)net * out_l + o] = v;
    } else {
        if (l == 0 && p.mask) v *= p.mask[(net / p.E) * p.m_sp + (net % p.E) * p.m_se + i];
        p.dw[l][((long long)net * in_l + i) * out_l + o] = v;
    }
}

__global__ void __launch_bounds__(DW_THREADS, 1) train_dw_kernel(const DwP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const Geom& g = p.g;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DW_BAR_OFF);
    uint64_t* bar_rfull = bars;                      // [2] tx
    uint64_t* bar_rempty = bars + 2;                 // [2] 16 warps
    uint64_t* bar_ofull = bars + 4;                  // [2] 16 warps
    uint64_t* bar_oempty = bars + 6;                 // [2] commit
    uint64_t* bar_acc = bars + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int net = blockIdx.y;
    // ---- which (layer, feature tile, column tile, batch slice) ------------------------------------------------------------
    const int slice = blockIdx.x % p.ksplit;
    const int w = blockIdx.x / p.ksplit;
    int l = 0;
    while (w >= p.work_off[l + 1]) ++l;
    const int wi = w - p.work_off[l];
    const int mt = wi / p.n_tiles[l], nt = wi % p.n_tiles[l];
    const int a_cols = l == 0 ? g.K0P : g.HP;            // saved input tile of layer l: x or a_{l-1}
    const int z_cols = l == g.L ? g.NH : g.HP;           // dz_l
    const long long a_off = l == 0 ? g.off_x : g.off_a[l - 1];
    const long long z_off = g.off_dz[l];
    const int ga = min(32, a_cols / 4 - mt * 32);         // 4-feature groups of this CTA's A rows
    const int NT = min(DW_NT, z_cols - nt * DW_NT);      // output columns of this CTA
    const int gb = NT / 4;
    const int t0 = slice * p.tiles_per_slice, t1 = min(p.tiles_per_net, t0 + p.tiles_per_slice);
    const int n_sub = max(0, t1 - t0) * (TF_ROWS / DW_SUB);

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_rfull[s], 1);
            mbar_init(&bar_rempty[s], TF_EPI_WARPS);
            mbar_init(&bar_ofull[s], TF_EPI_WARPS);
            mbar_init(&bar_oempty[s], 1);
        }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == TF_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // rows of the transposed operands that no feature group fills must not hold NaN patterns: zero the operand stages once
    for (uint32_t i = threadIdx.x; i < DW_OP_STAGES * DW_OP / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem + DW_RAW_STAGES * DW_RAW)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const float* base = p.scratch + ((long long)net * p.tiles_per_net) * g.item_floats;

    if (warp == TF_WARP_W) {
        // ===================== raw tile producer: one bulk copy per operand and 32-row slice ================================
        TRACE_DECL;
        for (int i = 0; i < n_sub; ++i) {
            const int s = i & 1;
            const int t = t0 + i / 4, qr = i % 4;
            const float* item = base + (long long)t * g.item_floats;
            if (lane == 0) {
                if (i >= 2) mbar_wait_sleep(&bar_rempty[s], ((i >> 1) - 1) & 1, 32);
                TRACE(0);
                mbar_expect_tx(&bar_rfull[s], (uint32_t)(ga + gb) * 512u);
                uint8_t* dst = smem + s * DW_RAW;
                tma_bulk_g2s(dst, item + a_off + ((long long)(qr * (a_cols / 4) + mt * 32) * 32) * 4, (uint32_t)ga * 512u, &bar_rfull[s]);
                tma_bulk_g2s(dst + DW_RAW_A, item + z_off + ((long long)(qr * (z_cols / 4) + nt * (DW_NT / 4)) * 32) * 4, (uint32_t)gb * 512u,
                             &bar_rfull[s]);
            }
        }
    } else if (warp == TF_WARP_MMA) {
        // ===================== MMA issuer: A = features x rows, B = columns x rows, both K-major after the transpose =============
        const uint32_t idesc = umma_idesc_tf32(TF_ROWS, NT);
        TRACE_DECL;
        for (int i = 0; i < n_sub; ++i) {
            const int s = i & 1;
            mbar_wait(&bar_ofull[s], (i >> 1) & 1);
            tc_fence_after();
            TRACE(2);
            const uint32_t op = smem_u32(smem + DW_RAW_STAGES * DW_RAW + s * DW_OP);
            const uint32_t a_hi = op, a_lo = op + DW_OP_A, b_hi = op + 2 * DW_OP_A, b_lo = b_hi + DW_OP_B;
#pragma unroll
            for (int ks = 0; ks < DW_SUB / 8; ++ks) {
                const uint32_t ko = (uint32_t)ks * 2 * DW_KG;
                const uint64_t ah = umma_desc(a_hi + ko, DW_KG, DW_MG), bh = umma_desc(b_hi + ko, DW_KG, DW_MG);
                const uint32_t first = (i == 0 && ks == 0) ? 0u : 1u;
                const uint64_t al = umma_desc(a_lo + ko, DW_KG, DW_MG), bl = umma_desc(b_lo + ko, DW_KG, DW_MG);
                tf32_mma_ss_uni(tmem, al, bh, idesc, first);
                tf32_mma_ss_uni(tmem, ah, bl, idesc, 1u);
                tf32_mma_ss_uni(tmem, ah, bh, idesc, 1u);
            }
            umma_commit_uni(&bar_oempty[s]);
        }
        umma_commit_uni(bar_acc);
    } else {
        // ===================== transposers, then the epilogue ====================================================================
        const int tid = threadIdx.x;  // 0..511
        TRACE_DECL;
        for (int i = 0; i < n_sub; ++i) {
            const int s = i & 1;
            if (warp == 0) TRACE(1);
            mbar_wait(&bar_rfull[s], (i >> 1) & 1);
            if (warp == 0) TRACE(1);
            if (i >= 2) mbar_wait(&bar_oempty[s], ((i >> 1) - 1) & 1);
            if (warp == 0) TRACE(1);
            const uint8_t* raw = smem + s * DW_RAW;
            const uint32_t op = smem_u32(smem + DW_RAW_STAGES * DW_RAW + s * DW_OP);
            for (int idx = tid; idx < (ga + gb) * 32; idx += TF_EPI_WARPS * 32) {
                const int grp = idx >> 5, r = idx & 31;  // r == lane
                const bool isa = grp < ga;
                const int gl = isa ? grp : grp - ga;
                const float4 v = *reinterpret_cast<const float4*>(raw + (isa ? 0u : DW_RAW_A) + (uint32_t)gl * 512u + (uint32_t)r * 16u);
                const uint32_t hi_base = op + (isa ? 0u : 2 * DW_OP_A) + (uint32_t)(gl >> 1) * DW_MG + (uint32_t)(r >> 2) * DW_KG +
                                         (uint32_t)(gl & 1) * 64u + (uint32_t)(r & 3) * 4u;
                const uint32_t lo_base = hi_base + (isa ? DW_OP_A : DW_OP_B);
                const float e4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t h = __float_as_uint(e4[j]) & 0xFFFFE000u;
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(hi_base + (uint32_t)j * 16u), "r"(h) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(lo_base + (uint32_t)j * 16u), "r"(__float_as_uint(e4[j] - __uint_as_float(h))) : "memory");
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_ofull[s]);
                mbar_arrive(&bar_rempty[s]);
            }
            if (warp == 0) TRACE(1);
        }
        // ---- epilogue: TMEM lane = feature row of this tile, columns = output features --------------------------------------
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        if (warp == 0) TRACE(1);
        const int q = warp & 3, cset = warp >> 2;
        const int i_row = mt * TF_ROWS + q * 32 + lane;
        float* wsp = p.ksplit > 1 ? p.ws + ((long long)slice * gridDim.y + net) * p.ext_total + p.ext_off[l] : nullptr;
        const int zc = l == g.L ? g.NH : p.out_l[l];  // columns of the extended gradient in the workspace
        for (int c = cset; c < NT / 16; c += 4) {
            uint32_t rawv[16];
            __syncwarp();
            v5_ld16_issue(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c * 16, rawv);
            v3_ld_wait();
            if (n_sub == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) rawv[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = nt * DW_NT + c * 16 + j;
                if (wsp) {
                    if (i_row <= p.in_l[l] && n < zc) wsp[(long long)i_row * zc + n] = __uint_as_float(rawv[j]);
                } else {
                    dw_store(p, net, l, i_row, n, __uint_as_float(rawv[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TF_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

// sums the batch-slice partials in slice order (deterministic) and scatters them to the gradient tensors
__global__ void train_dw_reduce_kernel(const DwP p, int nets) {
    const Geom& g = p.g;
    const long long total = (long long)nets * p.ext_total;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int net = (int)(idx / p.ext_total);
        const long long r = idx % p.ext_total;
        int l = 0;
        while (l < g.L && r >= p.ext_off[l + 1]) ++l;
        const int zc = l == g.L ? g.NH : p.out_l[l];
        const int i = (int)((r - p.ext_off[l]) / zc), n = (int)((r - p.ext_off[l]) % zc);
        float acc = 0.f;
        for (int s = 0; s < p.ksplit; ++s) acc += p.ws[((long long)s * nets + net) * p.ext_total + r];
        dw_store(p, net, l, i, n, acc);
    }
}

}  // namespace tr
}  // namespace cmrl

using namespace cmrl;
using namespace cmrl::tr;

#ifdef CMRL_TRACE
extern "C" int cmrl_trace_set(long long* buf) {
    CMRL_CUDA(cudaMemcpyToSymbol(cmrl::tr::g_trace_buf, &buf, sizeof(buf)));
    return 0;
}
#endif

extern "C" long long cmrl_train_fused_sizes(int L, int in_size, int hidden, int head_outputs, int P, int E, int B, long long* blob_floats,
                                            long long* scratch_floats, long long* dw_ws_floats) {
    Geom g;
    if (make_geom(g, L, in_size, hidden, head_outputs) != 0) return -1;
    const long long nets = (long long)P * E, tiles = (B + TF_ROWS - 1) / TF_ROWS;
    if (blob_floats) *blob_floats = g.blob_floats;
    if (scratch_floats) *scratch_floats = nets * tiles * g.item_floats;
    if (dw_ws_floats) {
        long long ext = 0;
        for (int l = 0; l <= L; ++l) ext += (long long)((l == 0 ? in_size : hidden) + 1) * (l == L ? g.NH : hidden);
        *dw_ws_floats = ext * nets;  // per batch slice
    }
    return 0;
}

extern "C" int cmrl_train_pack_tf32(const float* const* weights, const float* const* biases, int L, int P, int E, int in_size, int hidden,
                                    int head_outputs, const float* mask, long long m_sp, long long m_se, float* packed, void* stream) {
    PackP p{};
    if (make_geom(p.g, L, in_size, hidden, head_outputs) != 0) return -1;
    CMRL_REQUIRE(weights && biases && packed, "train_pack_tf32: null pointer");
    for (int l = 0; l <= L; ++l) {
        CMRL_REQUIRE(weights[l] && biases[l], "train_pack_tf32: null layer pointer");
        p.w[l] = weights[l];
        p.b[l] = biases[l];
    }
    p.mask = mask; p.m_sp = m_sp; p.m_se = m_se; p.P = P; p.E = E; p.out = packed;
    CMRL_REQUIRE(P * E <= 65535, "train_pack_tf32: too many nets");
    const int per = (p.g.HP / 4) * p.g.HP;
    dim3 grid((per + 255) / 256, P * E, p.g.n_ph);
    train_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    CMRL_CHECK_LAUNCH("train_pack_tf32");
    return 0;
}

extern "C" int cmrl_train_fused(const float* packed, int L, int P, int E, int B, int O, int A, int hidden, int head_dim, int layout, int act,
                                int residual, int passes, const float* obs, long long obs_se, const float* actn, long long act_se,
                                const float* target, long long tgt_se, int tgt_ld, const float* min_logvar, const float* max_logvar,
                                int bounds_n, float reg, float g_scale, float* loss, float* scratch, void* stream) {
    TrainP p{};
    const int Dn = layout == CMRL_HEAD_PARALLEL ? 1 : head_dim;
    if (make_geom(p.g, L, O + A, hidden, Dn) != 0) return -1;
    CMRL_REQUIRE(packed && obs && actn && target && min_logvar && max_logvar && scratch, "train_fused: null pointer");
    CMRL_REQUIRE(layout == CMRL_HEAD_PLAIN || layout == CMRL_HEAD_PARALLEL, "train_fused: bad head layout");
    CMRL_REQUIRE(layout == CMRL_HEAD_PLAIN ? P == 1 : P == head_dim, "train_fused: P must be 1 (plain head) or head_dim (parallel head)");
    CMRL_REQUIRE(passes == 1 || passes == 3, "train_fused: passes must be 1 (TF32) or 3 (3xTF32)");
    CMRL_REQUIRE(B >= 1 && E >= 1, "train_fused: empty batch");
    CMRL_REQUIRE(bounds_n == 1 || bounds_n == head_dim, "train_fused: logvar bounds must have 1 or head_dim entries");
    CMRL_REQUIRE(act == CMRL_ACT_IDENTITY || act == CMRL_ACT_RELU || act == CMRL_ACT_SILU, "train_fused: unsupported activation");
    p.packed = packed; p.P = P; p.E = E; p.B = B; p.O = O; p.A = A; p.layout = layout; p.act = act; p.residual = residual; p.passes = passes;
    p.obs = obs; p.actn = actn; p.target = target; p.obs_se = obs_se; p.act_se = act_se; p.tgt_se = tgt_se; p.tgt_ld = tgt_ld;
    p.min_logvar = min_logvar; p.max_logvar = max_logvar; p.bounds_n = bounds_n; p.reg = reg; p.g_scale = g_scale;
    p.loss = loss; p.loss_ld = head_dim; p.scratch = scratch;
    p.tiles_per_net = (B + TF_ROWS - 1) / TF_ROWS;
    p.a_bytes = (uint32_t)(p.g.HP / 4) * TF_GROUP;
    p.stage_bytes = 128u * (uint32_t)p.g.HP;
    int stages = (int)((SMEM_LIMIT - 256 - p.a_bytes) / p.stage_bytes);
    if (stages > 6) stages = 6;
    CMRL_REQUIRE(stages >= 2, "train_fused: hidden width %d leaves no room for a weight ring", hidden);
    p.stages = stages;
    p.bar_off = p.a_bytes + (uint32_t)stages * p.stage_bytes;
    const uint32_t smem_bytes = p.bar_off + (2 * stages + 2) * 8;
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(train_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_once.done(attr_dev);
    }
    const long long items = (long long)P * E * p.tiles_per_net;
    CMRL_REQUIRE(items < (1ll << 30), "train_fused: too many row tiles");
    int grid = (int)(items < num_sms() ? items : num_sms());
    train_fused_kernel<<<grid, TF_THREADS, smem_bytes, (cudaStream_t)stream>>>(p);
    CMRL_CHECK_LAUNCH("train_fused");
    return 0;
}

extern "C" int cmrl_train_dw(const float* scratch, int L, int P, int E, int B, int in_size, int hidden, int head_outputs,
                             float* const* dw, float* const* db, const float* mask, long long m_sp, long long m_se, float* workspace,
                             long long workspace_floats, void* stream) {
    DwP p{};
    if (make_geom(p.g, L, in_size, hidden, head_outputs) != 0) return -1;
    CMRL_REQUIRE(scratch && dw && db, "train_dw: null pointer");
    const Geom& g = p.g;
    p.scratch = scratch; p.P = P; p.E = E; p.mask = mask; p.m_sp = m_sp; p.m_se = m_se;
    p.tiles_per_net = (B + TF_ROWS - 1) / TF_ROWS;
    int work = 0;
    long long ext = 0;
    for (int l = 0; l <= L; ++l) {
        CMRL_REQUIRE(dw[l] && db[l], "train_dw: null gradient pointer");
        p.dw[l] = dw[l]; p.db[l] = db[l];
        p.in_l[l] = l == 0 ? in_size : hidden;
        p.out_l[l] = l == L ? 2 * head_outputs : hidden;
        const int z_cols = l == L ? g.NH : g.HP;
        const int m_tiles = (p.in_l[l] + 1 + TF_ROWS - 1) / TF_ROWS;
        p.n_tiles[l] = (z_cols + DW_NT - 1) / DW_NT;
        p.work_off[l] = work;
        work += m_tiles * p.n_tiles[l];
        p.ext_off[l] = ext;
        ext += (long long)(p.in_l[l] + 1) * (l == L ? g.NH : p.out_l[l]);
    }
    p.work_off[L + 1] = work;
    p.ext_total = ext;
    const int nets = P * E;
    CMRL_REQUIRE(nets <= 65535, "train_dw: too many nets");
    // batch slices: enough CTAs for ~2 per SM, at least one row tile each, bounded by the workspace
    int ksplit = 1;
    if (p.tiles_per_net > 1 && workspace != nullptr) {
        const long long ctas = (long long)work * nets;
        long long want = (2LL * num_sms() + ctas - 1) / ctas;
        if (want > p.tiles_per_net) want = p.tiles_per_net;
        const long long fit = workspace_floats / (ext * nets);
        if (want > fit) want = fit;
        if (want > 64) want = 64;
        if (want >= 2) ksplit = (int)want;
    }
    p.tiles_per_slice = (p.tiles_per_net + ksplit - 1) / ksplit;
    ksplit = (p.tiles_per_net + p.tiles_per_slice - 1) / p.tiles_per_slice;
    p.ksplit = ksplit;
    p.ws = ksplit > 1 ? workspace : nullptr;
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(train_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_once.done(attr_dev);
    }
    dim3 grid(work * ksplit, nets);
    train_dw_kernel<<<grid, DW_THREADS, DW_SMEM, (cudaStream_t)stream>>>(p);
    CMRL_CHECK_LAUNCH("train_dw");
    if (ksplit > 1) {
        const long long total = ext * nets;
        int nb = (int)((total + 255) / 256);
        if (nb > 8 * num_sms()) nb = 8 * num_sms();
        train_dw_reduce_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(p, nets);
        CMRL_CHECK_LAUNCH("train_dw.reduce");
    }
    return 0;
}
