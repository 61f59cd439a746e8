// Tensor-core tier: the whole selected-member MLP of one mechanism (4 x (EnsembleLinear + act) + Gaussian
// head + sampling + scatter; cmrl/models/layers.py:60-65, plain_transition.py:100-122,
// external_mask_transition.py:141-169, fake_env.py:195-215) fused into ONE persistent kernel on tcgen05:
//
//   * operands fp16 (11-bit significand = TF32's; kind::f16), accumulation fp32 in TMEM;
//   * weights pre-packed (cmrl_pack_layer_f16) into the UMMA K-major no-swizzle core-matrix layout with the
//     bias folded in as an extra K row (the A operand carries a 1.0 column) and the causal input mask folded
//     into W1, so a layer's B operand is one contiguous blob staged by a single TMA bulk copy (UBLKCP);
//   * activations never leave the SM: TMEM accumulator -> registers (tcgen05.ld) -> SiLU/ReLU -> fp16 ->
//     shared memory in the A-operand layout of the next layer;
//   * the head epilogue applies the soft-clamped log-variance, the residual, the supplied N(0,1) noise and
//     scatters next-obs / reward / terminal to the env-ordered output.
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lanes 32w..32w+31 = rows of the 128-row tile),
// warp 4 MMA issuer + TMEM allocator, warp 5 TMA producer.
#include "tc_common.cuh"

namespace cmrl {
namespace tc {

// work item -> (p, member e, first sorted row, rows in tile); items enumerate p-major, then member tiles
__device__ __forceinline__ bool decode_item(const RolloutP& p, int item, int tiles_total, const int* s_off, int& pp,
                                            int& e, int& row0, int& nrows) {
    if (item >= p.P * tiles_total) return false;
    pp = item / tiles_total;
    int t = item - pp * tiles_total;
    for (e = 0; e < p.E; ++e) {
        int cnt = s_off[e + 1] - s_off[e];
        int nt = (cnt + TILE_M - 1) / TILE_M;
        if (t < nt) {
            row0 = s_off[e] + t * TILE_M;
            nrows = min(TILE_M, cnt - t * TILE_M);
            return true;
        }
        t -= nt;
    }
    return false;
}

__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_rollout_f16_kernel(const RolloutP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int L = p.num_layers;
    uint8_t* w_buf[2] = {smem, smem + p.w_stage_bytes};
    uint8_t* a_buf = smem + 2 * p.w_stage_bytes;
    uint8_t* tail = a_buf + p.a_chunks * A_CHUNK_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);  // [0,1] w_full, [2,3] w_empty, [4] a_ready, [5] d_full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 64);
    int* s_off = reinterpret_cast<int*>(tail + 96);  // [E+1]

    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + 2;
    uint64_t* a_ready = bars + 4;
    uint64_t* d_full = bars + 5;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i <= p.E; i += blockDim.x) s_off[i] = p.offsets[i];
    if (threadIdx.x == 0) {
        mbar_init(&w_full[0], 1);
        mbar_init(&w_full[1], 1);
        mbar_init(&w_empty[0], 1);
        mbar_init(&w_empty[1], 1);
        mbar_init(a_ready, 4);  // one elected arrive per epilogue warp
        mbar_init(d_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int tiles_total = 0;
    for (int e = 0; e < p.E; ++e) tiles_total += (s_off[e + 1] - s_off[e] + TILE_M - 1) / TILE_M;

    if (warp == 5) {
        // ===================== TMA producer: one contiguous blob per (item, layer) =====================
        if (lane == 0) {
            uint32_t n_load = 0;
            for (int item = blockIdx.x;; item += gridDim.x) {
                int pp, e, row0, nrows;
                if (!decode_item(p, item, tiles_total, s_off, pp, e, row0, nrows)) break;
                const __half* blob = p.packed + (long long)(pp * p.E + e) * p.blob_halfs;
                for (int l = 0; l <= L; ++l, ++n_load) {
                    const int s = n_load & 1;
                    mbar_wait(&w_empty[s], ((n_load >> 1) & 1) ^ 1);
                    const uint32_t bytes = (uint32_t)p.layer_n[l] * p.layer_k[l] * 2;
                    mbar_expect_tx(&w_full[s], bytes);
                    tma_bulk_g2s(w_buf[s], blob + p.layer_off[l], bytes, &w_full[s]);
                }
            }
        }
    } else if (warp == 4) {
        // ===================== MMA issuer (one elected lane) ============================================
        if (lane == 0) {
            uint32_t n_load = 0, n_a = 0;
            for (int item = blockIdx.x;; item += gridDim.x) {
                int pp, e, row0, nrows;
                if (!decode_item(p, item, tiles_total, s_off, pp, e, row0, nrows)) break;
                for (int l = 0; l <= L; ++l, ++n_load, ++n_a) {
                    const int s = n_load & 1;
                    mbar_wait(a_ready, n_a & 1);                 // A operand of this layer is in smem, D is drained
                    mbar_wait(&w_full[s], (n_load >> 1) & 1);    // weights landed
                    tc_fence_after();
                    const int K = p.layer_k[l], Nn = p.layer_n[l];
                    const uint32_t idesc = umma_idesc_f16(TILE_M, Nn);
                    const uint32_t d_tmem = tmem_base + (l == L ? HEAD_TMEM_COL : 0);
                    const uint32_t a_addr = smem_u32(a_buf), b_addr = smem_u32(w_buf[s]);
                    const uint32_t b_sbo = (uint32_t)(K / 8) * 128;
                    for (int ks = 0; ks < K / 16; ++ks) {
                        uint64_t adesc = umma_desc(a_addr + ks * 2 * A_CHUNK_BYTES, A_CHUNK_BYTES, 128);
                        uint64_t bdesc = umma_desc(b_addr + ks * 256, 128, b_sbo);
                        umma_f16(d_tmem, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
                    }
                    umma_commit(&w_empty[s]);  // weights stage free once these MMAs retire
                    umma_commit(d_full);       // accumulator ready for the epilogue
                }
            }
        }
    } else {
        // ===================== epilogue warps: rows = TMEM lanes =========================================
        const int r = threadIdx.x;  // row within the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        uint32_t n_d = 0;
        for (int item = blockIdx.x;; item += gridDim.x) {
            int pp, e, row0, nrows;
            if (!decode_item(p, item, tiles_total, s_off, pp, e, row0, nrows)) break;
            const bool valid = r < nrows;
            const int env = valid ? p.perm[row0 + r] : -1;
            // ---- layer-0 A operand: [obs | act | 1 | 0...] in fp16, K chunks of 8 --------------------------
            {
                const int in = p.O + p.A;
                const int K0 = p.layer_k[0];
                for (int c = 0; c < K0 / 8; ++c) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int k = c * 8 + j;
                        float x = 0.0f;
                        if (valid) {
                            if (k < p.O) x = p.obs[(long long)env * p.O + k];
                            else if (k < in) x = p.actn[(long long)env * p.A + (k - p.O)];
                            else if (k == in) x = 1.0f;
                        }
                        v[j] = x;
                    }
                    uint4 q = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                    *reinterpret_cast<uint4*>(a_buf + c * A_CHUNK_BYTES + r * 16) = q;
                }
            }
            tc_fence_before();    // earlier tcgen05.ld of this thread are ordered before the MMA that reuses TMEM
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);

            for (int l = 0; l < L; ++l, ++n_d) {
                mbar_wait(d_full, n_d & 1);
                tc_fence_after();
                const int Nn = p.layer_n[l];
                for (int c = 0; c < Nn / 16; ++c) {
                    float v[16];
                    tmem_ld16(tmem_base + lane_base + c * 16, v);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        v[j] = act_fast(v[j], p.act);
                        if (c * 16 + j == p.H) v[j] = 1.0f;  // the ones column that carries the next layer's bias
                    }
                    uint4 q0 = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                    uint4 q1 = make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
                    *reinterpret_cast<uint4*>(a_buf + (2 * c) * A_CHUNK_BYTES + r * 16) = q0;
                    *reinterpret_cast<uint4*>(a_buf + (2 * c + 1) * A_CHUNK_BYTES + r * 16) = q1;
                }
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_ready);
            }
            // ---- head: Gaussian head + residual + sampling + scatter (fake_env.py:205-214) -------------------
            mbar_wait(d_full, n_d & 1);
            ++n_d;
            tc_fence_after();
            const int D = p.head_dim;
            const int Nh = p.layer_n[L];
            if (p.layout == CMRL_HEAD_PARALLEL) {
                float v[16];
                tmem_ld16(tmem_base + lane_base + HEAD_TMEM_COL, v);
                if (valid) {
                    float m = v[0];
                    if (p.residual) m += p.obs[(long long)env * p.O + pp];
                    float out = m;
                    if (p.gaussian && p.noise) {
                        int bi = p.bounds_n == 1 ? 0 : pp;
                        float lv = soft_clamp_logvar(v[1], p.min_logvar[bi], p.max_logvar[bi]);
                        out = m + sqrtf(expf(lv)) * p.noise[(long long)env * D + pp];
                    }
                    p.pred[(long long)env * D + pp] = out;
                }
            } else {
                // PLAIN: packed head columns are interleaved (2d = mean_d, 2d+1 = raw logvar_d; deterministic: d).
                for (int c = 0; c < Nh / 16; ++c) {
                    float v[16];
                    tmem_ld16(tmem_base + lane_base + HEAD_TMEM_COL + c * 16, v);
                    if (p.gaussian) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const int d = (c * 16 + j) >> 1;
                            if (valid && d < D) {
                                float m = v[j];
                                if (p.residual) m += p.obs[(long long)env * p.O + d];
                                if (p.noise) {
                                    const int bi = p.bounds_n == 1 ? 0 : d;
                                    const float lv = soft_clamp_logvar(v[j + 1], p.min_logvar[bi], p.max_logvar[bi]);
                                    m += sqrtf(expf(lv)) * p.noise[(long long)env * D + d];
                                }
                                p.pred[(long long)env * D + d] = m;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int d = c * 16 + j;
                            if (valid && d < D) {
                                float m = v[j];
                                if (p.residual) m += p.obs[(long long)env * p.O + d];
                                p.pred[(long long)env * D + d] = m;
                            }
                        }
                    }
                }
            }
        }
    }

    // ---- teardown ---------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ----------------------------------------------------------------------------------------------- packing
// One layer of every net: fp32 W [G,in,out] (+ bias [G,1,out], + 0/1 mask on the input rows) -> fp16 B operand
// [n_pad rows (out)][k_pad (in + bias slot)] in the K-major core-matrix layout:
//   half index = (n/8)*(k_pad*8) + (k/8)*64 + (n%8)*8 + (k%8)
__global__ void pack_layer_f16_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                      const float* __restrict__ mask, long long m_sp, long long m_se, int E, int in,
                                      int out, int k_pad, int n_pad, int interleave_d, float scale,
                                      __half* __restrict__ dst, long long net_stride) {
    const int g = blockIdx.y;
    const long long total = (long long)n_pad * k_pad;
    const float* wg = w + (long long)g * in * out;
    const float* mg = mask ? mask + (g / E) * m_sp + (g % E) * m_se : nullptr;
    __half* d = dst + g * net_stride;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // iterate in destination order (coalesced stores)
        int kk = (int)(i % 8);
        int nn = (int)((i / 8) % 8);
        long long rest = i / 64;
        int kc = (int)(rest % (k_pad / 8));
        int ng = (int)(rest / (k_pad / 8));
        int n = ng * 8 + nn, k = kc * 8 + kk;
        float v = 0.0f;
        if (n < out) {
            // Gaussian PLAIN heads are stored with (mean_d, logvar_d) in adjacent columns (2d, 2d+1)
            const int src = interleave_d ? ((n & 1) ? interleave_d + (n >> 1) : (n >> 1)) : n;
            if (k < in) {
                v = wg[(long long)k * out + src];
                if (mg) v *= mg[k];
            } else if (k == in && b) {
                v = b[(long long)g * out + src];
            }
        }
        v = fminf(fmaxf(v * scale, -65504.0f), 65504.0f);
        d[i] = __float2half_rn(v);
    }
}

}  // namespace tc
}  // namespace cmrl

namespace cmrl {
namespace tc {
int launch_rollout_v3(const RolloutP& p, cudaStream_t st);
int launch_rollout_v4(const RolloutP& p, cudaStream_t st);
int launch_rollout_v5(const RolloutP& p, cudaStream_t st);
int launch_rollout_v6(const RolloutP& p, cudaStream_t st);
}
}  // namespace cmrl

using namespace cmrl;

static long long* g_dbg_buffer = nullptr;
static int g_dbg_flags = 0;
extern "C" void cmrl_debug_set_flags(int f) { g_dbg_flags = f; }
// profiling hook (not part of the public header): device buffer [grid][16] of per-role cycle counters, or NULL
extern "C" void cmrl_debug_set_buffer(void* buf) { g_dbg_buffer = (long long*)buf; }

extern "C" int cmrl_pack_layer_f16(const float* w, const float* b, const float* mask, long long m_sp, long long m_se,
                                   int P, int E, int in_size, int out_size, int k_pad, int n_pad, int interleave_d,
                                   float scale, void* dst_half, long long net_stride_halfs, void* stream) {
    CMRL_REQUIRE(w && dst_half, "pack_layer_f16: null pointer");
    CMRL_REQUIRE(k_pad % 16 == 0 && n_pad % 16 == 0 && k_pad >= in_size + 1 && n_pad >= out_size,
                 "pack_layer_f16: k_pad must be a multiple of 16 >= in+1, n_pad a multiple of 16 >= out");
    CMRL_REQUIRE(P * E <= 65535, "pack_layer_f16: too many nets");
    CMRL_REQUIRE(interleave_d == 0 || out_size == 2 * interleave_d, "pack_layer_f16: interleave_d must be out_size / 2");
    long long total = (long long)n_pad * k_pad;
    int bx = (int)((total + 255) / 256);
    if (bx > 1024) bx = 1024;
    dim3 grid(bx, P * E);
    tc::pack_layer_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, b, mask, m_sp, m_se, E, in_size, out_size, k_pad,
                                                                    n_pad, interleave_d, scale, (__half*)dst_half, net_stride_halfs);
    CMRL_CHECK_LAUNCH("pack_layer_f16");
    return 0;
}

extern "C" int cmrl_mlp_rollout_f16(const void* packed, long long blob_halfs, const int* layer_off_halfs,
                                    const int* layer_k, const int* layer_n, int num_layers, int P, int E, int N, int O,
                                    int A, int H, int head_dim, int layout, int gaussian, int act, const float* obs,
                                    const float* actn, const int* perm, const int* offsets, int residual,
                                    const float* min_logvar, const float* max_logvar, int bounds_n, const float* noise,
                                    float* pred, int variant, void* stream) {
    CMRL_REQUIRE(packed && obs && perm && offsets && pred, "mlp_rollout_f16: null pointer");
    CMRL_REQUIRE(num_layers >= 1 && num_layers <= 7, "mlp_rollout_f16: 1..7 hidden layers supported");
    CMRL_REQUIRE(E >= 1 && E <= 64, "mlp_rollout_f16: E must be <= 64");
    CMRL_REQUIRE(!gaussian || (min_logvar && max_logvar), "mlp_rollout_f16: gaussian head needs logvar bounds");
    if (N == 0) return 0;
    tc::RolloutP p{};
    p.packed = (const __half*)packed;
    p.blob_halfs = blob_halfs;
    const int HP = layer_n[0];
    for (int l = 0; l <= num_layers; ++l) {
        p.layer_off[l] = layer_off_halfs[l];
        p.layer_k[l] = layer_k[l];
        p.layer_n[l] = layer_n[l];
        CMRL_REQUIRE(layer_k[l] % 16 == 0 && layer_n[l] % 16 == 0 && layer_n[l] <= (variant == 6 ? 512 : 256) && layer_n[l] >= 16,
                     "mlp_rollout_f16: layer %d dims must be multiples of 16 with N <= 256 (512 for variant 6)", l);
        if (variant != 6) {
            if (l >= 1) CMRL_REQUIRE(layer_k[l] == HP, "mlp_rollout_f16: hidden width must be uniform");
            if (l < num_layers) CMRL_REQUIRE(layer_n[l] == HP, "mlp_rollout_f16: hidden width must be uniform");
        }
    }
    CMRL_REQUIRE(variant == 6 || H < HP, "mlp_rollout_f16: padded width must leave a bias slot (H < HP)");
    CMRL_REQUIRE(layout == CMRL_HEAD_PLAIN ? layer_n[num_layers] >= (gaussian ? 2 : 1) * head_dim : true,
                 "mlp_rollout_f16: head too narrow");
    p.num_layers = num_layers;
    p.P = P;
    p.E = E;
    p.N = N;
    p.O = O;
    p.A = A;
    p.H = H;
    p.head_dim = head_dim;
    p.layout = layout;
    p.gaussian = gaussian;
    p.act = act;
    p.obs = obs;
    p.actn = actn;
    p.perm = perm;
    p.offsets = offsets;
    p.residual = residual;
    p.min_logvar = min_logvar;
    p.max_logvar = max_logvar;
    p.bounds_n = bounds_n;
    p.noise = noise;
    p.pred = pred;
    p.dbg = g_dbg_buffer;
    p.dbg_flags = g_dbg_flags;
    if (variant == 6) return tc::launch_rollout_v6(p, (cudaStream_t)stream);
    if (variant == 5) return tc::launch_rollout_v5(p, (cudaStream_t)stream);
    if (variant == 4) return tc::launch_rollout_v4(p, (cudaStream_t)stream);
    if (variant == 3) return tc::launch_rollout_v3(p, (cudaStream_t)stream);
    CMRL_REQUIRE(variant == 1, "mlp_rollout_f16: variant must be 1 (streamed weights, single CTA), 4 (weight-stationary CTA pairs, 256-row tiles), 3 (pairs, two 128-row tiles in flight), 5 (pairs, two 256-row tiles, A operand partly in TMEM) or 6 (pairs, streamed weights, wide hidden)");
    size_t w_stage = 0;
    int kmax = 0;
    for (int l = 0; l <= num_layers; ++l) {
        size_t b = (size_t)layer_n[l] * layer_k[l] * 2;
        if (b > w_stage) w_stage = b;
        if (layer_k[l] > kmax) kmax = layer_k[l];
    }
    w_stage = (w_stage + 1023) / 1024 * 1024;
    p.w_stage_bytes = (uint32_t)w_stage;
    p.a_chunks = kmax / 8;
    const size_t a_bytes = (size_t)p.a_chunks * tc::A_CHUNK_BYTES;
    const size_t smem = 2 * w_stage + a_bytes + 96 + sizeof(int) * (E + 1) + 16;
    CMRL_REQUIRE(smem <= 232448, "mlp_rollout_f16: hidden width %d needs %zu B of shared memory (> 227 KB)", H, smem);
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(tc::mlp_rollout_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_once.done(attr_dev);
    }
    int items = P * ((N + tc::TILE_M - 1) / tc::TILE_M + E);
    int grid = num_sms();
    if (grid > items) grid = items;
    tc::mlp_rollout_f16_kernel<<<grid, tc::NUM_THREADS, smem, (cudaStream_t)stream>>>(p);
    CMRL_CHECK_LAUNCH("mlp_rollout_f16");
    return 0;
}
