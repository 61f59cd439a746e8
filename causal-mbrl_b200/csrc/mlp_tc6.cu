// Tensor-core tier, v6: WIDE hidden layers (H a multiple of 16 up to 512, e.g. the 100-obs / 512-hidden causal model):
// the weights of a layer (0.5 MB at H = 512) do not fit shared memory next to the activations, so they are STREAMED.
//
// CTA pairs (tcgen05.mma cta_group::2, M = 256: 128 rows per CTA), one 256-row tile at a time:
//   * the fp16 A operand of the current layer lives in shared memory (128 rows x (H + 16) columns per CTA: the extra K
//     step holds a constant ones column = the bias input, so the accumulator needs exactly H <= 512 TMEM columns);
//   * a producer warp streams the B operand of every layer in 64-deep K chunks of one 256-column block (16 KB per CTA and
//     stage, 4 stages, 1 KB TMA bulk copies straight out of the packed per-layer layout), a forwarder warp tells the
//     leader CTA when both halves of a stage have landed, the MMA warp issues 4 MMAs per stage;
//   * after the last MMA of a layer the 16 epilogue warps convert the accumulator (bias already in, activation, fp16)
//     back into the A operand in place; the head layer's accumulator goes through the Gaussian head / sampling code.
// MMAs and epilogue of one tile do not overlap (the accumulator fills the tensor memory): expected tensor-pipe duty
// ~2/3 at H = 512.
#include "tc_common.cuh"

namespace cmrl {
namespace tc {

constexpr int V6_THREADS = 608;  // 19 warps: 0-15 epilogue / input loader / head, 16 MMA issuer, 17 weight producer, 18 forwarder
constexpr int V6_TILE = 256;
constexpr int V6_ROWS = 128;
constexpr int V6_CHUNK = V6_ROWS * 16;  // bytes of one 8-element K chunk of the A operand
constexpr int V6_KC = 64;               // K depth of one weight stage
constexpr int V6_STAGES = 4;
constexpr int V6_STAGE_BYTES = 16 * (V6_KC / 8) * 128;  // 16 n-groups x 8 k-groups x 128 B
constexpr int V6_EPI_WARPS = 16, V6_WARP_MMA = 16, V6_WARP_W = 17, V6_WARP_F = 18;

struct V6Lay {
    uint32_t a_off, w_off, bar_off, total;
};

enum {
    V6_W_FULL = 0,    // [stage] local (tx)
    V6_W_READY = 4,   // [stage] leader, count 2 (one forwarder per CTA)
    V6_W_EMPTY = 8,   // [stage] both (commit)
    V6_A_READY = 12,  // leader, count 2 * 16 warps
    V6_D_FULL = 13,   // both (commit)
    V6_BAR_COUNT = 14,
};

// contiguous range of 256-row tiles in (p, member, tile) order; same walk in every role
struct V6Iter {
    int item, item_end, pp, e, t;
};
struct V6Tile {
    int net, pp, row0, nrows;
};
__device__ __forceinline__ int v6_tiles_of(const int* s_off, int e) { return (s_off[e + 1] - s_off[e] + V6_TILE - 1) / V6_TILE; }
__device__ __forceinline__ void v6_iter_init(V6Iter& it, const RolloutP& p, int item_begin, int item_end, int tiles_total,
                                             const int* s_off) {
    it.item = item_begin;
    it.item_end = item_end;
    it.pp = tiles_total > 0 ? item_begin / tiles_total : 0;
    int t = item_begin - it.pp * tiles_total;
    it.e = 0;
    while (it.e < p.E && t >= v6_tiles_of(s_off, it.e)) {
        t -= v6_tiles_of(s_off, it.e);
        ++it.e;
    }
    it.t = t;
}
__device__ __forceinline__ bool v6_iter_next(V6Iter& it, const RolloutP& p, const int* s_off, V6Tile& tl) {
    if (it.item >= it.item_end) return false;
    while (it.t >= v6_tiles_of(s_off, it.e)) {
        it.t = 0;
        if (++it.e >= p.E) {
            it.e = 0;
            ++it.pp;
        }
    }
    const int cnt = s_off[it.e + 1] - s_off[it.e];
    tl.pp = it.pp;
    tl.net = it.pp * p.E + it.e;
    tl.row0 = s_off[it.e] + it.t * V6_TILE;
    tl.nrows = min(V6_TILE, cnt - it.t * V6_TILE);
    ++it.t;
    ++it.item;
    return true;
}

// The weight stream of one tile: for every layer, every 256-column block of its N, every 64-deep K chunk -> one stage.
// `fn(l, nb, width, kc, kgroups)`: width = columns of the block (multiple of 16), kgroups = 8-wide K groups in the chunk.
template <class F>
__device__ __forceinline__ void v6_for_each_stage(const RolloutP& p, F&& fn) {
    for (int l = 0; l <= p.num_layers; ++l) {
        const int N = p.layer_n[l], K = p.layer_k[l];
        for (int nb = 0; nb * 256 < N; ++nb) {
            const int width = min(256, N - nb * 256);
            for (int kc = 0; kc * V6_KC < K; ++kc) fn(l, nb, width, kc, min(V6_KC, K - kc * V6_KC) / 8);
        }
    }
}

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(V6_THREADS, 1)
    mlp_rollout_f16_v6_kernel(const RolloutP p, const V6Lay lay) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_off[65];
    const int L = p.num_layers;
    const int K0 = p.layer_k[0];
    const int NH = p.layer_n[L];
    const int HN = p.layer_n[0];  // hidden width (accumulator columns), multiple of 16
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
    const uint32_t rank = v3_ctarank();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i <= p.E; i += blockDim.x) s_off[i] = p.offsets[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < V6_STAGES; ++s) {
            mbar_init(&bars[V6_W_FULL + s], 1);
            mbar_init(&bars[V6_W_READY + s], 2);
            mbar_init(&bars[V6_W_EMPTY + s], 1);
        }
        mbar_init(&bars[V6_A_READY], 2 * V6_EPI_WARPS);
        mbar_init(&bars[V6_D_FULL], 1);
        fence_barrier_init();
    }
    if (warp == V6_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    // the constant bias-input K step of the hidden / head layers: column H of the A operand = 1, columns H+1 .. H+15 = 0
    if (threadIdx.x < V6_ROWS) {
        uint8_t* a = smem + lay.a_off + (size_t)(p.H / 8) * V6_CHUNK + threadIdx.x * 16;
        *reinterpret_cast<uint4*>(a) = make_uint4(0x00003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(a + V6_CHUNK) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    v3_cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    int tiles_total = 0;
    for (int e = 0; e < p.E; ++e) tiles_total += v6_tiles_of(s_off, e);
    const int n_items = p.P * tiles_total;
    const int n_clusters = gridDim.x / 2;
    const int cid = blockIdx.x / 2;
    const int item_begin = (int)((long long)cid * n_items / n_clusters);
    const int item_end = (int)((long long)(cid + 1) * n_items / n_clusters);
    V6Iter it;
    v6_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
    V6Tile tl;

    if (warp == V6_WARP_W) {
        // ===================== weight producer: runs up to V6_STAGES stages ahead ==================================================
        uint32_t n_st = 0;
        while (v6_iter_next(it, p, s_off, tl)) {
            const __half* blob = p.packed + (long long)tl.net * p.blob_halfs;
            v6_for_each_stage(p, [&](int l, int nb, int width, int kc, int kg) {
                const int s = n_st % V6_STAGES;
                mbar_wait(&bars[V6_W_EMPTY + s], ((n_st / V6_STAGES) & 1) ^ 1);
                const int ng = width / 16;  // 8-row groups of this CTA's half of the block
                const uint32_t seg = (uint32_t)kg * 128;
                if (lane == 0) mbar_expect_tx(&bars[V6_W_FULL + s], (uint32_t)ng * seg);
                __syncwarp();
                const int n0 = nb * 256 + (int)rank * (width / 2);  // first output column of this CTA's half
                for (int g = lane; g < ng; g += 32) {
                    const __half* src = blob + p.layer_off[l] + (size_t)(n0 / 8 + g) * ((size_t)p.layer_k[l] * 8) + (size_t)kc * (V6_KC * 8);
                    tma_bulk_g2s(smem + lay.w_off + s * V6_STAGE_BYTES + g * seg, src, seg, &bars[V6_W_FULL + s]);
                }
                ++n_st;
            });
        }
    } else if (warp == V6_WARP_F) {
        // ===================== forwarder: a stage of this CTA has landed -> tell the leader ===================================
        if (lane == 0) {
            uint32_t n_st = 0;
            while (v6_iter_next(it, p, s_off, tl)) {
                v6_for_each_stage(p, [&](int, int, int, int, int) {
                    const int s = n_st % V6_STAGES;
                    mbar_wait(&bars[V6_W_FULL + s], (n_st / V6_STAGES) & 1);
                    v3_arrive_leader(&bars[V6_W_READY + s]);
                    ++n_st;
                });
            }
        }
    } else if (warp == V6_WARP_MMA) {
        // ===================== MMA issuer (leader CTA) ===========================================================================
        if (rank == 0) {
            uint32_t n_st = 0, n_a = 0;
            const uint64_t ad_a = umma_desc(smem_u32(smem + lay.a_off), V6_CHUNK, 128);
            const uint32_t a_hi = (uint32_t)(ad_a >> 32);
            while (v6_iter_next(it, p, s_off, tl)) {
                int cur_l = -1;
                v6_for_each_stage(p, [&](int l, int nb, int width, int kc, int kg) {
                    if (l != cur_l) {  // first stage of a layer: its A operand must be complete
                        if (cur_l >= 0) v3_commit_uni(&bars[V6_D_FULL]);
                        mbar_wait(&bars[V6_A_READY], n_a & 1);
                        ++n_a;
                        tc_fence_after();
                        cur_l = l;
                    }
                    const int s = n_st % V6_STAGES;
                    mbar_wait(&bars[V6_W_READY + s], (n_st / V6_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t idesc = umma_idesc_f16(V6_TILE, width);
                    const uint64_t bd = umma_desc(smem_u32(smem + lay.w_off + s * V6_STAGE_BYTES), 128, (uint32_t)kg * 128);
                    uint32_t b_lo = (uint32_t)bd;
                    const uint32_t b_hi = (uint32_t)(bd >> 32);
                    uint32_t a_lo = (uint32_t)ad_a + (uint32_t)(kc * (V6_KC / 8)) * (V6_CHUNK / 16);
                    const uint32_t d_tmem = tmem_base + (uint32_t)nb * 256u;
#pragma unroll 1
                    for (int j = 0; j < kg / 2; ++j, a_lo += 2 * V6_CHUNK / 16, b_lo += 16)
                        v3_mma_uni(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, (kc > 0 || j > 0) ? 1u : 0u);
                    v3_commit_uni(&bars[V6_W_EMPTY + s]);
                    ++n_st;
                });
                v3_commit_uni(&bars[V6_D_FULL]);
            }
        }
    } else {
        // ===================== epilogue warps: input loader, per-layer conversion, head ===========================================
        const int q = warp & 3, cset = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int in = p.O + p.A;
        const int D = p.head_dim;
        uint32_t n_d = 0;
        const uint32_t a_row = smem_u32(smem + lay.a_off) + (uint32_t)(row * 16);
        auto publish = [&]() {
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) v3_arrive_leader(&bars[V6_A_READY]);
        };
        while (v6_iter_next(it, p, s_off, tl)) {
            const int grow = (int)rank * V6_ROWS + row;
            const bool valid = grow < tl.nrows;
            const long long env = valid ? p.perm[tl.row0 + grow] : 0;
            // ---- layer-0 operand: this warp set writes the 8-wide K chunks cset, cset + 4, ... of its 32 rows ---------------
#pragma unroll 1
            for (int c = cset; c < K0 / 8; c += 4) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = c * 8 + j;
                    float x = 0.0f;
                    if (valid) {
                        if (k < p.O) x = p.obs[env * p.O + k];
                        else if (k < in) x = p.actn[env * p.A + (k - p.O)];
                        else if (k == in) x = 1.0f;
                    }
                    v[j] = x;
                }
                v3_sts128(a_row + (uint32_t)c * V6_CHUNK, pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
            }
            publish();
            // ---- hidden layers: accumulator -> activation -> fp16 A operand, in place ------------------------------------------
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                mbar_wait(&bars[V6_D_FULL], n_d & 1);
                ++n_d;
                tc_fence_after();
#pragma unroll 1
                for (int c = cset; c < HN / 16; c += 4) {
                    uint32_t raw[16], hh[8];
                    __syncwarp();
                    v5_ld16_issue(tmem_base + lane_off + (uint32_t)c * 16, raw);
                    v3_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) hh[j] = v3_act_pack<ACT>(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
                    const uint32_t dst = a_row + (uint32_t)(2 * c) * V6_CHUNK;
                    v3_sts128(dst, hh[0], hh[1], hh[2], hh[3]);
                    v3_sts128(dst + V6_CHUNK, hh[4], hh[5], hh[6], hh[7]);
                }
                publish();
            }
            // ---- head: one row per lane, handled by warp set 0; everybody waits (the next tile's input overwrites A) --------
            mbar_wait(&bars[V6_D_FULL], n_d & 1);
            ++n_d;
            tc_fence_after();
            if (cset == 0) {
                const int n_out = p.layout == CMRL_HEAD_PARALLEL ? 1 : D;
                const int stride = p.gaussian ? 2 : 1;
#pragma unroll 1
                for (int d = 0; d < n_out; ++d) {
                    uint32_t r0, r1 = 0;
                    __syncwarp();
                    if (p.gaussian)
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(tmem_base + lane_off + (uint32_t)(d * stride)));
                    else
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(tmem_base + lane_off + (uint32_t)d));
                    v3_ld_wait();
                    if (valid) {
                        const int od = p.layout == CMRL_HEAD_PARALLEL ? tl.pp : d;
                        float m = __uint_as_float(r0);
                        if (p.residual) m += p.obs[env * p.O + od];
                        if (p.gaussian && p.noise) {
                            const int bi = p.bounds_n == 1 ? 0 : od;
                            const float lv = soft_clamp_logvar(__uint_as_float(r1), p.min_logvar[bi], p.max_logvar[bi]);
                            m += sqrtf(expf(lv)) * p.noise[env * D + od];
                        }
                        p.pred[(long long)(tl.net % p.E) * p.pred_se + env * D + od] = m;
                    }
                }
            }
            (void)NH;
        }
    }

    // ---- teardown ------------------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    v3_cluster_sync();
    if (warp == V6_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

int launch_rollout_v6(const RolloutP& p, cudaStream_t st) {
    const int L = p.num_layers;
    const int HN = p.layer_n[0];
    CMRL_REQUIRE(p.H % 16 == 0 && HN == p.H && HN <= 512, "mlp_rollout_f16 v6: hidden width must be a multiple of 16, <= 512 (got %d)", p.H);
    CMRL_REQUIRE(p.layer_k[0] <= p.H, "mlp_rollout_f16 v6: padded input width %d exceeds the hidden width", p.layer_k[0]);
    int kmax = p.layer_k[0];
    for (int l = 1; l <= L; ++l) {
        CMRL_REQUIRE(p.layer_k[l] == p.H + 16, "mlp_rollout_f16 v6: layer %d must be packed with K = H + 16 (un-widened layout)", l);
        if (l < L) CMRL_REQUIRE(p.layer_n[l] == HN, "mlp_rollout_f16 v6: hidden width must be uniform");
        kmax = p.layer_k[l];
    }
    CMRL_REQUIRE(p.layer_n[L] <= 256, "mlp_rollout_f16 v6: head wider than 256 columns");
    V6Lay lay{};
    uint32_t off = 0;
    lay.a_off = off;
    off += (uint32_t)(kmax / 8) * V6_CHUNK;
    lay.w_off = off;
    off += V6_STAGES * V6_STAGE_BYTES;
    lay.bar_off = off;
    off += V6_BAR_COUNT * 8;
    lay.total = off;
    CMRL_REQUIRE(lay.total <= 232448 - 1024, "mlp_rollout_f16 v6: needs %u B of shared memory", lay.total);
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v6_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v6_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v6_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v6_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        attr_once.done(attr_dev);
    }
    int max_items = p.P * ((p.N + V6_TILE - 1) / V6_TILE + p.E);
    int clusters = num_sms() / 2;
    if (clusters > max_items) clusters = max_items;
    if (clusters < 1) clusters = 1;
    if (p.act == CMRL_ACT_SILU)
        mlp_rollout_f16_v6_kernel<CMRL_ACT_SILU><<<clusters * 2, V6_THREADS, lay.total, st>>>(p, lay);
    else if (p.act == CMRL_ACT_SILU_HALF)
        mlp_rollout_f16_v6_kernel<CMRL_ACT_SILU_HALF><<<clusters * 2, V6_THREADS, lay.total, st>>>(p, lay);
    else if (p.act == CMRL_ACT_RELU)
        mlp_rollout_f16_v6_kernel<CMRL_ACT_RELU><<<clusters * 2, V6_THREADS, lay.total, st>>>(p, lay);
    else
        mlp_rollout_f16_v6_kernel<CMRL_ACT_IDENTITY><<<clusters * 2, V6_THREADS, lay.total, st>>>(p, lay);
    CMRL_CHECK_LAUNCH("mlp_rollout_f16_v6");
    return 0;
}

}  // namespace tc
}  // namespace cmrl
