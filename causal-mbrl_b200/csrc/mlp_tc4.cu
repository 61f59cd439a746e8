// Tensor-core tier, v4: weight-stationary CTA pairs, ONE 256-row tile per pair (tcgen05.mma cta_group::2, M=256),
// layer l+1's MMAs chase layer l's epilogue in 64-column groups.
//
// Why not two 128-row tiles in flight (v3)?  Measured (profiles/tools/tmem_contention.cu): with per-K-step operand
// addresses a cta_group::2 MMA with M=128 takes 71 cycles alone and ~100 cycles next to a running epilogue -- the B
// operand traffic per MMA does not shrink with M -- while M=256 stays at 105 cycles for twice the rows.  Per row the
// M=256 MMA is 2x cheaper, which outweighs the better overlap of the ping-pong schedule.
//
// Schedule per tile and layer (128 rows per CTA): 16 epilogue warps (4 per TMEM lane quarter, warp set s takes the
// 16-column chunks c = 4g + s) convert accumulator buffer l&1 into the fp16 A operand of layer l+1; after round g the
// 64-column group g is published and the MMA warp issues K steps 4g..4g+3 of layer l+1 into the other accumulator
// buffer.  Only the last group (one K step for H=200) is exposed after the epilogue.  The epilogue is SFU-bound
// (16 SFU lanes per SM: one tanh per activation = 1 664 cycles per layer) against 1 365 cycles of MMA.
//
// Warp roles per CTA (768 threads): 0-15 epilogue, 16-19 head epilogue (Gaussian head, residual, sampling, scatter;
// one row per lane), 20 MMA issuer (leader CTA) + TMEM allocation, 21 weight TMA producer, 22-23 layer-0 operand
// loaders (two rows per lane, one tile ahead).
#include "tc_common.cuh"

namespace cmrl {
namespace tc {

constexpr int V4_THREADS = 768;
constexpr int V4_TILE = 256;            // rows per tile per CTA pair
constexpr int V4_ROWS = 128;            // rows per CTA
constexpr int V4_CHUNK = V4_ROWS * 16;  // bytes of one 8-element K chunk of the A operand
constexpr int V4_EPI_WARPS = 16, V4_HEAD_WARP0 = 16, V4_WARP_MMA = 20, V4_WARP_W = 21, V4_WARP_X = 22;
constexpr uint32_t V4_D1_COL = 208, V4_DH_COL = 416;  // second accumulator buffer; head accumulators (2 x 32 columns)

struct V4Smem {
    uint32_t w_off[8];
    uint32_t a_off;
    uint32_t x_off[2];
    uint32_t bar_off;
    uint32_t total;
};

enum {
    V4_W_FULL = 0,    // local (tx)
    V4_W_READY = 1,   // leader, count 2
    V4_W_EMPTY = 2,   // both (commit)
    V4_X_READY = 3,   // leader [2], count 4 (2 loader warps x 2 CTAs)
    V4_X_EMPTY = 5,   // both [2] (commit)
    V4_A_READY = 7,   // leader [4 groups], count 2 * 16 warps
    V4_D_FULL = 11,   // both [2] (commit)
    V4_DH_FULL = 13,  // both [2] (commit)
    V4_DH_EMPTY = 15, // leader [2], count 2 * 4 head warps
    V4_BAR_COUNT = 17,
};

// work list of a cluster: contiguous range of 256-row tiles in (p, member, tile) order
struct V4Iter {
    int item, item_end, pp, e, t;
};
struct V4Tile {
    int net, pp, row0, nrows;
};
__device__ __forceinline__ int v4_tiles_of(const int* s_off, int e) { return (s_off[e + 1] - s_off[e] + V4_TILE - 1) / V4_TILE; }
__device__ __forceinline__ void v4_iter_init(V4Iter& it, const RolloutP& p, int item_begin, int item_end, int tiles_total,
                                             const int* s_off) {
    it.item = item_begin;
    it.item_end = item_end;
    it.pp = tiles_total > 0 ? item_begin / tiles_total : 0;
    int t = item_begin - it.pp * tiles_total;
    it.e = 0;
    while (it.e < p.E && t >= v4_tiles_of(s_off, it.e)) {
        t -= v4_tiles_of(s_off, it.e);
        ++it.e;
    }
    it.t = t;
}
__device__ __forceinline__ bool v4_iter_next(V4Iter& it, const RolloutP& p, const int* s_off, V4Tile& tl) {
    if (it.item >= it.item_end) return false;
    while (it.t >= v4_tiles_of(s_off, it.e)) {
        it.t = 0;
        if (++it.e >= p.E) {
            it.e = 0;
            ++it.pp;
        }
    }
    const int cnt = s_off[it.e + 1] - s_off[it.e];
    tl.pp = it.pp;
    tl.net = it.pp * p.E + it.e;
    tl.row0 = s_off[it.e] + it.t * V4_TILE;
    tl.nrows = min(V4_TILE, cnt - it.t * V4_TILE);
    ++it.t;
    ++it.item;
    return true;
}

__device__ __forceinline__ void v4_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(V4_THREADS, 1)
    mlp_rollout_f16_v4_kernel(const RolloutP p, const V4Smem sm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_off[65];
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    const int K0 = p.layer_k[0];
    const int NH = p.layer_n[L];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bar_off);
    const uint32_t rank = v3_ctarank();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_chunks = HP / 16;               // 16-column accumulator chunks = K steps of the next layer
    const int n_groups = (n_chunks + 3) >> 2;   // 64-column hand-off groups

    for (int i = threadIdx.x; i <= p.E; i += blockDim.x) s_off[i] = p.offsets[i];
    if (threadIdx.x == 0) {
        mbar_init(&bars[V4_W_FULL], 1);
        mbar_init(&bars[V4_W_READY], 2);
        mbar_init(&bars[V4_W_EMPTY], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[V4_X_READY + i], 4);
            mbar_init(&bars[V4_X_EMPTY + i], 1);
            mbar_init(&bars[V4_D_FULL + i], 1);
            mbar_init(&bars[V4_DH_FULL + i], 1);
            mbar_init(&bars[V4_DH_EMPTY + i], 2 * 4);
        }
        for (int g = 0; g < 4; ++g) mbar_init(&bars[V4_A_READY + g], 2 * V4_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == V4_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    v3_cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    int tiles_total = 0;
    for (int e = 0; e < p.E; ++e) tiles_total += v4_tiles_of(s_off, e);
    const int n_items = p.P * tiles_total;
    const int n_clusters = gridDim.x / 2;
    const int cid = blockIdx.x / 2;
    const int item_begin = (int)((long long)cid * n_items / n_clusters);
    const int item_end = (int)((long long)(cid + 1) * n_items / n_clusters);
    V4Iter it;
    v4_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
    V4Tile tl;

    if (warp == V4_WARP_W) {
        // ===================== weight producer: this CTA's half of every layer of the current net =====================
        if (lane == 0) {
            int cur_net = -1;
            uint32_t n_w = 0;
            while (v4_iter_next(it, p, s_off, tl)) {
                if (tl.net == cur_net) continue;
                if (cur_net != -1) mbar_wait(&bars[V4_W_EMPTY], (n_w - 1) & 1);
                const __half* blob = p.packed + (long long)tl.net * p.blob_halfs;
                uint32_t total = 0;
                for (int l = 0; l <= L; ++l) total += (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l] * 2;
                mbar_expect_tx(&bars[V4_W_FULL], total);
                for (int l = 0; l <= L; ++l) {
                    const uint32_t half_halfs = (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l];
                    tma_bulk_g2s(smem + sm.w_off[l], blob + p.layer_off[l] + (size_t)rank * half_halfs, half_halfs * 2,
                                 &bars[V4_W_FULL]);
                }
                mbar_wait(&bars[V4_W_FULL], n_w & 1);
                v3_arrive_leader(&bars[V4_W_READY]);
                ++n_w;
                cur_net = tl.net;
            }
        }
    } else if (warp == V4_WARP_MMA) {
        // ===================== MMA issuer: whole warp walks the schedule, one elected lane issues =====================
        if (rank == 0) {
            int cur_net = -1;
            uint32_t n_w = 0, n_x = 0, n_epi = 0, n_h = 0;
            const uint32_t idesc_h = umma_idesc_f16(V4_TILE, HP), idesc_head = umma_idesc_f16(V4_TILE, NH);
            const uint32_t b_sbo0 = (uint32_t)(K0 / 8) * 128, b_sbo = (uint32_t)(HP / 8) * 128;
            const uint64_t ad_a = umma_desc(smem_u32(smem + sm.a_off), V4_CHUNK, 128);
            const uint32_t a_hi = (uint32_t)(ad_a >> 32);
            const bool timing = p.dbg != nullptr && lane == 0;
            long long t_wa = 0, t_issue = 0, t_wx = 0, t_all = timing ? clock64() : 0, tt = 0;
            while (v4_iter_next(it, p, s_off, tl)) {
                if (tl.net != cur_net) {
                    if (cur_net != -1) v3_commit_uni(&bars[V4_W_EMPTY]);
                    mbar_wait(&bars[V4_W_READY], n_w & 1);
                    ++n_w;
                    cur_net = tl.net;
                }
                // ---- layer 0 ------------------------------------------------------------------------------------------
                {
                    const int xb = n_x & 1;
                    if (timing) tt = clock64();
                    mbar_wait(&bars[V4_X_READY + xb], (n_x >> 1) & 1);
                    if (timing) t_wx += clock64() - tt;
                    tc_fence_after();
                    const uint64_t ad0 = umma_desc(smem_u32(smem + sm.x_off[xb]), V4_CHUNK, 128);
                    const uint64_t bd0 = umma_desc(smem_u32(smem + sm.w_off[0]), 128, b_sbo0);
                    uint32_t a_lo = (uint32_t)ad0, b_lo = (uint32_t)bd0;
                    for (int ks = 0; ks < K0 / 16; ++ks, a_lo += 2 * V4_CHUNK / 16, b_lo += 16)
                        v3_mma_uni(tmem_base, a_lo, (uint32_t)(ad0 >> 32), b_lo, (uint32_t)(bd0 >> 32), idesc_h, ks > 0);
                    v3_commit_uni(&bars[V4_X_EMPTY + xb]);
                    v3_commit_uni(&bars[V4_D_FULL + 0]);
                    ++n_x;
                }
                // ---- layers 1..L (L = head): K steps follow the epilogue's 64-column groups ---------------------------------
                for (int l = 1; l <= L; ++l) {
                    const bool head = (l == L);
                    const uint64_t bd0 = umma_desc(smem_u32(smem + sm.w_off[l]), 128, b_sbo);
                    const uint32_t b_hi = (uint32_t)(bd0 >> 32);
                    const uint32_t d_tmem = tmem_base + (head ? V4_DH_COL + (n_h & 1) * 32 : ((l & 1) ? V4_D1_COL : 0u));
                    const uint32_t idesc = head ? idesc_head : idesc_h;
                    if (head) mbar_wait(&bars[V4_DH_EMPTY + (n_h & 1)], ((n_h >> 1) & 1) ^ 1);
                    uint32_t a_lo = (uint32_t)ad_a, b_lo = (uint32_t)bd0;
                    int ks = 0;
                    for (int g = 0; g < n_groups; ++g) {
                        if (timing) tt = clock64();
                        mbar_wait(&bars[V4_A_READY + g], n_epi & 1);
                        if (timing) {
                            long long t = clock64();
                            t_wa += t - tt;
                            tt = t;
                        }
                        tc_fence_after();
                        const int ks1 = min(n_chunks, 4 * g + 4);
                        for (; ks < ks1; ++ks, a_lo += 2 * V4_CHUNK / 16, b_lo += 16)
                            v3_mma_uni(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, ks > 0);
                        if (timing) t_issue += clock64() - tt;
                    }
                    if (head) {
                        v3_commit_uni(&bars[V4_DH_FULL + (n_h & 1)]);
                        ++n_h;
                    } else {
                        v3_commit_uni(&bars[V4_D_FULL + (l & 1)]);
                    }
                    ++n_epi;
                }
            }
            if (timing) {
                long long* d = p.dbg + (long long)blockIdx.x * 16;
                d[0] = clock64() - t_all;
                d[1] = t_wx;
                d[2] = t_wa;
                d[3] = t_issue;
                d[4] = item_end - item_begin;
            }
        }
    } else if (warp >= V4_WARP_X) {
        // ===================== layer-0 operand loaders: 64 threads x 2 rows, one tile ahead ===========================
        const int lt = threadIdx.x - V4_WARP_X * 32;
        const int in = p.O + p.A;
        uint32_t n_x = 0;
        while (v4_iter_next(it, p, s_off, tl)) {
            const int xb = n_x & 1;
            mbar_wait(&bars[V4_X_EMPTY + xb], ((n_x >> 1) & 1) ^ 1);
            ++n_x;
            uint8_t* xbuf = smem + sm.x_off[xb];
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
                const int r = lt + hr * 64;
                const int grow = (int)rank * V4_ROWS + r;
                const bool valid = grow < tl.nrows;
                const long long env = valid ? p.perm[tl.row0 + grow] : 0;
                for (int c = 0; c < K0 / 8; ++c) {
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
                    *reinterpret_cast<uint4*>(xbuf + c * V4_CHUNK + r * 16) =
                        make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) v3_arrive_leader(&bars[V4_X_READY + xb]);
        }
    } else if (warp >= V4_HEAD_WARP0) {
        // ===================== head warps: one row per lane ==================================================================
        const int q = warp - V4_HEAD_WARP0;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int D = p.head_dim;
        uint32_t n_h = 0;
        while (v4_iter_next(it, p, s_off, tl)) {
            const int grow = (int)rank * V4_ROWS + row;
            const bool valid = grow < tl.nrows;
            const long long env = valid ? p.perm[tl.row0 + grow] : 0;
            float res0 = 0.0f, nz0 = 0.0f;
            if (p.layout == CMRL_HEAD_PARALLEL && valid) {  // prefetch before the wait
                if (p.residual) res0 = p.obs[env * p.O + tl.pp];
                if (p.gaussian && p.noise) nz0 = p.noise[env * D + tl.pp];
            }
            const int hb = n_h & 1;
            mbar_wait(&bars[V4_DH_FULL + hb], (n_h >> 1) & 1);
            ++n_h;
            tc_fence_after();
            const uint32_t h_addr = tmem_base + lane_base + V4_DH_COL + hb * 32;
            for (int c0 = 0; c0 < NH; c0 += 16) {
                uint32_t raw[16];
                __syncwarp();
                v4_ld16_issue(h_addr + c0, raw);
                v3_ld_wait();
                if (valid) {
                    if (p.layout == CMRL_HEAD_PARALLEL) {
                        if (c0 == 0) {
                            float out = __uint_as_float(raw[0]) + res0;
                            if (p.gaussian && p.noise) {
                                const int bi = p.bounds_n == 1 ? 0 : tl.pp;
                                const float lv = soft_clamp_logvar(__uint_as_float(raw[1]), p.min_logvar[bi], p.max_logvar[bi]);
                                out += sqrtf(expf(lv)) * nz0;
                            }
                            p.pred[env * D + tl.pp] = out;
                        }
                    } else if (p.gaussian) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const int d = (c0 + j) >> 1;
                            if (d < D) {
                                float m = __uint_as_float(raw[j]);
                                if (p.residual) m += p.obs[env * p.O + d];
                                if (p.noise) {
                                    const int bi = p.bounds_n == 1 ? 0 : d;
                                    const float lv = soft_clamp_logvar(__uint_as_float(raw[j + 1]), p.min_logvar[bi], p.max_logvar[bi]);
                                    m += sqrtf(expf(lv)) * p.noise[env * D + d];
                                }
                                p.pred[env * D + d] = m;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int d = c0 + j;
                            if (d < D) {
                                float m = __uint_as_float(raw[j]);
                                if (p.residual) m += p.obs[env * p.O + d];
                                p.pred[env * D + d] = m;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) v3_arrive_leader(&bars[V4_DH_EMPTY + hb]);
        }
    } else {
        // ===================== epilogue warps ================================================================================
        // quarter q = TMEM lanes 32q.. = rows of the tile; warp set s converts the 16-column chunks c = 4g + s, g = 0,1,..
        // (one tcgen05.ld.x16, 16 activations, two 16-byte stores = K chunks 2c, 2c+1 of the next layer's A operand) and
        // publishes group g after its chunk.  The next chunk's TMEM load is issued before the current one is converted.
        const int q = warp & 3, cset = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cset * 16;
        const uint32_t a_addr = smem_u32(smem + sm.a_off) + (uint32_t)(2 * cset * V4_CHUNK + row * 16);
        const int my_n = (n_chunks - cset + 3) >> 2;  // chunks of this warp
        // ones column (fp16 1.0 at hidden unit H = the next layer's bias input): owned by the warp whose chunk holds it
        const int ones_c = p.H >> 4;
        const bool owns_ones = (ones_c & 3) == cset && ones_c < n_chunks;
        const int ones_i = ones_c >> 2;
        const uint32_t ones_addr = smem_u32(smem + sm.a_off) + (uint32_t)((p.H >> 3) * V4_CHUNK + row * 16 + (p.H & 7) * 2);
        uint32_t n_d0 = 0, n_d1 = 0;
        uint32_t ra[16], rb[16];
        const bool timing = p.dbg != nullptr && threadIdx.x == 0;
        long long t_wd = 0, t_loop = 0, t0 = timing ? clock64() : 0, tt = 0;
        auto convert_store = [&](int i, const uint32_t(&raw)[16]) {
            uint32_t hh[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) hh[j] = v3_act_pack<ACT>(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
            const uint32_t dst = a_addr + (uint32_t)(i * 8 * V4_CHUNK);
            v3_sts128(dst, hh[0], hh[1], hh[2], hh[3]);
            v3_sts128(dst + V4_CHUNK, hh[4], hh[5], hh[6], hh[7]);
            if (owns_ones && i == ones_i) v3_sts16(ones_addr, (uint16_t)0x3C00);
        };
        auto publish = [&](int g) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) v3_arrive_leader(&bars[V4_A_READY + g]);
        };
        while (v4_iter_next(it, p, s_off, tl)) {
            for (int l = 0; l < L; ++l) {
                if (timing) tt = clock64();
                if (l & 1) {
                    mbar_wait(&bars[V4_D_FULL + 1], n_d1 & 1);
                    ++n_d1;
                } else {
                    mbar_wait(&bars[V4_D_FULL + 0], n_d0 & 1);
                    ++n_d0;
                }
                if (timing) {
                    long long t = clock64();
                    t_wd += t - tt;
                    tt = t;
                }
                tc_fence_after();
                const uint32_t d_addr = t_addr + ((l & 1) ? V4_D1_COL : 0u);
                __syncwarp();
                if (my_n > 0) v4_ld16_issue(d_addr, ra);
                for (int g = 0; g < n_groups; g += 2) {
                    if (g < my_n) {
                        v3_ld_wait();
                        if (g + 1 < my_n) v4_ld16_issue(d_addr + (g + 1) * 64, rb);
                        else tc_fence_before();
                        convert_store(g, ra);
                    }
                    publish(g);
                    if (g + 1 < n_groups) {
                        if (g + 1 < my_n) {
                            v3_ld_wait();
                            if (g + 2 < my_n) v4_ld16_issue(d_addr + (g + 2) * 64, ra);
                            else tc_fence_before();
                            convert_store(g + 1, rb);
                        }
                        publish(g + 1);
                    }
                }
                if (timing) t_loop += clock64() - tt;
            }
        }
        if (timing) {
            long long* d = p.dbg + (long long)blockIdx.x * 16;
            d[8] = clock64() - t0;
            d[9] = t_wd;
            d[10] = t_loop;
        }
    }

    // ---- teardown ------------------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    v3_cluster_sync();
    if (warp == V4_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

int launch_rollout_v4(const RolloutP& p, cudaStream_t st) {
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    CMRL_REQUIRE(HP <= 208 && HP % 16 == 0, "mlp_rollout_f16 v4: padded hidden width must be <= 208 (hidden <= 207)");
    CMRL_REQUIRE(p.layer_n[L] <= 32, "mlp_rollout_f16 v4: head wider than 32 columns (use variant 3)");
    V4Smem sm{};
    uint32_t off = 0;
    for (int l = 0; l <= L; ++l) {
        sm.w_off[l] = off;
        off += (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l] * 2;
        off = (off + 127) / 128 * 128;
    }
    sm.a_off = off;
    off += (uint32_t)(HP / 8) * V4_CHUNK;
    for (int i = 0; i < 2; ++i) {
        sm.x_off[i] = off;
        off += (uint32_t)(p.layer_k[0] / 8) * V4_CHUNK;
    }
    sm.bar_off = off;
    off += V4_BAR_COUNT * 8;
    sm.total = off;
    CMRL_REQUIRE(sm.total <= 232448 - 1024, "mlp_rollout_f16 v4: needs %u B of shared memory", sm.total);
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v4_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v4_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        attr_once.done(attr_dev);
    }
    int max_items = p.P * ((p.N + V4_TILE - 1) / V4_TILE + p.E);
    int clusters = num_sms() / 2;
    if (clusters > max_items) clusters = max_items;
    if (clusters < 1) clusters = 1;
    if (p.act == CMRL_ACT_SILU)
        mlp_rollout_f16_v4_kernel<CMRL_ACT_SILU><<<clusters * 2, V4_THREADS, sm.total, st>>>(p, sm);
    else if (p.act == CMRL_ACT_SILU_HALF)
        mlp_rollout_f16_v4_kernel<CMRL_ACT_SILU_HALF><<<clusters * 2, V4_THREADS, sm.total, st>>>(p, sm);
    else if (p.act == CMRL_ACT_RELU)
        mlp_rollout_f16_v4_kernel<CMRL_ACT_RELU><<<clusters * 2, V4_THREADS, sm.total, st>>>(p, sm);
    else
        mlp_rollout_f16_v4_kernel<CMRL_ACT_IDENTITY><<<clusters * 2, V4_THREADS, sm.total, st>>>(p, sm);
    CMRL_CHECK_LAUNCH("mlp_rollout_f16_v4");
    return 0;
}

}  // namespace tc
}  // namespace cmrl
