// Tensor-core tier, v3: weight-stationary CTA pairs with TWO row tiles in flight ("ping-pong").
//
// Measured on B200 (profiles/tools/*.cu): a 16-element SiLU epilogue chunk costs ~140 cycles per SM sub-partition
// (SFU bound, tanh.approx = 8 cycles per warp instruction), a [256 x 208 x 16] cta_group::2 MMA 105 cycles, and the
// two do not slow each other down.  A single tile per pair therefore leaves the tensor pipe idle while the epilogue
// converts the layer and vice versa (dependency chain inside one tile).  Here each pair works on two independent
// 128-row tiles (64 rows per CTA, tcgen05.mma cta_group::2 with M=128, which runs at the full rate: 53.8 cycles per
// K step for N=208): while the 16 epilogue warps convert tile a's layer l, the tensor pipe runs tile b's layer l.
//
// Accumulator layout of M=128 / cta_group::2 (probed with profiles/tools/tmem_layout.cu): each CTA holds its 64 rows
// as 128 TMEM lanes x N/2 columns -- lanes 0-63 = rows 0-63, columns [0, N/2); lanes 64-127 = rows 0-63, columns
// [N/2, N).  So all four lane quarters take part in the epilogue.
//
// Warp roles per CTA (768 threads): 0-15 epilogue (quarter = warp % 4, set = warp / 4 takes the 8-column chunks
// c = set mod 4), 16-19 head epilogue (Gaussian head, residual, sampling, scatter), 20/21 MMA issuers of tile slot
// 0/1 (leader CTA; a [128 x 208 x 16] MMA lasts 54 cycles but costs ~100 cycles of single-thread issue work, so one
// issuer per slot keeps the tensor pipe fed; warp 20 also owns the TMEM allocation), 22 weight TMA producer,
// 23 layer-0 operand loader (two rows per lane).
#include "tc_common.cuh"

namespace cmrl {
namespace tc {

constexpr int V3_THREADS = 768;
constexpr int V3_TILE = 128;    // rows per tile per CTA pair
constexpr int V3_ROWS = 64;     // rows per tile per CTA
constexpr int V3_CHUNK = V3_ROWS * 16;  // bytes of one 8-element K chunk of the A operand (64 rows)
constexpr int V3_EPI_WARPS = 16, V3_HEAD_WARP0 = 16, V3_WARP_MMA = 20 /* and 21: one issuer per tile slot */, V3_WARP_W = 22, V3_WARP_X = 23;
// TMEM columns: hidden-layer accumulators of tile slot 0 / 1 (N/2 <= 128 columns each), then the head accumulators
__device__ __forceinline__ uint32_t v3_d_col(int s) { return (uint32_t)s * 128u; }
__device__ __forceinline__ uint32_t v3_dh_col(int s) { return 256u + (uint32_t)s * 64u; }

struct V3Smem {
    uint32_t w_off[8];
    uint32_t a_off[2];      // A operand per tile slot
    uint32_t x_off[2][2];   // layer-0 operand [stage][slot]
    uint32_t bar_off;
    uint32_t total;
};

enum {
    V3_W_FULL = 0,    // local (tx)
    V3_W_READY = 1,   // leader, count 2
    V3_W_EMPTY = 2,   // both (one commit per MMA issuer -> count 2)
    V3_X_READY = 3,   // leader [stage*2+slot], count 2 (one loader warp per CTA)
    V3_X_EMPTY = 7,   // both   [stage*2+slot] (commit)
    V3_A_READY = 11,  // leader [slot], count 2 * 16 warps
    V3_D_FULL = 13,   // both   [slot] (commit)
    V3_DH_FULL = 15,  // both   [slot] (commit)
    V3_DH_EMPTY = 17, // leader [slot], count 2 * 4 head warps
    V3_BAR_COUNT = 19,
};

// The work list of a cluster is a contiguous range of 128-row tiles in (p, member, tile) order; consecutive tiles of
// the same net are processed two at a time.  Every role walks the list with this iterator (decode once, then step).
struct V3Pair {
    int net, pp, n;       // net index, sub-net index, tiles in this pair (1 or 2)
    int row0[2], nrows[2];
};
struct V3Iter {
    int item, item_end, pp, e, t;  // next tile: sub-net pp, member e, tile t of that member
};
__device__ __forceinline__ int v3_tiles_of(const int* s_off, int e) { return (s_off[e + 1] - s_off[e] + V3_TILE - 1) / V3_TILE; }
__device__ __forceinline__ void v3_iter_init(V3Iter& it, const RolloutP& p, int item_begin, int item_end, int tiles_total,
                                             const int* s_off) {
    it.item = item_begin;
    it.item_end = item_end;
    it.pp = tiles_total > 0 ? item_begin / tiles_total : 0;
    int t = item_begin - it.pp * tiles_total;
    it.e = 0;
    while (it.e < p.E && t >= v3_tiles_of(s_off, it.e)) {
        t -= v3_tiles_of(s_off, it.e);
        ++it.e;
    }
    it.t = t;
}
// returns false when the range is exhausted
__device__ __forceinline__ bool v3_iter_next(V3Iter& it, const RolloutP& p, const int* s_off, V3Pair& pr) {
    if (it.item >= it.item_end) return false;
    while (it.t >= v3_tiles_of(s_off, it.e)) {  // skip exhausted / empty members
        it.t = 0;
        if (++it.e >= p.E) {
            it.e = 0;
            ++it.pp;
        }
    }
    const int nt = v3_tiles_of(s_off, it.e);
    const int cnt = s_off[it.e + 1] - s_off[it.e];
    pr.pp = it.pp;
    pr.net = it.pp * p.E + it.e;
    pr.n = (it.t + 1 < nt && it.item + 1 < it.item_end) ? 2 : 1;
    for (int s = 0; s < pr.n; ++s) {
        pr.row0[s] = s_off[it.e] + (it.t + s) * V3_TILE;
        pr.nrows[s] = min(V3_TILE, cnt - (it.t + s) * V3_TILE);
    }
    it.t += pr.n;
    it.item += pr.n;
    return true;
}

// MMA issuer role of tile slot SLOT.  A function template (slot = compile-time constant) so that every descriptor is
// provably warp-uniform: ptxas then keeps them in uniform registers (UIADD3) instead of moving vector registers to
// uniform ones (R2UR) in front of every UTCHMMA -- the issue path is a single thread's dependent instruction chain
// and directly bounds the MMA rate (54-cycle MMAs).
template <int SLOT>
__device__ __forceinline__ void v3_mma_role(const RolloutP& p, const V3Smem& sm, uint8_t* smem, uint64_t* bars, const int* s_off,
                                            uint32_t tmem_base, uint32_t rank, int lane, int item_begin, int item_end,
                                            int tiles_total) {
    constexpr int s = SLOT;
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    const int K0 = p.layer_k[0];
    const int NH = p.layer_n[L];
        if (rank == 0) {
            int cur_net = -1;
            uint32_t n_w = 0, n_pair = 0, n_a = 0, n_h = 0, n_x[2] = {0, 0};
            const uint32_t idesc_h = umma_idesc_f16(V3_TILE, HP), idesc_head = umma_idesc_f16(V3_TILE, NH);
            const uint32_t b_sbo0 = (uint32_t)(K0 / 8) * 128, b_sbo = (uint32_t)(HP / 8) * 128;
            const uint64_t ad_a = umma_desc(smem_u32(smem + sm.a_off[s]), V3_CHUNK, 128);
            const uint32_t a_hi = (uint32_t)(ad_a >> 32);
            V3Pair pr;
            V3Iter it;
            v3_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
            for (; v3_iter_next(it, p, s_off, pr); ++n_pair) {
                if (pr.net != cur_net) {
                    if (cur_net != -1) v3_commit_uni(&bars[V3_W_EMPTY]);
                    mbar_wait(&bars[V3_W_READY], n_w & 1);
                    ++n_w;
                    cur_net = pr.net;
                }
                if (s >= pr.n) continue;
                const int st = n_pair & 1;
                // ---- layer 0 ------------------------------------------------------------------------------------------
                {
                    mbar_wait(&bars[V3_X_READY + st * 2 + s], n_x[st] & 1);
                    ++n_x[st];
                    tc_fence_after();
                    const uint64_t ad0 = umma_desc(smem_u32(smem + sm.x_off[st][s]), V3_CHUNK, 128);
                    const uint64_t bd0 = umma_desc(smem_u32(smem + sm.w_off[0]), 128, b_sbo0);
                    uint32_t a_lo = (uint32_t)ad0, b_lo = (uint32_t)bd0;
                    const uint32_t b_hi = (uint32_t)(bd0 >> 32);
                    for (int ks = 0; ks < K0 / 16; ++ks, a_lo += 2 * V3_CHUNK / 16, b_lo += 16)
                        v3_mma_uni(tmem_base + v3_d_col(s), a_lo, (uint32_t)(ad0 >> 32), b_lo, b_hi, idesc_h, ks > 0);
                    v3_commit_uni(&bars[V3_X_EMPTY + st * 2 + s]);
                    v3_commit_uni(&bars[V3_D_FULL + s]);
                }
                // ---- layers 1..L (L = head) ------------------------------------------------------------------------------
                for (int l = 1; l <= L; ++l) {
                    const bool head = (l == L);
                    const uint64_t bd0 = umma_desc(smem_u32(smem + sm.w_off[l]), 128, b_sbo);
                    mbar_wait(&bars[V3_A_READY + s], n_a & 1);
                    ++n_a;
                    if (head) {  // the head accumulator of this slot must have been drained by the head warps
                        mbar_wait(&bars[V3_DH_EMPTY + s], (n_h & 1) ^ 1);
                        ++n_h;
                    }
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (head ? v3_dh_col(s) : v3_d_col(s));
                    const uint32_t idesc = head ? idesc_head : idesc_h;
                    uint32_t a_lo = (uint32_t)ad_a, b_lo = (uint32_t)bd0;
                    const uint32_t b_hi = (uint32_t)(bd0 >> 32);
                    const int nks = HP / 16;
                    v3_mma_uni(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, 0u);
                    for (int ks = 1; ks < nks; ++ks) {
                        a_lo += 2 * V3_CHUNK / 16;
                        b_lo += 16;
                        v3_mma_uni(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
                    }
                    v3_commit_uni(head ? &bars[V3_DH_FULL + s] : &bars[V3_D_FULL + s]);
                }
            }
        }
}

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(V3_THREADS, 1)
    mlp_rollout_f16_v3_kernel(const RolloutP p, const V3Smem sm) {
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

    for (int i = threadIdx.x; i <= p.E; i += blockDim.x) s_off[i] = p.offsets[i];
    if (threadIdx.x == 0) {
        mbar_init(&bars[V3_W_FULL], 1);
        mbar_init(&bars[V3_W_READY], 2);
        mbar_init(&bars[V3_W_EMPTY], 2);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars[V3_X_READY + i], 2);
            mbar_init(&bars[V3_X_EMPTY + i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[V3_A_READY + i], 2 * V3_EPI_WARPS);
            mbar_init(&bars[V3_D_FULL + i], 1);
            mbar_init(&bars[V3_DH_FULL + i], 1);
            mbar_init(&bars[V3_DH_EMPTY + i], 2 * 4);
        }
        fence_barrier_init();
    }
    if (warp == V3_WARP_MMA) {
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
    for (int e = 0; e < p.E; ++e) tiles_total += (s_off[e + 1] - s_off[e] + V3_TILE - 1) / V3_TILE;
    const int n_items = p.P * tiles_total;
    const int n_clusters = gridDim.x / 2;
    const int cid = blockIdx.x / 2;
    const int item_begin = (int)((long long)cid * n_items / n_clusters);
    const int item_end = (int)((long long)(cid + 1) * n_items / n_clusters);

    if (warp == V3_WARP_W) {
        // ===================== weight producer ===================================================================
        if (lane == 0) {
            int cur_net = -1;
            uint32_t n_w = 0;
            V3Pair pr;
            V3Iter it;
            v3_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
            while (v3_iter_next(it, p, s_off, pr)) {
                if (pr.net == cur_net) continue;
                if (cur_net != -1) mbar_wait(&bars[V3_W_EMPTY], (n_w - 1) & 1);
                const __half* blob = p.packed + (long long)pr.net * p.blob_halfs;
                uint32_t total = 0;
                for (int l = 0; l <= L; ++l) total += (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l] * 2;
                mbar_expect_tx(&bars[V3_W_FULL], total);
                for (int l = 0; l <= L; ++l) {
                    const uint32_t half_halfs = (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l];
                    tma_bulk_g2s(smem + sm.w_off[l], blob + p.layer_off[l] + (size_t)rank * half_halfs, half_halfs * 2,
                                 &bars[V3_W_FULL]);
                }
                mbar_wait(&bars[V3_W_FULL], n_w & 1);
                v3_arrive_leader(&bars[V3_W_READY]);
                ++n_w;
                cur_net = pr.net;
            }
        }
    } else if (warp == V3_WARP_MMA) {
        v3_mma_role<0>(p, sm, smem, bars, s_off, tmem_base, rank, lane, item_begin, item_end, tiles_total);
    } else if (warp == V3_WARP_MMA + 1) {
        v3_mma_role<1>(p, sm, smem, bars, s_off, tmem_base, rank, lane, item_begin, item_end, tiles_total);
    } else if (warp == V3_WARP_X) {
        // ===================== layer-0 operand loader: two rows per lane, one pair ahead ===============================
        const int in = p.O + p.A;
        uint32_t n_pair = 0, n_x[4] = {0, 0, 0, 0};
        V3Pair pr;
        V3Iter it;
        v3_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
        for (; v3_iter_next(it, p, s_off, pr); ++n_pair) {
            const int st = n_pair & 1;
            for (int s = 0; s < pr.n; ++s) {
                mbar_wait(&bars[V3_X_EMPTY + st * 2 + s], (n_x[st * 2 + s] & 1) ^ 1);
                ++n_x[st * 2 + s];
                uint8_t* xbuf = smem + sm.x_off[st][s];
#pragma unroll
                for (int hr = 0; hr < 2; ++hr) {
                    const int r = lane + hr * 32;
                    const int grow = (int)rank * V3_ROWS + r;
                    const bool valid = grow < pr.nrows[s];
                    const long long env = valid ? p.perm[pr.row0[s] + grow] : 0;
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
                        *reinterpret_cast<uint4*>(xbuf + c * V3_CHUNK + r * 16) =
                            make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) v3_arrive_leader(&bars[V3_X_READY + st * 2 + s]);
            }
        }
    } else if (warp >= V3_HEAD_WARP0) {
        // ===================== head warps: Gaussian head + residual + sampling + scatter ============================
        // accumulator of the head (N_head columns, interleaved (mean_d, logvar_d)): lanes 0-63 hold columns
        // [0, NH/2), lanes 64-127 columns [NH/2, NH) of rows 0-63
        const int q = warp - V3_HEAD_WARP0;
        const int row = (q & 1) * 32 + lane;
        const int col_half = (q >> 1) * (NH / 2);
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int D = p.head_dim;
        uint32_t n_h[2] = {0, 0};
        V3Pair pr;
        V3Iter it;
        v3_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
        while (v3_iter_next(it, p, s_off, pr)) {
            for (int s = 0; s < pr.n; ++s) {
                const int grow = (int)rank * V3_ROWS + row;
                const bool valid = grow < pr.nrows[s];
                const long long env = valid ? p.perm[pr.row0[s] + grow] : 0;
                mbar_wait(&bars[V3_DH_FULL + s], n_h[s] & 1);
                ++n_h[s];
                tc_fence_after();
                for (int c0 = 0; c0 < NH / 2; c0 += 8) {
                    uint32_t raw[8];
                    __syncwarp();
                    v3_ld8_issue(tmem_base + lane_base + v3_dh_col(s) + c0, raw);
                    v3_ld_wait();
                    if (!valid) continue;
                    if (p.layout == CMRL_HEAD_PARALLEL) {
                        if (col_half + c0 == 0) {
                            float out = __uint_as_float(raw[0]);
                            if (p.residual) out += p.obs[env * p.O + pr.pp];
                            if (p.gaussian && p.noise) {
                                const int bi = p.bounds_n == 1 ? 0 : pr.pp;
                                const float lv = soft_clamp_logvar(__uint_as_float(raw[1]), p.min_logvar[bi], p.max_logvar[bi]);
                                out += sqrtf(expf(lv)) * p.noise[env * D + pr.pp];
                            }
                            p.pred[(long long)(pr.net % p.E) * p.pred_se + env * D + pr.pp] = out;
                        }
                    } else if (p.gaussian) {
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {
                            const int d = (col_half + c0 + j) >> 1;
                            if (d < D) {
                                float m = __uint_as_float(raw[j]);
                                if (p.residual) m += p.obs[env * p.O + d];
                                if (p.noise) {
                                    const int bi = p.bounds_n == 1 ? 0 : d;
                                    const float lv = soft_clamp_logvar(__uint_as_float(raw[j + 1]), p.min_logvar[bi], p.max_logvar[bi]);
                                    m += sqrtf(expf(lv)) * p.noise[env * D + d];
                                }
                                p.pred[(long long)(pr.net % p.E) * p.pred_se + env * D + d] = m;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int d = col_half + c0 + j;
                            if (d < D) {
                                float m = __uint_as_float(raw[j]);
                                if (p.residual) m += p.obs[env * p.O + d];
                                p.pred[(long long)(pr.net % p.E) * p.pred_se + env * D + d] = m;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) v3_arrive_leader(&bars[V3_DH_EMPTY + s]);
            }
        }
    } else {
        // ===================== epilogue warps ============================================================================
        // quarter q: TMEM lanes 32q..32q+31 = rows 32(q&1).. of the tile, accumulator columns of half h = q>>1, i.e.
        // hidden units [h*HP/2, (h+1)*HP/2).  Set s converts the 8-column chunks c = s, s+4, ... : one tcgen05.ld.x8,
        // 8 activations, one 16-byte store = K chunk (h*HP/16 + c) of the next layer's A operand.  All of a warp's loads
        // are issued before the single wait; the loop body is kept to ~4 instructions per element (the epilogue is
        // issue / SFU bound: 16 SFU lanes per SM).
        const int q = warp & 3, cset = warp >> 2;
        const int row = (q & 1) * 32 + lane;
        const int h = q >> 1;
        const int n_chunks = HP / 16;                      // 8-column chunks per half
        const int my_n = (n_chunks - cset + 3) >> 2;       // chunks of this warp: cset, cset+4, ...
        const int kc0 = h * n_chunks + cset;               // K chunk of the first one
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cset * 8;
        const uint32_t a_addr0 = smem_u32(smem + sm.a_off[0]) + (uint32_t)(kc0 * V3_CHUNK + row * 16);
        const uint32_t a_addr1 = smem_u32(smem + sm.a_off[1]) + (uint32_t)(kc0 * V3_CHUNK + row * 16);
        // does this thread own the ones column (fp16 1.0 at hidden unit H, the next layer's bias input)?
        const int ones_chunk = p.H >> 3;
        const int ones_rel = ones_chunk - kc0;
        const bool owns_ones = ones_rel >= 0 && (ones_rel & 3) == 0 && (ones_rel >> 2) < my_n && ones_chunk < (h + 1) * n_chunks;
        const uint32_t ones_off = (uint32_t)(ones_rel * V3_CHUNK + (p.H & 7) * 2);
        uint32_t n_d0 = 0, n_d1 = 0;
        // A warp's chunks are processed in two halves (chunks {0,1} and {2,3}).  tcgen05.wait::ld waits for ALL of the
        // thread's outstanding loads, so the loads of the next half are issued right after the wait for the current
        // one and before its conversion: the TMEM load latency hides behind the SFU-bound conversion, also across the
        // two tiles of the pair (when the other tile's accumulator is already complete).
        uint32_t r0[2][8], r1[2][8];  // half 0 / half 1 staging registers
        auto issue_half = [&](int s, int hf, uint32_t(&dst)[2][8]) {
            const uint32_t d_addr = t_addr + v3_d_col(s) + (uint32_t)hf * 64;
            __syncwarp();
            if (2 * hf < my_n) v3_ld8_issue(d_addr, dst[0]);
            if (2 * hf + 1 < my_n) v3_ld8_issue(d_addr + 32, dst[1]);
        };
        auto convert_half = [&](int s, int hf, const uint32_t(&src)[2][8]) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (2 * hf + i < my_n) {
                    uint32_t hh[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) hh[j] = v3_act_pack<ACT>(__uint_as_float(src[i][2 * j]), __uint_as_float(src[i][2 * j + 1]));
                    v3_sts128((s ? a_addr1 : a_addr0) + (uint32_t)((2 * hf + i) * 4 * V3_CHUNK), hh[0], hh[1], hh[2], hh[3]);
                    // the ones column (fp16 1.0 = the next layer's bias input) goes out with the chunk that holds it, so it
                    // is in place before that chunk's half is published
                    if (owns_ones && (ones_rel >> 2) == 2 * hf + i) v3_sts16((s ? a_addr1 : a_addr0) + ones_off, (uint16_t)0x3C00);
                }
            }
        };
        V3Pair pr;
        V3Iter it;
        v3_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
        while (v3_iter_next(it, p, s_off, pr)) {
            const int n_steps = L * pr.n;  // tile-step ts -> layer ts / n, slot ts % n
            bool have0 = false;            // half-0 loads of the current tile-step already in flight?
            for (int ts = 0; ts < n_steps; ++ts) {
                const int s = (pr.n == 2) ? (ts & 1) : 0;
                // ---- half 0 ---------------------------------------------------------------------------------------------
                if (!have0) {
                    if (s) {
                        mbar_wait(&bars[V3_D_FULL + 1], n_d1 & 1);
                        ++n_d1;
                    } else {
                        mbar_wait(&bars[V3_D_FULL + 0], n_d0 & 1);
                        ++n_d0;
                    }
                    tc_fence_after();
                    issue_half(s, 0, r0);
                }
                v3_ld_wait();
                issue_half(s, 1, r1);
                convert_half(s, 0, r0);
                // ---- half 1 (prefetch the other tile's half 0 if its accumulator is already complete) -----------------
                v3_ld_wait();
                have0 = false;
                if (pr.n == 2 && ts + 1 < n_steps) {
                    const int s2 = s ^ 1;
                    if (s2 ? mbar_test(&bars[V3_D_FULL + 1], n_d1 & 1) : mbar_test(&bars[V3_D_FULL + 0], n_d0 & 1)) {
                        if (s2) ++n_d1; else ++n_d0;
                        tc_fence_after();
                        issue_half(s2, 0, r0);
                        have0 = true;
                    }
                }
                convert_half(s, 1, r1);
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) v3_arrive_leader(&bars[V3_A_READY + s]);
            }
        }
    }

    // ---- teardown ----------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    v3_cluster_sync();
    if (warp == V3_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

int launch_rollout_v3(const RolloutP& p, cudaStream_t st) {
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    CMRL_REQUIRE(HP <= 256 && HP % 16 == 0, "mlp_rollout_f16 v3: padded hidden width must be <= 256");
    CMRL_REQUIRE(p.layer_n[L] <= 128, "mlp_rollout_f16 v3: head wider than 128 columns");
    V3Smem sm{};
    uint32_t off = 0;
    for (int l = 0; l <= L; ++l) {
        sm.w_off[l] = off;
        off += (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l] * 2;
        off = (off + 127) / 128 * 128;
    }
    for (int s = 0; s < 2; ++s) {
        sm.a_off[s] = off;
        off += (uint32_t)(HP / 8) * V3_CHUNK;
    }
    for (int stg = 0; stg < 2; ++stg)
        for (int s = 0; s < 2; ++s) {
            sm.x_off[stg][s] = off;
            off += (uint32_t)(p.layer_k[0] / 8) * V3_CHUNK;
        }
    sm.bar_off = off;
    off += V3_BAR_COUNT * 8;
    sm.total = off;
    CMRL_REQUIRE(sm.total <= 232448 - 1024, "mlp_rollout_f16 v3: hidden width %d needs %u B of shared memory (> 227 KB)", p.H, sm.total);
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v3_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        attr_once.done(attr_dev);
    }
    int max_items = p.P * ((p.N + V3_TILE - 1) / V3_TILE + p.E);
    int clusters = num_sms() / 2;
    if (clusters > (max_items + 1) / 2) clusters = (max_items + 1) / 2;
    if (clusters < 1) clusters = 1;
    if (p.act == CMRL_ACT_SILU)
        mlp_rollout_f16_v3_kernel<CMRL_ACT_SILU><<<clusters * 2, V3_THREADS, sm.total, st>>>(p, sm);
    else if (p.act == CMRL_ACT_SILU_HALF)
        mlp_rollout_f16_v3_kernel<CMRL_ACT_SILU_HALF><<<clusters * 2, V3_THREADS, sm.total, st>>>(p, sm);
    else if (p.act == CMRL_ACT_RELU)
        mlp_rollout_f16_v3_kernel<CMRL_ACT_RELU><<<clusters * 2, V3_THREADS, sm.total, st>>>(p, sm);
    else
        mlp_rollout_f16_v3_kernel<CMRL_ACT_IDENTITY><<<clusters * 2, V3_THREADS, sm.total, st>>>(p, sm);
    CMRL_CHECK_LAUNCH("mlp_rollout_f16_v3");
    return 0;
}

}  // namespace tc
}  // namespace cmrl
