// Tensor-core tier, v5: weight-stationary CTA pairs, TWO 256-row tiles in flight, part of the A operand in tensor memory.
//
// v4 (one 256-row tile per pair, MMAs chasing the epilogue) spends ~1 000 of ~3 700 cycles per layer in the hand-off
// round trip epilogue -> remote mbarrier -> MMA issue -> MMA -> commit -> tcgen05.ld, which nothing overlaps because
// layer l+1 of a tile needs all of layer l.  v3 overlapped that with a second tile, but at M=128 per MMA (operand
// bound, ~100 cycles for half the rows).  Two M=256 tiles need 2 x 52 KB of A operand next to 137 KB of resident
// weights -- more shared memory than an SM has.  Here the first KT K steps of each tile's A operand live in TENSOR
// MEMORY (tcgen05.mma with the A operand in TMEM; the epilogue writes them with tcgen05.st), the remaining K steps in
// shared memory.  For H=200 (HP=208, 13 K steps): accumulators 2 x 208 columns, A operand 2 x 32 columns (KT=4), head
// accumulators 2 x 16 columns = 512 TMEM columns; shared memory 133 KB weights + 2 x 36 KB A + 8 KB layer-0 operands.
//
// Schedule: while the 16 epilogue warps convert slot a's layer l (SFU bound, ~1 820 cycles), the tensor pipe runs slot
// b's layer l (13 MMAs x 105 cycles) and the hand-off latency of slot a hides behind slot b's epilogue.
//
// Warp roles per CTA (768 threads): 0-15 epilogue (TMEM lane quarter = warp % 4, set = warp / 4 takes the 16-column
// chunks c = set + 4i), 16-19 head epilogue (Gaussian head, residual, sampling, scatter; one row per lane), 20/21 MMA
// issuers of tile slot 0/1 (leader CTA; warp 20 owns the TMEM allocation), 22 weight TMA producer, 23 layer-0 operand
// loader (four rows per lane, one tile pair ahead).
#include "tc_common.cuh"

namespace cmrl {
namespace tc {

constexpr int V5_THREADS = 768;
constexpr int V5_TILE = 256;            // rows per tile per CTA pair
constexpr int V5_ROWS = 128;            // rows per CTA
constexpr int V5_CHUNK = V5_ROWS * 16;  // bytes of one 8-element K chunk of the A operand
constexpr int V5_EPI_WARPS = 16, V5_HEAD_WARP0 = 16, V5_WARP_MMA = 20 /* and 21 */, V5_WARP_W = 22, V5_WARP_X = 23;

struct V5Lay {
    uint32_t w_off[8];
    uint32_t a_off[2];   // shared-memory part of the A operand per slot: K steps kt .. HP/16-1
    uint32_t x_off[2];   // layer-0 operand per slot
    uint32_t bar_off;
    uint32_t total;
    uint32_t d_col[2], at_col[2], dh_col[2];  // TMEM columns: accumulator, A operand (K steps 0..kt-1), head accumulator
    int kt;
};

enum {
    V5_W_FULL = 0,     // local (tx)
    V5_W_READY = 1,    // leader, count 2
    V5_W_EMPTY = 2,    // both (one commit per MMA issuer -> count 2)
    V5_X_READY = 3,    // leader [slot], count 2 (one loader warp per CTA)
    V5_X_EMPTY = 5,    // both   [slot] (commit)
    V5_A_READY = 7,    // leader [slot], count 2 * 16 warps
    V5_D_FULL = 9,     // both   [slot] (commit)
    V5_DH_FULL = 11,   // both   [slot] (commit)
    V5_DH_EMPTY = 13,  // leader [slot], count 2 * 4 head warps
    V5_BAR_COUNT = 15,
};

// The work list of a cluster is a contiguous range of 256-row tiles in (p, member, tile) order; consecutive tiles of
// the same net are processed two at a time.  Every role walks the list with this iterator.
struct V5Pair {
    int net, pp, n;  // net index, sub-net index, tiles in this pair (1 or 2)
    int row0[2], nrows[2];
};
struct V5Iter {
    int item, item_end, pp, e, t;
};
__device__ __forceinline__ int v5_tiles_of(const int* s_off, int e) { return (s_off[e + 1] - s_off[e] + V5_TILE - 1) / V5_TILE; }
__device__ __forceinline__ void v5_iter_init(V5Iter& it, const RolloutP& p, int item_begin, int item_end, int tiles_total,
                                             const int* s_off) {
    it.item = item_begin;
    it.item_end = item_end;
    it.pp = tiles_total > 0 ? item_begin / tiles_total : 0;
    int t = item_begin - it.pp * tiles_total;
    it.e = 0;
    while (it.e < p.E && t >= v5_tiles_of(s_off, it.e)) {
        t -= v5_tiles_of(s_off, it.e);
        ++it.e;
    }
    it.t = t;
}
__device__ __forceinline__ bool v5_iter_next(V5Iter& it, const RolloutP& p, const int* s_off, V5Pair& pr) {
    if (it.item >= it.item_end) return false;
    while (it.t >= v5_tiles_of(s_off, it.e)) {  // skip exhausted / empty members
        it.t = 0;
        if (++it.e >= p.E) {
            it.e = 0;
            ++it.pp;
        }
    }
    const int nt = v5_tiles_of(s_off, it.e);
    const int cnt = s_off[it.e + 1] - s_off[it.e];
    pr.pp = it.pp;
    pr.net = it.pp * p.E + it.e;
    pr.n = (it.t + 1 < nt && it.item + 1 < it.item_end) ? 2 : 1;
    for (int s = 0; s < pr.n; ++s) {
        pr.row0[s] = s_off[it.e] + (it.t + s) * V5_TILE;
        pr.nrows[s] = min(V5_TILE, cnt - (it.t + s) * V5_TILE);
    }
    it.t += pr.n;
    it.item += pr.n;
    return true;
}

// MMA issuer of tile slot SLOT (compile-time slot: every descriptor is provably warp-uniform, see mlp_tc3.cu)
template <int SLOT>
__device__ __forceinline__ void v5_mma_role(const RolloutP& p, const V5Lay& lay, uint8_t* smem, uint64_t* bars, const int* s_off,
                                            uint32_t tmem_base, uint32_t rank, int lane, int item_begin, int item_end,
                                            int tiles_total) {
    constexpr int s = SLOT;
    if (rank != 0) return;
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    const int K0 = p.layer_k[0];
    const int NH = p.layer_n[L];
    const int n_ks = HP / 16, kt = lay.kt;
    const uint32_t ns_mma = 0u;
    int cur_net = -1;
    uint32_t n_w = 0, n_a = 0, n_h = 0, n_x = 0;
    const uint32_t idesc_h = umma_idesc_f16(V5_TILE, HP), idesc_head = umma_idesc_f16(V5_TILE, NH);
    const uint32_t b_sbo0 = (uint32_t)(K0 / 8) * 128, b_sbo = (uint32_t)(HP / 8) * 128;
    const uint64_t ad_a = umma_desc(smem_u32(smem + lay.a_off[s]), V5_CHUNK, 128);
    const uint32_t a_hi = (uint32_t)(ad_a >> 32);
    const uint32_t d_t = tmem_base + lay.d_col[s], dh_t = tmem_base + lay.dh_col[s], a_t = tmem_base + lay.at_col[s];
    V5Pair pr;
    V5Iter it;
    v5_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
    while (v5_iter_next(it, p, s_off, pr)) {
        if (pr.net != cur_net) {
            if (cur_net != -1) v3_commit_uni(&bars[V5_W_EMPTY]);
            mbar_wait_sleep(&bars[V5_W_READY], n_w & 1, ns_mma);
            ++n_w;
            cur_net = pr.net;
        }
        if (s >= pr.n) continue;
        // ---- layer 0 ------------------------------------------------------------------------------------------------
        {
            mbar_wait_sleep(&bars[V5_X_READY + s], n_x & 1, ns_mma);
            ++n_x;
            tc_fence_after();
            const uint64_t ad0 = umma_desc(smem_u32(smem + lay.x_off[s]), V5_CHUNK, 128);
            const uint64_t bd0 = umma_desc(smem_u32(smem + lay.w_off[0]), 128, b_sbo0);
            uint32_t a_lo = (uint32_t)ad0, b_lo = (uint32_t)bd0;
            const uint32_t b_hi = (uint32_t)(bd0 >> 32);
#pragma unroll 1
            for (int ks = 0; ks < K0 / 16; ++ks, a_lo += 2 * V5_CHUNK / 16, b_lo += 16)
                v3_mma_uni(d_t, a_lo, (uint32_t)(ad0 >> 32), b_lo, b_hi, idesc_h, ks > 0);
            v3_commit_uni(&bars[V5_X_EMPTY + s]);
            v3_commit_uni(&bars[V5_D_FULL + s]);
        }
        // ---- layers 1..L (L = head): K steps 0..kt-1 read A from tensor memory, the rest from shared memory ----------
        for (int l = 1; l <= L; ++l) {
            const bool head = (l == L);
            const uint64_t bd0 = umma_desc(smem_u32(smem + lay.w_off[l]), 128, b_sbo);
            mbar_wait_sleep(&bars[V5_A_READY + s], n_a & 1, ns_mma);
            ++n_a;
            if (head) {  // the head accumulator of this slot must have been drained by the head warps
                mbar_wait_sleep(&bars[V5_DH_EMPTY + s], (n_h & 1) ^ 1, ns_mma);
                ++n_h;
            }
            tc_fence_after();
            const uint32_t d_tmem = head ? dh_t : d_t;
            const uint32_t idesc = head ? idesc_head : idesc_h;
            uint32_t a_lo = (uint32_t)ad_a, b_lo = (uint32_t)bd0, at = a_t;
            const uint32_t b_hi = (uint32_t)(bd0 >> 32);
            int ks = 0;
            // rolled loops: the issue path has slack (105-cycle MMAs) and the kernel has to stay small (instruction caches)
#pragma unroll 1
            for (; ks < kt; ++ks, at += 8, b_lo += 16) v5_mma_ts_uni(d_tmem, at, b_lo, b_hi, idesc, ks > 0);
#pragma unroll 1
            for (; ks < n_ks; ++ks, a_lo += 2 * V5_CHUNK / 16, b_lo += 16) v3_mma_uni(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, ks > 0);
            v3_commit_uni(head ? &bars[V5_DH_FULL + s] : &bars[V5_D_FULL + s]);
        }
    }
}

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(V5_THREADS, 1)
    mlp_rollout_f16_v5_kernel(const RolloutP p, const V5Lay lay) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_off[65];
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    const int K0 = p.layer_k[0];
    const int NH = p.layer_n[L];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
    const uint32_t rank = v3_ctarank();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t ns_slack = 0u;  // poll interval of the roles with slack (head, weights, layer-0 loader)

    for (int i = threadIdx.x; i <= p.E; i += blockDim.x) s_off[i] = p.offsets[i];
    if (threadIdx.x == 0) {
        mbar_init(&bars[V5_W_FULL], 1);
        mbar_init(&bars[V5_W_READY], 2);
        mbar_init(&bars[V5_W_EMPTY], 2);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[V5_X_READY + i], 2);
            mbar_init(&bars[V5_X_EMPTY + i], 1);
            mbar_init(&bars[V5_A_READY + i], 2 * V5_EPI_WARPS);
            mbar_init(&bars[V5_D_FULL + i], 1);
            mbar_init(&bars[V5_DH_FULL + i], 1);
            mbar_init(&bars[V5_DH_EMPTY + i], 2 * 4);
        }
        fence_barrier_init();
    }
    if (warp == V5_WARP_MMA) {
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
    for (int e = 0; e < p.E; ++e) tiles_total += v5_tiles_of(s_off, e);
    const int n_items = p.P * tiles_total;
    const int n_clusters = gridDim.x / 2;
    const int cid = blockIdx.x / 2;
    const int item_begin = (int)((long long)cid * n_items / n_clusters);
    const int item_end = (int)((long long)(cid + 1) * n_items / n_clusters);

    if (warp == V5_WARP_W) {
        // ===================== weight producer: this CTA's half of every layer of the current net =====================
        if (lane == 0) {
            int cur_net = -1;
            uint32_t n_w = 0;
            V5Pair pr;
            V5Iter it;
            v5_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
            while (v5_iter_next(it, p, s_off, pr)) {
                if (pr.net == cur_net) continue;
                if (cur_net != -1) mbar_wait_sleep(&bars[V5_W_EMPTY], (n_w - 1) & 1, ns_slack);
                const __half* blob = p.packed + (long long)pr.net * p.blob_halfs;
                uint32_t total = 0;
                for (int l = 0; l <= L; ++l) total += (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l] * 2;
                mbar_expect_tx(&bars[V5_W_FULL], total);
                for (int l = 0; l <= L; ++l) {
                    const uint32_t half_halfs = (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l];
                    tma_bulk_g2s(smem + lay.w_off[l], blob + p.layer_off[l] + (size_t)rank * half_halfs, half_halfs * 2,
                                 &bars[V5_W_FULL]);
                }
                mbar_wait_sleep(&bars[V5_W_FULL], n_w & 1, ns_slack);
                v3_arrive_leader(&bars[V5_W_READY]);
                ++n_w;
                cur_net = pr.net;
            }
        }
    } else if (warp == V5_WARP_MMA) {
        v5_mma_role<0>(p, lay, smem, bars, s_off, tmem_base, rank, lane, item_begin, item_end, tiles_total);
    } else if (warp == V5_WARP_MMA + 1) {
        v5_mma_role<1>(p, lay, smem, bars, s_off, tmem_base, rank, lane, item_begin, item_end, tiles_total);
    } else if (warp == V5_WARP_X) {
        // ===================== layer-0 operand loader: four rows per lane, one pair ahead ==============================
        const int in = p.O + p.A;
        uint32_t n_x[2] = {0, 0};
        V5Pair pr;
        V5Iter it;
        v5_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
        while (v5_iter_next(it, p, s_off, pr)) {
            for (int s = 0; s < pr.n; ++s) {
                mbar_wait_sleep(&bars[V5_X_EMPTY + s], (n_x[s] & 1) ^ 1, ns_slack);
                ++n_x[s];
                uint8_t* xbuf = smem + lay.x_off[s];
                long long envs[4];
#pragma unroll
                for (int hr = 0; hr < 4; ++hr) {  // the four index loads first (independent), then the rows one by one
                    const int grow = (int)rank * V5_ROWS + lane + hr * 32;
                    envs[hr] = grow < pr.nrows[s] ? (long long)p.perm[pr.row0[s] + grow] : -1;
                }
#pragma unroll
                for (int hr = 0; hr < 4; ++hr) {
                    const int r = lane + hr * 32;
                    const long long env = envs[hr];
#pragma unroll 1
                    for (int c = 0; c < K0 / 8; ++c) {
                        uint32_t h2[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float v[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int k = c * 8 + 2 * j + u;
                                float x = 0.0f;
                                if (env >= 0) {
                                    if (k < in) x = k < p.O ? p.obs[env * p.O + k] : p.actn[env * p.A + (k - p.O)];
                                    else if (k == in) x = 1.0f;
                                }
                                v[u] = x;
                            }
                            h2[j] = pack_h2(v[0], v[1]);
                        }
                        *reinterpret_cast<uint4*>(xbuf + c * V5_CHUNK + r * 16) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) v3_arrive_leader(&bars[V5_X_READY + s]);
            }
        }
    } else if (warp >= V5_HEAD_WARP0) {
        // ===================== head warps: one row per lane ================================================================
        const int q = warp - V5_HEAD_WARP0;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int D = p.head_dim;
        const uint32_t dh0 = lay.dh_col[0], dh1 = lay.dh_col[1];
        uint32_t hph = 0;  // phase parity of DH_FULL[slot] in bit `slot`
        V5Pair pr;
        V5Iter it;
        v5_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
        while (v5_iter_next(it, p, s_off, pr)) {
            for (int s = 0; s < pr.n; ++s) {
                const int grow = (int)rank * V5_ROWS + row;
                const bool valid = grow < pr.nrows[s];
                const long long env = valid ? p.perm[pr.row0[s] + grow] : 0;
                float res0 = 0.0f, nz0 = 0.0f;
                if (p.layout == CMRL_HEAD_PARALLEL && valid) {  // prefetch before the wait
                    if (p.residual) res0 = p.obs[env * p.O + pr.pp];
                    if (p.gaussian && p.noise) nz0 = p.noise[env * D + pr.pp];
                }
                mbar_wait_sleep(&bars[V5_DH_FULL + s], (hph >> s) & 1, ns_slack);
                hph ^= 1u << s;
                tc_fence_after();
                const uint32_t h_addr = tmem_base + lane_base + (s ? dh1 : dh0);
                // one output per iteration (rolled: the head has a whole tile time of slack, code size has not)
                const int n_out = p.layout == CMRL_HEAD_PARALLEL ? 1 : D;
                const int stride = p.gaussian ? 2 : 1;
#pragma unroll 1
                for (int d = 0; d < n_out; ++d) {
                    uint32_t r0, r1 = 0;
                    __syncwarp();
                    if (p.gaussian)
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(h_addr + (uint32_t)(d * stride)));
                    else
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(h_addr + (uint32_t)d));
                    v3_ld_wait();
                    if (valid) {
                        const int od = p.layout == CMRL_HEAD_PARALLEL ? pr.pp : d;
                        float m = __uint_as_float(r0);
                        if (p.layout == CMRL_HEAD_PARALLEL) m += res0;
                        else if (p.residual) m += p.obs[env * p.O + d];
                        if (p.gaussian && p.noise) {
                            const int bi = p.bounds_n == 1 ? 0 : od;
                            const float lv = soft_clamp_logvar(__uint_as_float(r1), p.min_logvar[bi], p.max_logvar[bi]);
                            m += sqrtf(expf(lv)) * (p.layout == CMRL_HEAD_PARALLEL ? nz0 : p.noise[env * D + d]);
                        }
                        p.pred[(long long)(pr.net % p.E) * p.pred_se + env * D + od] = m;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) v3_arrive_leader(&bars[V5_DH_EMPTY + s]);
            }
        }
    } else {
        // ===================== epilogue warps ====================================================================================
        // quarter q = TMEM lanes 32q.. = rows of the tile; set cset converts the 16-column chunks c = cset + 4i: one
        // tcgen05.ld.x16, 16 activations, eight packed registers = K step c of the next layer's A operand, written with one
        // tcgen05.st.x8 (c < kt) or two 16-byte shared-memory stores.
        //
        // The loop is deliberately minimal -- one register buffer, no software pipelining, rolled: the 16 SFU lanes of the SM
        // are the shared bottleneck (4 warps per sub-partition, 128 of every 512 cycles each), so a warp's own TMEM-load
        // latency hides behind the other three warps, while a large unrolled body (several KB per stage) thrashes the 6 KB
        // L0 instruction cache with 16 warps at different points of it (measured: 2x slower).
        const int q = warp & 3, cset = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int n_chunks = HP / 16;
        const int kt = lay.kt;
        const uint32_t d_addr0 = tmem_base + lane_off + lay.d_col[0];
        const uint32_t d_addr1 = tmem_base + lane_off + lay.d_col[1];
        const uint32_t at_addr0 = tmem_base + lane_off + lay.at_col[0];
        const uint32_t at_addr1 = tmem_base + lane_off + lay.at_col[1];
        const uint32_t as_addr0 = smem_u32(smem + lay.a_off[0]) + (uint32_t)(row * 16);
        const uint32_t as_addr1 = smem_u32(smem + lay.a_off[1]) + (uint32_t)(row * 16);
        // ones column (fp16 1.0 at hidden unit H = the next layer's bias input); its K step is kept in shared memory
        // (kt <= H / 16, see the launcher) so that it is one 16-bit store after the chunk that holds it
        const int ones_c = p.H >> 4;
        const uint32_t ones_off = (uint32_t)((2 * (ones_c - kt) + ((p.H >> 3) & 1)) * V5_CHUNK + (p.H & 7) * 2);
        uint32_t dph = 0;  // phase parity of D_FULL[slot] in bit `slot`

        V5Pair pr;
        V5Iter it;
        v5_iter_init(it, p, item_begin, item_end, tiles_total, s_off);
        while (v5_iter_next(it, p, s_off, pr)) {
            const int n_steps = L * pr.n;  // tile-step ts -> layer ts / n, slot ts % n
#pragma unroll 1
            for (int ts = 0; ts < n_steps; ++ts) {
                const int s = (pr.n == 2) ? (ts & 1) : 0;
                mbar_wait(&bars[V5_D_FULL + s], (dph >> s) & 1);
                dph ^= 1u << s;
                tc_fence_after();
                const uint32_t d_addr = s ? d_addr1 : d_addr0, at_addr = s ? at_addr1 : at_addr0, as_addr = s ? as_addr1 : as_addr0;
#pragma unroll 1
                for (int c = cset; c < n_chunks; c += 4) {
                    uint32_t raw[16], hh[8];
                    __syncwarp();
                    v5_ld16_issue(d_addr + (uint32_t)c * 16, raw);
                    v3_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) hh[j] = v3_act_pack<ACT>(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
                    if (c < kt) {
                        v5_st8(at_addr + (uint32_t)c * 8, hh);
                    } else {
                        const uint32_t dst = as_addr + (uint32_t)((c - kt) * 2 * V5_CHUNK);
                        v3_sts128(dst, hh[0], hh[1], hh[2], hh[3]);
                        v3_sts128(dst + V5_CHUNK, hh[4], hh[5], hh[6], hh[7]);
                        if (c == ones_c) v3_sts16(as_addr + ones_off, (uint16_t)0x3C00);
                    }
                }
                v5_st_wait();
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) v3_arrive_leader(&bars[V5_A_READY + s]);
            }
        }
    }

    // ---- teardown ------------------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    v3_cluster_sync();
    if (warp == V5_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// tensor-memory and shared-memory layout for this model; false when it does not fit (the caller then uses mlp_tc3.cu)
static bool v5_layout(const RolloutP& p, V5Lay& lay) {
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    const int NH = p.layer_n[L];
    if (HP % 16 != 0 || NH % 16 != 0 || 2 * HP + 2 * NH > 512) return false;
    const int n_ks = HP / 16;
    // TMEM columns: [D slot 0 | D slot 1 | A (K steps 0..kt-1) slot 0 | slot 1 | head slot 0 | slot 1]
    const int d1 = HP;
    int kt = (512 - d1 - HP - 2 * NH) / 16;  // A-operand K steps per slot that fit the remaining TMEM columns
    if (kt > n_ks) kt = n_ks;
    if (kt > p.H / 16) kt = p.H / 16;  // the K step that holds the ones column (hidden unit H) stays in shared memory
    if (kt < 0) return false;
    lay.kt = kt;
    lay.d_col[0] = 0;
    lay.d_col[1] = (uint32_t)d1;
    lay.at_col[0] = (uint32_t)(d1 + HP);
    lay.at_col[1] = lay.at_col[0] + 8u * kt;
    lay.dh_col[0] = lay.at_col[0] + 16u * kt;
    lay.dh_col[1] = lay.dh_col[0] + (uint32_t)NH;
    uint32_t off = 0;
    for (int l = 0; l <= L; ++l) {
        lay.w_off[l] = off;
        off += (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l] * 2;
        off = (off + 127) / 128 * 128;
    }
    for (int s = 0; s < 2; ++s) {
        lay.a_off[s] = off;
        off += (uint32_t)(n_ks - kt) * 2 * V5_CHUNK;
    }
    for (int s = 0; s < 2; ++s) {
        lay.x_off[s] = off;
        off += (uint32_t)(p.layer_k[0] / 8) * V5_CHUNK;
    }
    lay.bar_off = off;
    off += V5_BAR_COUNT * 8;
    lay.total = off;
    return lay.total <= 232448 - 1024;
}

bool rollout_v5_supports(const RolloutP& p) {
    V5Lay lay{};
    return v5_layout(p, lay);
}

int launch_rollout_v5(const RolloutP& p, cudaStream_t st) {
    V5Lay lay{};
    CMRL_REQUIRE(v5_layout(p, lay), "mlp_rollout_f16 v5: two 256-row tiles of this model (hidden %d, head %d columns) do not fit tensor / shared memory",
                 p.layer_n[0], p.layer_n[p.num_layers]);
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v5_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v5_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v5_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v5_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        attr_once.done(attr_dev);
    }
    int max_items = p.P * ((p.N + V5_TILE - 1) / V5_TILE + p.E);
    int clusters = num_sms() / 2;
    if (clusters > (max_items + 1) / 2) clusters = (max_items + 1) / 2;
    if (clusters < 1) clusters = 1;
    if (p.act == CMRL_ACT_SILU)
        mlp_rollout_f16_v5_kernel<CMRL_ACT_SILU><<<clusters * 2, V5_THREADS, lay.total, st>>>(p, lay);
    else if (p.act == CMRL_ACT_SILU_HALF)
        mlp_rollout_f16_v5_kernel<CMRL_ACT_SILU_HALF><<<clusters * 2, V5_THREADS, lay.total, st>>>(p, lay);
    else if (p.act == CMRL_ACT_RELU)
        mlp_rollout_f16_v5_kernel<CMRL_ACT_RELU><<<clusters * 2, V5_THREADS, lay.total, st>>>(p, lay);
    else
        mlp_rollout_f16_v5_kernel<CMRL_ACT_IDENTITY><<<clusters * 2, V5_THREADS, lay.total, st>>>(p, lay);
    CMRL_CHECK_LAUNCH("mlp_rollout_f16_v5");
    return 0;
}

}  // namespace tc
}  // namespace cmrl
