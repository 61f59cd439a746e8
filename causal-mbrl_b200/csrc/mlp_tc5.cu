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

// Several mechanisms (transition, reward mech, termination mech) can share ONE launch: the groups' layer tables, hidden
// width and activation agree (checked on the host), everything else -- weights, buckets, noise, head layout, output -- is per
// group.  The work list of the launch is group 0's tiles, then group 1's, ...
constexpr int V5_MAX_GROUPS = 3;
struct V5Groups {
    int G;
    RolloutP g[V5_MAX_GROUPS];
};

// The work list of a cluster is a contiguous range of 256-row tiles in (group, p, member, tile) order; consecutive tiles of
// the same net are processed two at a time.  One thread per CTA cuts the cluster's range into SEGMENTS (a run of tiles of one
// net) in shared memory when the kernel starts; every role then walks the segments with a cursor of two integers.  (The
// general iterator -- groups, sub-nets, empty members -- inlined into seven roles doubled the kernel and cost 20 % through
// the instruction caches; un-inlined it kept its state in local memory, which misses the L1 next to 220 KB of shared memory.)
struct V5Seg {
    int g, pp, netl, net;  // group, sub-net, net index inside the group, net index unique over the launch
    int t0, t1;            // tiles [t0, t1) of the member's tiles belong to this cluster
    int row_base, cnt;     // first bucketed row and number of rows of the member
};
constexpr int V5_MAX_SEGS = 40;
struct V5Pair {
    int g, net, netl, pp, n;  // n: tiles in this pair (1 or 2)
    int row0[2], nrows[2];
};
struct V5Cur {
    int i;
};
__device__ __forceinline__ int v5_tiles_of(const int* s_off, int e) { return (s_off[e + 1] - s_off[e] + V5_TILE - 1) / V5_TILE; }
// One thread per CTA lists the cluster's work when the kernel starts: SEGMENTS (a run of tiles of one net) and, per tile pair,
// one word = segment << 24 | two-tile flag << 23 | first tile.  The roles only index these tables: the kernel's roles must
// stay small -- the general iterator (groups, sub-nets, empty members) inlined into seven roles doubled the kernel and cost
// 20 % through the instruction caches; un-inlined it kept its state in local memory, which misses the L1 next to 220 KB of
// shared memory.
constexpr int V5_MAX_PAIRS = 1024;
// Executed by ONE WARP when the kernel starts (a single thread walking 35 nets cost ~2 us of a 180 us launch): lane i takes
// nets i, i + 32, ...; warp scans give every net its first pair unit, its segment slot and its first pair slot.  The range
// is counted in PAIR units (tiles 2k, 2k+1 of a net): a cluster boundary never splits a pair, so the only single tiles are
// the last ones of nets with an odd tile count.
__device__ __noinline__ void v5_build_work(const V5Groups* gs, const int (*s_off)[65], int n_clusters, int cid, V5Seg* segs,
                                           uint32_t* pairs, int* n_segs_out, int* n_pairs_out, int lane) {
    int total_nets = 0;
    for (int g = 0; g < gs->G; ++g) total_nets += gs->g[g].P * gs->g[g].E;
    auto decode = [&](int id, int& g, int& pp, int& e) {
        g = 0;
        while (id >= gs->g[g].P * gs->g[g].E) {
            id -= gs->g[g].P * gs->g[g].E;
            ++g;
        }
        pp = id / gs->g[g].E;
        e = id - pp * gs->g[g].E;
    };
    // pass 1: pair units of the whole launch
    int mine = 0;
    for (int id = lane; id < total_nets; id += 32) {
        int g, pp, e;
        decode(id, g, pp, e);
        mine += (v5_tiles_of(s_off[g], e) + 1) / 2;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    const int u_begin = (int)((long long)cid * mine / n_clusters), u_end = (int)((long long)(cid + 1) * mine / n_clusters);
    // pass 2: nets in order, 32 at a time
    int cur_base = 0, seg_base = 0, pair_base = 0;
    bool overflow = false;
    for (int id0 = 0; id0 < total_nets; id0 += 32) {
        const int id = id0 + lane;
        int g = 0, pp = 0, e = 0, nt = 0;
        if (id < total_nets) {
            decode(id, g, pp, e);
            nt = v5_tiles_of(s_off[g], e);
        }
        const int nu = (nt + 1) / 2;
        int incl = nu;  // inclusive scan of the units
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int cur = cur_base + incl - nu;
        const int lo = max(u_begin, cur), hi = min(u_end, cur + nu);
        const bool hit = lo < hi;
        const int t0 = hit ? 2 * (lo - cur) : 0, t1 = hit ? min(nt, 2 * (hi - cur)) : 0;
        const int np = (t1 - t0 + 1) / 2;
        const unsigned ball = __ballot_sync(0xffffffffu, hit);
        const int seg = seg_base + __popc(ball & ((1u << lane) - 1u));
        int pincl = np;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, pincl, o);
            if (lane >= o) pincl += v;
        }
        const int pfirst = pair_base + pincl - np;
        if (hit) {
            if (seg >= V5_MAX_SEGS || pfirst + np > V5_MAX_PAIRS) {
                overflow = true;
            } else {
                int net_base = 0;
                for (int h = 0; h < g; ++h) net_base += gs->g[h].P * gs->g[h].E;
                segs[seg].g = g; segs[seg].pp = pp; segs[seg].netl = pp * gs->g[g].E + e; segs[seg].net = net_base + pp * gs->g[g].E + e;
                segs[seg].t0 = t0; segs[seg].t1 = t1;
                segs[seg].row_base = s_off[g][e]; segs[seg].cnt = s_off[g][e + 1] - s_off[g][e];
                int k = pfirst;
                for (int t = t0; t < t1; t += 2) pairs[k++] = ((uint32_t)seg << 24) | ((uint32_t)(t + 1 < t1 ? 1 : 0) << 23) | (uint32_t)t;
            }
        }
        cur_base += __shfl_sync(0xffffffffu, incl, 31);
        seg_base += __popc(ball);
        pair_base += __shfl_sync(0xffffffffu, pincl, 31);
    }
    overflow = __any_sync(0xffffffffu, overflow);
    if (lane == 0) {
        *n_segs_out = overflow ? -1 : seg_base;
        *n_pairs_out = pair_base;
    }
}
__device__ __forceinline__ void v5_cur_init(V5Cur& c, const V5Seg*, int) { c.i = 0; }
__device__ __forceinline__ bool v5_next(V5Cur& c, const V5Seg* segs, int n_pairs, V5Pair& pr, const uint32_t* pairs) {
    if (c.i >= n_pairs) return false;
    const uint32_t w = pairs[c.i++];
    const V5Seg& sg = segs[w >> 24];
    const int t = (int)(w & 0x7FFFFFu);
    pr.g = sg.g; pr.pp = sg.pp; pr.netl = sg.netl; pr.net = sg.net;
    pr.n = 1 + (int)((w >> 23) & 1u);
    pr.row0[0] = sg.row_base + t * V5_TILE;
    pr.row0[1] = pr.row0[0] + V5_TILE;
    pr.nrows[0] = min(V5_TILE, sg.cnt - t * V5_TILE);
    pr.nrows[1] = min(V5_TILE, sg.cnt - (t + 1) * V5_TILE);
    return true;
}

// MMA issuer of tile slot SLOT (compile-time slot: every descriptor is provably warp-uniform, see mlp_tc3.cu)
template <int SLOT>
__device__ __forceinline__ void v5_mma_role(const RolloutP& p, const V5Lay& lay, uint8_t* smem, uint64_t* bars, const V5Seg* segs, int nseg,
                                            const uint32_t* pairs, int npairs, uint32_t tmem_base, uint32_t rank, int lane) {  // p: group 0 (all groups share the layer tables)
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
    V5Cur it;
    v5_cur_init(it, segs, nseg);
    while (v5_next(it, segs, npairs, pr, pairs)) {
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
    mlp_rollout_f16_v5_kernel(const V5Groups gs_param, const V5Lay lay) {
    const RolloutP& p = gs_param.g[0];  // layer tables, hidden width and activation are the same in every group
    // the per-group tables are read through shared memory by the (un-inlined) work-list iterator and the roles with slack: the
    // kernel must stay small (instruction caches: inlining the grouped iterator into seven roles doubled it and cost 20 %)
    __shared__ V5Groups s_gs;
    for (int i = threadIdx.x; i < (int)(sizeof(V5Groups) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&s_gs)[i] = reinterpret_cast<const uint32_t*>(&gs_param)[i];
    __syncthreads();
    const V5Groups* gs = &s_gs;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_off[V5_MAX_GROUPS][65];
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    const int K0 = p.layer_k[0];
    const int NH = p.layer_n[L];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
    const uint32_t rank = v3_ctarank();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t ns_slack = 0u;  // poll interval of the roles with slack (head, weights, layer-0 loader)

    for (int g = 0; g < gs->G; ++g)
        for (int i = threadIdx.x; i <= gs->g[g].E; i += blockDim.x) s_off[g][i] = gs->g[g].offsets[i];
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

    // ---- this cluster's share of the work list, as segments in shared memory (built by one thread) -----------------------------
    __shared__ V5Seg s_segs[V5_MAX_SEGS];
    __shared__ uint32_t s_pairs[V5_MAX_PAIRS];
    __shared__ int s_nseg, s_npairs;
    if (warp == 1) v5_build_work(gs, s_off, gridDim.x / 2, blockIdx.x / 2, s_segs, s_pairs, &s_nseg, &s_npairs, lane);
    __syncthreads();
    const V5Seg* segs = s_segs;
    const uint32_t* pairs = s_pairs;
    const int nseg = s_nseg, npairs = s_npairs;
    if (nseg < 0) __trap();  // more nets per cluster than the segment table holds: never silently drop work

    if (warp == V5_WARP_W) {
        // ===================== weight producer: this CTA's half of every layer of the current net =====================
        if (lane == 0) {
            int cur_net = -1;
            uint32_t n_w = 0;
            V5Pair pr;
            V5Cur it;
            v5_cur_init(it, segs, nseg);
            while (v5_next(it, segs, npairs, pr, pairs)) {
                if (pr.net == cur_net) continue;
                if (cur_net != -1) mbar_wait_sleep(&bars[V5_W_EMPTY], (n_w - 1) & 1, ns_slack);
                const __half* blob = gs->g[pr.g].packed + (long long)pr.netl * gs->g[pr.g].blob_halfs;
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
        v5_mma_role<0>(p, lay, smem, bars, segs, nseg, pairs, npairs, tmem_base, rank, lane);
    } else if (warp == V5_WARP_MMA + 1) {
        v5_mma_role<1>(p, lay, smem, bars, segs, nseg, pairs, npairs, tmem_base, rank, lane);
    } else if (warp == V5_WARP_X) {
        // ===================== layer-0 operand loader: four rows per lane, one pair ahead ==============================
        uint32_t n_x[2] = {0, 0};
        V5Pair pr;
        V5Cur it;
        v5_cur_init(it, segs, nseg);
        while (v5_next(it, segs, npairs, pr, pairs)) {
            // this group's inputs, in registers (the group table lives in shared memory)
            struct {
                const float *obs, *actn;
                const int* perm;
                int O, A;
            } p = {gs->g[pr.g].obs, gs->g[pr.g].actn, gs->g[pr.g].perm, gs->g[pr.g].O, gs->g[pr.g].A};
            const int in = p.O + p.A;
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
        const uint32_t dh0 = lay.dh_col[0], dh1 = lay.dh_col[1];
        uint32_t hph = 0;  // phase parity of DH_FULL[slot] in bit `slot`
        V5Pair pr;
        V5Cur it;
        v5_cur_init(it, segs, nseg);
        while (v5_next(it, segs, npairs, pr, pairs)) {
            // this group's head description, in registers (the group table lives in shared memory)
            const RolloutP& gq = gs->g[pr.g];
            struct {
                const float *obs, *noise, *min_logvar, *max_logvar;
                const int* perm;
                float* pred;
                long long pred_se;
                int O, E, head_dim, layout, gaussian, residual, bounds_n;
            } p = {gq.obs, gq.noise, gq.min_logvar, gq.max_logvar, gq.perm, gq.pred, gq.pred_se,
                   gq.O, gq.E, gq.head_dim, gq.layout, gq.gaussian, gq.residual, gq.bounds_n};
            const int D = p.head_dim;
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
                        p.pred[(long long)(pr.netl % p.E) * p.pred_se + env * D + od] = m;
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
        V5Cur it;
        v5_cur_init(it, segs, nseg);
        while (v5_next(it, segs, npairs, pr, pairs)) {
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
    return lay.total <= 232448 - 7424;  // (7 KB of static shared memory: group tables, bucket offsets, segment and pair tables)
}

bool rollout_v5_supports(const RolloutP& p) {
    V5Lay lay{};
    return v5_layout(p, lay);
}

// one launch for G mechanisms whose layer tables agree (G = 1: the plain rollout of one mechanism)
int launch_rollout_v5_groups(const RolloutP* groups, int G, cudaStream_t st) {
    CMRL_REQUIRE(G >= 1 && G <= V5_MAX_GROUPS, "mlp_rollout_f16 v5: 1..%d mechanisms per launch", V5_MAX_GROUPS);
    const RolloutP& p = groups[0];
    V5Lay lay{};
    CMRL_REQUIRE(v5_layout(p, lay), "mlp_rollout_f16 v5: two 256-row tiles of this model (hidden %d, head %d columns) do not fit tensor / shared memory",
                 p.layer_n[0], p.layer_n[p.num_layers]);
    V5Groups gs{};
    gs.G = G;
    long long max_items = 0;
    for (int g = 0; g < G; ++g) {
        const RolloutP& q = groups[g];
        CMRL_REQUIRE(q.num_layers == p.num_layers && q.H == p.H && q.act == p.act, "mlp_rollout_f16 v5: grouped mechanisms must share depth, width and activation");
        for (int l = 0; l <= p.num_layers; ++l)
            CMRL_REQUIRE(q.layer_k[l] == p.layer_k[l] && q.layer_n[l] == p.layer_n[l] && q.layer_off[l] == p.layer_off[l],
                         "mlp_rollout_f16 v5: grouped mechanisms must share the packed layer tables (layer %d differs)", l);
        CMRL_REQUIRE(q.E >= 1 && q.E <= 64, "mlp_rollout_f16 v5: E must be <= 64");
        gs.g[g] = q;
        max_items += (long long)q.P * ((q.N + V5_TILE - 1) / V5_TILE + q.E);
    }
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v5_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7424));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v5_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7424));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v5_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7424));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v5_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 7424));
        attr_once.done(attr_dev);
    }
    long long clusters = num_sms() / 2;
    if (clusters > (max_items + 1) / 2) clusters = (max_items + 1) / 2;
    if (clusters < 1) clusters = 1;
    const int grid = (int)clusters * 2;
    if (p.act == CMRL_ACT_SILU)
        mlp_rollout_f16_v5_kernel<CMRL_ACT_SILU><<<grid, V5_THREADS, lay.total, st>>>(gs, lay);
    else if (p.act == CMRL_ACT_SILU_HALF)
        mlp_rollout_f16_v5_kernel<CMRL_ACT_SILU_HALF><<<grid, V5_THREADS, lay.total, st>>>(gs, lay);
    else if (p.act == CMRL_ACT_RELU)
        mlp_rollout_f16_v5_kernel<CMRL_ACT_RELU><<<grid, V5_THREADS, lay.total, st>>>(gs, lay);
    else
        mlp_rollout_f16_v5_kernel<CMRL_ACT_IDENTITY><<<grid, V5_THREADS, lay.total, st>>>(gs, lay);
    CMRL_CHECK_LAUNCH("mlp_rollout_f16_v5");
    return 0;
}

int launch_rollout_v5(const RolloutP& p, cudaStream_t st) { return launch_rollout_v5_groups(&p, 1, st); }

}  // namespace tc
}  // namespace cmrl
