// Tensor-core tier, v2: weight-stationary CTA pairs.
//
// Two CTAs of a cluster form one tcgen05 `cta_group::2` unit working on 256 env rows (128 per CTA):
//   * each CTA keeps HALF of every layer's B operand (output features split across the pair) RESIDENT in shared
//     memory for the whole run of tiles of one net, so steady-state weight traffic is zero (a streamed design would
//     need ~9 KB/clk of L2 bandwidth at tensor peak, above the ~6.3 KB/clk the L2 delivers);
//   * accumulators are double-buffered in TMEM (2 x 208 + 16 head columns of 512), and the epilogue hands the next
//     layer's A operand to the MMA warp in 64-column groups, so layer l+1's MMAs run while layer l's epilogue is
//     still converting the remaining columns;
//   * layer-0 operands (obs | act | 1) are gathered one tile ahead by two loader warps.
//
// Warp roles per CTA (640 threads): 0-15 epilogue (TMEM lane quarter = warp % 4; the four warps of a quarter interleave
// the 16-column chunks so SFU / TMEM-load / shared-store latencies overlap), 16 MMA issuer (leader CTA) + TMEM alloc,
// 17 weight TMA producer, 18-19 layer-0 operand loaders.
#include <cooperative_groups.h>

#include "tc_common.cuh"

namespace cmrl {
namespace tc {

constexpr int V2_THREADS = 640;
constexpr int V2_EPI_WARPS = 16;    // 4 per SM sub-partition (set s = warp / 4 takes the chunks c = s mod 4)
constexpr int V2_WARP_MMA = 16, V2_WARP_W = 17, V2_WARP_X = 18;
constexpr int V2_TILE = 256;        // rows per CTA pair
constexpr int V2_GROUP_COLS = 64;   // epilogue -> MMA handoff granularity (columns of the next layer's K)
constexpr uint32_t V2_D1_COL = 208; // second accumulator buffer (requires HP <= 208 for now)
constexpr uint32_t V2_DH_COL = 416; // head accumulator

struct V2Smem {
    // byte offsets inside dynamic smem, identical in both CTAs (descriptors are shared by the pair)
    uint32_t w_off[8];   // per layer: this CTA's half of the B operand
    uint32_t a_off;      // A operand buffer (HP/8 chunks)
    uint32_t x_off[2];   // layer-0 operand double buffer
    uint32_t bar_off;    // mbarriers
    uint32_t total;
};

// barrier indices (uint64_t each)
enum {
    BAR_W_FULL = 0,   // local: weights landed (tx)
    BAR_W_READY = 1,  // leader: both CTAs' weights landed (count 2)
    BAR_W_EMPTY = 2,  // both: MMAs of the previous net retired (commit multicast)
    BAR_X_READY = 3,  // leader [2]: layer-0 operand written by both CTAs (count 4)
    BAR_X_EMPTY = 5,  // both [2]: layer-0 MMA retired
    BAR_A_READY = 7,  // leader [4]: column group g of the next A operand written by both CTAs (count 8)
    BAR_D_FULL = 11,  // both [2]: accumulator buffer complete
    BAR_DH_FULL = 13, // both: head accumulator complete
    BAR_COUNT = 14,
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster.  Default (CTA-scope release) on purpose:
// `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR per arrive (measured: 4x slower epilogue).  The data the
// arrive publishes (A-operand chunks) is written and consumed inside the same CTA (by its own tensor unit, after a
// fence.proxy.async); only the signal crosses the pair.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    unsigned long long spins = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (when all MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// activation on a packed fp16x2 pre-activation.  SiLU(z) = z*sigmoid(z) = 0.5z*(1 + tanh(0.5z)): one MUFU op
// (tanh.approx.f16x2) per TWO elements instead of ex2 + rcp per element -- the epilogue is MUFU/issue bound.
template <int ACT>
__device__ __forceinline__ uint32_t act_h2(uint32_t z) {
    if (ACT == CMRL_ACT_SILU) {
        uint32_t hz, t, y;
        asm("mul.f16x2 %0, %1, %2;" : "=r"(hz) : "r"(z), "r"(0x38003800u));  // 0.5 * z
        asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(hz));
        asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(y) : "r"(hz), "r"(t));      // hz * t + hz
        return y;
    } else if (ACT == CMRL_ACT_SILU_HALF) {
        uint32_t t, y;
        asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(z));
        asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(y) : "r"(z), "r"(t));
        return y;
    } else if (ACT == CMRL_ACT_RELU) {
        uint32_t y;
        asm("max.f16x2 %0, %1, %2;" : "=r"(y) : "r"(z), "r"(0u));
        return y;
    }
    return z;
}

// fp32 variant: MUFU.TANH (fp32) issues at the full SFU rate, MUFU.TANH.F16 measured at half of it on sm_100a
template <int ACT>
__device__ __forceinline__ uint32_t act_pack(float a, float b) {
    if (ACT == CMRL_ACT_SILU) {
        float ha = 0.5f * a, hb = 0.5f * b, ta, tb;
        asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(ha));
        asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(hb));
        return pack_h2(fmaf(ha, ta, ha), fmaf(hb, tb, hb));
    } else if (ACT == CMRL_ACT_SILU_HALF) {
        float ta, tb;
        asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
        asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
        return pack_h2(fmaf(a, ta, a), fmaf(b, tb, b));
    } else if (ACT == CMRL_ACT_RELU) {
        return pack_h2(fmaxf(a, 0.0f), fmaxf(b, 0.0f));
    }
    return pack_h2(a, b);
}

// item -> (p, e, first sorted row of the 256-row tile, rows in tile)
__device__ __forceinline__ bool v2_decode(const RolloutP& p, int item, int tiles_total, const int* s_off, int& pp, int& e,
                                          int& row0, int& nrows) {
    pp = item / tiles_total;
    int t = item - pp * tiles_total;
    for (e = 0; e < p.E; ++e) {
        int cnt = s_off[e + 1] - s_off[e];
        int nt = (cnt + V2_TILE - 1) / V2_TILE;
        if (t < nt) {
            row0 = s_off[e] + t * V2_TILE;
            nrows = min(V2_TILE, cnt - t * V2_TILE);
            return true;
        }
        t -= nt;
    }
    return false;
}

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(V2_THREADS, 1)
    mlp_rollout_f16_v2_kernel(const RolloutP p, const V2Smem sm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_off[65];
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    const int K0 = p.layer_k[0];
    const int NH = p.layer_n[L];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + sm.bar_off);
    uint8_t* a_buf = smem + sm.a_off;
    const uint32_t rank = cluster_ctarank();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_groups = (HP + V2_GROUP_COLS - 1) / V2_GROUP_COLS;

    for (int i = threadIdx.x; i <= p.E; i += blockDim.x) s_off[i] = p.offsets[i];
    if (threadIdx.x == 0) {
        mbar_init(&bars[BAR_W_FULL], 1);
        mbar_init(&bars[BAR_W_READY], 2);
        mbar_init(&bars[BAR_W_EMPTY], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[BAR_X_READY + i], 4);
            mbar_init(&bars[BAR_X_EMPTY + i], 1);
            mbar_init(&bars[BAR_D_FULL + i], 1);
        }
        for (int g = 0; g < 4; ++g) mbar_init(&bars[BAR_A_READY + g], 2 * V2_EPI_WARPS);
        mbar_init(&bars[BAR_DH_FULL], 1);
        fence_barrier_init();
    }
    if (warp == V2_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    int tiles_total = 0;
    for (int e = 0; e < p.E; ++e) tiles_total += (s_off[e + 1] - s_off[e] + V2_TILE - 1) / V2_TILE;
    const int n_items = p.P * tiles_total;
    const int n_clusters = gridDim.x / 2;
    const int cid = blockIdx.x / 2;
    const int item_begin = (int)((long long)cid * n_items / n_clusters);
    const int item_end = (int)((long long)(cid + 1) * n_items / n_clusters);

    if (warp == V2_WARP_W) {
        // ===================== weight producer: this CTA's half of every layer of the current net ==============
        if (lane == 0) {
            int cur_net = -1;
            uint32_t n_w = 0;
            for (int item = item_begin; item < item_end; ++item) {
                int pp, e, row0, nrows;
                v2_decode(p, item, tiles_total, s_off, pp, e, row0, nrows);
                const int net = pp * p.E + e;
                if (net == cur_net) continue;
                if (cur_net != -1) mbar_wait(&bars[BAR_W_EMPTY], (n_w - 1) & 1);
                const __half* blob = p.packed + (long long)net * p.blob_halfs;
                uint32_t total = 0;
                for (int l = 0; l <= L; ++l) total += (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l] * 2;
                mbar_expect_tx(&bars[BAR_W_FULL], total);
                for (int l = 0; l <= L; ++l) {
                    const uint32_t half_halfs = (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l];
                    tma_bulk_g2s(smem + sm.w_off[l], blob + p.layer_off[l] + (size_t)rank * half_halfs, half_halfs * 2,
                                 &bars[BAR_W_FULL]);
                }
                mbar_wait(&bars[BAR_W_FULL], n_w & 1);
                mbar_arrive_remote(&bars[BAR_W_READY], 0);
                ++n_w;
                cur_net = net;
            }
        }
    } else if (warp == V2_WARP_MMA) {
        // ===================== MMA issuer: one lane of the leader CTA ============================================
        if (rank == 0 && lane == 0) {
            int cur_net = -1;
            uint32_t n_w = 0, n_x = 0, n_epi = 0;
            long long t_wx = 0, t_wa = 0, t_issue = 0, t_all0 = clock64(), tt;
            const uint32_t a_addr = smem_u32(a_buf);
            const uint32_t idesc_h = umma_idesc_f16(256, HP);
            const uint32_t idesc_head = umma_idesc_f16(256, NH);
            for (int item = item_begin; item < item_end; ++item) {
                int pp, e, row0, nrows;
                v2_decode(p, item, tiles_total, s_off, pp, e, row0, nrows);
                const int net = pp * p.E + e;
                if (net != cur_net) {
                    if (cur_net != -1) umma_commit_2cta(&bars[BAR_W_EMPTY]);
                    mbar_wait_cluster(&bars[BAR_W_READY], n_w & 1);
                    ++n_w;
                    cur_net = net;
                }
                // ---- layer 0: A = gathered [obs|act|1] --------------------------------------------------------
                const int xb = n_x & 1;
                mbar_wait_cluster(&bars[BAR_X_READY + xb], (n_x >> 1) & 1);
                tc_fence_after();
                {
                    const uint32_t x_addr = smem_u32(smem + sm.x_off[xb]);
                    const uint32_t b_addr = smem_u32(smem + sm.w_off[0]);
                    const uint32_t b_sbo = (uint32_t)(K0 / 8) * 128;
                    const uint64_t ad0 = umma_desc(x_addr, A_CHUNK_BYTES, 128), bd0 = umma_desc(b_addr, 128, b_sbo);
                    for (int ks = 0; ks < K0 / 16; ++ks)
                        umma_f16_2cta(tmem_base, ad0 + (uint64_t)(ks * (2 * A_CHUNK_BYTES / 16)), bd0 + (uint64_t)(ks * (256 / 16)),
                                      idesc_h, ks > 0);
                    umma_commit_2cta(&bars[BAR_X_EMPTY + xb]);
                    umma_commit_2cta(&bars[BAR_D_FULL + 0]);
                }
                // ---- layers 1..L-1 and the head: A produced by the epilogue in column groups --------------------
                for (int l = 1; l <= L; ++l) {
                    const bool head = (l == L);
                    const uint32_t d_tmem = tmem_base + (head ? V2_DH_COL : ((l & 1) ? V2_D1_COL : 0u));
                    const uint32_t b_addr = smem_u32(smem + sm.w_off[l]);
                    const uint32_t b_sbo = (uint32_t)(HP / 8) * 128;
                    const uint32_t idesc = head ? idesc_head : idesc_h;
                    const uint64_t ad0 = umma_desc(a_addr, A_CHUNK_BYTES, 128), bd0 = umma_desc(b_addr, 128, b_sbo);
                    for (int g = 0; g < n_groups; ++g) {
                        if (p.dbg) tt = clock64();
                        mbar_wait_cluster(&bars[BAR_A_READY + g], n_epi & 1);
                        if (p.dbg) {
                            long long t = clock64();
                            t_wa += t - tt;
                            tt = t;
                        }
                        tc_fence_after();
                        const int ks0 = g * (V2_GROUP_COLS / 16);
                        const int ks1 = min(HP / 16, ks0 + V2_GROUP_COLS / 16);
                        for (int ks = ks0; ks < ks1; ++ks)  // start-address field advances in 16-byte units
                            umma_f16_2cta(d_tmem, ad0 + (uint64_t)(ks * (2 * A_CHUNK_BYTES / 16)), bd0 + (uint64_t)(ks * (256 / 16)),
                                          idesc, ks > 0);
                        if (p.dbg) t_issue += clock64() - tt;
                    }
                    umma_commit_2cta(head ? &bars[BAR_DH_FULL] : &bars[BAR_D_FULL + (l & 1)]);
                    ++n_epi;
                }
                ++n_x;
            }
            if (p.dbg) {
                long long* d = p.dbg + (long long)blockIdx.x * 16;
                d[0] = clock64() - t_all0;
                d[1] = t_wx;
                d[2] = t_wa;
                d[3] = t_issue;
                d[4] = item_end - item_begin;
            }
        }
    } else if (warp >= V2_WARP_X) {
        // ===================== layer-0 operand loaders (64 threads, 2 rows each), one tile ahead =================
        const int lt = threadIdx.x - V2_WARP_X * 32;
        uint32_t n_x = 0;
        const int in = p.O + p.A;
        for (int item = item_begin; item < item_end; ++item, ++n_x) {
            int pp, e, row0, nrows;
            v2_decode(p, item, tiles_total, s_off, pp, e, row0, nrows);
            const int xb = n_x & 1;
            mbar_wait(&bars[BAR_X_EMPTY + xb], ((n_x >> 1) & 1) ^ 1);
            uint8_t* xbuf = smem + sm.x_off[xb];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = lt + h * 64;
                const int grow = (int)rank * TILE_M + r;
                const bool valid = grow < nrows;
                const long long env = valid ? p.perm[row0 + grow] : 0;
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
                    *reinterpret_cast<uint4*>(xbuf + c * A_CHUNK_BYTES + r * 16) =
                        make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&bars[BAR_X_READY + xb], 0);
        }
    } else {
        // ===================== epilogue warps: rows = TMEM lanes ==================================================
        // 16 warps = 4 per SM sub-partition.  Warp (quarter q, set s) converts the 16-column chunks c = 4g + s of TMEM
        // lanes 32q..32q+31, so every 64-column handoff group g gets exactly one chunk from each set and is published
        // (to the MMA issuer of the leader CTA) as soon as the four sets are through it.
        const int quarter = warp & 3, cset = warp >> 2;
        const int r = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t n_full0 = 0, n_full1 = 0, n_h = 0;
        const int n_chunks = HP / 16;
        uint8_t* a_row = a_buf + r * 16;
        const bool timing = p.dbg != nullptr && threadIdx.x == 0;
        long long t_wd = 0, t_loop = 0, t_head = 0, t0 = timing ? clock64() : 0, tt = 0;
        for (int item = item_begin; item < item_end; ++item) {
            int pp, e, row0, nrows;
            v2_decode(p, item, tiles_total, s_off, pp, e, row0, nrows);
            const int grow = (int)rank * TILE_M + r;
            const bool valid = grow < nrows;
            const long long env = valid ? p.perm[row0 + grow] : 0;
            for (int l = 0; l < L; ++l) {
                const int b = l & 1;
                if (timing) tt = clock64();
                if (b) {
                    mbar_wait(&bars[BAR_D_FULL + 1], n_full1 & 1);
                    ++n_full1;
                } else {
                    mbar_wait(&bars[BAR_D_FULL + 0], n_full0 & 1);
                    ++n_full0;
                }
                if (timing) {
                    long long t = clock64();
                    t_wd += t - tt;
                    tt = t;
                }
                tc_fence_after();
                const uint32_t d_addr = tmem_base + lane_base + (b ? V2_D1_COL : 0u);
                uint32_t ra[16], rb[16];
                auto convert_store = [&](int c, const uint32_t(&raw)[16]) {
                    uint32_t h[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) h[j] = act_pack<ACT>(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
                    *reinterpret_cast<uint4*>(a_row + (2 * c) * A_CHUNK_BYTES) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(a_row + (2 * c + 1) * A_CHUNK_BYTES) = make_uint4(h[4], h[5], h[6], h[7]);
                    if (c == (p.H >> 4))  // the ones column (fp16 1.0) that carries the next layer's bias
                        *reinterpret_cast<uint16_t*>(a_row + (p.H >> 3) * A_CHUNK_BYTES + (p.H & 7) * 2) = 0x3C00;
                };
                auto publish = [&](int g) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&bars[BAR_A_READY + g], 0);
                };
                __syncwarp();
                int c = cset;
                if (c < n_chunks) tmem_ld16_issue(d_addr + c * 16, ra);
                for (int g = 0; g < n_groups; g += 2) {
                    if (c < n_chunks) {
                        tmem_ld_wait();
                        if (c + 4 < n_chunks) tmem_ld16_issue(d_addr + (c + 4) * 16, rb);
                        else tc_fence_before();
                        convert_store(c, ra);
                    }
                    publish(g);
                    c += 4;
                    if (g + 1 < n_groups) {
                        if (c < n_chunks) {
                            tmem_ld_wait();
                            if (c + 4 < n_chunks) tmem_ld16_issue(d_addr + (c + 4) * 16, ra);
                            else tc_fence_before();
                            convert_store(c, rb);
                        }
                        publish(g + 1);
                        c += 4;
                    }
                }
                if (timing) t_loop += clock64() - tt;
            }
            if (cset != 0) continue;  // the head epilogue (one row per lane) is done by the first warp set
            // ---- head -------------------------------------------------------------------------------------------
            if (timing) tt = clock64();
            const int D = p.head_dim;
            float res0 = 0.0f, nz0 = 0.0f;
            if (p.layout == CMRL_HEAD_PARALLEL && valid) {  // prefetch before the wait
                if (p.residual) res0 = p.obs[env * p.O + pp];
                if (p.gaussian && p.noise) nz0 = p.noise[env * D + pp];
            }
            mbar_wait(&bars[BAR_DH_FULL], n_h & 1);
            ++n_h;
            tc_fence_after();
            const uint32_t h_addr = tmem_base + lane_base + V2_DH_COL;
            if (p.layout == CMRL_HEAD_PARALLEL) {
                float v[16];
                tmem_ld16(h_addr, v);
                if (valid) {
                    float out = v[0] + res0;
                    if (p.gaussian && p.noise) {
                        const int bi = p.bounds_n == 1 ? 0 : pp;
                        const float lv = soft_clamp_logvar(v[1], p.min_logvar[bi], p.max_logvar[bi]);
                        out += sqrtf(expf(lv)) * nz0;
                    }
                    p.pred[env * D + pp] = out;
                }
            } else {
                for (int cc = 0; cc < NH / 16; ++cc) {  // interleaved head columns: 2d = mean_d, 2d+1 = raw logvar_d
                    float v[16];
                    tmem_ld16(h_addr + cc * 16, v);
                    if (p.gaussian) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const int d = (cc * 16 + j) >> 1;
                            if (valid && d < D) {
                                float m = v[j];
                                if (p.residual) m += p.obs[env * p.O + d];
                                if (p.noise) {
                                    const int bi = p.bounds_n == 1 ? 0 : d;
                                    const float lv = soft_clamp_logvar(v[j + 1], p.min_logvar[bi], p.max_logvar[bi]);
                                    m += sqrtf(expf(lv)) * p.noise[env * D + d];
                                }
                                p.pred[env * D + d] = m;
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int d = cc * 16 + j;
                            if (valid && d < D) {
                                float m = v[j];
                                if (p.residual) m += p.obs[env * p.O + d];
                                p.pred[env * D + d] = m;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            if (timing) t_head += clock64() - tt;
        }
        if (timing) {
            long long* d = p.dbg + (long long)blockIdx.x * 16;
            d[8] = clock64() - t0;
            d[9] = t_wd;
            d[10] = t_loop;
            d[12] = t_head;
        }
    }

    // ---- teardown: both CTAs must be done with TMEM / peer smem before either exits -------------------------
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == V2_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

int launch_rollout_v2(const RolloutP& p, cudaStream_t st) {
    const int L = p.num_layers;
    const int HP = p.layer_n[0];
    CMRL_REQUIRE(HP <= 208 && HP % 16 == 0, "mlp_rollout_f16 v2: padded hidden width must be <= 208 (hidden <= 207)");
    CMRL_REQUIRE(p.layer_n[L] + V2_DH_COL <= 512, "mlp_rollout_f16 v2: head wider than %d columns", 512 - (int)V2_DH_COL);
    V2Smem sm{};
    uint32_t off = 0;
    for (int l = 0; l <= L; ++l) {
        sm.w_off[l] = off;
        off += (uint32_t)(p.layer_n[l] / 2) * p.layer_k[l] * 2;
        off = (off + 127) / 128 * 128;
    }
    sm.a_off = off;
    off += (uint32_t)(HP / 8) * A_CHUNK_BYTES;
    for (int i = 0; i < 2; ++i) {
        sm.x_off[i] = off;
        off += (uint32_t)(p.layer_k[0] / 8) * A_CHUNK_BYTES;
    }
    sm.bar_off = off;
    off += BAR_COUNT * 8;
    sm.total = off;
    CMRL_REQUIRE(sm.total <= 232448 - 1024, "mlp_rollout_f16 v2: needs %u B of shared memory", sm.total);
    static bool attr_set = false;
    if (!attr_set) {
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(mlp_rollout_f16_v2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        attr_set = true;
    }
    int max_items = p.P * ((p.N + V2_TILE - 1) / V2_TILE + p.E);
    int clusters = num_sms() / 2;
    if (clusters > max_items) clusters = max_items;
    if (clusters < 1) clusters = 1;
    if (p.act == CMRL_ACT_SILU)
        mlp_rollout_f16_v2_kernel<CMRL_ACT_SILU><<<clusters * 2, V2_THREADS, sm.total, st>>>(p, sm);
    else if (p.act == CMRL_ACT_SILU_HALF)
        mlp_rollout_f16_v2_kernel<CMRL_ACT_SILU_HALF><<<clusters * 2, V2_THREADS, sm.total, st>>>(p, sm);
    else if (p.act == CMRL_ACT_RELU)
        mlp_rollout_f16_v2_kernel<CMRL_ACT_RELU><<<clusters * 2, V2_THREADS, sm.total, st>>>(p, sm);
    else
        mlp_rollout_f16_v2_kernel<CMRL_ACT_IDENTITY><<<clusters * 2, V2_THREADS, sm.total, st>>>(p, sm);
    CMRL_CHECK_LAUNCH("mlp_rollout_f16_v2");
    return 0;
}

}  // namespace tc
}  // namespace cmrl
