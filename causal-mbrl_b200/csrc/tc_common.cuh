// Inline-PTX helpers shared by the tcgen05 kernels (mbarrier, TMA bulk copy, UMMA descriptors, TMEM access).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace cmrl {
namespace tc {

constexpr int TILE_M = 128;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t HEAD_TMEM_COL = 256;
constexpr unsigned long long SPIN_LIMIT = 1ull << 27;  // ~ seconds; a broken pipeline traps instead of hanging the GPU

struct RolloutP {
    const __half* packed;      // [nets][blob]
    long long blob_halfs;      // per-net blob size (halves)
    int layer_off[8];          // offset (halves) of each layer's B blob inside a net blob
    int layer_k[8];            // padded K per layer (multiple of 16)
    int layer_n[8];            // padded N per layer (multiple of 16)
    int num_layers;            // hidden layers L; layer index L is the head
    int P, E, N, O, A, H;
    int head_dim;              // D: outputs per env this mech predicts (O, or 1)
    int layout;                // CMRL_HEAD_PLAIN / CMRL_HEAD_PARALLEL
    int gaussian;              // 0 = deterministic head
    int act;
    const float* obs;          // [N,O] env order
    const float* actn;         // [N,A]
    const int* perm;           // [N] sorted row -> env
    const int* offsets;        // [E+1]
    int residual;
    const float* min_logvar;
    const float* max_logvar;
    int bounds_n;
    const float* noise;        // [N,D] env order or NULL
    float* pred;               // [N,D] env order
    long long pred_se;         // floats between the outputs of consecutive MEMBERS (0: one [N,D] array; E*N all-member rows: N*D)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
// same, with a sleep between polls: for roles with slack (a spinning warp takes issue slots from the epilogue warps of
// its SM sub-partition)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns) {
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
        if (ns) __nanosleep(ns);
        if (++spins > SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE ("interleave"): core matrix = 8 rows x 16 bytes
// contiguous; LBO = byte step between the two K core matrices of one MMA, SBO = byte step between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// kind::f16 instruction descriptor: D fp32, A/B fp16, both K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    __syncwarp();
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float act_fast(float z, int act) {
    if (act == CMRL_ACT_RELU) return fmaxf(z, 0.0f);
    if (act == CMRL_ACT_SILU) return __fdividef(z, 1.0f + __expf(-z));
    if (act == CMRL_ACT_SILU_HALF) return __fdividef(2.0f * z, 1.0f + __expf(-2.0f * z));  // z holds half the pre-activation
    return z;
}

// two fp32 -> packed fp16x2 (a in the low half), round-to-nearest, saturating to +-65504 instead of inf
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}


// ---- helpers shared by the CTA-pair (cta_group::2) kernels -------------------------------------------------------
__device__ __forceinline__ uint32_t v3_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void v3_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void v3_arrive_leader(uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, 0;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void v3_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, descriptors passed as (lo, hi) words: only the low word (start-address field) changes between K steps, so the
// issue loop is two 32-bit adds per MMA (a single thread issues dependent instructions at ~1 per 4-6 cycles; a long
// per-MMA address computation makes the ISSUE, not the tensor pipe, the bottleneck -- measured 173 cycles per MMA)
__device__ __forceinline__ void v3_mma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "setp.ne.b32 p, %6, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-uniform variants: executed by ALL lanes of the (converged) MMA warp with warp-uniform operands; one elected
// lane issues.  Keeping the control flow uniform lets ptxas keep descriptors in uniform registers instead of emitting
// an ELECT / R2UR.BROADCAST waterfall loop around every UTCHMMA.
__device__ __forceinline__ void v3_mma_uni(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred pe, p;\n"
        ".reg .b64 da, db;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "setp.ne.b32 p, %6, 0;\n"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void v3_commit_uni(uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .pred pe;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void v3_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void v3_ld8_issue(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking phase test
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void v3_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void v3_sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void v3_sts16(uint32_t saddr, uint16_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(v) : "memory");
}

// ---- A operand in tensor memory (".ts" form): fp16 pairs packed in 32-bit columns, lane = row, 8 columns per K step ---
__device__ __forceinline__ void v5_mma_ts_uni(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred pe, p;\n"
        ".reg .b64 db;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void v5_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void v5_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void v5_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

template <int ACT>
__device__ __forceinline__ uint32_t v3_act_pack(float a, float b) {
    if (ACT == CMRL_ACT_SILU) {  // z*sigmoid(z) = 0.5z*(1 + tanh(0.5z)): one SFU op per element
        float ha = 0.5f * a, hb = 0.5f * b, ta, tb;
        asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(ha));
        asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(hb));
        return pack_h2(fmaf(ha, ta, ha), fmaf(hb, tb, hb));
    } else if (ACT == CMRL_ACT_SILU_HALF) {  // a, b are already z/2 (weights packed with scale 0.5): no multiply
        float ta, tb;
        asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
        asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
        return pack_h2(fmaf(a, ta, a), fmaf(b, tb, b));
    } else if (ACT == CMRL_ACT_RELU) {
        return pack_h2(fmaxf(a, 0.0f), fmaxf(b, 0.0f));
    }
    return pack_h2(a, b);
}


// ---- kind::tf32 (cta_group::1) helpers shared by the training kernels (gemm_tc.cu, train_tc.cu) -----------------------
// kind::tf32 instruction descriptor: D fp32, A/B tf32, both K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// warp-uniform forms (all lanes execute, one elected lane issues; see v3_mma_uni)
__device__ __forceinline__ void tf32_mma_ss_uni(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred pe, p;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory: lane = row, one 32-bit column per k (8 columns per K step) -- profiles/r2_tf32_probe.txt
__device__ __forceinline__ void tf32_mma_ts_uni(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred pe, p;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_uni(uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .pred pe;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr));
}


// cta_group::2 kind::tf32 (CTA pair, M = 256: 128 rows per CTA, the B operand split over the pair's shared memories)
__device__ __forceinline__ void tf32_mma2_ss_uni(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred pe, p;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, descriptors as (lo, hi) words: only the start-address field in the low word changes between MMAs, so the issue loop is
// one 32-bit add per operand (computing a 64-bit descriptor per MMA in the issuing thread cost ~170 cycles per MMA and paced the
// whole training kernel: profiles/r2_train_trace_*.txt)
__device__ __forceinline__ void tf32_mma2_lohi_uni(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                   uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred pe, p;\n"
        ".reg .b64 da, db;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "setp.ne.b32 p, %6, 0;\n"
        "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in tensor memory (each CTA of the pair supplies its own 128 rows), B descriptor as (lo, hi) words
__device__ __forceinline__ void tf32_mma2_ts_lohi_uni(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                      uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred pe, p;\n"
        ".reg .b64 db;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "@pe tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// multicast forms for 2-CTA clusters of independent (cta_group::1) tiles that share one weight stream
__device__ __forceinline__ void tma_bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc_uni(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "{\n"
        ".reg .pred pe;\n"
        "elect.sync _|pe, 0xffffffff;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}
// activation / derivative with the fast exp and divide intrinsics (a few ulp: far inside the 1e-5 tier)
__device__ __forceinline__ float act_fwd_fast(float z, int act) {
    if (act == CMRL_ACT_RELU) return fmaxf(z, 0.0f);
    if (act == CMRL_ACT_SILU) return __fdividef(z, 1.0f + __expf(-z));
    return z;
}
__device__ __forceinline__ float act_grad_fast(float z, int act) {
    if (act == CMRL_ACT_RELU) return z > 0.0f ? 1.0f : 0.0f;
    if (act == CMRL_ACT_SILU) {
        const float s = __fdividef(1.0f, 1.0f + __expf(-z));
        return s * (1.0f + z * (1.0f - s));
    }
    return 1.0f;
}

}  // namespace tc
}  // namespace cmrl
