// K1, tensor-core tier for TRAINING (dense, non-ragged): grouped per-net GEMMs on tcgen05 kind::tf32.
//
//   fwd : C[b,o] = act(sum_i (x*mask)[b,i] W[i,o] + bias[o])
//   dx  : C[b,i] = (sum_o dz[b,o] W[i,o]) * act'(z_prev[b,i])
//   dw  : C[i,o] = sum_b (x*mask)[b,i] dz[b,o];  db[o] = sum_b dz[b,o]
//
// The operands are plain row-major fp32 tensors with arbitrary strides (GemmP, gemm_common.cuh), so they are not
// TMA-shaped: all 512 threads gather 16-byte K groups from global memory (vector loads where the K direction is
// contiguous, otherwise four coalesced scalar loads), split every value into tf32 hi + lo parts in registers
// ("3xTF32": a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32-grade; passes = 1 keeps hi only) and store them in the
// K-major no-swizzle core-matrix layout the MMA reads (8 rows x 16 B; K groups padded by 16 B against bank conflicts).
// One thread issues the MMAs of a 32-wide K chunk (M=128 rows x N<=256 columns, fp32 accumulators in TMEM) while
// everybody already loads the next chunk (two shared-memory stages, tcgen05.commit -> mbarrier frees a stage).  db comes
// out of the same MMAs: the A operand of the dw GEMM carries a row of ones at index `in`.
//
// At the reference batch size (256 rows per member) the GEMMs are latency bound (14 CTAs per launch); the tensor pipe
// is what makes one CTA's 128 x 208 x 200 tile take ~3 us instead of the ~18 us of the FFMA kernel.
#include "gemm_common.cuh"
#include "tc_common.cuh"

namespace cmrl {

using namespace tc;

constexpr int TG_THREADS = 512;
constexpr int TG_BM = 128;
constexpr int TG_KC = 32;              // fp32 elements per K chunk = 4 MMA K steps of 8
constexpr int TG_GROUPS = TG_KC / 4;   // 16-byte K groups per row per chunk
constexpr int TG_MAX_NT = 256;

struct TgLay {
    int NT;               // N tile (multiple of 16, <= 256)
    int parts;            // 1: hi only (TF32), 2: hi + lo (3xTF32)
    uint32_t a_gstride;   // bytes between K groups of the A tile (128 rows x 16 B + 16 B pad)
    uint32_t b_gstride;
    uint32_t a_part, b_part;  // bytes of one hi (or lo) tile
    uint32_t stage;           // bytes of one stage
    uint32_t bar_off, total;
    uint32_t tmem_cols;
    int ksplit;           // dw only: the K (= batch) range is cut into ksplit slices, one CTA each, partial tiles go to `ws`
    int k_per_split;      // multiple of TG_KC
    float* ws;            // [ksplit][G][Mtot][N] partial results (Mtot = M + 1 with the db row)
    int m_tot;
};

// activation / derivative with the fast exp and divide intrinsics (a few ulp: far inside the 1e-5 tier; the precise
// expf + IEEE divide of common.cuh cost ~30 instructions per element and a third of the tile time in this kernel)
__device__ __forceinline__ float tg_act_fwd(float z, int act) {
    if (act == CMRL_ACT_RELU) return fmaxf(z, 0.0f);
    if (act == CMRL_ACT_SILU) return __fdividef(z, 1.0f + __expf(-z));
    return z;
}
__device__ __forceinline__ float tg_act_grad(float z, int act) {
    if (act == CMRL_ACT_RELU) return z > 0.0f ? 1.0f : 0.0f;
    if (act == CMRL_ACT_SILU) {
        const float s = __fdividef(1.0f, 1.0f + __expf(-z));
        return s * (1.0f + z * (1.0f - s));
    }
    return 1.0f;
}

template <int MODE>
__global__ void __launch_bounds__(TG_THREADS, 1) gemm_tc_kernel(const GemmP p, const TgLay lay) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);  // [0],[1] stage free; [2] accumulator complete
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = lay.NT;

    const int pe_p = blockIdx.z / p.E, pe_e = blockIdx.z % p.E;
    const int split = lay.ksplit > 1 ? (int)blockIdx.x % lay.ksplit : 0;
    const int m_tile = lay.ksplit > 1 ? (int)blockIdx.x / lay.ksplit : (int)blockIdx.x;
    const int kbase = split * lay.k_per_split;
    const int Kloc = lay.ksplit > 1 ? min(p.K - kbase, lay.k_per_split) : p.K;  // this CTA's slice of the K range
    const int m0 = m_tile * TG_BM, n0 = blockIdx.y * NT;
    const float* A = p.A + pe_p * p.a_sp + pe_e * p.a_se + (long long)kbase * p.a_sk;
    const float* Am = p.Amask ? p.Amask + pe_p * p.am_sp + pe_e * p.am_se + (long long)kbase * p.am_sk : nullptr;
    const float* Bp = p.B + pe_p * p.b_sp + pe_e * p.b_se + (long long)kbase * p.b_sk;
    const long long c_off = pe_p * p.c_sp + pe_e * p.c_se;
    const long long g = (long long)pe_p * p.E + pe_e;
    const bool ones_row = (MODE == MODE_DW) && p.bsum != nullptr;  // A row index M (= in) is all ones -> C row M = db

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(lay.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // ---- loader: a work item is one 16-byte K group (4 consecutive k) of one tile row.  Everything that does not change
    // from chunk to chunk (row validity, global pointer, shared-memory offset) is computed once: the loop body is a
    // handful of instructions per item (the first version recomputed the index maths per item and chunk and was
    // instruction bound at ~10 us per chunk).
    const bool a_kfast = (p.a_sk == 1);
    const bool b_kfast = (p.b_sk == 1);
    const bool a_vec = a_kfast && Am == nullptr && (p.a_sm % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool b_vec = b_kfast && (p.b_sn % 4 == 0) && ((reinterpret_cast<uintptr_t>(Bp) & 15) == 0);
    constexpr int A_ITEMS = TG_BM * TG_GROUPS / TG_THREADS;      // 2
    constexpr int B_ITEMS = TG_MAX_NT * TG_GROUPS / TG_THREADS;  // 4 (items beyond NT rows are inactive)
    const float* a_ptr[A_ITEMS];
    const float* am_ptr[A_ITEMS];
    const float* b_ptr[B_ITEMS];
    uint32_t a_soff[A_ITEMS], b_soff[B_ITEMS];
    int a_k[A_ITEMS], b_k[B_ITEMS];          // k offset of the item inside a chunk
    int a_kind[A_ITEMS];                     // 0 = zero row, 1 = data row, 2 = ones row (db)
    bool b_on[B_ITEMS], b_data[B_ITEMS];
#pragma unroll
    for (int j = 0; j < A_ITEMS; ++j) {
        const int it = tid + j * TG_THREADS;
        const int c = a_kfast ? it % TG_GROUPS : it / TG_BM;
        const int m = a_kfast ? it / TG_GROUPS : it % TG_BM;
        const int gm = m0 + m;
        a_k[j] = c * 4;
        a_kind[j] = gm < p.M ? 1 : ((ones_row && gm == p.M) ? 2 : 0);
        a_ptr[j] = A + (long long)(gm < p.M ? gm : 0) * p.a_sm + (long long)(c * 4) * p.a_sk;
        am_ptr[j] = Am ? Am + (long long)(gm < p.M ? gm : 0) * p.am_sm + (long long)(c * 4) * p.am_sk : nullptr;
        a_soff[j] = (uint32_t)c * lay.a_gstride + (uint32_t)m * 16;
    }
#pragma unroll
    for (int j = 0; j < B_ITEMS; ++j) {
        const int it = tid + j * TG_THREADS;
        const int c = b_kfast ? it % TG_GROUPS : it / NT;
        const int n = b_kfast ? it / TG_GROUPS : it % NT;
        const int gn = n0 + n;
        b_on[j] = it < NT * TG_GROUPS;
        b_data[j] = b_on[j] && gn < p.N;
        b_k[j] = c * 4;
        b_ptr[j] = Bp + (long long)(gn < p.N ? gn : 0) * p.b_sn + (long long)(c * 4) * p.b_sk;
        b_soff[j] = (uint32_t)c * lay.b_gstride + (uint32_t)n * 16;
    }
    const long long a_step = (long long)TG_KC * p.a_sk, am_step = (long long)TG_KC * p.am_sk, b_step = (long long)TG_KC * p.b_sk;
    float4 ra0[A_ITEMS], rb0[B_ITEMS], ra1[A_ITEMS], rb1[B_ITEMS];  // two chunks in flight (register double buffer)

    // full: the whole chunk lies inside K (no per-element bound checks)
    auto load_chunk = [&](int k0, bool full, float4(&ra)[A_ITEMS], float4(&rb)[B_ITEMS]) {
#pragma unroll
        for (int j = 0; j < A_ITEMS; ++j) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_kind[j] == 1) {
                if (a_vec && full) {
                    v = *reinterpret_cast<const float4*>(a_ptr[j]);
                } else {
                    float e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float x = 0.f;
                        if (full || k0 + a_k[j] + u < Kloc) {
                            x = a_ptr[j][(long long)u * p.a_sk];
                            if (Am) x *= am_ptr[j][(long long)u * p.am_sk];
                        }
                        e[u] = x;
                    }
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
            } else if (a_kind[j] == 2) {
                const int kk = k0 + a_k[j];
                v = make_float4(kk < Kloc ? 1.f : 0.f, kk + 1 < Kloc ? 1.f : 0.f, kk + 2 < Kloc ? 1.f : 0.f, kk + 3 < Kloc ? 1.f : 0.f);
            }
            ra[j] = v;
            a_ptr[j] += a_step;
            if (Am) am_ptr[j] += am_step;
        }
#pragma unroll
        for (int j = 0; j < B_ITEMS; ++j) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b_data[j]) {
                if (b_vec && full) {
                    v = *reinterpret_cast<const float4*>(b_ptr[j]);
                } else {
                    float e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = (full || k0 + b_k[j] + u < Kloc) ? b_ptr[j][(long long)u * p.b_sk] : 0.f;
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
            rb[j] = v;
            b_ptr[j] += b_step;
        }
    };
    auto split_store = [&](uint32_t hi_addr, uint32_t lo_addr, const float4& v) {
        const uint32_t b0 = __float_as_uint(v.x), b1 = __float_as_uint(v.y), b2 = __float_as_uint(v.z), b3 = __float_as_uint(v.w);
        const uint32_t h0 = b0 & 0xFFFFE000u, h1 = b1 & 0xFFFFE000u, h2 = b2 & 0xFFFFE000u, h3 = b3 & 0xFFFFE000u;  // tf32 = top 19 bits
        v3_sts128(hi_addr, h0, h1, h2, h3);
        if (lay.parts == 2) {
            v3_sts128(lo_addr, __float_as_uint(v.x - __uint_as_float(h0)), __float_as_uint(v.y - __uint_as_float(h1)),
                      __float_as_uint(v.z - __uint_as_float(h2)), __float_as_uint(v.w - __uint_as_float(h3)));
        }
    };
    const uint32_t smem_base = smem_u32(smem);
    auto store_chunk = [&](int s, const float4(&ra)[A_ITEMS], const float4(&rb)[B_ITEMS]) {
        const uint32_t base = smem_base + (uint32_t)s * lay.stage;
        const uint32_t a_hi = base, a_lo = base + lay.a_part;
        const uint32_t b_hi = base + (uint32_t)lay.parts * lay.a_part, b_lo = b_hi + lay.b_part;
#pragma unroll
        for (int j = 0; j < A_ITEMS; ++j) split_store(a_hi + a_soff[j], a_lo + a_soff[j], ra[j]);
#pragma unroll
        for (int j = 0; j < B_ITEMS; ++j)
            if (b_on[j]) split_store(b_hi + b_soff[j], b_lo + b_soff[j], rb[j]);
    };

    // ---- main loop: chunk kc is stored from one register set while the loads of chunk kc+1 (other set) are in flight;
    // two shared-memory stages, the MMAs of chunk kc run while chunk kc+1 is stored
    const int nk = (Kloc + TG_KC - 1) / TG_KC;
    const uint32_t idesc = umma_idesc_tf32(TG_BM, NT);
    load_chunk(0, TG_KC <= Kloc, ra0, rb0);
    if (nk > 1) load_chunk(TG_KC, 2 * TG_KC <= Kloc, ra1, rb1);
    for (int kc = 0; kc < nk; ++kc) {
        const int s = kc & 1;
        if (kc >= 2) mbar_wait(&bars[s], ((kc >> 1) - 1) & 1);  // the MMAs that read this stage two chunks ago are done
        if (s == 0) store_chunk(0, ra0, rb0); else store_chunk(1, ra1, rb1);
        fence_proxy_async();
        __syncthreads();
        if (kc + 2 < nk) {
            if (s == 0) load_chunk((kc + 2) * TG_KC, (kc + 3) * TG_KC <= Kloc, ra0, rb0);
            else load_chunk((kc + 2) * TG_KC, (kc + 3) * TG_KC <= Kloc, ra1, rb1);
        }
        if (tid == 0) {
            tc_fence_after();
            const uint32_t base = smem_base + (uint32_t)s * lay.stage;
            const uint32_t a_hi = base, a_lo = base + lay.a_part;
            const uint32_t b_hi = base + (uint32_t)lay.parts * lay.a_part, b_lo = b_hi + lay.b_part;
            const int ksteps = min(TG_KC / 8, (Kloc - kc * TG_KC + 7) / 8);
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t ao = (uint32_t)(2 * ks) * lay.a_gstride, bo = (uint32_t)(2 * ks) * lay.b_gstride;
                const uint64_t dah = umma_desc(a_hi + ao, lay.a_gstride, 128), dbh = umma_desc(b_hi + bo, lay.b_gstride, 128);
                const uint32_t first = (kc == 0 && ks == 0) ? 0u : 1u;
                if (lay.parts == 2) {  // small terms first
                    const uint64_t dal = umma_desc(a_lo + ao, lay.a_gstride, 128), dbl = umma_desc(b_lo + bo, lay.b_gstride, 128);
                    umma_tf32(tmem, dal, dbh, idesc, first);
                    umma_tf32(tmem, dah, dbl, idesc, 1u);
                    umma_tf32(tmem, dah, dbh, idesc, 1u);
                } else {
                    umma_tf32(tmem, dah, dbh, idesc, first);
                }
            }
            umma_commit(&bars[s]);
            if (kc + 1 == nk) umma_commit(&bars[2]);
        }
    }
    mbar_wait(&bars[2], 0);
    tc_fence_after();

    // ---- epilogue: TMEM -> registers -> shared memory (the operand stages are free now) -> coalesced global stores.
    // A thread owns one accumulator ROW in TMEM; writing its 16 consecutive columns straight to global memory makes every
    // store instruction touch 32 different rows (32 sectors, serialised in the LSU: measured ~30 us per tile), so the
    // tile is transposed through shared memory and written row by row with the lanes along the columns.
    {
        float* otile = reinterpret_cast<float*>(smem);
        const int ostride = NT + 4;  // floats; 16-byte aligned rows, conflict-free for the float4 writes below
        const int q = warp & 3, part = warp >> 2;
        const int r = q * 32 + lane;
        const int n_chunks = NT / 16;
        for (int c = part; c < n_chunks; c += TG_THREADS / 128) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c * 16, v);
            float4* dst = reinterpret_cast<float4*>(otile + (size_t)r * ostride + c * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        __syncthreads();
        // four columns per lane where the output rows allow 16-byte accesses (the scalar write-out -- one element, its 64-bit
        // address arithmetic, bias load and two stores per iteration -- was 45 % of the kernel's samples)
        const bool vec_ok = lay.ksplit == 1 && (p.N % 4 == 0) && (p.c_sm % 4 == 0) && (c_off % 4 == 0) &&
                            ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (!p.C2 || (reinterpret_cast<uintptr_t>(p.C2) & 15) == 0) &&
                            (!p.emul || (reinterpret_cast<uintptr_t>(p.emul) & 15) == 0) &&
                            (!p.bias || (reinterpret_cast<uintptr_t>(p.bias + g * p.N) & 15) == 0);
        for (int row = warp; row < TG_BM; row += TG_THREADS / 32) {
            const int gm = m0 + row;
            const bool row_ok = gm < p.M;
            const bool row_db = ones_row && gm == p.M;
            if (!row_ok && !row_db) continue;
            if (vec_ok && row_ok) {
                for (int col = lane * 4; col < NT; col += 128) {
                    const int gn = n0 + col;
                    if (gn >= p.N) continue;  // N % 4 == 0: a group of four is entirely inside or outside
                    float4 x = *reinterpret_cast<const float4*>(&otile[(size_t)row * ostride + col]);
                    const long long off = c_off + (long long)gm * p.c_sm + gn;
                    if (MODE == MODE_FWD) {
                        if (p.bias) {
                            const float4 bb = *reinterpret_cast<const float4*>(p.bias + g * p.N + gn);
                            x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                        }
                        if (p.C2) *reinterpret_cast<float4*>(p.C2 + off) = x;
                        *reinterpret_cast<float4*>(p.C + off) =
                            make_float4(tg_act_fwd(x.x, p.act), tg_act_fwd(x.y, p.act), tg_act_fwd(x.z, p.act), tg_act_fwd(x.w, p.act));
                    } else if (MODE == MODE_DX) {
                        if (p.emul) {
                            const float4 zz = *reinterpret_cast<const float4*>(p.emul + off);
                            x.x *= tg_act_grad(zz.x, p.act); x.y *= tg_act_grad(zz.y, p.act);
                            x.z *= tg_act_grad(zz.z, p.act); x.w *= tg_act_grad(zz.w, p.act);
                        }
                        *reinterpret_cast<float4*>(p.C + off) = x;
                    } else {
                        if (p.accumulate) {
                            const float4 cc = *reinterpret_cast<const float4*>(p.C + off);
                            x.x += cc.x; x.y += cc.y; x.z += cc.z; x.w += cc.w;
                        }
                        *reinterpret_cast<float4*>(p.C + off) = x;
                    }
                }
                continue;
            }
            for (int col = lane; col < NT; col += 32) {
                const int gn = n0 + col;
                if (gn >= p.N) continue;
                float x = otile[(size_t)row * ostride + col];
                if (lay.ksplit > 1) {  // partial tile of this K slice; splitk_reduce_kernel applies accumulate / db
                    lay.ws[(((long long)split * gridDim.z + blockIdx.z) * lay.m_tot + gm) * p.N + gn] = x;
                    continue;
                }
                if (row_ok) {
                    const long long off = c_off + (long long)gm * p.c_sm + gn;
                    if (MODE == MODE_FWD) {
                        if (p.bias) x += p.bias[g * p.N + gn];
                        if (p.C2) p.C2[off] = x;
                        p.C[off] = tg_act_fwd(x, p.act);
                    } else if (MODE == MODE_DX) {
                        if (p.emul) x *= tg_act_grad(p.emul[off], p.act);
                        p.C[off] = x;
                    } else {
                        p.C[off] = p.accumulate ? p.C[off] + x : x;
                    }
                } else {
                    const long long off = g * p.N + gn;
                    p.bsum[off] = p.accumulate ? p.bsum[off] + x : x;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(lay.tmem_cols) : "memory");
}

// sums the K-slice partials in slice order (deterministic) into dw (rows < M) and db (row M)
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int S, long long G, int m_tot, int M, int N, float* __restrict__ C,
                                     float* __restrict__ bsum, int accumulate) {
    const long long total = G * m_tot * N, slice = total;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += ws[s * slice + i];
        const int n = (int)(i % N);
        const int r = (int)((i / N) % m_tot);
        const long long g = i / ((long long)N * m_tot);
        float* dst = r < M ? &C[(g * M + r) * N + n] : &bsum[g * N + n];
        *dst = accumulate ? *dst + acc : acc;
    }
}

template <int MODE>
int launch_gemm_tc(const GemmP& p, int passes, cudaStream_t st, const char* name, float* workspace, long long workspace_floats) {
    if (p.row_offsets != nullptr) return -2;
    if (MODE == MODE_DW && p.bsum != nullptr && p.M % TG_BM == 0) return -2;  // no spare tile row for the ones row (db)
    if (p.M < 1 || p.N < 1 || p.K < 1) return -2;
    TgLay lay{};
    int NT = (p.N + 15) / 16 * 16;
    if (NT > TG_MAX_NT) NT = TG_MAX_NT;
    {   // few row tiles (the reference batch size: 256 rows per member -> 14 CTAs): split N so that the per-element
        // epilogue (bias, activation with expf, stores) and the operand loads spread over the whole chip
        const int m_rows0 = p.M + ((MODE == MODE_DW && p.bsum) ? 1 : 0);
        const long long ctas = (long long)((m_rows0 + TG_BM - 1) / TG_BM) * p.P * p.E;
        const int want = (int)((num_sms() + ctas - 1) / ctas);           // N tiles for ~one CTA per SM
        const bool will_split_k = MODE == MODE_DW && workspace != nullptr && p.K >= 2048;  // parallelism comes from K slices instead
        if (want > 1 && !will_split_k) {
            int nt = ((p.N + want - 1) / want + 15) / 16 * 16;
            if (nt < 32) nt = 32;
            if (nt < NT) NT = nt;
        }
    }
    lay.NT = NT;
    lay.parts = passes == 3 ? 2 : 1;
    lay.a_gstride = TG_BM * 16 + 16;
    lay.b_gstride = (uint32_t)NT * 16 + 16;
    lay.a_part = TG_GROUPS * lay.a_gstride;
    lay.b_part = TG_GROUPS * lay.b_gstride;
    lay.stage = (uint32_t)lay.parts * (lay.a_part + lay.b_part);
    lay.stage = (lay.stage + 127) / 128 * 128;
    lay.bar_off = 2 * lay.stage;
    lay.total = lay.bar_off + 128;
    const uint32_t otile = (uint32_t)TG_BM * (uint32_t)(NT + 4) * 4;  // epilogue staging tile re-uses the operand stages
    if (lay.total < otile) lay.total = otile;
    lay.tmem_cols = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256;
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        CMRL_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE_DW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 1024));
        attr_once.done(attr_dev);
    }
    CMRL_REQUIRE(lay.total <= 232448 - 1024, "%s: tensor-core tile needs %u B of shared memory", name, lay.total);
    const int m_rows = p.M + ((MODE == MODE_DW && p.bsum) ? 1 : 0);
    dim3 grid((m_rows + TG_BM - 1) / TG_BM, (p.N + NT - 1) / NT, p.P * p.E);
    CMRL_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "%s: too many nets/columns for one launch", name);
    lay.ksplit = 1;
    if (MODE == MODE_DW && workspace != nullptr && p.K >= 2048) {
        // dw: K is the batch.  Cut it into slices so that the launch fills the chip; partial tiles go to the caller's
        // workspace and are summed in slice order (contiguous dw [G,M,N] / db [G,N] as the Python layer passes them)
        const long long ctas = (long long)grid.x * grid.y * grid.z;
        int S = (int)((2LL * num_sms() + ctas - 1) / ctas);
        const long long per_slice = (long long)grid.z * m_rows * p.N;
        if (S > workspace_floats / per_slice) S = (int)(workspace_floats / per_slice);
        if (S > 32) S = 32;
        const bool contiguous = p.c_sm == p.N && p.c_se == (long long)p.M * p.N && (p.P == 1 || p.c_sp == (long long)p.E * p.M * p.N);
        if (S >= 2 && contiguous) {
            int kps = ((p.K + S - 1) / S + TG_KC - 1) / TG_KC * TG_KC;
            S = (p.K + kps - 1) / kps;
            lay.ksplit = S;
            lay.k_per_split = kps;
            lay.ws = workspace;
            lay.m_tot = m_rows;
            grid.x *= S;
        }
    }
    gemm_tc_kernel<MODE><<<grid, TG_THREADS, lay.total, st>>>(p, lay);
    CMRL_CHECK_LAUNCH(name);
    if (lay.ksplit > 1) {
        const long long total = (long long)grid.z * m_rows * p.N;
        int nb = (int)((total + 255) / 256);
        if (nb > 8 * num_sms()) nb = 8 * num_sms();
        splitk_reduce_kernel<<<nb, 256, 0, st>>>(workspace, lay.ksplit, grid.z, m_rows, p.M, p.N, p.C, p.bsum, p.accumulate);
        CMRL_CHECK_LAUNCH("ens_linear_bwd_dw.splitk_reduce");
    }
    return 0;
}

template int launch_gemm_tc<MODE_FWD>(const GemmP&, int, cudaStream_t, const char*, float*, long long);
template int launch_gemm_tc<MODE_DX>(const GemmP&, int, cudaStream_t, const char*, float*, long long);
template int launch_gemm_tc<MODE_DW>(const GemmP&, int, cudaStream_t, const char*, float*, long long);

}  // namespace cmrl
