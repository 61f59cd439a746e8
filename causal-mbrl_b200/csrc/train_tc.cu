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
constexpr int TF_THREADS = 20 * 32;
constexpr int TF_MAX_PH = 16;
constexpr int TF_MAX_LAYERS = 8;  // hidden layers + head
constexpr uint32_t TF_GROUP = TF_ROWS * 16;  // bytes of one 4-wide K group of a 128-row operand tile
constexpr uint32_t SMEM_LIMIT = 232448 - 1024;

enum { PH_FWD = 0, PH_HEAD = 1, PH_DX = 2 };

// Event trace of CTA (0,0) for profiles/tools/train_trace.py: compiled in only with -DCMRL_TRACE (a separate tool build of
// this file; the shipped library carries none of it).  Stamps are clock64 values in [role][1024] slots.
#ifdef CMRL_TRACE
__device__ long long* g_trace_buf = nullptr;
// tool build only -- experiments that INVALIDATE the results, to attribute time: 1 no weight copies, 2 no MMAs, 4 no global
// loads / stores in the epilogue, 8 no A-ring writes
__device__ int g_trace_flags = 0;
#define XFLAG(bit) (g_trace_flags & (bit))
#define TRACE_DECL int trace_n = 0
#define TRACE_FINE(on)                                                                                              \
    do {                                                                                                              \
        if ((on) && g_trace_buf != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && fine_n < 1024) \
            g_trace_buf[3 * 1024 + fine_n++] = clock64();                                                             \
    } while (0)
#define TRACE(role)                                                                                                   \
    do {                                                                                                              \
        if (g_trace_buf != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && trace_n < 1024) \
            g_trace_buf[(role) * 1024 + trace_n++] = clock64();                                                       \
    } while (0)
// per-chunk stamps of the MMA issuer (row 4), phase 1 of the first work item
#define TRACE_MMA(on)                                                                                                 \
    do {                                                                                                              \
        if ((on) && g_trace_buf != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0 && mma_n < 1024) \
            g_trace_buf[4 * 1024 + mma_n++] = clock64();                                                              \
    } while (0)
#else
#define XFLAG(bit) 0
#define TRACE_DECL
#define TRACE(role)
#define TRACE_FINE(on)
#define TRACE_MMA(on)
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
    // saved tiles per (net, row tile), floats: x | z_l (l < L) | dz_l (l <= L)
    long long off_x, off_z[TF_MAX_LAYERS], off_dz[TF_MAX_LAYERS + 1], item_floats;
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
    for (int l = 0; l < L; ++l) { g.off_z[l] = o; o += (long long)g.HP * TF_ROWS; }
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
        // chunk = [CTA half of the output rows][hi | lo][4 K groups][N/2 rows][4 floats]: each CTA of a pair streams its half
        const int half = n >= N / 2 ? 1 : 0, nl = n - half * (N / 2);
        float* dst = blob + (long long)c * 32 * N + (long long)half * 16 * N + ((long long)gi * (N / 2) + nl) * 4;
        uint32_t h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __float_as_uint(v[j]) & 0xFFFFE000u;
        *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(h[0]), __uint_as_float(h[1]), __uint_as_float(h[2]), __uint_as_float(h[3]));
        *reinterpret_cast<float4*>(dst + 8 * N) = make_float4(v[0] - __uint_as_float(h[0]), v[1] - __uint_as_float(h[1]),
                                                               v[2] - __uint_as_float(h[2]), v[3] - __uint_as_float(h[3]));
    }
}

// ================================================= fused forward + dz chain ================================================
// Shared memory: [A ring: R slots of one 16-deep K chunk x 128 rows, hi | lo][weight ring: S stages][barriers].
// Tensor memory: two accumulators (phases alternate, so the epilogue of phase p reads one while the MMAs of phase p+1 --
// which chase that epilogue chunk by chunk through the A ring -- fill the other) + the head accumulator.
constexpr int TF_A_SLOTS = 5;
constexpr uint32_t TF_A_PART = 4 * TF_GROUP;   // bytes of the hi (or lo) part of an A chunk: 4 K groups x 128 rows x 16 B
constexpr uint32_t TF_A_SLOT = 2 * TF_A_PART;
// SHARED-MEMORY BANDWIDTH bounds this kernel, not the tensor pipe (profiles/r2_train_trace_*.txt): a kind::tf32 MMA of
// 128 x 208 x 8 reads 4 KB of A and 6.6 KB of B in its 104 cycles (102 B/clk of the ~128 an SM has), three of them per K step,
// next to the weight stream landing and the epilogue writing the next A.  CTAs therefore run as cta_group::2 PAIRS on two row
// tiles of the same net (M = 256): each CTA holds and streams only HALF of every weight chunk (the B operand is split over the
// pair), which halves both the B reads and the weight-stream writes per SM.
constexpr int TF_WARP_FWD = 18;                // reports landed weight stages to the leader CTA
constexpr int TF_WARP_AW = 19;                 // leader: waits for A chunks on behalf of the MMA issuer
// An mbarrier wait costs the waiting thread ~300 cycles even when the phase completed long ago.  Two of them per 16-deep K chunk
// in the issuing thread (A chunk published, weight stage landed) paced the whole kernel at ~1.2k cycles per chunk against 624
// cycles of MMAs (3xTF32) or 208 (TF32).  Two helper warps therefore do the mbarrier waits AHEAD of the issuer and publish plain
// progress counters in shared memory; the issuer polls those with ordinary loads (~30 cycles).

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
    int tiles_per_net;   // real row tiles per net
    int pairs_per_net;   // cluster work items per net; tile slots in the scratch = 2 * pairs_per_net
    int stages;
    int lo_tmem;  // the lo parts of the A chunks live in tensor memory (columns left over by the accumulators) instead of shared memory
    uint32_t stage_bytes, w_off, bar_off;
    // inference (rollout) mode: rows of member e are the envs perm[row_off[e] .. row_off[e+1]) (cmrl_member_bucket); only the
    // forward phases run, nothing is saved, the head writes pred[env, :] = mean (+ obs) + sqrt(exp(logvar)) * noise[env, :]
    const int *perm, *row_off;
    const float* noise;
    float* pred;
    long long pred_se;  // floats between the outputs of consecutive members (0: one [N,D] array)
    int D;
};
constexpr int TF_MAX_E = 64;
// pair (cluster work item) -> (sub-net pp, member e, index of the tile pair inside the member's rows); pref = prefix sums of the
// members' pair counts (uniform in training, by bucket size in rollouts)
struct PairInfo {
    int pp, e, tp;
};
__device__ __forceinline__ PairInfo pair_info(const int* pref, int E, int pair) {
    const int per = pref[E];
    PairInfo r;
    r.pp = pair / per;
    const int rem = pair - r.pp * per;
    int e = 0;
    while (pref[e + 1] <= rem) ++e;
    r.e = e;
    r.tp = rem - pref[e];
    return r;
}

// bars: [0,S) W_FULL (local, tx), [S,2S) W_EMPTY (both CTAs, commit), [2S,3S) W_READY (leader, both halves landed),
//       [3S,3S+R) A_FULL (leader, 2 x 4 warps), [3S+R,3S+2R) A_EMPTY (both, commit), 3S+2R D_FULL (both, commit)

// Saved tile layout: [row quarter q][col / 4][32 rows][4 floats].  Inside a quarter it is the K-major operand layout
// (a 4-column group of 32 rows is 512 contiguous bytes, so a warp's store is one contiguous 512-byte run), and a
// 32-row slice of any range of column groups is ONE contiguous block: the weight-gradient kernel fetches it with a single
// TMA bulk copy (512-byte copies, one per group, cost ~180 cycles each and made that kernel 10x slower).
__device__ __forceinline__ long long tile_off(int q, int lane, int ncg, int cg) { return ((long long)(q * ncg + cg) * 32 + lane) * 4; }
__device__ __forceinline__ void save_chunk(float* tile, int q, int lane, int ncg, int c, const float* v) {
    if (XFLAG(4)) return;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(tile + tile_off(q, lane, ncg, c * 4 + j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

// A-ring bookkeeping of an epilogue warp: chunk `seq` of the global chunk sequence lives in slot seq % R
struct ARing {
    uint32_t smem0;  // shared-memory address of slot 0 + this thread's row offset
    uint32_t tmem0;  // tensor-memory address (this thread's lane) of the lo ring, 16 columns per slot
    uint64_t *full, *empty;
};
// returns the slot index of chunk `seq` once the MMAs that read the slot's previous chunk are complete
__device__ __forceinline__ uint32_t aring_acquire(const ARing& r, uint32_t seq) {
    const uint32_t slot = seq % TF_A_SLOTS, k = seq / TF_A_SLOTS;
    if (k >= 1) mbar_wait(&r.empty[slot], (k - 1) & 1);
    return slot;
}
template <bool LO_TMEM>
__device__ __forceinline__ void aring_publish(const ARing& r, uint32_t seq, int lane) {
    if (LO_TMEM) {
        v5_st_wait();
        tc_fence_before();
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) v3_arrive_leader(&r.full[seq % TF_A_SLOTS]);  // the leader CTA issues the pair's MMAs
}
// 16 values of this thread's row = one K chunk of the next GEMM: tf32 hi parts in shared memory (four 16-byte groups), lo parts in
// tensor memory (16 columns of this thread's lane) or, where the accumulators leave no columns, in shared memory as well
template <bool LO_TMEM>
__device__ __forceinline__ void split_store_chunk(const ARing& r, uint32_t slot, const float* v) {
    if (XFLAG(8)) return;
    const uint32_t dst = r.smem0 + slot * TF_A_SLOT;
    uint32_t lo[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            h[u] = __float_as_uint(v[4 * j + u]) & 0xFFFFE000u;
            lo[4 * j + u] = __float_as_uint(v[4 * j + u] - __uint_as_float(h[u]));
        }
        v3_sts128(dst + (uint32_t)j * TF_GROUP, h[0], h[1], h[2], h[3]);
        if (!LO_TMEM) v3_sts128(dst + TF_A_PART + (uint32_t)j * TF_GROUP, lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
    }
    if (LO_TMEM) tmem_st16(r.tmem0 + slot * 16, lo);
}
// one element pair (k, k + 1) of a chunk (the head writes its dz two values at a time)
template <bool LO_TMEM>
__device__ __forceinline__ void split_store_pair(const ARing& r, uint32_t slot, int k, float x, float y) {
    const uint32_t hx = __float_as_uint(x) & 0xFFFFE000u, hy = __float_as_uint(y) & 0xFFFFE000u;
    const uint32_t lx = __float_as_uint(x - __uint_as_float(hx)), ly = __float_as_uint(y - __uint_as_float(hy));
    const uint32_t at = r.smem0 + slot * TF_A_SLOT + (uint32_t)((k & 15) / 4) * TF_GROUP + (uint32_t)(k & 3) * 4;
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(at), "r"(hx), "r"(hy) : "memory");
    if (LO_TMEM) tmem_st2(r.tmem0 + slot * 16 + (uint32_t)(k & 15), lx, ly);
    else asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(at + TF_A_PART), "r"(lx), "r"(ly) : "memory");
}

// ACT is a template parameter: with a run-time activation switch the unrolled 16-element epilogue bodies compiled to a chain of
// branches per ELEMENT and the epilogue (not the tensor pipe, not the weight stream) paced the kernel at ~12k cycles per phase
// (profiles/r2_train_trace_attribution.txt).
template <int ACT, bool LO_TMEM, bool INFER>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TF_THREADS, 1) train_fused_kernel(const TrainP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile uint32_t a_ready_seq, w_ready_n;  // leader: chunks whose A operand / weight stage are known to be ready
    __shared__ int s_pref[TF_MAX_E + 1], s_rowoff[TF_MAX_E + 1];
    const Geom& g = p.g;
    const int S = p.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
    uint64_t* bar_wfull = bars;
    uint64_t* bar_wempty = bars + S;
    uint64_t* bar_wready = bars + 2 * S;
    uint64_t* bar_afull = bars + 3 * S;
    uint64_t* bar_aempty = bars + 3 * S + TF_A_SLOTS;
    uint64_t* bar_dfull = bars + 3 * S + 2 * TF_A_SLOTS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = v3_ctarank();

    if (INFER) {
        if ((int)threadIdx.x <= p.E) s_rowoff[threadIdx.x] = p.row_off[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        s_pref[0] = 0;
        for (int e = 0; e < p.E; ++e) {
            const int tiles = INFER ? (s_rowoff[e + 1] - s_rowoff[e] + TF_ROWS - 1) / TF_ROWS : 0;
            s_pref[e + 1] = s_pref[e] + (INFER ? (tiles + 1) / 2 : p.pairs_per_net);
        }
        a_ready_seq = 0;
        w_ready_n = 0;
        for (int s = 0; s < S; ++s) {
            // the leader's "stage landed" barrier also counts the peer's report that ITS half landed: one wait per chunk tells
            // the leader's forwarder that the whole stage is there (a separate both-halves barrier meant two waits in a row;
            // measured neutral -- the forwarder is not what paces the kernel -- but one barrier less per stage)
            mbar_init(&bar_wfull[s], rank == 0 ? 2 : 1);
            mbar_init(&bar_wempty[s], 1);
            mbar_init(&bar_wready[s], 2);  // (unused since round 2)
        }
        for (int s = 0; s < TF_A_SLOTS; ++s) {
            mbar_init(&bar_afull[s], 8);  // the four lane quarters of one column set, in both CTAs
            mbar_init(&bar_aempty[s], 1);
        }
        mbar_init(bar_dfull, 1);
        fence_barrier_init();
    }
    if (warp == TF_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    v3_cluster_sync();  // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t col_head = (uint32_t)(2 * g.HP), col_alo = (uint32_t)(2 * g.HP + g.NH);
    const int hi_only = p.passes == 1;
    const int n_pairs = p.P * s_pref[p.E];  // work items of a cluster: two row tiles of one net
    // A cluster takes a contiguous range of the work list (net-major), like the f16 rollout kernel: neighbouring clusters work on
    // different nets.  (Measured neutral against a strided assignment: the weight stream is not what paces the kernel -- with the
    // weight copies disabled the event trace of a 4 096-row step is unchanged, profiles/r2_train_trace_4096.txt.)
    const int n_clusters = gridDim.x / 2, cid = blockIdx.x / 2;
    const int pair0 = (int)((long long)cid * n_pairs / n_clusters), pair_end = (int)((long long)(cid + 1) * n_pairs / n_clusters);
    constexpr int pair_step = 1;

    if (warp == TF_WARP_W) {
        // ===================== weight producer: this CTA's half (N/2 output rows, hi | lo) of every 16-deep K chunk =================
        if (lane == 0) {
            uint32_t n = 0;
            TRACE_DECL;
            for (int pair = pair0; pair < pair_end; pair += pair_step) {
                const PairInfo pi = pair_info(s_pref, p.E, pair);
                const float* blob = p.packed + (long long)(pi.pp * p.E + pi.e) * g.blob_floats;
                for (int ph = 0; ph < g.n_ph; ++ph) {
                    const uint32_t half = 64u * (uint32_t)g.ph_n[ph];          // bytes of one CTA's half of a chunk: hi | lo
                    const uint32_t bytes = hi_only ? half / 2 : half;           // the lo part is not read by plain TF32
                    const float* src = blob + g.ph_off[ph];
                    const int nch = g.ph_k[ph] / 16;
                    for (int c = 0; c < nch; ++c, ++n) {
                        const uint32_t s = n % (uint32_t)S;
                        if (n >= (uint32_t)S) mbar_wait_sleep(&bar_wempty[s], ((n / (uint32_t)S) - 1) & 1, 32);
                        TRACE(0);
                        if (XFLAG(1)) {
                            mbar_arrive(&bar_wfull[s]);
                            continue;
                        }
                        mbar_expect_tx(&bar_wfull[s], bytes);
                        tma_bulk_g2s(smem + p.w_off + s * p.stage_bytes,
                                     reinterpret_cast<const uint8_t*>(src + (long long)c * 32 * g.ph_n[ph]) + rank * half, bytes, &bar_wfull[s]);
                    }
                }
            }
        }
    } else if (warp == TF_WARP_FWD) {
        // ===================== forwarder: tells the leader CTA that this CTA's half of a weight stage has landed ===================
        if (lane == 0) {
            uint32_t n = 0;
            for (int pair = pair0; pair < pair_end; pair += pair_step)
                for (int ph = 0; ph < g.n_ph; ++ph)
                    for (int c = 0; c < g.ph_k[ph] / 16; ++c, ++n) {
                        const uint32_t s = n % (uint32_t)S;
                        mbar_wait(&bar_wfull[s], (n / (uint32_t)S) & 1);
                        if (rank != 0) {
                            v3_arrive_leader(&bar_wfull[s]);  // the second arrival of the leader's barrier for this stage
                        } else {  // own bytes and the peer's report are in -> tell the issuer through the plain counter
                            __threadfence_block();
                            w_ready_n = n + 1;
                        }
                    }
        }
    } else if (warp == TF_WARP_AW) {
        // ===================== A-chunk waiter (leader): all 8 producer warps of a chunk have published it ==========================
        if (lane == 0 && rank == 0) {
            uint32_t seq = 0;
            for (int pair = pair0; pair < pair_end; pair += pair_step)
                for (int ph = 0; ph < g.n_ph; ++ph)
                    for (int c = 0; c < g.ph_k[ph] / 16; ++c, ++seq) {
                        mbar_wait(&bar_afull[seq % TF_A_SLOTS], (seq / TF_A_SLOTS) & 1);
                        __threadfence_block();
                        a_ready_seq = seq + 1;
                    }
        }
    } else if (warp == TF_WARP_MMA) {
        // ===================== MMA issuer: chases the epilogue through the A ring ================================================
        uint32_t n = 0, seq = 0, n_phase = 0;
        TRACE_DECL;
#ifdef CMRL_TRACE
        int mma_n = 0;
#endif
        // Descriptors are kept as (lo, hi) words: hi = stride field | version, lo = start address | leading-dimension field; an MMA
        // costs the issuing thread two 32-bit adds.
        const uint32_t a0 = smem_u32(smem), w0 = smem_u32(smem + p.w_off);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);                       // SBO = 128 B, descriptor version 1
        const uint32_t a_lbo = (TF_GROUP >> 4) << 16;
        for (int pair = pair0; rank == 0 && pair < pair_end; pair += pair_step) {  // the leader issues for the pair
            for (int ph = 0; ph < g.n_ph; ++ph, ++n_phase) {
                const int N = g.ph_n[ph];
                const uint32_t idesc = umma_idesc_tf32(2 * TF_ROWS, N);
                const uint32_t d_t = tmem + (g.ph_kind[ph] == PH_HEAD ? col_head : (n_phase & 1) ? (uint32_t)g.HP : 0u);
                const uint32_t b_lbo = 8u * (uint32_t)N;                           // bytes between K groups of this CTA's N/2 rows
                const uint32_t b_part = (4u * b_lbo) >> 4;                         // hi -> lo part of a half chunk (16-byte units)
                const uint32_t b_ks = (2u * b_lbo) >> 4, b_lbo_f = (b_lbo >> 4) << 16;
                const int nch = g.ph_k[ph] / 16;
                TRACE(2);
#pragma unroll 1
                for (int c = 0; c < nch; ++c, ++n, ++seq) {
                    const uint32_t s = n % (uint32_t)S, slot = seq % TF_A_SLOTS;
                    TRACE_MMA(n_phase == 1);
                    {
                        unsigned long long spins = 0;
                        while (a_ready_seq <= seq)
                            if (++spins > SPIN_LIMIT) __trap();
                        if (c == 0) TRACE(2);
                        TRACE_MMA(n_phase == 1);
                        while (w_ready_n <= n)
                            if (++spins > SPIN_LIMIT) __trap();
                    }
                    if (c == 0) TRACE(2);
                    if (c == nch - 1) TRACE(2);
                    TRACE_MMA(n_phase == 1);
                    __threadfence_block();
                    fence_proxy_async();  // (measured free: dropping it changes nothing)
                    tc_fence_after();
                    TRACE_MMA(n_phase == 1);
                    const uint32_t ah = ((a0 + slot * TF_A_SLOT) >> 4) | a_lbo, al = ah + (TF_A_PART >> 4);
                    const uint32_t bh = ((w0 + s * p.stage_bytes) >> 4) | b_lbo_f, bl = bh + b_part;
                    const uint32_t first = c == 0 ? 0u : 1u;
                    if (!XFLAG(2)) {
                        if (!hi_only) {  // 3xTF32, small terms first
                            const uint32_t alt = tmem + col_alo + slot * 16;  // lo part of the A chunk in tensor memory
                            if (LO_TMEM) tf32_mma2_ts_lohi_uni(d_t, alt, bh, desc_hi, idesc, first);
                            else tf32_mma2_lohi_uni(d_t, al, desc_hi, bh, desc_hi, idesc, first);
                            tf32_mma2_lohi_uni(d_t, ah, desc_hi, bl, desc_hi, idesc, 1u);
                            tf32_mma2_lohi_uni(d_t, ah, desc_hi, bh, desc_hi, idesc, 1u);
                            if (LO_TMEM) tf32_mma2_ts_lohi_uni(d_t, alt + 8, bh + b_ks, desc_hi, idesc, 1u);
                            else tf32_mma2_lohi_uni(d_t, al + (2 * TF_GROUP >> 4), desc_hi, bh + b_ks, desc_hi, idesc, 1u);
                            tf32_mma2_lohi_uni(d_t, ah + (2 * TF_GROUP >> 4), desc_hi, bl + b_ks, desc_hi, idesc, 1u);
                            tf32_mma2_lohi_uni(d_t, ah + (2 * TF_GROUP >> 4), desc_hi, bh + b_ks, desc_hi, idesc, 1u);
                        } else {
                            tf32_mma2_lohi_uni(d_t, ah, desc_hi, bh, desc_hi, idesc, first);
                            tf32_mma2_lohi_uni(d_t, ah + (2 * TF_GROUP >> 4), desc_hi, bh + b_ks, desc_hi, idesc, 1u);
                        }
                    }
                    TRACE_MMA(n_phase == 1);
                    v3_commit_uni(&bar_wempty[s]);  // cta_group::2 commits arrive in both CTAs of the pair
                    v3_commit_uni(&bar_aempty[slot]);
                    TRACE_MMA(n_phase == 1);
                }
                v3_commit_uni(bar_dfull);
                TRACE(2);
            }
        }
    } else {
        // ===================== epilogue warps: TMEM lane quarter q, column set cset =============================================
        const int q = warp & 3, cset = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        ARing ring{smem_u32(smem) + (uint32_t)row * 16, tmem + lane_off + col_alo, bar_afull, bar_aempty};
        const int Dn = g.Dn;
        // the next layer's bias input: hidden unit H is 1.0 -- patched into the chunk that holds it after the chunk is stored
        const int ones_c = g.H >> 4, ones_i = g.H & 15;
        uint32_t n_d = 0, seq_base = 0, n_phase = 0;
        TRACE_DECL;
#ifdef CMRL_TRACE
        int fine_n = 0;
#endif
        for (int pair = pair0; pair < pair_end; pair += pair_step) {
            const PairInfo pi = pair_info(s_pref, p.E, pair);
            const int pp = pi.pp, e = pi.e, net = pp * p.E + e;
            const int tile = pi.tp * 2 + (int)rank;  // may be a padding tile: no valid row
            const int grow = tile * TF_ROWS + row;
            const bool valid = grow < (INFER ? s_rowoff[e + 1] - s_rowoff[e] : p.B);
            // training: row grow of member e's dense batch; rollout: the env this bucketed row stands for
            const long long rr = !valid ? 0 : INFER ? (long long)p.perm[s_rowoff[e] + grow] : grow;
            float* sc = INFER ? nullptr : p.scratch + ((long long)net * 2 * p.pairs_per_net + tile) * g.item_floats;
            // ---- layer-0 operand: [obs | action | 1 | 0...] -------------------------------------------------------------------
            {
                const float* po = p.obs + (INFER ? 0 : e * p.obs_se) + rr * p.O;
                const float* pa = p.actn + (INFER ? 0 : e * p.act_se) + rr * p.A;
#pragma unroll 1
                for (int c = cset; c < g.K0P / 16; c += 4) {
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
                    if (!INFER) save_chunk(sc + g.off_x, q, lane, g.K0P / 4, c, v);
                    split_store_chunk<LO_TMEM>(ring, aring_acquire(ring, seq_base + c), v);
                    aring_publish<LO_TMEM>(ring, seq_base + c, lane);
                }
                seq_base += g.K0P / 16;
            }
            for (int ph = 0; ph < g.n_ph; ++ph, ++n_phase) {
                const int kind = g.ph_kind[ph], l = g.ph_layer[ph];
                const int out_chunks = ph + 1 < g.n_ph ? g.ph_k[ph + 1] / 16 : 0;  // A chunks this epilogue hands to the next GEMM
                const bool has_work = kind == PH_HEAD ? cset == 0 : cset < g.HP / 16;
                // Every warp waits for every phase, also one it has no chunk of: a parity wait is only meaningful when the
                // previous phase is known to be complete (a warp that skipped the head's wait walked straight through the
                // next one and read a stale accumulator).
                if (warp == 0) TRACE(1);
                mbar_wait(bar_dfull, n_d & 1);
                ++n_d;
                if (has_work) {
                    tc_fence_after();
                    if (warp == 0) TRACE(1);
                    const uint32_t acc_t = tmem + lane_off + ((n_phase & 1) ? (uint32_t)g.HP : 0u);
                    if (kind == PH_FWD) {
                        // All 16 warps reach this loop together, so an un-pipelined body runs the SM's resources one after the
                        // other (TMEM read, global stores, MUFU, shared-memory stores: ~3.3k cycles per round of 4 chunks).  Order
                        // per chunk: the NEXT chunk's TMEM load is in flight, the activations (MUFU) are issued BEFORE the global
                        // stores of z so that both pipes work at the same time.
                        float* zt = INFER ? nullptr : sc + g.off_z[l];
                        uint32_t raw[16];
                        __syncwarp();
                        v5_ld16_issue(acc_t + (uint32_t)cset * 16, raw);
#pragma unroll 1
                        for (int c = cset; c < g.HP / 16; c += 4) {
                            float z[16], a[16];
                            TRACE_FINE(warp == 0 && ph == 1);
                            v3_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) z[i] = __uint_as_float(raw[i]);
                            if (c + 4 < g.HP / 16) {
                                __syncwarp();
                                v5_ld16_issue(acc_t + (uint32_t)(c + 4) * 16, raw);
                            }
                            TRACE_FINE(warp == 0 && ph == 1);
#pragma unroll
                            for (int i = 0; i < 16; ++i) a[i] = act_fwd_fast(z[i], ACT);
                            TRACE_FINE(warp == 0 && ph == 1);
                            if (!INFER) save_chunk(zt, q, lane, g.HP / 4, c, z);  // pre-activations: the dz chain and the dw kernel re-activate
                            TRACE_FINE(warp == 0 && ph == 1);
                            if (c == ones_c) {  // the next layer's bias input: hidden unit H is 1.0 (a warp-uniform patch of one element)
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    if (i == ones_i) a[i] = 1.0f;
                            }
                            const uint32_t slot = aring_acquire(ring, seq_base + c);
                            TRACE_FINE(warp == 0 && ph == 1);
                            split_store_chunk<LO_TMEM>(ring, slot, a);
                            TRACE_FINE(warp == 0 && ph == 1);
                            aring_publish<LO_TMEM>(ring, seq_base + c, lane);
                            TRACE_FINE(warp == 0 && ph == 1);
                        }
                    } else if (INFER) {
                        // ---- rollout head: mean (+ residual) + sqrt(exp(soft-clamped logvar)) * noise, scattered to the env's row
                        // (fake_env.py:205-215 over plain_transition.py:111-120); columns 2d = mean_d, 2d+1 = logvar_d
                        const uint32_t h_t = tmem + lane_off + col_head;
#pragma unroll 1
                        for (int d = 0; d < Dn; ++d) {
                            uint32_t r0, r1;
                            __syncwarp();
                            tmem_ld2(h_t + (uint32_t)(2 * d), r0, r1);
                            v3_ld_wait();
                            if (valid) {
                                const int od = p.layout == CMRL_HEAD_PARALLEL ? pp : d;
                                float m = __uint_as_float(r0);
                                if (p.residual) m += p.obs[rr * p.O + od];
                                if (p.noise) {
                                    const int bi = p.bounds_n == 1 ? 0 : od;
                                    const float lv = soft_clamp_logvar(__uint_as_float(r1), p.min_logvar[bi], p.max_logvar[bi]);
                                    m += sqrtf(expf(lv)) * p.noise[rr * p.D + od];
                                }
                                p.pred[(long long)e * p.pred_se + rr * p.D + od] = m;
                            }
                        }
                    } else if (kind == PH_DX) {
                        // dz = (dz_next . W^T) * act'(z): act'(z) does not depend on the accumulator, so it is evaluated (MUFU) while
                        // the TMEM load is in flight; the A chunk is published before dz is saved; the next chunk's z is fetched last.
                        const float* zt = sc + g.off_z[l];
                        float* dt = sc + g.off_dz[l];
                        float z[16];
                        auto load_z = [&](int c) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (XFLAG(4)) {
                                    z[4 * j] = z[4 * j + 1] = z[4 * j + 2] = z[4 * j + 3] = 0.5f;
                                    continue;
                                }
                                const float4 t4 = *reinterpret_cast<const float4*>(zt + tile_off(q, lane, g.HP / 4, c * 4 + j));
                                z[4 * j] = t4.x; z[4 * j + 1] = t4.y; z[4 * j + 2] = t4.z; z[4 * j + 3] = t4.w;
                            }
                        };
                        load_z(cset);
#pragma unroll 1
                        for (int c = cset; c < g.HP / 16; c += 4) {
                            uint32_t raw[16];
                            __syncwarp();
                            v5_ld16_issue(acc_t + (uint32_t)c * 16, raw);
#pragma unroll
                            for (int i = 0; i < 16; ++i) z[i] = act_grad_fast(z[i], ACT);
                            v3_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) z[i] *= __uint_as_float(raw[i]);
                            if (out_chunks > 0) {
                                split_store_chunk<LO_TMEM>(ring, aring_acquire(ring, seq_base + c), z);
                                aring_publish<LO_TMEM>(ring, seq_base + c, lane);
                            }
                            save_chunk(dt, q, lane, g.HP / 4, c, z);
                            if (c + 4 < g.HP / 16) load_z(c + 4);
                        }
                    } else {
                        // ---- Gaussian head + NLL + their gradient, one output per iteration (columns 2d = mean_d, 2d+1 = logvar_d);
                        // the NH-wide dz of the head is chunk 0 (and 1) of the next GEMM
                        float* dt = sc + g.off_dz[g.L];
                        const uint32_t h_t = tmem + lane_off + col_head;
                        uint32_t slot = 0;
#pragma unroll 1
                        for (int d = 0; d < g.NH / 2; ++d) {
                            if ((d & 7) == 0) slot = aring_acquire(ring, seq_base + (d >> 3));
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
                            split_store_pair<LO_TMEM>(ring, slot, k, dm, dl);
                            if ((d & 7) == 7) aring_publish<LO_TMEM>(ring, seq_base + (d >> 3), lane);
                        }
                    }
                    tc_fence_before();  // orders this warp's accumulator reads before the MMAs that overwrite it two phases on
                    if (warp == 0) TRACE(1);
                }
                seq_base += out_chunks;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    v3_cluster_sync();  // no CTA leaves while its peer may still arrive on its barriers or read its shared memory
    if (warp == TF_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ================================================= weight gradients ========================================================
constexpr int DW_THREADS = 18 * 32;  // (the dw kernel has no forwarder warp)
constexpr int DW_NT = 112;          // output columns per CTA
constexpr int DW_SUB = 32;          // rows (= K) per pipeline step
constexpr int DW_DRAIN = 4;         // pipeline steps accumulated in tensor memory before the sum moves to registers (128 rows, 48 MMAs)
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
    int P, E, tiles_per_net, tile_slots, act;  // real row tiles / tile slots per net in the scratch (padded to cluster pairs)
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

template <int ACT>
__global__ void __launch_bounds__(DW_THREADS, 1) train_dw_kernel(const DwP p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    const Geom& g = p.g;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DW_BAR_OFF);
    uint64_t* bar_rfull = bars;                      // [2] tx
    uint64_t* bar_rempty = bars + 2;                 // [2] 16 warps
    uint64_t* bar_ofull = bars + 4;                  // [2] 16 warps
    uint64_t* bar_oempty = bars + 6;                 // [2] commit
    uint64_t* bar_accfull = bars + 8;                // [2] commit: the MMAs of an accumulation chunk are complete
    uint64_t* bar_accempty = bars + 10;              // [2] 16 warps: the chunk was added to the register sums
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int net = blockIdx.y;
    // ---- which (layer, feature tile, column tile, batch slice) ------------------------------------------------------------
    const int slice = blockIdx.x % p.ksplit;
    const int w = blockIdx.x / p.ksplit;
    int l = 0;
    while (w >= p.work_off[l + 1]) ++l;
    const int wi = w - p.work_off[l];
    const int mt = wi / p.n_tiles[l], nt = wi % p.n_tiles[l];
    const int a_cols = l == 0 ? g.K0P : g.HP;            // saved input tile of layer l: x, or z_{l-1} (re-activated here)
    const int z_cols = l == g.L ? g.NH : g.HP;           // dz_l
    const long long a_off = l == 0 ? g.off_x : g.off_z[l - 1];
    const long long z_off = g.off_dz[l];
    const int ga = min(32, a_cols / 4 - mt * 32);         // 4-feature groups of this CTA's A rows
    const int NT = min(DW_NT, z_cols - nt * DW_NT);      // output columns of this CTA
    const int gb = NT / 4;
    const int t0 = slice * p.tiles_per_slice, t1 = min(p.tiles_per_net, t0 + p.tiles_per_slice);
    const int n_sub = max(0, t1 - t0) * (TF_ROWS / DW_SUB);
    const bool activate = l > 0;  // layer 0 reads the saved inputs, the others re-activate saved pre-activations
    const int ones_f = l == 0 ? -1 : g.H - mt * TF_ROWS;  // feature (local to this tile) that is the bias input of layer l

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_rfull[s], 1);
            mbar_init(&bar_rempty[s], TF_EPI_WARPS);
            mbar_init(&bar_ofull[s], TF_EPI_WARPS);
            mbar_init(&bar_oempty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_accfull[s], 1);
            mbar_init(&bar_accempty[s], TF_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == TF_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
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
    const float* base = p.scratch + ((long long)net * p.tile_slots) * g.item_floats;

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
            const int ch = i / DW_DRAIN, cb = ch & 1;  // accumulation chunk and its tensor-memory buffer
            if (i % DW_DRAIN == 0 && ch >= 2) mbar_wait(&bar_accempty[cb], ((ch >> 1) - 1) & 1);
            mbar_wait(&bar_ofull[s], (i >> 1) & 1);
            tc_fence_after();
            TRACE(2);
            const uint32_t op = smem_u32(smem + DW_RAW_STAGES * DW_RAW + s * DW_OP);
            const uint32_t a_hi = op, a_lo = op + DW_OP_A, b_hi = op + 2 * DW_OP_A, b_lo = b_hi + DW_OP_B;
            const uint32_t d_t = tmem + (uint32_t)cb * 128u;
#pragma unroll
            for (int ks = 0; ks < DW_SUB / 8; ++ks) {
                const uint32_t ko = (uint32_t)ks * 2 * DW_KG;
                const uint64_t ah = umma_desc(a_hi + ko, DW_KG, DW_MG), bh = umma_desc(b_hi + ko, DW_KG, DW_MG);
                const uint32_t first = (i % DW_DRAIN == 0 && ks == 0) ? 0u : 1u;
                const uint64_t al = umma_desc(a_lo + ko, DW_KG, DW_MG), bl = umma_desc(b_lo + ko, DW_KG, DW_MG);
                tf32_mma_ss_uni(d_t, al, bh, idesc, first);
                tf32_mma_ss_uni(d_t, ah, bl, idesc, 1u);
                tf32_mma_ss_uni(d_t, ah, bh, idesc, 1u);
            }
            umma_commit_uni(&bar_oempty[s]);
            if (i % DW_DRAIN == DW_DRAIN - 1 || i == n_sub - 1) umma_commit_uni(&bar_accfull[cb]);
        }
    } else {
        // ===================== transposers, then the epilogue ====================================================================
        const int tid = threadIdx.x;  // 0..511
        // Running sums of this thread's part of the tile (TMEM lane = feature row q*32+lane, 16-column chunks cset and cset+4).
        // The tensor cores accumulate with truncation: the error of one accumulator grows LINEARLY with the number of MMAs
        // chained into it (measured 9e-6 relative after 256 rows, 4.8e-5 after 2 048), so a chunk of DW_DRAIN sub-steps is
        // accumulated in tensor memory and then added, rounded to nearest, to fp32 registers -- two accumulators alternate,
        // and a chunk is fetched two sub-steps after its last MMA was issued, so nothing waits.
        const int q = warp & 3, cset = warp >> 2;
        float sum[2][16];
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[k][j] = 0.f;
        const int n_chunks = (n_sub + DW_DRAIN - 1) / DW_DRAIN;
        int drained = 0;
        auto drain = [&](int ch) {
            const int cb = ch & 1;
            mbar_wait(&bar_accfull[cb], (ch >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = cset + 4 * k;
                if (c < NT / 16) {
                    uint32_t rawv[16];
                    __syncwarp();
                    v5_ld16_issue(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cb * 128u + (uint32_t)c * 16, rawv);
                    v3_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum[k][j] += __uint_as_float(rawv[j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_accempty[cb]);
        };
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
                float e0 = v.x, e1 = v.y, e2 = v.z, e3 = v.w;
                if (isa && activate) {  // the layer input is the activation of the saved pre-activations (+ its bias column of ones)
                    e0 = act_fwd_fast(e0, ACT); e1 = act_fwd_fast(e1, ACT); e2 = act_fwd_fast(e2, ACT); e3 = act_fwd_fast(e3, ACT);
                    if ((ones_f >> 2) == gl) {
                        const int oj = ones_f & 3;
                        e0 = oj == 0 ? 1.0f : e0; e1 = oj == 1 ? 1.0f : e1; e2 = oj == 2 ? 1.0f : e2; e3 = oj == 3 ? 1.0f : e3;
                    }
                }
                auto put = [&](uint32_t j, float x) {
                    const uint32_t h = __float_as_uint(x) & 0xFFFFE000u;
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(hi_base + j * 16u), "r"(h) : "memory");
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(lo_base + j * 16u), "r"(__float_as_uint(x - __uint_as_float(h))) : "memory");
                };
                put(0, e0); put(1, e1); put(2, e2); put(3, e3);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_ofull[s]);
                mbar_arrive(&bar_rempty[s]);
            }
            if (warp == 0) TRACE(1);
            if (i >= 1 && drained < (i - 1) / DW_DRAIN) drain(drained++);
        }
        while (drained < n_chunks) drain(drained++);
        // ---- epilogue: this thread's sums = feature row q*32+lane of the tile.  Rows of the gradient are far apart in memory, so
        // the tile goes through shared memory and is written out with the lanes along the columns.  The operand stages are
        // reused for it: the last chunk's accumulator-full barrier says every MMA that read them is complete.
        if (warp == 0) TRACE(1);
        asm volatile("bar.sync 1, %0;" ::"r"(TF_EPI_WARPS * 32) : "memory");  // every epilogue warp has seen that barrier
        float* otile = reinterpret_cast<float*>(smem + DW_RAW_STAGES * DW_RAW);
        constexpr int OS = DW_NT + 4;  // floats per staged row (16-byte aligned, conflict-free for the float4 writes)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int c = cset + 4 * k;
            if (c < NT / 16) {
                float4* dst = reinterpret_cast<float4*>(otile + (q * 32 + lane) * OS + c * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_float4(sum[k][4 * j], sum[k][4 * j + 1], sum[k][4 * j + 2], sum[k][4 * j + 3]);
            }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(TF_EPI_WARPS * 32) : "memory");  // the 16 epilogue warps only
        float* wsp = p.ksplit > 1 ? p.ws + ((long long)slice * gridDim.y + net) * p.ext_total + p.ext_off[l] : nullptr;
        const int zc = l == g.L ? g.NH : p.out_l[l];  // columns of the extended gradient in the workspace
        for (int r = warp; r < TF_ROWS; r += TF_EPI_WARPS) {
            const int i_row = mt * TF_ROWS + r;
            if (i_row > p.in_l[l]) break;
            for (int cc = lane; cc < NT; cc += 32) {
                const int n = nt * DW_NT + cc;
                const float v = otile[r * OS + cc];
                if (wsp) {
                    if (n < zc) wsp[(long long)i_row * zc + n] = v;
                } else {
                    dw_store(p, net, l, i_row, n, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TF_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
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
extern "C" int cmrl_trace_flags(int flags) {
    CMRL_CUDA(cudaMemcpyToSymbol(cmrl::tr::g_trace_flags, &flags, sizeof(flags)));
    return 0;
}
#endif

extern "C" long long cmrl_train_fused_sizes(int L, int in_size, int hidden, int head_outputs, int P, int E, int B, long long* blob_floats,
                                            long long* scratch_floats, long long* dw_ws_floats) {
    Geom g;
    if (make_geom(g, L, in_size, hidden, head_outputs) != 0) return -1;
    const long long nets = (long long)P * E, tiles = (B + TF_ROWS - 1) / TF_ROWS;
    if (blob_floats) *blob_floats = g.blob_floats;
    if (scratch_floats) *scratch_floats = nets * ((tiles + 1) / 2 * 2) * g.item_floats;  // tile slots padded to cluster pairs
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

// shared-memory plan, kernel attributes and launch of train_fused_kernel (training step or rollout mode)
template <bool INFER>
static int launch_fused_t(TrainP& p, int act, int clusters, uint32_t smem_bytes, cudaStream_t st) {
#define CMRL_TF_LAUNCH(A)                                                                              \
    do {                                                                                               \
        if (p.lo_tmem) train_fused_kernel<A, true, INFER><<<2 * clusters, TF_THREADS, smem_bytes, st>>>(p); \
        else train_fused_kernel<A, false, INFER><<<2 * clusters, TF_THREADS, smem_bytes, st>>>(p);          \
    } while (0)
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(train_fused_kernel<CMRL_ACT_IDENTITY, false, INFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CMRL_CUDA(cudaFuncSetAttribute(train_fused_kernel<CMRL_ACT_RELU, false, INFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CMRL_CUDA(cudaFuncSetAttribute(train_fused_kernel<CMRL_ACT_SILU, false, INFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CMRL_CUDA(cudaFuncSetAttribute(train_fused_kernel<CMRL_ACT_IDENTITY, true, INFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CMRL_CUDA(cudaFuncSetAttribute(train_fused_kernel<CMRL_ACT_RELU, true, INFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CMRL_CUDA(cudaFuncSetAttribute(train_fused_kernel<CMRL_ACT_SILU, true, INFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_once.done(attr_dev);
    }
    if (act == CMRL_ACT_SILU) CMRL_TF_LAUNCH(CMRL_ACT_SILU);
    else if (act == CMRL_ACT_RELU) CMRL_TF_LAUNCH(CMRL_ACT_RELU);
    else CMRL_TF_LAUNCH(CMRL_ACT_IDENTITY);
#undef CMRL_TF_LAUNCH
    CMRL_CHECK_LAUNCH(INFER ? "mlp_rollout_tf32" : "train_fused");
    return 0;
}
static int launch_fused(TrainP& p, int act, long long max_pairs, bool infer, cudaStream_t st) {
    p.w_off = TF_A_SLOTS * TF_A_SLOT;
    p.stage_bytes = 64u * (uint32_t)p.g.HP;  // one CTA's half (N/2 rows) of a 16-deep chunk, hi | lo
    int stages = (int)((SMEM_LIMIT - 512 - p.w_off) / p.stage_bytes);
    if (stages > 10) stages = 10;
    CMRL_REQUIRE(stages >= 2, "train_fused: hidden width %d leaves no room for a weight ring", p.g.H);
    p.stages = stages;
    p.bar_off = p.w_off + (uint32_t)stages * p.stage_bytes;
    const uint32_t smem_bytes = p.bar_off + (3 * stages + 2 * TF_A_SLOTS + 1) * 8;
    // tensor-memory columns: two accumulators, the head accumulator, then (if they fit) the lo parts of the A-chunk ring
    p.lo_tmem = 2 * p.g.HP + p.g.NH + 16 * TF_A_SLOTS <= 512 ? 1 : 0;
    CMRL_REQUIRE(max_pairs < (1ll << 30), "train_fused: too many row tiles");
    const int clusters = (int)(max_pairs < num_sms() / 2 ? (max_pairs < 1 ? 1 : max_pairs) : num_sms() / 2);
    return infer ? launch_fused_t<true>(p, act, clusters, smem_bytes, st) : launch_fused_t<false>(p, act, clusters, smem_bytes, st);
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
    CMRL_REQUIRE(E <= TF_MAX_E, "train_fused: at most %d members", TF_MAX_E);
    CMRL_REQUIRE(bounds_n == 1 || bounds_n == head_dim, "train_fused: logvar bounds must have 1 or head_dim entries");
    CMRL_REQUIRE(act == CMRL_ACT_IDENTITY || act == CMRL_ACT_RELU || act == CMRL_ACT_SILU, "train_fused: unsupported activation");
    p.packed = packed; p.P = P; p.E = E; p.B = B; p.O = O; p.A = A; p.layout = layout; p.act = act; p.residual = residual; p.passes = passes;
    p.obs = obs; p.actn = actn; p.target = target; p.obs_se = obs_se; p.act_se = act_se; p.tgt_se = tgt_se; p.tgt_ld = tgt_ld;
    p.min_logvar = min_logvar; p.max_logvar = max_logvar; p.bounds_n = bounds_n; p.reg = reg; p.g_scale = g_scale;
    p.loss = loss; p.loss_ld = head_dim; p.scratch = scratch;
    p.tiles_per_net = (B + TF_ROWS - 1) / TF_ROWS;
    p.pairs_per_net = (p.tiles_per_net + 1) / 2;
    return launch_fused(p, act, (long long)P * E * p.pairs_per_net, false, (cudaStream_t)stream);
}

/* Rollout (inference) mode of the same kernel: the fp32-grade tier of VecFakeEnv.get_dynamics_predict on tensor cores. */
extern "C" int cmrl_mlp_rollout_tf32(const float* packed, int L, int P, int E, int N, int O, int A, int hidden, int head_dim, int layout,
                                     int act, int residual, const float* obs, const float* actn, const int* perm, const int* offsets,
                                     const float* min_logvar, const float* max_logvar, int bounds_n, const float* noise, float* pred,
                                     long long pred_member_stride, void* stream) {
    TrainP p{};
    const int Dn = layout == CMRL_HEAD_PARALLEL ? 1 : head_dim;
    if (make_geom(p.g, L, O + A, hidden, Dn) != 0) return -1;
    p.g.n_ph = L + 1;  // forward phases only; the blob keeps the training layout (the backward operands are not read)
    CMRL_REQUIRE(packed && obs && actn && perm && offsets && pred, "mlp_rollout_tf32: null pointer");
    CMRL_REQUIRE(layout == CMRL_HEAD_PLAIN || layout == CMRL_HEAD_PARALLEL, "mlp_rollout_tf32: bad head layout");
    CMRL_REQUIRE(layout == CMRL_HEAD_PLAIN ? P == 1 : P == head_dim, "mlp_rollout_tf32: P must be 1 (plain head) or head_dim (parallel head)");
    CMRL_REQUIRE(E >= 1 && E <= TF_MAX_E && N >= 0, "mlp_rollout_tf32: 1..%d members", TF_MAX_E);
    CMRL_REQUIRE(!noise || (min_logvar && max_logvar && (bounds_n == 1 || bounds_n == head_dim)),
                 "mlp_rollout_tf32: sampling needs logvar bounds with 1 or head_dim entries");
    CMRL_REQUIRE(act == CMRL_ACT_IDENTITY || act == CMRL_ACT_RELU || act == CMRL_ACT_SILU, "mlp_rollout_tf32: unsupported activation");
    if (N == 0) return 0;
    p.packed = packed; p.P = P; p.E = E; p.B = 0; p.O = O; p.A = A; p.layout = layout; p.act = act; p.residual = residual; p.passes = 3;
    p.obs = obs; p.actn = actn; p.min_logvar = min_logvar; p.max_logvar = max_logvar; p.bounds_n = bounds_n;
    p.perm = perm; p.row_off = offsets; p.noise = noise; p.pred = pred; p.pred_se = pred_member_stride; p.D = head_dim;
    // every member's rows round up to whole tile pairs: at most N / 256 + E pairs per sub-net
    return launch_fused(p, act, (long long)P * (N / (2 * TF_ROWS) + E), true, (cudaStream_t)stream);
}

extern "C" int cmrl_train_dw(const float* scratch, int L, int P, int E, int B, int in_size, int hidden, int head_outputs, int act,
                             float* const* dw, float* const* db, const float* mask, long long m_sp, long long m_se, float* workspace,
                             long long workspace_floats, void* stream) {
    DwP p{};
    if (make_geom(p.g, L, in_size, hidden, head_outputs) != 0) return -1;
    CMRL_REQUIRE(scratch && dw && db, "train_dw: null pointer");
    const Geom& g = p.g;
    CMRL_REQUIRE(act == CMRL_ACT_IDENTITY || act == CMRL_ACT_RELU || act == CMRL_ACT_SILU, "train_dw: unsupported activation");
    p.scratch = scratch; p.P = P; p.E = E; p.act = act; p.mask = mask; p.m_sp = m_sp; p.m_se = m_se;
    p.tiles_per_net = (B + TF_ROWS - 1) / TF_ROWS;
    p.tile_slots = (p.tiles_per_net + 1) / 2 * 2;
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
    // Batch slices.  The launch takes ceil(CTAs / SMs) waves (one CTA per SM); a CTA costs its row tiles plus a fixed part
    // (set-up, pipeline fill, staging and writing its 128 x 112 tile: ~0.8 of a row tile, measured with the event trace), and
    // slicing adds the reduction launch and ks x (workspace write + read).  The first version of this model charged slicing
    // one tile time in total: at 4096 rows per member it cut the batch into 32 slices, 3 584 CTAs of one row tile each plus
    // 400 MB of workspace traffic -- 0.28 ms where one wave of 112 CTAs takes 0.12 ms.
    int ksplit = 1;
    if (p.tiles_per_net > 1 && workspace != nullptr) {
        const long long ctas = (long long)work * nets, sms = num_sms();
        long long fit = workspace_floats / (ext * nets);
        if (fit > 64) fit = 64;
        if (fit > p.tiles_per_net) fit = p.tiles_per_net;
        const double fixed = 0.8, reduce_launch = 1.0, reduce_per_slice = 0.35;
        double best = (double)((ctas + sms - 1) / sms) * (p.tiles_per_net + fixed);
        for (long long ks = 2; ks <= fit; ++ks) {
            const double cost = (double)((ctas * ks + sms - 1) / sms) * ((p.tiles_per_net + ks - 1) / ks + fixed) + reduce_launch + reduce_per_slice * ks;
            if (cost < best) {
                best = cost;
                ksplit = (int)ks;
            }
        }
    }
    p.tiles_per_slice = (p.tiles_per_net + ksplit - 1) / ksplit;
    ksplit = (p.tiles_per_net + p.tiles_per_slice - 1) / p.tiles_per_slice;
    p.ksplit = ksplit;
    p.ws = ksplit > 1 ? workspace : nullptr;
    static cmrl::PerDeviceOnce attr_once;
    int attr_dev = 0;
    if (attr_once.pending(&attr_dev)) {
        CMRL_CUDA(cudaFuncSetAttribute(train_dw_kernel<CMRL_ACT_IDENTITY>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CMRL_CUDA(cudaFuncSetAttribute(train_dw_kernel<CMRL_ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        CMRL_CUDA(cudaFuncSetAttribute(train_dw_kernel<CMRL_ACT_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_once.done(attr_dev);
    }
    dim3 grid(work * ksplit, nets);
    if (act == CMRL_ACT_SILU) train_dw_kernel<CMRL_ACT_SILU><<<grid, DW_THREADS, DW_SMEM, (cudaStream_t)stream>>>(p);
    else if (act == CMRL_ACT_RELU) train_dw_kernel<CMRL_ACT_RELU><<<grid, DW_THREADS, DW_SMEM, (cudaStream_t)stream>>>(p);
    else train_dw_kernel<CMRL_ACT_IDENTITY><<<grid, DW_THREADS, DW_SMEM, (cudaStream_t)stream>>>(p);
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
