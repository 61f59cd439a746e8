// K5: the numpy tail of VecFakeEnv.step_wait (cmrl/models/fake_env.py:97-142,186-215) as HBM-bound
// device kernels: member draw + Gaussian noise (Philox), bucketing of envs by drawn member, input
// gather, head + sampling + scatter, MOPO penalty, and the episode bookkeeping / reset tail that
// writes SB3-layout buffers.
#include "common.cuh"

namespace cmrl {

constexpr int RT = 256;

static inline int blocks_for(long long n, int threads = RT) {
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    return (int)b;
}

// ------------------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__device__ __forceinline__ float u32_to_unit_open(uint32_t x) {  // (0,1]
    return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// Box-Muller with the fast intrinsics: the draws are statistical (rng_mode="device"), an absolute error of ~1e-6 in a
// standard normal is far below anything a rollout can see, and the accurate logf / sincospif made this kernel spend four times
// the instructions of its two Philox blocks.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float* n0, float* n1) {
    const float u = u32_to_unit_open(a), v = u32_to_unit_open(b);
    const float r = sqrtf(-2.0f * __logf(u));
    float s, c;
    __sincosf(6.283185307179586f * v - 3.141592653589793f, &s, &c);  // argument in (-pi, pi]: the intrinsic's accurate range
    *n0 = r * c;
    *n1 = r * s;
}

// One thread per env.  Counter block 0 of an env's Philox stream: word x draws the member, y/z and w/(block 1).x feed the first
// two Box-Muller pairs; every further block feeds two more pairs.  D = 5 takes two blocks (three with one block per four
// normals and a separate member block).  The kernel is bound by integer issue slots (Philox4x32-10: ~90 instructions per block),
// not by the 1 + 4 D bytes it writes per env.
__global__ void draw_member_noise_kernel(unsigned long long seed, unsigned long long stream_id,
                                         const unsigned long long* __restrict__ stream_base, int N,
                                         const uint8_t* __restrict__ elite, int n_elite, int D,
                                         uint8_t* __restrict__ member_idx, float* __restrict__ noise) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (stream_base) stream_id += *stream_base;
    uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t s_lo = (uint32_t)stream_id, s_hi = (uint32_t)(stream_id >> 32);
    uint4 r = philox4x32_10(make_uint4((uint32_t)n, 0u, s_lo, s_hi), key);
    if (member_idx) member_idx[n] = elite[(uint32_t)(((unsigned long long)r.x * (unsigned long long)n_elite) >> 32)];
    if (noise) {
        // word queue: r.y r.z r.w | next block x y z w | ...
        uint32_t w[3] = {r.y, r.z, r.w};
        int have = 3, blk = 1;
        float* dst = noise + (long long)n * D;
        for (int d0 = 0; d0 < D; d0 += 2) {
            if (have < 2) {
                const uint4 q = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)blk++, s_lo, s_hi), key);
                if (have == 1) {  // one word left over: pair it with q.x, keep q.y q.z q.w
                    float z0, z1;
                    box_muller(w[0], q.x, &z0, &z1);
                    dst[d0] = z0;
                    if (d0 + 1 < D) dst[d0 + 1] = z1;
                    w[0] = q.y; w[1] = q.z; w[2] = q.w;
                    have = 3;
                    continue;
                }
                w[0] = q.x; w[1] = q.y; w[2] = q.z;  // (have == 0; q.w is dropped: keeps the queue at three words)
                have = 3;
            }
            float z0, z1;
            box_muller(w[0], w[1], &z0, &z1);
            dst[d0] = z0;
            if (d0 + 1 < D) dst[d0 + 1] = z1;
            w[0] = w[2];
            have -= 2;
        }
    }
}

__global__ void draw_uniform_index_kernel(unsigned long long seed, unsigned long long stream_id,
                                          const unsigned long long* __restrict__ stream_base, int N, unsigned int upper,
                                          int* __restrict__ out) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (stream_base) stream_id += *stream_base;
    uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint4 r = philox4x32_10(make_uint4((uint32_t)n, 0x7fffffffu, (uint32_t)stream_id, (uint32_t)(stream_id >> 32)), key);
    out[n] = (int)(((unsigned long long)r.x * (unsigned long long)upper) >> 32);
}

// ------------------------------------------------------------------------------------------- bucketing
constexpr int MAX_E = 256;

// Counting sort of envs by drawn member in two launches, no global atomics, STABLE (ascending env id inside a member's
// range, so the permutation is deterministic).  The env range is cut into W contiguous warp chunks (8 warps per block).
//   pass 1: every warp counts its chunk per member (match.any groups the lanes of one member) -> counts[W][E]
//   pass 2: every block reduces counts to its members' offsets and its warps' start positions, then every warp walks its
//           chunk again and writes perm[start + rank]
constexpr int BUCKET_WARPS = 8;
constexpr int BUCKET_MAX_W = CMRL_BUCKET_SCRATCH_WARPS;

__device__ __forceinline__ void bucket_warp_range(int N, int W, int gw, int& n0, int& n1) {
    const int wchunk = (N + W - 1) / W;
    n0 = min(N, gw * wchunk);
    n1 = min(N, n0 + wchunk);
}

__global__ void bucket_count_kernel(const uint8_t* __restrict__ idx, int N, int E, int W, int* __restrict__ counts) {
    extern __shared__ int h[];  // [BUCKET_WARPS][E]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < BUCKET_WARPS * E; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const int gw = blockIdx.x * BUCKET_WARPS + warp;
    int n0, n1;
    bucket_warp_range(N, W, gw, n0, n1);
    int* hw = h + warp * E;
    for (int nb = n0; nb < n1; nb += 32) {
        const int n = nb + lane;
        const int e = n < n1 ? (int)idx[n] : 0x7fffffff;
        const unsigned grp = __match_any_sync(0xffffffffu, e);
        if (e < E && lane == __ffs(grp) - 1) hw[e] += __popc(grp);
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BUCKET_WARPS * E; i += blockDim.x) counts[(size_t)blockIdx.x * BUCKET_WARPS * E + i] = h[i];
}

__global__ void bucket_place_kernel(const uint8_t* __restrict__ idx, int N, int E, int W, const int* __restrict__ counts,
                                    int* __restrict__ perm, int* __restrict__ offsets) {
    extern __shared__ int sm[];      // before[E] | total[E] | cur[BUCKET_WARPS][E] | red[8]
    int* before = sm;                // envs of member e in the chunks of earlier blocks
    int* total = sm + E;
    int* cur = sm + 2 * E;
    int* red = sm + 2 * E + BUCKET_WARPS * E;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w_first = blockIdx.x * BUCKET_WARPS;
    if (E >= 64) {  // one thread per member walks all warp chunks
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            int bsum = 0, tsum = 0;
            for (int w = 0; w < W; ++w) {
                const int c = counts[(size_t)w * E + e];
                tsum += c;
                if (w < w_first) bsum += c;
            }
            before[e] = bsum;
            total[e] = tsum;
        }
    } else {  // few members: block-wide reduction per member
        for (int e = 0; e < E; ++e) {
            int bsum = 0, tsum = 0;
            for (int w = threadIdx.x; w < W; w += blockDim.x) {
                const int c = counts[(size_t)w * E + e];
                tsum += c;
                if (w < w_first) bsum += c;
            }
            for (int o = 16; o > 0; o >>= 1) {
                bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
                tsum += __shfl_xor_sync(0xffffffffu, tsum, o);
            }
            if (lane == 0) {
                red[warp] = bsum;
                red[BUCKET_WARPS + warp] = tsum;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int b2 = 0, t2 = 0;
                for (int i = 0; i < BUCKET_WARPS; ++i) {
                    b2 += red[i];
                    t2 += red[BUCKET_WARPS + i];
                }
                before[e] = b2;
                total[e] = t2;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // exclusive scan over members (E <= 256) -> start of every member's range
        int run = 0;
        for (int e = 0; e < E; ++e) {
            const int t = total[e];
            total[e] = run;
            if (blockIdx.x == 0) offsets[e] = run;
            run += t;
        }
        if (blockIdx.x == 0) offsets[E] = run;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        int run = total[e] + before[e];
        for (int w = 0; w < BUCKET_WARPS; ++w) {
            cur[w * E + e] = run;
            if (w_first + w < W) run += counts[(size_t)(w_first + w) * E + e];
        }
    }
    __syncthreads();
    const int gw = w_first + warp;
    int n0, n1;
    bucket_warp_range(N, W, gw, n0, n1);
    int* cw = cur + warp * E;
    for (int nb = n0; nb < n1; nb += 32) {
        const int n = nb + lane;
        const int e = n < n1 ? (int)idx[n] : 0x7fffffff;
        const unsigned grp = __match_any_sync(0xffffffffu, e);
        const int leader = __ffs(grp) - 1;
        int start = 0;
        if (e < E && lane == leader) {
            start = cw[e];
            cw[e] = start + __popc(grp);
        }
        start = __shfl_sync(0xffffffffu, start, leader);
        if (e < E) perm[start + __popc(grp & ((1u << lane) - 1u))] = n;
        __syncwarp();
    }
}

__global__ void bump_u64_kernel(unsigned long long* c, unsigned long long delta) { *c += delta; }

__global__ void gather_inputs_kernel(const float* __restrict__ obs, const float* __restrict__ act,
                                     const int* __restrict__ perm, int N, int O, int A, float* __restrict__ x) {
    const int in = O + A;
    const long long total = (long long)N * in;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % in);
        long long r = i / in;
        long long env = perm ? perm[r] : r;
        x[i] = c < O ? obs[env * O + c] : act[env * A + (c - O)];
    }
}

__global__ void sample_scatter_kernel(const float* __restrict__ raw, int layout, const int* __restrict__ perm, int N,
                                      int D, const float* __restrict__ obs, int obs_ld, const float* __restrict__ mn,
                                      const float* __restrict__ mx, int bounds_n, const float* __restrict__ noise,
                                      float* __restrict__ pred) {
    const long long total = (long long)N * D;
    const bool gaussian = mn != nullptr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int d = (int)(i % D);
        long long r = i / D;
        long long env = perm ? perm[r] : r;
        float m, lvraw = 0.0f;
        if (layout == CMRL_HEAD_PLAIN) {
            const float* row = raw + r * (gaussian ? 2 * D : D);
            m = row[d];
            if (gaussian) lvraw = row[D + d];
        } else {
            const float* row = raw + ((long long)d * N + r) * (gaussian ? 2 : 1);
            m = row[0];
            if (gaussian) lvraw = row[1];
        }
        if (obs) m += obs[env * obs_ld + d];
        float out = m;
        if (gaussian && noise) {
            int bi = bounds_n == 1 ? 0 : d;
            float lv = soft_clamp_logvar(lvraw, mn[bi], mx[bi]);
            out = m + sqrtf(expf(lv)) * noise[env * D + d];
        }
        pred[env * D + d] = out;
    }
}

__global__ void select_sample_kernel(const float* __restrict__ mean, const float* __restrict__ logvar,
                                     const uint8_t* __restrict__ idx, int E, int N, int D,
                                     const float* __restrict__ noise, float* __restrict__ pred) {
    const long long total = (long long)N * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long n = i / D;
        long long src = (long long)idx[n] * N * D + i;
        float out = mean[src];
        if (noise && logvar) out += sqrtf(expf(logvar[src])) * noise[i];
        pred[i] = out;
    }
}

// one thread per env (O is small): mean over members, then max member distance
__global__ void mopo_penalty_kernel(const float* __restrict__ mean, int E, int N, int O, float* __restrict__ penalty) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float best = 0.0f;
    // two passes over [E,O] per env; avg recomputed per member to keep registers O-independent
    for (int e = 0; e < E; ++e) {
        float d2 = 0.0f;
        for (int o = 0; o < O; ++o) {
            float avg = 0.0f;
            for (int k = 0; k < E; ++k) avg += mean[((long long)k * N + n) * O + o];
            avg /= (float)E;
            float diff = mean[((long long)e * N + n) * O + o] - avg;
            d2 = fmaf(diff, diff, d2);
        }
        best = fmaxf(best, sqrtf(d2));
    }
    penalty[n] = best;
}

// A block takes TAIL_ENVS consecutive envs.  Phase 1, one thread per env: the episode bookkeeping and the per-env outputs
// (4-byte and 1-byte arrays, coalesced), and where the env's next observation comes from (itself, or a reset row) into shared
// memory.  Phase 2, all threads over the block's FLAT range of TAIL_ENVS x O floats: next_obs is read once with 16-byte
// accesses and written to terminal_obs and (patched at reset envs) obs_out the same way -- the first version walked each
// env's O floats with one thread (stride-O scalar accesses: 5 wavefronts per instruction at O = 5, 0.47 of the copy bandwidth).
constexpr int TAIL_ENVS = 256;
__global__ void __launch_bounds__(TAIL_ENVS) env_step_tail_kernel(
    const float* __restrict__ next_obs, const float* __restrict__ reward, const float* __restrict__ terminal,
    const uint8_t* __restrict__ terminal_u8, const float* __restrict__ penalty, float penalty_coeff, int* __restrict__ env_len,
    int max_steps, const float* __restrict__ reset_obs, const int* __restrict__ reset_index, int N, int O,
    float* __restrict__ obs_out, float* __restrict__ reward_out, uint8_t* __restrict__ done_out, uint8_t* __restrict__ trunc_out,
    float* __restrict__ terminal_obs, int* __restrict__ n_done, int vec) {
    __shared__ int s_src[TAIL_ENVS];  // -1: the env keeps its next observation; otherwise the row of reset_obs it restarts from
    const int n0 = blockIdx.x * TAIL_ENVS;
    const int n = n0 + threadIdx.x;
    bool done = false;
    if (n < N) {
        const int len = env_len[n] + 1;
        const bool trunc = len >= max_steps;
        const bool term = (terminal && terminal[n] != 0.0f) || (terminal_u8 && terminal_u8[n] != 0);
        done = term || trunc;
        float r = reward[n];
        if (penalty) r -= penalty[n] * penalty_coeff;
        reward_out[n] = r;
        done_out[n] = done ? 1 : 0;
        trunc_out[n] = trunc ? 1 : 0;
        env_len[n] = done ? 0 : len;
        s_src[threadIdx.x] = (done && reset_obs) ? (reset_index ? reset_index[n] : n) : -1;
    }
    const int n_done_blk = __syncthreads_count(done);
    if (n_done && threadIdx.x == 0 && n_done_blk) atomicAdd(n_done, n_done_blk);
    const int cnt = min(TAIL_ENVS, N - n0);
    const int total = cnt * O;                      // floats of this block's flat range
    const long long base = (long long)n0 * O;       // a multiple of 4 floats from the arrays' start (TAIL_ENVS = 256)
    const int nv = vec ? total / 4 : 0;
    for (int i = threadIdx.x; i < nv; i += TAIL_ENVS) {
        float4 v = *reinterpret_cast<const float4*>(next_obs + base + 4 * i);
        if (terminal_obs) *reinterpret_cast<float4*>(terminal_obs + base + 4 * i) = v;
        int e = (4 * i) / O, o = 4 * i - e * O;
        float c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int src = s_src[e];
            if (src >= 0) c[j] = reset_obs[(long long)src * O + o];
            if (++o == O) {
                o = 0;
                ++e;
            }
        }
        *reinterpret_cast<float4*>(obs_out + base + 4 * i) = make_float4(c[0], c[1], c[2], c[3]);
    }
    for (int i = 4 * nv + threadIdx.x; i < total; i += TAIL_ENVS) {
        const int e = i / O, o = i - e * O;
        const float v = next_obs[base + i];
        if (terminal_obs) terminal_obs[base + i] = v;
        const int src = s_src[e];
        obs_out[base + i] = src >= 0 ? reset_obs[(long long)src * O + o] : v;
    }
}

}  // namespace cmrl

using namespace cmrl;

extern "C" int cmrl_draw_member_noise(unsigned long long seed, unsigned long long stream_id,
                                      const unsigned long long* stream_base_dev, int N, const uint8_t* elite, int n_elite,
                                      int D, uint8_t* member_idx, float* noise, void* stream) {
    CMRL_REQUIRE(N >= 0 && D >= 0, "draw_member_noise: bad dims");
    CMRL_REQUIRE(!member_idx || (elite && n_elite >= 1), "draw_member_noise: elite list required");
    if (N == 0) return 0;
    draw_member_noise_kernel<<<blocks_for(N), RT, 0, (cudaStream_t)stream>>>(seed, stream_id, stream_base_dev, N, elite,
                                                                           n_elite, D, member_idx, noise);
    CMRL_CHECK_LAUNCH("draw_member_noise");
    return 0;
}

extern "C" int cmrl_draw_uniform_index(unsigned long long seed, unsigned long long stream_id,
                                       const unsigned long long* stream_base_dev, int N, int upper, int* out,
                                       void* stream) {
    CMRL_REQUIRE(out && upper >= 1 && N >= 0, "draw_uniform_index: bad arguments");
    if (N == 0) return 0;
    draw_uniform_index_kernel<<<blocks_for(N), RT, 0, (cudaStream_t)stream>>>(seed, stream_id, stream_base_dev, N,
                                                                            (unsigned)upper, out);
    CMRL_CHECK_LAUNCH("draw_uniform_index");
    return 0;
}

extern "C" int cmrl_counter_add(unsigned long long* counter_dev, unsigned long long delta, void* stream) {
    CMRL_REQUIRE(counter_dev, "counter_add: null pointer");
    bump_u64_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, delta);
    CMRL_CHECK_LAUNCH("counter_add");
    return 0;
}

extern "C" int cmrl_member_bucket(const uint8_t* member_idx, int N, int E, int* perm, int* offsets, int* counters,
                                  void* stream) {
    CMRL_REQUIRE(member_idx && perm && offsets && counters, "member_bucket: null pointer");
    CMRL_REQUIRE(E >= 1 && E <= MAX_E, "member_bucket: E must be in [1,%d]", MAX_E);
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        CMRL_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int) * (E + 1), st));
        return 0;
    }
    int W = (N + 63) / 64;  // >= 64 envs per warp chunk
    if (W > BUCKET_MAX_W) W = BUCKET_MAX_W;
    W = (W + BUCKET_WARPS - 1) / BUCKET_WARPS * BUCKET_WARPS;
    const int nb = W / BUCKET_WARPS;
    bucket_count_kernel<<<nb, BUCKET_WARPS * 32, sizeof(int) * BUCKET_WARPS * E, st>>>(member_idx, N, E, W, counters);
    CMRL_CHECK_LAUNCH("member_bucket.count");
    bucket_place_kernel<<<nb, BUCKET_WARPS * 32, sizeof(int) * (2 * E + BUCKET_WARPS * E + 2 * BUCKET_WARPS), st>>>(
        member_idx, N, E, W, counters, perm, offsets);
    CMRL_CHECK_LAUNCH("member_bucket.place");
    return 0;
}

extern "C" int cmrl_gather_inputs(const float* obs, const float* act, const int* perm, int N, int O, int A,
                                  float* x_sorted, void* stream) {
    CMRL_REQUIRE(obs && x_sorted && (act || A == 0), "gather_inputs: null pointer");
    long long total = (long long)N * (O + A);
    if (total == 0) return 0;
    int nb = blocks_for(total);
    if (nb > 16 * num_sms()) nb = 16 * num_sms();
    gather_inputs_kernel<<<nb, RT, 0, (cudaStream_t)stream>>>(obs, act, perm, N, O, A, x_sorted);
    CMRL_CHECK_LAUNCH("gather_inputs");
    return 0;
}

extern "C" int cmrl_sample_scatter(const float* raw_sorted, int layout, const int* perm, int N, int D, const float* obs,
                                   int obs_ld, const float* min_logvar, const float* max_logvar, int bounds_n,
                                   const float* noise, float* pred, void* stream) {
    CMRL_REQUIRE(raw_sorted && pred, "sample_scatter: null pointer");
    CMRL_REQUIRE(layout == CMRL_HEAD_PLAIN || layout == CMRL_HEAD_PARALLEL, "sample_scatter: bad layout");
    CMRL_REQUIRE((min_logvar == nullptr) == (max_logvar == nullptr), "sample_scatter: need both bounds or none");
    CMRL_REQUIRE(!min_logvar || bounds_n == 1 || bounds_n == D, "sample_scatter: bounds must have 1 or D entries");
    long long total = (long long)N * D;
    if (total == 0) return 0;
    int nb = blocks_for(total);
    if (nb > 16 * num_sms()) nb = 16 * num_sms();
    sample_scatter_kernel<<<nb, RT, 0, (cudaStream_t)stream>>>(raw_sorted, layout, perm, N, D, obs, obs_ld, min_logvar,
                                                             max_logvar, bounds_n, noise, pred);
    CMRL_CHECK_LAUNCH("sample_scatter");
    return 0;
}

extern "C" int cmrl_select_sample(const float* mean, const float* logvar, const uint8_t* member_idx, int E, int N,
                                  int D, const float* noise, float* pred, void* stream) {
    CMRL_REQUIRE(mean && member_idx && pred, "select_sample: null pointer");
    long long total = (long long)N * D;
    if (total == 0) return 0;
    int nb = blocks_for(total);
    if (nb > 16 * num_sms()) nb = 16 * num_sms();
    select_sample_kernel<<<nb, RT, 0, (cudaStream_t)stream>>>(mean, logvar, member_idx, E, N, D, noise, pred);
    CMRL_CHECK_LAUNCH("select_sample");
    return 0;
}

extern "C" int cmrl_mopo_penalty(const float* mean, int E, int N, int O, float* penalty, void* stream) {
    CMRL_REQUIRE(mean && penalty, "mopo_penalty: null pointer");
    if (N == 0) return 0;
    mopo_penalty_kernel<<<blocks_for(N), RT, 0, (cudaStream_t)stream>>>(mean, E, N, O, penalty);
    CMRL_CHECK_LAUNCH("mopo_penalty");
    return 0;
}

extern "C" int cmrl_env_step_tail(const float* next_obs, const float* reward, const float* terminal,
                                  const uint8_t* terminal_u8, const float* penalty, float penalty_coeff, int* env_len,
                                  int max_episode_steps, const float* reset_obs, const int* reset_index, int N, int O,
                                  float* obs_out, float* reward_out, uint8_t* done, uint8_t* truncated,
                                  float* terminal_obs, int* n_done, void* stream) {
    CMRL_REQUIRE(next_obs && reward && env_len && obs_out && reward_out && done && truncated,
                 "env_step_tail: null pointer");
    if (N == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int vec = ((((uintptr_t)next_obs | (uintptr_t)obs_out | (uintptr_t)terminal_obs) & 15) == 0) ? 1 : 0;
    env_step_tail_kernel<<<blocks_for(N, TAIL_ENVS), TAIL_ENVS, 0, st>>>(next_obs, reward, terminal, terminal_u8, penalty, penalty_coeff, env_len,
                                                                          max_episode_steps, reset_obs, reset_index, N, O, obs_out, reward_out,
                                                                          done, truncated, terminal_obs, n_done, vec);
    CMRL_CHECK_LAUNCH("env_step_tail");
    return 0;
}
