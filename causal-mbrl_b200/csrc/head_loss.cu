// K2 (Gaussian head), K3 (NLL / MSE losses, per-member hold-out MSE sums) and the fused
// L2-Adam step.  All HBM-bound elementwise / warp-reduction kernels: one thread per output
// element, grid-stride, coalesced on the innermost dim.
#include "common.cuh"

namespace cmrl {

constexpr int EW_THREADS = 256;

static inline int ew_blocks(long long n, int per_thread = 1) {
    long long b = (n + (long long)EW_THREADS * per_thread - 1) / ((long long)EW_THREADS * per_thread);
    long long cap = (long long)num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// raw element offsets of (mean, logvar) for output element (e,b,d)
__device__ __forceinline__ void head_raw_index(int layout, bool gaussian, int E, int B, int D, int e, int b, int d,
                                               long long* i_mean, long long* i_lv) {
    if (layout == CMRL_HEAD_PLAIN) {
        long long row = ((long long)e * B + b) * (gaussian ? 2 * D : D);
        *i_mean = row + d;
        *i_lv = row + D + d;
    } else {
        long long row = (((long long)d * E + e) * B + b) * (gaussian ? 2 : 1);
        *i_mean = row;
        *i_lv = row + 1;
    }
}

__global__ void gaussian_head_fwd_kernel(const float* __restrict__ raw, int layout, int E, int B, int D,
                                         const float* __restrict__ obs, long long obs_se, int obs_ld,
                                         const float* __restrict__ mn, const float* __restrict__ mx, int bounds_n,
                                         float* __restrict__ mean, float* __restrict__ logvar) {
    const long long total = (long long)E * B * D;
    const bool gaussian = logvar != nullptr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int d = (int)(i % D);
        long long eb = i / D;
        int b = (int)(eb % B), e = (int)(eb / B);
        long long im, il;
        head_raw_index(layout, gaussian, E, B, D, e, b, d, &im, &il);
        float m = raw[im];
        if (obs) m += obs[e * obs_se + (long long)b * obs_ld + d];
        mean[i] = m;
        if (gaussian) {
            int bi = bounds_n == 1 ? 0 : d;
            logvar[i] = soft_clamp_logvar(raw[il], mn[bi], mx[bi]);
        }
    }
}

__global__ void gaussian_head_bwd_kernel(const float* __restrict__ raw, int layout, int E, int B, int D,
                                         const float* __restrict__ mn, const float* __restrict__ mx, int bounds_n,
                                         const float* __restrict__ g_mean, const float* __restrict__ g_logvar,
                                         float* __restrict__ d_raw, float* d_min, float* d_max) {
    const long long total = (long long)E * B * D;
    const bool gaussian = g_logvar != nullptr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int d = (int)(i % D);
        long long eb = i / D;
        int b = (int)(eb % B), e = (int)(eb / B);
        long long im, il;
        head_raw_index(layout, gaussian, E, B, D, e, b, d, &im, &il);
        d_raw[im] = g_mean ? g_mean[i] : 0.0f;
        if (gaussian) {
            int bi = bounds_n == 1 ? 0 : d;
            float r = raw[il], lo = mn[bi], hi = mx[bi];
            float lv1 = hi - softplusf_(hi - r);
            // d lv / d lv1 = sigmoid(lv1 - min) ; d lv1 / d raw = sigmoid(max - raw)
            // (softplus' = sigmoid; above the threshold torch's softplus is the identity -> derivative 1)
            float s_lo = (lv1 - lo) > 20.0f ? 1.0f : sigmoidf_(lv1 - lo);
            float s_hi = (hi - r) > 20.0f ? 1.0f : sigmoidf_(hi - r);
            float g = g_logvar[i];
            d_raw[il] = g * s_lo * s_hi;
            if (d_max) atomicAdd(&d_max[bi], g * s_lo * (1.0f - s_hi));
            if (d_min) atomicAdd(&d_min[bi], g * (1.0f - s_lo));
        }
    }
}

__global__ void nll_loss_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ logvar,
                                    const float* __restrict__ target, long long t_se, int E, long long n, float reg,
                                    float* __restrict__ loss) {
    const long long total = (long long)E * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long e = i / n, r = i - e * n;
        float diff = mean[i] - target[e * t_se + r];
        float lv = logvar[i];
        loss[i] = diff * diff * expf(-lv) + lv + reg;
    }
}

__global__ void nll_loss_bwd_kernel(const float* __restrict__ mean, const float* __restrict__ logvar,
                                    const float* __restrict__ target, long long t_se, int E, long long n,
                                    const float* __restrict__ g_loss, float g_scalar, float* __restrict__ g_mean,
                                    float* __restrict__ g_logvar) {
    const long long total = (long long)E * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long e = i / n, r = i - e * n;
        float g = g_loss ? g_loss[i] : g_scalar;
        float diff = mean[i] - target[e * t_se + r];
        float iv = expf(-logvar[i]);
        g_mean[i] = g * 2.0f * diff * iv;
        g_logvar[i] = g * (1.0f - diff * diff * iv);
    }
}

__global__ void mse_loss_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ target, long long t_se,
                                    int E, long long n, float* __restrict__ loss) {
    const long long total = (long long)E * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long e = i / n, r = i - e * n;
        float diff = mean[i] - target[e * t_se + r];
        loss[i] = diff * diff;
    }
}

// gradient of the un-reduced squared error: g_mean = 2 (mean - target) g_loss
__global__ void mse_loss_bwd_kernel(const float* __restrict__ mean, const float* __restrict__ target, long long t_se, int E,
                                    long long n, const float* __restrict__ g_loss, float* __restrict__ g_mean) {
    const long long total = (long long)E * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long e = i / n, r = i - e * n;
        g_mean[i] = 2.0f * (mean[i] - target[e * t_se + r]) * g_loss[i];
    }
}

// grid = (blocks, E): per-member sum of squared errors, fp32 per thread -> fp64 warp/block -> one atomic per block
__global__ void mse_member_sums_kernel(const float* __restrict__ mean, const float* __restrict__ target, long long t_se,
                                       long long n, double* __restrict__ sums) {
    const int e = blockIdx.y;
    const float* m = mean + (long long)e * n;
    const float* t = target + (long long)e * t_se;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float diff = m[i] - t[i];
        acc += (double)(diff * diff);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double warp_acc[EW_THREADS / 32];
    if ((threadIdx.x & 31) == 0) warp_acc[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < EW_THREADS / 32; ++w) s += warp_acc[w];
        atomicAdd(&sums[e], s);
    }
}

// Adam with the step counter in device memory (CUDA-graph replayable): t = step_dev[0] + 1; the last block to finish
// advances the counter (step_dev[1] is its arrival ticket).  16-byte accesses when the four arrays are aligned.
__device__ __forceinline__ float adam_l2_one(float pi, float g, float& mi, float& vi, float lr_over_bc1, float sqrt_bc2, float beta1,
                                             float beta2, float eps, float wd, float gscale) {
    const float gi = fmaf(wd, pi, g * gscale);
    mi = fmaf(1.0f - beta1, gi - mi, mi);
    vi = fmaf((1.0f - beta2) * gi, gi, beta2 * vi);
    return pi - lr_over_bc1 * (mi / (sqrtf(vi) / sqrt_bc2 + eps));
}
__device__ __forceinline__ void adam_l2_range(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, long long n, int vec, float lr_over_bc1, float sqrt_bc2,
                                              float beta1, float beta2, float eps, float wd, float gscale) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec ? n / 4 : 0;
    for (long long i = tid; i < n4; i += nth) {
        const float4 pi = reinterpret_cast<const float4*>(p)[i], gi = reinterpret_cast<const float4*>(g)[i];
        float4 mi = reinterpret_cast<const float4*>(m)[i], vi = reinterpret_cast<const float4*>(v)[i], o;
        o.x = adam_l2_one(pi.x, gi.x, mi.x, vi.x, lr_over_bc1, sqrt_bc2, beta1, beta2, eps, wd, gscale);
        o.y = adam_l2_one(pi.y, gi.y, mi.y, vi.y, lr_over_bc1, sqrt_bc2, beta1, beta2, eps, wd, gscale);
        o.z = adam_l2_one(pi.z, gi.z, mi.z, vi.z, lr_over_bc1, sqrt_bc2, beta1, beta2, eps, wd, gscale);
        o.w = adam_l2_one(pi.w, gi.w, mi.w, vi.w, lr_over_bc1, sqrt_bc2, beta1, beta2, eps, wd, gscale);
        reinterpret_cast<float4*>(m)[i] = mi;
        reinterpret_cast<float4*>(v)[i] = vi;
        reinterpret_cast<float4*>(p)[i] = o;
    }
    for (long long i = 4 * n4 + tid; i < n; i += nth) {
        float mi = m[i], vi = v[i];
        p[i] = adam_l2_one(p[i], g[i], mi, vi, lr_over_bc1, sqrt_bc2, beta1, beta2, eps, wd, gscale);
        m[i] = mi;
        v[i] = vi;
    }
}
__global__ void adam_l2_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                               long long n, float lr_over_bc1, float beta1, float beta2, float eps, float wd, float gscale,
                               float sqrt_bc2, int vec) {
    adam_l2_range(p, g, m, v, n, vec, lr_over_bc1, sqrt_bc2, beta1, beta2, eps, wd, gscale);
}
__global__ void adam_l2_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                   float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                                   float wd, float gscale, int* step_dev, int vec) {
    __shared__ float s_lr_over_bc1, s_sqrt_bc2;
    if (threadIdx.x == 0) {
        const double t = (double)(step_dev[0] + 1);
        s_lr_over_bc1 = (float)((double)lr / (1.0 - pow((double)beta1, t)));
        s_sqrt_bc2 = (float)sqrt(1.0 - pow((double)beta2, t));
    }
    __syncthreads();
    adam_l2_range(p, g, m, v, n, vec, s_lr_over_bc1, s_sqrt_bc2, beta1, beta2, eps, wd, gscale);
    last_block_add(step_dev, 1);
}

// acc[0] += sum(x[0..n)) in fp64, deterministic: per-block partial sums (fixed strided assignment), the last block to finish
// adds them up in block order.  ws: [blocks] doubles then two ints (unused, arrival ticket), zero-initialised by the caller once.
constexpr int SUM_BLOCKS = 64, SUM_THREADS = 512;
__global__ void sum_acc_f64_kernel(const float* __restrict__ x, long long n, double* __restrict__ acc, double* __restrict__ ws) {
    __shared__ double sh[SUM_THREADS / 32];
    __shared__ bool last;
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += (double)x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    int* ticket = reinterpret_cast<int*>(ws + SUM_BLOCKS);
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < SUM_THREADS / 32; ++w) t += sh[w];
        ws[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(reinterpret_cast<unsigned*>(ticket + 1), 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            double tot = 0.0;
            for (unsigned b = 0; b < gridDim.x; ++b) tot += reinterpret_cast<volatile double*>(ws)[b];
            acc[0] += tot;
            ticket[1] = 0;
        }
    }
}

// Bootstrap batch assembly on the device (BootstrapIterator.__next__ + _consolidate_batches,
// cmrl/util/transition_iterator.py:14-28,131-139): row b of member e = transitions[member_idx[e, order[start + b]]]
__global__ void bootstrap_gather_kernel(const float* __restrict__ obs, const float* __restrict__ act,
                                        const float* __restrict__ target, const int* __restrict__ member_idx,
                                        const int* __restrict__ order, int* start_dev, int advance, int B, int N,
                                        int O, int A, int D, float* __restrict__ o_obs, float* __restrict__ o_act,
                                        float* __restrict__ o_target) {
    const int W = O + A + D;
    const int e = blockIdx.y;
    const int start = *start_dev;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * W; i += gridDim.x * blockDim.x) {
        const int b = i / W, c = i - b * W;
        const int pos = min(start + b, N - 1);
        const long long src = member_idx[(long long)e * N + order[pos]];
        const long long row = (long long)e * B + b;
        if (c < O) o_obs[row * O + c] = obs[src * O + c];
        else if (c < O + A) o_act[row * A + (c - O)] = act[src * A + (c - O)];
        else o_target[row * D + (c - O - A)] = target[src * D + (c - O - A)];
    }
    if (advance) last_block_add(start_dev, advance);
}

}  // namespace cmrl

using namespace cmrl;

extern "C" int cmrl_adam_l2_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                     float lr, float beta1, float beta2, float eps, float weight_decay,
                                     float grad_scale, int* step_dev, void* stream) {
    CMRL_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_dev, "adam_l2_step_dev: null pointer");
    if (n == 0) return 0;
    const int vec = ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0) ? 1 : 0;
    adam_l2_dev_kernel<<<ew_blocks(n, 4), EW_THREADS, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                               beta2, eps, weight_decay, grad_scale, step_dev, vec);
    CMRL_CHECK_LAUNCH("adam_l2_step_dev");
    return 0;
}

extern "C" int cmrl_sum_acc_f64(const float* x, long long n, double* acc, double* workspace, void* stream) {
    CMRL_REQUIRE(x && acc && workspace && n >= 0, "sum_acc_f64: null pointer");
    if (n == 0) return 0;
    long long blocks = (n + SUM_THREADS * 8 - 1) / (SUM_THREADS * 8);
    if (blocks > SUM_BLOCKS) blocks = SUM_BLOCKS;
    sum_acc_f64_kernel<<<(int)blocks, SUM_THREADS, 0, (cudaStream_t)stream>>>(x, n, acc, workspace);
    CMRL_CHECK_LAUNCH("sum_acc_f64");
    return 0;
}

extern "C" int cmrl_bootstrap_gather(const float* obs, const float* act, const float* target, const int* member_idx,
                                     const int* order, int* start_dev, int advance, int E, int B, int N, int O, int A,
                                     int D, float* out_obs, float* out_act, float* out_target, void* stream) {
    CMRL_REQUIRE(obs && act && target && member_idx && order && start_dev && out_obs && out_act && out_target,
                 "bootstrap_gather: null pointer");
    CMRL_REQUIRE(E >= 1 && B >= 1 && N >= 1, "bootstrap_gather: bad dims");
    int bx = (B * (O + A + D) + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid(bx, E);
    bootstrap_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(obs, act, target, member_idx, order, start_dev, advance, B, N,
                                                                  O, A, D, out_obs, out_act, out_target);
    CMRL_CHECK_LAUNCH("bootstrap_gather");
    return 0;
}

extern "C" int cmrl_gaussian_head_fwd(const float* raw, int layout, int E, int B, int D, const float* obs,
                                      long long obs_se, int obs_ld, const float* min_logvar, const float* max_logvar,
                                      int bounds_n, float* mean, float* logvar, void* stream) {
    CMRL_REQUIRE(raw && mean, "gaussian_head_fwd: null pointer");
    CMRL_REQUIRE(layout == CMRL_HEAD_PLAIN || layout == CMRL_HEAD_PARALLEL, "gaussian_head_fwd: bad layout");
    CMRL_REQUIRE(E >= 1 && B >= 0 && D >= 1, "gaussian_head_fwd: bad dims");
    CMRL_REQUIRE(!logvar || (min_logvar && max_logvar && (bounds_n == 1 || bounds_n == D)),
                 "gaussian_head_fwd: logvar bounds must have 1 or D entries");
    CMRL_REQUIRE(!obs || obs_ld >= D, "gaussian_head_fwd: residual needs obs_ld >= D");
    long long total = (long long)E * B * D;
    if (total == 0) return 0;
    gaussian_head_fwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
        raw, layout, E, B, D, obs, obs_se, obs_ld, min_logvar, max_logvar, bounds_n, mean, logvar);
    CMRL_CHECK_LAUNCH("gaussian_head_fwd");
    return 0;
}

extern "C" int cmrl_gaussian_head_bwd(const float* raw, int layout, int E, int B, int D, const float* min_logvar,
                                      const float* max_logvar, int bounds_n, const float* g_mean,
                                      const float* g_logvar, float* d_raw, float* d_min, float* d_max, void* stream) {
    CMRL_REQUIRE(raw && d_raw, "gaussian_head_bwd: null pointer");
    CMRL_REQUIRE(layout == CMRL_HEAD_PLAIN || layout == CMRL_HEAD_PARALLEL, "gaussian_head_bwd: bad layout");
    CMRL_REQUIRE(!g_logvar || (min_logvar && max_logvar && (bounds_n == 1 || bounds_n == D)),
                 "gaussian_head_bwd: logvar bounds must have 1 or D entries");
    long long total = (long long)E * B * D;
    if (total == 0) return 0;
    gaussian_head_bwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
        raw, layout, E, B, D, min_logvar, max_logvar, bounds_n, g_mean, g_logvar, d_raw, d_min, d_max);
    CMRL_CHECK_LAUNCH("gaussian_head_bwd");
    return 0;
}

extern "C" int cmrl_nll_loss_fwd(const float* mean, const float* logvar, const float* target, long long t_se, int E,
                                 long long n_per_member, float reg, float* loss, void* stream) {
    CMRL_REQUIRE(mean && logvar && target && loss, "nll_loss_fwd: null pointer");
    long long total = (long long)E * n_per_member;
    if (total == 0) return 0;
    nll_loss_fwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(mean, logvar, target, t_se, E,
                                                                                 n_per_member, reg, loss);
    CMRL_CHECK_LAUNCH("nll_loss_fwd");
    return 0;
}

extern "C" int cmrl_nll_loss_bwd(const float* mean, const float* logvar, const float* target, long long t_se, int E,
                                 long long n_per_member, const float* g_loss, float g_scalar, float* g_mean,
                                 float* g_logvar, void* stream) {
    CMRL_REQUIRE(mean && logvar && target && g_mean && g_logvar, "nll_loss_bwd: null pointer");
    long long total = (long long)E * n_per_member;
    if (total == 0) return 0;
    nll_loss_bwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
        mean, logvar, target, t_se, E, n_per_member, g_loss, g_scalar, g_mean, g_logvar);
    CMRL_CHECK_LAUNCH("nll_loss_bwd");
    return 0;
}

extern "C" int cmrl_mse_loss_fwd(const float* mean, const float* target, long long t_se, int E, long long n_per_member,
                                 float* loss, void* stream) {
    CMRL_REQUIRE(mean && target && loss, "mse_loss_fwd: null pointer");
    long long total = (long long)E * n_per_member;
    if (total == 0) return 0;
    mse_loss_fwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(mean, target, t_se, E, n_per_member,
                                                                                 loss);
    CMRL_CHECK_LAUNCH("mse_loss_fwd");
    return 0;
}

extern "C" int cmrl_mse_loss_bwd(const float* mean, const float* target, long long t_se, int E, long long n_per_member,
                                 const float* g_loss, float* g_mean, void* stream) {
    CMRL_REQUIRE(mean && target && g_loss && g_mean, "mse_loss_bwd: null pointer");
    long long total = (long long)E * n_per_member;
    if (total == 0) return 0;
    mse_loss_bwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(mean, target, t_se, E, n_per_member, g_loss,
                                                                                 g_mean);
    CMRL_CHECK_LAUNCH("mse_loss_bwd");
    return 0;
}

extern "C" int cmrl_mse_member_sums(const float* mean, const float* target, long long t_se, int E,
                                    long long n_per_member, double* sums, void* stream) {
    CMRL_REQUIRE(mean && target && sums, "mse_member_sums: null pointer");
    CMRL_REQUIRE(E >= 1 && E <= 65535, "mse_member_sums: bad E");
    if (n_per_member == 0) return 0;
    int bx = ew_blocks(n_per_member, 4);
    int cap = (num_sms() * 8 + E - 1) / E;
    if (bx > cap) bx = cap;
    dim3 grid(bx, E);
    mse_member_sums_kernel<<<grid, EW_THREADS, 0, (cudaStream_t)stream>>>(mean, target, t_se, n_per_member, sums);
    CMRL_CHECK_LAUNCH("mse_member_sums");
    return 0;
}

extern "C" int cmrl_adam_l2_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                 float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                 int step, void* stream) {
    CMRL_REQUIRE(param && grad && exp_avg && exp_avg_sq, "adam_l2_step: null pointer");
    CMRL_REQUIRE(step >= 1, "adam_l2_step: step is 1-based");
    if (n == 0) return 0;
    double bc1 = 1.0 - pow((double)beta1, (double)step);
    double bc2 = 1.0 - pow((double)beta2, (double)step);
    const int vec = ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0) ? 1 : 0;
    adam_l2_kernel<<<ew_blocks(n, 4), EW_THREADS, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, (float)((double)lr / bc1), beta1, beta2, eps, weight_decay, grad_scale,
        (float)sqrt(bc2), vec);
    CMRL_CHECK_LAUNCH("adam_l2_step");
    return 0;
}
