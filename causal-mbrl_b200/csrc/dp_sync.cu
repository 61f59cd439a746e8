// Data-parallel gradient exchange over NVLink peer memory, fused with the L2-Adam step.
//
// Replaces `torch.distributed.all_reduce(flat_grad)` + `cmrl_adam_l2_step_dev` of a data-parallel training step
// (the reference trains on one device: base_dynamics.py:173-188; the exchange has no counterpart there) with ONE kernel
// per rank.  The 3.5 MB gradient arena of a 7 x (4 x 200) ensemble is far below the size at which link bandwidth
// matters (NVSwitch: every peer at full rate), so the exchange is built for latency: two store-only phases, no
// remote loads, no kernel boundary, no host round trip.
//
//   phase 1 (reduce-scatter by push): rank r stores its slice-q gradients into rank q's inbox[r]  (q != r)
//   phase 2: rank q sums inbox[0..W) + its own slice in rank order -- so every rank receives bit-identical sums --
//            and stores the reduced slice into every peer's `reduced` arena
//   phase 3: every rank applies Adam to all slices from its `reduced` arena (and mirrors the sums into `grad`)
//
// Every CTA b owns the same sub-range of every slice on every rank, so all waits are CTA b of rank r on CTA b of
// the other ranks (one flag word per (source rank, CTA), monotonically increasing epochs kept in device memory: the
// kernel is CUDA-graph replayable).  A wait can only be on a kernel of another GPU that is running or about to start.
// Waits give up after ~20 s and raise the communicator's error word instead of hanging the device.
#include <string.h>

#include "common.cuh"

namespace cmrl {
namespace dp {

constexpr int MAX_WORLD = 16;
constexpr int DP_THREADS = 512;
constexpr int DP_MAX_CTAS = 148;

struct Layout {
    long long slice_cap;  // floats per slice the buffers were sized for
    size_t inbox_off, reduced_off, flag1_off, flag2_off, epoch_off, err_off, total;
};

struct Comm {
    int rank, world, ctas, device;
    long long cap;
    Layout lay;
    uint8_t* local;
    uint8_t* peer[MAX_WORLD];
    bool connected;
};

static Layout make_layout(long long cap, int world) {
    Layout l{};
    l.slice_cap = ((cap + world - 1) / world + 3) / 4 * 4;
    size_t off = 0;
    l.inbox_off = off;
    off += (size_t)world * l.slice_cap * 4;
    l.reduced_off = off;
    off += (size_t)world * l.slice_cap * 4;
    l.flag1_off = off;
    off += (size_t)MAX_WORLD * DP_MAX_CTAS * 4;
    l.flag2_off = off;
    off += (size_t)MAX_WORLD * DP_MAX_CTAS * 4;
    l.epoch_off = off;
    off += (size_t)DP_MAX_CTAS * 4;
    l.err_off = off;
    off += 16;
    l.total = (off + 255) / 256 * 256;
    return l;
}

struct KParams {
    int rank, world;
    long long n, S, T;  // floats in the arena, floats per slice, floats per CTA and slice (multiples of 4)
    Layout lay;
    uint8_t* peer[MAX_WORLD];
    float *param, *grad, *m, *v;  // param == nullptr: exchange only
    float lr, beta1, beta2, eps, wd, gscale;
    int* step_dev;  // [0] steps taken, [1] arrival ticket of the last-block update
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// threads 0..W-1 (except `rank`) each wait for one source rank's flag; false after ~20 s
__device__ __forceinline__ bool wait_flags(const uint32_t* flags, int W, int rank, int b, uint32_t epoch, uint32_t* err) {
    if ((int)threadIdx.x < W && (int)threadIdx.x != rank) {
        const uint32_t* f = flags + (size_t)threadIdx.x * DP_MAX_CTAS + b;
        const unsigned long long t0 = globaltimer_ns();
        // epochs only grow; the signed difference tolerates wrap-around
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            __nanosleep(40);
            if (globaltimer_ns() - t0 > 20000000000ull) {
                atomicExch(err, 1u + (uint32_t)threadIdx.x);
                break;
            }
        }
    }
    __syncthreads();
    return *(volatile uint32_t*)err == 0;
}

struct AdamC {
    float lr_over_bc1, sqrt_bc2;
};
__device__ __forceinline__ float adam_one(const KParams& p, const AdamC& c, float pi, float g, float& mi, float& vi) {
    const float gi = fmaf(p.wd, pi, g * p.gscale);
    mi = fmaf(1.0f - p.beta1, gi - mi, mi);
    vi = fmaf((1.0f - p.beta2) * gi, gi, p.beta2 * vi);
    return pi - c.lr_over_bc1 * (mi / (sqrtf(vi) / c.sqrt_bc2 + p.eps));
}
// grad[gi..gi+4) <- a (the summed gradient) and, with ADAM, the update of those four parameters; `full`: all four exist
template <bool ADAM>
__device__ __forceinline__ void apply4(const KParams& p, const AdamC& c, long long gi, float4 a, bool full) {
    if (full) {
        *reinterpret_cast<float4*>(p.grad + gi) = a;
        if (ADAM) {
            const float4 w = *reinterpret_cast<const float4*>(p.param + gi);
            float4 m = *reinterpret_cast<const float4*>(p.m + gi), v = *reinterpret_cast<const float4*>(p.v + gi);
            float4 o;
            o.x = adam_one(p, c, w.x, a.x, m.x, v.x);
            o.y = adam_one(p, c, w.y, a.y, m.y, v.y);
            o.z = adam_one(p, c, w.z, a.z, m.z, v.z);
            o.w = adam_one(p, c, w.w, a.w, m.w, v.w);
            *reinterpret_cast<float4*>(p.m + gi) = m;
            *reinterpret_cast<float4*>(p.v + gi) = v;
            *reinterpret_cast<float4*>(p.param + gi) = o;
        }
    } else {  // the last, partial vector of an arena whose length is not a multiple of four
        const float av[4] = {a.x, a.y, a.z, a.w};
        for (int u = 0; u < 4; ++u) {
            if (gi + u >= p.n) break;
            p.grad[gi + u] = av[u];
            if (ADAM) {
                float m = p.m[gi + u], v = p.v[gi + u];
                p.param[gi + u] = adam_one(p, c, p.param[gi + u], av[u], m, v);
                p.m[gi + u] = m;
                p.v[gi + u] = v;
            }
        }
    }
}
__device__ __forceinline__ float4 load4_clipped(const float* src, long long gi, long long n) {
    if (gi + 3 < n) return *reinterpret_cast<const float4*>(src + gi);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gi < n) x.x = src[gi];
    if (gi + 1 < n) x.y = src[gi + 1];
    if (gi + 2 < n) x.z = src[gi + 2];
    return x;
}

// Work is counted in 16-byte vectors.  Tv = vectors per CTA and slice; loops over (peer slice, vector) pairs are flattened
// and run in batches of DP_BATCH independent loads, because every loop here is one or two memory round trips long.
constexpr int DP_BATCH = 4;

template <bool ADAM>
__global__ void __launch_bounds__(DP_THREADS, 1) allreduce_adam_kernel(const KParams p) {
    const int b = blockIdx.x, W = p.world, rank = p.rank, tid = threadIdx.x;
    uint8_t* local = p.peer[rank];
    uint32_t* ep = reinterpret_cast<uint32_t*>(local + p.lay.epoch_off) + b;
    uint32_t* err = reinterpret_cast<uint32_t*>(local + p.lay.err_off);
    const uint32_t epoch = *ep + 1;
    __shared__ AdamC s_c;
    if (ADAM && tid == 0) {
        const double t = (double)(p.step_dev[0] + 1);
        s_c.lr_over_bc1 = (float)((double)p.lr / (1.0 - pow((double)p.beta1, t)));
        s_c.sqrt_bc2 = (float)sqrt(1.0 - pow((double)p.beta2, t));
    }
    // this CTA's sub-range of every slice, in slice-local floats: [lo, hi)
    const long long lo = (long long)b * p.T, hi = min((long long)(b + 1) * p.T, p.S);
    const int Tv = hi > lo ? (int)((hi - lo) / 4) : 0;  // S and T are multiples of 4
    const int pairs = (W - 1) * Tv;

    // ---- phase 1: push my gradients of every other rank's slice into that rank's inbox[rank] -----------------------------
    for (int j0 = tid; j0 < pairs; j0 += DP_BATCH * DP_THREADS) {
        float4 x[DP_BATCH];
#pragma unroll
        for (int u = 0; u < DP_BATCH; ++u) {
            const int j = j0 + u * DP_THREADS;
            if (j < pairs) {
                const int q = (rank + 1 + j / Tv) % W;
                x[u] = load4_clipped(p.grad, (long long)q * p.S + lo + 4ll * (j % Tv), p.n);
            }
        }
#pragma unroll
        for (int u = 0; u < DP_BATCH; ++u) {
            const int j = j0 + u * DP_THREADS;
            if (j < pairs) {
                const int q = (rank + 1 + j / Tv) % W;
                float* inbox = reinterpret_cast<float*>(p.peer[q] + p.lay.inbox_off) + (size_t)rank * p.lay.slice_cap;
                *reinterpret_cast<float4*>(inbox + lo + 4ll * (j % Tv)) = x[u];  // slots past n carry zeros
            }
        }
    }
    __syncthreads();  // the release below is cumulative over the stores this barrier orders before it
    if (tid < W && tid != rank)
        st_release_sys(reinterpret_cast<uint32_t*>(p.peer[tid] + p.lay.flag1_off) + (size_t)rank * DP_MAX_CTAS + b, epoch);

    // ---- phase 2: reduce my slice in rank order, publish it to every peer (and update my slice) ---------------------------------
    bool ok = wait_flags(reinterpret_cast<const uint32_t*>(local + p.lay.flag1_off), W, rank, b, epoch, err);
    const AdamC c = s_c;
    if (ok) {
        const float* inbox0 = reinterpret_cast<const float*>(local + p.lay.inbox_off);
        const long long base = (long long)rank * p.S;
        for (int i = tid; i < Tv; i += DP_THREADS) {
            const long long sl = lo + 4ll * i, gi = base + sl;
            if (gi >= p.n) break;
            float4 x[MAX_WORLD];
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r) {
                if (r < W) x[r] = r == rank ? load4_clipped(p.grad, gi, p.n) : __ldcg(reinterpret_cast<const float4*>(inbox0 + (size_t)r * p.lay.slice_cap + sl));
            }
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < MAX_WORLD; ++r) {
                if (r < W) {
                    acc.x += x[r].x;
                    acc.y += x[r].y;
                    acc.z += x[r].z;
                    acc.w += x[r].w;
                }
            }
            for (int dq = 1; dq < W; ++dq) {
                const int q = (rank + dq) % W;
                float* red = reinterpret_cast<float*>(p.peer[q] + p.lay.reduced_off) + (size_t)rank * p.lay.slice_cap;
                *reinterpret_cast<float4*>(red + sl) = acc;
            }
            apply4<ADAM>(p, c, gi, acc, gi + 3 < p.n);
        }
    }
    __syncthreads();
    if (tid < W && tid != rank)
        st_release_sys(reinterpret_cast<uint32_t*>(p.peer[tid] + p.lay.flag2_off) + (size_t)rank * DP_MAX_CTAS + b, epoch);

    // ---- phase 3: the other ranks' reduced slices -> grad (+ Adam) ------------------------------------------------------------
    ok = wait_flags(reinterpret_cast<const uint32_t*>(local + p.lay.flag2_off), W, rank, b, epoch, err);
    if (ok) {
        const float* red0 = reinterpret_cast<const float*>(local + p.lay.reduced_off);
        for (int j0 = tid; j0 < pairs; j0 += 2 * DP_THREADS) {
            float4 x[2];
            long long gi[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = j0 + u * DP_THREADS;
                gi[u] = p.n;
                if (j < pairs) {
                    const int q = (rank + 1 + j / Tv) % W;
                    const long long sl = lo + 4ll * (j % Tv);
                    gi[u] = (long long)q * p.S + sl;
                    x[u] = __ldcg(reinterpret_cast<const float4*>(red0 + (size_t)q * p.lay.slice_cap + sl));
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (gi[u] < p.n) apply4<ADAM>(p, c, gi[u], x[u], gi[u] + 3 < p.n);
        }
    }
    if (tid == 0) *ep = epoch;
    if (ADAM) last_block_add(p.step_dev, 1);
}

static int launch(Comm* c, float* param, float* grad, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float wd, float gscale, int* step_dev, cudaStream_t st) {
    CMRL_REQUIRE(c && c->connected, "allreduce_grads: communicator not connected (cmrl_dp_comm_connect)");
    CMRL_REQUIRE(grad && n >= 0 && n <= c->cap, "allreduce_grads: %lld floats exceed the communicator's capacity %lld", n, c->cap);
    CMRL_REQUIRE(((uintptr_t)grad & 15) == 0, "allreduce_grads: the gradient arena must be 16-byte aligned");
    if (n == 0) return 0;
    KParams p{};
    p.rank = c->rank;
    p.world = c->world;
    p.n = n;
    p.S = ((n + c->world - 1) / c->world + 3) / 4 * 4;
    p.T = ((p.S + c->ctas - 1) / c->ctas + 3) / 4 * 4;
    p.lay = c->lay;
    for (int r = 0; r < c->world; ++r) p.peer[r] = c->peer[r];
    p.param = param;
    p.grad = grad;
    p.m = m;
    p.v = v;
    p.lr = lr;
    p.beta1 = beta1;
    p.beta2 = beta2;
    p.eps = eps;
    p.wd = wd;
    p.gscale = gscale;
    p.step_dev = step_dev;
    if (param) allreduce_adam_kernel<true><<<c->ctas, DP_THREADS, 0, st>>>(p);
    else allreduce_adam_kernel<false><<<c->ctas, DP_THREADS, 0, st>>>(p);
    CMRL_CHECK_LAUNCH("allreduce_grads");
    return 0;
}

}  // namespace dp
}  // namespace cmrl

using namespace cmrl;
using cmrl::dp::Comm;

extern "C" int cmrl_dp_comm_create(int rank, int world, long long max_floats, void** comm_out, unsigned char* ipc_handle_out) {
    CMRL_REQUIRE(comm_out && ipc_handle_out, "dp_comm_create: null pointer");
    CMRL_REQUIRE(world >= 1 && world <= dp::MAX_WORLD && rank >= 0 && rank < world && max_floats >= 1, "dp_comm_create: bad rank / world / size");
    Comm* c = new Comm();
    memset(c, 0, sizeof(Comm));
    c->rank = rank;
    c->world = world;
    c->cap = max_floats;
    c->lay = dp::make_layout(max_floats, world);
    // One CTA per SM of a B200, the same count on every rank.  A CTA waits only for the CTA of the same index on the other
    // ranks and CTAs are dispatched in index order, so a grid that is not fully resident (SMs busy with another stream's
    // kernel) still makes progress.
    c->ctas = dp::DP_MAX_CTAS;
    cudaGetDevice(&c->device);
    cudaError_t e = cudaMalloc((void**)&c->local, c->lay.total);
    if (e != cudaSuccess) {
        set_error("dp_comm_create: cudaMalloc(%zu): %s", c->lay.total, cudaGetErrorString(e));
        delete c;
        return (int)e;
    }
    cudaMemset(c->local, 0, c->lay.total);
    c->peer[rank] = c->local;
    cudaIpcMemHandle_t h;
    e = world > 1 ? cudaIpcGetMemHandle(&h, c->local) : cudaSuccess;
    if (e != cudaSuccess) {
        set_error("dp_comm_create: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        cudaFree(c->local);
        delete c;
        return (int)e;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (world > 1) memcpy(ipc_handle_out, &h, 64);
    else memset(ipc_handle_out, 0, 64);
    c->connected = world == 1;
    *comm_out = c;
    return 0;
}

extern "C" int cmrl_dp_comm_connect(void* comm, const unsigned char* all_handles) {
    Comm* c = static_cast<Comm*>(comm);
    CMRL_REQUIRE(c && all_handles, "dp_comm_connect: null pointer");
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || c->peer[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + 64 * r, 64);
        void* ptr = nullptr;
        CMRL_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[r] = static_cast<uint8_t*>(ptr);
    }
    c->connected = true;
    return 0;
}

extern "C" int cmrl_dp_comm_error(void* comm, int* err_out) {
    Comm* c = static_cast<Comm*>(comm);
    CMRL_REQUIRE(c && err_out, "dp_comm_error: null pointer");
    uint32_t e = 0;
    CMRL_CUDA(cudaMemcpy(&e, c->local + c->lay.err_off, 4, cudaMemcpyDeviceToHost));
    *err_out = (int)e;
    return 0;
}

extern "C" int cmrl_dp_comm_destroy(void* comm) {
    Comm* c = static_cast<Comm*>(comm);
    if (!c) return 0;
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->local);
    delete c;
    return 0;
}

extern "C" int cmrl_allreduce_grads(void* comm, float* grad, long long n, void* stream) {
    return dp::launch(static_cast<Comm*>(comm), nullptr, grad, nullptr, nullptr, n, 0, 0, 0, 0, 0, 1.0f, nullptr, (cudaStream_t)stream);
}

extern "C" int cmrl_allreduce_grads_adam(void* comm, float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                         float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                         int* step_dev, void* stream) {
    CMRL_REQUIRE(param && exp_avg && exp_avg_sq && step_dev, "allreduce_grads_adam: null pointer");
    return dp::launch(static_cast<Comm*>(comm), param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, grad_scale,
                      step_dev, (cudaStream_t)stream);
}
