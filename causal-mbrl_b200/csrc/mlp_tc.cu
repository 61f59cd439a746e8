// Tensor-core tier: the whole selected-member MLP of one mechanism (4 x (EnsembleLinear + act) + Gaussian
// head + sampling + scatter; cmrl/models/layers.py:60-65, plain_transition.py:100-122,
// external_mask_transition.py:141-169, fake_env.py:195-215) fused into ONE persistent kernel on tcgen05:
//
//   * operands fp16 (11-bit significand = TF32's; kind::f16), accumulation fp32 in TMEM;
//   * weights pre-packed (cmrl_pack_layer_f16) into the UMMA K-major no-swizzle core-matrix layout with the
//     bias folded in as an extra K row (the A operand carries a 1.0 column) and the causal input mask folded
//     into W1, so a layer's B operand is one contiguous blob staged by a single TMA bulk copy (UBLKCP);
//   * activations never leave the SM: TMEM accumulator -> registers (tcgen05.ld) -> SiLU/ReLU -> fp16 ->
//     shared memory in the A-operand layout of the next layer;
//   * the head epilogue applies the soft-clamped log-variance, the residual, the supplied N(0,1) noise and
//     scatters next-obs / reward / terminal to the env-ordered output.
//
// This file holds the weight packing and the entry point; the kernels live in mlp_tc5.cu (default: CTA pairs, two 256-row tiles
// in flight), mlp_tc3.cu (heads wider than the default kernel takes) and mlp_tc6.cu (hidden layers wider than shared memory).
#include "tc_common.cuh"

namespace cmrl {
namespace tc {

// ----------------------------------------------------------------------------------------------- packing
// One layer of every net: fp32 W [G,in,out] (+ bias [G,1,out], + 0/1 mask on the input rows) -> fp16 B operand
// [n_pad rows (out)][k_pad (in + bias slot)] in the K-major core-matrix layout:
//   half index = (n/8)*(k_pad*8) + (k/8)*64 + (n%8)*8 + (k%8)
__global__ void pack_layer_f16_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                      const float* __restrict__ mask, long long m_sp, long long m_se, int E, int in,
                                      int out, int k_pad, int n_pad, int interleave_d, float scale,
                                      __half* __restrict__ dst, long long net_stride) {
    const int g = blockIdx.y;
    const long long total = (long long)n_pad * k_pad;
    const float* wg = w + (long long)g * in * out;
    const float* mg = mask ? mask + (g / E) * m_sp + (g % E) * m_se : nullptr;
    __half* d = dst + g * net_stride;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // iterate in destination order (coalesced stores)
        int kk = (int)(i % 8);
        int nn = (int)((i / 8) % 8);
        long long rest = i / 64;
        int kc = (int)(rest % (k_pad / 8));
        int ng = (int)(rest / (k_pad / 8));
        int n = ng * 8 + nn, k = kc * 8 + kk;
        float v = 0.0f;
        if (n < out) {
            // Gaussian PLAIN heads are stored with (mean_d, logvar_d) in adjacent columns (2d, 2d+1)
            const int src = interleave_d ? ((n & 1) ? interleave_d + (n >> 1) : (n >> 1)) : n;
            if (k < in) {
                v = wg[(long long)k * out + src];
                if (mg) v *= mg[k];
            } else if (k == in && b) {
                v = b[(long long)g * out + src];
            }
        }
        v = fminf(fmaxf(v * scale, -65504.0f), 65504.0f);
        d[i] = __float2half_rn(v);
    }
}

}  // namespace tc
}  // namespace cmrl

namespace cmrl {
namespace tc {
int launch_rollout_v3(const RolloutP& p, cudaStream_t st);
int launch_rollout_v5(const RolloutP& p, cudaStream_t st);
int launch_rollout_v5_groups(const RolloutP* groups, int G, cudaStream_t st);
int launch_rollout_v6(const RolloutP& p, cudaStream_t st);
bool rollout_v5_supports(const RolloutP& p);
}
}  // namespace cmrl

using namespace cmrl;

extern "C" int cmrl_pack_layer_f16(const float* w, const float* b, const float* mask, long long m_sp, long long m_se,
                                   int P, int E, int in_size, int out_size, int k_pad, int n_pad, int interleave_d,
                                   float scale, void* dst_half, long long net_stride_halfs, void* stream) {
    CMRL_REQUIRE(w && dst_half, "pack_layer_f16: null pointer");
    CMRL_REQUIRE(k_pad % 16 == 0 && n_pad % 16 == 0 && k_pad >= in_size + 1 && n_pad >= out_size,
                 "pack_layer_f16: k_pad must be a multiple of 16 >= in+1, n_pad a multiple of 16 >= out");
    CMRL_REQUIRE(P * E <= 65535, "pack_layer_f16: too many nets");
    CMRL_REQUIRE(interleave_d == 0 || out_size == 2 * interleave_d, "pack_layer_f16: interleave_d must be out_size / 2");
    long long total = (long long)n_pad * k_pad;
    int bx = (int)((total + 255) / 256);
    if (bx > 1024) bx = 1024;
    dim3 grid(bx, P * E);
    tc::pack_layer_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, b, mask, m_sp, m_se, E, in_size, out_size, k_pad,
                                                                    n_pad, interleave_d, scale, (__half*)dst_half, net_stride_halfs);
    CMRL_CHECK_LAUNCH("pack_layer_f16");
    return 0;
}

static int fill_rollout_p(tc::RolloutP& p, const void* packed, long long blob_halfs, const int* layer_off_halfs, const int* layer_k,
                          const int* layer_n, int num_layers, int P, int E, int N, int O, int A, int H, int head_dim, int layout,
                          int gaussian, int act, const float* obs, const float* actn, const int* perm, const int* offsets, int residual,
                          const float* min_logvar, const float* max_logvar, int bounds_n, const float* noise, float* pred,
                          long long pred_member_stride, int variant) {
    CMRL_REQUIRE(packed && obs && perm && offsets && pred, "mlp_rollout_f16: null pointer");
    CMRL_REQUIRE(num_layers >= 1 && num_layers <= 7, "mlp_rollout_f16: 1..7 hidden layers supported");
    CMRL_REQUIRE(E >= 1 && E <= 64, "mlp_rollout_f16: E must be <= 64");
    CMRL_REQUIRE(!gaussian || (min_logvar && max_logvar), "mlp_rollout_f16: gaussian head needs logvar bounds");
    p.packed = (const __half*)packed;
    p.blob_halfs = blob_halfs;
    const int HP = layer_n[0];
    for (int l = 0; l <= num_layers; ++l) {
        p.layer_off[l] = layer_off_halfs[l];
        p.layer_k[l] = layer_k[l];
        p.layer_n[l] = layer_n[l];
        CMRL_REQUIRE(layer_k[l] % 16 == 0 && layer_n[l] % 16 == 0 && layer_n[l] <= ((variant == 6 || (variant == 0 && layer_n[0] == H)) ? 512 : 256) && layer_n[l] >= 16,
                     "mlp_rollout_f16: layer %d dims must be multiples of 16 with N <= 256 (512 for variant 6)", l);
        if (!(variant == 6 || (variant == 0 && layer_n[0] == H))) {
            if (l >= 1) CMRL_REQUIRE(layer_k[l] == HP, "mlp_rollout_f16: hidden width must be uniform");
            if (l < num_layers) CMRL_REQUIRE(layer_n[l] == HP, "mlp_rollout_f16: hidden width must be uniform");
        }
    }
    CMRL_REQUIRE(variant == 6 || (variant == 0 && HP == H) || H < HP, "mlp_rollout_f16: padded width must leave a bias slot (H < HP)");
    CMRL_REQUIRE(layout == CMRL_HEAD_PLAIN ? layer_n[num_layers] >= (gaussian ? 2 : 1) * head_dim : true,
                 "mlp_rollout_f16: head too narrow");
    p.num_layers = num_layers;
    p.P = P;
    p.E = E;
    p.N = N;
    p.O = O;
    p.A = A;
    p.H = H;
    p.head_dim = head_dim;
    p.layout = layout;
    p.gaussian = gaussian;
    p.act = act;
    p.obs = obs;
    p.actn = actn;
    p.perm = perm;
    p.offsets = offsets;
    p.residual = residual;
    p.min_logvar = min_logvar;
    p.max_logvar = max_logvar;
    p.bounds_n = bounds_n;
    p.noise = noise;
    p.pred = pred;
    p.pred_se = pred_member_stride;
    return 0;
}

extern "C" int cmrl_mlp_rollout_f16(const void* packed, long long blob_halfs, const int* layer_off_halfs,
                                    const int* layer_k, const int* layer_n, int num_layers, int P, int E, int N, int O,
                                    int A, int H, int head_dim, int layout, int gaussian, int act, const float* obs,
                                    const float* actn, const int* perm, const int* offsets, int residual,
                                    const float* min_logvar, const float* max_logvar, int bounds_n, const float* noise,
                                    float* pred, long long pred_member_stride, int variant, void* stream) {
    if (N == 0) return 0;
    tc::RolloutP p{};
    const int rc = fill_rollout_p(p, packed, blob_halfs, layer_off_halfs, layer_k, layer_n, num_layers, P, E, N, O, A, H, head_dim, layout,
                                  gaussian, act, obs, actn, perm, offsets, residual, min_logvar, max_logvar, bounds_n, noise, pred,
                                  pred_member_stride, variant);
    if (rc != 0) return rc;
    // variant 0: chosen by shape -- un-widened packing (hidden layers too wide for the weight-stationary kernels) -> 6; two tiles'
    // accumulators and heads fit tensor memory and the operands fit shared memory -> 5; otherwise 3
    if (variant == 0) variant = layer_n[0] == H ? 6 : (tc::rollout_v5_supports(p) ? 5 : 3);
    if (variant == 6) return tc::launch_rollout_v6(p, (cudaStream_t)stream);
    if (variant == 5) return tc::launch_rollout_v5(p, (cudaStream_t)stream);
    CMRL_REQUIRE(variant == 3, "mlp_rollout_f16: variant must be 0 (by shape), 5 (CTA pairs, two 256-row tiles, A operand partly in TMEM), "
                               "3 (pairs, two 128-row tiles in flight) or 6 (pairs, streamed weights, wide hidden)");
    return tc::launch_rollout_v3(p, (cudaStream_t)stream);
}

/* Several mechanisms of one step (transition, reward mech, termination mech) in ONE launch of the weight-stationary pair kernel:
 * see the header.  Returns -2 (and sets the error text) when the mechanisms cannot share a launch -- the caller then launches
 * them one by one. */
extern "C" int cmrl_mlp_rollout_f16_groups(const cmrl_rollout_group* groups, int n_groups, void* stream) {
    CMRL_REQUIRE(groups && n_groups >= 1 && n_groups <= 3, "mlp_rollout_f16_groups: 1..3 groups");
    tc::RolloutP ps[3];
    for (int g = 0; g < n_groups; ++g) {
        const cmrl_rollout_group& q = groups[g];
        ps[g] = tc::RolloutP{};
        CMRL_REQUIRE(q.N >= 1, "mlp_rollout_f16_groups: empty group");
        const int rc = fill_rollout_p(ps[g], q.packed, q.blob_halfs, q.layer_off_halfs, q.layer_k, q.layer_n, q.num_layers, q.P, q.E, q.N, q.O,
                                      q.A, q.H, q.head_dim, q.layout, q.gaussian, q.act, q.obs, q.actn, q.perm, q.offsets, q.residual,
                                      q.min_logvar, q.max_logvar, q.bounds_n, q.noise, q.pred, 0, 5);
        if (rc != 0) return rc;
    }
    if (!tc::rollout_v5_supports(ps[0])) {
        set_error("mlp_rollout_f16_groups: the pair kernel with two 256-row tiles does not take this model shape");
        return -2;
    }
    return tc::launch_rollout_v5_groups(ps, n_groups, (cudaStream_t)stream);
}
