/*
 * cmrl_b200.h -- C ABI of libcmrl_b200.so: the sm_100a kernels behind the cmrl
 * ensemble-dynamics hot path (polixir/causal-mbrl, reference tree /root/reference).
 *
 * The reference has no FFI: its "operators" are torch/numpy calls inside Python
 * classes.  Each entry point below names the reference call site it replaces
 * (file:line relative to /root/reference).  The Python host mirror
 * (causal-mbrl_b200/, importable as `cmrl_b200`) binds these with ctypes; see
 * INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller unless the name ends in
 *    `_host`.  The library never allocates, frees or retains caller memory.
 *  - All tensors are dense row-major fp32 unless stated.  Group dims: P = parallel
 *    sub-networks (ExternalMaskTransition: one per output dim), E = ensemble members.
 *    A "net" is one (p,e) pair.  A stride of 0 broadcasts that dim.
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and
 *    performs no host synchronisation.
 *  - Return value: 0 ok; <0 invalid argument / unsupported shape; >0 a cudaError_t.
 *    cmrl_last_error() returns a thread-local message for the last non-zero return.
 *  - No CPU fallback exists: on a machine without an sm_100 device every compute entry
 *    returns cudaErrorNoDevice / cudaErrorNoKernelImageForDevice.
 */
#ifndef CMRL_B200_H
#define CMRL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMRL_B200_ABI_VERSION 1

/* CMRL_ACT_SILU_HALF (fused tensor-core rollout only): SiLU evaluated on HALVED pre-activations, h + h*tanh(h) with
 * h = z/2 -- the hidden layers must have been packed with scale 0.5 (cmrl_pack_layer_f16), which is exact in binary
 * floating point and saves one multiply per activation in the epilogue. */
enum cmrl_act { CMRL_ACT_IDENTITY = 0, CMRL_ACT_RELU = 1, CMRL_ACT_SILU = 2, CMRL_ACT_SILU_HALF = 3 };
/* layout of the head's raw output (see cmrl_gaussian_head_fwd) */
enum cmrl_head_layout { CMRL_HEAD_PLAIN = 0, CMRL_HEAD_PARALLEL = 1 };

int cmrl_abi_version(void);
const char* cmrl_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long cmrl_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * K1  EnsembleLinearLayer.forward  (cmrl/models/layers.py:60-65)  and
 *     ParallelEnsembleLinearLayer.forward (layers.py:106-111):  y = act(x.matmul(W) + b),
 *     with the activation module of plain_transition.py:69-83 fused, and the input mask of
 *     ExternalMaskTransition.mask_input (external_mask_transition.py:116-139) fused into the
 *     A-operand load instead of materialising repeat(P) * mask (:149-150).
 *
 *  x    [P?,E?,B,in]   element strides x_sp / x_se between sub-nets / members (0 = shared)
 *  w    [P,E,in,out]   b [P,E,1,out] or NULL
 *  y    [P,E,B,out]    post-activation;  z: same shape, pre-activation, may be NULL
 *  mask NULL, or 0/1 floats addressed mask[p*m_sp + e*m_se + b*m_sb + i]
 *  row_offsets NULL, or int32 [E+1] (device): ragged mode -- member e owns rows
 *       [row_offsets[e], row_offsets[e+1]) of a flat [B,in] x and [P,B,out] y (x_se ignored);
 *       used by selected-member rollouts (VecFakeEnv.get_dynamics_predict, fake_env.py:195-215).
 *  tier (all three K1 entry points): 0 = fp32 FFMA, ascending-k accumulation (row results independent of
 *       tile placement); 1 = tcgen05 kind::tf32 (1e-3 tier); 3 = 3xTF32 error-compensated tcgen05 (operands split
 *       into tf32 hi + lo, three MMAs per K step: fp32-grade, 1e-5 tier).  Ragged calls always use tier 0.
 * ------------------------------------------------------------------------------------------ */
int cmrl_ens_linear_fwd(const float* x, long long x_sp, long long x_se, const float* w, const float* b,
                        float* y, float* z, int P, int E, int B, int in_size, int out_size,
                        const float* mask, long long m_sp, long long m_se, long long m_sb,
                        const int* row_offsets, int act, int tier, void* stream);

/* Backward of K1 w.r.t. the input, fused with the previous layer's activation derivative
 * (autograd of layers.py:61 + the activation; base_dynamics.py:185 `.backward()`):
 *   dx[p,e,b,i] = (sum_o dz[p,e,b,o] * w[p,e,i,o]) * act'(z_prev[p,e,b,i])
 * z_prev NULL or act_prev == IDENTITY -> plain dz.W^T. */
int cmrl_ens_linear_bwd_dx(const float* dz, const float* w, const float* z_prev, float* dx, int P, int E,
                           int B, int in_size, int out_size, int act_prev, int tier, void* stream);

/* Backward of K1 w.r.t. weight and bias:
 *   dw[p,e,i,o] = sum_b (x*mask)[p,e,b,i] * dz[p,e,b,o];   db[p,e,0,o] = sum_b dz[p,e,b,o]
 * db may be NULL.  accumulate != 0 adds into dw/db instead of overwriting.
 * workspace (optional, caller-owned device floats; NULL / 0 = none): with tier 1 / 3 and B >= 2048 the batch (= the K
 * dimension of this GEMM) is cut into up to 32 slices that run as separate CTAs; each slice needs
 * P*E*(in_size+1)*out_size floats and the slices are summed in order (deterministic). */
int cmrl_ens_linear_bwd_dw(const float* x, long long x_sp, long long x_se, const float* mask, long long m_sp,
                           long long m_se, long long m_sb, const float* dz, float* dw, float* db, int P,
                           int E, int B, int in_size, int out_size, int accumulate, int tier, float* workspace,
                           long long workspace_floats, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  Gaussian head: split, double soft-clamp, residual (and the ExternalMask transpose)
 *     plain_transition.py:111-120, external_mask_transition.py:154-167, plain_reward_mech.py:86-92
 *   PLAIN    raw [E,B,2D] (deterministic: [E,B,D]);  PARALLEL raw [D(=P),E,B,2] (det.: [D,E,B,1])
 *   mean/logvar out [E,B,D];  lv = max - softplus(max - raw_lv); lv = min + softplus(lv - min)
 *   (softplus beta=1, threshold=20);  min_logvar/max_logvar: [D] (bounds_n == D) or [1].
 *   obs (for the residual): [E?,B,obs_ld] with member stride obs_se (0 = shared); NULL = no residual.
 *   logvar == NULL -> deterministic head (mean only).
 * ------------------------------------------------------------------------------------------ */
int cmrl_gaussian_head_fwd(const float* raw, int layout, int E, int B, int D, const float* obs, long long obs_se,
                           int obs_ld, const float* min_logvar, const float* max_logvar, int bounds_n,
                           float* mean, float* logvar, void* stream);

/* Backward of K2: given g_mean, g_logvar [E,B,D] (g_logvar may be NULL) writes d_raw in raw's layout and,
 * when d_min/d_max are non-NULL, ACCUMULATES (atomicAdd) the gradients of the bounds into them. */
int cmrl_gaussian_head_bwd(const float* raw, int layout, int E, int B, int D, const float* min_logvar,
                           const float* max_logvar, int bounds_n, const float* g_mean, const float* g_logvar,
                           float* d_raw, float* d_min, float* d_max, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  losses.  gaussian_nll (cmrl/models/util.py:14-37, reduce=False) + the constant of
 *     EnsembleMLP.get_nll_loss (mlp.py:72-76):  loss = (mu-y)^2 * exp(-lv) + lv + reg
 *     EnsembleMLP.get_mse_loss (mlp.py:68-70):  loss = (mu-y)^2
 *     target: [E?,n_per_member] with member stride t_se (0 = same target for all members).
 * ------------------------------------------------------------------------------------------ */
int cmrl_nll_loss_fwd(const float* mean, const float* logvar, const float* target, long long t_se, int E,
                      long long n_per_member, float reg, float* loss, void* stream);
/* g_mean = g * 2 (mu-y) exp(-lv);  g_logvar = g * (1 - (mu-y)^2 exp(-lv)).  g_loss NULL -> g = g_scalar
 * (the `.mean()` of base_dynamics.py:185 gives g_scalar = 1/(E*B*D)). */
int cmrl_nll_loss_bwd(const float* mean, const float* logvar, const float* target, long long t_se, int E,
                      long long n_per_member, const float* g_loss, float g_scalar, float* g_mean,
                      float* g_logvar, void* stream);
int cmrl_mse_loss_fwd(const float* mean, const float* target, long long t_se, int E, long long n_per_member,
                      float* loss, void* stream);
/* Its gradient (autograd of `loss.mean().backward()` for deterministic mechanisms): g_mean = 2 (mean - target) g_loss. */
int cmrl_mse_loss_bwd(const float* mean, const float* target, long long t_se, int E, long long n_per_member,
                      const float* g_loss, float* g_mean, void* stream);
/* Hold-out MSE for elite selection (plain_dynamics.py:74,82 `evaluate(...).mean(dim=(1,2))`):
 * sums[e] += sum_i (mean[e,i]-target[e,i])^2 in fp64 (caller zeroes `sums` [E] doubles, divides by n). */
int cmrl_mse_member_sums(const float* mean, const float* target, long long t_se, int E, long long n_per_member,
                         double* sums, void* stream);

/* torch.optim.Adam as built at base_dynamics.py:58-63 (L2-coupled weight decay, bias correction,
 * eps outside the sqrt), one launch over a flat parameter arena.  grad_scale multiplies the gradient
 * first (data-parallel averaging). `step` is the 1-based step count after this update. */
int cmrl_adam_l2_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, float grad_scale, int step,
                      void* stream);

/* Same update with the step counter kept in device memory, so a captured CUDA graph can be replayed step after step.
 * step_dev points at TWO ints: [0] = steps already taken (the call increments it: the last block of the kernel to finish
 * does, there is no trailing launch), [1] = that block's arrival ticket, zero between calls. */
int cmrl_adam_l2_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                          float beta1, float beta2, float eps, float weight_decay, float grad_scale, int* step_dev,
                          void* stream);

/* acc[0] += sum of x[0..n) in fp64, deterministic (fixed assignment of elements to blocks, partial sums added in block
 * order by the last block to finish): the running epoch loss of BaseDynamics.train (`loss.mean()` of every batch,
 * base_dynamics.py:181-187) without a host round trip.  workspace: 66 doubles, zeroed once by the caller. */
int cmrl_sum_acc_f64(const float* x, long long n, double* acc, double* workspace, void* stream);

/* Device-side bootstrap batch assembly (BootstrapIterator.__next__ + _consolidate_batches,
 * cmrl/util/transition_iterator.py:14-28,131-139): for member e and row b < B
 *   src = member_idx[e, order[min(*start_dev + b, N-1)]];  out_*[e,b,:] = {obs, act, target}[src,:]
 * obs [N,O], act [N,A], target [N,D] are the device-resident training set; member_idx int32 [E,N], order int32 [N].
 * start_dev points at TWO ints: [0] = first row, [1] = arrival ticket (zero between calls); advance != 0 is added to
 * start_dev[0] by the last block of the gather to finish (graph-replayable epoch loop, no trailing launch). */
int cmrl_bootstrap_gather(const float* obs, const float* act, const float* target, const int* member_idx,
                          const int* order, int* start_dev, int advance, int E, int B, int N, int O, int A, int D,
                          float* out_obs, float* out_act, float* out_target, void* stream);

/* ------------------------------------------------------------------------------------------
 * K5  rollout step (VecFakeEnv.step_wait, cmrl/models/fake_env.py:97-142)
 * ------------------------------------------------------------------------------------------ */
/* Device RNG (throughput mode; the parity mode supplies numpy draws): Philox4x32-10 keyed by
 * (seed, stream_id), counter = (env, draw).  member_idx[n] = elite[u32 * n_elite >> 32] (uint8),
 * noise [N,D] ~ N(0,1) by Box-Muller.  Replaces Generator.choice + Generator.normal (fake_env.py:205-214,
 * mlp.py:40-44).  noise may be NULL (deterministic rollouts).  stream_base_dev: NULL, or a device counter whose value
 * is added to stream_id (so a captured CUDA graph draws fresh numbers on every replay; see cmrl_counter_add). */
int cmrl_draw_member_noise(unsigned long long seed, unsigned long long stream_id,
                           const unsigned long long* stream_base_dev, int N, const uint8_t* elite, int n_elite, int D,
                           uint8_t* member_idx, float* noise, void* stream);

/* out[n] = uniform integer in [0, upper) (Philox): reset-row draw replacing np.random.randint at
 * fake_env.py:154,177 in throughput mode. */
int cmrl_draw_uniform_index(unsigned long long seed, unsigned long long stream_id,
                            const unsigned long long* stream_base_dev, int N, int upper, int* out, void* stream);
/* *counter_dev += delta (one-thread kernel): advances the RNG stream base inside a captured step graph. */
int cmrl_counter_add(unsigned long long* counter_dev, unsigned long long delta, void* stream);

/* Counting sort of envs by drawn member: perm [N] lists env ids grouped by member (ascending env id inside a
 * member: the sort is stable and deterministic), offsets [E+1].
 * counters: int32 [CMRL_BUCKET_SCRATCH_WARPS * E] scratch (per-warp-chunk counts), overwritten by the call. */
#define CMRL_BUCKET_SCRATCH_WARPS 512
int cmrl_member_bucket(const uint8_t* member_idx, int N, int E, int* perm, int* offsets, int* counters,
                       void* stream);

/* x_sorted[r, :] = concat(obs[perm[r]], act[perm[r]])  (torch.concat at plain_transition.py:108 on the
 * bucketed order).  perm NULL = identity. */
int cmrl_gather_inputs(const float* obs, const float* act, const int* perm, int N, int O, int A, float* x_sorted,
                       void* stream);

/* get_dynamics_predict (fake_env.py:195-215) on bucketed rows: for sorted row r (env = perm[r]) apply the
 * Gaussian head to raw_sorted (PLAIN [N,2D] or PARALLEL [D,N,2]; deterministic heads: [N,D] / [D,N,1] with
 * min_logvar NULL), add the residual obs[env] when obs != NULL, and write
 *   pred[env, d] = mean + sqrt(exp(logvar)) * noise[env, d]      (noise NULL -> mean). */
int cmrl_sample_scatter(const float* raw_sorted, int layout, const int* perm, int N, int D, const float* obs,
                        int obs_ld, const float* min_logvar, const float* max_logvar, int bounds_n,
                        const float* noise, float* pred, void* stream);

/* All-member variant: mean/logvar [E,N,D] from query (base_dynamics.py:190-204), member_idx [N]:
 *   pred[n,d] = mean[idx[n],n,d] + sqrt(exp(logvar[idx[n],n,d])) * noise[n,d]. */
int cmrl_select_sample(const float* mean, const float* logvar, const uint8_t* member_idx, int E, int N, int D,
                       const float* noise, float* pred, void* stream);

/* VecFakeEnv.get_penalty (fake_env.py:186-193): penalty[n] = max_e || mean[e,n,:] - avg_e mean[:,n,:] ||_2 */
int cmrl_mopo_penalty(const float* mean, int E, int N, int O, float* penalty, void* stream);

/* Step tail (fake_env.py:116-142), SB3 layout outputs:
 *   reward_out = reward - penalty_coeff * penalty (penalty NULL -> no penalty)
 *   env_len += 1; truncated = env_len >= max_episode_steps; done = (terminal != 0) | truncated
 *   terminal_obs = next_obs (pre-reset copy, the infos["terminal_observation"] rows)
 *   obs_out = done ? reset_obs[reset_index ? reset_index[n] : n] : next_obs;  env_len = done ? 0 : env_len
 *   terminal: float [N] (learned mech sample, non-zero => done: reference quirk) or NULL;
 *   terminal_u8: uint8 [N] flags from a termination_fn, or NULL.  n_done: int32 [1], incremented by #done. */
int cmrl_env_step_tail(const float* next_obs, const float* reward, const float* terminal, const uint8_t* terminal_u8,
                       const float* penalty, float penalty_coeff, int* env_len, int max_episode_steps,
                       const float* reset_obs, const int* reset_index, int N, int O, float* obs_out,
                       float* reward_out, uint8_t* done, uint8_t* truncated, float* terminal_obs, int* n_done,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core tier (tcgen05, fp16 operands / fp32 accumulation; 1e-3 parity tier)
 * ------------------------------------------------------------------------------------------ */
/* pack_weights: one layer of every net, fp32 state_dict tensors -> the private padded fp16 device layout
 * (UMMA K-major core-matrix order, bias folded in as K row `in_size`, optional 0/1 input mask
 * mask[p*m_sp + e*m_se + i] folded into the rows: exact per SURVEY Appendix B).  dst: half*, net g's layer
 * blob at dst + g*net_stride_halfs, n_pad*k_pad halves.  interleave_d = D for the Gaussian PLAIN head layer
 * (output columns re-ordered to 2d = mean_d, 2d+1 = logvar_d), 0 otherwise.  scale multiplies weights and bias
 * (1, or 0.5 for the hidden layers of a CMRL_ACT_SILU_HALF rollout).  Never visible in state_dict. */
int cmrl_pack_layer_f16(const float* w, const float* b, const float* mask, long long m_sp, long long m_se, int P,
                        int E, int in_size, int out_size, int k_pad, int n_pad, int interleave_d, float scale,
                        void* dst_half, long long net_stride_halfs, void* stream);

/* Fused selected-member rollout of ONE mechanism (VecFakeEnv.get_dynamics_predict, fake_env.py:195-215, over
 * PlainTransition / ExternalMaskTransition / PlainRewardMech / PlainTerminationMech forward): for every
 * bucketed env row (perm/offsets from cmrl_member_bucket) run the drawn member's whole MLP on tcgen05 tensor
 * cores with TMA-staged packed weights and write pred[env,:] = mean (+obs) + sqrt(exp(logvar)) * noise[env,:].
 * layer_*: arrays of num_layers+1 entries (hidden layers, then the head).
 * variant: 0 = chosen from the shape (the default of the Python host): 5 = weight-stationary CTA pairs (cta_group::2,
 * M=256) with two 256-row tiles in flight and part of the A operand in tensor memory, where two tiles' accumulators and heads
 * fit tensor memory; 3 = pairs with two 128-row tiles in flight (wider heads); 6 = pairs with the weights streamed in K chunks
 * (hidden layers too wide for shared memory, packed without widening: layer_n[0] == H).
 * pred_member_stride: 0 for one [N,D] output; with ALL-MEMBER buckets (perm = 0..N-1 repeated E times, offsets[e] = e*N: every
 * member evaluates every env -- what the MOPO penalty of fake_env.py:186-193 needs) pass N*D and a [E,N,D] output. */
int cmrl_mlp_rollout_f16(const void* packed, long long blob_halfs, const int* layer_off_halfs, const int* layer_k,
                         const int* layer_n, int num_layers, int P, int E, int N, int O, int A, int H, int head_dim,
                         int layout, int gaussian, int act, const float* obs, const float* actn, const int* perm,
                         const int* offsets, int residual, const float* min_logvar, const float* max_logvar,
                         int bounds_n, const float* noise, float* pred, long long pred_member_stride, int variant, void* stream);

/* One launch for several mechanisms of a VecFakeEnv step (fake_env.py:97-115: transition, reward mech, termination mech are three
 * independent get_dynamics_predict calls on the same observations and actions).  The groups must share depth, hidden width,
 * activation and packed layer tables (checked); weights, buckets, noise, head layout and output are per group.  Fields as the
 * arguments of cmrl_mlp_rollout_f16; the layer tables point at HOST arrays of num_layers + 1 ints. */
typedef struct cmrl_rollout_group {
    const void* packed;
    long long blob_halfs;
    const int *layer_off_halfs, *layer_k, *layer_n;
    int num_layers, P, E, N, O, A, H, head_dim, layout, gaussian, act, residual, bounds_n;
    const float *obs, *actn;
    const int *perm, *offsets;
    const float *min_logvar, *max_logvar, *noise;
    float* pred;
} cmrl_rollout_group;
int cmrl_mlp_rollout_f16_groups(const cmrl_rollout_group* groups, int n_groups, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused training step (BaseDynamics.train, base_dynamics.py:181-186: `loss = get_nll_loss(...); loss.mean().backward()`
 * over PlainTransition / ExternalMaskTransition / PlainRewardMech / PlainTerminationMech) on tcgen05 kind::tf32 tensor
 * cores.  passes = 3 -> 3xTF32 error compensation (fp32-grade, 1e-5 tier), passes = 1 -> plain TF32 (1e-3 tier).
 * Three launches per step: pack (weights -> streamed tf32 hi/lo B operands), fused (forward, Gaussian head, NLL and its
 * gradient, dz chain; one persistent CTA per (net, 128-row tile)), dw (every layer's weight / bias gradient).
 * L = hidden layers (the head is layer L); in_size = O + A; head_outputs = outputs PER NET (PLAIN head: D; PARALLEL
 * head: 1).  Limits: hidden <= 239, head_outputs <= 16, in_size <= hidden.  Returns <0 for unsupported shapes.
 * ------------------------------------------------------------------------------------------ */
/* sizes (floats) of the packed-weight blob per net, of the saved-activation scratch for a batch of B rows per net, and of
 * ONE batch slice of the dw workspace (pass a multiple of it to cmrl_train_dw to let large batches split over CTAs) */
long long cmrl_train_fused_sizes(int L, int in_size, int hidden, int head_outputs, int P, int E, int B,
                                 long long* blob_floats, long long* scratch_floats, long long* dw_ws_floats);
/* weights_host / biases_host: HOST arrays of L+1 device pointers ([P,E,in,out] / [P,E,1,out], layers.py:42-58,86-104).
 * mask: NULL or the 0/1 input mask of the first layer addressed mask[p*m_sp + e*m_se + i] (folded into W1). */
int cmrl_train_pack_tf32(const float* const* weights_host, const float* const* biases_host, int L, int P, int E,
                         int in_size, int hidden, int head_outputs, const float* mask, long long m_sp, long long m_se,
                         float* packed, void* stream);
/* obs [E,B,O] / actn [E,B,A] / target [E,B,tgt_ld] with member strides *_se (floats).  loss [E,B,head_dim] (may be
 * NULL) = gaussian_nll(reduce=False) + reg (util.py:14-37, mlp.py:72-76); the gradient is that of
 * g_scale * sum(loss) (g_scale = 1 / (E * global_rows * head_dim) reproduces loss.mean().backward()).
 * scratch: cmrl_train_fused_sizes floats; it carries the first layer's input, every hidden layer's pre-activations and
 * every layer's dz to cmrl_train_dw (which re-applies `act` to the pre-activations). */
int cmrl_train_fused(const float* packed, int L, int P, int E, int B, int O, int A, int hidden, int head_dim,
                     int layout, int act, int residual, int passes, const float* obs, long long obs_se,
                     const float* actn, long long act_se, const float* target, long long tgt_se, int tgt_ld,
                     const float* min_logvar, const float* max_logvar, int bounds_n, float reg, float g_scale,
                     float* loss, float* scratch, void* stream);
/* Rollout mode of the same kernel -- the fp32-grade (3xTF32) tier of the fused selected-member rollout: same contract as
 * cmrl_mlp_rollout_f16 (fake_env.py:195-215: perm / offsets from cmrl_member_bucket, pred[env,:] = mean (+obs) +
 * sqrt(exp(logvar)) * noise[env,:], noise == NULL: the mean) for Gaussian heads, with the weights packed by
 * cmrl_train_pack_tf32 (the backward operands of that blob are not read).  Only the forward phases run and nothing is saved. */
int cmrl_mlp_rollout_tf32(const float* packed, int L, int P, int E, int N, int O, int A, int hidden, int head_dim, int layout,
                          int act, int residual, const float* obs, const float* actn, const int* perm, const int* offsets,
                          const float* min_logvar, const float* max_logvar, int bounds_n, const float* noise, float* pred,
                          long long pred_member_stride, void* stream);
/* dw_host / db_host: HOST arrays of L+1 device pointers to the gradient tensors (overwritten, not accumulated).
 * mask as in cmrl_train_pack_tf32 (gradient rows of masked inputs are zero, as in the reference's masked forward).
 * workspace: optional device floats for deterministic batch-slice partial sums. */
int cmrl_train_dw(const float* scratch, int L, int P, int E, int B, int in_size, int hidden, int head_outputs, int act,
                  float* const* dw_host, float* const* db_host, const float* mask, long long m_sp, long long m_se,
                  float* workspace, long long workspace_floats, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (SURVEY 8b item 6 `allreduce_grads`, 8e row 2) ----------
 * The reference trains on one device (base_dynamics.py:173-188: loss.mean().backward(); optimizer.step()); with the
 * rows of every ensemble batch sharded over W ranks, the per-rank gradients of the flat arena are summed between
 * backward() and step().  The opaque handle is this library's own communicator: a cudaMalloc'ed exchange buffer per rank
 * that every peer maps through CUDA IPC (handles are exchanged by the caller, e.g. with torch.distributed.all_gather).
 *
 * cmrl_dp_comm_create: allocates the buffer for arenas of up to max_floats floats on the current device and writes its
 *   64-byte cudaIpcMemHandle_t to ipc_handle_out.  cmrl_dp_comm_connect: all_handles = world x 64 bytes in rank order.
 * cmrl_allreduce_grads: grad[0..n) <- sum over ranks, in rank order, bit-identical on every rank; ONE launch.
 * cmrl_allreduce_grads_adam: the same exchange fused with cmrl_adam_l2_step_dev (identical arithmetic; step_dev as there:
 *   two ints).  Both are stream-ordered and CUDA-graph replayable (epochs live in device memory); every
 *   rank must make the same sequence of calls with the same n.  A peer that does not arrive within ~20 s raises the
 *   communicator's error word (cmrl_dp_comm_error, synchronises) instead of hanging the device. */
int cmrl_dp_comm_create(int rank, int world, long long max_floats, void** comm_out, unsigned char* ipc_handle_out);
int cmrl_dp_comm_connect(void* comm, const unsigned char* all_handles);
int cmrl_dp_comm_error(void* comm, int* err_out);
int cmrl_dp_comm_destroy(void* comm);
int cmrl_allreduce_grads(void* comm, float* grad, long long n, void* stream);
int cmrl_allreduce_grads_adam(void* comm, float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, float grad_scale, int* step_dev,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMRL_B200_H */
