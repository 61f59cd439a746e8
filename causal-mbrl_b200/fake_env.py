"""VecFakeEnv with the reference's interface (cmrl/models/fake_env.py:23-230): a Stable-Baselines3 VecEnv whose
transitions come from the learned ensemble dynamics.  Observations, episode lengths and the model stay resident on
the GPU; one `step_wait` is a fixed sequence of libcmrl_b200 launches:

    member draw (+ Gaussian noise) -> bucket envs by member -> gather inputs -> K1 layers on the SELECTED member's
    rows only -> head + sampling + scatter -> (reward / termination mechs likewise) -> step tail

The reference evaluates all E members for every env and then indexes one (fake_env.py:195-215); evaluating only the
drawn member gives identical outputs (rows are independent) at 1/E of the FLOPs.  With `penalty_coeff != 0` the MOPO
penalty needs every member's mean, so the transition takes the all-member path.

Extra (non-reference) attributes, all optional:
    rng_mode     "numpy"  (default): member choice and noise drawn by `self.generator` in the reference's order --
                           bit-exact member choice / done flags against the reference for the same seed;
                 "device": Philox draws on the GPU (statistically equivalent, no host RNG work).
    precision    "fp32"   (default): 1e-5 relative parity with the reference -- the fused tcgen05 kernel with kind::tf32
                           operands split into hi + lo parts (3xTF32, fp32 accumulation) for Gaussian models it takes
                           (hidden width up to ~240), the CUDA-core FFMA kernels otherwise;
                 "ffma":  always the CUDA-core kernels (fp32 FMA in ascending-k order: bit-identical to `query`);
                 "f16":   the fused tcgen05 kernel with fp16 operands (fp32 accumulation), 1e-3 relative parity.
    infos_mode   "dicts"  (default): one fresh dict per env like the reference;
                 "shared": envs that are neither done nor truncated share one read-only dict (O(#done) host work).
"""
from typing import Any, Dict, List, Optional, Type

import numpy as np
import torch

from . import _lib, ops

try:  # the real SB3 base class when it is installed; an interface-compatible stand-in otherwise
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv
except Exception:  # pragma: no cover - SB3 is absent in the build image

    class VecEnv:  # type: ignore
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs = num_envs
            self.observation_space = observation_space
            self.action_space = action_space

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()


_MECHS = ("transition", "reward_mech", "termination_mech")


class VecFakeEnv(VecEnv):
    def __init__(self, num_envs: int, observation_space, action_space):
        super(VecFakeEnv, self).__init__(num_envs=num_envs, observation_space=observation_space, action_space=action_space)
        self.has_set_up = False

        self.penalty_coeff = None
        self.deterministic = None
        self.max_episode_steps = None

        self.dynamics = None
        self.reward_fn = None
        self.termination_fn = None
        self.learned_reward = None
        self.learned_termination = None
        self.get_init_obs_fn = None
        self.replay_buffer = None
        self.generator = np.random.default_rng()
        self.device = None
        self.logger = None

        self._current_batch_obs = None
        self._current_batch_action = None
        self._reset_by_buffer = True

        # B200-side knobs and state
        self.rng_mode = "numpy"
        self.infos_mode = "dicts"
        # device-resident f16 steps: transition + reward (+ termination) mechs in ONE kernel launch.  "auto": where it was measured
        # to pay -- three mechanisms, or few envs (launch bound); at 100 000 envs with two mechanisms one launch per mechanism
        # is as fast (0.243 vs 0.247 ms), with three it is 6 % slower than the grouped launch, at 1 000 envs 5 % slower
        self.group_mechs = "auto"
        self.precision = "fp32"  # "fp32" / "ffma": 1e-5 parity tiers (3xTF32 tensor cores / CUDA cores); "f16": tcgen05 tier (1e-3 parity)
        self.use_graph = True   # device rng_mode: replay the per-step launches from a CUDA graph
        self.return_views = False    # step(): hand out the pinned observation buffer itself instead of a copy (opt-in)
        self.overlap_prepare = True  # device rng_mode: reward / termination / reset draws + buckets on a side stream
        self._side_stream = None
        self._copy_stream = None
        self._stream_cache = {}
        self._shared_infos = None
        # graph-replayed steps draw and bucket the transition's members for the NEXT step beside this step's reward
        # mech (~18 us of small launches off the critical path); see _prepare_next_transition
        self.prep_ahead = True
        self._ahead_bufs = None      # persistent (idx, noise) of the transition draw
        self._ahead_family = None    # token of the graph family whose last replay filled them for the coming step
        self._prep_event = None
        self._prep_reset = None
        self._tail_graphs = {}
        self._obs_flip = False
        self.init_obs_pool_size = None  # device rng_mode + get_init_obs_fn: size of the pre-drawn reset pool
        self._device_seed = 0
        self._draw_counter = 0
        self._obs_dev = None
        self._env_len_dev = None
        self._rng_base = None  # device counter added to the Philox stream ids (advanced inside captured step graphs)
        self._step_graph = None
        self._step_graph_key = None
        self._step_graph_ahead = None
        self._pool_dev = None
        self._pool_rows = 0
        self._bufs = None

    # ------------------------------------------------------------------ set-up (fake_env.py:60-92)
    def set_up(self, dynamics, reward_fn=None, termination_fn=None, get_init_obs_fn=None, real_replay_buffer=None,
               penalty_coeff: float = 0.0, deterministic=False, max_episode_steps=1000, logger=None):  # fmt: skip
        self.dynamics = dynamics
        self.penalty_coeff = penalty_coeff
        self.deterministic = deterministic
        self.max_episode_steps = max_episode_steps

        self.reward_fn = reward_fn
        self.termination_fn = termination_fn
        assert self.dynamics.learned_reward or reward_fn
        assert self.dynamics.learned_termination or termination_fn
        self.learned_reward = self.dynamics.learned_reward
        self.learned_termination = self.dynamics.learned_termination
        self.get_init_obs_fn = get_init_obs_fn
        self.replay_buffer = real_replay_buffer
        self.logger = logger

        assert self.get_init_obs_fn or self.replay_buffer
        self._reset_by_buffer = self.replay_buffer is not None

        self.device = dynamics.device
        idx = torch.device(self.device).index
        self._device_index = torch.cuda.current_device() if idx is None else idx
        self.has_set_up = True
        self._alloc()

    def _alloc(self):
        # a (repeated) set_up starts from scratch: captured graphs hold the addresses of the previous buffers
        self._step_graph = None
        self._predict_graphs = None
        self._predict_graph = None
        self._tail_graphs = {}
        self._ahead_bufs = None
        self._ahead_family = None
        self._pool_dev = None
        self._pool_rows = 0
        if self._rng_base is not None:
            self._rng_base.zero_()
        N, dev = self.num_envs, self.device
        tr = self.dynamics.transition
        O, A = tr.obs_size, tr.action_size
        b = {}
        b["act"] = torch.empty((N, A), device=dev, dtype=torch.float32)
        b["act_pin"] = torch.empty((N, A), dtype=torch.float32).pin_memory()
        b["next_obs"] = torch.empty((N, O), device=dev, dtype=torch.float32)
        b["reward"] = torch.empty((N, 1), device=dev, dtype=torch.float32)
        b["terminal"] = torch.empty((N, 1), device=dev, dtype=torch.float32)
        # reward [N] f32 | done [N] u8 | truncated [N] u8 live in ONE buffer so that they come back in one copy
        b["tail_out"] = torch.empty(6 * N, device=dev, dtype=torch.uint8)
        b["reward_out"] = b["tail_out"][:4 * N].view(torch.float32)
        b["done"] = b["tail_out"][4 * N:5 * N]
        b["trunc"] = b["tail_out"][5 * N:]
        b["term_obs"] = torch.empty((N, O), device=dev, dtype=torch.float32)
        b["n_done"] = torch.zeros(1, device=dev, dtype=torch.int32)
        E = self.dynamics.ensemble_num
        for k in ("", "@transition", "@reward_mech", "@termination_mech"):  # one bucket workspace per mechanism (+ a shared one)
            b["perm" + k] = torch.empty(N, device=dev, dtype=torch.int32)
            b["offsets" + k] = torch.empty(E + 1, device=dev, dtype=torch.int32)
            b["counters" + k] = torch.empty(ops.BUCKET_SCRATCH_WARPS * E, device=dev, dtype=torch.int32)
        # pinned host mirrors for the per-step device->host read of the SB3 return values
        b["h_obs"] = torch.empty((N, O), dtype=torch.float32).pin_memory()
        b["h_obs2"] = torch.empty((N, O), dtype=torch.float32).pin_memory()
        b["h_term_u8"] = torch.empty(N, dtype=torch.uint8).pin_memory()
        b["term_u8"] = torch.empty(N, device=dev, dtype=torch.uint8)
        b["h_term_obs"] = torch.empty((N, O), dtype=torch.float32).pin_memory()
        b["h_tail_out"] = torch.empty(6 * N, dtype=torch.uint8).pin_memory()
        b["h_reward"] = b["h_tail_out"][:4 * N].view(torch.float32)
        b["h_done"] = b["h_tail_out"][4 * N:5 * N]
        b["h_trunc"] = b["h_tail_out"][5 * N:]
        self._bufs = b
        self._obs_dev = torch.zeros((N, O), device=dev, dtype=torch.float32)
        self._env_len_dev = torch.zeros(N, device=dev, dtype=torch.int32)
        self._elite_cache = {}

    @property
    def _envs_length(self):
        if self._env_len_dev is None:
            return np.zeros(self.num_envs, dtype=int)
        return self._env_len_dev.cpu().numpy().astype(int)

    @_envs_length.setter
    def _envs_length(self, value):
        if self._env_len_dev is not None:
            self._env_len_dev.copy_(torch.as_tensor(np.asarray(value), dtype=torch.int32))

    # ------------------------------------------------------------------ stepping (fake_env.py:94-142)
    def step_async(self, actions: np.ndarray) -> None:
        self._current_batch_action = actions

    def _elite_u8(self, mech):
        model = getattr(self.dynamics, mech)
        key = tuple(int(e) for e in model.elite_members)
        cached = self._elite_cache.get(mech)
        if cached is None or cached[0] != key:
            cached = (key, torch.tensor(key, dtype=torch.uint8, device=self.device))
            self._elite_cache[mech] = cached
        return cached[1]

    def _draw(self, mech, D):
        """Member index [N] (uint8, device) and N(0,1) noise [N,D] (device) for one mechanism, consuming
        `self.generator` exactly like get_dynamics_predict (fake_env.py:205-214) in numpy mode."""
        N = self.num_envs
        model = getattr(self.dynamics, mech)
        if self.rng_mode == "numpy":
            idx = model.get_random_index(N, self.generator)
            idx_dev = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.uint8)).to(self.device, non_blocking=True)
            noise_dev = None
            if not self.deterministic:
                noise = self.generator.normal(size=(N, D)).astype(np.float32)
                noise_dev = torch.from_numpy(noise).to(self.device, non_blocking=True)
            return idx_dev, noise_dev
        if self.rng_mode == "device":
            self._draw_counter += 1
            return ops.draw_member_noise(self._device_seed, self._draw_counter, N, self._elite_u8(mech), D,
                                         want_noise=not self.deterministic, stream_base=self._rng_base)  # fmt: skip
        raise ValueError("rng_mode must be 'numpy' or 'device'")

    def _selected_path(self, model, need_all):
        mask = model._kernel_mask()
        per_sample_mask = mask is not None and (mask.dim() == 4 or (mask.dim() == 3 and mask.shape[1] != model.ensemble_num))
        return not (need_all or per_sample_mask)

    def _prepare(self, mech, need_all=False):
        """Everything of a mechanism's prediction that does not depend on the observations: the member / noise draw and
        (selected-member path) the counting sort of the envs by drawn member, into this mechanism's own workspace."""
        model = getattr(self.dynamics, mech)
        idx, noise = self._draw(mech, model._head_dim)
        bucket = None
        if self._selected_path(model, need_all):
            b = self._bufs
            bucket = ops.member_bucket(idx, model.ensemble_num, b["perm@" + mech], b["offsets@" + mech], b["counters@" + mech])
        return idx, noise, bucket

    def _predict(self, mech, obs, act, all_member_out=None, prep=None):
        """next value of `mech`'s variable for every env: [N,D] device tensor."""
        model = getattr(self.dynamics, mech)
        N, D = self.num_envs, model._head_dim
        if prep is not None:
            idx, noise, bucket = prep
            if bucket is not None:
                return self.predict_selected(model, obs, act, idx, noise, bucket)
        else:
            idx, noise = self._draw(mech, D)
        mask = model._kernel_mask()
        per_sample_mask = mask is not None and (mask.dim() == 4 or (mask.dim() == 3 and mask.shape[1] != model.ensemble_num))
        if all_member_out is not None and not per_sample_mask:
            means = self._all_member_means(model, obs, act)
            if means is not None:
                # MOPO penalty (fake_env.py:186-193) needs every member's mean: one pass of the fused tensor-core kernel over
                # the all-member buckets (E x N rows), then the drawn member's sample through the selected-member pass
                all_member_out["mean"] = means
                return self.predict_selected(model, obs, act, idx, noise)
        if all_member_out is not None or per_sample_mask:
            with torch.no_grad():
                mean, logvar = model.forward(obs.unsqueeze(0), act.unsqueeze(0))
            if all_member_out is not None:
                all_member_out["mean"] = mean
            return ops.select_sample(mean, logvar, idx, noise)
        return self.predict_selected(model, obs, act, idx, noise)

    def _all_member_bufs(self, E, N, D, device):
        key = (E, N, D, str(device))
        cached = getattr(self, "_all_member_buckets", None)
        if cached is None or cached[0] != key:
            perm = torch.arange(N, device=device, dtype=torch.int32).repeat(E)
            offsets = (torch.arange(E + 1, device=device, dtype=torch.int64) * N).to(torch.int32)
            out = torch.empty((E, N, D), device=device, dtype=torch.float32)
            cached = self._all_member_buckets = (key, perm, offsets, out)
        return cached[1:]

    def _all_member_means(self, model, obs, act):
        """[E,N,D] means of EVERY member on the tensor-core tier of `self.precision`, or None when that tier does not take the
        model (then the all-member CUDA-core forward runs).  The buckets are trivial: member e's rows are envs 0..N-1."""
        if getattr(model, "repeat_times", None) is not None or model.deterministic:
            return None
        E, N, D = model.ensemble_num, obs.shape[0], model._head_dim
        perm, offsets, out = self._all_member_bufs(E, N, D, obs.device)
        mn, mx = model._bounds()
        if self.precision == "f16":
            return ops.mlp_rollout_f16(model.packed_f16(), model._parallel_num(), E, obs, act, perm, offsets, model.hid_size, D,
                                       model._head_layout, True, model._act_kind, model._residual(), mn, mx, None, out=out,
                                       all_members=True)  # fmt: skip
        if self.precision == "fp32":
            pack = model.packed_tf32()
            if pack is not None and pack.ok:
                return ops.mlp_rollout_tf32(pack, obs, act, perm, offsets, model._act_kind, model._residual(), mn, mx, None,
                                            out=out, all_members=True)  # fmt: skip
        return None

    def _prebuild_pack(self, model):
        """The packed operand copies of the active precision tier, built outside a graph capture (they allocate and launch)."""
        if model is self.dynamics.transition and self.penalty_coeff and getattr(model, "repeat_times", None) is None:
            self._all_member_bufs(model.ensemble_num, self.num_envs, model._head_dim, self.device)  # MOPO: all-member buckets
        model = getattr(model, "one_step_transition", model)  # ForwardEulerTransition rolls its one-step model out
        if self.precision == "f16":
            model.packed_f16()
        elif self.precision == "fp32" and hasattr(model, "packed_tf32"):
            model.packed_tf32()

    def predict_selected(self, model, obs, act, idx, noise, bucket=None):
        """Selected-member evaluation: bucket envs by drawn member, run the MLP on each member's rows only."""
        b = self._bufs
        N, D, E = obs.shape[0], model._head_dim, model.ensemble_num
        perm, offsets = bucket if bucket is not None else ops.member_bucket(idx, E, b["perm"], b["offsets"], b["counters"])
        repeat = getattr(model, "repeat_times", None)
        if repeat is not None:
            # ForwardEulerTransition (forward_euler.py:32-41): the drawn member's one-step MEAN is re-applied k times, the last
            # application supplies the log-variance the sample is drawn with
            cur = obs
            for k in range(int(repeat)):
                cur = self.predict_selected(model.one_step_transition, cur, act, idx, noise if k == repeat - 1 else None,
                                            bucket=(perm, offsets))
            return cur
        if self.precision == "f16":
            mn, mx = model._bounds()
            return ops.mlp_rollout_f16(model.packed_f16(), model._parallel_num(), E, obs, act, perm, offsets,
                                       model.hid_size, D, model._head_layout, not model.deterministic, model._act_kind,
                                       model._residual(), mn, mx, noise)  # fmt: skip
        assert self.precision in ("fp32", "ffma"), "precision must be 'fp32', 'ffma' or 'f16'"
        if self.precision == "fp32":
            # fp32-grade on tensor cores: kind::tf32 with hi/lo split operands (3xTF32) where the fused kernel takes the model
            pack = model.packed_tf32()
            if pack is not None and pack.ok:
                mn, mx = model._bounds()
                return ops.mlp_rollout_tf32(pack, obs, act, perm, offsets, model._act_kind, model._residual(), mn, mx, noise)
        h = ops.gather_inputs(obs, act, perm)
        params = model._stack_params()
        L = model.num_layers
        mask = model._kernel_mask()
        mask_str = ops.mask_strides(mask, model._parallel_num(), E, N) if mask is not None else (0, 0, 0)
        for i in range(L):
            h = ops.ens_linear_fwd(h, params[2 * i], params[2 * i + 1], model._act_kind, mask if i == 0 else None,
                                   mask_str, row_offsets=offsets)  # fmt: skip
        raw = ops.ens_linear_fwd(h, params[2 * L], params[2 * L + 1], ops.ACT_IDENTITY, row_offsets=offsets)
        mn, mx = model._bounds()
        return ops.sample_scatter(raw, model._head_layout, perm, N, D, obs if model._residual() else None, mn, mx, noise)

    def _predict_all(self, obs, act, need_all, after_transition=None, before_transition=None, ahead=None):
        """transition (+ reward / termination mechs): the device part of a step before any host callback.
        `after_transition(next_obs)` is called as soon as the transition prediction is enqueued, `before_transition()`
        after the draws / buckets (which do not read `act`) and before the first kernel that does."""
        all_out = {} if need_all else None
        preps = self._prepare_all(need_all, ahead=ahead)
        if before_transition is not None:
            before_transition()
        if after_transition is None and not need_all and self._group_on() and self.precision == "f16" and preps:
            # device-resident steps: every learned mechanism in ONE launch of the fused kernel (the reward mech alone is a
            # 60 us launch for a fifth of the transition's tiles: pipeline fill / drain and idle SMs at its tail)
            self._join_prepare()
            grouped = self._predict_grouped(obs, act, preps)
            if grouped is not None:
                if ahead is not None:
                    self._prepare_next_transition(ahead)
                    torch.cuda.current_stream().wait_stream(self._side_stream)
                return grouped + (None,)
        next_obs = self._predict("transition", obs, act, all_out, prep=preps.get("transition"))
        if after_transition is not None:
            after_transition(next_obs)
        if ahead is not None:
            self._prepare_next_transition(ahead)
        self._join_prepare()
        reward = self._predict("reward_mech", obs, act, prep=preps.get("reward_mech")) if self.learned_reward else None
        terminal_f = (self._predict("termination_mech", obs, act, prep=preps.get("termination_mech"))
                      if self.learned_termination else None)
        penalty = ops.mopo_penalty(all_out["mean"]) if need_all else None
        if ahead is not None:
            torch.cuda.current_stream().wait_stream(self._side_stream)  # the next step's draws / buckets
        return next_obs, reward, terminal_f, penalty

    def _group_on(self):
        if self.group_mechs == "auto":
            return (int(self.learned_reward) + int(self.learned_termination) >= 2) or self.num_envs <= 16384
        return bool(self.group_mechs)

    def _predict_grouped(self, obs, act, preps):
        """(next_obs, reward, terminal) from one grouped launch, or None when the mechanisms cannot share one (a mechanism
        without a selected-member bucket, a wrapper, different packed layer tables, a shape the pair kernel does not take)."""
        mechs = ["transition"] + [m for m, on in (("reward_mech", self.learned_reward), ("termination_mech", self.learned_termination)) if on]
        if len(mechs) < 2:
            return None
        groups = []
        for mech in mechs:
            model = getattr(self.dynamics, mech)
            prep = preps.get(mech)
            if prep is None or prep[2] is None or getattr(model, "repeat_times", None) is not None:
                return None
            idx, noise, (perm, offsets) = prep
            mn, mx = model._bounds()
            groups.append(dict(pack=model.packed_f16(), P=model._parallel_num(), E=model.ensemble_num, obs=obs, act=act, perm=perm,
                               offsets=offsets, H=model.hid_size, head_dim=model._head_dim, layout=model._head_layout,
                               gaussian=not model.deterministic, act_kind=model._act_kind, residual=model._residual(),
                               min_logvar=mn, max_logvar=mx, noise=noise))  # fmt: skip
        outs = ops.mlp_rollout_f16_groups(groups)
        if outs is None:
            return None
        out = dict(zip(mechs, outs))
        return out["transition"], out.get("reward_mech"), out.get("termination_mech")

    # ---- transition draws one step ahead (graph-replayed steps only) ---------------------------------------------
    def _ahead_ok(self, need_all):
        return (self.prep_ahead and self.overlap_prepare and self.rng_mode == "device"
                and self._selected_path(self.dynamics.transition, need_all))

    def _ahead_new(self, reset_source, grouped=False):
        """Capture-time description of a pipelined step graph: the transition kernel reads persistent draw / bucket
        buffers that the PREVIOUS replay filled (Philox stream id of this step's draw = t_id + device counter; the
        draw for the next step uses t_id + n, n = draws per step, because the graph advances the counter by n).
        grouped: the device-resident step launches all mechanisms in one kernel (`group_mechs`), so the reward /
        termination mechs' draws and buckets are made a step ahead as well ("mechs"; their stream ids follow the
        transition's in the fixed order of `_prepare_all`)."""
        tr = self.dynamics.transition
        mechs = []
        if grouped and self._group_on() and self.precision == "f16":
            mechs = [m for m, on in (("reward_mech", self.learned_reward), ("termination_mech", self.learned_termination)) if on]
        more = self.__dict__.setdefault("_ahead_more", {})
        for m in mechs:
            d = getattr(self.dynamics, m)._head_dim
            cur = more.get(m)
            if cur is None or cur[0].numel() != self.num_envs or (cur[1] is None) != bool(self.deterministic):
                more[m] = (torch.empty(self.num_envs, device=self.device, dtype=torch.uint8),
                           None if self.deterministic else torch.empty((self.num_envs, d), device=self.device))
                self._ahead_family = None
        if self._ahead_bufs is None or self._ahead_bufs[0].numel() != self.num_envs or \
                (self._ahead_bufs[1] is None) != bool(self.deterministic):
            idx = torch.empty(self.num_envs, device=self.device, dtype=torch.uint8)
            noise = None if self.deterministic else torch.empty((self.num_envs, tr._head_dim), device=self.device)
            self._ahead_bufs = (idx, noise)
            self._ahead_family = None
        n = 1 + int(self.learned_reward) + int(self.learned_termination) + int(bool(reset_source))
        return {"t_id": self._draw_counter + 1, "n": n, "mechs": mechs}

    def _ahead_draw(self, stream_id, mechs=()):
        tr = self.dynamics.transition
        b = self._bufs
        ops.draw_member_noise(self._device_seed, stream_id, self.num_envs, self._elite_u8("transition"), tr._head_dim,
                              want_noise=not self.deterministic, stream_base=self._rng_base, out=self._ahead_bufs)
        ops.member_bucket(self._ahead_bufs[0], tr.ensemble_num, b["perm@transition"], b["offsets@transition"],
                          b["counters@transition"])
        for k, m in enumerate(mechs):  # stream ids in the order of `_prepare_all`: transition, reward, termination
            model = getattr(self.dynamics, m)
            bufs = self._ahead_more[m]
            ops.draw_member_noise(self._device_seed, stream_id + 1 + k, self.num_envs, self._elite_u8(m), model._head_dim,
                                  want_noise=not self.deterministic, stream_base=self._rng_base, out=bufs)
            ops.member_bucket(bufs[0], model.ensemble_num, b["perm@" + m], b["offsets@" + m], b["counters@" + m])

    def _prepare_next_transition(self, ahead):
        """(inside a capture) after the transition kernel consumed the buffers: refill them for the next replay, on the
        side stream, concurrent with the reward / termination mechs"""
        side = self._side_stream
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._ahead_draw(ahead["t_id"] + ahead["n"], ahead.get("mechs", ()))

    def _ahead_before_replay(self, ahead):
        """the buffers hold this step's draws only if the previous thing that advanced the device counter was a replay
        of the same graph family; otherwise draw them now (first step, after a re-capture, another step flavour)"""
        if ahead is None:
            self._ahead_family = None
            return
        if self._ahead_family != ahead["family"]:
            self._ahead_draw(ahead["t_id"], ahead.get("mechs", ()))
        self._ahead_family = ahead["family"]

    def _prepare_all(self, need_all, reset_source=False, ahead=None):
        """device rng_mode: the draws and member buckets of every mechanism (and the reset-row draw) are independent of
        the transition network, so all but the transition's own run on a side stream while the fused transition kernel
        computes (~20 us of small launches per step off the critical path; fork / join through events, also under
        CUDA-graph capture).  The Philox stream ids are assigned here in the fixed order transition, reward,
        termination, reset -- independent of the execution order.  Returns {mech: (idx, noise, bucket)}."""
        self._prep_reset = None
        if self.rng_mode != "device":
            return {}
        if ahead is not None:  # drawn by the previous replay (or _ahead_before_replay); the stream id is consumed all the same
            self._draw_counter += 1
            assert self._draw_counter == ahead["t_id"]
            b = self._bufs
            preps = {"transition": self._ahead_bufs + ((b["perm@transition"], b["offsets@transition"]),)}
            for m in ahead.get("mechs", ()):  # drawn a step ahead too (grouped launch): consume their stream ids in order
                self._draw_counter += 1
                preps[m] = self._ahead_more[m] + ((b["perm@" + m], b["offsets@" + m]),)
        else:
            self._ahead_family = None  # the transition's bucket buffers are rewritten
            preps = {"transition": self._prepare("transition", need_all)}
        others = [m for m, on in (("reward_mech", self.learned_reward), ("termination_mech", self.learned_termination))
                  if on and m not in preps]
        if not self.overlap_prepare:
            for m in others:
                preps[m] = self._prepare(m)
            if reset_source:
                self._prep_reset = self._device_reset_source()
            return preps
        main = torch.cuda.current_stream()
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        side = self._side_stream
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for m in others:
                preps[m] = self._prepare(m)
                for x in preps[m][:2]:
                    if x is not None:
                        x.record_stream(main)  # allocated on the side stream, consumed on the main one
            if reset_source:
                self._prep_reset = self._device_reset_source()
                self._prep_reset[1].record_stream(main)
            # joined through an event, not the stream: with `ahead` more work follows on the side stream
            self._prep_event = torch.cuda.Event()
            self._prep_event.record(side)
        self._prep_forked = True
        return preps

    def _join_prepare(self):
        if getattr(self, "_prep_forked", False):
            torch.cuda.current_stream().wait_event(self._prep_event)
            self._prep_forked = False

    def _predict_all_graphed(self, obs, act, need_all, host_next_buf=None, act_pin=None):
        """Same, replayed from a CUDA graph in device rng_mode (the graph is re-captured when packed weights, elite
        sets or the precision changed).  host_next_buf: pinned [N,O] tensor that receives next_obs -- the copy is a
        node of the graph on a branch that starts right after the transition kernel, so it overlaps the reward /
        termination mechanisms (one graph per destination buffer: the host side ping-pongs between two).  With
        host_next_buf, `act_pin` (pinned [N,A]) is uploaded into `act` by a node on the copy branch beside the draws."""
        graphs = getattr(self, "_predict_graphs", None)
        if graphs is None:
            graphs = self._predict_graphs = {}
        slot = 0 if host_next_buf is None else host_next_buf.data_ptr()
        self._predict_graph = graphs.get(slot)
        try:
            return self._predict_all_graphed_one(obs, act, need_all, host_next_buf, act_pin)
        finally:
            graphs[slot] = self._predict_graph

    def _graph_state_key(self, need_all):
        """What a captured step graph bakes in besides the env's fixed buffers: the tier, the sampling mode and, per
        mechanism, the weights (token: torch version counters + the fused optimiser's counter -- a graph captured
        before a `learn()` must not be replayed after it: the fp16 copies it points at are stale), the packed copy, the
        elite list and the parameter arena."""
        parts = [self.precision, self.deterministic, need_all]
        for m in self.dynamics.learn_mech:
            model = getattr(self.dynamics, m)
            parts.append((model.weights_token(), id(model._packed.get("f16")), id(model._packed.get("tf32")), id(model._elite_members),
                          model._flat.data_ptr() if model._flat is not None else 0, bool(model.deterministic)))
        parts.append(self._group_on() if self.has_set_up else bool(self.group_mechs))
        return tuple(parts)

    def _predict_all_graphed_one(self, obs, act, need_all, host_next_buf, act_pin=None):
        key = self._graph_state_key(need_all)
        cached = getattr(self, "_predict_graph", None)
        if cached is None or cached[0] != key:
            if self._rng_base is None:
                self._rng_base = torch.zeros(1, dtype=torch.int64, device=self.device)
            for mech in self.dynamics.learn_mech:
                self._elite_u8(mech)
                self._prebuild_pack(getattr(self.dynamics, mech))
            key = self._graph_state_key(need_all)  # again: the packed copies may have been rebuilt
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            c0 = self._draw_counter
            l0 = _lib.launch_count()
            ahead = self._ahead_new(False) if self._ahead_ok(need_all) else None
            copy_stream = None
            if host_next_buf is not None:
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self.device)
                copy_stream = self._copy_stream

            # an EXTERNAL event becomes an event-record node of the graph: the host waits for the next_obs copy alone and
            # runs its callbacks while the reward / termination mechanisms are still executing
            copied = torch.cuda.Event(external=True) if copy_stream is not None else None

            def early_copy(next_obs):
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):
                    host_next_buf.copy_(next_obs, non_blocking=True)
                    copied.record(copy_stream)

            def join_upload():
                torch.cuda.current_stream().wait_stream(copy_stream)

            with ops.graph_capture(g):
                if copy_stream is not None:  # action upload on the copy branch, joined before the transition kernel
                    copy_stream.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(copy_stream):
                        act.copy_(act_pin, non_blocking=True)
                out = self._predict_all(obs, act, need_all, early_copy if copy_stream is not None else None,
                                        join_upload if copy_stream is not None else None, ahead=ahead)
                ops.counter_add(self._rng_base, self._draw_counter - c0)
                if copy_stream is not None:
                    torch.cuda.current_stream().wait_stream(copy_stream)
            if ahead is not None:
                assert self._draw_counter - c0 == ahead["n"]
                ahead["family"] = ("predict", key, ahead["t_id"], ahead["n"])
            # every predict graph (one per pinned destination buffer) bakes the same Philox stream ids; the device
            # counter the graphs advance makes them unique per step
            self._draw_counter = c0
            cached = (key, g, out, _lib.launch_count() - l0, copied, ahead)
            self._predict_graph = cached
        self._ahead_before_replay(cached[5])
        cached[1].replay()
        _lib.note_replayed(cached[3])
        self._next_obs_copied = cached[4]
        return cached[2]

    def _current_stream(self):
        """torch.cuda.current_stream() costs ~30 us of Python per call (device-index lookups); the raw handle is one
        C call, so the Stream objects are cached by handle."""
        raw = torch._C._cuda_getCurrentRawStream(self._device_index)
        st = self._stream_cache.get(raw)
        if st is None:
            if len(self._stream_cache) > 16:
                self._stream_cache.clear()
            st = self._stream_cache[raw] = torch.cuda.current_stream(self._device_index)
        return st

    def step_wait(self):
        assert self.has_set_up, "fake-env has not set up"
        assert len(self._current_batch_action.shape) == 2  # batch, action_dim
        b = self._bufs
        N = self.num_envs
        obs = self._obs_dev
        stream = self._current_stream()
        if self._current_batch_obs is None:  # device-resident steps ran in between
            self.sync_host_obs()
        b["act_pin"].numpy()[...] = self._current_batch_action
        act = b["act"]

        need_all = self.penalty_coeff != 0
        device_rng = self.rng_mode == "device"
        h_next = b["h_obs"] if self._obs_flip else b["h_obs2"]
        self._obs_flip = not self._obs_flip
        host_callbacks = not (self.learned_reward and self.learned_termination)
        copied_in_graph = False
        if device_rng and self.use_graph:
            # the action upload and the next_obs read-back are nodes of the predict graph (beside the draws / buckets and
            # beside the reward / termination mechanisms), with or without host callbacks
            next_obs, reward, terminal_f, penalty = self._predict_all_graphed(obs, act, need_all, h_next, b["act_pin"])
            copied_in_graph = True
        else:
            act.copy_(b["act_pin"], non_blocking=True)
            next_obs, reward, terminal_f, penalty = self._predict_all(obs, act, need_all)
        if penalty is not None and self.logger is not None:
            self.logger.record_mean("rollout/penalty", penalty.mean().item())

        # ---- host callbacks (reward_fn / termination_fn of the real env), on pinned copies ----------------------
        # The predicted next_obs lands in one of two pinned ping-pong buffers (the other one still holds the current
        # obs for the callbacks); in device rng_mode that same buffer, patched at the few reset rows, becomes the
        # current obs of the next step -- the full post-reset obs is never read back a second time.
        host_next = None
        terminal_u8 = None
        if host_callbacks:
            if not copied_in_graph:
                h_next.copy_(next_obs, non_blocking=True)
                stream.synchronize()
            else:
                self._next_obs_copied.synchronize()  # the copy node of the predict graph; the other mechs keep running
            host_next = h_next.numpy()
            host_obs = self._current_batch_obs
            if not self.learned_reward:
                r = np.asarray(self.reward_fn(host_next, host_obs, self._current_batch_action), dtype=np.float32)
                reward = b["reward"]
                reward.copy_(torch.from_numpy(np.ascontiguousarray(r.reshape(N, 1))), non_blocking=False)
            if not self.learned_termination:
                t = np.asarray(self.termination_fn(host_next, host_obs, self._current_batch_action))
                flags = b["h_term_u8"].numpy().view(np.bool_)
                if t.dtype == np.bool_:
                    np.copyto(flags, t.reshape(N))  # 100 KB memcpy; the generic compare below casts through int64 (~60 us)
                else:
                    np.not_equal(t.reshape(N), 0, out=flags)
                terminal_u8 = b["term_u8"]  # uploaded by tail_and_readback()

        need_term_obs = host_next is None and not copied_in_graph

        def tail_and_readback():
            """reset-row draw, fused tail, device -> host copies of the SB3 return values (all asynchronous)"""
            if terminal_u8 is not None:
                terminal_u8.copy_(b["h_term_u8"], non_blocking=True)
            reset_rows = reset_index = None
            if device_rng:
                reset_rows, reset_index = self._device_reset_source(tail=True)
            ops.env_step_tail(next_obs, reward, terminal_f, terminal_u8, penalty, self.penalty_coeff, self._env_len_dev,
                              self.max_episode_steps, reset_rows, reset_index, obs, b["reward_out"], b["done"], b["trunc"],
                              b["term_obs"] if need_term_obs else None, None)  # fmt: skip
            if need_term_obs:  # fully learned eager step: the pre-reset next_obs comes back now
                h_next.copy_(b["term_obs"], non_blocking=True)
            b["h_tail_out"].copy_(b["tail_out"], non_blocking=True)  # reward, done, truncated

        if device_rng and self.use_graph and copied_in_graph and (self.learned_reward or reward is b["reward"]):
            # second half of the step as one graph replay (5-6 launches / copies otherwise issued one by one from Python);
            # every operand is a fixed buffer: the predict graph's outputs, the pinned mirrors, this step's h_next
            if self._pool_dev is None or (self._reset_by_buffer and self._pool_rows != self._upper_bound()):
                c_keep = self._draw_counter
                self._device_reset_source(tail=True)  # (re)builds the reset pool outside any capture
                self._draw_counter = c_keep
            # keyed by the ADDRESSES the graph bakes in (object ids are recycled once a re-captured predict graph's
            # outputs are freed: a stale entry could match new tensors that live somewhere else)
            ptr = lambda x: 0 if x is None else x.data_ptr()  # noqa: E731
            key = (ptr(next_obs), ptr(reward), ptr(terminal_f), terminal_u8 is not None, ptr(penalty), need_term_obs,
                   h_next.data_ptr() if need_term_obs else 0, self._pool_dev.data_ptr(), self._pool_rows,
                   self.penalty_coeff, self.max_episode_steps)
            tg = self._tail_graphs.get(key)
            if tg is None:
                if len(self._tail_graphs) >= 4:
                    self._tail_graphs.clear()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                l0 = _lib.launch_count()
                c_keep = self._draw_counter
                with ops.graph_capture(g):
                    tail_and_readback()
                self._draw_counter = c_keep  # as for the predict graphs: baked id + device counter
                tg = self._tail_graphs[key] = (g, _lib.launch_count() - l0)
            tg[0].replay()
            _lib.note_replayed(tg[1])
        else:
            tail_and_readback()

        # ---- device -> host: the SB3 return values ------------------------------------------------------------------
        ret_obs = None
        if host_next is not None and not self.return_views:
            ret_obs = host_next.copy()  # the returned copy is made while the tail kernel and the small copies run
        stream.synchronize()
        if host_next is None:
            host_next = h_next.numpy()
        if ret_obs is None:
            # return_views: the observation array handed out IS the pinned ping-pong buffer (valid until the step after
            # the next one, which is how SB3's collectors use it: `_last_obs` is replaced after every step)
            ret_obs = host_next if self.return_views else host_next.copy()
        batch_done = b["h_done"].numpy().view(np.bool_)
        batch_truncate = b["h_trunc"].numpy().view(np.bool_)
        done_idx = np.nonzero(batch_done)[0]
        terminal_rows = host_next[done_idx] if done_idx.size else None  # fancy index = copy, taken before the patch
        if done_idx.size:
            if device_rng:  # rows the fused tail wrote into the done envs
                if done_idx.size * 4 < N:
                    rows = ops.gather_rows(obs, torch.from_numpy(done_idx.astype(np.int32)).to(self.device)).cpu().numpy()
                else:
                    b["h_term_obs"].copy_(obs, non_blocking=True)
                    stream.synchronize()
                    rows = b["h_term_obs"].numpy()[done_idx]
            else:  # host reset in the reference's order (ascending env index; fake_env.py:131-135,171-181)
                rows = np.ascontiguousarray(self._host_reset_rows(done_idx.size), dtype=np.float32)
                obs.index_copy_(0, torch.from_numpy(done_idx).to(self.device), torch.from_numpy(rows).to(self.device))
            host_next[done_idx] = rows
            if ret_obs is not host_next:
                ret_obs[done_idx] = rows
        self._current_batch_obs = host_next

        infos = self._make_infos(batch_truncate, done_idx, terminal_rows)
        # return_views: the reward array is the pinned read-back buffer itself (valid until the next step -- SB3's
        # collectors add it to their buffers before stepping again); `done` is always a copy (on-policy collectors keep
        # it as `_last_episode_starts` across the next step)
        host_reward = b["h_reward"].numpy()
        return ret_obs, (host_reward if self.return_views else host_reward.copy()), batch_done.copy(), infos

    def step_device(self, actions: torch.Tensor, termination=None, _ahead=None):
        """Device-resident step for rollout collectors that keep their buffers on the GPU: `actions` is a CUDA
        tensor [N,A]; nothing is copied to the host and nothing synchronises.  Needs rng_mode == "device" and a
        learned reward mech; termination is the learned mech, or `termination(next_obs, obs, actions)` -> CUDA
        uint8/bool [N] tensor, or None (episodes end by truncation only).  Returns views of internal buffers
        (next_obs_after_reset [N,O], reward [N], done u8 [N], truncated u8 [N], terminal_obs [N,O]) that are
        overwritten by the next step (SB3 ReplayBuffer slabs copy them out on the device)."""
        assert self.has_set_up, "fake-env has not set up"
        assert self.rng_mode == "device" and self.learned_reward
        b = self._bufs
        obs = self._obs_dev
        need_all = self.penalty_coeff != 0
        all_out = {} if need_all else None
        # Once a device-side stream base exists (graph-replayed steps advance it), an eager step advances the same base
        # instead of the host counter, so that eager steps and replays never share a Philox stream id.
        eager_base = self._rng_base is not None and not torch.cuda.is_current_stream_capturing()
        c_start = self._draw_counter
        preps = self._prepare_all(need_all, reset_source=True, ahead=_ahead)
        grouped = None
        if not need_all and self._group_on() and self.precision == "f16" and preps:
            all_ahead = _ahead is not None and len(_ahead.get("mechs", ())) == int(self.learned_reward) + int(self.learned_termination)
            if not all_ahead:  # the other mechs' draws / buckets of this step are on the side stream
                self._join_prepare()
            grouped = self._predict_grouped(obs, actions, preps)  # every learned mechanism in one launch of the fused kernel
            if grouped is None and all_ahead:
                pass  # (falls through to the per-mechanism launches below, which join before they read the side stream's work)
        terminal_f = terminal_u8 = None
        if grouped is not None:
            next_obs, reward, terminal_f = grouped
            if _ahead is not None:
                self._prepare_next_transition(_ahead)
            self._join_prepare()  # the reset-row draw (side stream) before the tail reads it
        else:
            next_obs = self._predict("transition", obs, actions, all_out, prep=preps.get("transition"))
            if _ahead is not None:
                self._prepare_next_transition(_ahead)
            self._join_prepare()
            reward = self._predict("reward_mech", obs, actions, prep=preps.get("reward_mech"))
            if self.learned_termination:
                terminal_f = self._predict("termination_mech", obs, actions, prep=preps.get("termination_mech"))
        if not self.learned_termination and termination is not None:
            terminal_u8 = termination(next_obs, obs, actions).to(torch.uint8).contiguous()
        penalty = ops.mopo_penalty(all_out["mean"]) if need_all else None
        reset_rows, reset_index = self._prep_reset
        ops.env_step_tail(next_obs, reward, terminal_f, terminal_u8, penalty, self.penalty_coeff, self._env_len_dev,
                          self.max_episode_steps, reset_rows, reset_index, obs, b["reward_out"], b["done"], b["trunc"],
                          b["term_obs"], None)  # fmt: skip
        if _ahead is not None:
            torch.cuda.current_stream().wait_stream(self._side_stream)  # the next step's draws / buckets
        if eager_base:
            ops.counter_add(self._rng_base, self._draw_counter - c_start)
            self._draw_counter = c_start
            self._ahead_family = None  # buffers drawn ahead by a replay belong to ids this step has just used
        self._current_batch_obs = None  # host mirror is stale now (refreshed lazily by step_wait / sync_host_obs)
        return obs, b["reward_out"], b["done"], b["trunc"], b["term_obs"]

    def step_device_graphed(self, actions: torch.Tensor):
        """`step_device` captured once in a CUDA graph and replayed: one host call per step instead of ~15 launches.
        The Philox stream ids baked into the graph are offset by a device-side counter that the graph itself advances,
        so every replay draws fresh members / noise / reset rows.  Weights, elite lists and the reset pool are read at
        replay time from the same device buffers; the graph is re-captured when the packed weights or the reset pool
        changed, after a `learn()` (the graph points at the fp16 copies of the weights), a new elite set or other step settings."""
        b = self._bufs
        if self._reset_by_buffer and self._pool_dev is not None and self._pool_rows != self._upper_bound():
            self._step_graph = None
        if self._step_graph is not None and self._step_graph_key != self._step_graph_state():
            self._step_graph = None  # weights (a `learn()` in between), elites, tier or step settings changed
        if self._step_graph is None:
            assert self.rng_mode == "device" and self.learned_reward
            if self._rng_base is None:
                self._rng_base = torch.zeros(1, dtype=torch.int64, device=self.device)
            for mech in self.dynamics.learn_mech:  # everything that allocates or syncs happens before the capture
                self._elite_u8(mech)
                self._prebuild_pack(getattr(self.dynamics, mech))
            c_keep = self._draw_counter
            self._device_reset_source()  # builds the reset pool; its draw is discarded so that the graph-replayed
            self._draw_counter = c_keep  # steps use the same Philox streams as kernel-by-kernel `step_device` calls
            b["act"].copy_(actions)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            c0 = self._draw_counter
            l0 = _lib.launch_count()
            ahead = self._ahead_new(True, grouped=True) if self._ahead_ok(self.penalty_coeff != 0) else None
            with ops.graph_capture(g):
                self._graph_out = self.step_device(b["act"], _ahead=ahead)
                ops.counter_add(self._rng_base, self._draw_counter - c0)
            if ahead is not None:
                assert self._draw_counter - c0 == ahead["n"]
                ahead["family"] = ("step", id(g), ahead["t_id"], ahead["n"])
            self._step_graph_ahead = ahead
            self._step_graph_launches = _lib.launch_count() - l0
            self._step_graph_key = self._step_graph_state()
            self._step_graph = g
        else:
            b["act"].copy_(actions, non_blocking=True)
        self._ahead_before_replay(self._step_graph_ahead)
        self._step_graph.replay()
        _lib.note_replayed(self._step_graph_launches)
        self._current_batch_obs = None
        return self._graph_out

    def _step_graph_state(self):
        return self._graph_state_key(self.penalty_coeff != 0) + (self.penalty_coeff, self.max_episode_steps)

    def invalidate_step_graph(self):
        self._step_graph = None
        self._ahead_family = None

    def rollout_device(self, policy, n_steps: int, termination=None, out: Optional[Dict[str, torch.Tensor]] = None):
        """Device-resident multi-step rollout into SB3 `ReplayBuffer`-layout slabs: replaces the
        per-step host loop around `fake_env.py:97-142` + SB3's `_store_transition` for collectors that keep their
        data on the GPU.  `policy(obs [N,O] cuda) -> actions [N,A] cuda` is called once per step; nothing is copied
        to the host and nothing synchronises.  Returns (or fills `out` with) float32 CUDA tensors in the layout of
        `stable_baselines3.common.buffers.ReplayBuffer` with `n_envs = num_envs`:
        `observations`, `next_observations` [T,N,O] (next_observations is the PRE-reset next state, i.e. the
        `terminal_observation` for done envs, as `OffPolicyAlgorithm._store_transition` stores it), `actions`
        [T,N,A], `rewards`, `dones`, `timeouts` [T,N] (timeouts = `TimeLimit.truncated`).
        Needs rng_mode == "device" and a learned reward mech; `termination` as in `step_device` (with a callable
        the step is launched kernel by kernel, otherwise replayed from the captured graph)."""
        assert self.has_set_up, "fake-env has not set up"
        N, T = self.num_envs, int(n_steps)
        O, A = self._obs_dev.shape[1], self._bufs["act"].shape[1]
        dev = self.device
        if out is None:
            out = {
                "observations": torch.empty((T, N, O), device=dev), "next_observations": torch.empty((T, N, O), device=dev),
                "actions": torch.empty((T, N, A), device=dev), "rewards": torch.empty((T, N), device=dev),
                "dones": torch.empty((T, N), device=dev), "timeouts": torch.empty((T, N), device=dev),
            }  # fmt: skip
        for t in range(T):
            out["observations"][t].copy_(self._obs_dev)
            actions = policy(self._obs_dev)
            assert actions.is_cuda and actions.shape == (N, A), "policy must return a CUDA tensor [num_envs, action_dim]"
            actions = actions.to(torch.float32).contiguous()
            out["actions"][t].copy_(actions)
            if termination is None and self.use_graph:
                _, reward, done, trunc, term_obs = self.step_device_graphed(actions)
            else:
                _, reward, done, trunc, term_obs = self.step_device(actions, termination)
            out["next_observations"][t].copy_(term_obs)
            out["rewards"][t].copy_(reward)
            out["dones"][t].copy_(done)
            out["timeouts"][t].copy_(trunc)
        return out

    def sync_host_obs(self):
        """Refresh the host mirror of the current observations after device-resident stepping."""
        self._current_batch_obs = self._obs_dev.cpu().numpy()
        return self._current_batch_obs.copy()

    @staticmethod
    def extend_replay_buffer(replay_buffer, slabs: Dict[str, torch.Tensor]):
        """Append `rollout_device` slabs to an SB3-style `ReplayBuffer` (`observations/next_observations/actions/
        rewards/dones/timeouts` arrays of shape [buffer_size, n_envs, ...], `pos`, `full`, `buffer_size`), with the
        ring-buffer wrap-around of `ReplayBuffer.add` (one device->host copy per field instead of T x n_envs adds)."""
        T = slabs["rewards"].shape[0]
        cap = replay_buffer.buffer_size
        fields = [k for k in ("observations", "next_observations", "actions", "rewards", "dones", "timeouts")
                  if getattr(replay_buffer, k, None) is not None]  # fmt: skip
        host = {k: slabs[k].cpu().numpy() for k in fields}
        pos = int(replay_buffer.pos)
        t = 0
        while t < T:
            n = min(T - t, cap - pos)
            for k in fields:
                dst = getattr(replay_buffer, k)
                dst[pos:pos + n] = host[k][t:t + n].reshape((n,) + dst.shape[1:])
            pos += n
            t += n
            if pos == cap:
                replay_buffer.full = True
                pos = 0
        replay_buffer.pos = pos

    def _make_infos(self, batch_truncate, done_idx, terminal_rows):
        """`terminal_rows[k]` = pre-reset next_obs of env `done_idx[k]` (fake_env.py:126-135)."""
        N = self.num_envs
        if self.infos_mode == "shared":
            # every truncated env is also done (done = terminal | truncated), so only the done rows get their own dict;
            # a step without done envs hands out one cached list (building 100 000 references costs ~20 us)
            base = self._shared_infos
            if base is None or len(base) != N:
                base = self._shared_infos = [{"TimeLimit.truncated": np.False_}] * N
            if done_idx.size == 0:
                return base
            infos = list(base)
            for k, i in enumerate(done_idx):
                infos[i] = {"TimeLimit.truncated": batch_truncate[i], "terminal_observation": terminal_rows[k]}
            return infos
        infos = [{"TimeLimit.truncated": t} for t in batch_truncate]
        for k, i in enumerate(done_idx):
            infos[i]["terminal_observation"] = terminal_rows[k]
        return infos

    # ------------------------------------------------------------------ resets (fake_env.py:144-181)
    def _upper_bound(self):
        return self.replay_buffer.buffer_size if self.replay_buffer.full else self.replay_buffer.pos

    def _host_reset_rows(self, n):
        """n reset observations in the reference's draw order: one global `np.random.randint` per done env
        (buffer resets), or `get_init_obs_fn` (one batched call of size n instead of n calls of size 1)."""
        if self._reset_by_buffer:
            upper_bound = self._upper_bound()
            inds = np.array([np.random.randint(0, upper_bound) for _ in range(n)], dtype=np.int64)
            return self.replay_buffer.observations[inds, 0]
        assert self.get_init_obs_fn is not None
        return np.asarray(self.get_init_obs_fn(n)).reshape(n, -1)

    def _device_reset_source(self, tail=False):
        """(rows [R,O] device, reset_index [N] device) for the fused tail in device rng_mode.  tail=True (host-facing
        step): the draw may be replayed from its own CUDA graph, where its stream id is a capture-time constant plus the
        device-side base the predict graph advances -- a disjoint id range (bit 40) keeps it from ever coinciding with a
        member / noise draw of another step."""
        N = self.num_envs
        if self._reset_by_buffer:
            ub = self._upper_bound()
            if self._pool_dev is None or self._pool_rows != ub:
                rows = np.ascontiguousarray(self.replay_buffer.observations[:ub, 0], dtype=np.float32)
                self._pool_dev, self._pool_rows = torch.from_numpy(rows).to(self.device), ub
        else:
            size = self.init_obs_pool_size or max(4096, N)
            if self._pool_dev is None or self._pool_rows != size:
                rows = np.ascontiguousarray(np.asarray(self.get_init_obs_fn(size)).reshape(size, -1), dtype=np.float32)
                self._pool_dev, self._pool_rows = torch.from_numpy(rows).to(self.device), size
        self._draw_counter += 1
        idx = ops.draw_uniform_index(self._device_seed, self._draw_counter + ((1 << 40) if tail else 0), N, self._pool_rows,
                                     self.device, stream_base=self._rng_base)  # fmt: skip
        return self._pool_dev, idx

    def reset(self, *, seed: Optional[int] = None, return_info: bool = False, options: Optional[dict] = None):
        if self.has_set_up:
            if self._reset_by_buffer:
                upper_bound = self._upper_bound()
                batch_inds = np.random.randint(0, upper_bound, size=self.num_envs)
                self._current_batch_obs = self.replay_buffer.observations[batch_inds, 0]
            else:
                self._current_batch_obs = self.get_init_obs_fn(self.num_envs)
            self._current_batch_obs = np.ascontiguousarray(self._current_batch_obs, dtype=np.float32)
            self._obs_dev.copy_(torch.from_numpy(self._current_batch_obs))
            self._env_len_dev.zero_()
            return self._current_batch_obs.copy()

    def seed(self, seed: Optional[int] = None):
        self.generator = np.random.default_rng(seed=seed)
        self._device_seed = 0 if seed is None else int(seed)
        self._draw_counter = 0
        if self._rng_base is not None:
            self._rng_base.zero_()  # in place: same seed -> same draws again
        # seed and stream ids are baked into captured graphs
        self._step_graph = None
        self._predict_graphs = None
        self._predict_graph = None
        self._tail_graphs = {}
        self._ahead_family = None

    def close(self) -> None:
        pass

    def env_is_wrapped(self, wrapper_class: Type, indices=None) -> List[bool]:
        return [False for _ in range(self.num_envs)]

    def single_reset(self, idx):
        assert self.has_set_up, "fake-env has not set up"
        self._env_len_dev[idx] = 0
        row = np.asarray(self._host_reset_rows(1), dtype=np.float32).reshape(-1)
        self._current_batch_obs[idx] = row
        self._obs_dev[idx] = torch.from_numpy(row).to(self.device)

    def render(self, mode="human"):
        raise NotImplementedError

    @staticmethod
    def get_penalty(ensemble_batch_next_obs):
        """MOPO max-disagreement penalty (fake_env.py:186-193) on the K5 kernel; accepts [E,N,O] numpy or tensor."""
        t = torch.as_tensor(ensemble_batch_next_obs, dtype=torch.float32)
        was_np = not isinstance(ensemble_batch_next_obs, torch.Tensor)
        if not t.is_cuda:
            t = t.cuda()
        pen = ops.mopo_penalty(t)
        return pen.cpu().numpy() if was_np else pen

    def get_dynamics_predict(self, origin_predict: Dict, mech: str, deterministic: bool = False):
        """Reference-compatible entry (fake_env.py:195-215): sample one elite member per env from a `query` result
        ([E,N,D] numpy mean / logvar) with `self.generator`; the gather + sampling runs on the K5 kernel."""
        variable = self.dynamics.get_variable_by_mech(mech)
        ensemble_mean, ensemble_logvar = origin_predict[variable]["mean"], origin_predict[variable]["logvar"]
        batch_size = ensemble_mean.shape[1]
        random_index = getattr(self.dynamics, mech).get_random_index(batch_size, self.generator)
        idx = torch.from_numpy(np.ascontiguousarray(random_index, dtype=np.uint8)).to(self.device)
        mean = torch.as_tensor(ensemble_mean, dtype=torch.float32).to(self.device)
        if deterministic:
            return ops.select_sample(mean, None, idx, None).cpu().numpy()
        noise = self.generator.normal(size=ensemble_mean.shape[1:]).astype(np.float32)
        logvar = torch.as_tensor(ensemble_logvar, dtype=torch.float32).to(self.device)
        return ops.select_sample(mean, logvar, idx, torch.from_numpy(noise).to(self.device)).cpu().numpy()

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> List[Any]:
        pass

    def get_attr(self, attr_name: str, indices=None) -> List[Any]:
        pass

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        pass
