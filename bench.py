#!/usr/bin/env python
"""bench.py -- VecFakeEnv transitions/sec and ensemble train member-samples/sec on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--precision fp32|f16]

Headline workload (BASELINE.json configs[1], "C2"): ExternalMaskTransition (P=O=5 sub-nets x E=7 members, 4x200, SiLU,
fixed 2-D causal mask, 5 elites) + learned PlainRewardMech, N = 100 000 fake envs per GPU, one step = one VecFakeEnv step
over all envs.  Weak scaling: every rank owns its own envs and a full weight replica, no data-path collective.

JSON line (rank 0):
  value        device-resident transitions/s (whole job), inputs resident in HBM
  e2e          the same metric through VecFakeEnv.step_async/step_wait with host numpy in / out
  roofline     the fused rollout kernel alone (algorithmic FLOPs / CUDA-event time / measured bf16 peak)
  cpu_baseline the UNMODIFIED reference (baseline/_ref, installed by __graft_entry__.build) on this box's host cores
  train        BaseDynamics.train: member-samples/s at the reference batch size (256 rows per member), at 4 096 rows, and one
               epoch over C4's 10 M synthetic transitions (8 M training rows resident in HBM), each with a roofline
  configs      C1 (1 000 envs), C3 (transition + reward + termination, 125 000 envs per GPU), C5 (wide: O=100, A=20, E=16,
               H=512, 8 192 envs), the fp32-grade tier and the default SB3 contract (numpy RNG, per-env info dicts, copies)
  gpu_eager_reference   the unmodified reference on cuda:0 (torch eager): the same-box bar of SURVEY 2a
`--impl reference` times only the reference on the host cores and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

O, A, E, ELITE, H, L = 5, 1, 7, 5, 200, 4
SILU = {"_target_": "torch.nn.SiLU"}
FLOP_SUBNET = 2 * ((O + A) * H + 3 * H * H + H * 2)  # 243 200 per (net, sample), SURVEY 8d
FLOP_PLAIN = 2 * ((O + A) * H + 3 * H * H + H * 2 * O)  # 246 400: PlainTransition, head 2*O
FLOP_PER_TRANSITION_SELECTED = 5 * FLOP_SUBNET + FLOP_SUBNET  # 5 transition sub-nets + reward mech, drawn member only
FLOP_TRAIN_PLAIN = 3 * FLOP_PLAIN - 2 * (O + A) * H  # 736 800 per (member, sample): forward + dW + dX (no dX for layer 1)
METRIC = "vecfakeenv_transitions_per_sec"
WORKLOAD = "C2: ExternalMaskTransition P=5 E=7 (5 elites) 4x200 SiLU fixed 2-D mask + PlainRewardMech, %d fake envs per GPU"


def make_mask():
    rng = np.random.default_rng(0)
    mask = (rng.uniform(size=(O, O + A)) < 0.5).astype(np.float32)
    mask[np.arange(O), np.arange(O)] = 1.0
    return mask


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)  # fmt: skip
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 6:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s in sm if s > 0.5 * max(mx)] or sm
            out = {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}  # fmt: skip
        return out


# ------------------------------------------------------------------------------------------------ reference legs (baselines)
# The only places of this file that touch oracle/: the stubs that let the UNMODIFIED reference import (oracle/refshim.py)
# and, if the reference was not installed, its torch-CPU port.  Never called by the measured (native) code paths.
def _reference():
    from oracle import refshim

    if not refshim.available():
        return None
    try:
        return refshim.load_reference()
    except ImportError as e:  # an incomplete install must not take the bench line down: the port is timed and `kind` says so
        print("bench: reference install unusable (%s); timing the torch-CPU port instead" % e, file=sys.stderr)
        return None


def build_reference_c2(ref, n_envs, device="cpu", seed=0):
    import torch

    torch.manual_seed(seed)
    np.random.seed(seed)
    tr = ref.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=device)
    tr.set_input_mask(torch.from_numpy(make_mask()).to(device))
    rw = ref.PlainRewardMech(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=device)
    tr.set_elite_members(list(range(ELITE)))
    rw.set_elite_members(list(range(ELITE)))
    dyn = ref.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
    rng = np.random.default_rng(seed)
    env = ref.VecFakeEnv(n_envs, None, None)
    env.set_up(dyn, reward_fn=None, termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), dtype=bool),
               get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32), max_episode_steps=1000)  # fmt: skip
    env.seed(seed)
    env.reset()
    return env, rng


def time_reference_rollout(n_envs, steps, warmup, device="cpu"):
    """transitions/s of the reference's own VecFakeEnv.step (fake_env.py:94-142) for the C2 dynamics."""
    import torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ref = _reference()
    if ref is None:  # reference not installed (baseline/_ref missing): the torch-CPU port of the same step
        from oracle import cmrl_oracle as oc
        from oracle import ref_port_torch as rp

        rng = np.random.default_rng(0)
        tr = {k: torch.from_numpy(v) for k, v in oc.init_plain_params(rng, O + A, H, 2, E, L, P=O, bounds_shape=(O, 1, 1, 1)).items()}
        rw = {k: torch.from_numpy(v) for k, v in oc.init_plain_params(rng, O + A, H, 2, E, L, bounds_shape=(1,)).items()}
        env = rp.RefPortEnv([("extmask", tr, list(range(ELITE)), torch.from_numpy(make_mask())), ("reward", rw, list(range(ELITE)), None)],
                            n_envs, rng.normal(size=(n_envs, O)).astype(np.float32),
                            termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), dtype=bool), seed=0)  # fmt: skip
        kind = "port"
    else:
        env, rng = build_reference_c2(ref, n_envs, device)
        kind = "reference"
    acts = rng.uniform(-1, 1, size=(n_envs, A)).astype(np.float32)
    for _ in range(warmup):
        env.step(acts)
    if device != "cpu":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        env.step(acts)
    if device != "cpu":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt / steps * 1e3, threads, kind


def time_reference_train(B, n_batches=50, device="cpu"):
    """member-samples/s of the reference's BaseDynamics.train (base_dynamics.py:173-188) on PlainTransition, bs B."""
    import torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ref = _reference()
    if ref is None:
        return None
    torch.manual_seed(1)
    tr = ref.PlainTransition(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=device)
    dyn = ref.PlainEnsembleDynamics(tr, learned_reward=False)
    buf = _ReplayBufferFields(*synthetic_transitions(int(B * n_batches / 0.8), 7))
    train_it, _ = dyn.dataset_split(buf, 0.2, B)
    steps = len(train_it)
    if device != "cpu":  # one warm-up pass of a few batches is enough for cuBLAS to initialise
        dyn.train(dyn.dataset_split(_ReplayBufferFields(*synthetic_transitions(4 * B, 8)), 0.2, B)[0], "transition")
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    dyn.train(train_it, "transition")
    if device != "cpu":
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return {"value": E * B / dt, "unit": "member-samples/s", "ms_per_step": dt * 1e3, "cores": threads, "kind": "reference",
            "sample": "BaseDynamics.train of the unmodified reference over %d batches of %d rows x %d members on %s" % (steps, B, E, device)}  # fmt: skip


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_envs
    value, ms, threads, kind = time_reference_rollout(n, args.steps, args.warmup)
    sample = "%d envs x %d steps of the C2 step (VecFakeEnv.step of the %s: ExternalMask P=5,E=7 all-member forward + reward mech + " \
             "numpy tail + per-env info dicts)" % (n, args.steps, "unmodified reference" if kind == "reference" else "torch-CPU port")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.envs, "envs_per_gpu": args.envs, "precision": "fp32 (reference torch CPU)",
                   "sample": "each step advances a bounded sample of %d of the %d envs on the host cores" % (n, args.envs)},
        "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    if not args.no_train:
        line["train"] = time_reference_train(256, 30)
    print(json.dumps(line))


def gpu_eager_reference(dev, n_envs, steps):
    """The unmodified reference with device=cuda (torch eager: cuBLAS bmm + elementwise launches): SURVEY 2a's same-box bar."""
    out = {}
    try:
        v, ms, _, kind = time_reference_rollout(n_envs, steps, 2, device=dev)
        out["rollout"] = {"value": v, "unit": "transitions/s", "ms_per_step": ms, "envs": n_envs, "kind": kind,
                          "api": "reference VecFakeEnv.step, models on %s, host numpy in / out" % dev}  # fmt: skip
    except Exception as e:  # noqa: BLE001 -- a baseline leg must never take the bench line down
        out["rollout"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    try:
        out["train_bs256"] = time_reference_train(256, 60, device=dev)
        out["train_bs4096"] = time_reference_train(4096, 12, device=dev)
    except Exception as e:  # noqa: BLE001
        out["train_error"] = "%s: %s" % (type(e).__name__, str(e)[:200])
    return out


# ------------------------------------------------------------------------------------------------ synthetic data
def synthetic_transitions(N, seed):
    """obs ~ N(0,1), action ~ U(-1,1), next_obs = obs + 0.1 tanh(obs M + action K) + 0.01 eps."""
    rng = np.random.default_rng(seed)
    obs = rng.standard_normal(size=(N, O), dtype=np.float32)
    act = (rng.random(size=(N, A), dtype=np.float32) * 2 - 1).astype(np.float32)
    M = (rng.normal(size=(O, O)) / np.sqrt(O)).astype(np.float32)
    K = rng.normal(size=(A, O)).astype(np.float32)
    nxt = (obs + 0.1 * np.tanh(obs @ M + act @ K) + 0.01 * rng.standard_normal(size=(N, O), dtype=np.float32)).astype(np.float32)
    rew = (np.cos(obs[:, 0]) - 0.1 * (act**2).sum(-1)).astype(np.float32)
    return obs, act, nxt, rew, np.zeros(N, bool)


class _ReplayBufferFields:
    """The SB3 `ReplayBuffer` fields `BaseDynamics.dataset_split` reads (base_dynamics.py:131-138), n_envs = 1."""

    def __init__(self, obs, act, next_obs, rew, done):
        n = obs.shape[0]
        self.buffer_size, self.full, self.pos = n, True, 0
        self.observations, self.next_observations = obs[:, None, :], next_obs[:, None, :]
        self.actions, self.rewards, self.dones = act[:, None, :], rew.reshape(n, 1), done.reshape(n, 1)


# ------------------------------------------------------------------------------------------------ native arm
class _Ctx:
    """Process-wide plumbing of the native arm: device, ranks, barrier, L2 flush buffer, max-over-ranks reduction."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = "cuda:%d" % self.local_rank
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
        self.flush = torch.empty(256 * 1024 * 1024 // 4, device=self.dev, dtype=torch.float32)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def make_env(cm, ctx, dyn, N, obs_dim, precision, throughput_knobs=True, deterministic=False, host_termination=True):
    rng = np.random.default_rng(1000 + ctx.rank)
    env = cm.VecFakeEnv(N, None, None)
    kw = {}
    if host_termination:
        kw["termination_fn"] = lambda n, o, a: np.zeros((n.shape[0], 1), dtype=bool)
    env.set_up(dyn, get_init_obs_fn=lambda n: rng.normal(size=(n, obs_dim)).astype(np.float32), max_episode_steps=1000, **kw)
    if throughput_knobs:
        env.rng_mode = "device"      # Philox draws on the GPU (same distributions as numpy's, not the same numbers)
        env.infos_mode = "shared"    # one shared info dict on steps without a finished episode
        env.return_views = True      # step() hands out the pinned ring buffer instead of copying 2 MB per step
    env.precision = precision
    env.deterministic = deterministic
    env.seed(ctx.rank)
    env.reset()
    return env, rng


def time_device_steps(cm, ctx, env, act_dev, steps, warmup, graph=True):
    """Device-resident steps: CUDA events around every step, L2 flushed in between, max over ranks."""
    torch = ctx.torch
    step_fn = env.step_device_graphed if graph else env.step_device
    for _ in range(warmup):
        step_fn(act_dev)
    ctx.barrier()
    l0 = cm._lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s in range(steps):
        ctx.flush.fill_(float(s))
        ev[s][0].record()
        step_fn(act_dev)
        ev[s][1].record()
    ctx.barrier()
    launches = cm._lib.launch_count() - l0
    ms = ctx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    return env.num_envs * ctx.world * steps / (ms * 1e-3), ms / steps, int(launches)


def time_host_steps(ctx, env, act_host, steps, warmup):
    """VecFakeEnv.step with host numpy actions in and obs / reward / done / infos out (wall clock, max over ranks)."""
    for _ in range(warmup):
        env.step(act_host)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        env.step(act_host)
    ctx.torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    return env.num_envs * ctx.world * steps / dt, dt / steps * 1e3


def run_native(args):
    import torch

    import cmrl_b200 as cm

    ctx = _Ctx()
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    N = args.envs
    torch.manual_seed(0)
    np.random.seed(0)
    tr = cm.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    tr.set_input_mask(torch.from_numpy(make_mask()))
    rw = cm.PlainRewardMech(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    tr.set_elite_members(list(range(ELITE)))
    rw.set_elite_members(list(range(ELITE)))
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
    env, rng = make_env(cm, ctx, dyn, N, O, args.precision)
    act_host = rng.uniform(-1, 1, size=(N, A)).astype(np.float32)
    act_dev = torch.from_numpy(act_host).to(dev)

    sampler = ClockSampler(ctx.local_rank)
    sampler.start()  # samples SM clocks / throttle reasons across all timed regions of this process
    # ---- headline: device-resident steps (value), then end to end through VecFakeEnv.step (e2e) --------------------------
    value, ms_step, launches = time_device_steps(cm, ctx, env, act_dev, args.steps, args.warmup, args.graph)
    e2e_value, e2e_ms = time_host_steps(ctx, env, act_host, args.steps, args.warmup)
    h2d = N * A * 4
    d2h = N * O * 4 + N * 4 + 2 * N  # next_obs (for termination_fn; patched at reset rows it is also the returned obs), reward, done, trunc
    roof = dominant_kernel_roofline(cm, ctx, env, tr, act_dev, args)

    # ---- second half of the metric: training ---------------------------------------------------------------------------
    train = None
    if not args.no_train:
        train = time_train(cm, ctx, args, B=256, n_batches=200)
        train["roofline"] = train_kernel_roofline(cm, ctx, 256)
        train["bs4096"] = time_train(cm, ctx, args, B=4096, n_batches=24)
        train["bs4096"]["roofline"] = train_kernel_roofline(cm, ctx, 4096)
        if not args.quick or args.c4:
            train["c4"] = time_train_c4(cm, ctx, args)
        if not args.quick and not args.skip_c5:
            train["c5"] = time_train_c5(cm, ctx)

    # ---- the other configs of BASELINE.json + tiers / contracts -----------------------------------------------------------
    configs = {}
    if not args.quick:
        configs = other_configs(cm, ctx, args, dyn, act_host)
    clocks = sampler.stop()

    if rank == 0:
        cpu = None
        eager = None
        if not args.no_cpu_baseline:
            v, ms, threads, kind = time_reference_rollout(args.ref_envs, 5, 1)
            cpu = {"value": v, "unit": "transitions/s", "cores": threads, "kind": kind, "ms_per_step": ms,
                   "sample": "%d envs x 5 steps of the same C2 step through the %s VecFakeEnv.step on the host cores"
                             % (args.ref_envs, "unmodified reference's" if kind == "reference" else "torch-CPU port's")}  # fmt: skip
            if train is not None and world == 1:
                train["cpu_baseline"] = time_reference_train(256, 30)
            if world == 1 and not args.quick:
                eager = gpu_eager_reference(dev, N, 3)
        line = {
            "metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": roof.pop("dtype"), "data": "synthetic",
            "config": {
                "workload": WORKLOAD % N, "evaluation": "selected-member (drawn member only), device Philox draws",
                "envs_per_gpu": N, "precision": args.precision, "l2": "flushed (256 MB write) between timed steps",
                "flop_per_transition": FLOP_PER_TRANSITION_SELECTED,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "transitions/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "VecFakeEnv.step_async/step_wait, rng_mode=device, infos_mode=shared, return_views=True, host termination_fn"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "train": train,
            "configs": configs,
            "gpu_eager_reference": eager,
        }  # fmt: skip
        print(json.dumps(line))
    if world > 1:
        # CUDA graphs that captured the NCCL all-reduce (data-parallel training) still reference the communicator and can
        # stall an orderly destroy_process_group(): synchronise, flush and leave
        ctx.dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def dominant_kernel_roofline(cm, ctx, env, tr, act_dev, args, flop_subnet=FLOP_SUBNET, nets_per_env=O, kernel=None):
    """The fused rollout kernel of the transition mech alone: CUDA events around each launch, L2 flushed before it."""
    torch = ctx.torch
    pk, src = peaks()
    N = env.num_envs
    D = tr._head_dim
    if getattr(env, "precision", "fp32") == "f16":
        idx, noise = cm.ops.draw_member_noise(1, 1, N, env._elite_u8("transition"), D)
        obs = env._obs_dev
        perm, offsets = cm.ops.member_bucket(idx, tr.ensemble_num)
        bucket = (perm, offsets)
        for _ in range(3):
            env.predict_selected(tr, obs, act_dev, idx, noise, bucket)
        reps = max(3, min(args.steps, 20))
        # flush, event, kernel, event -- all reps queued before the one synchronize, so that an event pair brackets the
        # kernel and not the host-side launch latency of an idle GPU
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        torch.cuda.synchronize()
        for a, c in evs:
            ctx.flush.fill_(1.0)
            a.record()
            env.predict_selected(tr, obs, act_dev, idx, noise, bucket)
            c.record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(c) for a, c in evs) / reps
        flops = float(flop_subnet) * nets_per_env * N
        achieved = flops / (ms * 1e-3) / 1e12
        return {
            "dtype": "f16", "kernel": kernel or "mlp_rollout_f16_v5_kernel (fused 4x200 MLP + head + sampling, 5 sub-nets, tcgen05 kind::f16)",
            "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_tflops"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch at N = 100000 (profiles/r1_ncu_mlp_f16_v5.csv)
            "traffic": 11.70e6 if (N == 100000 and nets_per_env == O and kernel is None) else None, "traffic_unit": "bytes per launch (ncu --set full)",
            "peak_source": src + " bf16 burst",
            "flops_per_launch": flops, "ms_per_launch": ms,
            "note": "algorithmic FLOPs = %d per (sub-net, env) x %d sub-nets x N envs (drawn member only, padding not counted)" % (flop_subnet, nets_per_env),
        }  # fmt: skip
    # fp32 tier: grouped SGEMM hidden layer on the bucketed rows of all P sub-nets
    idx, _ = cm.ops.draw_member_noise(1, 1, N, env._elite_u8("transition"), O, want_noise=False)
    perm, offsets = cm.ops.member_bucket(idx, E)
    h = torch.randn(O, N, H, device=act_dev.device)
    w, b = tr.hidden_layers[1][0].weight.detach(), tr.hidden_layers[1][0].bias.detach()
    for _ in range(3):
        cm.ops.ens_linear_fwd(h, w, b, cm.ops.ACT_SILU, row_offsets=offsets)
    reps = max(3, min(args.steps, 10))
    tot = 0.0
    for r in range(reps):
        ctx.flush.fill_(1.0)
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        cm.ops.ens_linear_fwd(h, w, b, cm.ops.ACT_SILU, row_offsets=offsets)
        c.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(c)
    ms = tot / reps
    flops = 2.0 * H * H * N * O
    achieved = flops / (ms * 1e-3) / 1e12
    return {
        "dtype": "f32", "kernel": "grouped_sgemm_kernel<fwd> (K1 hidden layer 200x200, P=5 sub-nets, ragged by member)",
        "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
        "frac": achieved / pk["bf16_tflops"], "traffic": None, "peak_source": src + " bf16 burst",
        "flops_per_launch": flops, "ms_per_launch": ms,
        "note": "fp32 CUDA-core FFMA tier (1e-5 parity); fraction is against the bf16 tensor peak",
    }  # fmt: skip


def time_train(cm, ctx, args, B=256, n_batches=200):
    """Ensemble train member-samples/s.  C4-shaped: PlainTransition E=7 (4x200, SiLU, O=5, A=1), B rows per member per GPU,
    fp32-grade tier, one device-resident epoch (on-device bootstrap gather + CUDA-graph replay of the optimiser step) over a
    synthetic training set.  N > 1: data parallel, the global batch (B x world rows per member) is sharded by rows and the
    flat gradient arena is all-reduced once per step."""
    torch = ctx.torch
    world, dev = ctx.world, ctx.dev
    Bg = B * world
    buf = _ReplayBufferFields(*synthetic_transitions(int(Bg * n_batches / 0.8), 7))
    torch.manual_seed(1)
    tr = cm.PlainTransition(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False)
    if world > 1:
        dyn.enable_data_parallel()
    train_it, _ = dyn.dataset_split(buf, 0.2, Bg)
    steps = len(train_it)
    dyn.train_epoch_device(train_it, "transition")  # warm-up epoch: uploads the set, captures the graph
    ctx.barrier()
    l0 = cm._lib.launch_count()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loss = dyn.train_epoch_device(train_it, "transition")
    c.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(a.elapsed_time(c)) / steps
    fused = cm.ops.fused_train_plan(tr, E, B) is not None
    return {"metric": "ensemble_train_member_samples_per_sec", "value": E * B * world / (ms * 1e-3), "unit": "member-samples/s",
            "ms_per_step": ms, "steps_per_epoch": steps, "epoch_mean_nll": loss, "gpu_launches_per_step": (cm._lib.launch_count() - l0) / max(steps, 1),
            "tflops_algorithmic": FLOP_TRAIN_PLAIN * E * B * world / (ms * 1e-3) / 1e12,
            "config": "C4-shaped: PlainTransition E=7 4x200 SiLU O=5 A=1, %d rows per member per GPU, %s, device-resident epoch "
                      "(on-device bootstrap gather + CUDA-graph replay of the whole optimiser step), %s"
                      % (B, "fused tcgen05 3xTF32 step (pack + forward/NLL/dz chain + all weight gradients: fp32-grade)" if fused
                         else "per-layer kernels (GEMM tier %s)" % cm.ops.TRAIN_GEMM_TIER,
                         "single GPU" if world == 1 else "data parallel over %d GPUs: global batch %d rows per member sharded by rows, "
                         "gradient sync per step: %s" % (world, Bg, "one kernel exchanging the flat gradient arena over NVLink peer memory, fused "
                                                          "with the Adam update (cmrl_allreduce_grads_adam)" if getattr(dyn, "_peer_sync", None) is not None
                                                          else "one ncclAllReduce of the flat gradient arena + the Adam kernel"))}  # fmt: skip


def train_kernel_roofline(cm, ctx, B):
    """The dominant kernel of a training step (train_fused_kernel: forward + NLL + dz chain on tcgen05) timed alone, cold L2."""
    torch = ctx.torch
    pk, src = peaks()
    torch.manual_seed(2)
    tr = cm.PlainTransition(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=ctx.dev)
    tr.flat_parameters()
    plan = cm.ops.fused_train_plan(tr, E, B)
    if plan is None:
        return None
    obs = torch.randn(E, B, O, device=ctx.dev)
    act = torch.rand(E, B, A, device=ctx.dev) * 2 - 1
    tgt = obs + 0.1 * torch.randn(E, B, O, device=ctx.dev)
    mn, mx = tr._bounds()
    for _ in range(3):
        plan.step(obs, act, tgt, mn, mx, 0.5, 1e-4)
    out = {}
    for name, fn in (("fused", plan.run_fused), ("dw", plan.run_dw)):
        reps = 10
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        torch.cuda.synchronize()
        for a, c in evs:
            ctx.flush.fill_(1.0)
            a.record()
            fn(obs, act, tgt, mn, mx, 0.5, 1e-4)
            c.record()
        torch.cuda.synchronize()
        out[name] = sum(a.elapsed_time(c) for a, c in evs) / reps
    # forward + dX flops run in train_fused_kernel, dW flops in train_dw_kernel (3F - 2*in*H per member-sample)
    f_fused = float(2 * FLOP_PLAIN - 2 * (O + A) * H) * E * B
    f_dw = float(FLOP_PLAIN) * E * B
    ach = f_fused / (out["fused"] * 1e-3) / 1e12
    return {
        "kernel": "train_fused_kernel (forward + Gaussian head + NLL + dz chain, tcgen05 kind::tf32, 3xTF32)", "bound": "tensor",
        "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
        # dram__bytes_read + dram__bytes_write per launch from the committed `ncu --set full` capture (profiles/r2_ncu_kernels.csv)
        "traffic": {256: 15.2e6, 4096: 180.8e6}.get(B), "traffic_unit": "bytes per launch (ncu --set full, profiles/r2_ncu_kernels.csv)",
        "tensor_pipe_active_pct_ncu": {256: 40.9, 4096: 40.4}.get(B),
        "peak_source": src + " bf16 burst", "flops_per_launch": f_fused, "ms_per_launch": out["fused"],
        "dw_kernel": {"kernel": "train_dw_kernel (all layers' weight / bias gradients)", "flops_per_launch": f_dw, "ms_per_launch": out["dw"],
                      "achieved": f_dw / (out["dw"] * 1e-3) / 1e12, "frac": f_dw / (out["dw"] * 1e-3) / 1e12 / pk["bf16_tflops"],
                      "traffic": {256: 12.2e6, 4096: 198.0e6}.get(B), "tensor_pipe_active_pct_ncu": {256: 15.7, 4096: 27.8}.get(B)},
        "note": "algorithmic FLOPs (2*MAC, padding not counted); 3xTF32 issues three kind::tf32 MMAs per product, each at half the "
                "bf16 rate, so 1/6 of the bf16 peak (%.0f TFLOP/s) is this tier's tensor-pipe ceiling" % (pk["bf16_tflops"] / 6),
    }  # fmt: skip


def time_train_c4(cm, ctx, args):
    """BASELINE.json configs[3]: one epoch over 10 M synthetic transitions (8 M training rows, 2 M hold-out) at the reference batch
    size, data parallel over the ranks; the training set and the [E, 8 M] bootstrap index table are resident in HBM."""
    torch = ctx.torch
    world, dev = ctx.world, ctx.dev
    N = args.c4_rows
    B = 256 * world  # reference batch size per member per GPU
    t0 = time.perf_counter()
    buf = _ReplayBufferFields(*synthetic_transitions(N, 11))
    torch.manual_seed(3)
    tr = cm.PlainTransition(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False)
    if world > 1:
        dyn.enable_data_parallel()
    train_it, val_it = dyn.dataset_split(buf, 0.2, B)
    t_host = time.perf_counter() - t0
    steps = len(train_it)
    ctx.barrier()
    t0 = time.perf_counter()
    loss = dyn.train_epoch_device(train_it, "transition")  # includes the upload of the set and the graph capture
    ctx.barrier()
    first = time.perf_counter() - t0
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loss = dyn.train_epoch_device(train_it, "transition")
    c.record()
    ctx.barrier()
    ms_epoch = ctx.max_over_ranks(a.elapsed_time(c))
    t0 = time.perf_counter()
    val = dyn.evaluate_member_mse(val_it, "transition")
    ctx.barrier()
    t_val = time.perf_counter() - t0
    ms = ms_epoch / steps
    return {"rows": N, "train_rows": train_it.num_stored, "val_rows": val_it.num_stored, "rows_per_member_per_gpu": 256, "steps_per_epoch": steps,
            "epoch_s": ms_epoch * 1e-3, "ms_per_step": ms, "value": E * 256 * world / (ms * 1e-3), "unit": "member-samples/s",
            "tflops_algorithmic": FLOP_TRAIN_PLAIN * E * 256 * world / (ms * 1e-3) / 1e12,
            "first_epoch_s_incl_upload_and_capture": first, "host_dataset_build_s": t_host, "holdout_eval_s": t_val,
            "epoch_mean_nll": loss, "holdout_mse_per_member": [float(x) for x in val.tolist()],
            "resident_bytes": int(train_it.num_stored * (2 * O + A) * 4 + E * train_it.num_stored * 4)}  # fmt: skip


def time_train_c5(cm, ctx, B=256, steps=4):
    """C5 training (BASELINE config 5: O = P = 100, A = 20, E = 16, H = 512, per-member masks: 1 600 nets, 1.36 G parameters).
    N > 1: NET SHARDING -- every rank trains its contiguous share of the 100 sub-networks on the whole batch, with no exchange
    at all during training (the sub-networks are independent); the same step is cut N ways, i.e. strong scaling."""
    torch = ctx.torch
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    O5, A5, E5, H5 = 100, 20, 16, 512
    p0, p1 = cm.net_shard_range(O5, world, rank)
    t0 = time.perf_counter()
    torch.manual_seed(5)
    rng5 = np.random.default_rng(0)
    m = cm.ExternalMaskTransition(O5, A5, ensemble_num=E5, elite_num=12, hid_size=H5, activation_fn_cfg=SILU, device=dev,
                                  subnet_range=(p0, p1) if world > 1 else None)  # fmt: skip
    mask = (rng5.uniform(size=(O5, E5, O5 + A5)) < 0.3).astype(np.float32)
    mask[np.arange(O5), :, np.arange(O5)] = 1.0
    m.set_input_mask(torch.from_numpy(mask))
    dyn = cm.PlainEnsembleDynamics(m, learned_reward=False)
    if world > 1:
        dyn.enable_net_sharding()
    obs = torch.randn(E5, B, O5, device=dev)
    act = torch.rand(E5, B, A5, device=dev) * 2 - 1
    tgt = m.target_columns(obs + 0.1 * torch.tanh(obs)).contiguous()
    for _ in range(2):
        dyn.train_step("transition", obs, act, tgt, want_loss=False)
    ctx.barrier()
    l0 = cm._lib.launch_count()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        dyn.train_step("transition", obs, act, tgt, want_loss=False)
    c.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(a.elapsed_time(c)) / steps
    flop = 3 * 2 * ((O5 + A5) * H5 + 3 * H5 * H5 + H5 * 2) * O5 * E5 * B
    out = {"workload": "ExternalMaskTransition O=P=100 A=20 E=16 H=512 per-member 3-D mask: one optimiser step, %d rows per member" % B,
           "value": E5 * B / (ms * 1e-3), "unit": "member-samples/s", "ms_per_step": ms, "tflops_algorithmic": flop / (ms * 1e-3) / 1e12,
           "gpu_launches_per_step": (cm._lib.launch_count() - l0) / steps, "train_gemm_tier": str(cm.ops.TRAIN_GEMM_TIER),
           "parallelism": "single GPU, all 100 sub-networks" if world == 1 else
                          "net sharding over %d GPUs (rank 0: sub-networks %d..%d), no gradient traffic; scaling: strong" % (world, p0, p1 - 1),
           "build_s": time.perf_counter() - t0}  # fmt: skip
    del dyn, m, obs, act, tgt
    torch.cuda.empty_cache()
    return out


def other_configs(cm, ctx, args, dyn_c2, act_host_c2):
    """C1, C3, C5 and the non-default tiers / contracts of C2 as named sub-records (short runs)."""
    torch = ctx.torch
    dev = ctx.dev
    out = {}
    steps = max(5, min(args.steps, 50))

    # ---- C1: PlainTransition + PlainRewardMech, 1 000 envs (conf/algorithm/mbpo.yaml:19) ----------------------------------------
    torch.manual_seed(4)
    tr = cm.PlainTransition(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    rw = cm.PlainRewardMech(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
    env, rng = make_env(cm, ctx, dyn, 1000, O, "f16")
    ah = rng.uniform(-1, 1, size=(1000, A)).astype(np.float32)
    v, ms, _ = time_device_steps(cm, ctx, env, torch.from_numpy(ah).to(dev), steps, 3)
    ev, ems = time_host_steps(ctx, env, ah, steps, 3)
    out["C1"] = {"workload": "PlainTransition + PlainRewardMech E=7 4x200 SiLU, 1 000 fake envs per GPU (launch / latency bound)",
                 "value": v, "ms_per_step": ms, "e2e": ev, "e2e_ms_per_step": ems, "unit": "transitions/s", "precision": "f16"}  # fmt: skip
    env_d, rng_d = make_env(cm, ctx, dyn, 1000, O, "fp32", throughput_knobs=False)
    ev, ems = time_host_steps(ctx, env_d, ah, steps, 3)
    out["C1"]["sb3_default_contract_fp32"] = {"e2e": ev, "e2e_ms_per_step": ems}

    # ---- C3: transition + reward + termination mechs, all learned, 125 000 envs per GPU (1 M over 8) ---------------------------
    n3 = args.c3_envs
    tm = cm.PlainTerminationMech(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    with torch.no_grad():  # the reference's rule is "non-zero => done" (SURVEY section 0): a head that fires rarely
        tm.mean_and_logvar.weight.zero_()
        tm.mean_and_logvar.bias.zero_()
    dyn3 = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw, learned_termination=True, termination_mech=tm)
    env3, rng3 = make_env(cm, ctx, dyn3, n3, O, "f16", deterministic=True, host_termination=False)
    ah3 = rng3.uniform(-1, 1, size=(n3, A)).astype(np.float32)
    v, ms, _ = time_device_steps(cm, ctx, env3, torch.from_numpy(ah3).to(dev), steps, 3)
    ev, ems = time_host_steps(ctx, env3, ah3, steps, 3)
    flop3 = FLOP_PLAIN + 2 * FLOP_SUBNET
    out["C3"] = {"workload": "PlainTransition + PlainRewardMech + PlainTerminationMech (all learned) E=7 4x200 SiLU, %d fake envs per GPU" % n3,
                 "value": v, "ms_per_step": ms, "e2e": ev, "e2e_ms_per_step": ems, "unit": "transitions/s", "precision": "f16",
                 "flop_per_transition": flop3, "tflops_algorithmic": flop3 * v / 1e12,
                 "frac_of_bf16_sustained": flop3 * v / ctx.world / 1e12 / peaks()[0]["bf16_tflops_sustained"]}  # fmt: skip
    del env3

    # ---- C2 on the fp32-grade tier, and through the default SB3 contract ---------------------------------------------------------
    N = args.envs
    env32, rng32 = make_env(cm, ctx, dyn_c2, N, O, "fp32")
    v, ms, _ = time_device_steps(cm, ctx, env32, torch.from_numpy(act_host_c2).to(dev), 5, 3)
    ev, ems = time_host_steps(ctx, env32, act_host_c2, 5, 3)
    out["C2_fp32_tier"] = {"workload": WORKLOAD % N, "value": v, "ms_per_step": ms, "e2e": ev, "e2e_ms_per_step": ems, "unit": "transitions/s",
                           "precision": "fp32 (the VecFakeEnv default, 1e-5 parity tier): train_fused_kernel in rollout mode, tcgen05 kind::tf32 "
                                        "with hi/lo split operands (3xTF32)",
                           "tflops_algorithmic": FLOP_PER_TRANSITION_SELECTED * v / 1e12}  # fmt: skip
    del env32
    envff, _ = make_env(cm, ctx, dyn_c2, N, O, "ffma")
    v, ms, _ = time_device_steps(cm, ctx, envff, torch.from_numpy(act_host_c2).to(dev), 3, 2)
    out["C2_ffma_tier"] = {"workload": WORKLOAD % N, "value": v, "ms_per_step": ms, "unit": "transitions/s",
                           "precision": "ffma (CUDA-core fp32 kernels, bit-identical to the all-member query path)"}  # fmt: skip
    del envff
    # MOPO-penalised rollouts (penalty_coeff != 0: every member's mean is needed, fake_env.py:186-193): all-member pass + drawn
    # member's sample on the tensor-core kernels; the CUDA-core tier runs the all-member forward of the per-layer kernels
    for prec, steps_p in (("f16", 10), ("fp32", 5), ("ffma", 2)):
        envp, _ = make_env(cm, ctx, dyn_c2, N, O, prec)
        envp.penalty_coeff = 1.0
        v, ms, _ = time_device_steps(cm, ctx, envp, torch.from_numpy(act_host_c2).to(dev), steps_p, 2)
        out["C2_mopo_penalty_" + prec] = {"workload": WORKLOAD % N + ", penalty_coeff = 1 (MOPO: all 7 members evaluated per env)",
                                          "value": v, "ms_per_step": ms, "unit": "transitions/s", "precision": prec}
        del envp
    for prec in ("f16", "fp32"):
        envd, _ = make_env(cm, ctx, dyn_c2, N, O, prec, throughput_knobs=False)
        ev, ems = time_host_steps(ctx, envd, act_host_c2, 3, 2)
        out["C2_sb3_default_contract_" + prec] = {
            "workload": WORKLOAD % N, "e2e": ev, "e2e_ms_per_step": ems, "unit": "transitions/s",
            "api": "VecFakeEnv.step defaults: rng_mode=numpy (the reference's host draws, bit-exact member choice), per-env info dicts, "
                   "returned arrays are copies; precision=%s" % prec}  # fmt: skip
        del envd

    # ---- C5: wide causal model, O = P = 100, A = 20, E = 16 (12 elites), H = 512, per-member mask, 8 192 envs ----------------------
    if not args.skip_c5:
        t0 = time.perf_counter()
        O5, A5, E5, H5, N5 = 100, 20, 16, 512, args.c5_envs
        torch.manual_seed(5)
        rng5 = np.random.default_rng(0)
        m5 = cm.ExternalMaskTransition(O5, A5, ensemble_num=E5, elite_num=12, hid_size=H5, activation_fn_cfg=SILU, device=dev)
        mask = (rng5.uniform(size=(O5, E5, O5 + A5)) < 0.3).astype(np.float32)
        mask[np.arange(O5), :, np.arange(O5)] = 1.0
        m5.set_input_mask(torch.from_numpy(mask))
        m5.set_elite_members(list(range(12)))
        dyn5 = cm.PlainEnsembleDynamics(m5, learned_reward=False)
        env5 = cm.VecFakeEnv(N5, None, None)
        env5.set_up(dyn5, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32),
                    termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
                    get_init_obs_fn=lambda n: rng5.normal(size=(n, O5)).astype(np.float32))  # fmt: skip
        env5.precision = "f16"
        env5.reset()
        act5 = torch.rand(N5, A5, device=dev) * 2 - 1
        flop5 = 2 * ((O5 + A5) * H5 + 3 * H5 * H5 + H5 * 2)
        r5 = dominant_kernel_roofline(cm, ctx, env5, m5, act5, args, flop_subnet=flop5, nets_per_env=O5,
                                      kernel="mlp_rollout_f16_v6_kernel (streamed weights, hidden 512, tcgen05 kind::f16)")
        r5.pop("dtype", None)
        out["C5"] = {"workload": "ExternalMaskTransition O=P=100 A=20 E=16 (12 elites) H=512 per-member 3-D mask, %d fake envs: selected-member "
                                 "prediction of the transition mech (100 sub-nets per env)" % N5,
                     "value": N5 / (r5["ms_per_launch"] * 1e-3), "unit": "transitions/s", "ms_per_step": r5["ms_per_launch"], "precision": "f16",
                     "roofline": r5, "build_s": time.perf_counter() - t0}  # fmt: skip
        del env5, dyn5, m5
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--envs", type=int, default=100_000)
    ap.add_argument("--ref-envs", type=int, default=10_000)
    ap.add_argument("--c3-envs", type=int, default=125_000)
    ap.add_argument("--c5-envs", type=int, default=8192)
    ap.add_argument("--c4-rows", type=int, default=10_000_000)
    ap.add_argument("--precision", default="f16", choices=["fp32", "f16"])
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch the step kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the reference timings on the host cores / torch eager (profiling runs)")
    ap.add_argument("--no-train", action="store_true", help="skip the training half of the metric")
    ap.add_argument("--quick", action="store_true", help="headline + short training legs only (no C1/C3/C4/C5 sub-records)")
    ap.add_argument("--skip-c5", action="store_true")
    ap.add_argument("--c4", action="store_true", help="with --quick: still run the 10 M-row C4 training epoch")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
