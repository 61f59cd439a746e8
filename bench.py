#!/usr/bin/env python
"""bench.py -- VecFakeEnv transitions/sec (+ ensemble train member-samples/sec) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--precision fp32|f16]

Workload (BASELINE.json configs[1], "C2"): ExternalMaskTransition (P=O=5 sub-nets x E=7 members, 4x200, SiLU,
fixed 2-D causal mask, 5 elites) + learned PlainRewardMech, N = 100 000 fake envs per GPU, one step = one
VecFakeEnv step over all envs.  Weak scaling: every rank owns its own 100 000 envs and a full weight replica,
no data-path collective.  Synthetic data, reference init (seed 0).

JSON line (rank 0): `value` = device-resident transitions/s (whole job); `e2e` = the same metric through
`VecFakeEnv.step_async/step_wait` with host numpy actions in / obs,reward,done,infos out; `roofline` for the
dominant kernel; `cpu_baseline` = the torch-CPU port of the reference step (oracle/ref_port_torch.py) on this
box's host cores; `train` = BaseDynamics.train_step member-samples/s (bs 256 x 7 members).
`--impl reference` times only that CPU port (all host threads) and prints the same line shape.
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
FLOP_PER_TRANSITION_SELECTED = 5 * FLOP_SUBNET + FLOP_SUBNET  # 5 transition sub-nets + reward mech, drawn member only
METRIC = "vecfakeenv_transitions_per_sec"


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
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}  # fmt: skip
        return out


# ------------------------------------------------------------------------------------------------ CPU reference port
def build_port_env(n_envs, seed=0):
    import torch

    from oracle import cmrl_oracle as oc
    from oracle import ref_port_torch as rp

    rng = np.random.default_rng(seed)
    tr = {k: torch.from_numpy(v) for k, v in oc.init_plain_params(rng, O + A, H, 2, E, L, P=O, bounds_shape=(O, 1, 1, 1)).items()}
    rw = {k: torch.from_numpy(v) for k, v in oc.init_plain_params(rng, O + A, H, 2, E, L, bounds_shape=(1,)).items()}
    mask = torch.from_numpy(make_mask())
    elite = list(range(ELITE))
    init = rng.normal(size=(n_envs, O)).astype(np.float32)
    env = rp.RefPortEnv([("extmask", tr, elite, mask), ("reward", rw, elite, None)], n_envs, init,
                        termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), dtype=bool), seed=seed)  # fmt: skip
    return env, rng


def time_port(n_envs, steps, warmup):
    import torch

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    env, rng = build_port_env(n_envs)
    acts = rng.uniform(-1, 1, size=(n_envs, A)).astype(np.float32)
    for _ in range(warmup):
        env.step(acts)
    t0 = time.perf_counter()
    for _ in range(steps):
        env.step(acts)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt / steps * 1e3, threads


WORKLOAD = "C2: ExternalMaskTransition P=5 E=7 (5 elites) 4x200 SiLU fixed 2-D mask + PlainRewardMech, %d fake envs per GPU"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_envs
    value, ms, threads = time_port(n, args.steps, args.warmup)
    sample = "%d envs x %d steps of the C2 step (ExternalMask P=5,E=7 all-member forward + reward mech + numpy tail)" % (n, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % args.envs, "envs_per_gpu": args.envs, "precision": "fp32 (reference torch CPU)",
                   "sample": "each step advances a bounded sample of %d of the %d envs on the host cores" % (n, args.envs)},
        "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import torch
    import torch.distributed as dist

    import cmrl_b200 as cm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N = args.envs
    torch.manual_seed(0)
    np.random.seed(0)
    tr = cm.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    tr.set_input_mask(torch.from_numpy(make_mask()))
    rw = cm.PlainRewardMech(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    tr.set_elite_members(list(range(ELITE)))
    rw.set_elite_members(list(range(ELITE)))
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
    rng = np.random.default_rng(1000 + rank)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), dtype=bool),
               get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32), max_episode_steps=1000)  # fmt: skip
    env.rng_mode = "device"
    env.infos_mode = "shared"
    env.return_views = True  # step() returns the pinned observation ring buffer itself (valid for two steps), no 2 MB host copy
    if hasattr(env, "precision"):
        env.precision = args.precision
    env.seed(rank)
    env.reset()
    act_host = rng.uniform(-1, 1, size=(N, A)).astype(np.float32)
    act_dev = torch.from_numpy(act_host).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev, dtype=torch.float32)  # > 126 MB L2

    # ---- device-resident steps (value) ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()  # samples SM clocks / throttle reasons across all timed regions of this process
    step_fn = env.step_device_graphed if args.graph else env.step_device
    for _ in range(args.warmup):
        step_fn(act_dev)
    barrier()
    launches0 = cm._lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s in range(args.steps):
        flush.fill_(float(s))  # L2 flush between timed iterations
        ev[s][0].record()
        step_fn(act_dev)
        ev[s][1].record()
    barrier()
    launches = cm._lib.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = N * world * args.steps / (dev_ms * 1e-3)

    # ---- end to end through VecFakeEnv.step (host buffers) ----------------------------------------------
    for _ in range(args.warmup):
        env.step(act_host)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        env.step(act_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N * world * args.steps / float(t.item())
    h2d = N * A * 4
    d2h = N * O * 4 + N * 4 + 2 * N  # next_obs (for termination_fn; patched at reset rows it is also the returned obs), reward, done, trunc

    # ---- dominant kernel: K1 hidden layer (200->200) on the bucketed rows, timed alone ------------------
    roof = dominant_kernel_roofline(cm, env, tr, act_dev, args, flush)

    # ---- training step (second half of the metric) ---------------------------------------------------------
    train = time_train(cm, dyn, dev, args, world, barrier)
    # throughput point of the SURVEY 8d sweep (4 096 rows per member): the tensor-core training GEMMs take over here
    train["bs4096"] = time_train(cm, dyn, dev, args, world, barrier, B=4096, n_batches=12, with_cpu=False)
    clocks = sampler.stop()

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            v, ms, threads = time_port(args.ref_envs, 2, 1)
            cpu = {"value": v, "unit": "transitions/s", "cores": threads, "kind": "port",
                   "sample": "%d envs x 2 steps of the same C2 step on the torch-CPU port of the reference" % args.ref_envs}  # fmt: skip
        line = {
            "metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": roof.pop("dtype"), "data": "synthetic",
            "config": {
                "workload": WORKLOAD % N, "evaluation": "selected-member (drawn member only), device Philox draws",
                "envs_per_gpu": N, "precision": args.precision, "l2": "flushed (256 MB write) between timed steps",
                "flop_per_transition": FLOP_PER_TRANSITION_SELECTED,
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "transitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "VecFakeEnv.step_async/step_wait, rng_mode=device, infos_mode=shared, return_views=True, host termination_fn"},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "train": train,
        }  # fmt: skip
        print(json.dumps(line))
    if world > 1:
        # CUDA graphs that captured the NCCL all-reduce (data-parallel training) still reference the communicator and can
        # stall an orderly destroy_process_group(): synchronise, flush and leave
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def dominant_kernel_roofline(cm, env, tr, act_dev, args, flush):
    import torch

    pk, src = peaks()
    N = env.num_envs
    if getattr(env, "precision", "fp32") == "f16":
        # tensor-core tier: the fused whole-MLP kernel of the transition mech (5 sub-nets x 4x200 + head per env)
        idx, noise = cm.ops.draw_member_noise(1, 1, N, env._elite_u8("transition"), O)
        obs = env._obs_dev
        for _ in range(3):
            env.predict_selected(tr, obs, act_dev, idx, noise)
        reps = max(3, args.steps)
        perm, offsets = cm.ops.member_bucket(idx, E)
        mn, mx = tr._bounds()
        pack = tr.packed_f16()
        # flush, event, kernel, event -- all reps queued before the one synchronize, so that an event pair brackets the
        # kernel and not the host-side launch latency of an idle GPU
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        torch.cuda.synchronize()
        for a, c in evs:
            flush.fill_(1.0)
            a.record()
            cm.ops.mlp_rollout_f16(pack, O, E, obs, act_dev, perm, offsets, H, O, cm.ops.HEAD_PARALLEL, True,
                                   cm.ops.ACT_SILU, True, mn, mx, noise)
            c.record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(c) for a, c in evs) / reps
        flops = float(FLOP_SUBNET) * O * N
        achieved = flops / (ms * 1e-3) / 1e12
        return {
            "dtype": "f16", "kernel": "mlp_rollout_f16_v5_kernel (fused 4x200 MLP + head + sampling, 5 sub-nets, tcgen05 kind::f16)",
            "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_tflops"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch at N = 100000 (profiles/r1_ncu_mlp_f16_v5.csv)
            "traffic": 11.70e6 if N == 100000 else None, "traffic_unit": "bytes per launch (ncu --set full)",
            "peak_source": src + " bf16 burst",
            "flops_per_launch": flops, "ms_per_launch": ms,
            "note": "algorithmic FLOPs = 243200 per (sub-net, env) x 5 sub-nets x N envs (drawn member only, padding not counted)",
        }  # fmt: skip
    # fp32 tier: grouped SGEMM hidden layer on the bucketed rows of all P sub-nets
    idx, _ = cm.ops.draw_member_noise(1, 1, N, env._elite_u8("transition"), O, want_noise=False)
    perm, offsets = cm.ops.member_bucket(idx, E)
    h = torch.randn(O, N, H, device=act_dev.device)
    w, b = tr.hidden_layers[1][0].weight.detach(), tr.hidden_layers[1][0].bias.detach()
    for _ in range(3):
        cm.ops.ens_linear_fwd(h, w, b, cm.ops.ACT_SILU, row_offsets=offsets)
    reps = max(3, args.steps)
    tot = 0.0
    for r in range(reps):
        flush.fill_(1.0)
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


class _ReplayBufferFields:
    """The SB3 `ReplayBuffer` fields `BaseDynamics.dataset_split` reads (base_dynamics.py:131-138), n_envs = 1."""

    def __init__(self, obs, act, next_obs, rew, done):
        n = obs.shape[0]
        self.buffer_size, self.full, self.pos = n, True, 0
        self.observations, self.next_observations = obs[:, None, :], next_obs[:, None, :]
        self.actions, self.rewards, self.dones = act[:, None, :], rew.reshape(n, 1), done.reshape(n, 1)


def time_train(cm, dyn_unused, dev, args, world, barrier, B=256, n_batches=200, with_cpu=True):
    """Second half of the metric: ensemble train member-samples/s.  C4-shaped: PlainTransition E=7 (4x200, SiLU, O=5,
    A=1), bs 256 per member, fp32 tier, one device-resident epoch (on-device bootstrap gather + CUDA-graph replay of
    the optimiser step) over a synthetic training set; replicas per GPU (gradient all-reduce is exercised by
    `enable_data_parallel`, not timed here)."""
    import torch
    import torch.distributed as dist

    Bg = B * world  # data parallel: the global batch is sharded by rows, B rows per member stay on every GPU (weak scaling)
    N = int(Bg * n_batches / 0.8)
    rng = np.random.default_rng(7)
    obs = rng.normal(size=(N, O)).astype(np.float32)
    act = rng.uniform(-1, 1, size=(N, A)).astype(np.float32)
    M = (rng.normal(size=(O, O)) / np.sqrt(O)).astype(np.float32)
    K = rng.normal(size=(A, O)).astype(np.float32)
    nxt = (obs + 0.1 * np.tanh(obs @ M + act @ K) + 0.01 * rng.normal(size=(N, O))).astype(np.float32)
    rew = (np.cos(obs[:, 0]) - 0.1 * (act**2).sum(-1)).astype(np.float32)
    buf = _ReplayBufferFields(obs, act, nxt, rew, np.zeros(N, bool))
    torch.manual_seed(1)
    tr = cm.PlainTransition(O, A, ensemble_num=E, elite_num=ELITE, hid_size=H, activation_fn_cfg=SILU, device=dev)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False)
    if world > 1:
        dyn.enable_data_parallel()  # one NCCL all-reduce (sum) of the flat gradient arena per step, inside the captured graph
    train_it, _ = dyn.dataset_split(buf, 0.2, Bg)
    steps = len(train_it)
    dyn.train_epoch_device(train_it, "transition")  # warm-up epoch: uploads the set, captures the graph
    barrier()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loss = dyn.train_epoch_device(train_it, "transition")
    c.record()
    barrier()
    t = torch.tensor([a.elapsed_time(c)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    out = {"metric": "ensemble_train_member_samples_per_sec", "value": E * B * world / (ms * 1e-3), "unit": "member-samples/s",
           "ms_per_step": ms, "steps_per_epoch": steps, "epoch_mean_nll": loss,
           "config": "C4-shaped: PlainTransition E=7 4x200 SiLU O=5 A=1, bs %d per member per GPU, fp32-grade tier "
                     "(GEMM tier %s: %s), device-resident epoch (on-device bootstrap gather + CUDA-graph replay), %s"
                     % (B, cm.ops.TRAIN_GEMM_TIER, "tcgen05 3xTF32 forward/dx" + (", split-K dw" if B >= 2048 else ", FFMA dw") if B >= cm.ops.TRAIN_GEMM_TC_MIN_ROWS
                        else "FFMA: latency bound at this batch size",
                        "single GPU" if world == 1 else "data parallel over %d GPUs: global batch %d rows per member sharded by rows, "
                        "one NCCL all-reduce of the flat gradient arena per step" % (world, Bg))}  # fmt: skip
    if with_cpu and not args.no_cpu_baseline and int(os.environ.get("RANK", "0")) == 0:
        out["cpu_baseline"] = time_port_train(B)
    return out


def time_port_train(B, n_steps=30):
    """The reference's training step (NLL + backward + torch.optim.Adam) on the host cores (torch-CPU port)."""
    import torch

    from oracle import cmrl_oracle as oc
    from oracle import ref_port_torch as rp

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    params = {k: torch.from_numpy(v).requires_grad_(not k.endswith("logvar"))
              for k, v in oc.init_plain_params(rng, O + A, H, 2 * O, E, L, bounds_shape=(1, O)).items()}  # fmt: skip
    optim = torch.optim.Adam([v for v in params.values() if v.requires_grad], lr=1e-4, weight_decay=1e-5, eps=1e-8)
    obs = torch.randn(E, B, O)
    act = torch.rand(E, B, A) * 2 - 1
    tgt = obs + 0.1 * torch.randn(E, B, O)
    for _ in range(3):
        rp.train_batch(params, optim, obs.clone(), act, tgt, rp._act("silu"))
    t0 = time.perf_counter()
    for _ in range(n_steps):
        rp.train_batch(params, optim, obs.clone(), act, tgt, rp._act("silu"))
    dt = (time.perf_counter() - t0) / n_steps
    return {"value": E * B / dt, "unit": "member-samples/s", "cores": threads, "kind": "port",
            "sample": "%d optimiser steps of bs %d x %d members on the torch-CPU port" % (n_steps, B, E)}  # fmt: skip


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--envs", type=int, default=100_000)
    ap.add_argument("--ref-envs", type=int, default=10_000)
    ap.add_argument("--precision", default="f16", choices=["fp32", "f16"])
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch the step kernel by kernel instead of replaying its CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU port timing (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
