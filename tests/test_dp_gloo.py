"""World-size-2 `gloo` tests (CPU) of the data-parallel host logic: row sharding, gradient all-reduce
equivalence (sum of shard gradients with the global loss scale == full-batch gradient) and the bench reference arm
under torchrun semantics (rank 0 alone prints)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden, params_of, rel_err


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import cmrl_oracle as oc
    import cmrl_b200 as cm
    from cmrl_b200.dynamics import dp_allreduce_member_sums, dp_row_slice

    g = load_golden("plain_train.npz")
    p = params_of(g, "p0.")
    obs, act, target = g["obs0"], g["act0"], g["target0"]
    B = obs.shape[1]
    sl = dp_row_slice(B, world, rank)
    # local gradient of the shard, scaled by the GLOBAL element count (train_step's grad_rows)
    _, grads = oc.plain_nll_backward(p, obs[:, sl], act[:, sl], target[:, sl])
    local_rows = sl.stop - sl.start
    flat = torch.cat([torch.from_numpy(grads[k] * np.float32(local_rows / B)).reshape(-1) for k in sorted(grads)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    # a dynamics object picks up the group and shards consistently
    dyn = cm.PlainEnsembleDynamics(cm.PlainTransition(5, 1, hid_size=8), learned_reward=False)
    dyn.enable_data_parallel()
    ok = dyn.dp_world == world and dyn.dp_rank == rank
    # sharded evaluation (evaluate_member_mse under data parallel): per-member squared-error sums of this rank's
    # hold-out rows, one [E] all-reduce -> the full-set sums on every rank (ragged: 1001 rows over 2 ranks)
    rng = np.random.default_rng(3)
    pred, tgt = rng.normal(size=(7, 1001, 5)), rng.normal(size=(7, 1001, 5))
    rows = dp_row_slice(1001, world, rank)
    sums = torch.from_numpy(((pred[:, rows] - tgt[:, rows]) ** 2).sum(axis=(1, 2)))
    dp_allreduce_member_sums(sums, dyn.dp_group)
    eval_err = rel_err(sums.numpy(), ((pred - tgt) ** 2).sum(axis=(1, 2)))
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    same_elites = all(np.array_equal(np.argsort(g.numpy()), np.argsort(sums.numpy())) for g in gathered)
    # rollout replicas: parameters and elite set of rank 0 reach every rank
    if rank != 0:
        with torch.no_grad():
            for q in dyn.transition.parameters():
                q.fill_(float(rank))
        dyn.transition.set_elite_members([6, 5, 4, 3, 2])
    dyn.broadcast_parameters(src=0)
    flat_w = torch.cat([q.detach().reshape(-1) for q in dyn.transition.parameters()])
    all_w = [torch.zeros_like(flat_w) for _ in range(world)]
    dist.all_gather(all_w, flat_w)
    elites = [None] * world
    dist.all_gather_object(elites, list(dyn.transition.elite_members))
    same_replicas = all(torch.equal(w, all_w[0]) for w in all_w) and all(e == elites[0] for e in elites) and \
        not bool((flat_w == 1.0).all())  # rank 0's initialisation, not rank 1's fill value
    if rank == 0:
        _, full = oc.plain_nll_backward(p, obs, act, target)
        ref = torch.cat([torch.from_numpy(full[k]).reshape(-1) for k in sorted(full)])
        out.put((rel_err(flat.numpy(), ref.numpy()), ok and same_elites and eval_err < 1e-12 and same_replicas))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and err < 1e-5


def _shard_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import cmrl_b200 as cm

    P = 5  # dealt 3 + 2 over two ranks
    p0, p1 = cm.net_shard_range(P, world, rank)
    torch.manual_seed(100 + rank)
    shard = cm.ExternalMaskTransition(P, 1, hid_size=8, subnet_range=(p0, p1))
    mask = (np.arange(P * (P + 1)).reshape(P, P + 1) % 3 != 0).astype(np.float32)
    shard.set_input_mask(torch.from_numpy(mask))
    shard.set_elite_members([4, 2, 0, 1, 3])
    dyn = cm.PlainEnsembleDynamics(shard, learned_reward=False)
    dyn.enable_net_sharding()
    ok = shard.is_net_shard and tuple(shard._kernel_mask().shape) == (p1 - p0, P + 1) and np.array_equal(shard._kernel_mask().numpy(), mask[p0:p1])
    ok = ok and shard._grad_dims() == P and shard._head_dim == p1 - p0
    mine = {k: v.clone() for k, v in shard.state_dict().items()}
    full = dyn.gather_net_shards()
    ok = ok and not full.is_net_shard and dyn.transition is full and list(full.elite_members) == [4, 2, 0, 1, 3]
    ok = ok and np.array_equal(full.input_mask.numpy(), mask)
    for k, v in full.state_dict().items():  # this rank's slice of every tensor came through unchanged
        ok = ok and v.shape[0] == P and torch.equal(v[p0:p1], mine[k])
    flat = torch.cat([v.reshape(-1) for v in full.state_dict().values()])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    ok = ok and all(torch.equal(b, both[0]) for b in both)
    if rank == 0:
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_net_shards_gather_into_the_full_transition():
    """Net sharding: ranks hold disjoint sub-network ranges of an ExternalMaskTransition; the gather
    assembles identical full models everywhere (host logic on gloo; the arithmetic of shards is a GPU test)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_dp_row_slices_cover_ragged_batches():
    from cmrl_b200.dynamics import dp_row_slice

    for rows in (0, 1, 7, 256, 257):
        for world in (1, 2, 4, 8):
            covered = np.concatenate([np.arange(rows)[dp_row_slice(rows, world, r)] for r in range(world)])
            assert np.array_equal(covered, np.arange(rows))
            sizes = [dp_row_slice(rows, world, r).stop - dp_row_slice(rows, world, r).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_bench_reference_arm_prints_one_line_on_rank0_only():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--ref-envs", "256", "--no-train"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "transitions/s" and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] in ("reference", "port") and rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["n_gpus"] == 2
