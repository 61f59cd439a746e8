"""Data-parallel training over NCCL, run under torchrun on N GPUs of one node:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 profiles/tools/dp_nccl_check.py
Every rank trains the same PlainTransition on its row shard of identical ensemble batches (one all-reduce of the flat
gradient arena per step); rank 0 also trains a copy on the full batches without data parallelism.  Prints the relative
difference of the final weights (fp32 round-off expected) and the time per DP step."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cmrl_b200 as cm
from cmrl_b200.dynamics import dp_row_slice

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
_logf = open(os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "gpurun_out", "dp_rank%d.log" % rank), "w") if os.path.isdir(os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "gpurun_out")) else sys.stderr


def mark(*a):
    print(*a, file=_logf, flush=True)

torch.cuda.set_device(local)
dev = "cuda:%d" % local
dist.init_process_group("nccl", device_id=torch.device(dev))
SILU = {"_target_": "torch.nn.SiLU"}
E, B, O, A, steps = 7, 256, 5, 1, 20
SYNC = os.environ.get("CMRL_DP_SYNC", "auto")  # auto (= peer: NVLink peer-memory exchange fused with Adam) | peer | nccl


def build():
    torch.manual_seed(0)
    np.random.seed(0)
    tr = cm.PlainTransition(O, A, ensemble_num=E, hid_size=200, activation_fn_cfg=SILU, device=dev)
    return cm.PlainEnsembleDynamics(tr, learned_reward=False, optim_lr=1e-3)


rng = np.random.default_rng(1)
batches = [(torch.from_numpy(rng.normal(size=(E, B, O)).astype(np.float32)).to(dev),
            torch.from_numpy(rng.uniform(-1, 1, size=(E, B, A)).astype(np.float32)).to(dev)) for _ in range(steps)]
dyn = build()
dyn.enable_data_parallel(sync=SYNC)
sl = dp_row_slice(B, world, rank)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for obs, act in batches:
    tgt = obs + 0.1 * torch.tanh(obs)
    dyn.train_step("transition", obs[:, sl].contiguous(), act[:, sl].contiguous(), tgt[:, sl].contiguous(), grad_rows=B)
torch.cuda.synchronize(); dist.barrier()
dt = (time.perf_counter() - t0) / steps
flat_dp = dyn.transition.flat_parameters()[0].clone()
# all ranks must hold identical weights
ref0 = flat_dp.clone(); dist.broadcast(ref0, 0)
same = bool(torch.equal(ref0, flat_dp))
if rank == 0:
    single = build()
    for obs, act in batches:
        tgt = obs + 0.1 * torch.tanh(obs)
        single.train_step("transition", obs, act, tgt)
    flat_1 = single.transition.flat_parameters()[0]
    init = build().transition.flat_parameters()[0]
    upd_err = float(((flat_dp - init) - (flat_1 - init)).abs().max() / (flat_1 - init).abs().max())
    print("DP (gradient sync: %s), world %d: weights identical on all ranks: %s; weights vs single-process full batch: rel %.2e; "
          "update (20 Adam steps) rel %.2e; %.3f ms per DP step (bs %d per member, eager launches)"
          % ("peer memory" if dyn._peer_sync is not None else "nccl", world, same, float((flat_dp - flat_1).abs().max() / flat_1.abs().max()), upd_err, dt * 1e3, B))
# ---- device-resident epoch (graph-captured step with the all-reduce inside) vs the same epoch on one process ----------

class DuckReplayBuffer:  # the SB3 ReplayBuffer fields dataset_split reads (base_dynamics.py:131-138)
    def __init__(self, obs, act, next_obs, rew, done):
        n = obs.shape[0]
        self.buffer_size, self.full, self.pos = n, True, 0
        self.observations, self.next_observations = obs[:, None, :], next_obs[:, None, :]
        self.actions, self.rewards, self.dones = act[:, None, :], rew.reshape(n, 1), done.reshape(n, 1)


def epoch_weights(dp):
    dyn2 = build()
    if dp:
        dyn2.enable_data_parallel(sync=SYNC)
    r2 = np.random.default_rng(3)
    n = 256 * 12
    o = r2.normal(size=(n, O)).astype(np.float32); a = r2.uniform(-1, 1, size=(n, A)).astype(np.float32)
    buf = DuckReplayBuffer(o, a, (o + 0.1 * np.tanh(o)).astype(np.float32), np.zeros(n, np.float32), np.zeros(n, bool))
    real = np.random.default_rng
    seeds = iter(range(500, 600))
    np.random.default_rng = lambda seed=None: real(next(seeds) if seed is None else seed)
    try:
        it, _ = dyn2.dataset_split(buf, 0.2, 256)
        mark('split done, dp', dp)
        loss = dyn2.train_epoch_device(it, "transition")
        mark('epoch 1 done', loss)
        loss = dyn2.train_epoch_device(it, "transition")
        mark('epoch 2 done', loss)
    finally:
        np.random.default_rng = real
    return dyn2.transition.flat_parameters()[0].clone(), loss

mark('starting DP device epochs')
w_dp, loss_dp = epoch_weights(True)
mark('DP device epochs done')
if rank == 0:
    w_1, loss_1 = epoch_weights(False)
    print("device-resident epochs (CUDA graph with the gradient exchange captured): DP loss %.6f vs single %.6f; weights rel %.2e"
          % (loss_dp, loss_1, float((w_dp - w_1).abs().max() / w_1.abs().max())))
same_t = torch.tensor([1 if same else 0], device=dev); dist.all_reduce(same_t, op=dist.ReduceOp.MIN)
assert int(same_t.item()) == 1

# ---- ragged last batch with FEWER rows than ranks: the rank with an empty slice still joins the all-reduce and the Adam step --------
dyn3 = build()
dyn3.enable_data_parallel(sync=SYNC)
rem = max(1, world - 1)  # rank world-1 gets no row
obs_r = torch.from_numpy(rng.normal(size=(E, rem, O)).astype(np.float32)).to(dev)
act_r = torch.from_numpy(rng.uniform(-1, 1, size=(E, rem, A)).astype(np.float32)).to(dev)
tgt_r = obs_r + 0.1 * torch.tanh(obs_r)
slr = dp_row_slice(rem, world, rank)
dyn3.train_step("transition", obs_r[:, slr].contiguous(), act_r[:, slr].contiguous(), tgt_r[:, slr].contiguous(), grad_rows=rem)
torch.cuda.synchronize()
w3 = dyn3.transition.flat_parameters()[0].clone()
ref3 = w3.clone(); dist.broadcast(ref3, 0)
ok3 = torch.tensor([1 if torch.equal(ref3, w3) else 0], device=dev); dist.all_reduce(ok3, op=dist.ReduceOp.MIN)
assert int(ok3.item()) == 1, "replicas diverged on a ragged batch with an empty shard"
if rank == 0:
    single3 = build()
    single3.train_step("transition", obs_r, act_r, tgt_r)
    w1 = single3.transition.flat_parameters()[0]
    print("ragged batch of %d rows over %d ranks (one empty shard): replicas identical, weights vs single process rel %.2e"
          % (rem, world, float((w3 - w1).abs().max() / w1.abs().max())))

# ---- sharded hold-out evaluation + elite selection over NCCL, and broadcast_parameters ---------------------------
from cmrl_b200.transition_iterator import TransitionIterator
from cmrl_b200.types import InteractionBatch

r4 = np.random.default_rng(9)
nv = 1001  # ragged over the ranks
o4 = r4.normal(size=(nv, O)).astype(np.float32); a4 = r4.uniform(-1, 1, size=(nv, A)).astype(np.float32)
val = TransitionIterator(InteractionBatch(o4, a4, (o4 + 0.1 * np.tanh(o4)).astype(np.float32), np.zeros(nv, np.float32), np.zeros(nv, bool)), 256,
                         shuffle_each_epoch=False)
mse_dp = dyn3.evaluate_member_mse(val, "transition")
dyn3.dp_group = None
mse_1 = dyn3.evaluate_member_mse(val, "transition")
dyn3.enable_data_parallel(sync=SYNC)
err4 = float((mse_dp - mse_1).abs().max() / mse_1.abs().max())
gath = [torch.zeros_like(mse_dp) for _ in range(world)]
dist.all_gather_object(gath, mse_dp)
assert all(torch.equal(g_, gath[0]) for g_ in gath), "ranks disagree on the hold-out scores"
assert err4 < 1e-5, err4
if rank != 0:
    with torch.no_grad():
        dyn3.transition.flat_parameters()[0].fill_(float(rank))
    dyn3.transition.set_elite_members([6, 5, 4, 3, 2])
dyn3.broadcast_parameters(src=0)
w5 = dyn3.transition.flat_parameters()[0]
ref5 = w5.clone(); dist.broadcast(ref5, 0)
ok5 = torch.tensor([1 if (torch.equal(ref5, w5) and list(dyn3.transition.elite_members) == list(build().transition.elite_members) or rank == 0) else 0], device=dev)
elites = [None] * world
dist.all_gather_object(elites, [int(e) for e in dyn3.transition.elite_members])
assert all(e == elites[0] for e in elites) and bool(torch.equal(ref5, w5)) and not bool((w5 == float(max(rank, 1))).all())
if rank == 0:
    print("sharded hold-out evaluation over NCCL: %d rows over %d ranks, [E] scores identical on every rank, vs single process rel %.2e; "
          "broadcast_parameters: replicas and elite sets equal rank 0's" % (nv, world, err4))
# ---- the exchange alone: this library's peer-memory kernel (with and without the fused Adam) against ncclAllReduce + the Adam kernel ----
from cmrl_b200 import ops
from cmrl_b200.peer_sync import PeerGradSync

n_arena = int(dyn.transition._n_trainable)
peer = dyn._peer_sync if dyn._peer_sync is not None else PeerGradSync(dist.group.WORLD, n_arena, dev)
g_buf = torch.randn(n_arena, device=dev) * 1e-3
p_buf, m_buf, v_buf = torch.randn(n_arena, device=dev), torch.zeros(n_arena, device=dev), torch.zeros(n_arena, device=dev)
step_dev = torch.zeros(2, dtype=torch.int32, device=dev)  # [steps taken, arrival ticket]
# correctness of the bare exchange against NCCL on the same data
g_ref = g_buf.clone(); dist.all_reduce(g_ref)
g_mine = g_buf.clone(); peer.allreduce(g_mine); torch.cuda.synchronize(); peer.check()
ex_err = float((g_mine - g_ref).abs().max() / g_ref.abs().max())
g_all = [torch.empty_like(g_mine) for _ in range(world)]; dist.all_gather(g_all, g_mine)
ex_same = all(torch.equal(x, g_all[0]) for x in g_all)


def timed(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3  # us, max over ranks


def nccl_then_adam():
    dist.all_reduce(g_buf)
    ops.adam_l2_step_dev(p_buf, g_buf, m_buf, v_buf, 1e-3, 0.9, 0.999, 1e-8, 1e-5, step_dev, 1.0 / world)


t_nccl = timed(lambda: dist.all_reduce(g_buf))
g_buf.normal_().mul_(1e-3)
t_nccl_adam = timed(nccl_then_adam)
g_buf.normal_().mul_(1e-3)
t_peer = timed(lambda: peer.allreduce(g_buf))
g_buf.normal_().mul_(1e-3)
t_peer_adam = timed(lambda: peer.allreduce_adam(p_buf, g_buf, m_buf, v_buf, 1e-3, 0.9, 0.999, 1e-8, 1e-5, step_dev, 1.0 / world))
t_adam = timed(lambda: ops.adam_l2_step_dev(p_buf, g_buf, m_buf, v_buf, 1e-3, 0.9, 0.999, 1e-8, 1e-5, step_dev, 1.0))
peer.check()
if rank == 0:
    print("gradient exchange of %d floats (%.2f MB) over %d GPUs, back-to-back launches, us per call (max over ranks): ncclAllReduce %.1f, "
          "ncclAllReduce + Adam kernel %.1f | peer-memory exchange %.1f, exchange fused with Adam %.1f | Adam kernel alone %.1f; "
          "peer exchange vs NCCL sum rel %.1e, identical on all ranks: %s"
          % (n_arena, n_arena * 4 / 1e6, world, t_nccl, t_nccl_adam, t_peer, t_peer_adam, t_adam, ex_err, ex_same))
assert ex_same and ex_err < 1e-5
dist.barrier()
torch.cuda.synchronize()
mark('done')
sys.stdout.flush()
os._exit(0)  # captured graphs still reference the NCCL communicator: skip the orderly teardown
