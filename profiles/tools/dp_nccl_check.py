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


def build():
    torch.manual_seed(0)
    np.random.seed(0)
    tr = cm.PlainTransition(O, A, ensemble_num=E, hid_size=200, activation_fn_cfg=SILU, device=dev)
    return cm.PlainEnsembleDynamics(tr, learned_reward=False, optim_lr=1e-3)


rng = np.random.default_rng(1)
batches = [(torch.from_numpy(rng.normal(size=(E, B, O)).astype(np.float32)).to(dev),
            torch.from_numpy(rng.uniform(-1, 1, size=(E, B, A)).astype(np.float32)).to(dev)) for _ in range(steps)]
dyn = build()
dyn.enable_data_parallel()
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
    print("DP over NCCL, world %d: weights identical on all ranks: %s; weights vs single-process full batch: rel %.2e; "
          "update (20 Adam steps) rel %.2e; %.3f ms per DP step (bs %d per member, eager launches)"
          % (world, same, float((flat_dp - flat_1).abs().max() / flat_1.abs().max()), upd_err, dt * 1e3, B))
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
        dyn2.enable_data_parallel()
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
    print("device-resident epochs (CUDA graph with the NCCL all-reduce captured): DP loss %.6f vs single %.6f; weights rel %.2e"
          % (loss_dp, loss_1, float((w_dp - w_1).abs().max() / w_1.abs().max())))
same_t = torch.tensor([1 if same else 0], device=dev); dist.all_reduce(same_t, op=dist.ReduceOp.MIN)
assert int(same_t.item()) == 1

# ---- ragged last batch with FEWER rows than ranks: the rank with an empty slice still joins the all-reduce and the Adam step --------
dyn3 = build()
dyn3.enable_data_parallel()
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
dyn3.enable_data_parallel()
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
dist.barrier()
torch.cuda.synchronize()
mark('done')
sys.stdout.flush()
os._exit(0)  # captured graphs still reference the NCCL communicator: skip the orderly teardown
