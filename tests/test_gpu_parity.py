"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against
 (a) golden outputs of the unmodified reference (tests/golden/*.npz) and
 (b) the CPU oracle (oracle/cmrl_oracle.py) on seeded random inputs.
Tolerances: fp32 tier 1e-5 relative (max|a-b| / max|b|) for outputs/losses; integer / flag outputs bit-exact."""
import os
import tempfile
import types
from unittest import mock

import numpy as np
import pytest
import torch

from conftest import elem_err, load_golden, params_of, record_max, rel_err
from oracle import cmrl_oracle as oc

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star: 1e-5 relative at fp32
SILU = {"_target_": "torch.nn.SiLU"}
DEV = "cuda:0"


@pytest.fixture(scope="module")
def cm():
    import cmrl_b200

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return cmrl_b200


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def load_sd(model, params):
    model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})


def act_cfg(name):
    return None if "relu" in name else SILU


# ------------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("P,E,B,I,O_", [(1, 7, 128, 6, 200), (1, 7, 33, 200, 200), (1, 7, 1, 200, 10), (5, 7, 40, 6, 24), (3, 4, 257, 24, 2), (1, 1, 5, 3, 1)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_ens_linear_fwd_bwd_vs_oracle(cm, P, E, B, I, O_, act):
    rng = np.random.default_rng(P * 1000 + B + act)
    lead = (E,) if P == 1 else (P, E)
    x = rng.normal(size=lead + (B, I)).astype(np.float32)
    w = (rng.normal(size=lead + (I, O_)) / np.sqrt(I)).astype(np.float32)
    b = rng.normal(size=lead + (1, O_)).astype(np.float32)
    y, z = cm.ops.ens_linear_fwd(t(x), t(w), t(b), act, want_z=True)
    z_ref = oc.ens_linear(x, w, b)
    y_ref = oc.activation(z_ref, act)
    assert rel_err(z.cpu().numpy(), z_ref) < TOL
    assert rel_err(y.cpu().numpy(), y_ref) < TOL
    dz = rng.normal(size=lead + (B, O_)).astype(np.float32)
    zprev = rng.normal(size=lead + (B, I)).astype(np.float32)
    dx = cm.ops.ens_linear_bwd_dx(t(dz), t(w), t(zprev), act)
    dx_ref = np.matmul(dz, np.swapaxes(w, -1, -2)) * oc.activation_grad(zprev, act)
    assert rel_err(dx.cpu().numpy(), dx_ref) < TOL
    dw, db = cm.ops.ens_linear_bwd_dw(t(x), t(dz), w.shape)
    assert rel_err(dw.cpu().numpy(), np.matmul(np.swapaxes(x, -1, -2), dz)) < TOL
    assert rel_err(db.cpu().numpy(), dz.sum(-2, keepdims=True)) < TOL


def test_ens_linear_shared_input_and_layer_modules(cm):
    """reference unit tests test_layers.py:25-33,51-54 (output shapes) + broadcast input rows."""
    lin = cm.EnsembleLinearLayer(5, 6, ensemble_num=4).to(DEV)
    x = torch.rand(4, 128, 5, device=DEV)
    assert lin(x).shape == (4, 128, 6)
    assert rel_err(lin(x).detach().cpu().numpy(), (x.matmul(lin.weight) + lin.bias).detach().cpu().numpy()) < TOL
    plin = cm.ParallelEnsembleLinearLayer(5, 6, parallel_num=3, ensemble_num=4).to(DEV)
    xp = torch.rand(3, 4, 128, 5, device=DEV)
    assert plin(xp).shape == (3, 4, 128, 6)
    x1 = torch.rand(128, 5, device=DEV)
    y = cm.ops.ens_linear_fwd(x1, lin.weight.detach(), lin.bias.detach())
    assert torch.equal(y, cm.ops.ens_linear_fwd(x1.expand(4, 128, 5), lin.weight.detach(), lin.bias.detach()))
    # autograd through the layer module
    lin.zero_grad()
    out = lin(x)
    out.square().sum().backward()
    ref = (x.matmul(lin.weight.detach().requires_grad_()) + lin.bias.detach())
    w2 = lin.weight.detach().clone().requires_grad_()
    b2 = lin.bias.detach().clone().requires_grad_()
    (x.matmul(w2) + b2).square().sum().backward()
    assert rel_err(lin.weight.grad.cpu().numpy(), w2.grad.cpu().numpy()) < TOL
    assert rel_err(lin.bias.grad.cpu().numpy(), b2.grad.cpu().numpy()) < TOL


# ------------------------------------------------------------------------------------------------ forward vs golden
@pytest.mark.parametrize("name", ["plain_c1_forward.npz", "plain_relu_forward.npz", "plain_noresidual_forward.npz"])
def test_plain_transition_forward_golden(cm, name):
    g = load_golden(name)
    p = params_of(g)
    E, B, O = g["obs"].shape
    A = g["act"].shape[-1]
    H = p["hidden_layers.0.0.weight"].shape[-1]
    m = cm.PlainTransition(O, A, ensemble_num=E, hid_size=H, activation_fn_cfg=act_cfg(name), residual="noresidual" not in name, device=DEV)
    load_sd(m, p)
    with torch.no_grad():
        mean, logvar = m.forward(t(g["obs"]), t(g["act"]))
        model_in = {"batch_obs": t(g["obs"]), "batch_action": t(g["act"])}
        nll = m.get_nll_loss(model_in, t(g["target"]))
        mse = m.get_mse_loss(model_in, t(g["target"]))
    assert mean.shape == logvar.shape == (E, B, O)
    assert rel_err(mean.cpu().numpy(), g["mean"]) < TOL
    assert rel_err(logvar.cpu().numpy(), g["logvar"]) < TOL
    assert rel_err(nll.cpu().numpy(), g["nll"]) < 3e-5
    assert rel_err(mse.cpu().numpy(), g["mse"]) < 3e-5
    # shared-row (query) path == replicated path, bit for bit
    if name == "plain_c1_forward.npz":
        with torch.no_grad():
            mean2, logvar2 = m.forward(t(g["obs"][:1]).expand(E, B, O), t(g["act"][:1]).expand(E, B, A))
        assert torch.equal(mean, mean2) and torch.equal(logvar, logvar2)


@pytest.mark.parametrize("name,cls", [("reward_forward.npz", "PlainRewardMech"), ("termination_forward.npz", "PlainTerminationMech")])
def test_reward_termination_forward_golden(cm, name, cls):
    g = load_golden(name)
    p = params_of(g)
    m = getattr(cm, cls)(5, 1, ensemble_num=7, hid_size=p["hidden_layers.0.0.weight"].shape[-1], activation_fn_cfg=SILU, device=DEV)
    load_sd(m, p)
    with torch.no_grad():
        mean, logvar = m.forward(t(g["obs"]), t(g["act"]))
        nll = m.get_nll_loss({"batch_obs": t(g["obs"]), "batch_action": t(g["act"])}, t(g["target"]))
    assert mean.shape == (7, 50, 1)
    assert rel_err(mean.cpu().numpy(), g["mean"]) < TOL and rel_err(logvar.cpu().numpy(), g["logvar"]) < TOL
    assert rel_err(nll.cpu().numpy(), g["nll"]) < 3e-5


@pytest.mark.parametrize("tag", ["2d", "3de", "3db", "4d", "wide"])
def test_external_mask_forward_golden(cm, tag):
    g = load_golden("extmask_%s_forward.npz" % tag)
    p = params_of(g)
    E, B, O = g["obs"].shape
    A = g["act"].shape[-1]
    m = cm.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=min(5, E - 1), hid_size=p["hidden_layers.0.0.weight"].shape[-1], activation_fn_cfg=SILU, device=DEV)
    load_sd(m, p)
    m.set_input_mask(torch.from_numpy(g["input_mask"]))
    with torch.no_grad():
        mean, logvar = m.forward(t(g["obs"]), t(g["act"]))
        nll = m.get_nll_loss({"batch_obs": t(g["obs"]), "batch_action": t(g["act"])}, t(g["target"]))
    assert mean.shape == logvar.shape == (E, B, O)
    assert rel_err(mean.cpu().numpy(), g["mean"]) < TOL and rel_err(logvar.cpu().numpy(), g["logvar"]) < TOL
    assert rel_err(nll.cpu().numpy(), g["nll"]) < 3e-5


def test_reference_invariants(cm):
    """test_basic_ensemble.py:40-72 and test_external_mask_ensemble.py:42-93 re-used as black-box tests."""
    torch.manual_seed(0)
    obs = torch.rand(7, 128, 11, device=DEV)
    act = torch.rand(7, 128, 3, device=DEV)
    det = cm.PlainTransition(11, 3, device=DEV, deterministic=True)
    mean, logvar = det.forward(obs, act)
    assert mean.shape == (7, 128, 11) and logvar is None
    for cls in (cm.PlainTransition, cm.ExternalMaskTransition):
        a = cls(11, 3, device=DEV, hid_size=64)
        mean, logvar = a.forward(obs, act)
        assert mean.shape == logvar.shape == (7, 128, 11)
        with tempfile.TemporaryDirectory() as d:
            a.save(d)
            b = cls(11, 3, device=DEV, hid_size=64)
            mean_b, _ = b.forward(obs, act)
            assert not (mean == mean_b).all()
            b.load(d)
            mean_b, logvar_b = b.forward(obs, act)
            assert (mean == mean_b).all() and (logvar == logvar_b).all()  # bit-equal after save -> load
            assert list(np.asarray(b.elite_members)) == list(np.asarray(a.elite_members))
    em = cm.ExternalMaskTransition(11, 3, device=DEV, hid_size=64)
    det_em = cm.ExternalMaskTransition(11, 3, device=DEV, hid_size=64, deterministic=True)
    m1, lv1 = det_em.forward(obs, act)
    assert m1.shape == (7, 128, 11) and lv1 is None
    mask = torch.ones(11, 14)
    em.set_input_mask(mask)
    mean, _ = em.forward(obs, act)
    mask2 = mask.clone()
    mask2[0] = 0
    em.set_input_mask(mask2)
    new_mean, _ = em.forward(obs, act)
    assert not (mean == new_mean).all()
    assert (mean[..., 1:] == new_mean[..., 1:]).all()  # sub-network independence, mask orientation [out,in]
    fe = cm.ForwardEulerTransition(cm.PlainTransition(11, 3, device=DEV, hid_size=32), repeat_times=3)
    mean, logvar = fe.forward(obs, act)
    assert mean.shape == logvar.shape == (7, 128, 11)


# ------------------------------------------------------------------------------------------------ training
def _build_for_train(cm, name, g):
    p = params_of(g, "p0.")
    H = p["hidden_layers.0.0.weight"].shape[-1]
    if name.startswith("reward"):
        m = cm.PlainRewardMech(5, 1, hid_size=H, activation_fn_cfg=SILU, device=DEV)
    elif name.startswith("extmask"):
        m = cm.ExternalMaskTransition(5, 1, hid_size=H, activation_fn_cfg=SILU, device=DEV)
        m.set_input_mask(torch.from_numpy(g["input_mask"]))
    else:
        m = cm.PlainTransition(5, 1, hid_size=H, activation_fn_cfg=act_cfg(name), learn_logvar_bounds="learnbounds" in name, device=DEV)
    load_sd(m, p)
    return m


@pytest.fixture(params=["auto", 3], ids=["auto_tier", "3xtf32"])
def train_tier(request, cm):
    """training GEMM tier: the size-based default, and the tcgen05 3xTF32 kernels forced for every GEMM"""
    old = cm.ops.TRAIN_GEMM_TIER
    cm.ops.TRAIN_GEMM_TIER = request.param
    yield request.param
    cm.ops.TRAIN_GEMM_TIER = old


@pytest.mark.parametrize("name", ["plain_train.npz", "plain_relu_train.npz", "reward_train.npz", "extmask_train.npz", "plain_learnbounds_train.npz"])
@pytest.mark.parametrize("path", ["fused", "layers", "autograd"])
def test_train_steps_golden(cm, name, path, train_tier):
    """fused: train_step on the fused tcgen05 3xTF32 kernels (train_tc.cu) where the model is eligible; layers: train_step on
    the per-layer kernels; autograd: get_nll_loss + backward through the autograd nodes."""
    g = load_golden(name)
    m = _build_for_train(cm, name, g)
    if path == "fused" and train_tier != "auto":
        pytest.skip("the GEMM tier only selects among the per-layer kernels")
    old_fused = cm.ops.TRAIN_FUSED
    cm.ops.TRAIN_FUSED = path == "fused"
    try:
        _train_steps_golden(cm, name, path, g, m)
    finally:
        cm.ops.TRAIN_FUSED = old_fused


def _train_steps_golden(cm, name, path, g, m):
    mech = "reward_mech" if name.startswith("reward") else "transition"
    tr = m if mech == "transition" else cm.PlainTransition(5, 1, hid_size=8, device=DEV)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=mech == "reward_mech", reward_mech=m if mech == "reward_mech" else None)
    optim = getattr(dyn, mech + "_optimizer")
    steps = int(g["steps"])
    for s in range(steps):
        obs, act, target = t(g["obs%d" % s]), t(g["act%d" % s]), t(g["target%d" % s])
        if path in ("fused", "layers"):
            loss = dyn.train_step(mech, obs, act, target)
            if path == "fused" and "learnbounds" not in name:
                assert cm.ops.fused_train_plan(m, obs.shape[0], obs.shape[1]) is not None, "fused tcgen05 step not taken"
        else:
            loss = m.get_nll_loss({"batch_obs": obs, "batch_action": act}, target)
            optim.zero_grad()
            loss.mean().backward()
            if s == 0:
                for k, prm in m.named_parameters():
                    if prm.requires_grad:
                        assert rel_err(prm.grad.cpu().numpy(), g["g0." + k]) < 1e-4, k
            optim.step()
        assert record_max("train %s %s loss" % (name, path), rel_err(loss.detach().cpu().numpy(), g["loss%d" % s])) < TOL
    ref = params_of(g, "p%d." % steps)
    sd = m.state_dict()
    for k in ref:
        mine = sd[k].cpu().numpy()
        if k.endswith("weight"):
            assert rel_err(mine, ref[k]) < TOL, k
        # the 3-step update itself (magnitude ~3e-4) must match
        assert rel_err(mine - g["p0." + k], ref[k] - g["p0." + k]) < 2e-3 or np.array_equal(mine, ref[k]), k


def test_learn_golden(cm):
    """PlainEnsembleDynamics.learn end to end against the reference run (same seeded iterators):
    elite indices bit-exact, final weights within 1e-4 (4 epochs of Adam steps)."""
    g = load_golden("learn_small.npz")
    tr = cm.PlainTransition(5, 1, hid_size=32, activation_fn_cfg=SILU, device=DEV)
    rw = cm.PlainRewardMech(5, 1, hid_size=32, activation_fn_cfg=SILU, device=DEV)
    load_sd(tr, params_of(g, "tr0."))
    load_sd(rw, params_of(g, "rw0."))
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw, optim_lr=1e-3)
    from oracle.refshim import DuckReplayBuffer

    buf = DuckReplayBuffer(g["obs"], g["act"], g["next_obs"], g["rew"], g["done"])
    seeds = iter(range(1000, 2000))
    real_default_rng = np.random.default_rng

    def seeded_default_rng(seed=None):
        return real_default_rng(next(seeds) if seed is None else seed)

    with mock.patch("numpy.random.default_rng", seeded_default_rng):
        dyn.learn(buf, validation_ratio=0.2, batch_size=64, longest_epoch=4, improvement_threshold=0.01, patience=3)
    assert dyn.total_epoch == {"transition": int(g["total_epoch_tr"]), "reward_mech": int(g["total_epoch_rw"])}
    assert list(map(int, tr.elite_members)) == list(map(int, g["tr_elite"]))
    assert list(map(int, rw.elite_members)) == list(map(int, g["rw_elite"]))
    for pre, m in (("tr1.", tr), ("rw1.", rw)):
        ref = params_of(g, pre)
        sd = m.state_dict()
        for k in ref:
            if k.endswith("weight"):
                assert rel_err(sd[k].cpu().numpy(), ref[k]) < 1e-4, k
    _, val_it = dyn.dataset_split(buf, 0.2, 64)
    assert rel_err(dyn.evaluate_member_mse(val_it, "transition").numpy(), g["tr_val_mse"]) < 1e-4
    assert rel_err(dyn.evaluate(val_it, "transition").mean(dim=(1, 2)).numpy(), g["tr_val_mse"]) < 1e-4
    assert rel_err(dyn.evaluate_member_mse(val_it, "reward_mech").numpy(), g["rw_val_mse"]) < 1e-4
    with tempfile.TemporaryDirectory() as d:
        dyn.save(d)
        assert sorted(os.listdir(d)) == ["plain_reward_mech.pth", "plain_transition.pth"]
        dyn.load(d)


def test_constraint_based_dynamics_learn_golden(cm):
    """ConstraintBasedDynamics.learn against the reference run (constraint_based_dynamics.py:83-183): the oracle causal mask
    is put on the masked transition before training, every member iterates its own PERMUTATION of the data
    (`bootstrap_permutes=True`), ragged last batch, checkpoint (with the mask) written at the end.  Epoch counts and elite
    indices bit-exact, weights and hold-out scores within 1e-4 after three epochs of Adam steps."""
    g = load_golden("constraint_learn.npz")
    from oracle.refshim import DuckReplayBuffer

    tr = cm.ExternalMaskTransition(5, 1, hid_size=24, activation_fn_cfg=SILU, device=DEV)
    rw = cm.PlainRewardMech(5, 1, hid_size=24, activation_fn_cfg=SILU, device=DEV)
    load_sd(tr, params_of(g, "tr0."))
    load_sd(rw, params_of(g, "rw0."))
    dyn = cm.ConstraintBasedDynamics(tr, learned_reward=True, reward_mech=rw, optim_lr=1e-3)
    assert dyn.transition_oracle_mask is None and tuple(dyn.transition_history_mask.shape) == (5, 6)
    assert not hasattr(dyn, "reward_mech_oracle_mask")  # only mechanisms with an input mask get one
    buf = DuckReplayBuffer(g["obs"], g["act"], g["next_obs"], g["rew"], g["done"])
    with pytest.raises(ValueError):  # no oracle mask yet: `set_input_mask(None)` fails as in the reference (util.py:64-69)
        dyn.learn(buf, longest_epoch=1, work_dir=None)
    dyn.set_oracle_mask("transition", g["oracle_mask"])
    seeds = iter(range(3000, 4000))
    real_default_rng = np.random.default_rng

    def seeded_default_rng(seed=None):
        return real_default_rng(next(seeds) if seed is None else seed)

    with tempfile.TemporaryDirectory() as d, mock.patch("numpy.random.default_rng", seeded_default_rng):
        dyn.learn(buf, validation_ratio=0.2, batch_size=64, bootstrap_permutes=True, longest_epoch=3,
                  improvement_threshold=0.01, patience=2, work_dir=d)
        assert sorted(os.listdir(d)) == list(g["files"])
        saved = torch.load(os.path.join(d, "external_mask_transition.pth"), weights_only=False)
    assert sorted(saved.keys()) == list(g["saved_keys"])
    assert np.array_equal(saved["input_mask"].cpu().numpy(), g["saved_input_mask"])
    assert np.array_equal(tr.input_mask.cpu().numpy(), g["oracle_mask"])
    assert dyn.total_epoch == {"transition": int(g["total_epoch_tr"]), "reward_mech": int(g["total_epoch_rw"])}
    assert list(map(int, tr.elite_members)) == list(map(int, g["tr_elite"]))
    assert list(map(int, rw.elite_members)) == list(map(int, g["rw_elite"]))
    for pre, m in (("tr1.", tr), ("rw1.", rw)):
        ref = params_of(g, pre)
        sd = m.state_dict()
        for k in ref:
            if k.endswith("weight"):
                assert record_max("constraint learn " + pre + k, rel_err(sd[k].cpu().numpy(), ref[k])) < 1e-4, k
    _, val_it = dyn.dataset_split(buf, 0.2, 64)
    assert rel_err(dyn.evaluate_member_mse(val_it, "transition").numpy(), g["tr_val_mse"]) < 1e-4
    assert rel_err(dyn.evaluate(val_it, "transition").mean(dim=(1, 2)).numpy(), g["tr_val_mse"]) < 1e-4
    assert rel_err(dyn.evaluate_member_mse(val_it, "reward_mech").numpy(), g["rw_val_mse"]) < 1e-4


# ------------------------------------------------------------------------------------------------ VecFakeEnv
def _build_env(cm, g, rng_mode="numpy"):
    learned_term, penalty_coeff, deterministic, max_steps = g["cfg"]
    learned_term, deterministic, max_steps = bool(learned_term), bool(deterministic), int(max_steps)
    O, A, E, H, N = 5, 1, 7, 32, g["reset_obs"].shape[0]
    if "input_mask" in g:
        tr = cm.ExternalMaskTransition(O, A, hid_size=H, activation_fn_cfg=SILU, device=DEV)
        tr.set_input_mask(torch.from_numpy(g["input_mask"]))
    else:
        tr = cm.PlainTransition(O, A, hid_size=H, activation_fn_cfg=SILU, device=DEV)
    rw = cm.PlainRewardMech(O, A, hid_size=H, activation_fn_cfg=SILU, device=DEV)
    tm = cm.PlainTerminationMech(O, A, hid_size=H, activation_fn_cfg=SILU, device=DEV) if learned_term else None
    load_sd(tr, params_of(g, "tr."))
    load_sd(rw, params_of(g, "rw."))
    tr.set_elite_members(list(g["tr_elite"]))
    rw.set_elite_members(list(g["rw_elite"]))
    if learned_term:
        load_sd(tm, params_of(g, "tm."))
        tm.set_elite_members(list(g["tm_elite"]))
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw, learned_termination=learned_term, termination_mech=tm)
    init_rng = np.random.default_rng(77)
    buf = None
    if "pool_obs" in g:  # resets draw rows of a partly filled real replay buffer with the global numpy RNG
        from oracle.refshim import DuckReplayBuffer

        pool = g["pool_obs"]
        buf = DuckReplayBuffer(pool, np.zeros((len(pool), A), np.float32), pool, np.zeros(len(pool), np.float32), np.zeros(len(pool), bool))
        buf.full, buf.pos = False, int(g["pool_pos"])
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(
        dyn,
        termination_fn=None if learned_term else (lambda next_obs, obs, act: np.abs(next_obs[:, 0:1]) > 1.5),
        get_init_obs_fn=None if buf is not None else (lambda n: init_rng.normal(size=(n, O)).astype(np.float32)),
        real_replay_buffer=buf,
        penalty_coeff=penalty_coeff,
        deterministic=deterministic,
        max_episode_steps=max_steps,
    )
    env.rng_mode = rng_mode
    env.seed(11)
    return env


@pytest.mark.parametrize(
    "name", ["fake_env_termfn.npz", "fake_env_truncate.npz", "fake_env_learned_term.npz", "fake_env_penalty.npz", "fake_env_deterministic.npz", "fake_env_extmask.npz",
             "fake_env_bufreset.npz"]
)
@pytest.mark.parametrize("infos_mode", ["dicts", "shared"])
def test_fake_env_step_golden(cm, name, infos_mode):
    """VecFakeEnv.step vs the reference run with the same seed: done / truncated flags (and hence member
    choice) bit-exact, obs / reward within 1e-5."""
    g = load_golden(name)
    env = _build_env(cm, g)
    env.infos_mode = infos_mode
    if "pool_obs" in g:
        np.random.seed(123)  # buffer resets use the global generator (fake_env.py:154,177): seeded like the fixture's run
    obs0 = env.reset()
    assert np.array_equal(obs0, g["reset_obs"])
    N = obs0.shape[0]
    for step in range(int(g["T"])):
        obs, rew, done, infos = env.step(g["actions"][step])
        assert obs.dtype == np.float32 and rew.dtype == np.float32 and done.dtype == bool
        assert obs.shape == (N, 5) and rew.shape == (N,) and done.shape == (N,) and len(infos) == N
        assert np.array_equal(done, g["done%d" % step])
        assert np.array_equal(np.asarray([i["TimeLimit.truncated"] for i in infos]), g["trunc%d" % step])
        assert rel_err(obs, g["obs%d" % step]) < TOL
        assert rel_err(rew, g["rew%d" % step]) < TOL
        for i, info in enumerate(infos):
            assert ("terminal_observation" in info) == bool(done[i])
            if done[i]:
                assert rel_err(info["terminal_observation"], g["termobs%d" % step][i]) < TOL


def test_selected_member_path_equals_all_member_path(cm):
    """Evaluating only the drawn member (ragged K1) is bit-identical to query + gather (reference order)."""
    g = load_golden("fake_env_extmask.npz")
    env = _build_env(cm, g)
    env.reset()
    rng = np.random.default_rng(3)
    N = env.num_envs
    for mech, D in (("transition", 5), ("reward_mech", 1)):
        model = getattr(env.dynamics, mech)
        obs = env._obs_dev
        act = t(rng.uniform(-1, 1, size=(N, 1)).astype(np.float32))
        idx = t(rng.choice(np.asarray(model.elite_members), size=N).astype(np.uint8))
        noise = t(rng.normal(size=(N, D)).astype(np.float32))
        env.precision = "ffma"  # the CUDA-core tier: same kernels and summation order as the all-member path
        sel = env.predict_selected(model, obs, act, idx, noise)
        with torch.no_grad():
            mean, logvar = model.forward(obs.unsqueeze(0), act.unsqueeze(0))
        allm = cm.ops.select_sample(mean, logvar, idx, noise)
        assert torch.equal(sel, allm)
        ref = oc.dynamics_predict(mean.cpu().numpy(), logvar.cpu().numpy(), idx.cpu().numpy().astype(np.int64), noise.cpu().numpy())
        assert rel_err(sel.cpu().numpy(), ref) < TOL
        # the default tier: the fused tcgen05 kernel with hi/lo split tf32 operands (3xTF32) -- fp32-grade, elementwise too
        env.precision = "fp32"
        assert model.packed_tf32().ok
        sel3 = env.predict_selected(model, obs, act, idx, noise)
        assert record_max("3xTF32 rollout vs all-member %s" % mech, rel_err(sel3.cpu().numpy(), ref)) < TOL
        assert record_max("3xTF32 rollout vs all-member %s elem" % mech, elem_err(sel3.cpu().numpy(), ref)) < 5 * TOL
        mean_only = env.predict_selected(model, obs, act, idx, None)  # deterministic envs: no sampling
        ref_m = np.take_along_axis(mean.cpu().numpy(), idx.cpu().numpy().astype(np.int64)[None, :, None], axis=0)[0]
        assert rel_err(mean_only.cpu().numpy(), ref_m) < TOL


def test_query_and_get_dynamics_predict(cm):
    g = load_golden("fake_env_termfn.npz")
    env = _build_env(cm, g)
    obs = g["reset_obs"]
    act = g["actions"][0]
    q = env.dynamics.query(torch.from_numpy(obs), torch.from_numpy(act))
    assert set(q) == {"batch_next_obs", "batch_reward"}
    tr = params_of(g, "tr.")
    o3, a3 = np.repeat(obs[None], 7, 0), np.repeat(act[None], 7, 0)
    mean, logvar = oc.plain_transition_forward(tr, o3, a3)
    assert q["batch_next_obs"]["mean"].shape == (7, obs.shape[0], 5)
    assert rel_err(q["batch_next_obs"]["mean"], mean) < TOL and rel_err(q["batch_next_obs"]["logvar"], logvar) < TOL
    env.seed(5)
    pred = env.get_dynamics_predict(q, "transition")
    gen = np.random.default_rng(5)
    idx = gen.choice(np.asarray(env.dynamics.transition.elite_members), size=obs.shape[0])
    noise = gen.normal(size=(obs.shape[0], 5)).astype(np.float32)
    assert rel_err(pred, oc.dynamics_predict(q["batch_next_obs"]["mean"], q["batch_next_obs"]["logvar"], idx, noise)) < TOL
    pen = cm.VecFakeEnv.get_penalty(q["batch_next_obs"]["mean"])
    assert rel_err(pen, oc.mopo_penalty(q["batch_next_obs"]["mean"])) < TOL


def test_device_rng_mode_statistics(cm):
    """Throughput mode: Philox member draw uniform over the elites, noise ~ N(0,1); env runs and resets."""
    elite = torch.tensor([0, 2, 3, 5, 6], dtype=torch.uint8, device=DEV)
    N = 1 << 20
    idx, noise = cm.ops.draw_member_noise(123, 1, N, elite, 6)
    counts = np.bincount(idx.cpu().numpy(), minlength=7)
    assert counts[1] == 0 and counts[4] == 0
    assert np.all(np.abs(counts[[0, 2, 3, 5, 6]] / N - 0.2) < 0.003)
    z = noise.cpu().numpy().astype(np.float64)
    assert abs(z.mean()) < 0.003 and abs(z.std() - 1) < 0.003
    assert abs((z**3).mean()) < 0.01 and abs((z**4).mean() - 3) < 0.03
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.005
    idx2, noise2 = cm.ops.draw_member_noise(123, 2, N, elite, 6)
    assert not torch.equal(noise, noise2)
    idx3, noise3 = cm.ops.draw_member_noise(123, 1, N, elite, 6)
    assert torch.equal(noise, noise3) and torch.equal(idx, idx3)
    u = cm.ops.draw_uniform_index(7, 3, N, 1000, DEV).cpu().numpy()
    assert u.min() == 0 and u.max() == 999 and abs(u.mean() - 499.5) < 2

    g = load_golden("fake_env_truncate.npz")
    env = _build_env(cm, g, rng_mode="device")
    env.reset()
    lengths = np.zeros(env.num_envs, dtype=int)
    for step in range(6):
        obs, rew, done, infos = env.step(g["actions"][step % 4])
        lengths += 1
        trunc = np.asarray([i["TimeLimit.truncated"] for i in infos])
        assert np.array_equal(trunc, lengths >= 3)
        assert np.all(done[trunc])
        lengths[done] = 0
        assert np.array_equal(env._envs_length, lengths)
        assert np.isfinite(obs).all() and np.isfinite(rew).all()


def test_member_bucket_and_tail_edge_cases(cm):
    rng = np.random.default_rng(0)
    for N, E in ((1, 1), (5, 7), (63, 3), (1000, 7), (100003, 16), (40000, 70), (70001, 200)):
        idx = rng.integers(0, E, size=N).astype(np.uint8)
        perm, offsets = cm.ops.member_bucket(t(idx), E)
        perm, offsets = perm.cpu().numpy(), offsets.cpu().numpy()
        assert np.array_equal(offsets, np.concatenate([[0], np.cumsum(np.bincount(idx, minlength=E))]))
        # the counting sort is stable: exactly numpy's stable argsort (a permutation, grouped by member, ascending env id)
        assert np.array_equal(perm, np.argsort(idx, kind="stable"))
    # empty member buckets (all envs choose member 3)
    perm, offsets = cm.ops.member_bucket(t(np.full(64, 3, np.uint8)), 7)
    assert offsets.cpu().tolist() == [0, 0, 0, 0, 64, 64, 64, 64]
    # wide observations take the warp-per-env tail
    N, O = 257, 40
    nxt = rng.normal(size=(N, O)).astype(np.float32)
    reset = rng.normal(size=(N, O)).astype(np.float32)
    term = (rng.uniform(size=N) < 0.3).astype(np.float32) * 0.25
    env_len = rng.integers(0, 5, size=N).astype(np.int32)
    rew = rng.normal(size=N).astype(np.float32)
    obs_out, rew_out = torch.empty(N, O, device=DEV), torch.empty(N, device=DEV)
    done, trunc = torch.empty(N, dtype=torch.uint8, device=DEV), torch.empty(N, dtype=torch.uint8, device=DEV)
    term_obs, n_done, len_dev = torch.empty(N, O, device=DEV), torch.zeros(1, dtype=torch.int32, device=DEV), t(env_len)
    cm.ops.env_step_tail(t(nxt), t(rew), t(term), None, None, 0.0, len_dev, 4, t(reset), None, obs_out, rew_out, done, trunc, term_obs, n_done)
    o_ref, r_ref, d_ref, t_ref, to_ref, l_ref = oc.env_step_tail(nxt, rew, term, env_len.astype(np.int64), 4, reset)
    assert np.array_equal(done.cpu().numpy().astype(bool), d_ref) and np.array_equal(trunc.cpu().numpy().astype(bool), t_ref)
    assert np.array_equal(obs_out.cpu().numpy(), o_ref) and np.array_equal(term_obs.cpu().numpy(), to_ref)
    assert np.array_equal(len_dev.cpu().numpy(), l_ref) and int(n_done.item()) == int(d_ref.sum())


def test_large_rollout_properties(cm):
    """BASELINE-size property check (C2: 100k envs, ExternalMask 4x200): selected path == all-member path on a
    random subset, and the step is deterministic for fixed draws."""
    torch.manual_seed(0)
    np.random.seed(0)
    N = 100_000
    tr = cm.ExternalMaskTransition(5, 1, activation_fn_cfg=SILU, device=DEV)
    rng = np.random.default_rng(0)
    mask = (rng.uniform(size=(5, 6)) < 0.5).astype(np.float32)
    mask[np.arange(5), np.arange(5)] = 1
    tr.set_input_mask(torch.from_numpy(mask))
    rw = cm.PlainRewardMech(5, 1, activation_fn_cfg=SILU, device=DEV)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool), get_init_obs_fn=lambda n: rng.normal(size=(n, 5)).astype(np.float32))
    env.reset()
    obs = env._obs_dev.clone()
    act = t(rng.uniform(-1, 1, size=(N, 1)).astype(np.float32))
    idx = t(rng.choice(np.asarray(tr.elite_members), size=N).astype(np.uint8))
    noise = t(rng.normal(size=(N, 5)).astype(np.float32))
    sub = torch.arange(0, N, 97, device=DEV)
    with torch.no_grad():
        mean, logvar = tr.forward(obs[sub].unsqueeze(0), act[sub].unsqueeze(0))
    ref = cm.ops.select_sample(mean, logvar, idx[sub].contiguous(), noise[sub].contiguous())
    env.precision = "ffma"  # CUDA-core tier: bit-identical to the all-member path
    a = env.predict_selected(tr, obs, act, idx, noise)
    b = env.predict_selected(tr, obs, act, idx, noise)
    assert torch.equal(a, b)
    assert torch.equal(a[sub], ref)
    env.precision = "fp32"  # default tier, 3xTF32 on tcgen05: deterministic, fp32-grade
    a3 = env.predict_selected(tr, obs, act, idx, noise)
    assert torch.equal(a3, env.predict_selected(tr, obs, act, idx, noise))
    assert record_max("3xTF32 rollout C2 100k envs vs all-member", rel_err(a3[sub].cpu().numpy(), ref.cpu().numpy())) < TOL
    # elementwise over all 500 000 outputs, floor 1e-3 of the largest: an output near zero (observation + delta cancel) carries the
    # absolute fp32 rounding of its O(1) terms, ~1e-7 of the output scale, i.e. up to ~1e-4 in this metric for either tier
    assert record_max("3xTF32 rollout C2 100k envs vs ffma tier elem", elem_err(a3.cpu().numpy(), a.cpu().numpy())) < 5e-4


# ------------------------------------------------------------------------------------------------ tcgen05 tier
TOL_TC = 1e-3  # north_star: 1e-3 relative at TF32 / 16-bit operands


def _tc_case(cm, kind, N, H, seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    O, A = 5, 1
    if kind == "extmask":
        m = cm.ExternalMaskTransition(O, A, hid_size=H, activation_fn_cfg=SILU, device=DEV)
        mask = (rng.uniform(size=(O, O + A)) < 0.5).astype(np.float32)
        mask[np.arange(O), np.arange(O)] = 1
        m.set_input_mask(torch.from_numpy(mask))
    elif kind == "plain":
        m = cm.PlainTransition(O, A, hid_size=H, activation_fn_cfg=SILU, device=DEV)
    elif kind == "plain_relu":
        m = cm.PlainTransition(O, A, hid_size=H, device=DEV)
    else:
        m = cm.PlainRewardMech(O, A, hid_size=H, activation_fn_cfg=SILU, device=DEV)
    with torch.no_grad():  # non-trivial biases / head so every term of the epilogue matters
        for prm in m.parameters():
            if prm.dim() >= 3 and prm.shape[-2] == 1:
                prm.add_(torch.randn_like(prm) * 0.1)
    dyn = cm.PlainEnsembleDynamics(m if kind != "reward" else cm.PlainTransition(O, A, hid_size=8, device=DEV),
                                   learned_reward=kind == "reward", reward_mech=m if kind == "reward" else None)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32), termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
               get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32))
    env.reset()
    obs = env._obs_dev
    act = t(rng.uniform(-1, 1, size=(N, A)).astype(np.float32))
    idx = t(rng.choice(np.asarray(m.elite_members), size=N).astype(np.uint8))
    noise = t(rng.normal(size=(N, m._head_dim)).astype(np.float32))
    return env, m, obs, act, idx, noise


@pytest.fixture(params=[None, 5, 3], ids=["by_shape", "pair256x2", "pingpong"])
def tc_variant(request, cm):
    old = cm.ops.TC_VARIANT
    cm.ops.TC_VARIANT = request.param
    yield request.param
    cm.ops.TC_VARIANT = old


@pytest.mark.parametrize("kind,N,H", [("plain", 3, 200), ("reward", 1, 200), ("extmask", 257, 200), ("plain", 300, 200), ("extmask", 1000, 200), ("reward", 777, 200), ("plain_relu", 130, 64),
                                      ("extmask", 5, 40), ("plain", 20000, 200), ("extmask", 100000, 200)])
def test_tcgen05_rollout_vs_fp32_tier(cm, tc_variant, kind, N, H):
    """The fused tcgen05 kernel (fp16 operands, fp32 accumulate) against the fp32 tier on the same draws:
    1e-3 relative on the predicted variable, and on the network's own contribution (pred - obs residual)."""
    env, m, obs, act, idx, noise = _tc_case(cm, kind, N, H, seed=N)
    env.precision = "fp32"
    ref = env.predict_selected(m, obs, act, idx, noise)
    env.precision = "f16"
    out = env.predict_selected(m, obs, act, idx, noise)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    print("tc variant %s %s N=%d H=%d rel_err %.2e" % (tc_variant, kind, N, H, rel_err(out.cpu().numpy(), ref.cpu().numpy())))
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < TOL_TC
    # deterministic part only (no noise): isolates mean from logvar
    env.precision = "fp32"
    ref_m = env.predict_selected(m, obs, act, idx, None)
    env.precision = "f16"
    out_m = env.predict_selected(m, obs, act, idx, None)
    assert rel_err(out_m.cpu().numpy(), ref_m.cpu().numpy()) < TOL_TC
    if m._residual():
        d_ref = (ref_m - obs).cpu().numpy()
        d_out = (out_m - obs).cpu().numpy()
        assert rel_err(d_out, d_ref) < 5e-3  # the head's own output (no residual to hide behind)
    # repeatable
    assert torch.equal(out, env.predict_selected(m, obs, act, idx, noise))


def test_tcgen05_pack_refreshes_on_weight_change(cm):
    env, m, obs, act, idx, noise = _tc_case(cm, "plain", 256, 200, seed=9)
    env.precision = "f16"
    a = env.predict_selected(m, obs, act, idx, noise)
    with torch.no_grad():
        m.mean_and_logvar.bias.add_(0.5)
    b = env.predict_selected(m, obs, act, idx, noise)
    assert not torch.equal(a, b)
    assert "hidden_layers.0.0.weight" in m.state_dict() and all(v.dtype == torch.float32 for v in m.state_dict().values())


# ------------------------------------------------------------------------------------------------ device-resident training
@pytest.mark.parametrize("use_graph", [False, True])
def test_device_epoch_matches_host_iterator_path(cm, use_graph):
    """train_epoch_device (on-device bootstrap gather + CUDA-graph replay, ragged last batch) performs the same
    updates as the reference-shaped `train` loop over the host BootstrapIterator."""
    g = load_golden("learn_small.npz")
    from oracle.refshim import DuckReplayBuffer

    buf = DuckReplayBuffer(g["obs"], g["act"], g["next_obs"], g["rew"], g["done"])
    results = []
    for device_path in (False, True):
        tr = cm.PlainTransition(5, 1, hid_size=32, activation_fn_cfg=SILU, device=DEV)
        load_sd(tr, params_of(g, "tr0."))
        dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False, optim_lr=1e-3)
        seeds = iter(range(50, 60))
        real = np.random.default_rng
        with mock.patch("numpy.random.default_rng", lambda seed=None: real(next(seeds) if seed is None else seed)):
            train_it, _ = dyn.dataset_split(buf, 0.2, 100)  # 512 train rows -> 5 full batches + ragged 12
        losses = []
        for epoch in range(3):
            if device_path:
                losses.append(dyn.train_epoch_device(train_it, "transition", use_graph=use_graph))
            else:
                losses.append(float(dyn.train(train_it, "transition").mean().item()))
        assert dyn.transition_optimizer.step_count == 18
        results.append((losses, {k: v.cpu().numpy() for k, v in tr.state_dict().items()}))
    (l0, w0), (l1, w1) = results
    assert rel_err(np.asarray(l1), np.asarray(l0)) < 1e-5
    for k in w0:
        assert rel_err(w1[k], w0[k]) < 1e-5, k


def test_device_epoch_recaptures_after_the_arena_moved(cm):
    """The captured training step bakes the addresses of the parameter arena and of the fused plan's buffers.  When the
    arena is re-homed between two epochs over the SAME data set (here: `_flatten_parameters()`, what `model.to()` leads to)
    the runner has to capture the step again: the second epoch must update the live parameters exactly like a dynamics
    that never used a graph."""
    g = load_golden("learn_small.npz")
    from oracle.refshim import DuckReplayBuffer

    buf = DuckReplayBuffer(g["obs"], g["act"], g["next_obs"], g["rew"], g["done"])
    results = []
    for use_graph in (False, True):
        tr = cm.PlainTransition(5, 1, hid_size=32, activation_fn_cfg=SILU, device=DEV)
        load_sd(tr, params_of(g, "tr0."))
        dyn = cm.PlainEnsembleDynamics(tr, learned_reward=False, optim_lr=1e-3)
        seeds = iter(range(50, 60))
        real = np.random.default_rng
        with mock.patch("numpy.random.default_rng", lambda seed=None: real(next(seeds) if seed is None else seed)):
            train_it, _ = dyn.dataset_split(buf, 0.2, 100)
        losses = [dyn.train_epoch_device(train_it, "transition", use_graph=use_graph)]
        old_ptr = tr.flat_parameters()[0].data_ptr()
        keep = tr.flat_parameters()  # the old arena stays allocated, so that the new one cannot reuse its address
        tr._flatten_parameters()
        assert tr.flat_parameters()[0].data_ptr() != old_ptr
        losses.append(dyn.train_epoch_device(train_it, "transition", use_graph=use_graph))
        del keep
        results.append((losses, {k: v.cpu().numpy() for k, v in tr.state_dict().items()}))
    (l0, w0), (l1, w1) = results
    assert rel_err(np.asarray(l1), np.asarray(l0)) < 1e-5
    for k in w0:
        assert rel_err(w1[k], w0[k]) < 1e-5, k


def test_step_device_graphed(cm):
    """CUDA-graph replay of the device-resident step: episode bookkeeping advances, every replay draws fresh noise."""
    g = load_golden("fake_env_truncate.npz")
    for precision in ("fp32", "f16"):
        env = _build_env(cm, g, rng_mode="device")
        env.precision = precision
        env.reset()
        act = t(g["actions"][0])
        seen = []
        lengths = np.zeros(env.num_envs, dtype=int)
        for step in range(7):
            obs, rew, done, trunc, term_obs = env.step_device_graphed(act)
            torch.cuda.synchronize()
            lengths += 1
            d = done.cpu().numpy().astype(bool)
            assert np.array_equal(trunc.cpu().numpy().astype(bool), lengths >= 3) and np.all(d[lengths >= 3])
            lengths[d] = 0
            assert np.array_equal(env._envs_length, lengths)
            assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
            seen.append(rew.cpu().numpy().copy())
        assert all(not np.array_equal(seen[i], seen[i + 1]) for i in range(6))
        # the all-replays reward noise is N(0, sigma): distinct draws -> low correlation between steps
        assert abs(np.corrcoef(seen[1], seen[2])[0, 1]) < 0.6


def test_tcgen05_wide_heads_and_hidden(cm):
    """Shapes outside the default pair kernel: O=20 (head 48 columns -> variant 3); hidden 250 (wider than the weight-stationary
    kernels take and not a multiple of 16) -> a loud error, never a silent fallback."""
    torch.manual_seed(0)
    np.random.seed(0)
    rng = np.random.default_rng(0)
    for O, A, H in ((20, 4, 200), (5, 1, 250), (17, 6, 128)):
        N = 3000
        m = cm.PlainTransition(O, A, hid_size=H, activation_fn_cfg=SILU, device=DEV)
        dyn = cm.PlainEnsembleDynamics(m, learned_reward=False)
        env = cm.VecFakeEnv(N, None, None)
        env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32), termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
                   get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32))
        env.reset()
        act = t(rng.uniform(-1, 1, size=(N, A)).astype(np.float32))
        idx = t(rng.choice(np.asarray(m.elite_members), size=N).astype(np.uint8))
        noise = t(rng.normal(size=(N, O)).astype(np.float32))
        env.precision = "fp32"
        ref = env.predict_selected(m, env._obs_dev, act, idx, noise)
        env.precision = "f16"
        if H > 207:
            with pytest.raises(cm._lib.CmrlB200Error):
                env.predict_selected(m, env._obs_dev, act, idx, noise)
            continue
        out = env.predict_selected(m, env._obs_dev, act, idx, noise)
        assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < TOL_TC


# ---------------------------------------------------------------------------------------------------------------------------
# device-resident multi-step rollout into SB3 ReplayBuffer-layout slabs
# ---------------------------------------------------------------------------------------------------------------------------
def _rollout_env(cm, N, seed, max_steps=7):
    torch.manual_seed(3)
    tr = cm.PlainTransition(4, 2, hid_size=64, activation_fn_cfg=SILU, device=DEV)
    rw = cm.PlainRewardMech(4, 2, hid_size=64, activation_fn_cfg=SILU, device=DEV)
    tm = cm.PlainTerminationMech(4, 2, hid_size=64, activation_fn_cfg=SILU, device=DEV)
    for m in (tr, rw, tm):  # the constructor draws the initial elite set from the global numpy RNG (mlp.py:27)
        m.set_elite_members([0, 2, 3, 5, 6])
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw, learned_termination=True, termination_mech=tm)
    rng = np.random.default_rng(5)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, get_init_obs_fn=lambda n: rng.normal(size=(n, 4)).astype(np.float32), max_episode_steps=max_steps)
    env.rng_mode = "device"
    env.seed(seed)
    env.reset()
    return env


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "f16"])
def test_rollout_device_slabs_equal_step_by_step(cm, precision):
    """`rollout_device` (graph-replayed steps written into [T,N,.] slabs) == the same steps taken one by one through
    `step_device`, bit for bit; slab layout follows SB3's ReplayBuffer (next_observations = pre-reset next state)."""
    N, T = 333, 12
    wp = torch.randn(4, 2, device=DEV)
    policy = lambda obs: torch.tanh(obs @ wp)
    env_a, env_b = _rollout_env(cm, N, 11), _rollout_env(cm, N, 11)
    env_a.precision = env_b.precision = precision
    slabs = env_a.rollout_device(policy, T)
    assert slabs["observations"].shape == (T, N, 4) and slabs["actions"].shape == (T, N, 2) and slabs["rewards"].shape == (T, N)
    for t in range(T):
        obs_before = env_b._obs_dev.clone()
        act = policy(env_b._obs_dev)
        obs, reward, done, trunc, term_obs = env_b.step_device(act)
        assert torch.equal(slabs["observations"][t], obs_before)
        assert torch.equal(slabs["actions"][t], act)
        assert torch.equal(slabs["next_observations"][t], term_obs)
        assert torch.equal(slabs["rewards"][t], reward)
        assert torch.equal(slabs["dones"][t], done.float()) and torch.equal(slabs["timeouts"][t], trunc.float())
    assert torch.equal(env_a._obs_dev, env_b._obs_dev)
    # SB3 invariants: where an env is not done its next_observation is the next step's observation; timeouts imply dones;
    # every env is truncated at max_episode_steps at the latest
    nd = slabs["dones"][:-1] == 0
    assert torch.equal(slabs["next_observations"][:-1][nd], slabs["observations"][1:][nd])
    assert bool((slabs["dones"] >= slabs["timeouts"]).all())
    assert bool((slabs["dones"].cumsum(0)[6] >= 1).all())
    # host step after device-resident stepping sees the current observations
    host_obs = env_a.sync_host_obs()
    assert np.array_equal(host_obs, env_a._obs_dev.cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "f16"])
def test_draws_one_step_ahead_same_numbers(cm, precision):
    """`prep_ahead` (graph-replayed steps draw / bucket the transition's members for the NEXT step) only moves launches:
    host-facing steps, device-resident graph steps, kernel-by-kernel steps and switches between them give bit-identical
    trajectories with and without it."""
    N = 500
    act = np.random.default_rng(1).uniform(-1, 1, size=(N, 2)).astype(np.float32)
    act_d = t(act)
    runs = []
    for ahead in (False, True):
        env = _rollout_env(cm, N, 7)
        env.precision, env.prep_ahead = precision, ahead
        with torch.no_grad():  # a termination head that never fires ("non-zero => done"): episodes end by truncation, so
            tm = env.dynamics.termination_mech  # the returned observations carry the transition's draws
            tm.mean_and_logvar.weight.zero_()
            tm.mean_and_logvar.bias.zero_()
        tm.deterministic = True  # (its own sampling noise would make the output non-zero)
        rec = []

        def host_steps(n):
            for _ in range(n):
                o, r, d, _infos = env.step(act)
                rec.extend([np.array(o), np.array(r), np.array(d)])

        def graph_steps(n):
            for _ in range(n):
                o, r, d, _tr, to = env.step_device_graphed(act_d)
                rec.extend([x.cpu().numpy().copy() for x in (o, r, d, to)])

        host_steps(4)
        graph_steps(3)
        host_steps(3)
        rec.extend([x.cpu().numpy().copy() for x in env.step_device(act_d)])  # kernel by kernel: rewrites the bucket buffers
        graph_steps(2)
        host_steps(2)
        runs.append(rec)
    assert len(runs[0]) == len(runs[1])
    for k, (a, b) in enumerate(zip(*runs)):
        assert np.array_equal(a, b), k
    assert not np.array_equal(runs[0][1], runs[0][4])  # fresh draws every step
    dones = [r for r in runs[0] if r.shape == (N,) and r.dtype in (np.bool_, np.uint8)]
    assert not dones[0].any() and any(d.all() for d in dones)  # episodes end by truncation only (max_episode_steps = 7)


@pytest.mark.gpu
def test_extend_replay_buffer_wraps_like_sb3(cm):
    N, T, cap = 16, 10, 14
    env = _rollout_env(cm, N, 2)
    slabs = env.rollout_device(lambda obs: torch.zeros(N, 2, device=DEV), T)
    buf = types.SimpleNamespace(buffer_size=cap, pos=9, full=False, observations=np.zeros((cap, N, 4), np.float32),
                                next_observations=np.zeros((cap, N, 4), np.float32), actions=np.zeros((cap, N, 2), np.float32),
                                rewards=np.zeros((cap, N), np.float32), dones=np.zeros((cap, N), np.float32),
                                timeouts=np.zeros((cap, N), np.float32))  # fmt: skip
    cm.VecFakeEnv.extend_replay_buffer(buf, slabs)
    assert buf.full and buf.pos == (9 + T) % cap
    rw = slabs["rewards"].cpu().numpy()
    assert np.array_equal(buf.rewards[9:14], rw[:5]) and np.array_equal(buf.rewards[0:5], rw[5:10])
    assert np.array_equal(buf.next_observations[0:5], slabs["next_observations"].cpu().numpy()[5:10])


# ---------------------------------------------------------------------------------------------------------------------------
# K1 tensor-core training tier (tcgen05 kind::tf32, 3xTF32) vs the fp32 FFMA tier
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("P,E,B,i,o", [(1, 7, 256, 6, 200), (1, 7, 256, 200, 200), (1, 7, 256, 200, 10), (5, 7, 100, 6, 200),
                                       (5, 3, 77, 200, 2), (2, 2, 300, 130, 300), (1, 1, 5, 3, 1), (1, 2, 1000, 128, 64),
                                       (1, 7, 4096, 200, 200), (2, 3, 2500, 64, 130)])  # the last two: split-K dw
def test_train_gemm_tensor_core_tiers(cm, P, E, B, i, o):
    """fwd (+pre-activation), dx (fused act'), dw (+db, accumulate) on tcgen05: tier 3 (3xTF32) is fp32-grade (1e-5 tier),
    tier 1 (plain TF32) is the 1e-3 tier.  Shapes cover K tails (6, 130, 3), N tails (10, 2, 300 > one 256-wide tile),
    M tails, shared / per-member inputs and input masks."""
    g = torch.Generator(device=DEV).manual_seed(P * 1000 + B + i)
    rnd = lambda *s: torch.randn(*s, device=DEV, generator=g)
    lead = (P, E) if P > 1 else (E,)
    w, b = rnd(*lead, i, o) / np.sqrt(i), rnd(*lead, 1, o)
    x = rnd(E, B, i) if P > 1 else rnd(*lead, B, i)  # P > 1: rows shared by the sub-nets (ExternalMask layer 1)
    mask = (torch.rand(P, i, device=DEV, generator=g) < 0.6).float() if P > 1 else None
    mstr = cm.ops.mask_strides(mask, P, E, B) if mask is not None else (0, 0, 0)
    dz = rnd(*lead, B, o)
    ref_y, ref_z = cm.ops.ens_linear_fwd(x, w, b, cm.ops.ACT_SILU, mask, mstr, want_z=True, tier=0)
    x_full = x if P == 1 else x.expand(P, E, B, i)
    ref_dx = cm.ops.ens_linear_bwd_dx(dz, w, None, cm.ops.ACT_IDENTITY, tier=0)
    z_in = rnd(*lead, B, i)
    ref_dx2 = cm.ops.ens_linear_bwd_dx(dz, w, z_in, cm.ops.ACT_SILU, tier=0)
    ref_dw, ref_db = cm.ops.ens_linear_bwd_dw(x_full, dz, w.shape, mask, mstr, tier=0)
    for tier, tol in ((3, 1e-5), (1, 2e-3)):
        y, z = cm.ops.ens_linear_fwd(x, w, b, cm.ops.ACT_SILU, mask, mstr, want_z=True, tier=tier)
        assert rel_err(z.cpu().numpy(), ref_z.cpu().numpy()) < tol and rel_err(y.cpu().numpy(), ref_y.cpu().numpy()) < tol
        dx = cm.ops.ens_linear_bwd_dx(dz, w, None, cm.ops.ACT_IDENTITY, tier=tier)
        assert rel_err(dx.cpu().numpy(), ref_dx.cpu().numpy()) < tol
        dx2 = cm.ops.ens_linear_bwd_dx(dz, w, z_in, cm.ops.ACT_SILU, tier=tier)
        assert rel_err(dx2.cpu().numpy(), ref_dx2.cpu().numpy()) < tol
        dw, db = cm.ops.ens_linear_bwd_dw(x_full, dz, w.shape, mask, mstr, tier=tier)
        assert rel_err(dw.cpu().numpy(), ref_dw.cpu().numpy()) < tol and rel_err(db.cpu().numpy(), ref_db.cpu().numpy()) < tol
        dw2, db2 = dw.clone(), db.clone()  # accumulate mode adds into the given buffers
        cm.ops.ens_linear_bwd_dw(x_full, dz, w.shape, mask, mstr, dw=dw2, db=db2, accumulate=True, tier=tier)
        assert rel_err(dw2.cpu().numpy(), 2 * ref_dw.cpu().numpy()) < tol and rel_err(db2.cpu().numpy(), 2 * ref_db.cpu().numpy()) < tol


# ---------------------------------------------------------------------------------------------------------------------------
# C5 at full size (O = P = 100, A = 20, E = 16, H = 512, per-member 3-D mask: 1 600 sub-nets, 1.36 G parameters)
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_c5_full_size_forward_and_rollout(cm):
    """Maximum configuration of BASELINE.json: parity against the oracle on sampled (sub-net, member) pairs (a sub-net's
    output column depends only on its own weights and mask row), selected-member rollout == all-member path bit for bit,
    and the tensor-core tier (streamed weights) against the fp32 tier."""
    O, A, E, H, B = 100, 20, 16, 512, 48
    torch.manual_seed(5)
    rng = np.random.default_rng(5)
    m = cm.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=12, hid_size=H, activation_fn_cfg=SILU, device=DEV)
    mask = (rng.uniform(size=(O, E, O + A)) < 0.3).astype(np.float32)
    mask[np.arange(O), :, np.arange(O)] = 1.0
    m.set_input_mask(torch.from_numpy(mask))
    m.set_elite_members(list(range(12)))
    assert sum(p.numel() for p in m.parameters()) > 1.3e9
    obs = rng.normal(size=(E, B, O)).astype(np.float32)
    act = rng.uniform(-1, 1, size=(E, B, A)).astype(np.float32)
    with torch.no_grad():
        mean, logvar = m.forward(t(obs), t(act))
    assert mean.shape == logvar.shape == (E, B, O)
    sd = {k: v for k, v in m.state_dict().items()}
    for p_i, e_i in ((0, 0), (57, 3), (99, 15)):
        sub = {}
        for k, v in sd.items():  # one (sub-net, member) slice as a P = E = 1 ExternalMask model for the oracle
            if v.dim() == 4:
                sub[k] = v[p_i:p_i + 1, e_i:e_i + 1].cpu().numpy()
        sub["min_logvar"], sub["max_logvar"] = sd["min_logvar"][p_i:p_i + 1].cpu().numpy(), sd["max_logvar"][p_i:p_i + 1].cpu().numpy()
        x = np.concatenate([obs[e_i], act[e_i]], -1) * mask[p_i, e_i]
        h = x[None, None]
        for l in range(4):
            h = oc.activation(oc.ens_linear(h, sub["hidden_layers.%d.0.weight" % l], sub["hidden_layers.%d.0.bias" % l]), oc.ACT_SILU)
        raw = oc.ens_linear(h, sub["mean_and_logvar.weight"], sub["mean_and_logvar.bias"])[0, 0]
        ref_mean = raw[:, 0] + obs[e_i][:, p_i]
        ref_lv = oc.soft_clamp_logvar(raw[:, 1], sub["min_logvar"].reshape(-1)[0], sub["max_logvar"].reshape(-1)[0])
        assert rel_err(mean[e_i, :, p_i].cpu().numpy(), ref_mean) < TOL
        assert rel_err(logvar[e_i, :, p_i].cpu().numpy(), ref_lv) < TOL
    # rollout: 1 024 envs, selected-member evaluation on the ragged fp32 kernels == query-and-gather
    N = 1024
    dyn = cm.PlainEnsembleDynamics(m, learned_reward=False)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32),
               termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
               get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32))
    env.reset()
    a_dev = t(rng.uniform(-1, 1, size=(N, A)).astype(np.float32))
    idx = t(rng.choice(np.arange(12), size=N).astype(np.uint8))
    noise = t(rng.normal(size=(N, O)).astype(np.float32))
    sel = env.predict_selected(m, env._obs_dev, a_dev, idx, noise)
    with torch.no_grad():
        am, alv = m.forward(env._obs_dev.unsqueeze(0), a_dev.unsqueeze(0))
    assert torch.equal(sel, cm.ops.select_sample(am, alv, idx, noise))
    env.precision = "f16"  # hidden 512: the streamed-weights tensor-core kernel (variant 6)
    out16 = env.predict_selected(m, env._obs_dev, a_dev, idx, noise)
    assert rel_err(out16.cpu().numpy(), sel.cpu().numpy()) < TOL_TC
    o2, r2, d2, infos = env.step(rng.uniform(-1, 1, size=(N, A)).astype(np.float32))
    assert o2.shape == (N, O) and r2.shape == (N,) and d2.shape == (N,) and len(infos) == N and np.isfinite(o2).all()


# ---------------------------------------------------------------------------------------------------------------------------
# the CMI-test network (causal_discovery/CMI_test.py) on the grouped masked-GEMM kernels
# ---------------------------------------------------------------------------------------------------------------------------
def _build_cmi(cm, g, prefix="p.", **kw):
    p = params_of(g, prefix)
    m = cm.TransitionConditionalMutualInformationTest(4, 2, ensemble_num=5, elite_num=3,
                                                      hid_size=p["hidden_layers.0.0.weight"].shape[-1], device=DEV, **kw)
    load_sd(m, p)
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("cmi_relu", {}), ("cmi_silu", {"activation_fn_cfg": SILU}),
                                     ("cmi_noresidual", {"activation_fn_cfg": SILU, "residual": False})])
def test_cmi_test_network_forward_golden(cm, name, kw):
    g = load_golden(name + "_forward.npz")
    m = _build_cmi(cm, g, **kw)
    assert m.parallel_num == 7 and torch.equal(m.input_mask.cpu(), torch.from_numpy(g["input_mask"]))
    with torch.no_grad():
        mean, logvar = m.forward(t(g["obs"]), t(g["act"]))
        nll = m.get_nll_loss({"batch_obs": t(g["obs"]), "batch_action": t(g["act"])}, t(g["target"]))
        mean4, logvar4 = m.forward(t(g["obs4"]), t(g["act"]))
    assert mean.shape == logvar.shape == nll.shape == (7, 5, 30, 4)
    assert rel_err(mean.cpu().numpy(), g["mean"]) < TOL and rel_err(logvar.cpu().numpy(), g["logvar"]) < TOL
    assert rel_err(nll.cpu().numpy(), g["nll"]) < 3e-5
    assert rel_err(mean4.cpu().numpy(), g["mean4"]) < TOL and rel_err(logvar4.cpu().numpy(), g["logvar4"]) < TOL
    # leave-one-out property: sub-net p ignores input column p (the last sub-net sees everything)
    obs2 = g["obs"].copy()
    obs2[..., 1] += 3.0
    with torch.no_grad():
        mean2, _ = m.forward(t(obs2), t(g["act"]))
    res = 3.0 if m.residual else 0.0  # the residual adds the raw observation back
    d = (mean2 - mean).cpu().numpy()
    d[..., 1] -= res
    assert np.abs(d[1]).max() < 1e-5 and np.abs(d[0]).max() > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cmi_train.npz", "cmi_learnbounds_train.npz"])
def test_cmi_test_network_train_golden(cm, name):
    """three optimiser steps (NLL.mean().backward() + Adam with L2) against the reference run"""
    g = load_golden(name)
    m = _build_cmi(cm, g, "p0.", activation_fn_cfg=SILU, learn_logvar_bounds="learnbounds" in name)
    optim = cm.FusedAdamL2(m, lr=1e-4, weight_decay=1e-5, eps=1e-8)
    steps = int(g["steps"])
    for s in range(steps):
        loss = m.get_nll_loss({"batch_obs": t(g["obs%d" % s]), "batch_action": t(g["act%d" % s])}, t(g["target%d" % s]))
        optim.zero_grad()
        loss.mean().backward()
        if s == 0:
            for k, prm in m.named_parameters():
                if prm.requires_grad:
                    assert rel_err(prm.grad.cpu().numpy(), g["g0." + k]) < 1e-4, k
        optim.step()
        assert record_max("cmi train %s loss" % name, rel_err(loss.detach().cpu().numpy(), g["loss%d" % s])) < 3e-5
    ref = params_of(g, "p%d." % steps)
    sd = m.state_dict()
    for k in ref:
        mine = sd[k].cpu().numpy()
        if k.endswith("weight"):
            assert rel_err(mine, ref[k]) < TOL, k
        assert rel_err(mine - g["p0." + k], ref[k] - g["p0." + k]) < 2e-3 or np.array_equal(mine, ref[k]), k


def _host_term_env(cm, N, seed, max_steps):
    torch.manual_seed(3)
    tr = cm.PlainTransition(4, 2, hid_size=64, activation_fn_cfg=SILU, device=DEV)
    rw = cm.PlainRewardMech(4, 2, hid_size=64, activation_fn_cfg=SILU, device=DEV)
    for m in (tr, rw):
        m.set_elite_members([0, 2, 3, 5, 6])
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
    rng = np.random.default_rng(5)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, termination_fn=lambda n, o, a: np.abs(n[:, :1]) > 1.5,
               get_init_obs_fn=lambda n: rng.normal(size=(n, 4)).astype(np.float32), max_episode_steps=max_steps)
    env.rng_mode = "device"
    env.seed(seed)
    env.reset()
    return env


@pytest.mark.gpu
@pytest.mark.parametrize("make", [_rollout_env, _host_term_env], ids=["learned_termination", "host_termination_fn"])
def test_step_return_views_equals_copies(cm, make):
    """`return_views=True` hands out the pinned observation ring buffer instead of a copy: same values, and the array of
    step t stays intact while step t+1 runs (two-buffer ring), which is what SB3's collectors rely on."""
    N = 500
    env_a, env_b = make(cm, N, 4, max_steps=3), make(cm, N, 4, max_steps=3)
    env_b.return_views = True
    rng = np.random.default_rng(0)
    prev = None
    for t_ in range(5):
        act = rng.uniform(-1, 1, size=(N, 2)).astype(np.float32)
        oa, ra, da, ia = env_a.step(act)
        ob, rb, db, ib = env_b.step(act)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db)
        # the host mirror (next_obs read back before the tail, patched at the done rows fetched through `ops.gather_rows` or the
        # full copy) is the observation state in HBM, reset rows included; both fetch paths occur (few / many done envs)
        assert np.array_equal(oa, env_a._obs_dev.cpu().numpy()) and np.array_equal(ob, env_b._obs_dev.cpu().numpy())
        for i in np.nonzero(da)[0][:20]:
            assert np.array_equal(ia[i]["terminal_observation"], ib[i]["terminal_observation"])
        if prev is not None:
            assert np.array_equal(prev[0], prev[1])  # last step's view was not overwritten by this step
        prev = (oa.copy(), ob)


# ---------------------------------------------------------------------------------------------------------------------------
# tensor-core tier for WIDE hidden layers (streamed weights, variant 6)
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind,O,A,H,N", [("plain", 5, 1, 240, 700), ("plain", 20, 4, 256, 3000), ("reward", 9, 3, 320, 1500),
                                          ("extmask", 8, 3, 512, 2500), ("plain_relu", 30, 6, 512, 513)])
def test_tcgen05_wide_hidden_streamed_weights(cm, kind, O, A, H, N):
    """hidden 208..512 (a multiple of 16): the streamed-weights pair kernel against the fp32 tier on the same draws"""
    torch.manual_seed(1)
    rng = np.random.default_rng(1)
    act_cfg_ = None if kind == "plain_relu" else SILU
    if kind == "extmask":
        m = cm.ExternalMaskTransition(O, A, hid_size=H, activation_fn_cfg=act_cfg_, device=DEV)
        mask = (rng.uniform(size=(O, 7, O + A)) < 0.5).astype(np.float32)
        m.set_input_mask(torch.from_numpy(mask))
        dyn = cm.PlainEnsembleDynamics(m, learned_reward=False)
    elif kind == "reward":
        tr = cm.PlainTransition(O, A, hid_size=16, device=DEV)
        m = cm.PlainRewardMech(O, A, hid_size=H, activation_fn_cfg=act_cfg_, device=DEV)
        dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=m)
    else:
        m = cm.PlainTransition(O, A, hid_size=H, activation_fn_cfg=act_cfg_, device=DEV)
        dyn = cm.PlainEnsembleDynamics(m, learned_reward=False)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32), termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
               get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32))
    env.reset()
    act = t(rng.uniform(-1, 1, size=(N, A)).astype(np.float32))
    idx = t(rng.choice(np.asarray(m.elite_members), size=N).astype(np.uint8))
    noise = t(rng.normal(size=(N, m._head_dim)).astype(np.float32))
    env.precision = "fp32"
    ref = env.predict_selected(m, env._obs_dev, act, idx, noise)
    ref_m = env.predict_selected(m, env._obs_dev, act, idx, None)
    env.precision = "f16"
    out = env.predict_selected(m, env._obs_dev, act, idx, noise)
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < TOL_TC
    out_m = env.predict_selected(m, env._obs_dev, act, idx, None)
    base = env._obs_dev if kind != "reward" else 0.0
    assert rel_err((out_m - base).cpu().numpy(), (ref_m - base).cpu().numpy()) < 6e-3  # raw head output alone


# ---------------------------------------------------------------------------------------------------------------------------
# rollout -> optimiser step -> rollout: captured step graphs and the packed fp16 weights follow the parameters
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_rollouts_after_an_optimiser_step_use_the_new_weights(cm):
    """The fused Adam kernel writes the parameter arena through raw pointers (no torch version bump).  The packed fp16
    copies and the captured step graphs (which bake the copies' addresses) must follow all the same: after a step that
    moves every weight by ~0.05 away from zero, the tcgen05 tier -- graph-replayed device steps and host-facing steps
    alike -- agrees with the fp32 tier evaluated on the NEW weights, which is far away from what the old weights predict
    (at this model's initialisation one such step changes next_obs by ~1.5e-2 relative, the second by ~9e-2)."""
    N = 300
    env = _rollout_env(cm, N, 3, max_steps=1000)
    dyn = env.dynamics
    tr, tm = dyn.transition, dyn.termination_mech
    with torch.no_grad():  # a termination head that never fires: no resets, the returned observations are the predictions
        tm.mean_and_logvar.weight.zero_()
        tm.mean_and_logvar.bias.zero_()
    tm.deterministic = True
    for m in (tr, dyn.reward_mech, tm):
        m.set_elite_members([1, 1, 1, 1, 1])  # every env draws member 1: predictions are a function of the weights alone
    env.precision, env.deterministic = "f16", True
    rng = np.random.default_rng(0)
    act_h = rng.uniform(-1, 1, size=(N, 2)).astype(np.float32)
    act = t(act_h)
    idx = torch.ones(N, dtype=torch.uint8, device=DEV)

    def fp32_prediction():
        env.precision = "fp32"
        ref = env.predict_selected(tr, env._obs_dev.clone(), act, idx, None).clone()
        env.precision = "f16"
        return ref.cpu().numpy()

    def big_optimiser_step():
        opt = dyn.transition_optimizer
        opt.zero_grad()
        # zero gradient + NEGATIVE L2: g = -p, so Adam moves every weight by ~lr away from zero (the network's output grows)
        opt.param_groups[0].update(lr=0.05, weight_decay=-1.0)
        opt.step()

    # device-resident graph steps
    ref = fp32_prediction()
    term_obs = env.step_device_graphed(act)[4].cpu().numpy().copy()
    assert rel_err(term_obs, ref) < TOL_TC
    before = fp32_prediction()
    big_optimiser_step()
    after = fp32_prediction()
    assert rel_err(after, before) > 5 * TOL_TC  # the step really changed what the model predicts
    term_obs = env.step_device_graphed(act)[4].cpu().numpy().copy()
    assert rel_err(term_obs, after) < 2 * TOL_TC
    # host-facing steps (predict graph + tail graph)
    ref = fp32_prediction()
    obs_h = np.array(env.step(act_h)[0])
    assert rel_err(obs_h, ref) < TOL_TC
    before = fp32_prediction()
    big_optimiser_step()
    after = fp32_prediction()
    assert rel_err(after, before) > 5 * TOL_TC
    obs_h = np.array(env.step(act_h)[0])
    assert rel_err(obs_h, after) < 2 * TOL_TC


# ------------------------------------------------------------------------------------------------ round 2: tighter parity
def _fused_vs_oracle(cm, H, B, act_cfg_, seed):
    """One fused tcgen05 (3xTF32) training step of a PlainTransition against the oracle's manual backward + Adam."""
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    m = cm.PlainTransition(5, 1, hid_size=H, activation_fn_cfg=act_cfg_, device=DEV)
    with torch.no_grad():
        for prm in m.parameters():
            if prm.dim() == 3 and prm.shape[-2] == 1:
                prm.add_(0.1 * torch.randn_like(prm))
    p0 = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}
    E = 7
    obs = rng.normal(size=(E, B, 5)).astype(np.float32)
    act = rng.uniform(-1, 1, size=(E, B, 1)).astype(np.float32)
    tgt = (obs + 0.1 * rng.normal(size=(E, B, 5))).astype(np.float32)
    dyn = cm.PlainEnsembleDynamics(m, learned_reward=False)
    assert cm.ops.fused_train_plan(m, E, B) is not None
    loss = dyn.train_step("transition", t(obs), t(act), t(tgt))
    loss_ref, grads = oc.plain_nll_backward(p0, obs, act, tgt, act_kind=oc.ACT_RELU if act_cfg_ is None else oc.ACT_SILU)
    tag = "fused train H=%d B=%d %s" % (H, B, "relu" if act_cfg_ is None else "silu")
    assert record_max(tag + " loss", rel_err(loss.cpu().numpy(), loss_ref)) < TOL
    # the NLL of a dimension is a mean of signed terms and can sit near zero: an fp32 sum carries an absolute error of ~1e-6 of the
    # largest loss there, so the elementwise figure (floor 1e-3 of the largest) is recorded and bounded at 1e-5 / floor = 1e-2
    assert record_max(tag + " loss elem", elem_err(loss.cpu().numpy(), loss_ref)) < 1e-2
    for k, gr in grads.items():
        mine = dict(m.named_parameters())[k].grad.cpu().numpy()
        assert record_max(tag + " grad " + k, rel_err(mine, gr)) < 3e-5, k
    oc.adam_step(p0, grads, {})
    for k, v in m.state_dict().items():
        if k.endswith("weight") or k.endswith("bias"):
            assert record_max(tag + " weights after Adam", rel_err(v.cpu().numpy(), p0[k])) < TOL, k


@pytest.mark.parametrize("H,B,cfg", [(200, 256, SILU), (200, 300, SILU), (200, 1500, SILU), (64, 130, None), (24, 1, SILU), (239, 257, SILU),
                                     (200, 12000, SILU)])
def test_fused_tcgen05_train_step_vs_oracle(cm, H, B, cfg):
    """The reference batch size and shape (C1: 4 x 200, 256 rows per member), a ragged batch, a batch of several row tiles,
    ReLU, a single row, the widest supported hidden layer, and a batch large enough (94 row tiles) for the weight-gradient
    GEMMs to split it into slices whose partial sums a second launch adds up."""
    _fused_vs_oracle(cm, H, B, cfg, seed=H + B)


def test_fused_train_refuses_loudly_outside_its_shapes(cm):
    m = cm.PlainTransition(5, 1, hid_size=256, device=DEV)  # padded width 272: two accumulators do not fit tensor memory
    assert cm.ops.fused_train_plan(m, 7, 64) is None  # -> the per-layer kernels run
    with pytest.raises(cm._lib.CmrlB200Error):
        cm._lib.call("cmrl_train_fused", None, 4, 1, 7, 64, 5, 1, 256, 5, 0, 2, 1, 3, None, 0, None, 0, None, 0, 5, None, None, 5,
                     0.0, 1.0, None, None, None)


def _selected_means(cm, env, model, obs, act, noise_value):
    """[E,N,D]: every member's prediction through the selected-member rollout path (all envs draw member e)."""
    E, N = model.ensemble_num, obs.shape[0]
    outs = []
    for e in range(E):
        idx = torch.full((N,), e, dtype=torch.uint8, device=DEV)
        noise = None if noise_value is None else torch.full((N, model._head_dim), float(noise_value), device=DEV)
        outs.append(env.predict_selected(model, obs, act, idx, noise).clone())
    return torch.stack(outs).cpu().numpy()


@pytest.mark.parametrize("name", ["plain_c1_forward.npz", "extmask_2d_forward.npz", "extmask_3de_forward.npz", "extmask_wide_forward.npz",
                                  "reward_forward.npz"])
def test_f16_rollout_tier_directly_against_reference_goldens(cm, name):
    """The tcgen05 f16 rollout kernel against the UNMODIFIED reference's outputs (not against the repo's own fp32 tier):
    mean, and mean + sqrt(exp(logvar)) (noise = 1), for every member; normwise and elementwise errors are recorded."""
    g = load_golden(name)
    p = params_of(g)
    E, B, O = g["obs"].shape
    A = g["act"].shape[-1]
    H = p["hidden_layers.0.0.weight"].shape[-1]
    if name.startswith("extmask"):
        m = cm.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=min(5, E - 1), hid_size=H, activation_fn_cfg=SILU, device=DEV)
        load_sd(m, p)
        m.set_input_mask(torch.from_numpy(g["input_mask"]))
    elif name.startswith("reward"):
        m = cm.PlainRewardMech(O, A, ensemble_num=E, hid_size=H, activation_fn_cfg=SILU, device=DEV)
        load_sd(m, p)
    else:
        m = cm.PlainTransition(O, A, ensemble_num=E, hid_size=H, activation_fn_cfg=SILU, device=DEV)
        load_sd(m, p)
    dyn = cm.PlainEnsembleDynamics(m if not name.startswith("reward") else cm.PlainTransition(O, A, hid_size=8, device=DEV), learned_reward=False)
    env = cm.VecFakeEnv(B, None, None)
    env.set_up(dyn, reward_fn=lambda n, o, a: np.zeros((n.shape[0], 1), np.float32), termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool),
               get_init_obs_fn=lambda n: np.zeros((n, O), np.float32))
    env.precision = "f16"
    # the goldens hold per-member inputs [E,B,*]; member e's rows go through with every env drawing member e
    means, samples = [], []
    for e in range(E):
        obs, act = t(g["obs"][e]), t(g["act"][e])
        idx = torch.full((B,), e, dtype=torch.uint8, device=DEV)
        means.append(env.predict_selected(m, obs, act, idx, None).cpu().numpy())
        samples.append(env.predict_selected(m, obs, act, idx, torch.ones((B, m._head_dim), device=DEV)).cpu().numpy())
    means, samples = np.stack(means), np.stack(samples)
    ref_sample = g["mean"] + np.sqrt(np.exp(g["logvar"]))
    assert record_max("f16 tier vs golden %s mean" % name, rel_err(means, g["mean"])) < 1e-3
    assert record_max("f16 tier vs golden %s mean+std" % name, rel_err(samples, ref_sample)) < 1e-3
    record_max("f16 tier vs golden %s mean elem" % name, elem_err(means, g["mean"]))
    # the network's own contribution (residual removed): the 11-bit operand error is visible here
    if hasattr(m, "residual") and m.residual:
        net, net_ref = means - g["obs"][..., :means.shape[-1]], g["mean"] - g["obs"][..., :means.shape[-1]]
        assert record_max("f16 tier vs golden %s network output (no residual)" % name, rel_err(net, net_ref)) < 5e-3


def _load_euler(cm, g, prefix):
    O, A, H = 5, 1, g[prefix + "hidden_layers.0.0.weight"].shape[-1]
    one = cm.PlainTransition(O, A, hid_size=H, activation_fn_cfg=SILU, device=DEV)
    load_sd(one, params_of(g, prefix))
    return one


def test_forward_euler_forward_and_gradients_golden(cm):
    """ForwardEulerTransition (forward_euler.py:32-41), k = 3: forward, NLL, gradients w.r.t. the inputs and the one-step
    model's parameters against the unmodified reference's autograd."""
    g = load_golden("euler_forward.npz")
    one = _load_euler(cm, g, "p.")
    fe = cm.ForwardEulerTransition(one, repeat_times=int(g["repeat_times"]))
    obs, act = t(g["obs"]).requires_grad_(True), t(g["act"]).requires_grad_(True)
    mean, logvar = fe.forward(obs, act)
    assert record_max("euler k=3 mean", rel_err(mean.detach().cpu().numpy(), g["mean"])) < TOL
    assert record_max("euler k=3 logvar", rel_err(logvar.detach().cpu().numpy(), g["logvar"])) < TOL
    nll = fe.get_nll_loss({"batch_obs": obs, "batch_action": act}, t(g["target"]))
    assert record_max("euler k=3 nll", rel_err(nll.detach().cpu().numpy(), g["nll"])) < 3e-5
    one.zero_grad()
    nll.mean().backward()
    assert record_max("euler k=3 d/d obs", rel_err(obs.grad.cpu().numpy(), g["g_obs"])) < 1e-4
    assert record_max("euler k=3 d/d action", rel_err(act.grad.cpu().numpy(), g["g_act"])) < 1e-4
    for k, prm in one.named_parameters():
        if prm.requires_grad:
            assert record_max("euler k=3 grad", rel_err(prm.grad.cpu().numpy(), g["g." + k])) < 1e-4, k


def test_forward_euler_train_golden(cm):
    """Three Adam steps THROUGH the wrapper (BaseDynamics.train_step takes the autograd path for it), k = 2."""
    g = load_golden("euler_train.npz")
    one = _load_euler(cm, g, "p0.one_step_transition.")
    fe = cm.ForwardEulerTransition(one, repeat_times=int(g["repeat_times"]))
    dyn = cm.PlainEnsembleDynamics(fe, learned_reward=False)
    steps = int(g["steps"])
    for s in range(steps):
        loss = dyn.train_step("transition", t(g["obs%d" % s]), t(g["act%d" % s]), t(g["target%d" % s]))
        assert record_max("euler train loss", rel_err(loss.detach().cpu().numpy(), g["loss%d" % s])) < 3e-5
    ref = params_of(g, "p%d.one_step_transition." % steps)
    sd = one.state_dict()
    for k in ref:
        if k.endswith("weight"):
            assert record_max("euler train weights", rel_err(sd[k].cpu().numpy(), ref[k])) < TOL, k
            assert rel_err(sd[k].cpu().numpy() - g["p0.one_step_transition." + k], ref[k] - g["p0.one_step_transition." + k]) < 2e-3, k
    # the device-resident epoch falls back to `train` for the wrapper (no layer stack of its own) instead of failing
    from oracle.refshim import DuckReplayBuffer

    rng = np.random.default_rng(0)
    buf = DuckReplayBuffer(rng.normal(size=(96, 5)).astype(np.float32), rng.uniform(-1, 1, size=(96, 1)).astype(np.float32),
                           rng.normal(size=(96, 5)).astype(np.float32), np.zeros(96, np.float32), np.zeros(96, bool))
    dyn.learn(buf, batch_size=32, longest_epoch=1, patience=1)


@pytest.mark.parametrize("precision", ["fp32", "f16"])
def test_forward_euler_fake_env_golden(cm, precision):
    """VecFakeEnv.step with the wrapper as the transition: the drawn member is re-applied k times (its own elite set, as in the
    reference: forward_euler.py:29-30 forwards set_elite_members, get_random_index reads the wrapper's)."""
    g = load_golden("fake_env_euler.npz")
    one = _load_euler(cm, g, "tr.")
    rw = cm.PlainRewardMech(5, 1, hid_size=g["rw.hidden_layers.0.0.weight"].shape[-1], activation_fn_cfg=SILU, device=DEV)
    load_sd(rw, params_of(g, "rw."))
    fe = cm.ForwardEulerTransition(one, repeat_times=int(g["repeat_times"]))
    fe._elite_members = list(g["fe_elite"])
    one.set_elite_members(list(g["tr_elite"]))
    rw.set_elite_members(list(g["rw_elite"]))
    dyn = cm.PlainEnsembleDynamics(fe, learned_reward=True, reward_mech=rw)
    N = g["reset_obs"].shape[0]
    init_rng = np.random.default_rng(78)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, reward_fn=None, termination_fn=lambda n, o, a: np.abs(n[:, 0:1]) > 1.5,
               get_init_obs_fn=lambda n: init_rng.normal(size=(n, 5)).astype(np.float32))
    env.precision = precision
    env.seed(13)
    assert np.array_equal(env.reset(), g["reset_obs"])
    tol = TOL if precision == "fp32" else 1e-3
    for s in range(int(g["T"])):
        obs, rew, done, infos = env.step(g["actions"][s])
        if precision == "fp32":
            assert np.array_equal(done, g["done%d" % s])
        if np.array_equal(done, g["done%d" % s]):
            assert record_max("euler fake env %s obs" % precision, rel_err(obs, g["obs%d" % s])) < tol
            assert record_max("euler fake env %s reward" % precision, rel_err(rew, g["rew%d" % s])) < tol
            term = np.stack([infos[i]["terminal_observation"] if done[i] else np.zeros(5, np.float32) for i in range(N)])
            assert rel_err(term, g["termobs%d" % s]) < tol
        else:  # a 1e-3 perturbation moved an env across the termination threshold: later steps are not comparable
            break


def test_net_shards_train_like_the_full_model(cm):
    """SURVEY 8e row 2 (net sharding): the sub-networks of an ExternalMaskTransition are independent, so shards holding
    dimensions [0,3) and [3,5) -- trained separately, no exchange -- reproduce the full model's slices (same kernels per net;
    the loss scale uses the FULL dimension count), and the shards' hold-out sums add up to the full model's score."""
    torch.manual_seed(5)
    rng = np.random.default_rng(5)
    P, A, E, B, H = 5, 1, 7, 96, 48
    mask = (rng.uniform(size=(P, P + A)) < 0.6).astype(np.float32)
    mask[np.arange(P), np.arange(P)] = 1
    full = cm.ExternalMaskTransition(P, A, hid_size=H, activation_fn_cfg=SILU, device=DEV)
    full.set_input_mask(torch.from_numpy(mask))
    shards = []
    for p0, p1 in ((0, 3), (3, 5)):
        sh = cm.ExternalMaskTransition(P, A, hid_size=H, activation_fn_cfg=SILU, device=DEV, subnet_range=(p0, p1))
        sh.set_input_mask(torch.from_numpy(mask))
        with torch.no_grad():
            for q_full, q_sh in zip(full.parameters(), sh.parameters()):
                q_sh.copy_(q_full[p0:p1])
        shards.append(sh)
    dyn_full = cm.PlainEnsembleDynamics(full, learned_reward=False, optim_lr=1e-3)
    dyn_sh = [cm.PlainEnsembleDynamics(sh, learned_reward=False, optim_lr=1e-3) for sh in shards]
    cm.ops.TRAIN_FUSED, keep = False, cm.ops.TRAIN_FUSED  # the full model on the per-layer kernels too: same arithmetic per net
    try:
        for step in range(3):
            obs = t(rng.normal(size=(E, B, P)).astype(np.float32))
            act = t(rng.uniform(-1, 1, size=(E, B, A)).astype(np.float32))
            tgt = obs + 0.1 * torch.tanh(obs)
            loss_full = dyn_full.train_step("transition", obs, act, tgt)
            for sh, d in zip(shards, dyn_sh):
                loss_sh = d.train_step("transition", obs, act, sh.target_columns(tgt).contiguous())
                p0, p1 = sh.subnet_range
                assert rel_err(loss_sh.cpu().numpy(), loss_full[..., p0:p1].cpu().numpy()) < TOL
    finally:
        cm.ops.TRAIN_FUSED = keep
    for sh in shards:
        p0, p1 = sh.subnet_range
        for (name, q_full), (_, q_sh) in zip(full.named_parameters(), sh.named_parameters()):
            assert record_max("net shard vs full model weights", rel_err(q_sh.detach().cpu().numpy(), q_full[p0:p1].detach().cpu().numpy())) < 1e-6, name
    # forward of a shard = its columns of the full forward (residual included)
    obs, act = t(rng.normal(size=(E, 33, P)).astype(np.float32)), t(rng.uniform(-1, 1, size=(E, 33, A)).astype(np.float32))
    with torch.no_grad():
        m_full, lv_full = full.forward(obs, act)
        for sh in shards:
            p0, p1 = sh.subnet_range
            m_sh, lv_sh = sh.forward(obs, act)
            assert rel_err(m_sh.cpu().numpy(), m_full[..., p0:p1].cpu().numpy()) < 1e-6
            assert rel_err(lv_sh.cpu().numpy(), lv_full[..., p0:p1].cpu().numpy()) < 1e-6
    # hold-out score: the shards' [E] sums over their dimensions add up to the full model's
    from cmrl_b200.transition_iterator import TransitionIterator
    from cmrl_b200.types import InteractionBatch

    n = 300
    o = rng.normal(size=(n, P)).astype(np.float32)
    a = rng.uniform(-1, 1, size=(n, A)).astype(np.float32)
    val = TransitionIterator(InteractionBatch(o, a, (o + 0.1 * np.tanh(o)).astype(np.float32), np.zeros(n, np.float32), np.zeros(n, bool)), 128)
    score_full = dyn_full.evaluate_member_mse(val, "transition").numpy().astype(np.float64)
    parts = [d.evaluate_member_mse(val, "transition").numpy().astype(np.float64) * (sh.subnet_range[1] - sh.subnet_range[0])
             for sh, d in zip(shards, dyn_sh)]
    assert rel_err(sum(parts) / P, score_full) < TOL
    with pytest.raises(cm._lib.CmrlB200Error):
        shards[0].packed_f16()


def test_small_training_kernels_round2(cm):
    """Deterministic fp64 loss accumulation (one block and many), Adam on sizes / offsets that leave the 16-byte path (tail of
    1-3 elements, unaligned arenas) with the step counter advanced by the kernel's last block, and the gradient exchange of a
    single-rank communicator (plain Adam through the exchange kernel)."""
    import ctypes

    rng = np.random.default_rng(11)
    for n in (1, 1000, 8960, 300_001):
        x = rng.normal(size=n).astype(np.float32)
        acc = torch.full((), 2.5, device=DEV, dtype=torch.float64)
        ws = torch.zeros(66, device=DEV, dtype=torch.float64)
        for _ in range(3):  # the arrival ticket of the workspace resets itself
            cm.ops.sum_acc_f64(t(x), acc, ws)
        ref = 2.5 + 3 * float(np.sum(x.astype(np.float64)))
        assert abs(float(acc.item()) - ref) <= 1e-12 * max(1.0, abs(ref)) * n ** 0.5 + 1e-12
    for n, off in ((7, 0), (1026, 0), (4099, 1), (65_537, 3)):
        base = [torch.zeros(n + 8, device=DEV) for _ in range(4)]
        p_, g_, m_, v_ = (b[off:off + n] for b in base)
        p0, g0 = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
        p_.copy_(t(p0)), g_.copy_(t(g0))
        step = torch.tensor([4, 0], dtype=torch.int32, device=DEV)
        cm.ops.adam_l2_step_dev(p_, g_, m_, v_, 1e-3, 0.9, 0.999, 1e-8, 1e-5, step, 0.5)
        # closed form of one Adam step from zero moments at step t = 5 (torch.optim.Adam, L2-coupled decay)
        gi = 0.5 * g0.astype(np.float64) + 1e-5 * p0.astype(np.float64)
        m1, v1 = 0.1 * gi, 0.001 * gi * gi
        upd = (1e-3 / (1 - 0.9 ** 5)) * m1 / (np.sqrt(v1) / np.sqrt(1 - 0.999 ** 5) + 1e-8)
        assert rel_err(p_.cpu().numpy(), p0 - upd) < 1e-6
        assert step.cpu().tolist() == [5, 0]
        assert float(base[0][:off].abs().sum()) == 0.0 and float(base[0][off + n:].abs().sum()) == 0.0  # nothing outside the arena
    # single-rank communicator: the exchange kernel degenerates to Adam + counter update
    comm, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
    n = 10_003
    cm._lib.call("cmrl_dp_comm_create", 0, 1, n, ctypes.byref(comm), handle)
    try:
        p_, g_, m_, v_ = t(rng.normal(size=n).astype(np.float32)), t(rng.normal(size=n).astype(np.float32)), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        q_, mq, vq = p_.clone(), m_.clone(), v_.clone()
        s1, s2 = torch.tensor([2, 0], dtype=torch.int32, device=DEV), torch.tensor([2, 0], dtype=torch.int32, device=DEV)
        g_keep = g_.clone()
        cm._lib.call("cmrl_allreduce_grads_adam", comm, p_.data_ptr(), g_.data_ptr(), m_.data_ptr(), v_.data_ptr(), n, 1e-3, 0.9, 0.999,
                     1e-8, 1e-5, 1.0, s1.data_ptr(), torch.cuda.current_stream().cuda_stream)
        cm.ops.adam_l2_step_dev(q_, g_keep, mq, vq, 1e-3, 0.9, 0.999, 1e-8, 1e-5, s2, 1.0)
        assert torch.equal(p_, q_) and torch.equal(m_, mq) and torch.equal(v_, vq) and torch.equal(g_, g_keep)
        assert s1.cpu().tolist() == [3, 0] == s2.cpu().tolist()
        err = ctypes.c_int(7)
        cm._lib.call("cmrl_dp_comm_error", comm, ctypes.byref(err))
        assert err.value == 0
    finally:
        cm._lib.call("cmrl_dp_comm_destroy", comm)


@pytest.mark.parametrize("kind", ["plain", "extmask"])
def test_all_member_means_on_the_tensor_core_tiers(cm, kind):
    """MOPO penalty path (fake_env.py:186-193): every member's mean for every env from ONE pass of the fused rollout kernels over
    all-member buckets ([E,N,D] output through the member stride), against the all-member CUDA-core forward; the penalty and the
    sampled next observation of a penalised step follow."""
    torch.manual_seed(21)
    rng = np.random.default_rng(21)
    O, A, N = 5, 1, 3001
    if kind == "plain":
        tr = cm.PlainTransition(O, A, hid_size=200, activation_fn_cfg=SILU, device=DEV)
    else:
        tr = cm.ExternalMaskTransition(O, A, hid_size=200, activation_fn_cfg=SILU, device=DEV)
        mask = (rng.uniform(size=(O, O + A)) < 0.6).astype(np.float32)
        mask[np.arange(O), np.arange(O)] = 1
        tr.set_input_mask(torch.from_numpy(mask))
    rw = cm.PlainRewardMech(O, A, hid_size=200, activation_fn_cfg=SILU, device=DEV)
    dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw)
    env = cm.VecFakeEnv(N, None, None)
    env.set_up(dyn, termination_fn=lambda n, o, a: np.zeros((n.shape[0], 1), bool), get_init_obs_fn=lambda n: rng.normal(size=(n, O)).astype(np.float32),
               penalty_coeff=0.7)
    env.reset()
    obs = env._obs_dev.clone()
    act = t(rng.uniform(-1, 1, size=(N, A)).astype(np.float32))
    with torch.no_grad():
        ref_mean, _ = tr.forward(obs.unsqueeze(0), act.unsqueeze(0))
    ref_mean = ref_mean.cpu().numpy()
    ref_pen = oc.mopo_penalty(ref_mean)
    for precision, tol in (("fp32", TOL), ("f16", TOL_TC)):
        env.precision = precision
        means = env._all_member_means(tr, obs, act)
        assert means is not None and tuple(means.shape) == (7, N, O)
        assert record_max("all-member means %s %s" % (kind, precision), rel_err(means.cpu().numpy(), ref_mean)) < tol
        pen = cm.ops.mopo_penalty(means).cpu().numpy()
        assert record_max("MOPO penalty %s %s" % (kind, precision), rel_err(pen, ref_pen)) < (1e-4 if precision == "fp32" else 5e-2)
    env.precision = "ffma"
    assert env._all_member_means(tr, obs, act) is None  # the CUDA-core tier keeps the all-member forward
    # (the penalised STEP through this path is pinned by the golden trajectory fake_env_penalty.npz of the real reference,
    # which the default precision now runs through these kernels)


def test_deterministic_mechanism_trains_through_the_mse_kernels(cm):
    """Deterministic mechanisms train on `get_mse_loss` (mlp.py:68-70): the loss and its gradient kernel against the closed form
    2 (mean - target) / numel applied by torch autograd to the same forward output, and one optimiser step of the dynamics."""
    torch.manual_seed(8)
    m = cm.PlainTransition(5, 1, hid_size=32, deterministic=True, activation_fn_cfg=SILU, device=DEV)
    obs, act = torch.randn(7, 40, 5, device=DEV), torch.rand(7, 40, 1, device=DEV)
    tgt = obs + 0.1 * torch.tanh(obs)
    loss = m.get_mse_loss({"batch_obs": obs, "batch_action": act}, tgt)
    mean, _ = m.forward(obs, act)
    assert rel_err(loss.detach().cpu().numpy(), ((mean - tgt) ** 2).detach().cpu().numpy()) < 1e-6
    g = torch.randn_like(loss)
    got = cm.ops.mse_loss_bwd(mean.detach(), tgt, g)
    assert rel_err(got.cpu().numpy(), (2.0 * (mean.detach() - tgt) * g).cpu().numpy()) < 1e-6
    # shared (stride-0) targets, as `evaluate` passes them
    tgt_shared = tgt[0].unsqueeze(0).expand(7, -1, -1)
    got = cm.ops.mse_loss_bwd(mean.detach(), tgt_shared, g)
    assert rel_err(got.cpu().numpy(), (2.0 * (mean.detach() - tgt_shared) * g).cpu().numpy()) < 1e-6
    dyn = cm.PlainEnsembleDynamics(m, learned_reward=False, optim_lr=1e-3)
    before = m.flat_parameters()[0].clone()
    l0 = dyn.train_step("transition", obs, act, tgt).mean().item()
    for _ in range(20):
        l1 = dyn.train_step("transition", obs, act, tgt).mean().item()
    assert l1 < l0 and not torch.equal(before, m.flat_parameters()[0])


@pytest.mark.parametrize("learned_termination", [False, True])
def test_grouped_mechanism_launch_equals_separate_launches(cm, learned_termination):
    """Device-resident f16 steps run transition + reward (+ termination) mechs in ONE launch of the fused kernel
    (cmrl_mlp_rollout_f16_groups): same tiles, same arithmetic -- trajectories are bit-identical to one launch per mechanism,
    kernel by kernel and graph-replayed, for the causal-mask (parallel head) and plain transitions."""
    for kind in ("extmask", "plain"):
        torch.manual_seed(31)
        rng = np.random.default_rng(31)
        O, A, N = 5, 1, 20_011
        if kind == "extmask":
            tr = cm.ExternalMaskTransition(O, A, hid_size=200, activation_fn_cfg=SILU, device=DEV)
            mask = (rng.uniform(size=(O, O + A)) < 0.6).astype(np.float32)
            mask[np.arange(O), np.arange(O)] = 1
            tr.set_input_mask(torch.from_numpy(mask))
        else:
            tr = cm.PlainTransition(O, A, hid_size=200, activation_fn_cfg=SILU, device=DEV)
        rw = cm.PlainRewardMech(O, A, hid_size=200, activation_fn_cfg=SILU, device=DEV)
        tm = cm.PlainTerminationMech(O, A, hid_size=200, activation_fn_cfg=SILU, device=DEV) if learned_termination else None
        dyn = cm.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw, learned_termination=learned_termination, termination_mech=tm)
        init = rng.normal(size=(N, O)).astype(np.float32)
        act = t(rng.uniform(-1, 1, size=(N, A)).astype(np.float32))
        runs = {}
        for grouped in (True, False):
            for graphed in (False, True):
                env = cm.VecFakeEnv(N, None, None)
                env.set_up(dyn, termination_fn=None if learned_termination else (lambda n, o, a: np.zeros((n.shape[0], 1), bool)),
                           get_init_obs_fn=lambda n: init[:n].copy(), max_episode_steps=3)
                env.rng_mode, env.precision, env.group_mechs = "device", "f16", grouped
                env.seed(5)
                env.reset()
                rec = []
                for _ in range(4):
                    out = env.step_device_graphed(act) if graphed else env.step_device(act)
                    rec.append([x.clone() for x in out])
                runs[(grouped, graphed)] = rec
        ref = runs[(False, False)]
        for key, rec in runs.items():
            for a_step, b_step in zip(rec, ref):
                for x, y in zip(a_step, b_step):
                    assert torch.equal(x, y), (kind, key)
        assert float(ref[-1][0].abs().sum()) > 0
        if not learned_termination:  # (a learned termination mech ends every episode at once: "non-zero => done", reference quirk 2)
            assert bool(ref[2][3].any())  # the step limit truncated episodes
