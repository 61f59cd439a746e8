"""The numpy oracle against the LIVE reference (imported unmodified through oracle/refshim.py) on randomly drawn
configurations the committed fixtures do not hold: layer counts, widths, ensemble sizes, batch sizes down to one row,
both activations, residual on/off, every mask rank.  Forward, un-reduced NLL and the gradient of `loss.mean()` (torch
autograd on the reference side, the oracle's manual backward on ours).  Runs where /root/reference is mounted (the build
container); skipped on the GPU box, where the committed fixtures are the pin."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import cmrl_oracle as oc
from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted (GPU box)")

TOL = 5e-5  # fp32 numpy restatement vs torch fp32 (summation order differs)
SILU = {"_target_": "torch.nn.SiLU"}


def _np_params(module):
    return {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        yield dict(O=int(rng.integers(2, 10)), A=int(rng.integers(1, 5)), H=int(rng.choice([8, 33, 64])),
                   L=int(rng.choice([1, 2, 4])), E=int(rng.choice([2, 3, 7])), B=int(rng.choice([1, 17, 50])),
                   silu=bool(rng.integers(0, 2)), residual=bool(rng.integers(0, 2)), seed=int(rng.integers(1 << 30)))


def _reference_grads(model, obs, act, target):
    model_in = {"batch_obs": torch.from_numpy(obs.copy()), "batch_action": torch.from_numpy(act.copy())}
    loss = model.get_nll_loss(model_in, torch.from_numpy(target))
    model.zero_grad()
    loss.mean().backward()
    return loss.detach().numpy(), {k: p.grad.numpy().copy() for k, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("case", list(_cases(8, 11)), ids=lambda c: "O%(O)dA%(A)dH%(H)dL%(L)dE%(E)dB%(B)d" % c)
def test_plain_transition_forward_loss_and_gradient(case):
    ref = refshim.load_reference()
    c = case
    torch.manual_seed(c["seed"])
    rng = np.random.default_rng(c["seed"])
    m = ref.PlainTransition(c["O"], c["A"], ensemble_num=c["E"], elite_num=min(2, c["E"]), num_layers=c["L"], hid_size=c["H"],
                            residual=c["residual"], activation_fn_cfg=SILU if c["silu"] else None)
    kind = oc.ACT_SILU if c["silu"] else oc.ACT_RELU
    obs = rng.normal(size=(c["E"], c["B"], c["O"])).astype(np.float32)
    act = rng.uniform(-1, 1, size=(c["E"], c["B"], c["A"])).astype(np.float32)
    target = (obs + 0.1 * rng.normal(size=obs.shape)).astype(np.float32)
    p = _np_params(m)
    with torch.no_grad():
        mean_r, logvar_r = m.forward(torch.from_numpy(obs.copy()), torch.from_numpy(act.copy()))
    mean, logvar = oc.plain_transition_forward(p, obs, act, kind, num_layers=c["L"], residual=c["residual"])
    assert rel_err(mean, mean_r.numpy()) < TOL and rel_err(logvar, logvar_r.numpy()) < TOL
    loss_r, grads_r = _reference_grads(m, obs, act, target)
    loss, grads = oc.plain_nll_backward(p, obs, act, target, kind, num_layers=c["L"], residual=c["residual"])
    assert rel_err(loss, loss_r) < TOL
    assert set(grads) == set(grads_r)
    for k in grads_r:
        assert grads[k].shape == grads_r[k].shape and rel_err(grads[k], grads_r[k]) < 2e-4, k


@pytest.mark.parametrize("mask_rank", ["2d", "3de", "3db", "4d"])
@pytest.mark.parametrize("case", list(_cases(3, 23)), ids=lambda c: "O%(O)dA%(A)dH%(H)dL%(L)dE%(E)dB%(B)d" % c)
def test_external_mask_transition_forward_loss_and_gradient(case, mask_rank):
    ref = refshim.load_reference()
    c = case
    torch.manual_seed(c["seed"])
    rng = np.random.default_rng(c["seed"])
    O, A, E, B = c["O"], c["A"], c["E"], c["B"]
    m = ref.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=min(2, E), num_layers=c["L"], hid_size=c["H"],
                                   residual=c["residual"], activation_fn_cfg=SILU if c["silu"] else None)
    kind = oc.ACT_SILU if c["silu"] else oc.ACT_RELU
    shape = {"2d": (O, O + A), "3de": (O, E, O + A), "3db": (O, B, O + A), "4d": (O, E, B, O + A)}[mask_rank]
    if mask_rank == "3db" and B == E:
        pytest.skip("a [P,B,in] mask with B == E is read as [P,E,in] by the reference")
    mask = (rng.uniform(size=shape) < 0.6).astype(np.float32)
    m.set_input_mask(torch.from_numpy(mask.copy()))
    obs = rng.normal(size=(E, B, O)).astype(np.float32)
    act = rng.uniform(-1, 1, size=(E, B, A)).astype(np.float32)
    target = (obs + 0.1 * rng.normal(size=obs.shape)).astype(np.float32)
    p = _np_params(m)
    with torch.no_grad():
        mean_r, logvar_r = m.forward(torch.from_numpy(obs.copy()), torch.from_numpy(act.copy()))
    mean, logvar = oc.external_mask_transition_forward(p, mask, obs, act, kind, c["L"], c["residual"])[:2]
    assert rel_err(mean, mean_r.numpy()) < TOL and rel_err(logvar, logvar_r.numpy()) < TOL
    loss_r, grads_r = _reference_grads(m, obs, act, target)
    loss, grads = oc.external_mask_nll_backward(p, mask, obs, act, target, kind, c["L"], c["residual"])
    assert rel_err(loss, loss_r) < TOL
    for k in grads_r:
        assert grads[k].shape == grads_r[k].shape and rel_err(grads[k], grads_r[k]) < 2e-4, k


@pytest.mark.parametrize("cls", ["PlainRewardMech", "PlainTerminationMech"])
@pytest.mark.parametrize("case", list(_cases(3, 37)), ids=lambda c: "O%(O)dA%(A)dH%(H)dL%(L)dE%(E)dB%(B)d" % c)
def test_reward_and_termination_mech_forward_loss_and_gradient(case, cls):
    ref = refshim.load_reference()
    c = case
    torch.manual_seed(c["seed"])
    rng = np.random.default_rng(c["seed"])
    m = getattr(ref, cls)(c["O"], c["A"], ensemble_num=c["E"], elite_num=min(2, c["E"]), num_layers=c["L"], hid_size=c["H"],
                          activation_fn_cfg=SILU if c["silu"] else None)
    kind = oc.ACT_SILU if c["silu"] else oc.ACT_RELU
    obs = rng.normal(size=(c["E"], c["B"], c["O"])).astype(np.float32)
    act = rng.uniform(-1, 1, size=(c["E"], c["B"], c["A"])).astype(np.float32)
    target = rng.normal(size=(c["E"], c["B"], 1)).astype(np.float32)
    p = _np_params(m)
    with torch.no_grad():
        mean_r, logvar_r = m.forward(torch.from_numpy(obs.copy()), torch.from_numpy(act.copy()))
    mean, logvar = oc.reward_mech_forward(p, obs, act, kind, num_layers=c["L"])
    assert rel_err(mean, mean_r.numpy()) < TOL and rel_err(logvar, logvar_r.numpy()) < TOL
    loss_r, grads_r = _reference_grads(m, obs, act, target)
    loss, grads = oc.plain_nll_backward(p, obs, act, target, kind, num_layers=c["L"], kind="reward")
    assert rel_err(loss, loss_r) < TOL
    for k in grads_r:
        assert grads[k].shape == grads_r[k].shape and rel_err(grads[k], grads_r[k]) < 2e-4, k


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_adam_l2_steps_match_torch_adam_as_the_reference_constructs_it(seed):
    """base_dynamics.py:58-63: torch.optim.Adam(lr, weight_decay (L2, coupled), eps) over a mech's parameters; several steps on
    tensors of assorted shapes, against oc.adam_step."""
    rng = np.random.default_rng(seed)
    shapes = [(7, 6, 20), (7, 1, 20), (3, 4, 5, 2), (1,), (2, 1, 1, 1)]
    lr, wd, eps = float(rng.choice([1e-4, 1e-3])), float(rng.choice([1e-5, 5e-4])), 1e-8
    tensors = [torch.nn.Parameter(torch.from_numpy(rng.normal(size=s).astype(np.float32))) for s in shapes]
    optim = torch.optim.Adam(tensors, lr=lr, weight_decay=wd, eps=eps)
    params = {str(i): t.detach().numpy().copy() for i, t in enumerate(tensors)}
    state = {}
    for _ in range(5):
        grads = {str(i): rng.normal(size=s).astype(np.float32) for i, s in enumerate(shapes)}
        for i, t in enumerate(tensors):
            t.grad = torch.from_numpy(grads[str(i)].copy())
        optim.step()
        oc.adam_step(params, grads, state, lr=lr, weight_decay=wd, eps=eps)
    for i, t in enumerate(tensors):
        assert rel_err(params[str(i)], t.detach().numpy()) < 1e-6, i


@pytest.mark.parametrize("seed", [0, 1])
def test_rollout_sampling_and_penalty_match_the_reference_fake_env(seed):
    """fake_env.py:186-215: `get_dynamics_predict` (member draw + Gaussian sample from the env's generator) and `get_penalty`."""
    ref = refshim.load_reference()
    rng = np.random.default_rng(seed)
    E, B, D = 7, 64, 5
    mean = rng.normal(size=(E, B, D)).astype(np.float32)
    logvar = rng.uniform(-6, 0.5, size=(E, B, D)).astype(np.float32)
    env = ref.VecFakeEnv(B, None, None)
    mlp = ref.EnsembleMLP(ensemble_num=E, elite_num=5)
    mlp.set_elite_members([0, 2, 3, 5, 6])
    env.dynamics = type("D", (), {"transition": mlp, "get_variable_by_mech": staticmethod(lambda mech: "batch_next_obs")})()
    env.generator = np.random.default_rng(seed + 100)
    out_ref = env.get_dynamics_predict({"batch_next_obs": {"mean": mean, "logvar": logvar}}, "transition")
    gen = np.random.default_rng(seed + 100)  # same stream: member draw, then the noise
    idx = mlp.get_random_index(B, gen)
    noise = gen.normal(size=(B, D)).astype(np.float32)
    assert rel_err(oc.dynamics_predict(mean, logvar, idx, noise), out_ref) < 1e-6
    assert rel_err(oc.mopo_penalty(mean), env.get_penalty(mean)) < 1e-6
