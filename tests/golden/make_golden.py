"""Generate golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/cmrl through oracle/refshim.py (stub modules for the absent
hydra/omegaconf/stable_baselines3/gym imports; no reference code is copied) and writes
small .npz files next to this script.  The GPU box has no /root/reference, so the tests
there compare against these files.

Everything is seeded; re-running reproduces the files bit-for-bit on the same torch/numpy.
"""
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

ref = refshim.load_reference()
SILU = {"_target_": "torch.nn.SiLU"}
torch.set_num_threads(1)  # deterministic CPU reductions


def sd_np(module, prefix="p."):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def make_inputs(rng, E, B, O, A, ensemble=True):
    if ensemble:
        obs = rng.normal(size=(E, B, O)).astype(np.float32)
        act = rng.uniform(-1, 1, size=(E, B, A)).astype(np.float32)
    else:
        obs = np.repeat(rng.normal(size=(1, B, O)).astype(np.float32), E, 0)
        act = np.repeat(rng.uniform(-1, 1, size=(1, B, A)).astype(np.float32), E, 0)
    return obs, act


def forward_fixture(name, model, rng, B, target_dim, ensemble=True, extra=None):
    E, O, A = model.ensemble_num, model.obs_size, model.action_size
    obs, act = make_inputs(rng, E, B, O, A, ensemble)
    target = (obs[..., :target_dim] + 0.1 * rng.normal(size=(E, B, target_dim))).astype(np.float32)
    with torch.no_grad():
        mean, logvar = model.forward(torch.from_numpy(obs.copy()), torch.from_numpy(act.copy()))
        model_in = {"batch_obs": torch.from_numpy(obs.copy()), "batch_action": torch.from_numpy(act.copy())}
        nll = model.get_nll_loss(model_in, torch.from_numpy(target))
        mse = model.get_mse_loss(model_in, torch.from_numpy(target))
    out = sd_np(model)
    out.update(obs=obs, act=act, target=target, mean=mean.numpy(), logvar=logvar.numpy(), nll=nll.numpy(), mse=mse.numpy())
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: v.shape for k, v in out.items() if not k.startswith("p.")})


def train_fixture(name, model, rng, B, target_dim, steps=3, extra=None):
    """`steps` iterations of base_dynamics.py:181-186 on fresh bootstrap-shaped batches."""
    E, O, A = model.ensemble_num, model.obs_size, model.action_size
    optim = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5, eps=1e-8)
    out = sd_np(model, "p0.")
    for s in range(steps):
        obs, act = make_inputs(rng, E, B, O, A, True)
        target = (obs[..., :target_dim] + 0.1 * rng.normal(size=(E, B, target_dim))).astype(np.float32)
        model_in = {"batch_obs": torch.from_numpy(obs.copy()), "batch_action": torch.from_numpy(act.copy())}
        loss = model.get_nll_loss(model_in, torch.from_numpy(target))
        optim.zero_grad()
        loss.mean().backward()
        if s == 0:
            for k, p in model.named_parameters():
                if p.grad is not None:
                    out["g0." + k] = p.grad.numpy().copy()
        optim.step()
        out["obs%d" % s] = obs
        out["act%d" % s] = act
        out["target%d" % s] = target
        out["loss%d" % s] = loss.detach().numpy().copy()
    out.update(sd_np(model, "p%d." % steps))
    out["steps"] = np.int64(steps)
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "train steps", steps)


def cmi_fixtures():
    """CMI-test network (causal_discovery/CMI_test.py:15-130): P = O+A+1 leave-one-out sub-nets, outputs [P,E,B,O]."""
    torch.manual_seed(11)
    np.random.seed(11)
    rng = np.random.default_rng(11)
    O, A, E, B = 4, 2, 5, 30
    for name, kw in (("cmi_relu", {}), ("cmi_silu", {"activation_fn_cfg": SILU}), ("cmi_noresidual", {"activation_fn_cfg": SILU, "residual": False})):
        m = ref.TransitionConditionalMutualInformationTest(O, A, ensemble_num=E, elite_num=3, hid_size=24, **kw)
        obs, act = make_inputs(rng, E, B, O, A, True)
        target = (obs + 0.1 * rng.normal(size=(E, B, O))).astype(np.float32)
        with torch.no_grad():
            mean, logvar = m.forward(torch.from_numpy(obs.copy()), torch.from_numpy(act.copy()))
            nll = m.get_nll_loss({"batch_obs": torch.from_numpy(obs.copy()), "batch_action": torch.from_numpy(act.copy())},
                                 torch.from_numpy(target))
            # 4-D observations (one per sub-net): forward accepts [P,E,B,O] (CMI_test.py:103-105)
            obs4 = rng.normal(size=(O + A + 1, E, B, O)).astype(np.float32)
            mean4, logvar4 = m.forward(torch.from_numpy(obs4.copy()), torch.from_numpy(act.copy()))
        out = sd_np(m)
        out.update(obs=obs, act=act, target=target, mean=mean.numpy(), logvar=logvar.numpy(), nll=nll.numpy(),
                   obs4=obs4, mean4=mean4.numpy(), logvar4=logvar4.numpy(), input_mask=m.input_mask.numpy())
        np.savez_compressed(os.path.join(HERE, name + "_forward.npz"), **out)
        print(name, mean.shape, nll.shape)
    m = ref.TransitionConditionalMutualInformationTest(O, A, ensemble_num=E, elite_num=3, hid_size=24, activation_fn_cfg=SILU)
    train_fixture("cmi_train.npz", m, rng, B=32, target_dim=O, steps=3)
    m = ref.TransitionConditionalMutualInformationTest(O, A, ensemble_num=E, elite_num=3, hid_size=24, activation_fn_cfg=SILU,
                                                       learn_logvar_bounds=True)
    train_fixture("cmi_learnbounds_train.npz", m, rng, B=32, target_dim=O, steps=3)


def euler_fixtures():
    """ForwardEulerTransition (transition/multi_step/forward_euler.py:8-41): k-fold re-application of a one-step model.
    Forward with input gradients (k = 3), three Adam steps through the wrapper (k = 2) and a VecFakeEnv trajectory."""
    torch.manual_seed(21)
    np.random.seed(21)
    rng = np.random.default_rng(21)
    O, A, E, B, H = 5, 1, 7, 28, 32
    one = ref.PlainTransition(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    fe = ref.ForwardEulerTransition(one, repeat_times=3)
    obs, act = make_inputs(rng, E, B, O, A, True)
    target = (obs + 0.3 * rng.normal(size=(E, B, O))).astype(np.float32)
    t_obs = torch.from_numpy(obs.copy()).requires_grad_(True)
    t_act = torch.from_numpy(act.copy()).requires_grad_(True)
    mean, logvar = fe.forward(t_obs, t_act)
    nll = fe.get_nll_loss({"batch_obs": t_obs, "batch_action": t_act}, torch.from_numpy(target))
    nll.mean().backward()
    out = sd_np(one)
    out.update(obs=obs, act=act, target=target, mean=mean.detach().numpy(), logvar=logvar.detach().numpy(),
               nll=nll.detach().numpy(), g_obs=t_obs.grad.numpy().copy(), g_act=t_act.grad.numpy().copy(),
               repeat_times=np.int64(3))
    for k, p in one.named_parameters():
        if p.grad is not None:
            out["g." + k] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "euler_forward.npz"), **out)
    print("euler_forward", mean.shape)

    one = ref.PlainTransition(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    fe = ref.ForwardEulerTransition(one, repeat_times=2)
    train_fixture("euler_train.npz", fe, rng, B=24, target_dim=O, steps=3, extra={"repeat_times": np.int64(2)})

    # VecFakeEnv trajectory with the wrapper as the transition (the wrapper keeps its OWN elite set: its
    # set_elite_members forwards to the one-step model, get_random_index reads the wrapper's, forward_euler.py:29-30)
    torch.manual_seed(22)
    np.random.seed(22)
    N, T = 40, 3
    one = ref.PlainTransition(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    rw = ref.PlainRewardMech(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    with torch.no_grad():
        for m in (one, rw):
            m.mean_and_logvar.bias.add_(torch.from_numpy(rng.normal(size=tuple(m.mean_and_logvar.bias.shape)).astype(np.float32)) * 0.3)
    fe = ref.ForwardEulerTransition(one, repeat_times=2)
    dyn = ref.PlainEnsembleDynamics(fe, learned_reward=True, reward_mech=rw)
    init_rng = np.random.default_rng(78)
    env = ref.VecFakeEnv(N, None, None)
    env.set_up(dyn, reward_fn=None, termination_fn=lambda n, o, a: np.abs(n[:, 0:1]) > 1.5,
               get_init_obs_fn=lambda n: init_rng.normal(size=(n, O)).astype(np.float32))
    env.seed(13)
    out = sd_np(one, "tr.")
    out.update(sd_np(rw, "rw."))
    out["fe_elite"] = np.asarray(fe.elite_members, dtype=np.int64)
    out["tr_elite"] = np.asarray(one.elite_members, dtype=np.int64)
    out["rw_elite"] = np.asarray(rw.elite_members, dtype=np.int64)
    if reset_by_buffer:
        out["pool_obs"] = pool
        out["pool_pos"] = np.int64(33)
        np.random.seed(123)  # the tests seed the global generator the same way right before reset()
    out["reset_obs"] = env.reset()
    actions = rng.uniform(-1, 1, size=(T, N, A)).astype(np.float32)
    out["actions"] = actions
    for t in range(T):
        o, r, d, infos = env.step(actions[t])
        out["obs%d" % t], out["rew%d" % t], out["done%d" % t] = o, r, d
        term = np.zeros((N, O), np.float32)
        for i, info in enumerate(infos):
            if d[i]:
                term[i] = info["terminal_observation"]
        out["termobs%d" % t] = term
    out["T"] = np.int64(T)
    out["repeat_times"] = np.int64(2)
    np.savez_compressed(os.path.join(HERE, "fake_env_euler.npz"), **out)
    print("fake_env_euler done counts", [int(out["done%d" % t].sum()) for t in range(T)])


def synth_dataset(rng, N, O, A):
    """SURVEY 8(d) synthetic transitions."""
    obs = rng.normal(size=(N, O)).astype(np.float32)
    act = rng.uniform(-1, 1, size=(N, A)).astype(np.float32)
    M = rng.normal(size=(O, O)).astype(np.float32) / np.sqrt(O)
    K = rng.normal(size=(A, O)).astype(np.float32)
    nxt = (obs + 0.1 * np.tanh(obs @ M + act @ K) + 0.01 * rng.normal(size=(N, O))).astype(np.float32)
    rew = (np.cos(obs[:, 0]) - 0.1 * (act**2).sum(-1)).astype(np.float32)
    done = np.zeros(N, dtype=bool)
    return obs, act, nxt, rew, done


def learn_fixture(name):
    torch.manual_seed(3)
    np.random.seed(3)
    rng = np.random.default_rng(3)
    O, A, E, H, N = 5, 1, 7, 32, 640
    tr = ref.PlainTransition(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    rw = ref.PlainRewardMech(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    dyn = ref.PlainEnsembleDynamics(tr, learned_reward=True, reward_mech=rw, optim_lr=1e-3)
    out = sd_np(tr, "tr0.")
    out.update(sd_np(rw, "rw0."))
    obs, act, nxt, rew, done = synth_dataset(rng, N, O, A)
    buf = refshim.DuckReplayBuffer(obs, act, nxt, rew, done)
    seeds = iter(range(1000, 2000))
    real_default_rng = np.random.default_rng

    def seeded_default_rng(seed=None):
        return real_default_rng(next(seeds) if seed is None else seed)

    with mock.patch("numpy.random.default_rng", seeded_default_rng):
        dyn.learn(buf, validation_ratio=0.2, batch_size=64, longest_epoch=4, improvement_threshold=0.01, patience=3)
    out.update(sd_np(tr, "tr1."))
    out.update(sd_np(rw, "rw1."))
    out.update(
        obs=obs,
        act=act,
        next_obs=nxt,
        rew=rew,
        done=done,
        tr_elite=np.asarray(tr.elite_members, dtype=np.int64),
        rw_elite=np.asarray(rw.elite_members, dtype=np.int64),
        total_epoch_tr=np.int64(dyn.total_epoch["transition"]),
        total_epoch_rw=np.int64(dyn.total_epoch["reward_mech"]),
    )
    # per-member validation MSE of the final models (what elite selection sorts)
    _, val_it = dyn.dataset_split(buf, 0.2, 64)
    out["tr_val_mse"] = dyn.evaluate(val_it, "transition").mean(dim=(1, 2)).numpy()
    out["rw_val_mse"] = dyn.evaluate(val_it, "reward_mech").mean(dim=(1, 2)).numpy()
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "elites", tr.elite_members, rw.elite_members, "epochs", dyn.total_epoch)


def fake_env_fixture(name, learned_termination, penalty_coeff, deterministic, max_episode_steps, transition_kind="plain",
                     reset_by_buffer=False):
    torch.manual_seed(5)
    np.random.seed(5)
    rng = np.random.default_rng(5)
    O, A, E, H, N, T = 5, 1, 7, 32, 48, 4
    if transition_kind == "plain":
        tr = ref.PlainTransition(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    else:
        tr = ref.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
        mask = (rng.uniform(size=(O, O + A)) < 0.5).astype(np.float32)
        mask[np.arange(O), np.arange(O)] = 1.0
        tr.set_input_mask(torch.from_numpy(mask))
    rw = ref.PlainRewardMech(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    tm = ref.PlainTerminationMech(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    # make the heads non-trivial (init biases are 0 -> outputs tiny)
    with torch.no_grad():
        for m in (tr, rw, tm):
            m.mean_and_logvar.bias.add_(torch.from_numpy(rng.normal(size=tuple(m.mean_and_logvar.bias.shape)).astype(np.float32)) * 0.3)
    dyn = ref.PlainEnsembleDynamics(
        tr, learned_reward=True, reward_mech=rw, learned_termination=learned_termination, termination_mech=tm if learned_termination else None
    )
    init_rng = np.random.default_rng(77)

    def get_init_obs_fn(n):
        return init_rng.normal(size=(n, O)).astype(np.float32)

    def termination_fn(next_obs, obs, act):
        return np.abs(next_obs[:, 0:1]) > 1.5

    pool = None
    if reset_by_buffer:
        # resets draw rows of a (partly filled) real replay buffer with the GLOBAL numpy RNG (fake_env.py:152-155,175-178)
        pool = init_rng.normal(size=(40, O)).astype(np.float32)
        buf = refshim.DuckReplayBuffer(pool, np.zeros((40, A), np.float32), pool, np.zeros(40, np.float32), np.zeros(40, bool))
        buf.full, buf.pos = False, 33
    env = ref.VecFakeEnv(N, None, None)
    env.set_up(
        dyn,
        reward_fn=None,
        termination_fn=None if learned_termination else termination_fn,
        get_init_obs_fn=None if reset_by_buffer else get_init_obs_fn,
        real_replay_buffer=buf if reset_by_buffer else None,
        penalty_coeff=penalty_coeff,
        deterministic=deterministic,
        max_episode_steps=max_episode_steps,
    )
    env.seed(11)
    out = sd_np(tr, "tr.")
    out.update(sd_np(rw, "rw."))
    if learned_termination:
        out.update(sd_np(tm, "tm."))
    if transition_kind != "plain":
        out["input_mask"] = mask
    out["tr_elite"] = np.asarray(tr.elite_members, dtype=np.int64)
    out["rw_elite"] = np.asarray(rw.elite_members, dtype=np.int64)
    out["tm_elite"] = np.asarray(tm.elite_members, dtype=np.int64)
    if reset_by_buffer:
        out["pool_obs"] = pool
        out["pool_pos"] = np.int64(33)
        np.random.seed(123)  # the tests seed the global generator the same way right before reset()
    out["reset_obs"] = env.reset()
    actions = rng.uniform(-1, 1, size=(T, N, A)).astype(np.float32)
    out["actions"] = actions
    for t in range(T):
        obs, rew, done, infos = env.step(actions[t])
        out["obs%d" % t] = obs
        out["rew%d" % t] = rew
        out["done%d" % t] = done
        out["trunc%d" % t] = np.asarray([i["TimeLimit.truncated"] for i in infos])
        term = np.zeros((N, O), np.float32)
        for i, info in enumerate(infos):
            assert ("terminal_observation" in info) == bool(done[i])
            if done[i]:
                term[i] = info["terminal_observation"]
        out["termobs%d" % t] = term
    out["T"] = np.int64(T)
    out["cfg"] = np.asarray([int(learned_termination), penalty_coeff, int(deterministic), max_episode_steps], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "done counts", [int(out["done%d" % t].sum()) for t in range(T)])


def main():
    # ---- forward fixtures -------------------------------------------------------------
    torch.manual_seed(0)
    np.random.seed(0)
    rng = np.random.default_rng(0)
    # C1 full size: PlainTransition E=7/5, 4x200, SiLU, O=5, A=1
    m = ref.PlainTransition(5, 1, ensemble_num=7, elite_num=5, hid_size=200, activation_fn_cfg=SILU)
    forward_fixture("plain_c1_forward.npz", m, rng, B=96, target_dim=5, ensemble=False)
    # reference unit-test shape (test_basic_ensemble.py:10-46): O=11, A=3, ReLU default, H=48
    m = ref.PlainTransition(11, 3, ensemble_num=7, hid_size=48)
    forward_fixture("plain_relu_forward.npz", m, rng, B=37, target_dim=11)
    m = ref.PlainTransition(5, 1, ensemble_num=7, hid_size=40, activation_fn_cfg=SILU, residual=False)
    forward_fixture("plain_noresidual_forward.npz", m, rng, B=33, target_dim=5)
    m = ref.PlainRewardMech(5, 1, ensemble_num=7, hid_size=40, activation_fn_cfg=SILU)
    forward_fixture("reward_forward.npz", m, rng, B=50, target_dim=1)
    m = ref.PlainTerminationMech(5, 1, ensemble_num=7, hid_size=40, activation_fn_cfg=SILU)
    forward_fixture("termination_forward.npz", m, rng, B=50, target_dim=1)
    # ExternalMask: 2-D [P,in], 3-D per-member [P,E,in], 3-D per-sample [P,B,in], 4-D masks
    O, A, E, B = 5, 1, 7, 40
    for tag, shape in (("2d", (O, O + A)), ("3de", (O, E, O + A)), ("3db", (O, B, O + A)), ("4d", (O, E, B, O + A))):
        m = ref.ExternalMaskTransition(O, A, ensemble_num=E, hid_size=24, activation_fn_cfg=SILU)
        mask = (rng.uniform(size=shape) < 0.5).astype(np.float32)
        m.set_input_mask(torch.from_numpy(mask))
        forward_fixture("extmask_%s_forward.npz" % tag, m, rng, B=B, target_dim=O, extra={"input_mask": mask})
    # C5-shaped (scaled down): O=P=9, A=3, E=4, H=40, per-member mask
    m = ref.ExternalMaskTransition(9, 3, ensemble_num=4, elite_num=3, hid_size=40, activation_fn_cfg=SILU)
    mask = (rng.uniform(size=(9, 4, 12)) < 0.3).astype(np.float32)
    m.set_input_mask(torch.from_numpy(mask))
    forward_fixture("extmask_wide_forward.npz", m, rng, B=24, target_dim=9, extra={"input_mask": mask})

    # ---- training-step fixtures ----------------------------------------------------------
    torch.manual_seed(1)
    rng = np.random.default_rng(1)
    m = ref.PlainTransition(5, 1, ensemble_num=7, hid_size=40, activation_fn_cfg=SILU)
    train_fixture("plain_train.npz", m, rng, B=32, target_dim=5)
    m = ref.PlainTransition(5, 1, ensemble_num=7, hid_size=24)  # ReLU
    train_fixture("plain_relu_train.npz", m, rng, B=20, target_dim=5)
    m = ref.PlainRewardMech(5, 1, ensemble_num=7, hid_size=40, activation_fn_cfg=SILU)
    train_fixture("reward_train.npz", m, rng, B=32, target_dim=1)
    m = ref.ExternalMaskTransition(5, 1, ensemble_num=7, hid_size=24, activation_fn_cfg=SILU)
    mask = (rng.uniform(size=(5, 6)) < 0.5).astype(np.float32)
    mask[np.arange(5), np.arange(5)] = 1.0
    m.set_input_mask(torch.from_numpy(mask))
    train_fixture("extmask_train.npz", m, rng, B=24, target_dim=5, extra={"input_mask": mask})
    m = ref.PlainTransition(5, 1, ensemble_num=7, hid_size=24, activation_fn_cfg=SILU, learn_logvar_bounds=True)
    train_fixture("plain_learnbounds_train.npz", m, rng, B=16, target_dim=5)

    # ---- learn() (early stopping + elite selection) ------------------------------------------
    learn_fixture("learn_small.npz")

    # ---- VecFakeEnv.step ------------------------------------------------------------------
    fake_env_fixture("fake_env_termfn.npz", False, 0.0, False, 1000)
    fake_env_fixture("fake_env_truncate.npz", False, 0.0, False, 3)
    fake_env_fixture("fake_env_learned_term.npz", True, 0.0, False, 1000)
    fake_env_fixture("fake_env_penalty.npz", False, 0.7, False, 1000)
    fake_env_fixture("fake_env_deterministic.npz", False, 0.0, True, 1000)
    fake_env_fixture("fake_env_extmask.npz", False, 0.0, False, 1000, transition_kind="extmask")


def constraint_fixtures():
    """ConstraintBasedDynamics.learn (constraint_based_dynamics.py:83-183): the oracle causal mask is set on the transition
    before training, bootstrap PERMUTATIONS per member, a ragged last batch, checkpoint written at the end."""
    import tempfile

    torch.manual_seed(7)
    np.random.seed(7)
    rng = np.random.default_rng(7)
    O, A, E, H, N = 5, 1, 7, 24, 650
    tr = ref.ExternalMaskTransition(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    rw = ref.PlainRewardMech(O, A, ensemble_num=E, elite_num=5, hid_size=H, activation_fn_cfg=SILU)
    dyn = ref.ConstraintBasedDynamics(tr, learned_reward=True, reward_mech=rw, optim_lr=1e-3)
    assert dyn.transition_oracle_mask is None and tuple(dyn.transition_history_mask.shape) == (O, O + A)
    assert not hasattr(dyn, "reward_mech_oracle_mask")
    mask = (rng.uniform(size=(O, O + A)) < 0.5).astype(np.float32)
    mask[np.arange(O), np.arange(O)] = 1.0
    dyn.set_oracle_mask("transition", mask)
    out = sd_np(tr, "tr0.")
    out.update(sd_np(rw, "rw0."))
    obs, act, nxt, rew, done = synth_dataset(rng, N, O, A)
    buf = refshim.DuckReplayBuffer(obs, act, nxt, rew, done)
    seeds = iter(range(3000, 4000))
    real_default_rng = np.random.default_rng

    def seeded_default_rng(seed=None):
        return real_default_rng(next(seeds) if seed is None else seed)

    with tempfile.TemporaryDirectory() as d, mock.patch("numpy.random.default_rng", seeded_default_rng):
        dyn.learn(buf, validation_ratio=0.2, batch_size=64, bootstrap_permutes=True, longest_epoch=3,
                  improvement_threshold=0.01, patience=2, work_dir=d)
        files = sorted(os.listdir(d))
        saved = torch.load(os.path.join(d, "external_mask_transition.pth"), weights_only=False)
    assert torch.equal(tr.input_mask, torch.from_numpy(mask))
    out.update(sd_np(tr, "tr1."))
    out.update(sd_np(rw, "rw1."))
    out.update(obs=obs, act=act, next_obs=nxt, rew=rew, done=done, oracle_mask=mask,
               saved_input_mask=saved["input_mask"].numpy(), saved_keys=np.asarray(sorted(saved.keys())),
               files=np.asarray(files),
               tr_elite=np.asarray(tr.elite_members, dtype=np.int64), rw_elite=np.asarray(rw.elite_members, dtype=np.int64),
               total_epoch_tr=np.int64(dyn.total_epoch["transition"]), total_epoch_rw=np.int64(dyn.total_epoch["reward_mech"]))
    _, val_it = dyn.dataset_split(buf, 0.2, 64)
    out["tr_val_mse"] = dyn.evaluate(val_it, "transition").mean(dim=(1, 2)).numpy()
    out["rw_val_mse"] = dyn.evaluate(val_it, "reward_mech").mean(dim=(1, 2)).numpy()
    np.savez_compressed(os.path.join(HERE, "constraint_learn.npz"), **out)
    print("constraint_learn.npz", "elites", tr.elite_members, rw.elite_members, "epochs", dyn.total_epoch, files)
    # VecFakeEnv whose resets come from the real replay buffer (global numpy RNG), many resets: short episodes + termination_fn
    fake_env_fixture("fake_env_bufreset.npz", False, 0.0, False, 2, reset_by_buffer=True)


if __name__ == "__main__":
    if "--only-constraint" in sys.argv:
        constraint_fixtures()
    elif "--only-cmi" in sys.argv:
        cmi_fixtures()
    elif "--only-euler" in sys.argv:
        euler_fixtures()
    else:
        main()
        cmi_fixtures()
        euler_fixtures()
        constraint_fixtures()
