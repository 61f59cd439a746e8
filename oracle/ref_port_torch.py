"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's CPU execution of the hot path, restated with the
same torch-CPU / numpy library calls the reference makes (so its timing is the reference's timing).

The reference is pure Python over torch and numpy and cannot travel to the GPU box (`/root/reference` is not
mounted there), so `bench.py`'s `cpu_baseline` leg and `bench.py --impl reference` time THIS port on the
box's host cores (`cpu_baseline.kind = "port"`).  Every function cites the reference lines it follows; the op
sequence is kept one-to-one, including the work the B200 path avoids (E-fold input repeat, P-fold repeat + mask
multiply, all-member evaluation, [E,N,D] device->numpy conversion, per-env Python loop).

Pinned by tests/test_oracle_golden.py::test_torch_port_* against the golden outputs of the unmodified
reference.  Never imported by the product package.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _act(kind):
    return {"relu": torch.relu, "silu": F.silu, "identity": (lambda x: x)}[kind]


def mlp_stack(params, x, act, num_layers=4):
    """hidden_layers + mean_and_logvar: x.matmul(W) + b then the activation module
    (layers.py:60-65 / :106-111, plain_transition.py:75-91,108-109)."""
    h = x
    for i in range(num_layers):
        h = act(h.matmul(params["hidden_layers.%d.0.weight" % i]) + params["hidden_layers.%d.0.bias" % i])
    return h.matmul(params["mean_and_logvar.weight"]) + params["mean_and_logvar.bias"]


def plain_forward(params, obs, action, act, residual=True, head=None):
    """PlainTransition.forward (plain_transition.py:100-122) / PlainRewardMech.forward
    (plain_reward_mech.py:75-94) for head = obs_size / 1."""
    raw = mlp_stack(params, torch.concat([obs, action], dim=-1), act)
    d = obs.shape[-1] if head is None else head
    mean = raw[..., :d]
    logvar = raw[..., d:]
    logvar = params["max_logvar"] - F.softplus(params["max_logvar"] - logvar)
    logvar = params["min_logvar"] + F.softplus(logvar - params["min_logvar"])
    if residual:
        mean += obs
    return mean, logvar


def external_mask_forward(params, mask, obs, action, act, residual=True):
    """ExternalMaskTransition.forward (external_mask_transition.py:141-169) with a 2-D / per-member mask."""
    P = obs.shape[-1]
    x = torch.concat([obs, action], dim=-1).repeat((P, 1, 1, 1))
    x *= mask[:, None, None, :] if mask.ndim == 2 else mask[:, :, None, :]
    raw = mlp_stack(params, x, act)
    mean = raw[..., :1]
    logvar = raw[..., 1:]
    logvar = params["max_logvar"] - F.softplus(params["max_logvar"] - logvar)
    logvar = params["min_logvar"] + F.softplus(logvar - params["min_logvar"])
    mean = torch.transpose(mean, 0, -1)[0]
    logvar = torch.transpose(logvar, 0, -1)[0]
    if residual:
        mean += obs
    return mean, logvar


class RefPortEnv:
    """VecFakeEnv.step_wait + BaseDynamics.query + get_dynamics_predict on the CPU
    (fake_env.py:97-142,195-215; base_dynamics.py:90-100,190-204)."""

    def __init__(self, mechs, num_envs, init_obs, max_episode_steps=1000, termination_fn=None, act="silu", seed=0):
        # mechs: list of (kind, params(dict of torch CPU tensors), elite list, mask-or-None); kind in
        # {"plain", "extmask", "reward", "termination"}; order = learn_mech order (base_dynamics.py:57-79)
        self.mechs = mechs
        self.N = num_envs
        self.obs = np.ascontiguousarray(init_obs, dtype=np.float32)
        self.env_len = np.zeros(num_envs, dtype=int)
        self.max_episode_steps = max_episode_steps
        self.termination_fn = termination_fn
        self.act = _act(act)
        self.generator = np.random.default_rng(seed)
        self.reset_rng = np.random.default_rng(seed + 1)
        self.E = mechs[0][1]["mean_and_logvar.weight"].shape[-3]

    def _predict(self, kind, params, elite, mask, obs_t, act_t):
        with torch.no_grad():
            if kind == "plain":
                mean, logvar = plain_forward(params, obs_t, act_t, self.act)
            elif kind == "extmask":
                mean, logvar = external_mask_forward(params, mask, obs_t, act_t, self.act)
            else:
                mean, logvar = plain_forward(params, obs_t, act_t, self.act, residual=False, head=1)
        ensemble_mean, ensemble_logvar = mean.cpu().numpy(), logvar.cpu().numpy()  # base_dynamics.py:199-200
        batch_size = ensemble_mean.shape[1]
        random_index = self.generator.choice(elite, size=batch_size)  # mlp.py:40-44
        ensemble_std = np.sqrt(np.exp(ensemble_logvar))
        ar = np.arange(batch_size)
        return ensemble_mean[random_index, ar] + ensemble_std[random_index, ar] * self.generator.normal(
            size=ensemble_mean.shape[1:]
        ).astype(np.float32)

    def step(self, actions):
        N = self.N
        obs_t = torch.from_numpy(self.obs).to(torch.float32).repeat([self.E, 1, 1])  # get_3d_tensor, is_ensemble=False
        act_t = torch.from_numpy(actions).to(torch.float32).repeat([self.E, 1, 1])
        preds = {}
        for kind, params, elite, mask in self.mechs:
            preds[kind] = self._predict(kind, params, elite, mask, obs_t, act_t)
        next_obs = preds.get("plain", preds.get("extmask"))
        reward = preds["reward"].reshape(N)
        if "termination" in preds:
            terminal = preds["termination"].reshape(N)
        else:
            terminal = np.asarray(self.termination_fn(next_obs, self.obs, actions)).reshape(N)
        self.obs = next_obs.copy()
        infos = [{} for _ in range(N)]
        self.env_len += 1
        truncate = self.env_len >= self.max_episode_steps
        done = np.logical_or(terminal, truncate)
        for idx in range(N):  # fake_env.py:131-135
            infos[idx]["TimeLimit.truncated"] = truncate[idx]
            if done[idx]:
                infos[idx]["terminal_observation"] = next_obs[idx].copy()
                self.env_len[idx] = 0
                self.obs[idx] = self.reset_rng.normal(size=(1, self.obs.shape[1])).astype(np.float32)
        return self.obs.copy(), reward.copy(), done.copy(), infos


def train_batch(params, optim, obs, action, target, act, kind="plain", mask=None):
    """One iteration of BaseDynamics.train (base_dynamics.py:181-186): NLL forward, .mean().backward(), Adam."""
    if kind == "extmask":
        mean, logvar = external_mask_forward(params, mask, obs, action, act)
    elif kind == "plain":
        mean, logvar = plain_forward(params, obs, action, act)
    else:
        mean, logvar = plain_forward(params, obs, action, act, residual=False, head=1)
    l2 = F.mse_loss(mean, target, reduction="none")
    loss = l2 * (-logvar).exp() + logvar
    loss = loss + 0.01 * (params["max_logvar"].sum() - params["min_logvar"].sum())
    optim.zero_grad()
    loss.mean().backward()
    optim.step()
    return loss.detach()
