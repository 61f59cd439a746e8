"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (`/root/reference/cmrl`).

The reference is pure Python; its hot path only fails to import because four
third-party packages are absent in this image.  This
module registers minimal stand-ins for those *imports* (no arithmetic lives in
them) and then imports the reference from where it lies.  Nothing is copied.

Used by `tests/golden/make_golden.py` (fixture generation, in the build
container only) and by tests that pin `oracle/cmrl_oracle.py` against the real
reference when `/root/reference` is mounted.  `/root/reference` does not exist
on the GPU box; `available()` is False there and callers must skip.

Never imported by the product package.
"""
import importlib
import os
import sys
import types

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The unmodified reference: the mounted tree in the build container, else the copy `__graft_entry__.build()` installed with
# pip into the git-ignored baseline/_ref/ (it travels to the GPU box, where bench.py's reference arm times it).
_INSTALLED = os.path.join(_REPO_ROOT, "baseline", "_ref")


def _default_root():
    env = os.environ.get("CMRL_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/cmrl"):
        return "/root/reference"
    return _INSTALLED


REFERENCE_ROOT = _default_root()


def mounted() -> bool:
    """True only where the reference SOURCE TREE is mounted (the build container): live-reference tests and the golden
    generator need it; the installed copy under baseline/_ref is for bench.py's timed reference arm."""
    return os.path.isdir("/root/reference/cmrl")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cmrl"))


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _instantiate(cfg, *args, **kwargs):
    # hydra.utils.instantiate as used at plain_transition.py:69-73: resolve
    # "_target_" and call it without arguments.
    target = cfg["_target_"]
    mod, _, attr = target.rpartition(".")
    return getattr(importlib.import_module(mod), attr)(*args, **kwargs)


class _VecEnv:
    # stand-in for stable_baselines3 VecEnv (fake_env.py:23-34 only stores the
    # three ctor args and relies on step = step_async + step_wait)
    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()


def install_stubs():
    class _Placeholder:
        pass

    if "omegaconf" not in sys.modules:
        _mod("omegaconf", DictConfig=dict, ListConfig=list, OmegaConf=_Placeholder)
    if "hydra" not in sys.modules:
        utils = _mod("hydra.utils", instantiate=_instantiate)
        _mod("hydra", utils=utils)
    if "stable_baselines3" not in sys.modules:
        _mod("stable_baselines3")
        _mod("stable_baselines3.common")
        _mod("stable_baselines3.common.logger", Logger=_Placeholder)
        _mod("stable_baselines3.common.buffers", ReplayBuffer=_Placeholder)
        _mod("stable_baselines3.common.vec_env", VecEnv=_VecEnv)
        _mod(
            "stable_baselines3.common.vec_env.base_vec_env",
            VecEnv=_VecEnv,
            VecEnvIndices=_Placeholder,
            VecEnvObs=_Placeholder,
            VecEnvStepReturn=_Placeholder,
        )
    if "gym" not in sys.modules:
        spaces = _mod("gym.spaces", Space=_Placeholder)
        core = _mod("gym.core", ActType=_Placeholder, ObsType=_Placeholder)
        _mod("gym", spaces=spaces, core=core, Wrapper=_Placeholder)


def load_reference():
    """Return a namespace with the reference's hot-path classes (unmodified code)."""
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    from cmrl.models.layers import EnsembleLinearLayer, ParallelEnsembleLinearLayer, truncated_normal_init
    from cmrl.models.util import gaussian_nll, truncated_normal_
    from cmrl.models.networks.mlp import EnsembleMLP
    from cmrl.models.transition.one_step.plain_transition import PlainTransition
    from cmrl.models.transition.one_step.external_mask_transition import ExternalMaskTransition
    from cmrl.models.transition.multi_step.forward_euler import ForwardEulerTransition
    from cmrl.models.reward_mech.plain_reward_mech import PlainRewardMech
    from cmrl.models.termination_mech.plain_termination_mech import PlainTerminationMech
    from cmrl.models.dynamics.base_dynamics import BaseDynamics
    from cmrl.models.dynamics.plain_dynamics import PlainEnsembleDynamics
    from cmrl.models.dynamics.constraint_based_dynamics import ConstraintBasedDynamics
    from cmrl.models.fake_env import VecFakeEnv
    from cmrl.util.transition_iterator import BootstrapIterator, TransitionIterator
    from cmrl.types import InteractionBatch
    from cmrl.models.causal_discovery.CMI_test import TransitionConditionalMutualInformationTest

    for k, v in list(locals().items()):
        if k not in ("ns",):
            setattr(ns, k, v)
    return ns


class DuckReplayBuffer:
    """Object with the SB3 ReplayBuffer fields the reference reads
    (base_dynamics.py:131-138, fake_env.py:153-155)."""

    def __init__(self, obs, act, next_obs, rew, done):
        n = obs.shape[0]
        self.buffer_size = n
        self.full = True
        self.pos = 0
        self.observations = obs[:, None, :]
        self.next_observations = next_obs[:, None, :]
        self.actions = act[:, None, :]
        self.rewards = rew.reshape(n, 1)
        self.dones = done.reshape(n, 1)
