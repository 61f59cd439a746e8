"""cmrl_b200 -- B200-native (sm_100a) implementation of the cmrl ensemble-dynamics hot path.

Drop-in for the reference classes on that path (polixir/causal-mbrl, `cmrl.models.*`): same constructor
signatures, attributes, state_dict keys and checkpoint files; the arithmetic runs in libcmrl_b200.so
(include/cmrl_b200.h).  Import as `cmrl_b200` (the on-disk directory is `causal-mbrl_b200/`).
"""
from . import _lib, ops  # noqa: F401
from .cmi_test import TransitionConditionalMutualInformationTest  # noqa: F401
from .dynamics import BaseDynamics, ConstraintBasedDynamics, PlainEnsembleDynamics, net_shard_range  # noqa: F401
from .fake_env import VecFakeEnv  # noqa: F401
from .layers import EnsembleLinearLayer, ParallelEnsembleLinearLayer, truncated_normal_init  # noqa: F401
from .mlp import EnsembleMLP, ExternalMaskEnsembleMLP  # noqa: F401
from .optim import FusedAdamL2  # noqa: F401
from .reward_mech import BaseRewardMech, PlainRewardMech  # noqa: F401
from .termination_mech import BaseTerminationMech, PlainTerminationMech  # noqa: F401
from .transition import BaseTransition, ExternalMaskTransition, ForwardEulerTransition, PlainTransition  # noqa: F401
from .transition_iterator import BootstrapIterator, TransitionIterator  # noqa: F401
from .types import InteractionBatch  # noqa: F401
from .util import gaussian_nll, to_tensor, truncated_normal_  # noqa: F401

__version__ = "0.1.0"
