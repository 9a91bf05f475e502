"""B200-native batched simulator for the multi-UAV / multi-target tracking environment of
tjuDavidWang/MARL-UAVs-Targets-Tracking: the environment hot path only (reset / step /
observation / reward), behind the reference's `Environment` API.  CUDA (sm_100a) through the
C ABI in include/uavsim.h; no CPU path."""
from ._cabi import MODE_MEAN, MODE_PMI, MODE_SELF, UavSimError  # noqa: F401
from .environment import BatchedEnvironment, Environment, params_from_config  # noqa: F401
from .pmi import PMINetwork, fold_pmi  # noqa: F401
from .distributed import shard_envs, reduce_episode_stats, episode_summary, bind_host_to_gpu  # noqa: F401
from .config import default_config  # noqa: F401
from .replay import PrioritizedReplayBuffer  # noqa: F401

__version__ = "0.1.0"
