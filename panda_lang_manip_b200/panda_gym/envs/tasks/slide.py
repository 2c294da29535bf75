"""Drop-in for reference panda_gym/envs/tasks/slide.py (class Slide): scene, sampling, observation, success and reward of the task are
implemented by the CUDA kernels (csrc/panda_env.cuh, panda_kernels.cuh env_reset); this class keeps the plug-in interface."""
from ._base import BuiltinTask


class Slide(BuiltinTask):
    name = "slide"
    default_threshold = 0.05

    def __init__(self, sim, reward_type="sparse", distance_threshold=None, **kwargs) -> None:

        super().__init__(sim, reward_type=reward_type, distance_threshold=distance_threshold, **kwargs)
