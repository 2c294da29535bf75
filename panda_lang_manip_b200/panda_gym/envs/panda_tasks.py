"""Drop-in for reference panda_gym/envs/panda_tasks.py:14-113: the six env classes (constructor signature of the fork:
``render: bool``; ``render_mode`` of the upstream docs is tolerated and ignored)."""
import numpy as np

from .core import RobotTaskEnv
from .robots.panda import Panda
from .tasks import Flip, PickAndPlace, Push, Reach, Slide, Stack
from ..pybullet import PyBullet


def _make(env, task_cls, task_name, block_gripper, render, reward_type, control_type, device, precision):
    sim = PyBullet(render=render)
    sim._bind(task_name, reward_type, control_type, device=device, precision=precision)
    robot = Panda(sim, block_gripper=block_gripper, base_position=np.array([-0.6, 0.0, 0.0]), control_type=control_type)
    task = task_cls(sim, get_ee_position=robot.get_ee_position, reward_type=reward_type) if task_name == "reach" else task_cls(sim, reward_type=reward_type)
    RobotTaskEnv.__init__(env, robot, task)


class PandaFlipEnv(RobotTaskEnv):
    def __init__(self, render: bool = False, reward_type: str = "sparse", control_type: str = "ee", render_mode=None, device: int = 0, precision: str = "f32") -> None:
        _make(self, Flip, "flip", False, render, reward_type, control_type, device, precision)


class PandaPickAndPlaceEnv(RobotTaskEnv):
    def __init__(self, render: bool = False, reward_type: str = "sparse", control_type: str = "ee", render_mode=None, device: int = 0, precision: str = "f32") -> None:
        _make(self, PickAndPlace, "pick_and_place", False, render, reward_type, control_type, device, precision)


class PandaPushEnv(RobotTaskEnv):
    def __init__(self, render: bool = False, reward_type: str = "sparse", control_type: str = "ee", render_mode=None, device: int = 0, precision: str = "f32") -> None:
        _make(self, Push, "push", True, render, reward_type, control_type, device, precision)


class PandaReachEnv(RobotTaskEnv):
    def __init__(self, render: bool = False, reward_type: str = "sparse", control_type: str = "ee", render_mode=None, device: int = 0, precision: str = "f32") -> None:
        _make(self, Reach, "reach", True, render, reward_type, control_type, device, precision)


class PandaSlideEnv(RobotTaskEnv):
    def __init__(self, render: bool = False, reward_type: str = "sparse", control_type: str = "ee", render_mode=None, device: int = 0, precision: str = "f32") -> None:
        _make(self, Slide, "slide", True, render, reward_type, control_type, device, precision)


class PandaStackEnv(RobotTaskEnv):
    def __init__(self, render: bool = False, reward_type: str = "sparse", control_type: str = "ee", render_mode=None, device: int = 0, precision: str = "f32") -> None:
        _make(self, Stack, "stack", False, render, reward_type, control_type, device, precision)
