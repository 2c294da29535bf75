"""Drop-in for reference panda_gym/envs/robots/panda.py:10-140 (class Panda)."""
from typing import Optional

import numpy as np

from ... import spaces
from ..core import PyBulletRobot
from ...pybullet import PyBullet


class Panda(PyBulletRobot):
    def __init__(self, sim: PyBullet, block_gripper: bool = False, base_position: Optional[np.ndarray] = None, control_type: str = "ee") -> None:
        base_position = base_position if base_position is not None else np.zeros(3)
        self.block_gripper = block_gripper
        self.control_type = control_type
        n_action = 3 if self.control_type == "ee" else 7
        n_action += 0 if self.block_gripper else 1
        action_space = spaces.Box(-1.0, 1.0, shape=(n_action,), dtype=np.float32)
        super().__init__(sim, body_name="panda", file_name="franka_panda/panda.urdf", base_position=base_position, action_space=action_space,
                         joint_indices=np.array([0, 1, 2, 3, 4, 5, 6, 9, 10]), joint_forces=np.array([87.0, 87.0, 87.0, 87.0, 12.0, 120.0, 120.0, 170.0, 170.0]))
        self.fingers_indices = np.array([9, 10])
        self.neutral_joint_values = np.array([0.00, 0.41, 0.00, -1.85, 0.00, 2.26, 0.79, 0.00, 0.00])
        self.ee_link = 11
        self.sim.set_lateral_friction(self.body_name, self.fingers_indices[0], lateral_friction=1.0)
        self.sim.set_lateral_friction(self.body_name, self.fingers_indices[1], lateral_friction=1.0)
        self.sim.set_spinning_friction(self.body_name, self.fingers_indices[0], spinning_friction=0.001)
        self.sim.set_spinning_friction(self.body_name, self.fingers_indices[1], spinning_friction=0.001)

    def set_action(self, action: np.ndarray) -> None:
        """panda.py:52-70 -- clipping, IK / joint targets and the finger target are evaluated inside the step kernel."""
        action = np.asarray(action, dtype=np.float32).copy()
        if action.shape != self.action_space.shape:
            raise ValueError(f"action must have shape {self.action_space.shape}")
        self.sim._pending_action = action

    def get_obs(self) -> np.ndarray:
        """panda.py:109-119"""
        o = (self.sim._last_obs[0] if self.sim._last_obs is not None else self.sim._reset_obs)["observation"]
        return np.asarray(o[: 6 if self.block_gripper else 7], dtype=np.float64)

    def reset(self) -> None:
        pass  # the neutral pose is written by the reset kernel together with the task placement (core.py:240-250)

    def set_joint_neutral(self) -> None:
        self.set_joint_angles(self.neutral_joint_values)

    def get_fingers_width(self) -> float:
        return self.sim.get_joint_angle(self.body_name, 9) + self.sim.get_joint_angle(self.body_name, 10)

    def get_ee_position(self) -> np.ndarray:
        return self.get_link_position(self.ee_link)

    def get_ee_velocity(self) -> np.ndarray:
        return self.get_link_velocity(self.ee_link)
