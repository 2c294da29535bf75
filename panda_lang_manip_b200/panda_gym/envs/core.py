"""Drop-in for reference panda_gym/envs/core.py: PyBulletRobot (:11-158), Task (:161-196), RobotTaskEnv (:199-335)."""
from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .. import spaces
from ..pybullet import PyBullet

try:  # gymnasium is optional here
    import gymnasium as _gym
    _EnvBase = _gym.Env
except Exception:
    _EnvBase = object


class PyBulletRobot(ABC):
    """core.py:11-158.  The URDF is not loaded: the Panda tree is compiled into the CUDA kernels."""

    def __init__(self, sim: PyBullet, body_name: str, file_name: str, base_position: np.ndarray, action_space, joint_indices: np.ndarray, joint_forces: np.ndarray) -> None:
        self.sim = sim
        self.body_name = body_name
        with self.sim.no_rendering():
            self._load_robot(file_name, base_position)
            self.setup()
        self.action_space = action_space
        self.joint_indices = joint_indices
        self.joint_forces = joint_forces

    def _load_robot(self, file_name: str, base_position: np.ndarray) -> None:
        self.sim.loadURDF(body_name=self.body_name, fileName=file_name, basePosition=base_position, useFixedBase=True)

    def setup(self) -> None:
        pass

    @abstractmethod
    def set_action(self, action: np.ndarray) -> None:
        ...

    @abstractmethod
    def get_obs(self) -> np.ndarray:
        ...

    @abstractmethod
    def reset(self) -> None:
        ...

    def get_link_position(self, link: int) -> np.ndarray:
        return self.sim.get_link_position(self.body_name, link)

    def get_link_velocity(self, link: int) -> np.ndarray:
        return self.sim.get_link_velocity(self.body_name, link)

    def get_joint_angle(self, joint: int) -> float:
        return self.sim.get_joint_angle(self.body_name, joint)

    def get_joint_velocity(self, joint: int) -> float:
        return self.sim.get_joint_velocity(self.body_name, joint)

    def control_joints(self, target_angles: np.ndarray) -> None:
        self.sim.control_joints(body=self.body_name, joints=self.joint_indices, target_angles=target_angles, forces=self.joint_forces)

    def set_joint_angles(self, angles: np.ndarray) -> None:
        self.sim.set_joint_angles(self.body_name, joints=self.joint_indices, angles=angles)

    def inverse_kinematics(self, link: int, position: np.ndarray, orientation: np.ndarray) -> np.ndarray:
        return self.sim.inverse_kinematics(self.body_name, link=link, position=position, orientation=orientation)


class Task(ABC):
    """core.py:161-196."""

    def __init__(self, sim: PyBullet) -> None:
        self.sim = sim
        self.goal = None

    @abstractmethod
    def reset(self) -> None:
        ...

    @abstractmethod
    def get_obs(self) -> np.ndarray:
        ...

    @abstractmethod
    def get_achieved_goal(self) -> np.ndarray:
        ...

    def get_goal(self) -> np.ndarray:
        if self.goal is None:
            raise RuntimeError("No goal yet, call reset() first")
        return self.goal.copy()

    @abstractmethod
    def is_success(self, achieved_goal: np.ndarray, desired_goal: np.ndarray, info: Dict[str, Any] = {}) -> np.ndarray:
        ...

    @abstractmethod
    def compute_reward(self, achieved_goal: np.ndarray, desired_goal: np.ndarray, info: Dict[str, Any] = {}) -> np.ndarray:
        ...


class RobotTaskEnv(_EnvBase):
    """core.py:199-335: junction of a robot and a task.  step() is one fused kernel launch on a batch of one environment;
    use panda_lang_manip_b200.PandaVecEnv for batches."""

    metadata = {"render_modes": ["human", "rgb_array"]}

    def __init__(self, robot: PyBulletRobot, task: Task) -> None:
        assert robot.sim == task.sim, "The robot and the task must belong to the same simulation."
        self.sim = robot.sim
        self.robot = robot
        self.task = task
        observation, _ = self.reset()  # required for init; seed can be changed later
        observation_shape = observation["observation"].shape
        achieved_goal_shape = observation["achieved_goal"].shape
        self.observation_space = spaces.Dict(
            dict(
                observation=spaces.Box(-10.0, 10.0, shape=observation_shape, dtype=np.float32),
                desired_goal=spaces.Box(-10.0, 10.0, shape=achieved_goal_shape, dtype=np.float32),
                achieved_goal=spaces.Box(-10.0, 10.0, shape=achieved_goal_shape, dtype=np.float32),
            )
        )
        self.action_space = self.robot.action_space
        self.compute_reward = self.task.compute_reward
        self._saved_goal = dict()

    def _get_obs(self) -> Dict[str, np.ndarray]:
        o = self.sim._last_obs[0] if self.sim._last_obs is not None else self.sim._reset_obs
        return {"observation": o["observation"].copy(), "achieved_goal": o["achieved_goal"].copy(), "desired_goal": self.task.get_goal().astype(np.float32)}

    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None) -> Tuple[Dict[str, np.ndarray], Dict[str, Any]]:
        if _EnvBase is not object:
            super().reset(seed=seed, options=options)
        # core.py:243-244: the task RNG is re-created from the seed on every reset (seed=None -> fresh OS entropy)
        self.task.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        with self.sim.no_rendering():
            self.robot.reset()
            self.task.reset()
        goal, objs = self.task.goal, getattr(self.task, "_object_positions", [])
        vec = self.sim._require()
        obs = vec.reset(goals=np.asarray(goal, dtype=np.float64)[None, :], object_positions=np.concatenate(objs)[None, :] if objs else None)
        self.sim._reset_obs = {k: v[0].cpu().numpy() for k, v in obs.items()}
        self.sim._last_obs = None
        observation = self._get_obs()
        info = {"is_success": bool(self.task.is_success(observation["achieved_goal"], self.task.get_goal()))}
        return observation, info

    def save_state(self) -> int:
        state_id = self.sim.save_state()
        self._saved_goal[state_id] = (self.task.goal, self.sim._last_obs, self.sim._reset_obs)
        return state_id

    def restore_state(self, state_id: int) -> None:
        self.sim.restore_state(state_id)
        self.task.goal, self.sim._last_obs, self.sim._reset_obs = self._saved_goal[state_id]

    def remove_state(self, state_id: int) -> None:
        self._saved_goal.pop(state_id)
        self.sim.remove_state(state_id)

    def step(self, action: np.ndarray) -> Tuple[Dict[str, np.ndarray], float, bool, bool, Dict[str, Any]]:
        self.robot.set_action(action)
        self.sim.step()
        observation = self._get_obs()
        terminated = bool(self.sim._last_obs[2])      # is_success, computed in the kernel on the float32 goals (core.py:285)
        truncated = False
        info = {"is_success": terminated}
        reward = float(self.sim._last_obs[1])         # compute_reward, same kernel (core.py:288)
        return observation, reward, terminated, truncated, info

    def close(self) -> None:
        self.sim.close()

    def render(self, *args, **kwargs):
        raise NotImplementedError("rendering is out of scope for the B200 backend (and broken in the reference fork, core.py:294-335)")
