"""Drop-in for the reference's sim facade, reference panda_gym/pybullet.py:16-799 (class PyBullet), backed by the CUDA
environment instead of a pybullet physics client.

The reference's step is `robot.set_action(a); sim.step()`; here the controller, the 20 sub-steps and the observation are one
fused kernel launch, so `Panda.set_action` parks the action on the facade and `PyBullet.step()` launches it.  Getters read the
device state (pg_get_state).  Rendering (`render`, `deproject`, `get_cam2world_transforms`, reference pybullet.py:70-264) is out of scope.
"""
from contextlib import contextmanager
from typing import Iterator, Optional

import numpy as np

from .. import _lib


class error(RuntimeError):
    """Stand-in for pybullet.error (raised on an unknown state id, reference test/save_and_restore_test.py:30-36)."""


class PyBullet:
    def __init__(self, render: bool = False, n_substeps: int = 20, background_color: Optional[np.ndarray] = None) -> None:
        if render:
            raise NotImplementedError("the B200 backend has no renderer (reference pybullet.py:149-264 is out of scope)")
        if n_substeps != 20:
            raise NotImplementedError("the fused step kernel is built for the reference's 20 sub-steps (pybullet.py:26)")
        self.n_substeps = n_substeps
        self.timestep = 1.0 / 500
        self._bodies_idx = {}
        self._vec = None
        self._pending_action = None
        self._pending_orientation = None
        self._task_name = None
        self._last_obs = None

    # ---- wiring (called by the env classes) ----------------------------------------------------------------------
    def _bind(self, task_name: str, reward_type: str, control_type: str, device: int = 0, precision: str = "f32") -> None:
        from ..vec_env import PandaVecEnv
        self._task_name = task_name
        self._vec = PandaVecEnv(task_name, 1, reward_type=reward_type, control_type=control_type, device=device, precision=precision, auto_reset=False)

    def _require(self):
        if self._vec is None:
            raise _lib.PandaB200Error("the sim facade is not bound to a task yet (construct it through a Panda*Env class)")
        return self._vec

    @property
    def dt(self):
        """reference pybullet.py:47-50"""
        return self.timestep * self.n_substeps

    def step(self) -> None:
        """reference pybullet.py:52-55 (20 x stepSimulation) -- fused with the pending Panda.set_action."""
        import torch
        vec = self._require()
        if self._pending_action is None:
            raise _lib.PandaB200Error("PyBullet.step() needs a pending action: call robot.set_action(action) first (panda.py:52-70)")
        a = torch.as_tensor(np.asarray(self._pending_action, dtype=np.float32)[None, :])
        tq = None if self._pending_orientation is None else torch.as_tensor(np.asarray(self._pending_orientation, dtype=np.float32)[None, :])
        obs, rew, term, trunc, _ = vec.step(a, target_orientation=tq)
        self._pending_orientation = None
        self._last_obs = ({k: v[0].cpu().numpy() for k, v in obs.items()}, float(rew[0]), bool(term[0]))
        self._pending_action = None

    def close(self) -> None:
        if self._vec is not None:
            self._vec.close()
            self._vec = None

    # ---- state snapshots (reference pybullet.py:61-68, 266-280) ----------------------------------------------------
    def save_state(self) -> int:
        return self._require().save_state()

    def restore_state(self, state_id: int) -> None:
        try:
            self._require().restore_state(state_id)
        except _lib.PandaB200Error as e:
            raise error(str(e))

    def remove_state(self, state_id: int) -> None:
        try:
            self._require().remove_state(state_id)
        except _lib.PandaB200Error as e:
            raise error(str(e))

    # ---- getters (reference pybullet.py:284-425) ---------------------------------------------------------------------
    def _state(self) -> np.ndarray:
        return self._require().get_state()[0].cpu().numpy()

    def _obj(self, body: str) -> np.ndarray:
        idx = {"object": 0, "object1": 0, "object2": 1}.get(body)
        if idx is None:
            raise KeyError(f"body {body!r} has no dynamic state in this scene")
        return self._state()[18 + 13 * idx: 18 + 13 * idx + 13]

    def get_base_position(self, body: str) -> np.ndarray:
        return self._obj(body)[0:3].copy()

    def get_base_orientation(self, body: str) -> np.ndarray:
        return self._obj(body)[3:7].copy()

    def get_base_rotation(self, body: str, type: str = "euler") -> np.ndarray:
        q = self.get_base_orientation(body)
        if type == "euler":
            return euler_from_quaternion(q)
        if type == "quaternion":
            return q
        raise ValueError("""type must be "euler" or "quaternion".""")

    def get_base_velocity(self, body: str) -> np.ndarray:
        return self._obj(body)[7:10].copy()

    def get_base_angular_velocity(self, body: str) -> np.ndarray:
        return self._obj(body)[10:13].copy()

    def get_joint_angle(self, body: str, joint: int) -> float:
        return float(self._state()[_dof(joint)])

    def get_joint_velocity(self, body: str, joint: int) -> float:
        return float(self._state()[9 + _dof(joint)])

    def get_link_position(self, body: str, link: int) -> np.ndarray:
        if link != 11 or self._last_obs is None and self._reset_obs is None:
            raise NotImplementedError("only the end-effector link (11) is exposed by the fused kernel")
        return np.asarray((self._last_obs[0] if self._last_obs else self._reset_obs)["observation"][0:3], dtype=np.float64)

    def get_link_velocity(self, body: str, link: int) -> np.ndarray:
        if link != 11:
            raise NotImplementedError("only the end-effector link (11) is exposed by the fused kernel")
        return np.asarray((self._last_obs[0] if self._last_obs else self._reset_obs)["observation"][3:6], dtype=np.float64)

    # ---- setters / control ---------------------------------------------------------------------------------------------
    def set_base_pose(self, body: str, position: np.ndarray, orientation: np.ndarray) -> None:
        """reference pybullet.py:427-439 (resetBasePositionAndOrientation: velocities are zeroed); ghost targets are ignored."""
        import torch
        if body.startswith("target"):
            return
        idx = {"object": 0, "object1": 0, "object2": 1}[body]
        orientation = np.asarray(orientation, dtype=np.float64)
        if len(orientation) == 3:
            orientation = quaternion_from_euler(orientation)
        s = self._require().get_state()
        row = s[0].cpu().numpy()
        row[18 + 13 * idx: 18 + 13 * idx + 13] = np.concatenate([np.asarray(position, dtype=np.float64), orientation / np.linalg.norm(orientation), np.zeros(6)])
        self._vec.set_state(torch.as_tensor(row[None, :]))

    def set_joint_angles(self, body: str, joints: np.ndarray, angles: np.ndarray) -> None:
        import torch
        row = self._require().get_state()[0].cpu().numpy()
        for j, a in zip(joints, angles):
            row[_dof(int(j))] = a
            row[9 + _dof(int(j))] = 0.0
        self._vec.set_state(torch.as_tensor(row[None, :]))

    def set_joint_angle(self, body: str, joint: int, angle: float) -> None:
        self.set_joint_angles(body, [joint], [angle])

    def control_joints(self, body: str, joints: np.ndarray, target_angles: np.ndarray, forces: np.ndarray) -> None:
        raise NotImplementedError("raw motor targets are produced inside the fused step kernel; drive the robot through Panda.set_action")

    def inverse_kinematics(self, body: str, link: int, position: np.ndarray, orientation: np.ndarray) -> np.ndarray:
        """reference pybullet.py:479-497 -- 20 DLS iterations on link 11 from the current joint state (runs on the device)."""
        if link != 11:
            raise NotImplementedError("inverse kinematics is implemented for the end-effector link (11)")
        q7 = self._require().inverse_kinematics(np.asarray(position, dtype=np.float64)[None], np.asarray(orientation, dtype=np.float64)[None])[0].cpu().numpy()
        st = self._state()
        return np.concatenate([q7, st[7:9]])

    # ---- scene construction (reference pybullet.py:531-799): the scenes are compiled into the kernels; these record names only
    def loadURDF(self, body_name: str, **kwargs) -> None:
        self._bodies_idx[body_name] = len(self._bodies_idx)

    def _create(self, body_name: str, *args, **kwargs) -> None:
        self._bodies_idx[body_name] = len(self._bodies_idx)

    create_box = create_cylinder = create_sphere = _create

    def create_plane(self, z_offset: float) -> None:
        self._create("plane")

    def create_table(self, length: float, width: float, height: float, x_offset: float = 0.0, **kwargs) -> None:
        self._create("table")

    def set_lateral_friction(self, body: str, link: int, lateral_friction: float) -> None:
        pass

    def set_spinning_friction(self, body: str, link: int, spinning_friction: float) -> None:
        pass

    def place_visualizer(self, target_position: np.ndarray, distance: float, yaw: float, pitch: float) -> None:
        pass

    @contextmanager
    def no_rendering(self) -> Iterator[None]:
        yield

    def render(self, *args, **kwargs):
        raise NotImplementedError("no renderer on the B200 backend")

    _reset_obs = None


def _dof(joint: int) -> int:
    m = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6, 9: 7, 10: 8}
    if joint not in m:
        raise ValueError(f"joint {joint} is fixed")
    return m[joint]


def euler_from_quaternion(q) -> np.ndarray:
    """pybullet getEulerFromQuaternion."""
    x, y, z, w = [float(v) for v in q]
    sarg = -2 * (x * z - w * y)
    if sarg <= -0.99999:
        return np.array([0.0, -0.5 * np.pi, 2 * np.arctan2(x, -y)])
    if sarg >= 0.99999:
        return np.array([0.0, 0.5 * np.pi, 2 * np.arctan2(-x, y)])
    return np.array([np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), np.arcsin(sarg), np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)])


def quaternion_from_euler(e) -> np.ndarray:
    r, p, y = [float(v) / 2 for v in e]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])
