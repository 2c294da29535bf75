"""Drop-in for the reference's sim facade, reference panda_gym/pybullet.py:16-799 (class PyBullet), backed by the CUDA
environment instead of a pybullet physics client.

The reference's step is `robot.set_action(a); sim.step()`; here the controller, the 20 sub-steps and the observation are one
fused kernel launch, so `Panda.set_action` parks the action on the facade and `PyBullet.step()` launches it.  Getters read the
device state (pg_get_state, pg_get_link_state).  Used without a task (``PyBullet()`` directly, as reference test/pybullet_test.py does)
the same class drives a bare world: robot + free bodies + raw joint motors (pg_create_bare / pg_set_motors / pg_sim_step).
"""
from contextlib import contextmanager
from typing import Iterator, Optional

import numpy as np

from .. import _lib


class error(RuntimeError):
    """Stand-in for pybullet.error (raised on an unknown state id, reference test/save_and_restore_test.py:30-36)."""


class PyBullet:
    """Two modes.  *Bound* (constructed through a Panda*Env class, `_bind`): backed by a one-env task handle; the controller, the
    sub-steps and the observation are one fused launch.  *Unbound* (``PyBullet()`` used directly, as the reference's own
    test/pybullet_test.py does): `loadURDF` / `create_box` / `create_cylinder` / `create_table` / `create_plane` describe a bare
    world (robot, <= 2 free bodies, table top at z = 0, ground plane) that is created on the device at first use and advanced by
    ``step()`` with the motors ``control_joints`` set -- the same sub-step kernels, no task."""

    def __init__(self, render: bool = False, n_substeps: int = 20, background_color: Optional[np.ndarray] = None) -> None:
        if render:
            raise NotImplementedError("the B200 backend has no on-screen GUI (reference pybullet.py:34 p.GUI); PyBullet.render gives depth / colour / point clouds off-screen")
        if int(n_substeps) < 1:
            raise ValueError("n_substeps must be >= 1")
        self.n_substeps = int(n_substeps)
        self.timestep = 1.0 / 500
        self._bodies_idx = {}
        self._vec = None
        self._pending_action = None
        self._pending_orientation = None
        self._task_name = None
        self._last_obs = None
        # unbound mode
        self._bare = None
        self._robot_base = None
        self._free = []                 # dynamic bodies in creation order: dicts for PandaBareWorld
        self._free_names = []
        self._table_rect = None
        self._ground_z = None
        self._device, self._precision = 0, "f32"

    # ---- wiring (called by the env classes) ----------------------------------------------------------------------
    def _bind(self, task_name: str, reward_type: str, control_type: str, device: int = 0, precision: str = "f32", **task_params) -> None:
        from ..vec_env import PandaVecEnv
        self._task_name = task_name
        self._vec = PandaVecEnv(task_name, 1, reward_type=reward_type, control_type=control_type, device=device, precision=precision, auto_reset=False,
                                n_substeps=self.n_substeps, **task_params)

    def _require(self):
        if self._vec is None:
            raise _lib.PandaB200Error("the sim facade is not bound to a task yet (construct it through a Panda*Env class)")
        return self._vec

    def _world(self):
        """The bare world of the unbound facade, (re)created on the device when bodies were added since its last use."""
        from ..bare_world import PandaBareWorld
        if self._bare is None:
            self._bare = PandaBareWorld(1, robot_base=self._robot_base, bodies=self._free, table_rect=self._table_rect, ground_z=self._ground_z,
                                        device=self._device, precision=self._precision)
        return self._bare

    def _invalidate(self) -> None:
        """A body was added: carry the state of the existing bodies over into a new world."""
        if self._bare is None:
            return
        old = self._bare.get_state()[0].cpu().numpy()
        mot = self._bare.get_motors()
        nb_old = self._bare.n_bodies
        self._bare.close(); self._bare = None
        w = self._world()
        st = w.get_state()[0].cpu().numpy()
        st[:18] = old[:18]; st[18:18 + 13 * nb_old] = old[18:18 + 13 * nb_old]
        import torch
        w.set_state(torch.as_tensor(st[None, :])); w.set_motors(mot)

    def _backend(self):
        return self._vec if self._vec is not None else self._world()

    @property
    def dt(self):
        """reference pybullet.py:47-50"""
        return self.timestep * self.n_substeps

    def step(self) -> None:
        """reference pybullet.py:52-55 (n_substeps x stepSimulation).  Bound: fused with the pending Panda.set_action.  Unbound: the
        bare world advances with the motors `control_joints` left."""
        import torch
        if self._vec is None:
            self._world().step(self.n_substeps)
            return
        vec = self._vec
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
        if self._bare is not None:
            self._bare.close()
            self._bare = None

    # ---- state snapshots (reference pybullet.py:61-68, 266-280) ----------------------------------------------------
    def save_state(self) -> int:
        return self._backend().save_state()

    def restore_state(self, state_id: int) -> None:
        try:
            self._backend().restore_state(state_id)
        except _lib.PandaB200Error as e:
            raise error(str(e))

    def remove_state(self, state_id: int) -> None:
        try:
            self._backend().remove_state(state_id)
        except _lib.PandaB200Error as e:
            raise error(str(e))

    # ---- getters (reference pybullet.py:284-425) ---------------------------------------------------------------------
    def _state(self) -> np.ndarray:
        return self._backend().get_state()[0].cpu().numpy()

    def _obj_index(self, body: str) -> int:
        if self._vec is not None:
            idx = {"object": 0, "object1": 0, "object2": 1}.get(body)
        else:
            idx = self._free_names.index(body) if body in self._free_names else None
        if idx is None:
            raise KeyError(f"body {body!r} has no dynamic state in this scene")
        return idx

    def _obj(self, body: str) -> np.ndarray:
        idx = self._obj_index(body)
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

    def _link(self, link: int) -> np.ndarray:
        """getLinkState of any link 0..11, computed on the device (pg_get_link_state): pos3 quat4 lin3 ang3."""
        if not 0 <= int(link) <= 11:
            raise ValueError(f"the Panda has links 0..11, got {link}")
        return self._backend().link_state(int(link))[0].cpu().numpy()

    def get_link_position(self, body: str, link: int) -> np.ndarray:
        """reference pybullet.py:351-362"""
        return self._link(link)[0:3].copy()

    def get_link_orientation(self, body: str, link: int) -> np.ndarray:
        """reference pybullet.py:364-375"""
        return self._link(link)[3:7].copy()

    def get_link_velocity(self, body: str, link: int) -> np.ndarray:
        """reference pybullet.py:377-388"""
        return self._link(link)[7:10].copy()

    def get_link_angular_velocity(self, body: str, link: int) -> np.ndarray:
        """reference pybullet.py:390-400"""
        return self._link(link)[10:13].copy()

    # ---- setters / control ---------------------------------------------------------------------------------------------
    def set_base_pose(self, body: str, position: np.ndarray, orientation: np.ndarray) -> None:
        """reference pybullet.py:427-439 (resetBasePositionAndOrientation: velocities are zeroed); ghost targets are ignored."""
        import torch
        if self._vec is not None and body.startswith("target"):
            return
        if self._vec is None and body not in self._free_names:
            if body in self._bodies_idx:
                return                       # ghost / static body: no dynamic state
            raise KeyError(body)
        idx = self._obj_index(body)
        orientation = np.asarray(orientation, dtype=np.float64)
        if len(orientation) == 3:
            orientation = quaternion_from_euler(orientation)
        be = self._backend()
        row = be.get_state()[0].cpu().numpy()
        row[18 + 13 * idx: 18 + 13 * idx + 13] = np.concatenate([np.asarray(position, dtype=np.float64), orientation / np.linalg.norm(orientation), np.zeros(6)])
        be.set_state(torch.as_tensor(row[None, :]))

    def set_joint_angles(self, body: str, joints: np.ndarray, angles: np.ndarray) -> None:
        import torch
        be = self._backend()
        row = be.get_state()[0].cpu().numpy()
        for j, a in zip(joints, angles):
            row[_dof(int(j))] = a
            row[9 + _dof(int(j))] = 0.0
        be.set_state(torch.as_tensor(row[None, :]))

    def set_joint_angle(self, body: str, joint: int, angle: float) -> None:
        self.set_joint_angles(body, [joint], [angle])

    def control_joints(self, body: str, joints: np.ndarray, target_angles: np.ndarray, forces: np.ndarray) -> None:
        """reference pybullet.py:462-477.  Unbound: POSITION_CONTROL motors of the bare world.  Bound: the motor targets of a task
        env are produced inside the fused step kernel from the action (Panda.set_action), so raw targets cannot be injected there."""
        if self._vec is not None:
            raise NotImplementedError("raw motor targets are produced inside the fused step kernel; drive the robot through Panda.set_action")
        self._world().control_joints(joints, target_angles, forces)

    def inverse_kinematics(self, body: str, link: int, position: np.ndarray, orientation: np.ndarray) -> np.ndarray:
        """reference pybullet.py:479-497 -- 20 DLS iterations on `link` from the current joint state (runs on the device); all nine joints."""
        pos, orn = np.asarray(position, dtype=np.float64), np.asarray(orientation, dtype=np.float64)
        if self._vec is not None:
            return self._vec.inverse_kinematics(pos[None], orn[None], link=int(link))[0].cpu().numpy()
        return self._world().inverse_kinematics(int(link), pos, orn)[0].cpu().numpy()

    # ---- scene construction (reference pybullet.py:510-799).  Bound: the six task scenes are compiled into the kernels and these
    # only record names.  Unbound: they describe the bare world.
    def loadURDF(self, body_name: str, **kwargs) -> None:
        self._bodies_idx[body_name] = len(self._bodies_idx)
        if self._vec is None and self._task_name is None:
            if "panda" not in str(kwargs.get("fileName", "franka_panda/panda.urdf")):
                raise NotImplementedError("the B200 backend simulates franka_panda/panda.urdf only")
            if self._robot_base is not None:
                raise NotImplementedError("one robot per world")
            self._robot_base = [float(x) for x in kwargs.get("basePosition", (0.0, 0.0, 0.0))]
            self._invalidate()

    def _add_free(self, body_name: str, desc: dict) -> None:
        if len(self._free) >= 2:
            raise NotImplementedError("a bare world holds at most two free bodies (the task scenes never need more)")
        self._free.append(desc); self._free_names.append(body_name)
        self._invalidate()

    def create_box(self, body_name: str, half_extents, mass: float, position, rgba_color=None, specular_color=None, ghost: bool = False,
                   lateral_friction: Optional[float] = None, spinning_friction: Optional[float] = None, texture: Optional[str] = None) -> None:
        """reference pybullet.py:531-582.  Ghosts and static (mass 0) boxes have no dynamic state; a dynamic box becomes a free body."""
        self._bodies_idx[body_name] = len(self._bodies_idx)
        if self._task_name is not None or ghost or float(mass) == 0.0:
            return
        self._add_free(body_name, {"shape": "box", "half_extents": [float(x) for x in half_extents], "mass": float(mass), "position": [float(x) for x in position],
                                   "lateral_friction": 0.5 if lateral_friction is None else float(lateral_friction)})

    def create_cylinder(self, body_name: str, radius: float, height: float, mass: float, position, rgba_color=None, specular_color=None, ghost: bool = False,
                        lateral_friction: Optional[float] = None, spinning_friction: Optional[float] = None) -> None:
        """reference pybullet.py:584-632"""
        self._bodies_idx[body_name] = len(self._bodies_idx)
        if self._task_name is not None or ghost or float(mass) == 0.0:
            return
        self._add_free(body_name, {"shape": "cylinder", "radius": float(radius), "height": float(height), "mass": float(mass), "position": [float(x) for x in position],
                                   "lateral_friction": 0.5 if lateral_friction is None else float(lateral_friction)})

    def create_sphere(self, body_name: str, radius: float, mass: float, position, rgba_color=None, specular_color=None, ghost: bool = False,
                      lateral_friction: Optional[float] = None, spinning_friction: Optional[float] = None) -> None:
        """reference pybullet.py:634-677.  The tasks only create ghost spheres (goal markers); dynamic spheres are not simulated."""
        self._bodies_idx[body_name] = len(self._bodies_idx)
        if self._task_name is None and not ghost and float(mass) != 0.0:
            raise NotImplementedError("dynamic spheres are not part of any panda_gym task scene; the B200 backend simulates boxes and cylinders")

    def create_plane(self, z_offset: float) -> None:
        """reference pybullet.py:726-739: top surface at z_offset"""
        self._bodies_idx["plane"] = len(self._bodies_idx)
        if self._task_name is None:
            self._ground_z = float(z_offset)
            self._invalidate()

    def create_table(self, length: float, width: float, height: float, x_offset: float = 0.0, lateral_friction: Optional[float] = None,
                     spinning_friction: Optional[float] = None) -> None:
        """reference pybullet.py:741-771: top at z = 0, centred in y"""
        self._bodies_idx["table"] = len(self._bodies_idx)
        if self._task_name is None:
            self._table_rect = [x_offset - length / 2, x_offset + length / 2, -width / 2, width / 2]
            self._invalidate()

    def set_lateral_friction(self, body: str, link: int, lateral_friction: float) -> None:
        """reference pybullet.py:773-785.  Bound: the task scenes' coefficients (fingers 1.0, puck 0.04, default 0.5) are scene
        constants.  Unbound: takes effect for free bodies (the world is rebuilt with the new coefficient)."""
        if self._vec is None and body in self._free_names:
            self._free[self._free_names.index(body)]["lateral_friction"] = float(lateral_friction)
            self._invalidate()

    def set_spinning_friction(self, body: str, link: int, spinning_friction: float) -> None:
        """reference pybullet.py:787-799 (the finger pads' 0.001 of panda.py:47-50 is a scene constant of the kernels)."""

    def place_visualizer(self, target_position: np.ndarray, distance: float, yaw: float, pitch: float) -> None:
        pass

    @contextmanager
    def no_rendering(self) -> Iterator[None]:
        yield

    def get_cam2world_transforms(self, width: int = 480, height: int = 480, target_position=None, distance: float = 1.4, yaw: float = 45, pitch: float = -30, roll: float = 0):
        """reference pybullet.py:70-107: (view_matrix, proj_matrix) as pybullet's 16-tuples (column-major) and inv(P V).  Host matrix
        algebra on 4x4s, as in the reference (computeViewMatrixFromYawPitchRoll, up axis 2; computeProjectionMatrixFOV 60 deg, 0.1, 100)."""
        t = np.zeros(3) if target_position is None else np.asarray(target_position, dtype=np.float64)
        V, P = view_matrix(t, distance, yaw, pitch, roll), projection_matrix(60.0, float(width) / height, 0.1, 100.0)
        return tuple(V.reshape(-1, order="F")), tuple(P.reshape(-1, order="F")), np.linalg.inv(P @ V)

    def deproject(self, depth, pixels, tran_pix_world, width: int = 480, height: int = 480):
        """reference pybullet.py:109-147: pixel coordinates (x, y) + depth-buffer image -> world points."""
        pixels = np.asarray(pixels)
        x = pixels[:, 0] * 1 / width * 2 - 1
        y = (height - pixels[:, 1]) * 1 / height * 2 - 1
        z = 2 * np.asarray(depth)[pixels[:, 1], pixels[:, 0]] - 1
        pts = (tran_pix_world @ np.stack([x, y, z, np.ones_like(z)], axis=1).T).T
        return (pts / pts[:, 3:4])[:, :3]

    def render(self, width: int = 480, height: int = 480, target_position=None, distance: float = 1.4, yaw: float = 45, pitch: float = -30, roll: float = 0, waypoints=None):
        """reference pybullet.py:149-264: (rgb, depth, points, colors, pixels_2d, waypoints_proj).  The image is ray-cast on the device
        (pg_render; the robot is drawn as its physics boxes, its meshes are not part of the reference tree); the compaction of the
        valid pixels into the point list is host-side boolean indexing, in the reference's row-major pixel order."""
        t = np.zeros(3) if target_position is None else np.asarray(target_position, dtype=np.float64)
        r = self._backend().render(width, height, t, distance, yaw, pitch, roll, crop=True, rgb=True, points=True)
        rgb, depth = r["rgb"][0].cpu().numpy(), r["depth"][0].cpu().numpy()
        valid = r["valid"][0].cpu().numpy().reshape(-1)
        points = r["points"][0].cpu().numpy().reshape(-1, 3)[valid].astype(np.float64)
        colors = rgb.reshape(-1, 3)[valid]
        rows, cols = np.divmod(np.nonzero(valid)[0], width)
        pixels_2d = np.stack([cols.astype(np.float64), rows.astype(np.float64)], axis=1)      # (x, y) of each kept pixel, y measured from the top
        waypoints_proj = []
        if waypoints is not None:
            PV = projection_matrix(60.0, float(width) / height, 0.1, 100.0) @ view_matrix(t, distance, yaw, pitch, roll)
            for p in waypoints:
                x, y, z, w = PV @ np.array([p[0], p[1], p[2], 1.0])
                waypoints_proj.append([int((x / w + 1) / 2 * width), int(height - (y / w + 1) / 2 * height)])
        return rgb, depth, points, colors, pixels_2d, waypoints_proj

    _reset_obs = None


def view_matrix(target, distance, yaw, pitch, roll) -> np.ndarray:
    """pybullet computeViewMatrixFromYawPitchRoll(upAxisIndex=2) as a 4x4 (row-major maths): the eye sits at R (0, -distance, 0) + target with
    R = Rz(yaw) Ry(roll) Rx(pitch), looking at the target, up = R e_z."""
    y, p, r = np.radians([yaw, pitch, roll])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(r), 0, np.sin(r)], [0, 1, 0], [-np.sin(r), 0, np.cos(r)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(p), -np.sin(p)], [0, np.sin(p), np.cos(p)]])
    R = Rz @ Ry @ Rx
    target = np.asarray(target, dtype=np.float64)
    eye, up = R @ np.array([0.0, -distance, 0.0]) + target, R @ np.array([0.0, 0.0, 1.0])
    f = (target - eye) / np.linalg.norm(target - eye)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    V = np.eye(4)
    V[0, :3], V[1, :3], V[2, :3] = s, u, -f
    V[:3, 3] = -V[:3, :3] @ eye
    return V


def projection_matrix(fov, aspect, near, far) -> np.ndarray:
    """pybullet computeProjectionMatrixFOV (the OpenGL perspective matrix)."""
    ys = 1.0 / np.tan(np.radians(fov) / 2)
    return np.array([[ys / aspect, 0, 0, 0], [0, ys, 0, 0], [0, 0, (near + far) / (near - far), 2 * near * far / (near - far)], [0, 0, -1, 0]])


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
