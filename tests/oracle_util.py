"""ctypes wrapper of the CPU oracle (oracle/libpanda_oracle.so) -- test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
TASKS = {"reach": 0, "push": 1, "slide": 2, "pick_and_place": 3, "stack": 4, "flip": 5, "bare": 6}
OBS_DIM = {"reach": 6, "push": 18, "slide": 18, "pick_and_place": 19, "stack": 31, "flip": 20}
GOAL_DIM = {"reach": 3, "push": 3, "slide": 3, "pick_and_place": 3, "stack": 6, "flip": 4}
NOBJ = {"reach": 0, "push": 1, "slide": 1, "pick_and_place": 1, "stack": 2, "flip": 1}
BLOCKED = {"reach", "push", "slide"}
D = ctypes.c_double
_lib = None


def build_oracle():
    so = os.path.join(ORACLE_DIR, "libpanda_oracle.so")
    src = [os.path.join(ORACLE_DIR, f) for f in ("panda_oracle.c", "panda_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)
    return so


def load_oracle():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_oracle())
        vp = ctypes.c_void_p
        lib.po_create.restype = vp; lib.po_create.argtypes = [ctypes.c_int, D, D, D]
        lib.po_env_create.restype = vp; lib.po_env_create.argtypes = [ctypes.c_int] * 3
        lib.po_env_sim.restype = vp; lib.po_env_sim.argtypes = [vp]
        lib.po_destroy.argtypes = [vp]; lib.po_env_destroy.argtypes = [vp]
        lib.po_step.argtypes = [vp, ctypes.c_int]
        lib.po_control_joint.argtypes = [vp, ctypes.c_int, D, D]
        lib.po_reset_joint.argtypes = [vp, ctypes.c_int, D]
        lib.po_get_joint.argtypes = [vp, ctypes.c_int, vp, vp]
        lib.po_get_link_state.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp]
        lib.po_inverse_kinematics.argtypes = [vp, ctypes.c_int, vp, vp, vp]
        lib.po_add_box.argtypes = [vp, D, D, D, D, vp]
        for f in ("po_set_base_pose", "po_get_base_pose", "po_get_base_velocity", "po_set_base_velocity"):
            getattr(lib, f).argtypes = [vp, ctypes.c_int, vp, vp]
        lib.po_env_reset.argtypes = [vp] * 6
        lib.po_env_step.argtypes = [vp] * 7
        lib.po_env_step_oriented.argtypes = [vp, vp, vp, D, D, vp, vp, vp, vp, vp]
        lib.po_env_set_state.argtypes = [vp] * 3; lib.po_env_get_state.argtypes = [vp] * 3
        lib.po_save_state.argtypes = [vp, vp]; lib.po_restore_state.argtypes = [vp, vp]
        lib.po_last_num_contacts.argtypes = [vp]; lib.po_last_iterations.argtypes = [vp]
        lib.po_mass_matrix.argtypes = [vp, vp]
        lib.po_env_set_params.argtypes = [vp, ctypes.c_int, D]
        lib.po_env_step_batch.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp]
        lib.po_env_set_full_state.argtypes = [vp, vp]; lib.po_env_get_full_state.argtypes = [vp, vp]
        lib.po_set_static.argtypes = [vp, vp, vp]; lib.po_set_joint_state.argtypes = [vp, vp, vp, vp]
        lib.po_set_object_shape.argtypes = [vp, ctypes.c_int, ctypes.c_int, D, D, D, D, D]
        for f in ("po_compute_reward_f32", "po_compute_reward_f64"):
            getattr(lib, f).argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_long]
        for f in ("po_is_success_f32", "po_is_success_f64"):
            getattr(lib, f).argtypes = [ctypes.c_int, vp, vp, vp, ctypes.c_long]
        _lib = lib
    return _lib


def P(a):
    return a.ctypes.data


class OracleSim:
    """panda_gym.pybullet.PyBullet-level view of the oracle (bare robot at the origin + optional boxes)."""

    def __init__(self, task="bare", base=(0.0, 0.0, 0.0)):
        self.lib = load_oracle()
        self.h = self.lib.po_create(TASKS[task], *[float(b) for b in base])

    def close(self):
        self.lib.po_destroy(self.h)

    def step(self, n=20):
        self.lib.po_step(self.h, n)

    def control_joint(self, link, target, force):
        self.lib.po_control_joint(self.h, link, float(target), float(force))

    def reset_joint(self, link, angle):
        self.lib.po_reset_joint(self.h, link, float(angle))

    def joint(self, link):
        q, qd = D(), D()
        self.lib.po_get_joint(self.h, link, ctypes.addressof(q), ctypes.addressof(qd))
        return q.value, qd.value

    def link_state(self, link):
        p, q, v, w = np.zeros(3), np.zeros(4), np.zeros(3), np.zeros(3)
        self.lib.po_get_link_state(self.h, link, P(p), P(q), P(v), P(w))
        return p, q, v, w

    def ik(self, link, pos, quat):
        out = np.zeros(9)
        pos, quat = np.asarray(pos, np.float64), np.asarray(quat, np.float64)
        self.lib.po_inverse_kinematics(self.h, link, P(pos), P(quat), P(out))
        return out

    def add_box(self, half, mass, pos):
        pos = np.asarray(pos, np.float64)
        return self.lib.po_add_box(self.h, float(half[0]), float(half[1]), float(half[2]), float(mass), P(pos))

    def set_base_pose(self, o, pos, quat):
        pos, quat = np.asarray(pos, np.float64), np.asarray(quat, np.float64)
        self.lib.po_set_base_pose(self.h, o, P(pos), P(quat))

    def base_pose(self, o):
        p, q = np.zeros(3), np.zeros(4)
        self.lib.po_get_base_pose(self.h, o, P(p), P(q))
        return p, q

    def base_velocity(self, o):
        v, w = np.zeros(3), np.zeros(3)
        self.lib.po_get_base_velocity(self.h, o, P(v), P(w))
        return v, w


class OracleEnv:
    """RobotTaskEnv-level view of the oracle for one environment."""

    def __init__(self, task, control_type="ee", reward_type="sparse", n_substeps=20, distance_threshold=None):
        self.lib = load_oracle()
        self.task = task
        self.h = self.lib.po_env_create(TASKS[task], 0 if control_type == "ee" else 1, 0 if reward_type == "sparse" else 1)
        if n_substeps != 20 or distance_threshold is not None:
            self.lib.po_env_set_params(self.h, int(n_substeps), float({"stack": 0.1, "flip": 0.2}.get(task, 0.05) if distance_threshold is None else distance_threshold))
        self.sim = self.lib.po_env_sim(self.h)
        self.obs = np.zeros(OBS_DIM[task], np.float32)
        self.ag = np.zeros(GOAL_DIM[task], np.float32)
        self.dg = np.zeros(GOAL_DIM[task], np.float32)
        self.action_dim = (3 if control_type == "ee" else 7) + (0 if task in BLOCKED else 1)

    def close(self):
        self.lib.po_env_destroy(self.h)

    def reset(self, goal, objpos=None):
        goal = np.ascontiguousarray(np.resize(np.asarray(goal, np.float64), 6))
        objpos = np.ascontiguousarray(np.resize(np.asarray(objpos if objpos is not None else np.zeros(6), np.float64), 6))
        self.lib.po_env_reset(self.h, P(goal), P(objpos), P(self.obs), P(self.ag), P(self.dg))
        return self.obs.copy(), self.ag.copy(), self.dg.copy()

    def step(self, action):
        a = np.ascontiguousarray(action, np.float32)
        r, t = np.zeros(1, np.float32), np.zeros(1, np.uint8)
        self.lib.po_env_step(self.h, P(a), P(self.obs), P(self.ag), P(self.dg), P(r), P(t))
        return self.obs.copy(), self.ag.copy(), self.dg.copy(), float(r[0]), bool(t[0])

    def step_oriented(self, action, quat, ee_scale=0.05, finger_scale=0.2):
        a = np.ascontiguousarray(action, np.float32)
        tq = np.ascontiguousarray(quat, np.float64)
        r, t = np.zeros(1, np.float32), np.zeros(1, np.uint8)
        self.lib.po_env_step_oriented(self.h, P(a), P(tq), float(ee_scale), float(finger_scale), P(self.obs), P(self.ag), P(self.dg), P(r), P(t))
        return self.obs.copy(), self.ag.copy(), self.dg.copy(), float(r[0]), bool(t[0])

    def joints(self):
        q, qd = np.zeros(9), np.zeros(9)
        self.lib.po_env_get_state(self.h, P(q), P(qd))
        return q, qd

    def set_joints(self, q, qd):
        q, qd = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(qd, np.float64)
        self.lib.po_env_set_state(self.h, P(q), P(qd))

    def object_state(self, o):
        p, q, v, w = np.zeros(3), np.zeros(4), np.zeros(3), np.zeros(3)
        self.lib.po_get_base_pose(self.sim, o, P(p), P(q)); self.lib.po_get_base_velocity(self.sim, o, P(v), P(w))
        return np.concatenate([p, q, v, w])

    def contacts(self):
        return self.lib.po_last_num_contacts(self.sim), self.lib.po_last_iterations(self.sim)

    def set_full_state(self, st):
        """q(9) qd(9) | per object pos3 quat4 lin3 ang3 | goal(G): a row of pg_get_state without its trailing step counter."""
        st = np.ascontiguousarray(st, np.float64)
        self.lib.po_env_set_full_state(self.h, P(st))

    def full_state(self):
        st = np.zeros(18 + 13 * NOBJ[self.task] + GOAL_DIM[self.task])
        self.lib.po_env_get_full_state(self.h, P(st))
        return st


class OracleBatch:
    """n oracle envs stepped together by one C call (po_env_step_batch); the scripted-policy drivers fork one of these per core."""

    def __init__(self, task, n, control_type="ee"):
        self.lib = load_oracle()
        self.task, self.n = task, n
        ct = 0 if control_type == "ee" else 1
        self.handles = (ctypes.c_void_p * n)(*[self.lib.po_env_create(TASKS[task], ct, 0) for _ in range(n)])
        self.na = (3 if control_type == "ee" else 7) + (0 if task in BLOCKED else 1)
        self.no, self.ng = OBS_DIM[task], GOAL_DIM[task]
        self.obs = np.zeros((n, self.no), np.float32); self.ag = np.zeros((n, self.ng), np.float32); self.dg = np.zeros((n, self.ng), np.float32)
        self.rew = np.zeros(n, np.float32); self.term = np.zeros(n, np.uint8)

    def reset(self, goals, objs):
        for i in range(self.n):
            g = np.ascontiguousarray(np.resize(np.asarray(goals[i], np.float64), 6)); o = np.ascontiguousarray(np.resize(np.asarray(objs[i], np.float64), 6))
            self.lib.po_env_reset(self.handles[i], P(g), P(o), P(self.obs[i]), P(self.ag[i]), P(self.dg[i]))
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float32)
        self.lib.po_env_step_batch(self.handles, self.n, self.na, self.no, self.ng, P(a), P(self.obs), P(self.ag), P(self.dg), P(self.rew), P(self.term))
        return self.obs, self.rew, self.term

    def close(self):
        for h in self.handles:
            self.lib.po_env_destroy(h)


def reward_np(task, reward_type, ag, dg):
    """The reference's own arithmetic: utils.distance / angle_distance (row-wise) + tasks compute_reward, via numpy."""
    if task == "flip":
        d = 1 - np.einsum("...i,...i->...", ag, dg) ** 2
    else:
        d = np.linalg.norm(ag - dg, axis=-1)
    thr = {"stack": 0.1, "flip": 0.2}.get(task, 0.05)
    if reward_type == "sparse":
        return -np.array(d > thr, dtype=np.float32), np.array(d < thr)
    return -d.astype(np.float32), np.array(d < thr)
