"""Generates tests/golden/*.npz by executing the REFERENCE's own Python code from /root/reference (read-only).

What can run here: panda_gym/utils.py (numpy only) and the Task classes' pure-numpy methods
(_sample_goal / _sample_object / is_success / compute_reward) once `gymnasium`, `pybullet*` are stubbed out and the task is
given a mock sim (no physics call is made by those methods).  `np.bool8` (removed in numpy 2) is aliased to np.bool_.
PyBullet itself is absent, so no trajectory goldens can be produced.

    python tests/golden/make_golden.py      # run in the build container only; the .npz files are committed
"""
import contextlib
import os
import sys
import types
from unittest import mock

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference_tasks():
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    gym = types.ModuleType("gymnasium")
    gym.Env = object
    gym.spaces = types.ModuleType("gymnasium.spaces")
    gym.spaces.Box = lambda *a, **k: None
    gym.spaces.Dict = lambda *a, **k: None
    gym.spaces.Space = object
    gym.utils = types.ModuleType("gymnasium.utils")
    gym.utils.seeding = types.ModuleType("gymnasium.utils.seeding")
    gym.utils.seeding.np_random = lambda seed=None: (np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed))), seed)
    gym.envs = types.ModuleType("gymnasium.envs")
    gym.envs.registration = types.ModuleType("gymnasium.envs.registration")
    gym.envs.registration.register = lambda **k: None
    for name, m in [("gymnasium", gym), ("gymnasium.spaces", gym.spaces), ("gymnasium.utils", gym.utils), ("gymnasium.utils.seeding", gym.utils.seeding),
                    ("gymnasium.envs", gym.envs), ("gymnasium.envs.registration", gym.envs.registration)]:
        sys.modules[name] = m
    for name in ["pybullet", "pybullet_data", "pybullet_utils", "pybullet_utils.bullet_client", "cv2"]:
        sys.modules[name] = mock.MagicMock()
    sys.path.insert(0, REF)
    from panda_gym.envs.tasks.flip import Flip
    from panda_gym.envs.tasks.pick_and_place import PickAndPlace
    from panda_gym.envs.tasks.push import Push
    from panda_gym.envs.tasks.reach import Reach
    from panda_gym.envs.tasks.slide import Slide
    from panda_gym.envs.tasks.stack import Stack
    from panda_gym import utils
    return dict(reach=Reach, push=Push, slide=Slide, pick_and_place=PickAndPlace, stack=Stack, flip=Flip), utils


def mock_sim():
    sim = mock.MagicMock()
    sim.no_rendering = lambda: contextlib.nullcontext()
    return sim


def main():
    tasks, utils = import_reference_tasks()
    rng = np.random.default_rng(20240101)
    out = {}
    for name, cls in tasks.items():
        for rt in ("sparse", "dense"):
            task = cls(mock_sim(), get_ee_position=lambda: np.zeros(3), reward_type=rt) if name == "reach" else cls(mock_sim(), reward_type=rt)
            G = {"stack": 6, "flip": 4}.get(name, 3)
            thr = task.distance_threshold
            m = 1024
            for dt in (np.float32, np.float64):
                dg = rng.uniform(-0.3, 0.3, (m, G)).astype(dt)
                ag = (dg + rng.normal(0, thr / np.sqrt(G), (m, G))).astype(dt)
                if name == "flip":
                    ag /= np.linalg.norm(ag, axis=-1, keepdims=True); dg /= np.linalg.norm(dg, axis=-1, keepdims=True)
                ag[:8] = dg[:8]
                ag[8, 0] = dg[8, 0] + dt(thr)
                if name == "flip":   # the reference's angle_distance is not batched (np.inner -> [m,m]); call it row by row, as step() does
                    rew = np.array([task.compute_reward(ag[i], dg[i], {}) for i in range(m)], dtype=np.float32)
                    suc = np.array([bool(task.is_success(ag[i], dg[i])) for i in range(m)])
                else:
                    rew = np.asarray(task.compute_reward(ag, dg, {}), dtype=np.float32)
                    suc = np.asarray(task.is_success(ag, dg), dtype=bool)
                    r1 = np.array([task.compute_reward(ag[i], dg[i], {}) for i in range(64)], dtype=np.float32)
                    assert r1.tobytes() == rew[:64].tobytes()          # batched == per-row in the reference itself
                k = f"{name}_{rt}_{np.dtype(dt).name}"
                out[k + "_ag"], out[k + "_dg"], out[k + "_reward"], out[k + "_success"] = ag, dg, rew, suc
        # seeded samplers (core.py:243-244: task.np_random = seeding.np_random(seed)[0]; draw order of Task.reset)
        task = cls(mock_sim(), get_ee_position=lambda: np.zeros(3)) if name == "reach" else cls(mock_sim())
        if name != "flip":
            goals, objs = [], []
            for seed in range(16):
                task.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
                goals.append(np.asarray(task._sample_goal(), dtype=np.float64))
                if name == "stack":
                    objs.append(np.concatenate(task._sample_objects()))
                elif name != "reach":
                    objs.append(np.asarray(task._sample_object(), dtype=np.float64))
            out[f"{name}_seeded_goals"] = np.array(goals)
            if objs:
                out[f"{name}_seeded_objects"] = np.array(objs)
    a = rng.normal(size=(256, 3)); b = rng.normal(size=(256, 3))
    out["utils_distance_a"], out["utils_distance_b"], out["utils_distance"] = a, b, utils.distance(a, b)
    qa = rng.normal(size=(64, 4)); qa /= np.linalg.norm(qa, axis=-1, keepdims=True)
    qb = rng.normal(size=(64, 4)); qb /= np.linalg.norm(qb, axis=-1, keepdims=True)
    out["utils_angle_a"], out["utils_angle_b"] = qa, qb
    out["utils_angle_distance_rows"] = np.array([utils.angle_distance(qa[i], qb[i]) for i in range(64)])
    np.savez_compressed(os.path.join(OUT, "reference_rewards.npz"), **out)
    print("wrote", os.path.join(OUT, "reference_rewards.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
