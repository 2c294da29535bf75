"""Golden trajectories from the UNMODIFIED reference running on the real engine -- only possible on a machine that has
pybullet==3.2.5 and gymnasium installed (neither is installable in the build container: SURVEY.md section 8c).  Run it there:

    python tests/golden/make_pybullet_golden.py [/path/to/reference]

It imports the reference package as is (np.bool8 is aliased for numpy >= 2), and for every task x control type records, for fixed
seeds and seeded action sequences: the reset observation, goal, object placement, and per step the action, the Dict observation,
reward, terminated, joint angles / velocities and object poses.  tests/test_pybullet_golden.py consumes the file when present: the
oracle (CPU) and the CUDA path (GPU) are then compared with the real engine at the north-star tolerances, and the oracle's header may
drop "parity unpinned" for whatever holds."""
import importlib.util
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pybullet_trajectories.npz")
TASKS = ["Reach", "Push", "Slide", "PickAndPlace", "Stack", "Flip"]


def main():
    if importlib.util.find_spec("pybullet") is None or importlib.util.find_spec("gymnasium") is None:
        print("pybullet / gymnasium are not installed here: no golden file written (tests/test_pybullet_golden.py will skip)")
        return 1
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    sys.path.insert(0, ref)
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    import gymnasium as gym
    import panda_gym  # noqa: F401  (registers the ids)
    out = {}
    for task in TASKS:
        for joints in (False, True):
            env_id = f"Panda{task}{'Joints' if joints else ''}-v3"
            env = gym.make(env_id)
            for seed in range(4):
                obs, _ = env.reset(seed=seed)
                rng = np.random.default_rng(1000 + seed)
                key = f"{env_id}/seed{seed}"
                rec = {"obs": [obs["observation"]], "ag": [obs["achieved_goal"]], "dg": [obs["desired_goal"]], "act": [], "rew": [], "term": [], "q": [], "qd": [], "obj": []}
                sim, robot = env.unwrapped.sim, env.unwrapped.robot
                for t in range(50):
                    a = rng.uniform(-1, 1, env.action_space.shape).astype(np.float32)
                    obs, r, term, trunc, _ = env.step(a)
                    rec["act"].append(a); rec["obs"].append(obs["observation"]); rec["ag"].append(obs["achieved_goal"]); rec["dg"].append(obs["desired_goal"])
                    rec["rew"].append(r); rec["term"].append(term)
                    rec["q"].append([sim.get_joint_angle("panda", j) for j in robot.joint_indices]); rec["qd"].append([sim.get_joint_velocity("panda", j) for j in robot.joint_indices])
                    objs = [b for b in ("object", "object1", "object2") if b in sim._bodies_idx]
                    rec["obj"].append(np.concatenate([np.concatenate([sim.get_base_position(b), sim.get_base_orientation(b), sim.get_base_velocity(b), sim.get_base_angular_velocity(b)]) for b in objs]) if objs else np.zeros(0))
                    if term or trunc:
                        break
                for k, v in rec.items():
                    out[f"{key}/{k}"] = np.asarray(v)
            env.close()
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays")
    return 0


if __name__ == "__main__":
    sys.exit(main())
