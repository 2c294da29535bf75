"""Parity statistics on larger samples than the pytest cases (writes JSON to stdout): `python tests/parity_report.py` on a GPU box.
The committed results of this script are profiles/r1_parity.json and profiles/r2_parity.json."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_parity import _rollout  # noqa: E402


def pct(x):
    return {"median": float(np.median(x)), "p90": float(np.percentile(x, 90)), "p99": float(np.percentile(x, 99)), "max": float(x.max()), "frac_below_1e-4": float((x < 1e-4).mean())}


def main():
    out = {}
    for ctrl in ("joints", "ee"):
        e = _rollout("reach", ctrl, n_envs=256, steps=50, precision="f32", seed=11)
        out[f"reach_{ctrl}_free_running_f32"] = {"envs": 256, "q_err_rad": pct(e["q_env"]), "ee_err_m": pct(e["ee_env"]), "reward_mismatch": e["rew"], "success_mismatch": e["succ"]}
        e = _rollout("reach", ctrl, n_envs=128, steps=50, precision="f32", seed=12, teacher=True)
        out[f"reach_{ctrl}_per_step_f32"] = {"envs": 128, "q_err_rad_max": float(e["q"]), "ee_err_m_max": float(e["ee"]), "qd_err_max": float(e["qd"]), "reward_mismatch": e["rew"], "success_mismatch": e["succ"]}
    for task in ("push", "slide", "pick_and_place", "stack", "flip"):
        e = _rollout(task, "ee", n_envs=128, steps=25, precision="f32", seed=13, teacher=True)
        out[f"{task}_per_step_f32"] = {"envs": 128, "q_err_rad_max": float(e["q"]), "obj_pose_err_max": float(e["obj"]), "obj_pose_err_median_env": float(np.median(e["obj_env"])),
                                       "reward_mismatch": e["rew"], "success_mismatch": e["succ"]}
        e = _rollout(task, "ee", n_envs=128, steps=25, precision="f32", seed=14)
        out[f"{task}_free_running_25_steps_f32"] = {"envs": 128, "q_err_rad": pct(e["q_env"]), "obj_pose_err": pct(e["obj_env"]), "reward_mismatch": e["rew"], "success_mismatch": e["succ"]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
