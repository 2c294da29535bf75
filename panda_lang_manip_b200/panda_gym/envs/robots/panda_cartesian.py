"""Drop-in for the fork's reference panda_gym/envs/robots/panda_cartesian.py (class Panda): unscaled ee / finger actions
(:67,:157), an EE orientation target, and the blocking motion primitives `move` (:98-122: 15 mm way-points + Slerp), `grasp` /
`release` (:124-145: 30 steps).  Each primitive step is one fused kernel launch on the facade's batch of one.

Differences, stated: `get_ee_orientation` reads the pose from the current joint state (the reference reads pybullet's one-sub-step
stale link cache); `grasp` closes by commanding -1 on the finger action instead of flipping `block_gripper` (both saturate the
finger motors against the object)."""
import numpy as np
from scipy.spatial.transform import Rotation as R
from scipy.spatial.transform import Slerp

from .panda_ori import Panda as _PandaOri


class Panda(_PandaOri):
    def _ensure_scale(self) -> None:
        if not getattr(self, "_scaled", False):
            self.sim._require().set_action_scale(1.0, 1.0)
            self._scaled = True

    def set_action(self, action: np.ndarray, euler_xyz=None) -> None:
        self._ensure_scale()
        super().set_action(np.clip(np.asarray(action, dtype=np.float32), -1.0, 1.0), euler_xyz)

    def get_ee_position(self) -> np.ndarray:
        return self.sim._require().ee_pose()[0, :3].cpu().numpy()

    def get_ee_orientation(self) -> np.ndarray:
        return R.from_quat(self.sim._require().ee_pose()[0, 3:].cpu().numpy()).as_euler("xyz", degrees=True)

    @staticmethod
    def get_waypoint(start_pt, target_pt, max_delta, num_steps=None):
        total_delta = target_pt - start_pt
        if num_steps is None:
            num_steps = np.linalg.norm(total_delta) // max_delta
            if np.linalg.norm(total_delta) % max_delta > 1e-3:
                num_steps += 1
        num_steps = max(int(num_steps), 1)
        delta = total_delta / num_steps
        return (lambda i: start_pt + delta * min(i, num_steps)), num_steps

    @staticmethod
    def get_ori(initial_euler, final_euler, num_steps):
        slerp = Slerp([1, max(num_steps, 2)], R.from_euler("xyz", [np.array(initial_euler, dtype=float), np.array(final_euler, dtype=float)], degrees=True))
        return lambda i: slerp(min(max(i, 1), max(num_steps, 2))).as_euler("xyz", degrees=True)

    def move(self, goal_pos, goal_euler) -> None:
        goal_pos = np.asarray(goal_pos, dtype=np.float64)
        start_pos, start_euler = self.get_ee_position(), self.get_ee_orientation()
        if np.linalg.norm(goal_pos - start_pos) < 0.03:
            gen_fn, num_steps = self.get_waypoint(start_pos, goal_pos, 0.015, num_steps=20)
        else:
            gen_fn, num_steps = self.get_waypoint(start_pos, goal_pos, 0.015)
        gen_ori_fn = self.get_ori(start_euler, goal_euler, num_steps)
        for i in range(1, num_steps + 1):
            action = np.concatenate([gen_fn(i) - self.get_ee_position(), [0.0]])      # hold the fingers where they are
            self.set_action(action, gen_ori_fn(i))
            self.sim.step()

    def grasp(self) -> None:
        for _ in range(30):
            self.set_action(np.array([0.0, 0.0, 0.0, -1.0]), self.get_ee_orientation())
            self.sim.step()

    def release(self, width: float = 1.0) -> None:
        for _ in range(30):
            self.set_action(np.array([0.0, 0.0, 0.0, width]), self.get_ee_orientation())
            self.sim.step()
