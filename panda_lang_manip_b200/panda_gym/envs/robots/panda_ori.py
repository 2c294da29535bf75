"""Drop-in for the fork's reference panda_gym/envs/robots/panda_ori.py (class Panda): `set_action(action, euler_xyz)` feeds a target
end-effector orientation (xyz Euler angles in degrees) to the IK instead of the fixed (1,0,0,0) (panda_ori.py:52-99)."""
import numpy as np

from .panda import Panda as _Panda


def quat_from_euler_xyz_deg(euler_xyz) -> np.ndarray:
    """scipy Rotation.from_euler('xyz', e, degrees=True).as_quat(): extrinsic x, then y, then z; (x, y, z, w)."""
    r, p, y = [np.deg2rad(float(v)) / 2 for v in euler_xyz]
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])


class Panda(_Panda):
    def set_action(self, action: np.ndarray, euler_xyz=None) -> None:
        super().set_action(action)
        self.sim._pending_orientation = None if euler_xyz is None else quat_from_euler_xyz_deg(euler_xyz)
