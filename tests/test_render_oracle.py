"""The render oracle (oracle/render_oracle.py) against closed-form answers and against the reference's own deprojection arithmetic
(reference panda_gym/pybullet.py:205-241 restated with explicit 4x4 matrices) -- CPU only."""
import numpy as np

from oracle import render_oracle as ro

TABLE = ("box", 2, [-0.3, 0.0, -0.2], np.eye(3), [0.55, 0.35, 0.2])


def test_centre_ray_hits_the_table_at_the_camera_distance():
    # the camera looks at the target (0,0,0), which lies on the table top: the two centre-adjacent pixels see depth ~ distance
    d, seg, pts, valid = ro.render([TABLE], 64, 64, distance=1.4)
    ze = lambda db: 2 * 100.0 * 0.1 / ((100.0 + 0.1) - (2 * db - 1) * (100.0 - 0.1))
    assert abs(ze(d[32, 32]) - 1.4) < 0.03 and seg[32, 32] == 2
    assert d[0, 0] == 1.0 and seg[0, 0] == 0 and not valid[0, 0]          # top-left corner: sky


def test_closed_form_deprojection_is_the_reference_arithmetic():
    d, seg, pts, valid = ro.render([TABLE, ("box", 3, [0.0, 0.1, 0.02], np.eye(3), [0.02, 0.02, 0.02])], 96, 80, distance=1.0, yaw=30, pitch=-40, crop=False)
    ref = ro.deproject_reference(d, 96, 80, distance=1.0, yaw=30, pitch=-40)
    assert np.allclose(pts[valid], ref[valid], atol=1e-9)
    # hit points of the table top lie on z = 0 up to the reference's half-pixel bias (it deprojects pixel corners, the image samples centres)
    top = valid & (seg == 2) & (pts[..., 2] > -0.012)                        # the table's top face (its sides are visible too, below it)
    assert top.sum() > 1500 and np.abs(pts[top][:, 2]).max() < 0.012         # 96 x 80 pixels at 1 m: half a pixel is ~1 cm on the slanted table
    assert (seg == 3).sum() > 10                                             # the cube is visible


def test_cylinder_and_rotated_box():
    c, s = np.cos(0.5), np.sin(0.5)
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    d, seg, pts, valid = ro.render([TABLE, ("cyl", 3, [0.0, 0.0, 0.015], np.eye(3), [0.03, 0.03, 0.015]), ("box", 4, [0.1, 0.1, 0.02], Rz, [0.02, 0.02, 0.02])], 128, 128, distance=0.6)
    assert (seg == 3).sum() > 20 and (seg == 4).sum() > 20
    puck = pts[(seg == 3) & valid]
    assert np.all(np.linalg.norm(puck[:, :2], axis=1) < 0.03 + 0.01) and puck[:, 2].max() < 0.03 + 0.01
