"""The oracle against the reference's own known-answer tests (reference test/pybullet_test.py, atol 1e-3 as there)."""
import numpy as np

from tests.oracle_util import OracleSim


def test_dt():                      # test/pybullet_test.py:34  -- 20 sub-steps of 1/500 s
    assert np.isclose(20 * (1.0 / 500.0), 0.04)


def test_free_fall_velocity():      # test/pybullet_test.py:57-65
    s = OracleSim()
    o = s.add_box([0.5, 0.5, 0.5], 1.0, [0.0, 0.0, 0.0])     # as the reference: at the origin; a bare world has no plane to land on
    s.step()
    v, _ = s.base_velocity(o)
    assert np.allclose(v, [0.0, 0.0, -0.392], atol=1e-3)
    s.close()


def test_get_link_position():       # :123-136
    s = OracleSim()
    assert np.allclose(s.link_state(1)[0], [0.000, 0.060, 0.373], atol=1e-3)
    s.close()


def _joint5_motion():
    s = OracleSim()
    s.control_joint(5, 0.3, 5.0)
    s.step()
    return s


def test_get_link_orientation():    # :139-153 -- passes only with the stale link-transform cache
    s = _joint5_motion()
    assert np.allclose(s.link_state(5)[1], [0.707, -0.02, 0.02, 0.707], atol=1e-3)
    s.close()


def test_get_link_velocity():       # :156-170
    s = _joint5_motion()
    assert np.allclose(s.link_state(5)[2], [-0.0068, 0.0000, 0.1186], atol=1e-3)
    s.close()


def test_get_link_angular_velocity():   # :173-187
    s = _joint5_motion()
    assert np.allclose(s.link_state(5)[3], [0.000, -2.969, 0.000], atol=1e-3)
    s.close()


def test_get_joint_angle():         # :190-204
    s = _joint5_motion()
    assert np.allclose(s.joint(5)[0], 0.063, atol=1e-3)
    s.close()


def test_set_base_pose():           # :207-218
    s = OracleSim()
    o = s.add_box([0.5, 0.5, 0.5], 1.0, [0.0, 0.0, 0.0])
    s.set_base_pose(o, [1.0, 1.0, 1.0], [0.707, -0.02, 0.02, 0.707])
    p, q = s.base_pose(o)
    assert np.allclose(p, [1.0, 1.0, 1.0], atol=1e-3) and np.allclose(q, [0.707, -0.02, 0.02, 0.707], atol=1e-3)
    s.close()


def test_set_joint_angles():        # :221-251
    s = OracleSim()
    s.reset_joint(3, 0.4); s.reset_joint(4, 0.5)
    assert np.isclose(s.joint(3)[0], 0.4, atol=1e-3) and np.isclose(s.joint(4)[0], 0.5, atol=1e-3)
    s.close()


def test_inverse_kinematics():      # :254-266
    s = OracleSim()
    q = s.ik(6, [0.4, 0.5, 0.6], [0.707, -0.02, 0.02, 0.707])
    assert np.allclose(q, [1.000, 1.223, -1.113, -0.021, -0.917, 0.666, -0.499, 0.0, 0.0], atol=1e-3)
    s.close()


def test_neutral_ee_pose():
    """Upstream panda-gym's documented Reach reset observation: EE at (0.0384, 0, 0.1974)."""
    from tests.oracle_util import OracleEnv
    e = OracleEnv("reach")
    obs, ag, dg = e.reset([0.1, 0.0, 0.1])
    assert np.allclose(obs[:3], [0.03844, 0.0, 0.19740], atol=2e-4) and np.allclose(obs[3:], 0.0)
    e.close()
