"""The math behind the bare world (pg_sim_step with raw joint motors), pg_get_link_state and pg_inverse_kinematics_link -- the
kernel's own device code compiled for the host (tests/hostcheck) -- against the reference's known-answer tests
(reference test/pybullet_test.py:56-65,123-204,254-266, atol 1e-3 as there) and against the oracle, without a GPU.  The GPU versions of
the same KATs, through the C ABI and the facade, are tests/test_gpu_reference_kat.py."""
import numpy as np
import pytest

from tests.oracle_util import OracleSim, P

ORIGIN = np.zeros(3)
VEL_MOTORS = np.tile(np.array([0.0, 1.0, 0.0, 0.0, 500.0]), (9, 1))      # loadURDF defaults: velocity motors, target 0, max impulse 1 / sub-step


def _joint5(hc, dbl):
    st = np.zeros(18)
    mot = VEL_MOTORS.copy(); mot[5] = [0.1, 1.0, 0.3, 0.0, 5.0]             # control_joints("panda", [5], [0.3], [5.0])
    hc.hc_bare_steps(dbl, 0, P(ORIGIN), P(st), P(np.ascontiguousarray(mot)), None, None, None, 20)
    q, qd = st[:9].copy(), st[9:18].copy()
    out = np.zeros(13)
    hc.hc_link_state(dbl, P(ORIGIN), 5, P(q), P(qd), P(q - qd / 500.0), P(out))
    return q, qd, out


@pytest.mark.parametrize("dbl", [1, 0])
def test_reference_kats_on_the_kernel_math(hostcheck, dbl):
    out = np.zeros(13); z = np.zeros(9)
    hostcheck.hc_link_state(dbl, P(ORIGIN), 1, P(z), P(z), P(z), P(out))
    assert np.allclose(out[:3], [0.000, 0.060, 0.373], atol=1e-3)                          # test/pybullet_test.py:135
    q, qd, ls = _joint5(hostcheck, dbl)
    assert np.allclose(ls[3:7], [0.707, -0.02, 0.02, 0.707], atol=1e-3)                    # :152 (stale link-transform cache)
    assert np.allclose(ls[7:10], [-0.0068, 0.0000, 0.1186], atol=1e-3)                     # :169
    assert np.allclose(ls[10:13], [0.000, -2.969, 0.000], atol=1e-3)                       # :186
    assert np.allclose(q[5], 0.063, atol=1e-3)                                             # :203
    ik = np.zeros(9)
    hostcheck.hc_ik_link(dbl, P(ORIGIN), 6, P(z), P(np.array([0.4, 0.5, 0.6])), P(np.array([0.707, -0.02, 0.02, 0.707])), P(ik))
    assert np.allclose(ik, [1.000, 1.223, -1.113, -0.021, -0.917, 0.666, -0.499, 0.0, 0.0], atol=1e-3)   # :265
    # free fall of a 1 kg unit box, no plane: :57-65
    st = np.zeros(31); st[18 + 6] = 1.0
    body = np.array([0.0, 0.5, 0.5, 0.5, 1.0, 0.5])
    hostcheck.hc_bare_steps(dbl, 1, P(np.array([0.0, 0.0, 1000.0])), P(st), P(np.ascontiguousarray(VEL_MOTORS)), P(body), None, None, 20)
    assert np.allclose(st[18 + 7:18 + 10], [0.0, 0.0, -0.392], atol=1e-3)


def test_link_state_all_links_vs_oracle(hostcheck):
    """Every link 0..11, random joint state and velocity: CoM position, orientation and both velocities equal the oracle's getLinkState
    (which takes the pose from the stale cache and the velocity from the fresh state)."""
    rng = np.random.default_rng(0)
    lo = np.array([-2.9, -1.8, -2.9, -3.0, -2.9, 0.0, -2.9, 0.0, 0.0]); hi = np.array([2.9, 1.8, 2.9, -0.1, 2.9, 3.7, 2.9, 0.04, 0.04])
    for trial in range(5):
        q, qd = rng.uniform(lo, hi), rng.uniform(-1, 1, 9) * np.array([1, 1, 1, 1, 1, 1, 1, 0.05, 0.05])
        s = OracleSim(base=(0.1, -0.2, 0.3))
        from tests.oracle_util import load_oracle
        lib = load_oracle()
        # oracle state: q, qd fresh; cache qc = q - qd dt
        import ctypes
        for d, link in enumerate([0, 1, 2, 3, 4, 5, 6, 9, 10]):
            s.reset_joint(link, q[d])
        lib.po_set_joint_state(s.h, P(q), P(qd), P(q - qd / 500.0))
        for link in range(12):
            out = np.zeros(13)
            hostcheck.hc_link_state(1, P(np.array([0.1, -0.2, 0.3])), link, P(q), P(qd), P(q - qd / 500.0), P(out))
            p, qt, v, w = s.link_state(link)
            sign = 1.0 if np.dot(qt, out[3:7]) >= 0 else -1.0
            assert np.allclose(out[:3], p, atol=1e-12) and np.allclose(sign * out[3:7], qt, atol=1e-12), (link, out[:7], p, qt)
            assert np.allclose(out[7:10], v, atol=1e-12) and np.allclose(out[10:13], w, atol=1e-12), (link, out[7:], v, w)
        s.close()


@pytest.mark.parametrize("link", [3, 6, 8, 9, 10, 11])
def test_ik_any_link_vs_oracle(hostcheck, link):
    rng = np.random.default_rng(link)
    for trial in range(4):
        q = np.array([0.0, 0.41, 0.0, -1.85, 0.0, 2.26, 0.79, 0.01, 0.02]) + rng.uniform(-0.3, 0.3, 9) * np.array([1, 1, 1, 1, 1, 1, 1, 0.02, 0.02])
        s = OracleSim()
        for d, l in enumerate([0, 1, 2, 3, 4, 5, 6, 9, 10]):
            s.reset_joint(l, q[d])
        p0, q0, _, _ = s.link_state(link)
        tgt = np.array([0.45, 0.1, 0.45]) + rng.uniform(-0.1, 0.1, 3)
        quat = rng.normal(size=4)
        want = s.ik(link, tgt, quat)
        got = np.zeros(9)
        hostcheck.hc_ik_link(1, P(ORIGIN), link, P(q), P(tgt), P(quat), P(got))
        assert np.allclose(got, want, atol=1e-9), (link, got, want)
        s.close()


def test_generic_motors_vs_oracle(hostcheck):
    """Mixed motors (some joints position-controlled with different forces, others on the default velocity motors) for 3 x 20 sub-steps."""
    rng = np.random.default_rng(3)
    s = OracleSim()
    q0 = np.array([0.1, 0.3, -0.2, -1.5, 0.2, 1.9, 0.5, 0.01, 0.01])
    for d, l in enumerate([0, 1, 2, 3, 4, 5, 6, 9, 10]):
        s.reset_joint(l, q0[d])
    st = np.zeros(18); st[:9] = q0
    mot = VEL_MOTORS.copy()
    for j, (d, l) in enumerate([(1, 1), (3, 3), (5, 5), (7, 9)]):
        tq, f = q0[d] + rng.uniform(-0.2, 0.2) * (0.05 if d >= 7 else 1), [87.0, 40.0, 12.0, 20.0][j]
        mot[d] = [0.1, 1.0, tq, 0.0, f]
        s.control_joint(l, tq, f)
    for rep in range(3):
        s.step(20)
        hostcheck.hc_bare_steps(1, 0, P(ORIGIN), P(st), P(np.ascontiguousarray(mot)), None, None, None, 20)
        qo = np.array([s.joint(l)[0] for l in [0, 1, 2, 3, 4, 5, 6, 9, 10]])
        assert np.allclose(st[:9], qo, atol=1e-7), (rep, st[:9] - qo)
    s.close()
