"""The kernel's formulation (csrc/*.cuh compiled for the host by tests/hostcheck, fp64 and fp32) against the oracle, without a GPU:
RNEA + CRBA + Cholesky vs ABA + impulse responses, operational-space contact rows vs generalized rows, fast (watched-limit) sweeps
vs full sweeps.  fp64 agreement is ~1e-12 until the first discrete solver event; fp32 is held to the GPU tests' tolerances."""
import ctypes

import numpy as np
import pytest

from tests.oracle_util import GOAL_DIM, NOBJ, OBS_DIM, OracleEnv, P

BASE = np.array([-0.6, 0.0, 0.0])
NEUTRAL = np.array([0.0, 0.41, 0.0, -1.85, 0.0, 2.26, 0.79, 0.0, 0.0])
TASK_ID = {"reach": 0, "push": 1, "slide": 2, "pick_and_place": 3, "stack": 4, "flip": 5}


def _rollout(hc, task, control, seed, steps, dbl, action_fn=None, full_obs=False, obj0_xy=None):
    rng = np.random.default_rng(seed)
    oe = OracleEnv(task, control)
    goal = np.zeros(6); goal[:3] = rng.uniform([-0.15, -0.15, 0.0], [0.15, 0.15, 0.2])
    if task == "stack":
        goal[3:] = goal[:3] + [0, 0, 0.04]
    if task == "flip":
        goal[:4] = [0, 0, 0, 1]
    objpos = np.array([rng.uniform(-0.15, 0.15), rng.uniform(-0.15, 0.15), 0.03 if task == "slide" else 0.02, rng.uniform(-0.15, 0.15), rng.uniform(-0.15, 0.15), 0.06])
    if obj0_xy is not None:
        objpos[:2] = obj0_xy
    obs, ag, dg = oe.reset(goal, objpos)
    st = np.zeros(50); st[:9] = NEUTRAL
    for o in range(NOBJ[task]):
        st[18 + 13 * o:18 + 13 * o + 3] = objpos[3 * o:3 * o + 3]; st[18 + 13 * o + 6] = 1
    st[44:] = goal
    worst_q = worst_obs = 0.0
    for t in range(steps):
        a = (action_fn(t, oe.action_dim) if action_fn else rng.uniform(-1, 1, oe.action_dim)).astype(np.float32)
        obs, ag, dg, r, s = oe.step(a)
        o2, a2, d2 = np.zeros(OBS_DIM[task], np.float32), np.zeros(GOAL_DIM[task], np.float32), np.zeros(GOAL_DIM[task], np.float32)
        r2, s2 = np.zeros(1, np.float32), np.zeros(1, np.uint8)
        hc.hc_env_step(dbl, TASK_ID[task], 0 if control == "ee" else 1, 0, P(BASE), P(st), P(a), P(o2), P(a2), P(d2), P(r2), P(s2))
        q, qd = oe.joints()
        worst_q = max(worst_q, np.abs(st[:9] - q).max()); worst_obs = max(worst_obs, np.abs(o2[:3] - obs[:3]).max())
        if full_obs:        # the whole observation: gripper, object poses and velocities
            worst_obs = max(worst_obs, np.abs(o2 - obs).max())
        assert np.float32(r2[0]).tobytes() == np.float32(r).tobytes() or abs(float(r2[0]) - r) < 1e-6
    oe.close()
    return worst_q, worst_obs


@pytest.mark.parametrize("task,control", [("reach", "joints"), ("reach", "ee"), ("push", "ee"), ("pick_and_place", "joints"), ("stack", "ee"), ("slide", "ee")])
def test_kernel_math_fp64_matches_oracle(hostcheck, task, control):
    for seed in range(2):
        wq, wo = _rollout(hostcheck, task, control, seed, 30, 1)
        assert wq < 1e-4 and wo < 1e-4, (task, control, seed, wq, wo)


def test_kernel_math_fp32_matches_oracle(hostcheck):
    wq, wo = _rollout(hostcheck, "reach", "ee", 0, 30, 0)
    assert wq < 2e-4 and wo < 2e-4


@pytest.mark.parametrize("task", ["push", "pick_and_place", "stack"])
def test_contact_rows_fp32_gripper_pressed_down(hostcheck, task):
    """The contact paths of the kernel math in fp32 (typed passes, operational-space rows, generic rows on the contact-point velocity)
    with the gripper driven into the table / the objects: full observation (object poses and velocities included) against the fp64
    oracle over 20 free-running steps."""
    def act(t, n):
        a = np.zeros(n); a[2] = -1.0; a[0] = 0.3 * np.sin(0.7 * t)
        if n > 3:
            a[3] = -1.0 if t > 6 else 1.0
        return a
    for seed in range(2):
        wq, wo = _rollout(hostcheck, task, "ee", seed, 20, 0, action_fn=act, full_obs=True, obj0_xy=(0.04 + 0.01 * seed, 0.0))   # the object sits under the gripper
        assert wq < 5e-4 and wo < 5e-3, (task, seed, wq, wo)


def test_joint_limit_engages_and_fast_sweep_falls_back(hostcheck):
    """Constant +1 on joint 1 drives it into its upper limit (1.8326): the watched row must trigger the full sweep and the
    result must still match the oracle, which always runs every row."""
    def act(t, n):
        a = np.zeros(n); a[1] = 1.0; a[3] = 1.0
        return a
    wq, wo = _rollout(hostcheck, "reach", "joints", 0, 50, 1, action_fn=act)
    assert wq < 1e-6, wq
    oe = OracleEnv("reach", "joints"); oe.reset(np.zeros(3))
    for t in range(50):
        oe.step(act(t, 7).astype(np.float32))
    q, _ = oe.joints()
    oe.close()
    assert q[1] > 1.8 and q[1] < 1.8326 + 5e-3 and q[3] > -0.02 and q[3] < 5e-3      # both joints are resting on their limits


def test_mass_matrix_inverse(hostcheck, oracle):
    rng = np.random.default_rng(0)
    oe = OracleEnv("reach", "joints")
    for _ in range(3):
        q = NEUTRAL + rng.uniform(-0.3, 0.3, 9); q[7:] = rng.uniform(0, 0.04, 2); qd = rng.uniform(-1, 1, 9)
        oe.set_joints(q, qd)
        mo = np.zeros(81); oracle.po_mass_matrix(oe.sim, P(mo))
        mh, qdd = np.zeros(81), np.zeros(9)
        hostcheck.hc_minv(1, P(BASE), P(np.ascontiguousarray(q)), P(np.ascontiguousarray(qd)), P(mh), P(qdd))
        assert np.abs(mh - mo).max() / np.abs(mo).max() < 1e-12
    oe.close()


@pytest.mark.parametrize("deg,off", [(0.0, 0.0), (3.0, 0.0005), (20.0, 0.002), (45.0, 0.0), (12.0, 0.015)])
def test_box_on_box_contacts_match_the_oracle(hostcheck, deg, off):
    """Stack with the second cube lying on the first, turned / shifted: the kernel math's box-box contacts (vertices against the reference
    face, edge against edge) against the oracle's, fp64, 10 env steps of a resting stack -- and the stack stays up."""
    from tests.contact_kats import state_row
    ang = np.radians(deg)
    row = state_row("stack", [([0.1, 0.05, 0.02], [0, 0, 0]), ([0.1 + off, 0.05 - off / 2, 0.06], [0, 0, 0], [0, 0, np.sin(ang / 2), np.cos(ang / 2)])],
                    [0.1, 0.05, 0.02, 0.1, 0.05, 0.06])
    oe = OracleEnv("stack", "joints"); oe.set_full_state(row)
    st = np.zeros(50); st[:44] = row[:44]; st[44:] = row[44:50]
    a = np.zeros(8, np.float32)
    for t in range(10):
        oe.step(a)
        o2, a2, d2 = np.zeros(OBS_DIM["stack"], np.float32), np.zeros(6, np.float32), np.zeros(6, np.float32)
        r2, s2 = np.zeros(1, np.float32), np.zeros(1, np.uint8)
        hostcheck.hc_env_step(1, TASK_ID["stack"], 1, 0, P(BASE), P(st), P(a), P(o2), P(a2), P(d2), P(r2), P(s2))
        ref = oe.full_state()
        assert np.abs(st[:44] - ref[:44]).max() < 1e-7, (t, np.abs(st[:44] - ref[:44]).max())
    oe.close()
    assert abs(st[33] - 0.06) < 3e-4 and abs(st[20] - 0.02) < 3e-4
