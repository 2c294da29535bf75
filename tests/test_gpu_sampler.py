"""The DEVICE sampler (Philox; what every auto-reset / bench run uses) against the reference's distributions, task by task: ranges,
uniformity (KS against the host sampler that reproduces the reference's draws, panda_gym/sampling.py, golden-tested against the
reference's own _sample_goal / _sample_object), PickAndPlace's 30 % on-table goals (pick_and_place.py:74-76), Stack's shared goal
noise and independent cube placements (stack.py:94-119), Flip's uniform rotations; task kwargs (ranges) as kernel parameters; seeded
resets (test/seed_test.py semantics)."""
import numpy as np
import pytest
from scipy import stats

from tests.oracle_util import GOAL_DIM, NOBJ

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
N = 65536


def _draws(task, **kw):
    import panda_lang_manip_b200 as p
    env = p.PandaVecEnv(task, N, seed=123, **kw)
    st = env.get_state().cpu().numpy()
    env.close()
    nobj, G = NOBJ[task], GOAL_DIM[task]
    return st[:, 18 + 13 * nobj:18 + 13 * nobj + G], [st[:, 18 + 13 * o:18 + 13 * o + 7] for o in range(nobj)], st


def _host(task, n=20000):
    from panda_lang_manip_b200.panda_gym.sampling import sample_reset
    rng = np.random.default_rng(0)
    gs, os_ = [], []
    for _ in range(n):
        g, o = sample_reset(task, rng)
        gs.append(g); os_.append(np.concatenate(o) if o else np.zeros(0))
    return np.array(gs), np.array(os_)


def _same_dist(a, b):
    # rounded to 1e-6: the device state is float32, so an atom of the distribution (PickAndPlace's 30 % goals at exactly z = 0.02)
    # would otherwise sit an ulp away from the host's float64 atom and register as a 0.3 jump in the KS statistic
    return stats.ks_2samp(np.round(a, 6), np.round(b, 6)).pvalue > 1e-3


@pytest.mark.parametrize("task", ["reach", "push", "slide", "pick_and_place", "stack", "flip"])
def test_device_sampler_distribution(task):
    goal, objs, st = _draws(task)
    hg, ho = _host(task)
    assert np.all(st[:, :9] == np.float32([0.0, 0.41, 0.0, -1.85, 0.0, 2.26, 0.79, 0.0, 0.0]).astype(np.float64)) and np.all(st[:, 9:18] == 0)     # neutral pose, at rest
    if task == "flip":
        assert np.allclose(np.linalg.norm(goal, axis=1), 1.0, atol=1e-6)
        for k in range(4):                               # uniform on S^3: every component has the same marginal as the host's normalised Gaussians
            assert _same_dist(goal[:, k], hg[:, k]), k
    else:
        lo, hi = hg.min(0), hg.max(0)
        for k in range(goal.shape[1]):
            assert goal[:, k].min() >= lo[k] - 1e-3 and goal[:, k].max() <= hi[k] + 1e-3, (k, goal[:, k].min(), goal[:, k].max())
            if hi[k] - lo[k] > 1e-6:
                assert _same_dist(goal[:, k], hg[:, k]), k
            else:
                assert np.allclose(goal[:, k], hg[0, k], atol=1e-6), k
    if task == "pick_and_place":
        on_table = np.isclose(goal[:, 2], 0.02, atol=1e-7).mean()
        assert abs(on_table - 0.3) < 0.01, on_table      # P(goal on the table) = 0.3
        assert abs(np.isclose(hg[:, 2], 0.02).mean() - 0.3) < 0.015
    if task == "stack":
        assert np.allclose(goal[:, 3:5], goal[:, 0:2], atol=1e-7) and np.allclose(goal[:, 5], 0.06, atol=1e-6)     # both goals share the noise
        c = np.corrcoef(objs[0][:, 0], objs[1][:, 0])[0, 1]
        assert abs(c) < 0.02, c                          # the two cubes are placed independently
        assert np.allclose(objs[1][:, 2], 0.06, atol=1e-6)
    for o, ob in enumerate(objs):
        for k in range(2):
            assert ob[:, k].min() >= -0.15 - 1e-6 and ob[:, k].max() <= 0.15 + 1e-6
            assert _same_dist(ob[:, k], ho[:, 3 * o + k]), (o, k)
        assert np.allclose(ob[:, 3:7], [0, 0, 0, 1])
        assert abs(np.corrcoef(ob[:, 0], ob[:, 1])[0, 1]) < 0.02 and abs(np.corrcoef(ob[:, 0], goal[:, 0])[0, 1]) < 0.02


def test_task_kwargs_are_kernel_parameters():
    """tasks/reach.py:15-23 (distance_threshold, goal_range), push.py:12-25 (goal_xy_range, obj_xy_range), PyBullet(n_substeps)."""
    import panda_lang_manip_b200 as p
    from panda_lang_manip_b200.panda_gym.sampling import default_ranges
    glo, ghi, _, _ = default_ranges("reach", goal_range=0.1)
    goal, _, _ = _draws("reach", goal_range_low=glo, goal_range_high=ghi)
    assert goal[:, 0].min() >= -0.05 - 1e-6 and goal[:, 0].max() <= 0.05 + 1e-6 and goal[:, 2].max() <= 0.1 + 1e-6 and goal[:, 2].max() > 0.095
    glo, ghi, olo, ohi = default_ranges("push", goal_xy_range=0.2, obj_xy_range=0.1)
    goal, objs, _ = _draws("push", goal_range_low=glo, goal_range_high=ghi, obj_range_low=olo[:2], obj_range_high=ohi[:2])
    assert abs(goal[:, 1]).max() <= 0.1 + 1e-6 and abs(goal[:, 1]).max() > 0.095 and abs(objs[0][:, 0]).max() <= 0.05 + 1e-6 and abs(objs[0][:, 0]).max() > 0.045
    # threshold: the in-step success / sparse reward follow it, and so do compute_reward / is_success of the env
    n = 2048
    a = p.PandaVecEnv("reach", n, control_type="joints", seed=1, auto_reset=False)
    b = p.PandaVecEnv("reach", n, control_type="joints", seed=1, auto_reset=False, distance_threshold=0.12)
    act = torch.zeros((n, 7), device="cuda")
    oa, ra, ta, _, _ = a.step(act); ob, rb, tb, _, _ = b.step(act)
    d = torch.linalg.norm(oa["achieved_goal"] - oa["desired_goal"], dim=-1)
    assert torch.equal(oa["observation"], ob["observation"])
    assert torch.equal(ta.bool(), d < 0.05) and torch.equal(tb.bool(), d < np.float32(0.12)) and int(tb.sum()) > int(ta.sum())
    assert torch.equal(rb, -(d > np.float32(0.12)).float())
    assert torch.equal(b.is_success(oa["achieved_goal"], oa["desired_goal"]), d < np.float32(0.12))
    for e in (a, b):
        e.close()


def test_n_substeps_and_threshold_against_oracle():
    """PyBullet(n_substeps=10) + distance_threshold=0.08 on PandaPush (ee control): per-step parity with the oracle run with the same
    parameters (the stale link-cache sub-step, the segment cut and the in-step reward all follow the parameter)."""
    import panda_lang_manip_b200 as p
    from tests.oracle_util import OracleEnv
    n, steps = 8, 20
    rng = np.random.default_rng(2)
    goals = np.stack([rng.uniform(-0.15, 0.15, n), rng.uniform(-0.15, 0.15, n), np.full(n, 0.02)], -1)
    objs = np.stack([rng.uniform(-0.15, 0.15, n), rng.uniform(-0.15, 0.15, n), np.full(n, 0.02)], -1)
    env = p.PandaVecEnv("push", n, control_type="ee", auto_reset=False, n_substeps=10, distance_threshold=0.08)
    env.reset(goals=goals, object_positions=objs)
    ors = [OracleEnv("push", "ee", n_substeps=10, distance_threshold=0.08) for _ in range(n)]
    for i, oe in enumerate(ors):
        oe.reset(goals[i], objs[i])
    worst = 0.0
    for t in range(steps):
        a = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
        obs, rew, term, _, _ = env.step(torch.from_numpy(a).cuda())
        st = env.get_state().cpu().numpy(); og = obs["observation"].cpu().numpy()
        for i, oe in enumerate(ors):
            ob, ag, dg, r, s = oe.step(a[i])
            full = oe.full_state()
            worst = max(worst, np.abs(st[i, :9] - full[:9]).max(), np.abs(og[i, :3] - ob[:3]).max())
            d = float(np.linalg.norm(ag.astype(np.float64) - dg.astype(np.float64)))
            if abs(d - 0.08) > 1e-3:
                assert float(rew[i]) == r and bool(term[i]) == s
            st[i, :-1] = full
        env.set_state(torch.from_numpy(st))
    assert worst < 1e-4, worst
    env.close()
    for oe in ors:
        oe.close()


def test_n_substeps_composes():
    """Two env steps of 10 sub-steps under joints control with zero actions track the same motor targets only at the first step;
    exactness check instead: a bare world stepped 20 sub-steps at once equals the same world stepped 2 x 10 (bit-identical)."""
    from panda_lang_manip_b200.bare_world import PandaBareWorld
    w1, w2 = PandaBareWorld(4), PandaBareWorld(4)
    for w in (w1, w2):
        w.control_joints([1, 3, 5], [0.3, -1.2, 1.0], [87.0, 87.0, 12.0])
    w1.step(20); w2.step(10); w2.step(10)
    assert torch.equal(w1.get_state()[:, :18], w2.get_state()[:, :18])
    w1.close(); w2.close()


def test_seeded_reset_is_keyed_by_the_seed_alone():
    import panda_lang_manip_b200 as p
    n = 1024
    env = p.PandaVecEnv("pick_and_place", n, seed=9)
    seeds = np.arange(n) % 7                               # only 7 distinct seeds
    env.reset(seeds=seeds)
    st = env.get_state().cpu().numpy()
    for k in range(7):
        rows = st[seeds == k]
        assert np.all(rows == rows[0]), k                   # equal seeds -> equal goal and cube placement, whatever the env index
    assert len({tuple(st[k, 18:34]) for k in range(7)}) == 7
    other = p.PandaVecEnv("pick_and_place", 7, seed=1234, env_id_offset=555)
    other.reset(seeds=np.arange(7))
    assert np.array_equal(other.get_state().cpu().numpy(), st[:7])      # ... and whatever the handle
    env.reset(seeds=seeds)
    assert np.array_equal(env.get_state().cpu().numpy(), st)           # same seed twice -> same reset (test/seed_test.py)
    env.close(); other.close()
