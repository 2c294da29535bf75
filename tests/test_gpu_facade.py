"""The panda_gym drop-in layer on the GPU, following the reference's own tests: test/envs_test.py (random steps on every id),
test/seed_test.py (seeded determinism), test/save_and_restore_test.py (bit-identical snapshot, error on a removed id)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _ids():
    import panda_lang_manip_b200.panda_gym as pg
    return sorted(pg.ENV_IDS)


@pytest.mark.parametrize("env_id", ["PandaReach-v3", "PandaReachJointsDense-v3", "PandaPush-v3", "PandaSlideJoints-v3", "PandaPickAndPlaceDense-v3",
                                    "PandaStack-v3", "PandaFlipJoints-v3"])
def test_random_steps(env_id):          # test/envs_test.py:6-14 (shortened from 1000 to 120 steps per id)
    import panda_lang_manip_b200.panda_gym as pg
    env = pg.make(env_id)
    obs, info = env.reset()
    assert set(obs) == {"observation", "achieved_goal", "desired_goal"} and obs["observation"].dtype == np.float32
    rng = np.random.default_rng(0)
    n_trunc = 0
    for _ in range(120):
        a = rng.uniform(-1, 1, env.action_space.shape).astype(np.float32)
        obs, reward, terminated, truncated, info = env.step(a)
        assert isinstance(reward, float) and isinstance(terminated, bool) and np.all(np.isfinite(obs["observation"]))
        assert info["is_success"] == terminated
        n_trunc += truncated
        if terminated or truncated:
            obs, info = env.reset()
    assert n_trunc >= 1
    env.close()


def test_all_24_ids_construct():
    import panda_lang_manip_b200.panda_gym as pg
    assert len(_ids()) == 24
    for env_id in _ids():
        env = pg.make(env_id)
        o, _ = env.reset(seed=1)
        assert env.observation_space["observation"].shape == o["observation"].shape
        env.close()


@pytest.mark.parametrize("env_id", ["PandaReach-v3", "PandaPush-v3", "PandaSlide-v3", "PandaPickAndPlace-v3", "PandaStack-v3"])
def test_seed_determinism(env_id):      # test/seed_test.py:7-122
    import panda_lang_manip_b200.panda_gym as pg
    env = pg.make(env_id)
    acts = np.random.default_rng(5).uniform(-1, 1, (6,) + env.action_space.shape).astype(np.float32)
    finals = []
    for rep in range(2):
        obs, _ = env.reset(seed=12345)
        for a in acts:
            obs, *_ = env.step(a)
        finals.append(obs)
    for k in finals[0]:
        assert np.allclose(finals[0][k], finals[1][k])
    env.close()


def test_seeded_goal_matches_reference_sampler():
    """reset(seed=k) draws the goal exactly as the reference does (golden from the reference's own _sample_goal)."""
    import os
    import panda_lang_manip_b200.panda_gym as pg
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_rewards.npz"))
    env = pg.make("PandaPickAndPlace-v3")
    for seed in range(4):
        obs, _ = env.reset(seed=seed)
        assert np.array_equal(obs["desired_goal"], gold["pick_and_place_seeded_goals"][seed].astype(np.float32))
        assert np.array_equal(obs["observation"][7:10], gold["pick_and_place_seeded_objects"][seed].astype(np.float32))
    env.close()


def test_save_and_restore():            # test/save_and_restore_test.py:9-36
    import panda_lang_manip_b200.panda_gym as pg
    from panda_lang_manip_b200.panda_gym.pybullet import error
    env = pg.make("PandaPickAndPlace-v3")
    env.reset(seed=3)
    a = np.array([0.3, -0.2, -0.5, 0.7], dtype=np.float32)
    sid = env.save_state()
    o1, *_ = env.step(a)
    env.reset()
    env.restore_state(sid)
    o2, *_ = env.step(a)
    for k in o1:
        assert np.array_equal(o1[k], o2[k])
    env.remove_state(sid)
    with pytest.raises(error):
        env.restore_state(sid)
    env.close()


def test_her_compute_reward_numpy_batch():
    """env.compute_reward(ag[N,G], dg[N,G], info) -- the HER call (core.py:226) -- on numpy batches, bit-exact vs numpy."""
    import panda_lang_manip_b200.panda_gym as pg
    env = pg.make("PandaPush-v3")
    rng = np.random.default_rng(0)
    dg = rng.uniform(-0.2, 0.2, (5000, 3)).astype(np.float32); ag = (dg + rng.normal(0, 0.03, (5000, 3))).astype(np.float32)
    r = env.compute_reward(ag, dg, {})
    want = -np.array(np.linalg.norm(ag - dg, axis=-1) > 0.05, dtype=np.float32)
    assert r.dtype == np.float32 and r.tobytes() == want.tobytes()
    assert np.array_equal(env.task.is_success(ag, dg), np.linalg.norm(ag - dg, axis=-1) < 0.05)
    assert float(env.compute_reward(ag[0], dg[0], {})) == float(want[0])
    env.close()


def test_panda_ori_robot_tilts_the_gripper():
    """The fork's panda_ori.Panda.set_action(action, euler_xyz): commanding a tilted target orientation rotates the hand."""
    import panda_lang_manip_b200.panda_gym as pg
    from panda_lang_manip_b200.panda_gym.envs.robots.panda_ori import Panda as PandaOri
    env = pg.make("PandaPickAndPlace-v3")
    env.robot.__class__ = PandaOri
    env.reset(seed=0)
    q0 = np.array([env.sim.get_joint_angle("panda", j) for j in range(7)])
    for _ in range(15):
        env.robot.set_action(np.zeros(4, np.float32), euler_xyz=[180.0, 0.0, 60.0])
        env.sim.step()
    q1 = np.array([env.sim.get_joint_angle("panda", j) for j in range(7)])
    assert abs(q1[6] - q0[6]) > 0.3           # the wrist joint turned towards the 60 degree yaw target
    env.close()


def test_panda_cartesian_motion_primitives_pick_up_the_cube():
    """The fork's scripted layer (robots/panda_cartesian.py: move / grasp / release): open, move over the cube, descend, grasp, lift."""
    import panda_lang_manip_b200.panda_gym as pg
    from panda_lang_manip_b200.panda_gym.envs.robots.panda_cartesian import Panda as PandaCartesian
    env = pg.make("PandaPickAndPlace-v3")
    env.robot.__class__ = PandaCartesian
    env.reset(seed=2)
    robot, sim = env.robot, env.sim
    cube = sim.get_base_position("object")
    down = [180.0, 0.0, 0.0]
    robot.release()
    robot.move(cube + np.array([0.0, 0.0, 0.10]), down)
    robot.move(cube + np.array([0.0, 0.0, 0.0]), down)
    assert np.linalg.norm(robot.get_ee_position() - cube) < 0.01
    robot.grasp()
    robot.move(cube + np.array([0.0, 0.0, 0.15]), down)
    lifted = sim.get_base_position("object")
    assert lifted[2] > cube[2] + 0.10, (cube, lifted)
    assert robot.get_fingers_width() > 0.03          # the fingers are held open by the 4 cm cube
    env.close()
