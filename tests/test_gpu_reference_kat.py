"""The reference's own known-answer tests, reference test/pybullet_test.py (29 tests, atol 1e-3 unless noted), run against the CUDA
kernels through the drop-in sim facade ``panda_gym.pybullet.PyBullet`` used exactly as the reference uses it: ``PyBullet()`` +
``loadURDF`` / ``create_box`` + ``control_joints`` + ``step`` + getters.  Each test names the reference lines it ports.  Both the fp32
product kernels and the fp64 debug instantiation are held to the reference's tolerance."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(params=["f32", "f64"])
def pybullet(request):
    from panda_lang_manip_b200.panda_gym.pybullet import PyBullet
    sim = PyBullet()
    sim._precision = request.param
    yield sim
    sim.close()


def _load_panda(sim):
    sim.loadURDF(body_name="panda", fileName="franka_panda/panda.urdf", basePosition=[0.0, 0.0, 0.0], useFixedBase=True)


def _box(sim):
    sim.create_box("my_box", [0.5, 0.5, 0.5], 1.0, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 1.0])


def test_construct_step_close_dt(pybullet):                 # :9-35
    pybullet.step()
    assert pybullet.dt == 0.04


def test_get_base_position(pybullet):                       # :45-53 (atol 1e-7)
    _box(pybullet)
    assert np.allclose(pybullet.get_base_position("my_box"), np.zeros(3), atol=1e-7)


def test_get_base_velocity(pybullet):                       # :56-65  free fall over one 20-sub-step step
    _box(pybullet)
    pybullet.step()
    assert np.allclose(pybullet.get_base_velocity("my_box"), [0.0, 0.0, -0.392], atol=1e-3)


def test_get_base_orientation_rotation_angular_velocity(pybullet):   # :68-98
    _box(pybullet)
    assert np.allclose(pybullet.get_base_orientation("my_box"), [0.0, 0.0, 0.0, 1.0], atol=1e-3)
    assert np.allclose(pybullet.get_base_rotation("my_box"), [0.0, 0.0, 0.0], atol=1e-3)
    assert np.allclose(pybullet.get_base_angular_velocity("my_box"), [0.0, 0.0, 0.0], atol=1e-3)


def test_load_urdf_and_control_joints(pybullet):            # :101-121
    _load_panda(pybullet)
    pybullet.control_joints("panda", [5], [0.3], [5.0])
    pybullet.step()


def test_get_link_position(pybullet):                       # :124-136
    _load_panda(pybullet)
    assert np.allclose(pybullet.get_link_position("panda", 1), [0.000, 0.060, 0.373], atol=1e-3)


def test_get_link_orientation(pybullet):                    # :139-153
    _load_panda(pybullet)
    pybullet.control_joints("panda", [5], [0.3], [5.0])
    pybullet.step()
    assert np.allclose(pybullet.get_link_orientation("panda", 5), [0.707, -0.02, 0.02, 0.707], atol=1e-3)


def test_get_link_velocity(pybullet):                       # :156-170
    _load_panda(pybullet)
    pybullet.control_joints("panda", [5], [0.3], [5.0])
    pybullet.step()
    assert np.allclose(pybullet.get_link_velocity("panda", 5), [-0.0068, 0.0000, 0.1186], atol=1e-3)


def test_get_link_angular_velocity(pybullet):               # :173-187
    _load_panda(pybullet)
    pybullet.control_joints("panda", [5], [0.3], [5.0])
    pybullet.step()
    assert np.allclose(pybullet.get_link_angular_velocity("panda", 5), [0.000, -2.969, 0.000], atol=1e-3)


def test_get_joint_angle(pybullet):                         # :190-204
    _load_panda(pybullet)
    pybullet.control_joints("panda", [5], [0.3], [5.0])
    pybullet.step()
    assert np.allclose(pybullet.get_joint_angle("panda", 5), 0.063, atol=1e-3)


def test_set_base_pose(pybullet):                           # :207-218
    _box(pybullet)
    pybullet.set_base_pose("my_box", [1.0, 1.0, 1.0], [0.707, -0.02, 0.02, 0.707])
    assert np.allclose(pybullet.get_base_position("my_box"), [1.0, 1.0, 1.0], atol=1e-3)
    assert np.allclose(pybullet.get_base_orientation("my_box"), [0.707, -0.02, 0.02, 0.707], atol=1e-3)


def test_set_joint_angle_and_angles(pybullet):              # :221-251
    _load_panda(pybullet)
    pybullet.set_joint_angle("panda", 3, 0.4)
    assert np.allclose(pybullet.get_joint_angle("panda", 3), 0.4, atol=1e-3)
    pybullet.set_joint_angles("panda", [3, 4], [0.4, 0.5])
    assert np.allclose(pybullet.get_joint_angle("panda", 3), 0.4, atol=1e-3)
    assert np.allclose(pybullet.get_joint_angle("panda", 4), 0.5, atol=1e-3)


def test_inverse_kinematics(pybullet):                      # :254-266  link 6, un-normalised target quaternion
    _load_panda(pybullet)
    joint_angles = pybullet.inverse_kinematics("panda", 6, [0.4, 0.5, 0.6], [0.707, -0.02, 0.02, 0.707])
    assert np.allclose(joint_angles, [1.000, 1.223, -1.113, -0.021, -0.917, 0.666, -0.499, 0.0, 0.0], atol=1e-3)


def test_scene_builders_and_friction_setters(pybullet):     # :269-323
    pybullet.place_visualizer([0.1, 0.2, 0.3], 5.0, 0.3, 0.4)
    pybullet.create_cylinder("my_cylinder", 0.5, 1.0, 1.0, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 1.0])
    pybullet.create_sphere("my_sphere", 0.5, 1.0, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 1.0], ghost=True)
    pybullet.create_plane(0.5)
    pybullet.create_table(0.5, 0.6, 0.4)
    pybullet.set_lateral_friction("my_cylinder", 0, 0.5)
    pybullet.set_spinning_friction("my_cylinder", 0, 0.5)
    pybullet.step()


def test_facade_getters_on_a_task_env_match_the_bare_world():
    """a16: get_link_position / orientation / velocity / angular_velocity for ANY link on a bound (task) facade are the same device
    code as on the bare world: same joint state -> same link states; the EE link reproduces the observation of the fused step."""
    from panda_lang_manip_b200.panda_gym.envs import PandaPushEnv
    from panda_lang_manip_b200.panda_gym.pybullet import PyBullet
    env = PandaPushEnv()
    env.reset(seed=3)
    rng = np.random.default_rng(0)
    for _ in range(3):
        obs, *_ = env.step(rng.uniform(-1, 1, 3).astype(np.float32))
    bare = PyBullet()
    bare.loadURDF(body_name="panda", fileName="franka_panda/panda.urdf", basePosition=[-0.6, 0.0, 0.0], useFixedBase=True)
    st = env.sim._state()
    w = bare._world()
    row = w.get_state()[0].cpu().numpy(); row[:18] = st[:18]
    w.set_state(torch.as_tensor(row[None, :]))
    for link in range(12):
        a = np.concatenate([env.sim.get_link_position("panda", link), env.sim.get_link_orientation("panda", link), env.sim.get_link_velocity("panda", link), env.sim.get_link_angular_velocity("panda", link)])
        b = np.concatenate([bare.get_link_position("panda", link), bare.get_link_orientation("panda", link), bare.get_link_velocity("panda", link), bare.get_link_angular_velocity("panda", link)])
        assert np.array_equal(a, b), link
    assert np.allclose(env.robot.get_ee_position(), obs["observation"][:3], atol=2e-6) and np.allclose(env.robot.get_ee_velocity(), obs["observation"][3:6], atol=2e-5)
    assert np.isclose(env.robot.get_fingers_width(), st[7] + st[8])
    with pytest.raises(NotImplementedError):
        env.sim.control_joints("panda", [5], [0.3], [5.0])      # task envs derive their motor targets inside the fused step
    bare.close(); env.close()


def test_bare_world_batch_matches_oracle():
    """pg_sim_step on a batch of bare worlds with different raw motor settings vs the oracle, 3 x 20 sub-steps, fp32 at 1e-4 rad."""
    from panda_lang_manip_b200.bare_world import PandaBareWorld
    from tests.oracle_util import OracleSim
    n = 64
    rng = np.random.default_rng(1)
    w = PandaBareWorld(n)
    q0 = np.array([0.0, 0.41, 0.0, -1.85, 0.0, 2.26, 0.79, 0.01, 0.01]) + rng.uniform(-0.2, 0.2, (n, 9)) * np.array([1, 1, 1, 1, 1, 1, 1, 0.03, 0.03])
    st = w.get_state().cpu().numpy(); st[:, :9] = q0
    w.set_state(torch.as_tensor(st))
    mot = w.get_motors().cpu().numpy()
    tq = q0 + rng.uniform(-0.15, 0.15, (n, 9)) * np.array([1, 1, 1, 1, 1, 1, 1, 0.05, 0.05])
    forces = np.array([87.0, 87.0, 87.0, 87.0, 12.0, 120.0, 120.0, 170.0, 170.0])
    pos_ctrl = rng.random((n, 9)) < 0.7
    for i in range(n):
        for d in range(9):
            if pos_ctrl[i, d]:
                mot[i, d] = [0.1, 1.0, tq[i, d], 0.0, forces[d]]
    w.set_motors(torch.as_tensor(mot))
    links = [0, 1, 2, 3, 4, 5, 6, 9, 10]
    sims = []
    for i in range(8):
        s = OracleSim()
        for d, l in enumerate(links):
            s.reset_joint(l, q0[i, d])
            if pos_ctrl[i, d]:
                s.control_joint(l, tq[i, d], forces[d])
        sims.append(s)
    for rep in range(3):
        w.step(20)
        got = w.get_state().cpu().numpy()
        ls = w.link_state(11).cpu().numpy()
        for i, s in enumerate(sims):
            s.step(20)
            qo = np.array([s.joint(l)[0] for l in links])
            assert np.abs(got[i, :9] - qo).max() < 1e-4, (rep, i, got[i, :9] - qo)
            p, qt, v, wv = s.link_state(11)
            assert np.abs(ls[i, :3] - p).max() < 1e-4 and np.abs(ls[i, 7:10] - v).max() < 2e-2
    for s in sims:
        s.close()
    w.close()
