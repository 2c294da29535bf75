"""Analytic known-answer tests for contact -- physics pins where reference (PyBullet) vectors cannot exist in this container.  Each
scenario is run through a backend-neutral `Runner` (oracle or CUDA path, same call sequence) and compared with a closed-form /
recurrence answer that follows from the reference's scene constants alone (dt 1/500, g 9.81, cube 4 cm / 1 kg,
lateral friction 0.5 default x 0.5 table = 0.25, Slide puck 0.04 x 0.5 = 0.02, Bullet's link damping 0.04)."""
import numpy as np

DT, G, KD = 1.0 / 500.0, 9.81, 0.04
NEUTRAL = np.array([0.0, 0.41, 0.0, -1.85, 0.0, 2.26, 0.79, 0.0, 0.0])


def state_row(task, objs, goal):
    """q(9) qd(9) | objects pos3 quat4 lin3 ang3 | goal: robot at the neutral pose (gripper 20 cm above the table, out of the way)."""
    row = [NEUTRAL, np.zeros(9)]
    for o in objs:
        pos, vel = o[0], o[1]
        quat = o[2] if len(o) > 2 else [0, 0, 0, 1]
        row += [np.asarray(pos, float), np.asarray(quat, float), np.asarray(vel, float), [0, 0, 0]]
    row.append(np.asarray(goal, float))
    return np.concatenate([np.asarray(r, float) for r in row])


def sliding_recurrence(v0, mu, substeps):
    """Velocity and travelled distance of a body sliding flat on the table: per sub-step Bullet's damping v *= 1 - k (1 + |v|) dt,
    then the friction impulse mu * (normal impulse m g dt) opposing the motion, semi-implicit Euler on the position."""
    v, x = v0, 0.0
    for _ in range(substeps):
        v = v * (1.0 - KD * (1.0 + abs(v)) * DT)
        v = max(0.0, v - mu * G * DT)
        x += v * DT
    return v, x


def run_all(make_runner):
    """make_runner(task) -> object with .set(row), .step(action) -> state row (same layout), .close().  Returns {name: (got, want, tol)}."""
    out = {}
    zero7 = np.zeros(7, np.float32)
    # 1. a cube at rest on the table stays put: height = half size, no drift, no spin (resting contact: 4 corner contacts, ERP, slop)
    r = make_runner("push")
    r.set(state_row("push", [([0.1, 0.05, 0.02], [0, 0, 0])], [0, 0, 0.02]))
    for _ in range(25):
        st = r.step(zero7)
    out["rest_height"] = (st[20], 0.02, 2e-4)
    out["rest_drift_xy"] = (float(np.abs(st[18:20] - [0.1, 0.05]).max()), 0.0, 1e-4)
    out["rest_speed"] = (float(np.abs(st[25:31]).max()), 0.0, 2e-3)
    out["rest_upright"] = (float(np.abs(st[21:24]).max()), 0.0, 1e-4)
    # 2. free fall from 5 cm: after one env step (20 sub-steps) z = z0 - g dt^2 n(n+1)/2 (semi-implicit Euler; damping < 1e-5), then it
    # lands and settles on the table within a second
    r.set(state_row("push", [([0.1, 0.05, 0.07], [0, 0, 0])], [0, 0, 0.02]))
    st = r.step(zero7)
    out["drop_first_step_z"] = (st[20], 0.07 - G * DT * DT * 210, 1e-4)
    for _ in range(24):
        st = r.step(zero7)
    out["drop_settled_z"] = (st[20], 0.02, 3e-4)
    out["drop_settled_speed"] = (float(np.abs(st[25:31]).max()), 0.0, 3e-3)
    # 3. a cube sliding at 0.5 m/s stops by Coulomb friction mu = 0.5 x 0.5: travelled distance ~ v0^2 / (2 mu g) (+ damping)
    r.set(state_row("push", [([-0.1, 0.0, 0.02], [0.5, 0, 0])], [0, 0, 0.02]))
    st = r.step(zero7)                                   # 20 sub-steps: still sliding
    v_want, _ = sliding_recurrence(0.5, 0.25, 20)
    out["cube_slide_velocity_after_one_step"] = (st[25], v_want, 0.01)
    for _ in range(11):
        st = r.step(zero7)
    _, x_want = sliding_recurrence(0.5, 0.25, 240)
    out["cube_slide_distance"] = (st[18] + 0.1, x_want, 0.03 * x_want)
    out["cube_slide_stops"] = (float(np.abs(st[25:28]).max()), 0.0, 2e-3)
    out["cube_slide_straight"] = (float(abs(st[19])), 0.0, 1e-3)
    # 3b. a cube dropped tilted about x lands on an edge and settles on the nearer face: 30 degrees falls back flat, 60 degrees rolls on to
    # the next face (quaternion (sin 45, 0, 0, cos 45)); either way it ends at rest at the half-size height
    for name, deg, qx in (("30", 30.0, 0.0), ("60", 60.0, np.sin(np.pi / 4))):
        ang = np.radians(deg)
        r.set(state_row("push", [([0.1, 0.05, 0.07], [0, 0, 0], [np.sin(ang / 2), 0, 0, np.cos(ang / 2)])], [0, 0, 0.02]))
        for _ in range(50):
            st = r.step(zero7)
        out[f"tilted_{name}_deg_drop_settles_height"] = (st[20], 0.02, 3e-4)
        out[f"tilted_{name}_deg_drop_settles_speed"] = (float(np.abs(st[25:31]).max()), 0.0, 3e-3)
        out[f"tilted_{name}_deg_drop_ends_on_a_face"] = (float(abs(st[21])), qx, 2e-3)
    r.close()
    # 4. the Slide puck (lateral friction 0.04 x table 0.5 = 0.02) decelerates at mu g: velocity after 10 env steps
    r = make_runner("slide")
    r.set(state_row("slide", [([0.0, 0.0, 0.015], [0.5, 0, 0])], [0.4, 0, 0.015]))
    for _ in range(10):
        st = r.step(zero7)
    v_want, x_want = sliding_recurrence(0.5, 0.02, 200)
    out["puck_velocity"] = (st[25], v_want, 0.01 * v_want)
    out["puck_distance"] = (st[18], x_want, 0.01 * x_want)
    out["puck_height"] = (st[20], 0.015, 2e-4)
    r.close()
    # 4b. PickAndPlace: the 1 kg cube between the closed fingers of the arm at its neutral pose (grasp target at (0.0384, 0, 0.1974), 18 cm above
    # the table) is held against gravity by friction alone (finger 1.0 x cube 0.5, 20 N per finger from the position motors: 20 N > m g) while
    # the gripper keeps closing, and drops when the gripper opens
    r = make_runner("pick_and_place")
    row = state_row("pick_and_place", [([0.0384, 0.0, 0.1974], [0, 0, 0])], [0, 0, 0.1])
    row[7] = row[8] = 0.0205
    r.set(row)
    act = np.zeros(8, np.float32); act[7] = -1.0
    for _ in range(30):
        st = r.step(act)
    out["grasped_cube_is_held_height"] = (st[20], 0.1974, 1e-3)
    out["grasped_cube_is_held_xy"] = (float(np.abs(st[18:20] - [0.0384, 0.0]).max()), 0.0, 1e-3)
    out["grasped_cube_is_held_speed"] = (float(np.abs(st[25:28]).max()), 0.0, 5e-3)
    act[7] = 1.0
    for _ in range(25):
        st = r.step(act)
    out["released_cube_lands_on_the_table"] = (st[20], 0.02, 1e-3)
    r.close()
    # 5. Stack (tasks/stack.py:30-62: a 2 kg and a 1 kg 4 cm cube): the second cube resting on the first stays there -- exactly aligned (every
    # vertex at a corner of the other cube's face: the reference-face rule decides the normal), shifted, and turned by 20 / 45 degrees about
    # z, where every vertex of either cube lies outside the other's face and only the edge-against-edge contacts carry it (8 crossings of
    # the two squares); a cube whose centre of mass overhangs the lower cube's edge (25 mm of 20) tips off and ends up on the table
    r = make_runner("stack")
    zero8 = np.zeros(8, np.float32)
    for name, ang, off in (("exactly_aligned", 0.0, 0.0), ("shifted_2mm", 0.0, 0.002), ("turned_45_deg", np.pi / 4, 0.002), ("turned_20_deg", np.radians(20.0), 0.0005)):
        quat = [0, 0, np.sin(ang / 2), np.cos(ang / 2)]
        r.set(state_row("stack", [([0.1, 0.05, 0.02], [0, 0, 0]), ([0.1 + off, 0.05 + off / 2, 0.06], [0, 0, 0], quat)], [0.1, 0.05, 0.02, 0.1, 0.05, 0.06]))
        for _ in range(25):
            st = r.step(zero8)
        out[f"stacked_{name}_upper_height"] = (st[33], 0.06, 3e-4)
        out[f"stacked_{name}_lower_height"] = (st[20], 0.02, 3e-4)
        out[f"stacked_{name}_upper_drift_xy"] = (float(np.abs(st[31:33] - [0.1 + off, 0.05 + off / 2]).max()), 0.0, 3e-4)
        out[f"stacked_{name}_upper_speed"] = (float(np.abs(st[38:44]).max()), 0.0, 5e-3)
        out[f"stacked_{name}_upper_keeps_its_yaw"] = (float(abs(st[37]) - np.cos(ang / 2)), 0.0, 1e-3)
    r.set(state_row("stack", [([0.1, 0.05, 0.02], [0, 0, 0]), ([0.125, 0.05, 0.06], [0, 0, 0])], [0.1, 0.05, 0.02, 0.1, 0.05, 0.06]))
    for _ in range(40):
        st = r.step(zero8)
    out["overhanging_cube_falls_to_the_table"] = (st[33], 0.02, 1e-3)
    out["overhanging_cube_leaves_the_lower_one_in_place"] = (float(np.abs(st[18:20] - [0.1, 0.05]).max()), 0.0, 2e-3)
    r.close()
    return out


def check(results):
    bad = {k: v for k, v in results.items() if not abs(v[0] - v[1]) <= v[2]}
    assert not bad, bad
