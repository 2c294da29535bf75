"""Scripted policies used for the success-rate parity check (north_star: "scripted-policy success rate within +-1 pp over 10k
episodes"), written once over an array namespace so that the numpy (oracle) and torch (GPU batch) versions are the same arithmetic.
The fork's own scripted layer is robots/panda_cartesian.py:98-145 (move / grasp / release); these are the equivalent closed-loop
scripts for the registered tasks.  `state` is a dict of per-env int arrays, updated in place."""
import os

import numpy as np


def _stack(xp, cols):
    return xp.stack(cols, -1)


def _goto(xp, ee, tgt, gain=0.05):
    return xp.clip((tgt - ee) / gain, -1.0, 1.0)


def scripted_pick_and_place(xp, obs, goal, phase, count):
    """obs [N,19] float32, goal [N,3]; phase/count [N] int arrays (state of the script, updated in place).  Returns actions [N,4]."""
    ee, obj = obs[:, 0:3], obs[:, 7:10]
    zero = xp.zeros_like(ee[:, 0])
    above = obj + _stack(xp, [zero, zero, zero + 0.08])
    carry = goal + (ee - obj)
    tgt = xp.where((phase == 0)[:, None], above, xp.where((phase == 1)[:, None], obj, xp.where((phase == 2)[:, None], ee, carry)))
    move = _goto(xp, ee, tgt)
    grip = xp.where(phase <= 1, zero + 1.0, zero - 1.0)
    act = xp.concatenate([move, grip[:, None]], -1)
    d_above = ((ee - above) ** 2).sum(-1) ** 0.5
    d_obj = ((ee - obj) ** 2).sum(-1) ** 0.5
    to1 = (phase == 0) & (d_above < 0.01)
    to2 = (phase == 1) & (d_obj < 0.008)
    count += (phase == 2)
    to3 = (phase == 2) & (count >= 4)
    phase += to1 * 1 + to2 * 1 + to3 * 1
    return act


def scripted_push(xp, obs, goal, phase, count):
    """PandaPush-v3 (blocked gripper: the closed fingers are a 2 x 4 cm block; 3-D ee actions): two axis-aligned face pushes -- get
    behind the cube in x, descend, push until the cube's x is the goal's, rise, then the same in y.  obs [N,18]: ee 0:3, cube 6:9.
    Phases: 0 hover / 1 descend / 2 push (x axis), 3 rise, 4 hover / 5 descend / 6 push (y axis), 7 done."""
    ee, obj = obs[:, 0:3], obs[:, 6:9]
    zero = xp.zeros_like(ee[:, 0])
    along_y = phase >= 4
    ph = phase - 4 * along_y
    ax = along_y * 1                                            # 0: x, 1: y
    e_ax = _stack(xp, [1.0 - ax + zero, ax + zero])             # unit vector of the push axis
    err = ((goal - obj)[:, 0:2] * e_ax).sum(-1)                 # signed distance still to cover along the axis
    sg = xp.where(err >= 0, zero + 1.0, zero - 1.0)
    stand = xp.where(along_y, zero + 0.075, zero + 0.06)        # stand-off behind the cube: the finger block is 2.1 cm deep in x, 4.2 cm in y
    behind_xy = obj[:, 0:2] - e_ax * (sg * stand)[:, None]
    hover = xp.concatenate([behind_xy, (zero + 0.10)[:, None]], -1)
    low = xp.concatenate([behind_xy, (zero + 0.022)[:, None]], -1)
    touch = xp.where(along_y, zero + 0.041, zero + 0.0305)     # centre distance at which the finger block touches the cube
    lat = obj[:, 0:2] * (1.0 - e_ax) + (obj[:, 0:2] - e_ax * (sg * touch)[:, None]) * e_ax   # centred on the cube across the axis, touching it along the axis ...
    lead = e_ax * (sg * xp.clip(xp.abs(err) * 0.25, 0.004, 0.03))[:, None]       # ... and lead along it: 0.5 m/s far away, slower close to the goal (the cube coasts v^2 / 2 mu g)
    push = xp.concatenate([lat + lead, (zero + 0.022)[:, None]], -1)
    rise = xp.concatenate([ee[:, 0:2] - e_ax * (sg * 0.03)[:, None], (zero + 0.08)[:, None]], -1)   # back off the cube while rising (a finger leaving a contact sideways kicks it)
    tgt = xp.where((ph == 0)[:, None], hover, xp.where((ph == 1)[:, None], low, xp.where((ph == 2)[:, None], push, rise)))
    move = _goto(xp, ee, tgt)
    # while pushing, cap the speed along the axis (0.75 m/s far from the goal, 0.1 m/s close to it): contact is stiff, a fast pusher kicks
    # the cube ahead of itself and the cube then coasts v^2 / (2 mu g) past the goal
    vmax = xp.clip(xp.abs(err) / 0.1, 0.15, 0.6)[:, None]
    capped = xp.concatenate([xp.where(e_ax > 0.5, xp.clip(move[:, 0:2], -vmax, vmax), move[:, 0:2]), move[:, 2:3]], -1)
    move = xp.where((ph == 2)[:, None], capped, move)
    move = xp.where((phase >= 7)[:, None], move * 0.0, move)
    d_hover = ((ee - hover) ** 2).sum(-1) ** 0.5
    d_low = ((ee - low) ** 2).sum(-1) ** 0.5
    small = xp.abs(err) < 0.008
    t01 = (ph == 0) & (d_hover < 0.03)
    t12 = (ph == 1) & (d_low < 0.012)
    t23 = (ph == 2) & small
    t34 = (phase == 3) & (ee[:, 2] > 0.055)
    skip = (ph == 0) & small & (phase < 7)                      # already there along this axis: skip its three phases (and the rise)
    adv = (t01 | t12 | t23 | t34) & (phase < 7) & ~skip
    phase += adv * 1 + skip * xp.where(along_y, 3, 4)
    return move


def scripted_stack(xp, obs, goal, phase, count):
    """PandaStack-v3: pick cube 1 -> goal 1, release, retreat upwards, pick cube 2 -> goal 2 (on top of cube 1), release.
    obs [N,31]: ee 0:3, finger width 6, cube1 pos 7:10, cube2 pos 19:22.  goal [N,6].  Phases 0-5 handle cube 1, 6-11 cube 2:
    (0) above, (1) descend, (2) close, (3) carry, (4) open, (5) rise."""
    ee = obs[:, 0:3]
    zero = xp.zeros_like(ee[:, 0])
    second = phase >= 6
    ph = phase - 6 * second
    obj = xp.where(second[:, None], obs[:, 19:22], obs[:, 7:10])
    g = xp.where(second[:, None], goal[:, 3:6], goal[:, 0:3])
    up = _stack(xp, [zero, zero, zero + 0.09])
    above = obj + up
    dxy = (((obj - g)[:, 0:2] ** 2).sum(-1) + 1e-12) ** 0.5
    arc = xp.clip(dxy * 1.5, 0.004, 0.05)                              # carried in an arc: 5 cm up while far, a few mm high on arrival (set down, not pressed in)
    carry = g + (ee - obj) + _stack(xp, [zero, zero, arc])
    rise = ee + up
    tgt = xp.where((ph == 0)[:, None], above, xp.where((ph == 1)[:, None], obj, xp.where((ph == 2)[:, None], ee, xp.where((ph == 3)[:, None], carry, xp.where((ph == 4)[:, None], ee, rise)))))
    move = _goto(xp, ee, tgt)
    move = xp.where((ph == 3)[:, None], xp.clip(move, -0.5, 0.5), move)          # carry at half speed: the cube pivots between two finger pads
    move = xp.where((phase >= 12)[:, None], move * 0.0, move)
    grip = xp.where((ph <= 1) | (ph == 5), zero + 1.0, xp.where(ph == 4, zero + 0.25, zero - 1.0))   # release gently: a finger snapping open kicks the cube
    act = xp.concatenate([move, grip[:, None]], -1)
    d_above = ((ee - above) ** 2).sum(-1) ** 0.5
    d_obj = ((ee - obj) ** 2).sum(-1) ** 0.5
    d_goal = ((obj - g) ** 2).sum(-1) ** 0.5
    timed = (ph == 2) | (ph == 4) | (ph == 5)
    count += timed * 1
    t01 = (ph == 0) & (d_above < 0.012)
    t12 = (ph == 1) & (d_obj < 0.008)
    t23 = (ph == 2) & (count >= 4)
    t34 = (ph == 3) & (d_goal < 0.01)
    t45 = (ph == 4) & (count >= 5)
    t56 = (ph == 5) & (count >= 3)
    adv = (t01 | t12 | t23 | t34 | t45 | t56) & (phase < 12)
    lost = (ph == 3) & (d_obj > 0.04) & (phase < 12)                   # dropped on the way: start over with this cube
    count *= (1 - (adv | lost) * 1)
    phase += adv * 1 - lost * 3
    return act


POLICIES = {"pick_and_place": scripted_pick_and_place, "push": scripted_push, "stack": scripted_stack}
EPISODE_STEPS = {"pick_and_place": 50, "push": 50, "stack": 100}


def sample_episodes(task, n, seed):
    """Goals and object placements of n episodes (the reference's ranges; PickAndPlace goals always in the air so that the grasp matters)."""
    rng = np.random.default_rng(seed)
    xy = lambda: np.stack([rng.uniform(-0.15, 0.15, n), rng.uniform(-0.15, 0.15, n)], -1)
    if task == "pick_and_place":
        goals = np.concatenate([xy(), (0.02 + rng.uniform(0.0, 0.2, n))[:, None]], -1)
        objs = np.concatenate([xy(), np.full((n, 1), 0.02)], -1)
    elif task == "push":
        goals = np.concatenate([xy(), np.full((n, 1), 0.02)], -1)
        objs = np.concatenate([xy(), np.full((n, 1), 0.02)], -1)
    else:
        g = xy()
        goals = np.concatenate([g, np.full((n, 1), 0.02), g, np.full((n, 1), 0.06)], -1)
        objs = np.concatenate([xy(), np.full((n, 1), 0.02), xy(), np.full((n, 1), 0.06)], -1)
    return goals, objs


def _oracle_slice(args):
    from tests.oracle_util import OracleBatch
    task, goals, objs, steps = args
    n = len(goals)
    ob = OracleBatch(task, n, "ee")
    obs = ob.reset(goals, objs)
    g32 = goals.astype(np.float32)
    phase, count = np.zeros(n, np.int64), np.zeros(n, np.int64)
    done = np.zeros(n, bool)
    for t in range(steps):
        a = POLICIES[task](np, obs.astype(np.float32), g32, phase, count).astype(np.float32)
        obs, rew, term = ob.step(a)
        done |= term.astype(bool)
    final = obs.copy()
    ob.close()
    return done, final


def oracle_success(task, goals, objs, steps=None, procs=None):
    """The scripted policy on the CPU oracle for every episode, one process per host core (fork; each steps its slice through one C
    call per env step).  Returns (success [n] bool, final observation [n, O])."""
    import multiprocessing as mp
    steps = steps or EPISODE_STEPS[task]
    procs = procs or max(1, min(os.cpu_count() or 1, 64))
    n = len(goals)
    chunks = [(task, goals[i::procs], objs[i::procs], steps) for i in range(procs) if len(goals[i::procs])]
    from tests.oracle_util import build_oracle
    build_oracle()
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        res = pool.map(_oracle_slice, chunks)
    done = np.zeros(n, bool); final = None
    for i, (d, f) in enumerate(res):
        done[i::procs] = d
        if final is None:
            final = np.zeros((n, f.shape[1]), np.float32)
        final[i::procs] = f
    return done, final
