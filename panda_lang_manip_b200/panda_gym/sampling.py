"""Goal / object sampling of the six tasks with the reference's RNG contract: the task's generator is
np.random.Generator(PCG64(SeedSequence(seed))) re-created on every reset (reference envs/core.py:243-244) and the draw order is
that of Task.reset (reach.py:47-54, push.py:69-87, slide.py:73-91, pick_and_place.py:65-85, stack.py:94-119, flip.py:63-80)."""
import numpy as np


def default_ranges(task: str, goal_range=None, goal_xy_range=None, goal_z_range=None, goal_x_offset=None, obj_xy_range=None):
    """(goal_range_low, goal_range_high, obj_range_low, obj_range_high) exactly as the reference constructors build them from their
    keyword arguments (reach.py:21-23, push.py:20-23, slide.py:22-25, pick_and_place.py:24-27, stack.py:20-23, flip.py:21-22)."""
    if task == "reach":
        g = 0.3 if goal_range is None else goal_range
        return np.array([-g / 2, -g / 2, 0]), np.array([g / 2, g / 2, g]), np.zeros(3), np.zeros(3)
    g = 0.3 if goal_xy_range is None else goal_xy_range
    o = 0.3 if obj_xy_range is None else obj_xy_range
    x = (0.4 if goal_x_offset is None else goal_x_offset) if task == "slide" else 0.0
    z = (0.2 if goal_z_range is None else goal_z_range) if task == "pick_and_place" else 0.0
    return np.array([-g / 2 + x, -g / 2, 0]), np.array([g / 2 + x, g / 2, z]), np.array([-o / 2, -o / 2, 0]), np.array([o / 2, o / 2, 0])


def sample_reset(task: str, rng: np.random.Generator, ranges=None):
    """Returns (goal[G], [object positions...]) drawn exactly as the reference draws them."""
    glo, ghi, olo, ohi = default_ranges(task) if ranges is None else ranges
    if task == "reach":
        return rng.uniform(glo, ghi), []
    if task == "push":
        goal = np.array([0.0, 0.0, 0.02]) + rng.uniform(glo, ghi)
        return goal, [np.array([0.0, 0.0, 0.02]) + rng.uniform(olo, ohi)]
    if task == "slide":
        goal = np.array([0.0, 0.0, 0.03]) + rng.uniform(glo, ghi)
        return goal.copy(), [np.array([0.0, 0.0, 0.03]) + rng.uniform(olo, ohi)]
    if task == "pick_and_place":
        goal = np.array([0.0, 0.0, 0.02])
        noise = rng.uniform(glo, ghi)
        if rng.random() < 0.3:
            noise[2] = 0.0
        goal += noise
        return goal, [np.array([0.0, 0.0, 0.02]) + rng.uniform(olo, ohi)]
    if task == "stack":
        noise = rng.uniform(glo, ghi)
        goal = np.concatenate((np.array([0.0, 0.0, 0.02]) + noise, np.array([0.0, 0.0, 0.06]) + noise))
        n1, n2 = rng.uniform(olo, ohi), rng.uniform(olo, ohi)
        return goal, [np.array([0.0, 0.0, 0.02]) + n1, np.array([0.0, 0.0, 0.06]) + n2]
    if task == "flip":
        # the reference draws the goal from scipy's unseeded global RNG (flip.py:71); here: a uniform rotation from the task RNG
        q = rng.normal(size=4)
        goal = q / np.linalg.norm(q)
        return goal, [np.array([0.0, 0.0, 0.02]) + rng.uniform(olo, ohi)]
    raise ValueError(task)
