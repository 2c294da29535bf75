"""Goal / object sampling of the six tasks with the reference's RNG contract: the task's generator is
np.random.Generator(PCG64(SeedSequence(seed))) re-created on every reset (reference envs/core.py:243-244) and the draw order is
that of Task.reset (reach.py:47-54, push.py:69-87, slide.py:73-91, pick_and_place.py:65-85, stack.py:94-119, flip.py:63-80)."""
import numpy as np


def sample_reset(task: str, rng: np.random.Generator):
    """Returns (goal[G], [object positions...]) drawn exactly as the reference draws them."""
    xy = np.array([0.15, 0.15, 0.0])
    if task == "reach":
        return rng.uniform(np.array([-0.15, -0.15, 0.0]), np.array([0.15, 0.15, 0.3])), []
    if task == "push":
        goal = np.array([0.0, 0.0, 0.02]) + rng.uniform(-xy, xy)
        return goal, [np.array([0.0, 0.0, 0.02]) + rng.uniform(-xy, xy)]
    if task == "slide":
        goal = np.array([0.0, 0.0, 0.03]) + rng.uniform(np.array([0.25, -0.15, 0.0]), np.array([0.55, 0.15, 0.0]))
        return goal.copy(), [np.array([0.0, 0.0, 0.03]) + rng.uniform(-xy, xy)]
    if task == "pick_and_place":
        goal = np.array([0.0, 0.0, 0.02])
        noise = rng.uniform(np.array([-0.15, -0.15, 0.0]), np.array([0.15, 0.15, 0.2]))
        if rng.random() < 0.3:
            noise[2] = 0.0
        goal += noise
        return goal, [np.array([0.0, 0.0, 0.02]) + rng.uniform(-xy, xy)]
    if task == "stack":
        noise = rng.uniform(-xy, xy)
        goal = np.concatenate((np.array([0.0, 0.0, 0.02]) + noise, np.array([0.0, 0.0, 0.06]) + noise))
        n1, n2 = rng.uniform(-xy, xy), rng.uniform(-xy, xy)
        return goal, [np.array([0.0, 0.0, 0.02]) + n1, np.array([0.0, 0.0, 0.06]) + n2]
    if task == "flip":
        # the reference draws the goal from scipy's unseeded global RNG (flip.py:71); here: a uniform rotation from the task RNG
        q = rng.normal(size=4)
        goal = q / np.linalg.norm(q)
        return goal, [np.array([0.0, 0.0, 0.02]) + rng.uniform(-xy, xy)]
    raise ValueError(task)
