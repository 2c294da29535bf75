"""Scripted pick-and-place policy (approach above the cube, descend, close, carry to the goal) used for the success-rate parity
check.  Written once over an array namespace so the numpy (oracle) and torch (GPU batch) versions are the same arithmetic."""
import numpy as np


def scripted_pick_and_place(xp, obs, goal, phase, count):
    """obs [N,19] float32, goal [N,3]; phase/count [N] int arrays (state of the script, updated in place).  Returns actions [N,4]."""
    ee, obj = obs[:, 0:3], obs[:, 7:10]
    zero = xp.zeros_like(ee[:, 0])
    above = obj + xp.stack([zero, zero, zero + 0.08], -1)
    carry = goal + (ee - obj)
    tgt = xp.where((phase == 0)[:, None], above, xp.where((phase == 1)[:, None], obj, xp.where((phase == 2)[:, None], ee, carry)))
    move = xp.clip((tgt - ee) / 0.05, -1.0, 1.0)
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
