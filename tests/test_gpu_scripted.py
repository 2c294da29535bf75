"""Scripted-policy success rates, GPU vs oracle on the SAME episodes (north_star: "scripted-policy success rate within +-1 pp over 10k
episodes"; there: against PyBullet, unavailable here -> against the oracle).  PickAndPlace (finger grasp): 10,000 episodes on the GPU
and all 10,000 replayed on the oracle (one process per host core); Push (box-table sliding contact): 10,000; Stack (two objects,
100-step episodes): 4,096.  Identical goals, placements and policy arithmetic (tests/scripted.py over numpy / torch)."""
import numpy as np
import pytest

from tests.scripted import EPISODE_STEPS, POLICIES, oracle_success, sample_episodes

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _gpu_success(task, goals, objs):
    import panda_lang_manip_b200 as p
    n, steps = len(goals), EPISODE_STEPS[task]
    env = p.PandaVecEnv(task, n, control_type="ee", auto_reset=False)
    obs = env.reset(goals=goals, object_positions=objs)
    g = torch.from_numpy(goals.astype(np.float32)).cuda()
    phase = torch.zeros(n, dtype=torch.int64, device="cuda"); count = torch.zeros_like(phase)
    done = torch.zeros(n, dtype=torch.bool, device="cuda")
    for t in range(steps):
        a = POLICIES[task](torch, obs["observation"], g, phase, count)
        obs, rew, term, trunc, _ = env.step(a)
        done |= term.bool()
    out = done.cpu().numpy()
    assert env.diverged() == 0
    env.close()
    return out


@pytest.mark.parametrize("task,n,min_rate", [("pick_and_place", 10000, 0.9), ("push", 10000, 0.6), ("stack", 4096, 0.4)])
def test_scripted_success_rate_matches_oracle(task, n, min_rate):
    goals, objs = sample_episodes(task, n, seed=7)
    gpu = _gpu_success(task, goals, objs)
    ref, _ = oracle_success(task, goals, objs)
    agree = (gpu == ref).mean()
    print(f"scripted {task}: gpu {gpu.mean():.4f} vs oracle {ref.mean():.4f} over the same {n} episodes; per-episode agreement {agree:.4f}")
    assert gpu.mean() > min_rate, gpu.mean()                  # the script actually solves the task (grasp / push / stack work)
    assert abs(gpu.mean() - ref.mean()) <= 0.01, (gpu.mean(), ref.mean())
    # per-episode agreement: the grasp and the push are robust (measured 1.0000 / 0.990); Stack's script carries the second cube in a grasp
    # that pivots about the two fingertip contact lines and sets it down while still moving, so whether the cube ends on top or beside is
    # decided at fp32 resolution in many episodes (measured 0.87; the oracle ends with the cube on top in 17 % of the episodes, the 70 %
    # "success" is the task's loose 0.1 threshold on the 6-D goal distance) -- the RATE still agrees
    assert agree >= {"pick_and_place": 0.97, "push": 0.95, "stack": 0.8}[task], agree
