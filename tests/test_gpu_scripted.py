"""Scripted-policy success on PandaPickAndPlace-v3 (finger grasp contact): 10k episodes on the GPU, the first 128 of them replayed on
the oracle with identical goals / object placements / policy.  north_star: success rate within +-1 pp over 10k episodes (there:
against PyBullet, which is unavailable here -> against the oracle, on the episodes the oracle can afford, plus the 10k-episode rate)."""
import numpy as np
import pytest

from tests.oracle_util import OracleEnv
from tests.scripted import scripted_pick_and_place

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_scripted_pick_and_place_success_rate():
    import panda_lang_manip_b200 as p
    n, n_ref, steps = 10000, 128, 50
    rng = np.random.default_rng(7)
    goals = np.stack([rng.uniform(-0.15, 0.15, n), rng.uniform(-0.15, 0.15, n), 0.02 + rng.uniform(0.0, 0.2, n)], -1)
    objs = np.stack([rng.uniform(-0.15, 0.15, n), rng.uniform(-0.15, 0.15, n), np.full(n, 0.02)], -1)
    env = p.PandaVecEnv("pick_and_place", n, control_type="ee", auto_reset=False)
    obs = env.reset(goals=goals, object_positions=objs)
    g = torch.from_numpy(goals.astype(np.float32)).cuda()
    phase = torch.zeros(n, dtype=torch.int64, device="cuda"); count = torch.zeros_like(phase)
    done = torch.zeros(n, dtype=torch.bool, device="cuda")
    for t in range(steps):
        a = scripted_pick_and_place(torch, obs["observation"], g, phase, count)
        obs, rew, term, trunc, _ = env.step(a)
        done |= term.bool()
    gpu_success = done.cpu().numpy()
    final_obj = obs["observation"][:n_ref, 7:10].cpu().numpy()
    env.close()
    rate = gpu_success.mean()
    ref_success = np.zeros(n_ref, bool)
    for i in range(n_ref):
        oe = OracleEnv("pick_and_place", "ee")
        o, ag, dg = oe.reset(goals[i], objs[i])
        ph, cnt = np.zeros(1, np.int64), np.zeros(1, np.int64)
        for t in range(steps):
            a = scripted_pick_and_place(np, o[None].astype(np.float32), goals[i][None].astype(np.float32), ph, cnt)[0].astype(np.float32)
            o, ag, dg, r, s = oe.step(a)
            ref_success[i] |= s
        oe.close()
    agree = (ref_success == gpu_success[:n_ref]).mean()
    print(f"scripted success: gpu {rate:.4f} over {n} episodes; oracle {ref_success.mean():.4f} vs gpu {gpu_success[:n_ref].mean():.4f} on the same {n_ref}; per-episode agreement {agree:.4f}")
    assert rate > 0.9, rate                                   # the grasp actually works (finger contact + friction carry the cube)
    assert abs(ref_success.mean() - gpu_success[:n_ref].mean()) <= 0.01 + 1.0 / n_ref and agree >= 0.97
