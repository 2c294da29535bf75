"""render kernel alone, for ncu: 256 PickAndPlace envs x 480 x 480"""
import sys, torch
sys.path.insert(0, '.')
import panda_lang_manip_b200 as p
env = p.PandaVecEnv("pick_and_place", 256, control_type="ee")
for _ in range(3):
    out = env.render(width=480, height=480)
torch.cuda.synchronize()
print({k: tuple(v.shape) for k, v in out.items()})
