"""Small run of every kernel (a smoke script; compute-sanitizer is closed on the GPU pool): step (identity map + sorted path), reset, snapshots, IK, link state,
bare world, rewards, HER relabel, render."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import panda_lang_manip_b200 as p
g = torch.Generator(device='cuda'); g.manual_seed(0)
for task, ctrl, n in (("reach", "joints", 300), ("reach", "ee", 4100), ("pick_and_place", "ee", 4100), ("stack", "ee", 200), ("slide", "joints", 100)):
    env = p.PandaVecEnv(task, n, control_type=ctrl)
    for t in range(3):
        env.step(torch.rand((n, env.action_dim), device='cuda', generator=g) * 2 - 1)
    sid = env.save_state(); env.step(torch.zeros((n, env.action_dim), device='cuda')); env.restore_state(sid); env.remove_state(sid)
    st = env.get_state(); env.set_state(st)
    env.reset()
    env.close()
ag = torch.rand((10007, 3), device='cuda'); dg = torch.rand((10007, 3), device='cuda')
p.compute_reward("reach", "sparse", ag, dg); p.is_success("reach", ag, dg)
src, gs = p.her_sample_indices(10007, 5000, 50)
p.her_relabel("push", "dense", ag, dg, src, gs, return_achieved=True)
pad = torch.rand((10007, 8), device='cuda')
p.her_relabel("stack", "sparse", pad[:, :6], pad[:, :6].clone(), src, gs)
w = p.PandaBareWorld(64, robot_base=(0.0, 0.0, 0.0), bodies=[{"shape": "box", "half_extents": (0.5, 0.5, 0.5), "mass": 1.0, "position": (0, 0, 5.0), "orientation": (0, 0, 0, 1), "lateral_friction": 0.5}], ground_z=0.0)
w.step(20); w.link_state(6); w.render(width=64, height=48)
w.close()
env = p.PandaVecEnv("push", 8)
env.render(width=64, height=48, segmentation=True)
env.close()
torch.cuda.synchronize()
print("all kernels ran")
