"""N steady-state steps of one configuration (for ncu launch lists): python scripts/steps.py <task> <ee|joints> <envs> <preroll> <steps>"""
import sys, torch
sys.path.insert(0, '.')
import panda_lang_manip_b200 as p
task, ctrl, n, pre, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
env = p.PandaVecEnv(task, n, control_type=ctrl)
g = torch.Generator(device='cuda'); g.manual_seed(0)
st = env.get_state(); st[:, -1] = torch.randint(0, env.max_episode_steps, (n,), device='cuda', generator=g).to(st.dtype); env.set_state(st)
acts = [torch.rand((n, env.action_dim), device='cuda', generator=g) * 2 - 1 for _ in range(8)]
for t in range(pre + steps):
    env.step(acts[t % 8])
torch.cuda.synchronize()
print("launches", p.kernel_launches())
