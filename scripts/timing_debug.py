"""Timing introspection (run under gpurun with PG_DEBUG_TIMING=1): distribution of per-warp cycles in the last launch."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import panda_lang_manip_b200 as p
task, ctrl, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
env = p.PandaVecEnv(task, n, control_type=ctrl)
g = torch.Generator(device='cuda'); g.manual_seed(0)
st = env.get_state(); st[:, -1] = torch.randint(0, env.max_episode_steps, (n,), device='cuda', generator=g).to(st.dtype); env.set_state(st)     # steady state: random episode phases
for t in range(int(sys.argv[4]) if len(sys.argv) > 4 else 70):
    a = torch.rand((n, env.action_dim), device='cuda', generator=g) * 2 - 1
    env.step(a)
torch.cuda.synchronize()
buf = np.zeros((n, 2), np.int64)
rc = env.lib.pg_debug_timing(env._h, buf.ctypes.data); assert rc == 0, rc
cyc = buf[:, 0].reshape(-1, 32); key = (buf[:, 1] & 0xffff).reshape(-1, 32); smid = (buf[:, 1] >> 16).reshape(-1, 32)[:, 0]
w = cyc.max(1)
nn = key & 31; rob = (key >> 5) & 1; cap = (key >> 6) & 1
print('warps', len(w), 'cycles: mean %.0f median %.0f p90 %.0f p99 %.0f max %.0f' % (w.mean(), np.median(w), np.percentile(w, 90), np.percentile(w, 99), w.max()))
order = np.argsort(-w)
print('slowest warps: cycles, sm, max n, #robot lanes, #capped lanes, n histogram')
for k in order[:12]:
    print(w[k], smid[k], nn[k].max(), rob[k].sum(), cap[k].sum(), np.bincount(nn[k], minlength=11))
print('by warp class (robot lanes > 0 / capped lanes > 16): mean cycles')
for r in (0, 1):
    for c in (0, 1):
        m = ((rob.sum(1) > 0) == r) & ((cap.sum(1) > 16) == c)
        if m.any(): print(' robot', r, 'capped', c, 'warps', m.sum(), 'mean cycles %.0f' % w[m].mean(), 'max %.0f' % w[m].max())
# per-SM load: sum of warp cycles
sm_tot = np.bincount(smid, weights=w, minlength=148)
print('per-SM sum of warp cycles: mean %.0f max %.0f;  per-SM max warp: mean %.0f' % (sm_tot.mean(), sm_tot.max(), np.mean([w[smid == s].max() for s in np.unique(smid)])))
