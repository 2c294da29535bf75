"""Scheduling introspection (run under gpurun): how well does the key written by launch k predict the work of launch k+1?"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import panda_lang_manip_b200 as p
task, ctrl, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
env = p.PandaVecEnv(task, n, control_type=ctrl)
g = torch.Generator(device='cuda'); g.manual_seed(0)
for t in range(17):
    a = torch.rand((n, env.action_dim), device='cuda', generator=g) * 2 - 1
    env.step(a)
torch.cuda.synchronize()
key = np.zeros(n, np.uint16); perm = np.zeros(n, np.int32)
env.lib.pg_debug_schedule(env._h, key.ctypes.data, perm.ctypes.data)
def bucket(k):      # perm_bucket of csrc/panda_kernels.cuh
    nn = k & 31; robot = (k >> 5) & 1; capped = (k >> 6) & 1; near = (k >> 7) & 1; full = (k >> 9) & 1; ngen = (k >> 10) & 15
    nq = np.where(nn <= 10, nn, 11 + np.minimum((nn - 11) >> 2, 2))
    cls = np.where(ngen > 0, 3 + np.minimum((ngen - 1) >> 1, 2), np.where(robot == 1, 2, near))
    return ((full * 6 + cls) * 14 + nq) * 2 + capped
assert np.array_equal(np.sort(perm), np.arange(n)), "perm is not a permutation"
k = key[perm].astype(np.int64)          # key after the launch, in thread order of that launch
b = bucket(k)
print('bucket histogram (bucket: count):', {int(x): int(c) for x, c in zip(*np.unique(b, return_counts=True))})
w = k.reshape(-1, 32)
cap = (w >> 6) & 1; rob = (w >> 5) & 1; nn = w & 31
print('capped frac', cap.mean(), ' warps with any capped', (cap.max(1) > 0).mean(), ' ideal', np.ceil(cap.sum() / 32) / len(w))
print('robot frac', rob.mean(), ' warps with any robot', (rob.max(1) > 0).mean(), ' ideal', np.ceil(rob.sum() / 32) / len(w))
print('mean n', nn.mean(), ' mean warp-max n', nn.max(1).mean())
print('fraction of threads whose bucket order is monotone:', np.mean(np.diff(b) <= 0))
