"""CPU statistics of the solver's per-sub-step work (kernel math compiled for the host, tests/hostcheck): share of env sub-steps with
robot contacts, contact counts, sweeps.  usage: python scripts/host_trace.py <task> <ee|joints> [envs] [steps]"""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
d = os.path.join(ROOT, "tests", "hostcheck"); so = os.path.join(d, "libhostcheck.so")
subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(d, "hostcheck.cpp")], check=True)
hc = ctypes.CDLL(so); vp, ci = ctypes.c_void_p, ctypes.c_int
hc.hc_env_step.argtypes = [ci, ci, ci, ci] + [vp] * 8
hc.hc_dbg_trace.argtypes = [vp]
P = lambda a: a.ctypes.data
task, ctrl = sys.argv[1], sys.argv[2]; nenv = int(sys.argv[3]) if len(sys.argv) > 3 else 64; nsteps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
TASK = {"reach": 0, "push": 1, "slide": 2, "pick_and_place": 3, "stack": 4, "flip": 5}[task]
OBS = [6, 18, 18, 19, 31, 20][TASK]; G = [3, 3, 3, 3, 6, 4][TASK]; NOBJ = [0, 1, 1, 1, 2, 1][TASK]
A = (3 if ctrl == "ee" else 7) + (0 if TASK in (0, 1, 2) else 1); MAXS = 100 if TASK == 4 else 50
BASE = np.array([-0.6, 0.0, 0.0]); NEUTRAL = np.array([0.0, 0.41, 0.0, -1.85, 0.0, 2.26, 0.79, 0.0, 0.0])
rng = np.random.default_rng(0)
def fresh():
    st = np.zeros(50); st[:9] = NEUTRAL
    z0 = 0.03 if TASK == 2 else 0.02
    for o in range(NOBJ):
        st[18 + 13 * o:18 + 13 * o + 3] = [rng.uniform(-0.15, 0.15), rng.uniform(-0.15, 0.15), z0 if o == 0 else 0.06]; st[18 + 13 * o + 6] = 1
    g = np.zeros(6); g[:3] = rng.uniform([-0.15, -0.15, 0.0], [0.15, 0.15, 0.2]); 
    if TASK == 5: g[:4] = [0, 0, 0, 1]
    if TASK == 4: g[3:] = g[:3] + [0, 0, 0.04]
    st[44:] = g
    return st
tr = []; buf = np.zeros(4096, np.int32)
for e in range(nenv):
    st = fresh(); age = int(rng.integers(0, MAXS))
    for t in range(nsteps):
        a = rng.uniform(-1, 1, A).astype(np.float32)
        o2, a2, d2 = np.zeros(OBS, np.float32), np.zeros(G, np.float32), np.zeros(G, np.float32); r2, s2 = np.zeros(1, np.float32), np.zeros(1, np.uint8)
        hc.hc_env_step(0, TASK, 0 if ctrl == "ee" else 1, 0, P(BASE), P(st), P(a), P(o2), P(a2), P(d2), P(r2), P(s2))
        n = hc.hc_dbg_trace(P(buf))
        if t >= 10: tr.append(buf[:n].copy())
        age += 1
        if s2[0] or age >= MAXS: st = fresh(); age = 0
tr = np.concatenate(tr); it = tr & 255; nc = (tr >> 8) & 255; nr = (tr >> 16) & 255
print(f"{task}/{ctrl}: {len(tr)} sub-steps; robot-contact share {np.mean(nr > 0):.3f}; any-contact share {np.mean(nc > 0):.3f}; capped share {np.mean(it >= 49):.3f}; mean sweeps {it.mean() + 1:.1f}")
for name, m in (("nr>0", nr > 0), ("nr==0,nc>0", (nr == 0) & (nc > 0)), ("nc==0", nc == 0)):
    if m.any(): print(f"  {name}: share {m.mean():.3f} mean contacts {nc[m].mean():.1f} (robot {nr[m].mean():.1f}) mean sweeps {it[m].mean() + 1:.1f} capped {np.mean(it[m] >= 49):.2f}")
print("  nc histogram:", np.bincount(nc, minlength=11)[:24])
hc.hc_dbg_fallbacks.restype = ctypes.c_long; hc.hc_dbg_full_starts.restype = ctypes.c_long
print("  watched-sweep fallbacks", hc.hc_dbg_fallbacks(), "solves started with the full sweep", hc.hc_dbg_full_starts())
