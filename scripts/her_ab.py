"""her_relabel timing for the current PG_HER_INFLIGHT: 1 M of 16 M x 6-D rows, random / episode-local unsorted / index-sorted"""
import os, sys, torch
sys.path.insert(0, '.')
import panda_lang_manip_b200 as p
dev = torch.device('cuda'); R, M, G = 1 << 24, 1 << 20, 6
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nag = torch.rand((R, G), device=dev); dgb = torch.rand((R, G), device=dev)
src = torch.randint(0, R, (M,), device=dev); gs = torch.where(torch.rand(M, device=dev) < 0.8, torch.randint(0, R, (M,), device=dev), torch.full((M,), -1, device=dev))
cases = {"random": (src, gs), "local_unsorted": p.her_sample_indices(R, M, 100, 0.8, device=dev, sort=False), "local_sorted": p.her_sample_indices(R, M, 100, 0.8, device=dev, sort=True)}
out = []
for name, (s, g) in cases.items():
    for _ in range(3): p.her_relabel("stack", "sparse", nag, dgb, s, g)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(15)]
    for a, b in evs:
        flush.zero_(); a.record(); p.her_relabel("stack", "sparse", nag, dgb, s, g); b.record()
    torch.cuda.synchronize()
    out.append("%s %.1f us" % (name, 1e3 * sorted(a.elapsed_time(b) for a, b in evs)[7]))
print("PG_HER_INFLIGHT=%s: " % os.environ.get("PG_HER_INFLIGHT", "default") + ", ".join(out))
