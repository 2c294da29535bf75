"""HER kernels alone, for ncu (scripts/prof_r2.sh): compute_reward on 32 M x 3-D and 1 M x 6-D rows, her_relabel on 1 M of 16 M x 6-D rows with
random indices, episode-local goals and an index-sorted batch."""
import sys, torch
sys.path.insert(0, '.')
import panda_lang_manip_b200 as p
dev = torch.device('cuda')
ag = torch.rand((1 << 25, 3), device=dev); dg = torch.rand((1 << 25, 3), device=dev)
for _ in range(2): p.compute_reward('reach', 'sparse', ag, dg)
del ag, dg
ag = torch.rand((1 << 20, 6), device=dev); dg = torch.rand((1 << 20, 6), device=dev)
for _ in range(2): p.compute_reward('stack', 'sparse', ag, dg)
R, M, G = 1 << 24, 1 << 20, 6
nag = torch.rand((R, G), device=dev); dgb = torch.rand((R, G), device=dev)
src = torch.randint(0, R, (M,), device=dev); gs = torch.where(torch.rand(M, device=dev) < 0.8, torch.randint(0, R, (M,), device=dev), torch.full((M,), -1, device=dev))
for _ in range(2): p.her_relabel('stack', 'sparse', nag, dgb, src, gs)
for srt in (False, True):
    s2, g2 = p.her_sample_indices(R, M, 100, 0.8, device=dev, sort=srt)
    for _ in range(2): p.her_relabel('stack', 'sparse', nag, dgb, s2, g2)
torch.cuda.synchronize()
