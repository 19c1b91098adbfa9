"""CPU baseline beside tools/bench_aux.py's TTA-merge timing (test infrastructure: it runs the oracle, the
float32-numpy port of engine/bbox_aug_vote.py).  python tests/tta_cpu_baseline.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vote_oracle

g = torch.Generator().manual_seed(9)
n_obj, n_aug = 100, 14
ctr = torch.rand((n_obj, 2), generator=g) * 900
size = 40 + torch.rand((n_obj, 2), generator=g) * 200
lab = torch.randint(1, 21, (n_obj,), generator=g)
which = torch.arange(n_obj).repeat(n_aug)
xy = ctr[which] + torch.randn((n_obj * n_aug, 2), generator=g) * 4
wh = size[which] * (1 + 0.05 * torch.randn((n_obj * n_aug, 2), generator=g))
tb = torch.cat([xy, xy + wh], 1)
ts = torch.randperm(n_obj * n_aug, generator=g).float() / (n_obj * n_aug) * 0.9 + 0.06
tl = lab[which]
for kind in ("vote", "soft-vote"):
    t0 = time.perf_counter()
    for _ in range(5):
        vote_oracle.merge_multi_scale(tb.numpy(), ts.numpy(), tl.numpy(), 81, merge_type=kind, vote_thresh=0.66)
    print(kind, "%.2f ms per image (numpy port, 1 core)" % ((time.perf_counter() - t0) / 5 * 1e3))
