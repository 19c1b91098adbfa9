"""Calibration: how fast does a plain device copy of the bench's logit tensors run (same bytes as
final_loss_kernel), with and without the L2 flush in front?  Measurement aid, not part of the product."""
import torch
torch.cuda.set_device(0)
dev = torch.device("cuda")
xs = [torch.randn(16, 80, h, w, device=dev) for (h, w) in ((100, 168), (50, 84), (25, 42), (13, 21), (7, 11))]
gs = [torch.empty_like(x) for x in xs]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
big = torch.randn(16 * 80 * 22400, device=dev)
gbig = torch.empty_like(big)
nbytes = sum(x.numel() for x in xs) * 8


def timed(fn, flush_first, reps=20):
    ts = []
    for _ in range(reps):
        if flush_first:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1000)
    ts.sort()
    return ts[len(ts) // 2]


def per_level():
    for x, g in zip(xs, gs):
        g.copy_(x)


for name, fn in (("5 level copies", per_level), ("one flat copy", lambda: gbig.copy_(big)),
                 ("flat read-only sum", lambda: big.sum()), ("flat write-only fill", lambda: gbig.fill_(1.0))):
    for fl in (False, True):
        us = timed(fn, fl)
        b = nbytes if "copy" in name or "copies" in name else nbytes / 2
        print("%-22s flush=%-5s %8.1f us  %7.0f GB/s" % (name, fl, us, b / us / 1e3))
# flush with a READ instead of a write (leaves clean lines in L2)
rbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def rflush():
    return rbuf.sum()
for name, fn in (("one flat copy", lambda: gbig.copy_(big)), ("flat read-only sum", lambda: big.sum())):
    ts = []
    for _ in range(20):
        rflush()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1000)
    ts.sort(); us = ts[10]
    b = nbytes if "copy" in name else nbytes / 2
    print("%-22s readflush   %8.1f us  %7.0f GB/s" % (name, us, b / us / 1e3))
