import sys, json, torch
sys.path.insert(0, '/root/repo')
from bench_post import measure_post
torch.cuda.set_device(0)
for n in (8, 16, 32, 64):
    d = measure_post(torch.device("cuda", 0), steps=6, warmup=3, images=n, with_cpu=False, with_dense=False)
    print(n, round(d["value"]), round(d["ms_per_step"], 4), {k: round(v, 1) for k, v in d["per_kernel_us"].items()})
