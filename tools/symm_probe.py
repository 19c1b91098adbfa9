"""Feasibility probe: does torch symmetric memory (peer-mapped buffers over NVLink) work on this box?
torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(64, dtype=torch.float64, device=dev)
    t.zero_()
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name)
    print(rank, "rendezvous ok; buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
    peer = hdl.get_buffer((rank + 1) % world, (64,), torch.float64)
    peer[rank] = float(rank + 1)          # store into the peer's memory
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    print(rank, "my buffer after peer writes:", t[:world].tolist(), flush=True)
    print(rank, "can_access_peer", torch.cuda.can_device_access_peer(lr, (lr + 1) % world))
except Exception as e:  # noqa
    import traceback; traceback.print_exc()
    print(rank, "SYMM FAILED", repr(e)[:300], flush=True)
dist.barrier()
os._exit(0)
