"""Probe: does torch symmetric memory (CUDA backend, P2P over NVLink) work on this box?
torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
t = symm_mem.empty((world, 1024), dtype=torch.float32, device=dev)
t.fill_(-1)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "backend", symm_mem.get_backend(dev), "ptrs", [hex(p) for p in hdl.buffer_ptrs], "dev table", hex(hdl.buffer_ptrs_dev),
      "multicast", hdl.has_multicast_support, hex(hdl.multicast_ptr) if hdl.has_multicast_support else None,
      "signal pad", hdl.signal_pad_size, flush=True)
hdl.barrier(channel=0)
for p in range(world):
    peer = hdl.get_buffer(p, (world, 1024), torch.float32)
    peer[rank].fill_(float(rank))
hdl.barrier(channel=0)
torch.cuda.synchronize()
ok = all(float(t[p].min()) == p and float(t[p].max()) == p for p in range(world))
# barrier latency
for _ in range(5):
    hdl.barrier(channel=0)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    hdl.barrier(channel=0)
e.record()
torch.cuda.synchronize()
print(rank, "peer writes ok:", ok, "barrier us:", a.elapsed_time(e) / 50 * 1e3, flush=True)
dist.destroy_process_group()
