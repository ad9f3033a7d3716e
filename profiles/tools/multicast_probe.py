"""Does this box give torch's symmetric memory an NVSwitch multicast mapping (needed for multimem.ld_reduce / multimem.st)?
torchrun --nproc-per-node 2 profiles/tools/multicast_probe.py"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
t = symm.empty(1 << 20, dtype=torch.float32, device=f"cuda:{local}")
t.fill_(rank + 1.0)
h = symm.rendezvous(t, dist.group.WORLD)
print(f"rank {rank}: multicast_ptr={getattr(h, 'multicast_ptr', None)} buffer_ptrs={[hex(p) for p in h.buffer_ptrs][:2]} signal_pad_ptrs={len(h.signal_pad_ptrs)}", flush=True)
dist.barrier()
try:
    torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
    torch.cuda.synchronize()
    print(f"rank {rank}: multimem_all_reduce_ ok, t[0]={float(t[0])} (expect {world * (world + 1) / 2})", flush=True)
    for n in (1 << 19, 1 << 16):
        v = t[:n]
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            torch.ops.symm_mem.multimem_all_reduce_(v, "sum", dist.group.WORLD.group_name)
        e1.record(); torch.cuda.synchronize()
        print(f"rank {rank}: multimem_all_reduce_ {n * 4 / 1e6:.2f} MB: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
        g = torch.zeros(n, device=f"cuda:{local}")
        e0.record()
        for _ in range(50):
            dist.all_reduce(g)
        e1.record(); torch.cuda.synchronize()
        print(f"rank {rank}: nccl all_reduce {n * 4 / 1e6:.2f} MB: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us", flush=True)
except Exception as ex:
    print(f"rank {rank}: multimem_all_reduce_ failed: {ex!r}", flush=True)
dist.destroy_process_group()
