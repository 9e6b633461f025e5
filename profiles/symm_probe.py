"""Probe torch symmetric memory on this stack: allocate, rendezvous, peer pointers, a peer write from a torch op."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ctl = dist.new_group(backend="gloo")
import torch.distributed._symmetric_memory as symm_mem
print(rank, "symm_mem api:", [n for n in dir(symm_mem) if not n.startswith("_")][:40], flush=True)
try:
    t = symm_mem.empty(1024, dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    try:
        hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    except Exception as e:
        print(rank, "rendezvous(group) failed:", repr(e)[:200], flush=True)
        symm_mem.enable_symm_mem_for_group(dist.group.WORLD.group_name)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "handle:", type(hdl).__name__, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
    peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
    torch.cuda.synchronize(); dist.barrier(group=ctl)
    peer[:4] = 100.0 + rank          # store into the peer's memory over NVLink
    torch.cuda.synchronize(); dist.barrier(group=ctl)
    print(rank, "local after peer write:", t[:6].tolist(), flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
dist.barrier(group=ctl)
sys.stdout.flush(); os._exit(0)
