"""Probe: sharded GEMV + NCCL all-gather at world 2, eager then graph-captured (bounded by the caller's timeout)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, torch.distributed as dist
import quant_gemm, bench_detail
from quant_gemm import sharded
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
t0 = time.time()
dist.init_process_group("nccl", device_id=dev)
ctl = dist.new_group(backend="gloo")  # eager control-plane collectives stay off the NCCL communicator used inside graphs
print(rank, "init", time.time() - t0, flush=True)
F, K = 4096, 4096
w = bench_detail.make_weights(torch, 2, F, K, 1, dev, seed=rank)[0]
aq = quant_gemm.quantize_q8_1(torch.randn((1, K), device=dev, generator=torch.Generator(device=dev).manual_seed(1)))
op = sharded.ShardedGemm(w, F * world, K, 2, flags=int(os.environ.get("FLAGS", "0x10"), 0))
out = torch.zeros((F * world, 1), device=dev)
op(aq, out=out); torch.cuda.synchronize()
print(rank, "eager ok", float(out.abs().sum()), flush=True)
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    op(aq, out=out); stream.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for _ in range(4):
            op(aq, out=out)
    print(rank, "captured", flush=True)
    for _ in range(3):
        g.replay()
    stream.synchronize()
print(rank, "graph ok", float(out.abs().sum()), flush=True)
dist.barrier(group=ctl); torch.cuda.synchronize()
print(rank, "done", flush=True)
sys.stdout.flush()
os._exit(0)
