"""Probe: fused peer-write sharded GEMV vs the NCCL all-gather baseline at world N (correctness + time)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), ROOT]
import torch, torch.distributed as dist
import quant_gemm, bench_detail
from quant_gemm import sharded
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ctl = dist.new_group(backend="gloo")
shapes = [(4096, 4096), (11008, 4096), (4096, 11008)] * 4
T = int(os.environ.get("TOK", "1"))
ws = [bench_detail.make_weights(torch, 2, F, K, 1, dev, seed=100 * i + rank)[0] for i, (F, K) in enumerate(shapes)]
aqs = {K: quant_gemm.quantize_q8_1(torch.randn((T, K), device=dev, generator=torch.Generator(device=dev).manual_seed(K))) for K in (4096, 11008)}
# baseline
base = [sharded.ShardedGemm(w, F * world, K, 2, flags=0x10) for w, (F, K) in zip(ws, shapes)]
ref = [op(aqs[K]).clone() for op, (F, K) in zip(base, shapes)]
torch.cuda.synchronize()
plan = sharded.PeerPlan(sum(F * world * T for F, K in shapes), len(shapes), dev, ctl_group=ctl)
ops = [sharded.ShardedGemvP2P(w, F * world, K, 2, T, plan, flags=0x10) for w, (F, K) in zip(ws, shapes)]
def step():
    for op, (F, K) in zip(ops, shapes):
        op(aqs[K])
    plan.end_step()
for _ in range(3):
    step()
torch.cuda.synchronize(); dist.barrier(group=ctl)
ok = all(torch.equal(op.out, r) for op, r in zip(ops, ref)) if not os.environ.get("QGEMM_PEER_DBG") else None
print(rank, "p2p == nccl:", ok, flush=True)
stream = torch.cuda.Stream(device=dev)
def timeit(fn, n=20):
    with torch.cuda.stream(stream):
        g = torch.cuda.CUDAGraph()
        fn(); stream.synchronize()
        with torch.cuda.graph(g, stream=stream):
            fn()
        for _ in range(3): g.replay()
        stream.synchronize(); dist.barrier(group=ctl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n): g.replay()
        e1.record(stream); stream.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n / len(shapes)
def base_step():
    for op, (F, K) in zip(base, shapes):
        op(aqs[K])
t_p2p = timeit(step)
ok2 = None
t_nccl = timeit(base_step)
print(rank, f"us per GEMV: fused peer-write {t_p2p:.2f}  nccl all-gather {t_nccl:.2f}  still equal {ok2}", flush=True)
dist.barrier(group=ctl)
sys.stdout.flush(); os._exit(0)
