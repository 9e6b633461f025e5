"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank builds the same seeded inputs, takes its row shard of the weights and runs the N-sharded operators of
quant_gemm.sharded; the gathered C on EVERY rank is compared with the CPU oracle (oracle/qgemm_oracle.c) on sampled
rows -- not with another run of our own kernels:
  * ShardedGemm                  local kernels + NCCL all-gather (the baseline)
  * ShardedGemmP2P, unicast      all-gather fused into the kernels' epilogues, one NVLink peer store per rank
  * ShardedGemmP2P, multicast    same through the NVLS multicast mapping (skipped where the fabric offers none)
for a decode shape (T = 1, 4) and a prefill shape (tcgen05 path).  Prints one OK line per case; exit code 0 iff all pass.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import datagen  # noqa: E402
import qgemm_oracle as qo  # noqa: E402


def main() -> int:
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctl = dist.new_group(backend="gloo")
    import quant_gemm
    from quant_gemm import sharded

    O = qo.Oracle()
    failures = 0
    cases = [(qo.Q4_0, 1, 4096, 4096), (qo.Q5_1, 4, 2048, 4096), (qo.Q8_0, 8, 1024, 2048), (qo.Q4_0, 256, 2048, 2048),
             (qo.Q5_0, 128, 1000, 1024)]
    nlaunch = 2 * len(cases)
    plans = {}
    for mc in (False, True):
        if not mc:
            os.environ["QGEMM_NO_MULTICAST"] = "1"
        else:
            os.environ.pop("QGEMM_NO_MULTICAST", None)
        plans[mc] = sharded.PeerPlan(sum(F * T for _, T, F, _ in cases) + 4096, len(cases), dev, ctl_group=ctl, multicast=mc)
    have_mc = bool(plans[True].mc_ptr)
    ops = {False: [], True: []}
    data = []
    for wt, T, F, K in cases:
        x, w = datagen.model_like(T, F, K, seed=1000 + F + T)     # same on every rank
        aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
        rows = np.unique(np.r_[0:4, F // 2 - 2:F // 2 + 2, F - 4:F, np.random.default_rng(F).integers(0, F, 24)])
        ref = O.gemm(wt, aq, wq[rows], layout="FT")
        dq = torch.from_numpy(wq).to(dev)
        shard = sharded.shard_weight(dq, world, rank).contiguous()
        data.append((wt, T, F, K, torch.from_numpy(aq).to(dev), shard, rows, ref))
        for mc in (False, True):
            ops[mc].append(sharded.ShardedGemmP2P(shard, F, K, wt, T, plans[mc]))

    def check(name, c, rows, ref, wt, T, F, K):
        nonlocal failures
        got = c[torch.from_numpy(rows).to(dev)].cpu().numpy()
        e = qo.max_norm_err(got, ref)
        ok = bool(np.isfinite(got).all()) and e <= 1e-5
        failures += 0 if ok else 1
        print(f"rank {rank}: {name} {qo.TYPE_NAMES[wt]} T={T} F={F} K={K}: max-norm err {e:.2e} {'OK' if ok else 'FAIL'}", flush=True)

    # baseline: NCCL all-gather
    for wt, T, F, K, daq, shard, rows, ref in data:
        c = sharded.ShardedGemm(shard, F, K, wt)(daq)
        torch.cuda.synchronize()
        check("nccl", c, rows, ref, wt, T, F, K)
    # fused: unicast peer stores, then NVLS multicast
    for mc in (False, True):
        if mc and not have_mc:
            if rank == 0:
                print("multicast mapping not available on this fabric: skipped", flush=True)
            continue
        plan = plans[mc]
        for step in range(2):   # two steps: the counters must also work the second time round
            outs = [op(d[4]) for op, d in zip(ops[mc], data)]
            plan.end_step()
            torch.cuda.synchronize()
            dist.barrier(group=ctl)
            for out, (wt, T, F, K, daq, shard, rows, ref) in zip(outs, data):
                check(f"fused-{'multicast' if mc else 'unicast'} step {step}", out, rows, ref, wt, T, F, K)
            dist.barrier(group=ctl)
    t = torch.tensor([failures], device=dev)
    dist.all_reduce(t)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"mgpu_check world={world}: {'ALL OK' if int(t.item()) == 0 else str(int(t.item())) + ' FAILURES'}", flush=True)
    dist.destroy_process_group()
    return 1 if int(t.item()) else 0


if __name__ == "__main__":
    sys.exit(main())
