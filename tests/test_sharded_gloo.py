"""World-size-2 (and 3, uneven) CPU tests of the N-sharded path's host logic over gloo: row ranges,
in-place placement of each rank's slice, all-gather assembly == the unsharded oracle result.
The local GEMM is injected (oracle) because the product path has no CPU implementation."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cases, T, K, wt, q):
    import sys
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "llama.cpp-quant-gemm_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import datagen
        import qgemm_oracle as qo
        from quant_gemm import sharded
        O = qo.Oracle()

        def oracle_gemm(weight_q, activation_q, Fr, T_, K_, wtype, flags, out):
            c = O.gemm(wtype, activation_q.numpy(), weight_q.numpy(), layout="FT", flags=flags)
            out.copy_(torch.from_numpy(c))
            return out

        results = []
        for F, align in cases:
            x, w = datagen.model_like(T, F, K, seed=5)      # same seed on every rank: replicated inputs
            aq, wq = O.quantize_q8_1(x), O.quantize_weight(wt, w)
            shard = sharded.shard_weight(torch.from_numpy(wq), world, rank, align)
            op = sharded.ShardedGemm(shard, F, K, wt, align=align, gemm_fn=oracle_gemm)
            assert op.ranges[0][0] == 0 and op.ranges[-1][1] == F
            out = op(torch.from_numpy(aq))
            ref = O.gemm(wt, aq, wq, layout="FT")
            results.append((bool((out.numpy().view(np.uint32) == ref.view(np.uint32)).all()), op.ranges, op.even))
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


def test_sharded_gemm_gathers_to_unsharded_result():
    world = 2
    cases = [(512, 128), (300, 64), (130, 128)]   # even split, uneven tail, one rank left empty-ish
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, 3, 256, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i, (F, align) in enumerate(cases):
        oks = [res[r][i][0] for r in range(world)]
        assert all(oks), (F, align)
        ranges = res[0][i][1]
        assert res[1][i][1] == ranges
        sizes = [b - a for a, b in ranges]
        assert sum(sizes) == F and all(s % align == 0 for s in sizes[:-1])
    assert res[0][0][2] is True and res[0][1][2] is False
