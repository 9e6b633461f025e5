"""N-sharded quantized GEMM across the GPUs of one node.

The output column block C[:, f0:f1] depends only on weight rows [f0, f1) and on all of the
activations, so weight rows are split into contiguous, tile-aligned ranges (the split itself is
qgemm_shard_range() of the C ABI), every rank runs the single-GPU kernels on its slice, and
the slices are combined with an all-gather over NCCL/NVLink.  In the [F, T] (ggml / python)
output orientation each rank's slice is one contiguous chunk of the gathered tensor, so the
local GEMM writes straight into its final position and the all-gather runs in place.

One process per GPU (torchrun); torch.distributed is plumbing only.  The reference has no
multi-GPU code: this module is new functionality with no reference counterpart.
"""
from __future__ import annotations

import os

from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import _lib

DEFAULT_ALIGN = 128  # rows: one tensor-core tile (kBN) so no rank gets a partial tile except the last


def shard_rows(F: int, world: int, rank: int, align: int = DEFAULT_ALIGN) -> tuple[int, int]:
    """[f0, f1) of `rank` -- contiguous, multiples of `align` except possibly the tail."""
    return _lib.shard_range(F, world, rank, align)


def shard_weight(weight_q: torch.Tensor, world: int, rank: int, align: int = DEFAULT_ALIGN) -> torch.Tensor:
    """Rows of a full [F, K/32, bytes] weight tensor owned by `rank` (a view)."""
    f0, f1 = shard_rows(weight_q.shape[0], world, rank, align)
    return weight_q[f0:f1]


class ShardedGemm:
    """C[F_total, T] = W[F_total, K] . A[T, K]^T with W row-sharded over `group`.

    weight_shard: this rank's rows [f0:f1) as uint8 [F_r, K/32, block bytes] (device tensor).
    gemm_fn(weight_q, activation_q, F_r, T, K, wtype, flags, out) -> out computes the local slice;
    it defaults to quant_gemm.gemm (the CUDA path).  Tests on CPU (gloo) inject the oracle here --
    the product default never falls back.
    """

    def __init__(self, weight_shard: torch.Tensor, F_total: int, K: int, wtype: int,
                 group: Optional[dist.ProcessGroup] = None, align: int = DEFAULT_ALIGN,
                 gemm_fn: Optional[Callable] = None, flags: int = 0):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.F_total, self.K, self.wtype, self.flags = F_total, K, wtype, flags
        self.ranges = [shard_rows(F_total, self.world, r, align) for r in range(self.world)]
        self.f0, self.f1 = self.ranges[self.rank]
        assert weight_shard.shape[0] == self.f1 - self.f0, "weight shard does not match this rank's row range"
        self.weight = weight_shard
        self.even = len({b - a for a, b in self.ranges}) == 1
        if gemm_fn is None:
            from . import gemm as _gemm
            gemm_fn = _gemm
        self.gemm_fn = gemm_fn

    def local(self, activation_q: torch.Tensor, out_full: torch.Tensor) -> torch.Tensor:
        """Compute this rank's rows directly inside the gathered buffer."""
        T = activation_q.shape[0]
        mine = out_full[self.f0:self.f1]
        if self.f1 > self.f0:
            self.gemm_fn(self.weight, activation_q, self.f1 - self.f0, T, self.K, self.wtype, self.flags, out=mine)
        return mine

    def gather(self, out_full: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return out_full
        mine = out_full[self.f0:self.f1]
        if self.even:
            dist.all_gather_into_tensor(out_full, mine, group=self.group)  # in place: slice r sits at offset r
        else:  # uneven tail: one broadcast per owner (collectives need equal chunk sizes)
            for r, (a, b) in enumerate(self.ranges):
                if b > a:
                    dist.broadcast(out_full[a:b], src=dist.get_global_rank(self.group, r) if self.group else r,
                                   group=self.group)
        return out_full

    def __call__(self, activation_q: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        T = activation_q.shape[0]
        if out is None:
            out = torch.empty((self.F_total, T), dtype=torch.float32, device=activation_q.device)
        self.local(activation_q, out)
        return self.gather(out)


# ------------------------------------------------------------------------------------------
# Decode: all-gather fused into the GEMV kernel (peer stores over NVLink, device-side flags)
# ------------------------------------------------------------------------------------------
class PeerPlan:
    """Symmetric (peer-mapped) memory shared by the sharded decode GEMVs of one step.

    Holds one gathered-output pool and the arrival counters; every ShardedGemvP2P of the step takes
    a slice of the pool and a launch index.  torch symmetric memory provides the peer pointers; the
    kernels do the rest (include/qgemm.h, qgemm_gemm_peers)."""

    def __init__(self, pool_floats: int, launches_per_step: int, device: torch.device,
                 group: Optional[dist.ProcessGroup] = None, ctl_group: Optional[dist.ProcessGroup] = None,
                 multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        assert self.world <= 8
        self.lps = launches_per_step
        self.pool = symm_mem.empty(max(pool_floats, 1024), dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(1024, dtype=torch.int32, device=device)
        self.pool.zero_()
        self.flags.zero_()
        self.pool_hdl = symm_mem.rendezvous(self.pool, self.group)
        self.flag_hdl = symm_mem.rendezvous(self.flags, self.group)
        self.pool_ptrs = list(self.pool_hdl.buffer_ptrs)
        self.flag_ptrs = list(self.flag_hdl.buffer_ptrs)
        # NVLS: one multicast mapping of the pool bound on every rank -- a store to it lands everywhere
        self.mc_ptr = 0
        if multicast and not os.environ.get("QGEMM_NO_MULTICAST"):
            self.mc_ptr = int(getattr(self.pool_hdl, "multicast_ptr", 0) or 0)
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        self.cursor = 0
        self.next_index = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=ctl_group if ctl_group is not None else self.group)  # zeroed everywhere before any signal

    def alloc(self, numel: int) -> int:
        off = self.cursor
        self.cursor += (numel + 63) // 64 * 64
        assert self.cursor <= self.pool.numel(), "PeerPlan pool too small"
        return off

    def peers_struct(self, elem_offset: int, launch_index: int, wait_index: Optional[int] = None) -> "_lib.QgemmPeers":
        ps = _lib.QgemmPeers()
        ps.world, ps.rank = self.world, self.rank
        for r in range(self.world):
            ps.C[r] = self.pool_ptrs[r] + 4 * elem_offset
            ps.flag[r] = self.flag_ptrs[r]
        ps.done, ps.step = self.done.data_ptr(), self.step.data_ptr()
        ps.launches_per_step, ps.launch_index = self.lps, launch_index
        ps.wait_index = launch_index if wait_index is None else wait_index
        ps.C_multicast = (self.mc_ptr + 4 * elem_offset) if self.mc_ptr else None
        return ps

    def end_step(self) -> None:
        """Wait until every launch of the step has landed from every rank, then advance the step."""
        assert self.next_index == self.lps or self.next_index == 0
        L = _lib.lib()
        st = torch.cuda.current_stream(self.pool.device).cuda_stream
        ps = self.peers_struct(0, 0)
        _lib.raise_on_error(L.qgemm_peer_wait(ps, st), "peer_wait")
        _lib.raise_on_error(L.qgemm_peer_step_advance(self.step.data_ptr(), st), "peer_step_advance")


class ShardedGemvP2P:
    """GEMM over row-sharded weights whose kernel writes its slice of C[F_total, T] into every rank's
    gathered buffer: the decode kernels for T <= 8, the tcgen05 epilogue for larger T (scratch from the
    registered default workspace or the stream's pool).  `out` is this rank's full gathered view (valid
    after the next launch's prologue or PeerPlan.end_step())."""

    def __init__(self, weight_shard: torch.Tensor, F_total: int, K: int, wtype: int, T: int, plan: PeerPlan,
                 align: int = DEFAULT_ALIGN, flags: int = 0, wait_index: Optional[int] = None):
        """wait_index: how many launches of the step must have landed everywhere before this one reads
        its activations (default: all earlier ones).  A model passes the index after the launch that
        produced this GEMV's input, so independent projections (q/k/v, gate/up) do not re-synchronise."""
        self.plan, self.K, self.wtype, self.T, self.flags = plan, K, wtype, T, flags
        if T > 8:
            self.flags |= _lib.GEMM_STREAM_ALLOC
        self.ranges = [shard_rows(F_total, plan.world, r, align) for r in range(plan.world)]
        self.f0, self.f1 = self.ranges[plan.rank]
        assert weight_shard.shape[0] == self.f1 - self.f0 > 0, "every rank needs a non-empty shard in peer mode"
        self.weight = weight_shard.contiguous()
        self.offset = plan.alloc(F_total * T)
        self.out = plan.pool[self.offset:self.offset + F_total * T].view(F_total, T)
        self.launch_index = plan.next_index
        plan.next_index += 1
        assert plan.next_index <= plan.lps
        # this rank's rows start at element f0 * T of the [F_total, T] buffer on every rank
        self.ps = plan.peers_struct(self.offset + self.f0 * T, self.launch_index, wait_index)

    def __call__(self, activation_q: torch.Tensor) -> torch.Tensor:
        L = _lib.lib()
        rc = L.qgemm_gemm_peers(self.wtype, activation_q.data_ptr(), self.weight.data_ptr(), self.ps, self.T,
                                self.f1 - self.f0, self.K, 1, self.T, self.flags,
                                torch.cuda.current_stream(self.weight.device).cuda_stream)
        _lib.raise_on_error(rc, "gemm_peers")
        return self.out


ShardedGemmP2P = ShardedGemvP2P   # same operator at prefill sizes


class ShardedGemvGroupP2P:
    """Grouped form of ShardedGemvP2P: several row-sharded matrices that share their activations
    (fused q/k/v, gate/up) in ONE launch and ONE cross-GPU arrival."""

    def __init__(self, weight_shards: list, F_totals: list, K: int, wtype: int, T: int, plan: PeerPlan,
                 align: int = DEFAULT_ALIGN, flags: int = 0, wait_index: Optional[int] = None):
        import ctypes as C
        self.plan, self.K, self.wtype, self.T, self.flags = plan, K, wtype, T, flags
        self.n = len(weight_shards)
        self.weights = [w.contiguous() for w in weight_shards]
        self.outs, offs, fs = [], [], []
        for w, Ft in zip(self.weights, F_totals):
            f0, f1 = shard_rows(Ft, plan.world, plan.rank, align)
            assert w.shape[0] == f1 - f0 > 0
            off = plan.alloc(Ft * T)
            self.outs.append(plan.pool[off:off + Ft * T].view(Ft, T))
            offs.append(off + f0 * T)
            fs.append(f1 - f0)
        self.launch_index = plan.next_index
        plan.next_index += 1
        assert plan.next_index <= plan.lps
        self.ps = plan.peers_struct(0, self.launch_index, wait_index)
        self._wp = (C.c_void_p * self.n)(*[w.data_ptr() for w in self.weights])
        self._fs = (C.c_int * self.n)(*fs)
        self._offs = (C.c_int64 * self.n)(*offs)

    def __call__(self, activation_q: torch.Tensor) -> list:
        rc = _lib.lib().qgemm_gemm_group_peers(self.wtype, activation_q.data_ptr(), self.n, self._wp, self._fs, self._offs,
                                               self.ps, self.T, self.K, 1, self.T, self.flags,
                                               torch.cuda.current_stream(self.weights[0].device).cuda_stream)
        _lib.raise_on_error(rc, "gemm_group_peers")
        return self.outs
