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
