"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on
the GPU box, gloo in the CPU tests).  The reference is single-process; what is added
here is exactly one exchange step — the bucketed all-reduce of the flat gradient arena — plus an
8-byte all-reduce of the non-PAD target count so that the N-rank step equals the 1-rank step on
the concatenated batch (CrossEntropyLoss is a mean over ALL non-PAD targets, train.py:327,90).

Buckets are contiguous ranges of the arena in the order backward finishes them (fc_out, layer L-1
.. 0, embedding+projection).  The engine records one CUDA event per bucket on the compute stream;
each all-reduce is launched on a side stream after its event, so it overlaps the rest of backward.
"""
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .engine import nvtx_range


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of a global batch owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def plan_buckets(segments: Sequence[Tuple[int, int]], cap_elems: int) -> List[Tuple[int, int]]:
    """Merge consecutive (offset, count) segments, given in readiness order and contiguous in
    DEcreasing address order, into buckets of at most cap_elems (a single larger segment stays one
    bucket).  Returns (offset, count) per bucket in readiness order."""
    out: List[Tuple[int, int]] = []
    cur_off, cur_cnt = None, 0
    for off, cnt in segments:
        if cur_off is not None and off + cnt == cur_off and cur_cnt + cnt <= cap_elems:
            cur_off, cur_cnt = off, cur_cnt + cnt
        else:
            if cur_off is not None:
                out.append((cur_off, cur_cnt))
            cur_off, cur_cnt = off, cnt
    if cur_off is not None:
        out.append((cur_off, cur_cnt))
    return out


def allreduce_flat(flat: torch.Tensor, buckets: Sequence[Tuple[int, int]], group=None) -> None:
    """Sum `flat` over ranks bucket by bucket (synchronous; device-agnostic, used by the CPU tests)."""
    for off, cnt in buckets:
        dist.all_reduce(flat[off:off + cnt], op=dist.ReduceOp.SUM, group=group)


class DataParallel:
    """Per-rank helper bound to one DecoderEngine."""

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.buckets = engine.grad_buckets()                      # readiness order
        self.events = [torch.cuda.Event() for _ in self.buckets]
        self.comm_stream = torch.cuda.Stream(device=engine.device)
        self._cnt = torch.zeros(1, device=engine.device, dtype=torch.float32)
        # Optional second communicator with few CTAs for the EARLY buckets (B200_DP_SLOW_CTAS = its maxCTAs, 0 = off): their
        # all-reduces have the whole rest of backward to hide under, so they may be slow, and every SM NCCL does not hold
        # is an SM the statically scheduled persistent GEMMs do not wait for; the last `fast_tail` buckets (exposed after
        # backward ends) keep the default communicator.
        self.slow_group = None
        self.fast_tail = int(os.environ.get("B200_DP_FAST_TAIL", "2"))
        slow_ctas = int(os.environ.get("B200_DP_SLOW_CTAS", "0"))
        if slow_ctas > 0 and self.world_size > 1 and dist.get_backend(group) == "nccl":
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = slow_ctas
            opts.config.min_ctas = 1
            self.slow_group = dist.new_group(ranks=list(range(self.world_size)), pg_options=opts)

    @staticmethod
    def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
        """(rank, local_rank, world_size) from the torchrun environment; initialises the group."""
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend == "nccl":
                torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
        return rank, local, world

    def broadcast_parameters(self) -> None:
        """Replicas must start identical (rank 0 wins)."""
        if self.world_size > 1:
            dist.broadcast(self.engine.params, src=0, group=self.group)
            self.engine.mark_weights_changed()

    def global_inv_count(self, loss_and_count: torch.Tensor) -> torch.Tensor:
        """Device scalar 1 / sum_over_ranks(n_valid): the scale of dlogits on every rank."""
        self._cnt.copy_(loss_and_count[1:2])
        dist.all_reduce(self._cnt, op=dist.ReduceOp.SUM, group=self.group)
        return 1.0 / self._cnt

    def global_loss(self, loss_and_count: torch.Tensor, inv: torch.Tensor) -> torch.Tensor:
        """[global mean loss, global count] from the per-rank [mean, count]."""
        s = (loss_and_count[0:1] * loss_and_count[1:2]).clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=self.group)
        return torch.cat([s * inv, self._cnt])

    def backward_and_allreduce(self, inv_count: torch.Tensor) -> None:
        """Backward one gradient bucket at a time; each bucket's all-reduce is launched on the comm
        stream behind a (torch) event recorded right after the bucket's last gradient write, so it
        overlaps the rest of backward.  Fork/join through torch events only: the whole sequence can
        be captured in a CUDA graph (train.GraphedTrainStep).  Returns True when the squared global gradient norm
        has been accumulated along the way (engine.norm_add_bucket), so that the optimizer can skip its own pass."""
        g = self.engine.grads
        cur = torch.cuda.current_stream()
        overlap_norm = self.engine.buckets_cover_arena() and os.environ.get("B200_OVERLAP_NORM", "1") != "0"
        if overlap_norm:
            self.engine.norm_begin()
        for i, ((off, cnt), ev) in enumerate(zip(self.buckets, self.events)):
            self.engine.backward_parts(i, i, inv_count)
            ev.record(cur)
            grp = self.slow_group if (self.slow_group is not None and i < len(self.buckets) - self.fast_tail) else self.group
            with torch.cuda.stream(self.comm_stream), nvtx_range("b200.dp.allreduce_bucket"):
                self.comm_stream.wait_event(ev)
                dist.all_reduce(g[off:off + cnt], op=dist.ReduceOp.SUM, group=grp)
                if overlap_norm:
                    self.engine.norm_add_bucket(i)     # global norm of the REDUCED gradients, bucket by bucket
        cur.wait_stream(self.comm_stream)
        return overlap_norm

    def allreduce_buckets(self) -> None:
        """Launch one all-reduce per bucket on the comm stream as soon as backward has finished
        it, then make the compute stream wait for all of them (before clip + AdamW)."""
        g = self.engine.grads
        with torch.cuda.stream(self.comm_stream):
            for (off, cnt), ev in zip(self.buckets, self.events):
                self.comm_stream.wait_event(ev)
                dist.all_reduce(g[off:off + cnt], op=dist.ReduceOp.SUM, group=self.group)
        torch.cuda.current_stream().wait_stream(self.comm_stream)
