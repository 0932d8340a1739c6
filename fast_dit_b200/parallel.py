"""Data-parallel execution of the denoiser path on one 8xB200 box.

Sampling shards by batch exactly as /root/reference/sample_ddp.py does (cited SD:line): one
process per GPU, weights replicated, every rank draws its own latents/labels from the seed
`global_seed * world_size + rank` (SD:57-58) and runs an independent p_sample_loop; images never
interact, so there is NO collective inside the loop — only the barriers around it (SD:92,141).
Sample k of iteration i on rank r gets the global index `i * per_iter_total + k * world + r`
(SD:136).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class ShardPlan:
    world_size: int
    rank: int
    per_proc_batch: int
    num_samples: int

    @property
    def global_batch(self) -> int:  # SD:97
        return self.per_proc_batch * self.world_size

    @property
    def total_samples(self) -> int:
        """num_samples rounded up to a multiple of the global batch (SD:99)."""
        return int(math.ceil(self.num_samples / self.global_batch) * self.global_batch)

    @property
    def samples_this_rank(self) -> int:  # SD:103
        return self.total_samples // self.world_size

    @property
    def iterations(self) -> int:  # SD:105
        return self.samples_this_rank // self.per_proc_batch

    def seed(self, global_seed: int) -> int:  # SD:57
        return global_seed * self.world_size + self.rank

    def global_index(self, iteration: int, k: int) -> int:
        """Index of the k-th image this rank produces in `iteration` (SD:134-138)."""
        return iteration * self.global_batch + k * self.world_size + self.rank

    def all_indices(self) -> list[int]:
        return [self.global_index(i, k) for i in range(self.iterations) for k in range(self.per_proc_batch)]


def make_cfg_batch(n: int, latent_size: int, num_classes: int, device, in_channels: int = 4, generator=None):
    """One iteration's inputs (SD:110-118): z ~ N(0,1), y ~ U{0..num_classes-1}, then the
    classifier-free-guidance doubling z = [z; z], y = [y; null]."""
    z = torch.randn(n, in_channels, latent_size, latent_size, device=device, generator=generator)
    y = torch.randint(0, num_classes, (n,), device=device, generator=generator)
    z = torch.cat([z, z], 0)
    y = torch.cat([y, torch.full((n,), num_classes, device=device, dtype=y.dtype)], 0)
    return z, y


@torch.no_grad()
def sample_shard(model, diffusion, plan: ShardPlan, *, latent_size: int, num_classes: int = 1000,
                 cfg_scale: float = 4.0, global_seed: int = 0, device="cuda", on_batch=None):
    """Run this rank's share of a sample_ddp job.  Returns [(global indices, latents [n,C,H,W])]
    per iteration, or streams them to on_batch(indices, latents)."""
    torch.manual_seed(plan.seed(global_seed))
    n = plan.per_proc_batch
    results = []
    for it in range(plan.iterations):
        z, y = make_cfg_batch(n, latent_size, num_classes, device, model.in_channels)
        samples = diffusion.p_sample_loop(model.forward_with_cfg, z.shape, z, clip_denoised=False,
                                          model_kwargs=dict(y=y, cfg_scale=cfg_scale), progress=False, device=device)
        samples, _ = samples.chunk(2, dim=0)  # drop the null-class half (SD:129)
        idx = [plan.global_index(it, k) for k in range(n)]
        if on_batch is not None:
            on_batch(idx, samples)
        else:
            results.append((idx, samples))
    return results


def init_from_env(backend: str = "nccl"):
    """torchrun plumbing: returns (rank, local_rank, world_size); initialises torch.distributed
    when WORLD_SIZE > 1."""
    import os

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local, world


class DataParallel(torch.nn.Module):
    """Data-parallel training wrapper for this package's DiT: the role torch's
    DistributedDataParallel plays in /root/reference/train_options/train_original.py:149 (cited TO:line).

    Weights are replicated (broadcast from rank 0 at construction, like DDP's constructor); every rank
    runs forward/backward on its own batch shard.  The model's backward writes gradients into one flat
    arena (training.GradArena) and calls `_sync(bucket)` as soon as a bucket — the final layer, each
    DiT block from last to first, the embedders — is complete; the bucket's contiguous slice is
    all-reduced (mean) asynchronously, so NCCL traffic over NVLink overlaps the rest of backward.  The
    compute stream waits for all collectives only at the end of backward."""

    def __init__(self, module, process_group=None, broadcast_parameters: bool = True, grad_dtype=torch.float32,
                 shard_optimizer: bool = False):
        """grad_dtype=torch.bfloat16 sends the buckets over NVLink as bf16 (half the bytes; the f32 arena is rounded
        once on the way out and refilled from the averaged bf16 values).  The default, f32, is what the reference's
        DistributedDataParallel does.

        shard_optimizer=True (needs an optim.FusedAdamWEMA on the module; f32 wire): the optimizer state is
        partitioned over the ranks.  Per bucket, the GEMM-weight regions (99.5 % of DiT-XL/2) are REDUCE-SCATTERED
        instead of all-reduced — each rank receives the mean of 1/W of every region, half the NVLink traffic of an
        all-reduce — the rank applies AdamW + EMA to that part only (the HBM-bound optimizer pass shrinks by W), and the
        updated bf16 weight shadows, the only copy of those weights the forward reads, are ALL-GATHERED back, issued
        bucket by bucket while backward is still running.  Biases, embedders, final linear and adaLN biases (0.5 %)
        stay replicated: all-reduced and updated on every rank.  Same arithmetic as the replicated optimizer; the f32
        master weights, Adam moments and EMA of a region live on its owner only until consolidate() gathers them
        (state_dict / ema_state_dict do that)."""
        super().__init__()
        import torch.distributed as dist

        if grad_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("grad_dtype must be torch.float32 or torch.bfloat16")
        if shard_optimizer and grad_dtype != torch.float32:
            raise ValueError("shard_optimizer reduce-scatters f32 gradients; grad_dtype must be torch.float32")
        self.shard_optimizer = bool(shard_optimizer)
        self._grad_parts = {}  # (lo, hi) of a sharded region -> this rank's averaged gradient part
        self.grad_dtype = grad_dtype
        self.module = module
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._pending = []
        # DITB200_DDP_DYNAMIC=1: dynamic GEMM tile scheduling while collectives are in flight.  Off by default: on
        # 2 GPUs it measured within box-to-box noise of the static schedule (data-parallel findings).
        self.dynamic_gemm = os.environ.get("DITB200_DDP_DYNAMIC") is not None
        self._dynamic_prev = None
        self.buckets_issued = []  # bucket keys in the order they were reduced (inspection / tests)
        if self.world > 1 and broadcast_parameters:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.detach(), src=0, group=process_group)
            # An optimizer built BEFORE this wrapper (bench.py's order) holds bf16 shadows and an EMA copy of the
            # pre-broadcast weights; with per-rank seeds (train.py: seed = global_seed * world + rank before the
            # model is built) those differ from rank 0's.  The reference re-syncs at train.py:179
            # (update_ema(ema, model, decay=0) after DDP); same here.
            flat = getattr(module, "_flat", None)
            if flat is not None:
                flat.resync_from_parameters()
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        if self.shard_optimizer:
            opt = getattr(module, "_flat", None)
            if opt is None or not hasattr(opt, "enable_sharding"):
                raise ValueError("shard_optimizer=True needs an optim.FusedAdamWEMA built on the module first")
            opt.enable_sharding(self)
        module._grad_sync = self._sync

    # ---------------------------------------------------------------- collectives of the sharded optimizer
    def part(self, lo, hi):
        """This rank's part [plo, phi) of the sharded region [lo, hi)."""
        n = (hi - lo) // self.world
        return lo + self.rank * n, lo + (self.rank + 1) * n

    def gather_region(self, flat, lo, hi, async_op=False):
        """All-gather: every rank contributes its part of flat[lo:hi] and receives the others'."""
        import torch.distributed as dist

        plo, phi = self.part(lo, hi)
        if self.world == 1:
            return None
        return dist.all_gather_into_tensor(flat[lo:hi], flat[plo:phi].clone(), group=self.pg, async_op=async_op)

    def _sync(self, key, arena):
        import torch.distributed as dist

        def finish(work, buf, avg_done):
            work.wait()  # the CURRENT stream waits for the collective
            if isinstance(avg_done, torch.Tensor):  # bf16 wire format: refill the f32 arena slice
                if avg_done.is_cuda:
                    avg_done.record_stream(torch.cuda.current_stream())  # may be the optimizer's update stream
                buf.copy_(avg_done)
            elif not avg_done:
                buf.div_(self.world)

        if key is None:  # end of backward: the compute stream must see every reduced bucket
            for work, buf, avg_done in self._pending:
                finish(work, buf, avg_done)
            self._pending = []
            if self._dynamic_prev is not None:
                from . import ops

                ops.set_gemm_dynamic(self._dynamic_prev)
                self._dynamic_prev = None
            return None
        self.buckets_issued.append(key)
        if self.world == 1:
            return None
        if self.dynamic_gemm and self._dynamic_prev is None and arena.bucket(key)[0].is_cuda:
            # from the first collective to the end of backward the GEMMs share the SMs with NCCL's CTAs
            from . import ops

            self._dynamic_prev = ops.set_gemm_dynamic(True)
        first = len(self._pending)
        if self.shard_optimizer:
            lay = arena.layout
            plan, works = [], []
            for lo, hi in lay.big[key]:  # reduce-scatter: this rank gets the mean of its part
                plo, phi = self.part(lo, hi)
                out = self._grad_parts.get((lo, hi))
                if out is None:
                    out = self._grad_parts[(lo, hi)] = torch.empty(phi - plo, device=arena.flat.device, dtype=torch.float32)
                cuda = arena.flat.is_cuda
                w = dist.reduce_scatter_tensor(out, arena.flat[lo:hi], op=dist.ReduceOp.AVG if cuda else dist.ReduceOp.SUM,
                                               group=self.pg, async_op=True)
                works.append((w, out, cuda))
                plan.append(("shard", lo, hi, out))
            for lo, hi in lay.small[key]:  # replicated tail: all-reduce, updated everywhere
                buf = arena.flat[lo:hi]
                cuda = buf.is_cuda
                w = dist.all_reduce(buf, op=dist.ReduceOp.AVG if cuda else dist.ReduceOp.SUM, group=self.pg, async_op=True)
                works.append((w, buf, cuda))
                plan.append(("full", lo, hi, buf))
            return ([lambda w=w, b=b, a=a: finish(w, b, a) for (w, b, a) in works], plan)
        for buf in arena.bucket(key):
            if buf.is_cuda and self.grad_dtype == torch.bfloat16:
                from . import ops

                wire = ops.cast_bf16(buf)
                work = dist.all_reduce(wire, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
                self._pending.append((work, buf, wire))
            elif buf.is_cuda:  # NCCL averages in the collective
                work = dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
                self._pending.append((work, buf, True))
            else:            # gloo (CPU tests of the host logic): sum, divide afterwards
                work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
                self._pending.append((work, buf, False))
        if getattr(self.module, "_bucket_ready", None) is not None:
            # an optimizer that updates during backward takes the bucket over: it waits for these collectives on ITS
            # stream, finishes the averaging there and applies the update (training._DiTFunction.backward's sync())
            mine, self._pending = self._pending[first:], self._pending[:first]
            return [lambda w=w, b=b, a=a: finish(w, b, a) for (w, b, a) in mine]
        return None

    def forward(self, *args, **kwargs):
        self.buckets_issued = []
        return self.module(*args, **kwargs)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.module, name)
