"""Multi-process (gloo, world_size 2, CPU) checks of the data-parallel host logic: the sampling
shard plan of sample_ddp.py and the max-over-ranks timing reduction bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_dit_b200.parallel import ShardPlan, make_cfg_batch


def test_shard_plan_matches_sample_ddp_arithmetic():
    # 50 000 FID samples, 32 per GPU, 8 GPUs (sample_ddp.py defaults)
    plans = [ShardPlan(8, r, 32, 50000) for r in range(8)]
    p0 = plans[0]
    assert p0.global_batch == 256 and p0.total_samples == 50176 and p0.samples_this_rank == 6272
    assert p0.iterations == 196
    idx = sorted(i for p in plans for i in p.all_indices())
    assert idx == list(range(50176)), "ranks tile the index space exactly once"
    assert [p.seed(3) for p in plans] == [24 + r for r in range(8)]
    # index formula of sample_ddp.py:134-138: i * world + rank + total
    assert plans[5].global_index(2, 7) == 7 * 8 + 5 + 2 * 256


def test_cfg_batch_layout():
    g = torch.Generator().manual_seed(0)
    z, y = make_cfg_batch(3, 8, 1000, "cpu", generator=g)
    assert z.shape == (6, 4, 8, 8) and torch.equal(z[:3], z[3:])
    assert y.shape == (6,) and (y[3:] == 1000).all() and (y[:3] < 1000).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from fast_dit_b200.parallel import init_from_env

    r, local, w = init_from_env("gloo")
    assert (r, w) == (rank, world) and dist.is_initialized()
    plan = ShardPlan(w, r, 4, 20)
    # every rank contributes its indices; together they must tile [0, total)
    mine = torch.tensor(plan.all_indices(), dtype=torch.long)
    gathered = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(gathered, mine)
    allidx = torch.cat(gathered).sort().values
    ok = torch.equal(allidx, torch.arange(plan.total_samples))
    # rank-local RNG streams differ (seed = global_seed * world + rank)
    torch.manual_seed(plan.seed(0))
    draw = torch.randn(4)
    draws = [torch.empty(4) for _ in range(w)]
    dist.all_gather(draws, draw)
    distinct = not torch.equal(draws[0], draws[1])
    # timing reduction used by bench.py: the slowest rank defines the step time
    t = torch.tensor([10.0 + 5.0 * r])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if r == 0:
        q.put((ok, distinct, float(t.item()), plan.total_samples))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, distinct, tmax, total = q.get(timeout=10)
    assert ok and distinct and tmax == 15.0 and total == 24


# ------------------------------------------------------------------ training: gradient buckets
def test_arena_layout_buckets_cover_every_parameter_once():
    from fast_dit_b200.models import DiT_models
    from fast_dit_b200.training import ArenaLayout

    m = DiT_models["DiT-S/4"](input_size=32)
    lay = ArenaLayout(m)
    covered = torch.zeros(lay.total, dtype=torch.int32)
    for key, slices in lay.buckets.items():
        for lo, hi in slices:
            covered[lo:hi] += 1
    assert int(covered.max()) == 1, "buckets overlap"
    for p in m.parameters():
        if p.requires_grad:
            off, n, shape = lay.offsets[id(p)]
            assert off % 4 == 0 and shape == tuple(p.shape)
            assert int(covered[off:off + n].min()) == 1, "a parameter is outside every bucket"
    # the adaLN weight region is exactly the batched [(6L+2) D, D] matrix, blocks first, final layer last
    lo, hi = lay.ranges["ada_w"]
    D = m.hidden_size
    assert lay.ada_rows == (6 * m.depth + 2) * D and hi - lo >= lay.ada_rows * D
    assert lay.offsets[id(m.blocks[3].adaLN_modulation[1].weight)][0] == lo + 3 * 6 * D * D
    assert lay.offsets[id(m.final_layer.adaLN_modulation[1].weight)][0] == lo + m.depth * 6 * D * D


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from fast_dit_b200.models import DiT_models
    from fast_dit_b200.parallel import DataParallel, init_from_env
    from fast_dit_b200.training import GradArena

    init_from_env("gloo")
    torch.manual_seed(rank)  # different initial weights per rank: the wrapper must broadcast rank 0's
    m = DiT_models["DiT-S/8"](input_size=32)
    ddp = DataParallel(m)
    w0 = m.blocks[0].attn.qkv.weight.detach().clone()
    ws = [torch.empty_like(w0) for _ in range(world)]
    dist.all_gather(ws, w0)
    same_weights = torch.equal(ws[0], ws[1])
    # emulate a backward: rank r's gradient arena holds (r + 1) everywhere; buckets are reduced in the order
    # backward completes them; afterwards every gradient must be the mean over ranks
    arena = GradArena(m)
    arena.flat.fill_(float(rank + 1))
    order = ["final_layer"] + [f"blocks.{i}" for i in range(m.depth - 1, -1, -1)] + ["embed"]
    for key in order:
        m._grad_sync(key, arena)
    m._grad_sync(None, arena)
    mean = sum(range(1, world + 1)) / world
    ok = all(bool((arena.view(p) == mean).all()) for p in m.parameters() if p.requires_grad)
    dist.barrier()
    if rank == 0:
        q.put((same_weights, ok, ddp.buckets_issued == order))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_buckets():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    same_weights, ok, order_ok = q.get(timeout=10)
    assert same_weights and ok and order_ok


# ------------------------------------------------------------------ training: optimizer state sharded over the ranks
class _FakeShardedOptimizer:
    """Stands in for optim.FusedAdamWEMA on the CPU: 'update' = write -gradient into the weights it is told to
    update, then hand the region back to the wrapper's all-gather."""

    def __init__(self, model):
        from fast_dit_b200.training import layout_for

        self.model, self.layout = model, layout_for(model)
        self.flat = torch.zeros(self.layout.total)
        self.updated = torch.zeros(self.layout.total, dtype=torch.int32)
        model._flat = self

    def resync_from_parameters(self):
        pass

    def enable_sharding(self, dp):
        self.dp = dp
        self.model._bucket_ready = self._bucket_ready

    def _bucket_ready(self, key, arena, after=None, plan=None):
        if key is None:
            return
        for fn in after or ():
            fn()
        for kind, lo, hi, grads in plan:
            if kind == "full":
                self.flat[lo:hi] = -grads
                self.updated[lo:hi] += 1
            else:
                plo, phi = self.dp.part(lo, hi)
                assert grads.numel() == phi - plo
                self.flat[plo:phi] = -grads
                self.updated[plo:phi] += 1
                self.dp.gather_region(self.flat, lo, hi)


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from fast_dit_b200.models import DiT_models
    from fast_dit_b200.parallel import DataParallel, init_from_env
    from fast_dit_b200.training import GradArena

    init_from_env("gloo")
    torch.manual_seed(0)
    m = DiT_models["DiT-S/8"](input_size=32)
    opt = _FakeShardedOptimizer(m)
    DataParallel(m, shard_optimizer=True)
    lay = opt.layout
    arena = GradArena(m)
    # rank r's gradient of arena element i is (r + 1) * (1 + i mod 7): the mean over ranks is (W + 1) / 2 * (1 + i mod 7)
    arena.flat.copy_((rank + 1.0) * (1.0 + torch.arange(lay.total) % 7))
    order = ["final_layer"] + [f"blocks.{i}" for i in range(m.depth - 1, -1, -1)] + ["embed"]
    for key in order:
        after, plan = m._grad_sync(key, arena)  # what training._DiTFunction.backward's sync() does
        m._bucket_ready(key, arena, after, plan)
    m._grad_sync(None, arena)
    want = -(world + 1) / 2.0 * (1.0 + torch.arange(lay.total) % 7)
    covered = torch.zeros(lay.total, dtype=torch.bool)
    for key in order:
        for lo, hi in lay.buckets[key]:
            covered[lo:hi] = True
    ok_values = bool(torch.equal(opt.flat[covered], want[covered]))  # after the all-gathers every rank holds every update
    # work split: a rank updated its part of every sharded region and all of every replicated one, nothing twice
    mine = int(opt.updated.sum())
    big = sum(hi - lo for k in order for lo, hi in lay.big[k])
    small = sum(hi - lo for k in order for lo, hi in lay.small[k])
    ok_split = int(opt.updated.max()) == 1 and mine == big // world + small
    dist.barrier()
    q.put((rank, ok_values, ok_split, big / (big + small)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_optimizer_collectives():
    """DataParallel(shard_optimizer=True) over gloo: reduce-scatter of the weight regions, all-reduce of the small
    replicated tail, this rank's parts handed to the optimizer, all-gather of what it updated."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok_values, ok_split, frac in res:
        assert ok_values and ok_split, (rank, ok_values, ok_split)
        assert frac > 0.9
