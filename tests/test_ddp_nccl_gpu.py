"""Data-parallel training over real NCCL (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_ddp_nccl_gpu.py
-m gpu`; skipped on one GPU).  The role torch's DistributedDataParallel plays in
train_options/train_original.py:149,204-209.

Two ranks, each seeded differently BEFORE the model is built (train.py: seed = global_seed * world + rank), the
optimizer built before the wrapper (bench.py's order):
  * after construction every rank holds rank 0's weights, bf16 shadows derived from them and an EMA equal to them
    (train.py:179 re-syncs the EMA after DDP the same way);
  * DataParallel's gradients equal the MEAN over ranks of the gradients each rank computes alone on its shard, for
    the f32 wire (the reference's DDP) and the bf16 wire;
  * one optimizer step later all ranks still hold identical weights and EMA."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist

        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        from fast_dit_b200 import DiT_models
        from fast_dit_b200.optim import FusedAdamWEMA
        from fast_dit_b200.parallel import DataParallel, init_from_env
        from fast_dit_b200.utils import rerandomise_zero_params

        r, local, w = init_from_env("nccl")
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        res = {}
        after_step = {}
        for wire, overlap in ((torch.float32, False), (torch.bfloat16, False), (torch.float32, True), (torch.bfloat16, True),
                              (torch.float32, "shard")):
            torch.manual_seed(7 * world + rank)  # rank-dependent initial weights
            m = DiT_models["DiT-S/4"](input_size=32, num_classes=1000, precision="bf16")
            rerandomise_zero_params(m, seed=1234 + rank)
            m = m.to(dev).train()
            opt = FusedAdamWEMA(m, lr=1e-3, weight_decay=0.0, ema_decay=0.99, overlap_backward=overlap is True)
            before = m.blocks[0].attn.qkv.weight.detach().clone()
            net = DataParallel(m, grad_dtype=wire, shard_optimizer=overlap == "shard")
            tag = ("f32" if wire == torch.float32 else "bf16") + ("+shard" if overlap == "shard" else "+overlap" if overlap else "")
            # ---- replicas agree after construction
            flat = opt.flat.clone()
            ref = flat.clone()
            dist.broadcast(ref, src=0)
            res[tag + ".weights_equal_rank0"] = bool(torch.equal(flat, ref))
            res[tag + ".weights_changed_on_nonzero_rank"] = bool(rank == 0 or not torch.equal(before, m.blocks[0].attn.qkv.weight))
            res[tag + ".shadow_matches"] = bool(torch.equal(opt.shadow, opt.flat.bfloat16()))
            res[tag + ".ema_matches"] = bool(torch.equal(opt.ema, opt.flat))
            # ---- gradients: alone on the own shard, then through the wrapper
            g = torch.Generator(device=dev).manual_seed(100 + rank)
            x = torch.randn(4, 4, 32, 32, device=dev, generator=g)
            t = torch.randint(0, 1000, (4,), device=dev, generator=g)
            y = torch.randint(0, 1000, (4,), device=dev, generator=g)
            dout = torch.randn(4, 8, 32, 32, device=dev, generator=g)
            m.eval()  # no label dropout: both passes see the same labels
            # (the wrapper and an overlapping optimizer install their bucket hooks on the module itself)
            hooks = (m._grad_sync, getattr(m, "_bucket_ready", None))
            m._grad_sync = m._bucket_ready = None
            m(x, t, y).backward(dout)
            m._grad_sync, m._bucket_ready = hooks
            alone = torch.cat([p.grad.flatten().float() for p in m.parameters() if p.grad is not None])
            mean = alone.clone()
            dist.all_reduce(mean, op=dist.ReduceOp.SUM)
            mean /= world
            m.zero_grad(set_to_none=True)
            w_before = opt.flat.clone()
            net(x, t, y).backward(dout)
            torch.cuda.synchronize()
            res[tag + ".updated_during_backward"] = bool(not torch.equal(w_before, opt.flat))
            got = torch.cat([p.grad.flatten().float() for p in m.parameters() if p.grad is not None])
            res[tag + ".grad_vs_mean"] = _rel(got, mean)
            res[tag + ".grad_vs_alone"] = _rel(got, alone)  # must NOT be small: the shards differ
            # ---- one step: replicas stay identical
            opt.step()
            opt.zero_grad()
            if overlap == "shard":
                # the bf16 shadows (what the forward reads) are complete on every rank without consolidating ...
                sh = opt.shadow.clone()
                ref = sh.clone()
                dist.broadcast(ref, src=0)
                res[tag + ".shadow_equal_after_step"] = bool(torch.equal(sh, ref))
                # ... the f32 state is not, and says so, until every rank consolidates
                try:
                    m.state_dict()
                    res[tag + ".state_dict_guard"] = False
                except Exception:
                    res[tag + ".state_dict_guard"] = True
                opt.consolidate()
                m.state_dict()
                res[tag + ".shadow_follows_weights"] = bool(torch.equal(opt.shadow, opt.flat.bfloat16()))
                res[tag + ".vs_replicated"] = _rel(opt.flat, after_step["f32"])
                res[tag + ".ema_vs_replicated"] = _rel(opt.ema, after_step["f32.ema"])
            after_step[tag], after_step[tag + ".ema"] = opt.flat.clone(), opt.ema.clone()
            flat = opt.flat.clone()
            ref = flat.clone()
            dist.broadcast(ref, src=0)
            res[tag + ".weights_equal_after_step"] = bool(torch.equal(flat, ref))
            ema = opt.ema.clone()
            ref = ema.clone()
            dist.broadcast(ref, src=0)
            res[tag + ".ema_equal_after_step"] = bool(torch.equal(ema, ref))
            del net, opt, m
        # ---- the reference's own wrapper: torch DistributedDataParallel around this package's DiT
        # (train_options/train_original.py:149 — INTEGRATION.md says it works unchanged: gradients reach the
        # parameters through autograd, where DDP's reducer hooks pick them up)
        torch.manual_seed(3)  # same weights on every rank, so that "alone" can be taken before DDP is built
        m = DiT_models["DiT-S/4"](input_size=32, num_classes=1000, precision="bf16")
        rerandomise_zero_params(m)
        m = m.to(dev).eval()
        g = torch.Generator(device=dev).manual_seed(200 + rank)
        x = torch.randn(4, 4, 32, 32, device=dev, generator=g)
        t = torch.randint(0, 1000, (4,), device=dev, generator=g)
        y = torch.randint(0, 1000, (4,), device=dev, generator=g)
        dout = torch.randn(4, 8, 32, 32, device=dev, generator=g)
        m(x, t, y).backward(dout)
        alone = torch.cat([p.grad.flatten().float() for p in m.parameters() if p.grad is not None])
        mean = alone.clone()
        dist.all_reduce(mean, op=dist.ReduceOp.SUM)
        mean /= world
        m.zero_grad(set_to_none=True)
        ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[local])
        ddp(x, t, y=y).backward(dout)
        torch.cuda.synchronize()
        got = torch.cat([p.grad.flatten().float() for p in m.parameters() if p.grad is not None])
        res["torch_ddp.grad_vs_mean"] = _rel(got, mean)
        res["torch_ddp.grad_vs_alone"] = _rel(got, alone)
        del ddp
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, res))
    except Exception as e:  # noqa: BLE001
        import traceback

        q.put((rank, {"error": f"{e!r}\n{traceback.format_exc()}"}))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_over_nccl_two_ranks():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank in range(world):
        res = out[rank]
        assert "error" not in res, res.get("error")
        print(f"rank {rank}: " + ", ".join(f"{k}={v if isinstance(v, bool) else f'{v:.2e}'}" for k, v in res.items()))
        assert res["f32+shard.shadow_equal_after_step"] and res["f32+shard.state_dict_guard"], (rank, res)
        assert res["f32+shard.shadow_follows_weights"], (rank, res)
        # same update as the replicated optimizer (the gradients differ run to run by the atomics' rounding, ~1e-3
        # relative, times lr = 1e-3 per step)
        assert res["f32+shard.vs_replicated"] < 1e-4 and res["f32+shard.ema_vs_replicated"] < 1e-4, (rank, res)
        for tag, tol in (("f32", 2e-3), ("bf16", 8e-3), ("f32+overlap", 2e-3), ("bf16+overlap", 8e-3), ("f32+shard", 2e-3)):
            if tag.endswith("shard"):
                continue  # its gradients live as parts on their owners: covered by the weight comparison above
            assert res[f"{tag}.updated_during_backward"] is tag.endswith("overlap"), (rank, tag)
            for k in ("weights_equal_rank0", "weights_changed_on_nonzero_rank", "shadow_matches", "ema_matches",
                      "weights_equal_after_step", "ema_equal_after_step"):
                assert res[f"{tag}.{k}"] is True, (rank, tag, k)
            # f32 wire: NCCL's average of f32 buckets vs our f32 mean; the tolerance covers the run-to-run spread of
            # the atomically accumulated weight gradients (3e-3 bound used for a single rank in test_backward_gpu.py)
            assert res[f"{tag}.grad_vs_mean"] < tol, (rank, tag, res[f"{tag}.grad_vs_mean"])
            assert res[f"{tag}.grad_vs_alone"] > 0.1, (rank, tag)
        assert res["torch_ddp.grad_vs_mean"] < 2e-3 and res["torch_ddp.grad_vs_alone"] > 0.1, (rank, res)
