"""Fused AdamW + EMA + bf16-shadow step over flat arenas.

The reference's training loop (train.py:153-161,206-207; train_options/train_original.py:147,155,210-211)
keeps `ema = deepcopy(model)`, runs `torch.optim.AdamW(lr=1e-4, weight_decay=0).step()` and then
`update_ema(ema, model)`, a Python loop of mul_/add_ pairs — roughly 1200 small launches per step, and
autocast re-casts every weight to bf16 in the next forward.  Here the parameters, gradients, Adam moments,
EMA weights and bf16 shadows all share one layout (training.ArenaLayout) and `step()` is ONE kernel
(csrc/optim.cu) that touches each byte once.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops
from .training import GradArena, layout_for


class FusedAdamWEMA:
    """opt = FusedAdamWEMA(model, lr=1e-4, weight_decay=0.0, ema_decay=0.9999)
    loss.backward(); opt.step(); opt.zero_grad()

    Construction moves the model's parameters into a flat f32 arena (each nn.Parameter becomes a view, so
    state_dict()/load_state_dict()/DDP keep working) and attaches the arena to the model, whose forward
    then reads its bf16 GEMM weights from the shadow arena this optimizer maintains."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, ema_decay=0.9999,
                 overlap_backward: bool = False):
        """overlap_backward=True applies the update of every gradient bucket (final layer, each block from last to
        first, embedders) on a side stream AS SOON AS backward has finished that bucket (after its all-reduce under
        data parallelism), so the HBM-bound optimizer pass runs underneath the tensor-bound backward GEMMs of the
        earlier blocks; step() then only joins the streams.  Same arithmetic as backward(); step().  Opt-in, because
        the parameters already hold the new values when backward() returns: no gradient accumulation over several
        backward calls and no gradient clipping between backward() and step() in this mode."""
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise L.Ditb200Error("FusedAdamWEMA needs the model on a CUDA device (there is no CPU path)")
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        self.layout = lay = layout_for(model)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.flat = torch.zeros(lay.total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p in self.params:
                v = lay.view(self.flat, p)
                v.copy_(p.data)
                p.data = v
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.ema = self.flat.clone() if ema_decay is not None else None  # ema = deepcopy(model) (train.py:153)
        self.shadow = torch.empty(lay.total, device=dev, dtype=torch.bfloat16) if model.precision == "bf16" else None
        self.step_count = 0
        self._views = None
        self._versions = None
        self.overlap_backward = bool(overlap_backward)
        self._side = None          # stream the early updates run on
        self._applied = None       # bucket keys already updated in the step that is being built (None: no step open)
        model._flat = self
        model._shadow = {}
        if self.overlap_backward:
            model._bucket_ready = self._bucket_ready
        self.refresh()

    # ------------------------------------------------------------------ what the model's forward reads
    def _snapshot(self):
        return tuple(p._version for p in self.params)

    def refresh(self):
        """Re-derive the bf16 shadows from the f32 master weights (after load_state_dict or any edit of
        the parameters that did not go through step())."""
        self.consolidate()  # sharded state: the non-owned f32 parts must be current first (collective, symmetric)
        if self.shadow is not None:
            ops.cast_bf16(self.flat, out=self.shadow)
        self._versions = self._snapshot()

    @torch.no_grad()
    def resync_from_parameters(self, reset_ema: bool = True):
        """The parameters were overwritten from outside (data-parallel broadcast, manual edits through .data, which
        do not bump tensor versions): rebuild the bf16 shadows and, like the reference's
        `update_ema(ema, model, decay=0)` (train.py:179), restart the EMA from the new weights."""
        self.refresh()
        if reset_ema and self.ema is not None:
            self.ema.copy_(self.flat)

    def shadows(self):
        if self._versions != self._snapshot():
            self.refresh()
        if self._views is None:
            m, lay = self.model, self.layout
            src = self.shadow if self.shadow is not None else self.flat
            lo, hi = lay.ranges["ada_w"]
            blo, bhi = lay.ranges["ada_b"]
            self._views = {
                "key": "flat",
                "w": [lay.view(src, w) for w in m._gemm_weights()],
                "ada_w": src[lo:lo + lay.ada_rows * m.hidden_size].view(lay.ada_rows, m.hidden_size),
                "ada_b": self.flat[blo:blo + lay.ada_rows],
            }
        return self._views

    # ------------------------------------------------------------------------------------ the step
    def _update(self, grads, lo, hi, offset=0):
        """The fused AdamW + EMA + bf16-shadow pass over arena elements [lo, hi); their gradients are
        grads[lo - offset : hi - offset] (offset != 0: `grads` is this rank's part of a reduce-scattered region)."""
        sl = slice(lo, hi)
        ops.adamw_ema(self.flat[sl], grads[lo - offset:hi - offset], self.exp_avg[sl], self.exp_avg_sq[sl],
                      self.ema[sl] if self.ema is not None else None, self.shadow[sl] if self.shadow is not None else None,
                      lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay,
                      step=self.step_count, ema_decay=self.ema_decay if self.ema is not None else 0.0)

    # ------------------------------------------------------------------ state partitioned over data-parallel ranks
    def enable_sharding(self, dp):
        """Called by parallel.DataParallel(shard_optimizer=True): from now on this rank updates only its part of
        every sharded region (dp.part) and the update runs during backward."""
        for key, regions in self.layout.big.items():
            for lo, hi in regions:
                if (hi - lo) % (dp.world * 64):
                    raise L.Ditb200Error(f"region {key} [{lo}, {hi}) does not divide into {dp.world} 64-float parts")
        self._dp = dp
        self._ag_pending = []
        self._stale = False  # non-owned parts of flat / moments / ema are out of date
        self.overlap_backward = True
        self.model._bucket_ready = self._bucket_ready

        def guard(module, prefix, keep_vars):
            # the reference saves checkpoints on rank 0 only (train.py:229-239): a collective hidden in state_dict()
            # would hang there, so ask for the explicit, all-rank consolidate() instead
            if self._stale:
                raise L.Ditb200Error("the optimizer state is sharded over the data-parallel ranks: call "
                                     "opt.consolidate() on EVERY rank before model.state_dict() / opt.state_dict()")
        self.model.register_state_dict_pre_hook(guard)

    def _require_consolidated(self):
        if getattr(self, "_dp", None) is not None and self._stale:
            raise L.Ditb200Error("the optimizer state is sharded over the data-parallel ranks: call opt.consolidate() "
                                 "on EVERY rank first")

    @torch.no_grad()
    def consolidate(self):
        """Gather every rank's parts of the f32 master weights, the Adam moments and the EMA, so that each rank
        holds the complete state again (checkpoints, evaluation of the f32 weights).  Collective: all ranks call it."""
        dp = getattr(self, "_dp", None)
        if dp is None or not self._stale:
            return
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        for regions in self.layout.big.values():
            for lo, hi in regions:
                for arr in (self.flat, self.exp_avg, self.exp_avg_sq, self.ema):
                    if arr is not None:
                        dp.gather_region(arr, lo, hi)
        self._stale = False

    @torch.no_grad()
    def _bucket_ready(self, key, arena, after=None, plan=None):
        """Called by the model's backward when bucket `key` is final (key None: backward is over).  `after`: what
        must run on the update stream first — the data-parallel wrapper's wait for the bucket's collectives.  `plan`
        (sharded optimizer): [("shard" | "full", lo, hi, gradients)] — the regions of the bucket this rank updates a
        part of / updates whole."""
        m = self.model
        if arena is not getattr(m, "_grad_arena", None):  # a scratch arena: the caller is accumulating gradients
            raise L.Ditb200Error("FusedAdamWEMA(overlap_backward=True) updates during backward: call "
                                 "zero_grad(set_to_none=True) between steps (no gradient accumulation in this mode)")
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        if key is None:
            for w in getattr(self, "_ag_pending", ()):  # the gathered bf16 shadows the next forward reads
                w.wait()
            if getattr(self, "_dp", None) is not None:
                self._ag_pending = []
            cur.wait_stream(self._side)  # whatever follows backward sees the gradients averaged and the weights updated
            return
        if self._applied is None:
            if self._versions != self._snapshot():
                self.refresh()
            self.step_count += 1
            self._applied = set()
        self._side.wait_stream(cur)  # the bucket's last gradient kernel has been queued on `cur`
        with torch.cuda.stream(self._side):
            for fn in after or ():
                fn()
            if plan is None:
                for lo, hi in self.layout.buckets[key]:
                    self._update(arena.flat, lo, hi)
            else:
                dp = self._dp
                for kind, lo, hi, grads in plan:
                    if kind == "full":
                        self._update(arena.flat, lo, hi)
                    else:
                        plo, phi = dp.part(lo, hi)
                        self._update(grads, plo, phi, offset=plo)
                        if self.shadow is not None:  # every rank's forward needs the whole region's bf16 weights
                            w = dp.gather_region(self.shadow, lo, hi, async_op=True)
                        else:
                            w = dp.gather_region(self.flat, lo, hi, async_op=True)
                        if w is not None:
                            self._ag_pending.append(w)
                self._stale = True
        self._applied.add(key)

    @torch.no_grad()
    def step(self):
        m = self.model
        if getattr(m, "_flat", None) is not self:
            raise L.Ditb200Error("the model's parameters were moved after FusedAdamWEMA was built; rebuild the optimizer")
        if self._applied is not None:  # updates were applied during backward: finish what is left and join
            arena = m._grad_arena
            torch.cuda.current_stream().wait_stream(self._side)
            for key, slices in self.layout.buckets.items():
                if key not in self._applied:
                    for lo, hi in slices:
                        self._update(arena.flat, lo, hi)
            self._applied = None
            return
        arena = getattr(m, "_grad_arena", None)
        if arena is None:
            arena = GradArena(m)
            m._grad_arena = arena
        for p in self.params:  # gradients that did not land in the arena (torch DDP, accumulation, no backward)
            if not arena.aliases(p):
                if p.grad is None:
                    arena.view(p).zero_()
                else:
                    arena.view(p).copy_(p.grad)
        if self._versions != self._snapshot():
            self.refresh()
        self.step_count += 1
        ops.adamw_ema(self.flat, arena.flat, self.exp_avg, self.exp_avg_sq, self.ema, self.shadow, lr=self.lr,
                      beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay,
                      step=self.step_count, ema_decay=self.ema_decay if self.ema is not None else 0.0)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    # --------------------------------------------------------------------------------- checkpoints
    def _named_views(self, arena):
        """{parameter name: view of `arena` (one of the flat f32 arenas) for that parameter}."""
        ids = {id(p): self.layout.view(arena, p) for p in self.params}
        return {k: ids[id(v)] for k, v in self.model.named_parameters() if id(v) in ids}

    def ema_state_dict(self):
        """state_dict of the EMA model (what train.py:233 saves under "ema"; download.py:26-29 loads it)."""
        if self.ema is None:
            raise L.Ditb200Error("this optimizer was built with ema_decay=None: there is no EMA model")
        self._require_consolidated()
        ids = {id(p): self.layout.view(self.ema, p) for p in self.params}
        return {k: ids.get(id(v), v).detach().clone() for k, v in self.model.state_dict(keep_vars=True).items()}

    @torch.no_grad()
    def load_ema_state_dict(self, sd):
        """Restore the EMA weights from a model-style state_dict (the "ema" entry of a train.py checkpoint,
        train.py:231-236).  Frozen entries (pos_embed) are ignored."""
        if self.ema is None:
            raise L.Ditb200Error("this optimizer was built with ema_decay=None: there is no EMA model")
        views = self._named_views(self.ema)
        missing = [k for k in views if k not in sd]
        if missing:
            raise KeyError(f"EMA state_dict lacks {missing[:3]}{'...' if len(missing) > 3 else ''}")
        for k, v in views.items():
            v.copy_(sd[k].to(v.device, torch.float32).reshape(v.shape))

    def state_dict(self):
        """Checkpoint of the optimizer in torch.optim.AdamW's format — {"state": {i: {"step", "exp_avg",
        "exp_avg_sq"}}, "param_groups": [...]} with i = position in model.parameters(), the numbering torch's
        AdamW(model.parameters()) uses (train.py:161, saved as "opt" at train.py:234; the frozen pos_embed takes a
        number but has no state) — plus the EMA weights under "ema" (per-parameter, cloned).  Every tensor is a copy."""
        self._require_consolidated()
        m, v = self._named_views(self.exp_avg), self._named_views(self.exp_avg_sq)
        names = [k for k, p in self.model.named_parameters()]
        step = torch.tensor(float(self.step_count))
        state = {i: {"step": step.clone(), "exp_avg": m[k].detach().clone(), "exp_avg_sq": v[k].detach().clone()}
                 for i, k in enumerate(names) if k in m}
        # every key this torch version's AdamW keeps in a param group (a loader fills absent ones with Adam's
        # defaults, which would silently turn decoupled weight decay off)
        group = dict(torch.optim.AdamW([torch.zeros(1)], lr=self.lr, betas=tuple(self.betas), eps=self.eps,
                                       weight_decay=self.weight_decay).state_dict()["param_groups"][0])
        group["params"] = list(range(len(names)))
        out = {"state": state, "param_groups": [group], "param_names": names}
        if self.ema is not None:
            out["ema"] = {k: t.detach().clone() for k, t in self._named_views(self.ema).items()}
        return out

    @torch.no_grad()
    def load_state_dict(self, sd):
        """Accepts this class's state_dict() and a plain torch.optim.AdamW state_dict over the same parameters (the
        "opt" entry of a reference checkpoint).  A missing "ema" entry leaves the EMA untouched: restore it with
        load_ema_state_dict(checkpoint["ema"])."""
        names = [k for k, p in self.model.named_parameters()]
        state = sd["state"]
        m, v = self._named_views(self.exp_avg), self._named_views(self.exp_avg_sq)
        if len(state) not in (0, len(m)) or any(int(i) >= len(names) or names[int(i)] not in m for i in state):
            raise ValueError(f"optimizer state holds {len(state)} parameters, the model has {len(m)} trainable ones")
        state = {int(i): st for i, st in state.items()}
        steps = set()
        for i, k in enumerate(names):
            if i not in state:
                continue
            st = state[i]
            m[k].copy_(st["exp_avg"].to(m[k].device, torch.float32).reshape(m[k].shape))
            v[k].copy_(st["exp_avg_sq"].to(v[k].device, torch.float32).reshape(v[k].shape))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ; the fused step keeps one counter")
        self.step_count = steps.pop() if steps else 0
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
        if "ema" in sd and self.ema is not None:
            self.load_ema_state_dict(sd["ema"])
