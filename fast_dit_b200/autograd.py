"""Autograd glue: torch.autograd.Function wrappers whose forward AND backward are libditb200
kernels.  PyTorch only records the graph; it computes nothing on this path."""
from __future__ import annotations

import torch

from . import ops


class _DiffusionLoss(torch.autograd.Function):
    """training_losses' terms (gaussian_diffusion.py:747-781) from the model output."""

    @staticmethod
    def forward(ctx, model_output, x0, x_t, noise, t, tables, vb_scale, opts):
        mo = model_output.float().contiguous()
        r = ops.training_losses(mo, x0, x_t, noise, t, tables, vb_scale, **opts)
        ctx.save_for_backward(mo, x0, x_t, noise, t)
        ctx.tables, ctx.vb_scale, ctx.opts = tables, vb_scale, dict(opts, want_pred=False)
        pred = r["pred_xstart"] if r["pred_xstart"] is not None else mo.new_empty(0)
        ctx.mark_non_differentiable(pred)
        return r["loss"], r["mse"], r["vb"], pred

    @staticmethod
    def backward(ctx, g_loss, g_mse, g_vb, _g_pred):
        mo, x0, x_t, noise, t = ctx.saved_tensors
        z = torch.zeros(mo.shape[0], device=mo.device, dtype=torch.float32)
        g_loss = z if g_loss is None else g_loss.float()
        w_mse = (g_loss + (z if g_mse is None else g_mse.float())).contiguous()
        w_vb = (g_loss + (z if g_vb is None else g_vb.float())).contiguous()
        r = ops.training_losses(mo, x0, x_t, noise, t, ctx.tables, ctx.vb_scale, w_mse=w_mse, w_vb=w_vb, **ctx.opts)
        return r["grad_model_out"], None, None, None, None, None, None, None


def diffusion_loss(model_output, x0, x_t, noise, t, tables, vb_scale=1.0, **opts):
    """(loss, mse, vb, pred_xstart) of the fused loss kernel; opts: mean_type, var_type, clip_denoised,
    vb_through_mean, want_pred (ops.training_losses)."""
    return _DiffusionLoss.apply(model_output, x0, x_t, noise, t, tables, vb_scale, opts)


def dit_forward_autograd(model, x, t, y):
    from .training import dit_forward_train

    return dit_forward_train(model, x, t, y)
