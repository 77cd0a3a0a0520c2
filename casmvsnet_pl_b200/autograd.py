"""Differentiable wrappers of the hot-path kernels: what the reference's
training step (train.py:99-127: forward, SL1 loss on depth_l, backward) needs.

  WarpCostFn   K1: d(cost volume)/d(features) by casmvs_warp_cost_bwd (grid_sample's backward
               chained with the variance / group-wise-correlation reduction).  Hypotheses are
               detached in the reference (models/mvsnet.py:231), projections are data.
  Conv3dFn     K2: the raw 3x3x3 (transposed) convolution; data gradient = one of the forward
               kernels with re-arranged weights, weight gradient = casmvs_conv3d_wgrad.  The
               norm-act that follows it in training mode is batch-statistics ABN
               (casmvsnet_pl_b200.norm_act: F.batch_norm + leaky_relu, torch autograd).
  RegressFn    K3: d(depth)/d(logits); confidence is computed under no_grad (mvsnet.py:179).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from .ops import CONV, CONV_TRANSPOSE, FP32, NHWC, _ptr, _stream, check


def _dense_ndhwc(x):
    """logical (B,C,D,h,w) -> dense (B,D,h,w,C)."""
    return ops.volume_storage(x)


class WarpCostFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, proj_mats, depth_values, num_groups):
        f = feats.detach()
        cost = ops.warp_cost(f, proj_mats, depth_values, num_groups, NHWC)
        B, V, C, h, w = f.shape
        f_cl = f.permute(0, 1, 3, 4, 2).contiguous()          # (B,V,h,w,C); no copy if channels-last
        ctx.save_for_backward(f_cl, proj_mats.contiguous(), depth_values.contiguous())
        ctx.G = num_groups
        return cost

    @staticmethod
    def backward(ctx, gcost):
        f_cl, proj, dv = ctx.saved_tensors
        B, V, h, w, C = f_cl.shape
        D = dv.shape[1]
        g = _dense_ndhwc(gcost)
        gf = torch.zeros_like(f_cl)
        with torch.cuda.device(f_cl.device):
            check(_lib.load().casmvs_warp_cost_bwd(_ptr(f_cl), _ptr(proj), _ptr(dv), _ptr(g), _ptr(gf),
                                                   B, V, C, D, h, w, ctx.G, _stream()), "warp_cost_bwd")
        return gf.permute(0, 1, 4, 2, 3), None, None, None


def _raw_conv(x, weight_torch, kind, stride, precision):
    cin, cout = (weight_torch.shape[1], weight_torch.shape[0]) if kind == CONV else \
        (weight_torch.shape[0], weight_torch.shape[1])
    wp = ops.pack_conv3d_weight(weight_torch, kind)
    y = ops.conv3d(x, wp, cin, cout, None, None, 1.0, None, kind, stride, precision)
    ops.release_weight_images(wp)          # one-shot weights: do not let operand images pile up
    return y


class Conv3dFn(torch.autograd.Function):
    """y = Conv3d(k3, pad 1, stride 1|2)(x) or ConvTranspose3d(k3, s2, p1, op1)(x), no bias."""

    @staticmethod
    def forward(ctx, x, weight, kind, stride, precision):
        xd, wd = x.detach(), weight.detach()
        y = _raw_conv(xd, wd, kind, stride, precision)
        ctx.save_for_backward(xd, wd)
        ctx.cfg = (kind, stride, precision)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        kind, stride, precision = ctx.cfg
        gx = gw = None
        gyd = _dense_ndhwc(gy)                                  # (B,Do,ho,wo,Cout)
        gy_l = ops.as_volume_view(gyd)
        xs = _dense_ndhwc(x)
        B, Di, hi, wi, cin = xs.shape
        _, Do, ho, wo, cout = gyd.shape
        if ctx.needs_input_grad[0]:
            if kind == CONV and stride == 1:
                # adjoint of a stride-1 correlation: correlate with the flipped kernel, channels swapped
                wt = w.flip(2, 3, 4).transpose(0, 1).contiguous()            # (Cin,Cout,3,3,3)
                g_in = gy_l
                if cout % 4:                                                 # prob head: Cout = 1
                    pad = 4 - cout % 4
                    g_in = ops.as_volume_view(torch.nn.functional.pad(gyd, (0, pad)))
                    wt = torch.nn.functional.pad(wt, (0, 0, 0, 0, 0, 0, 0, pad))
                gx = _raw_conv(g_in, wt, CONV, 1, precision)
            elif kind == CONV:
                # conv(k3,s2,p1) on even dims: its adjoint is ConvTranspose3d(k3,s2,p1,op1) with
                # the SAME weight tensor read as (in = Cout, out = Cin)
                gx = _raw_conv(gy_l, w, CONV_TRANSPOSE, 2, precision)
            else:
                # adjoint of the transposed conv: Conv3d(k3,s2,p1), weight read as (out=Cin, in=Cout)
                gx = _raw_conv(gy_l, w, CONV, 2, precision)
        if ctx.needs_input_grad[1]:
            lib = _lib.load()
            with torch.cuda.device(x.device):
                if kind == CONV:
                    dw = torch.zeros(27, cin, cout, device=x.device, dtype=torch.float32)
                    check(lib.casmvs_conv3d_wgrad(_ptr(xs), _ptr(gyd), _ptr(dw), B, cin, cout, Di, hi,
                                                  wi, Do, ho, wo, stride, _stream()), "conv3d_wgrad")
                    gw = dw.view(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2).contiguous()
                else:
                    # y[2i-1+k] += x[i] w[k]: the output gradient plays the strided input
                    dw = torch.zeros(27, cout, cin, device=x.device, dtype=torch.float32)
                    check(lib.casmvs_conv3d_wgrad(_ptr(gyd), _ptr(xs), _ptr(dw), B, cout, cin, Do, ho,
                                                  wo, Di, hi, wi, 2, _stream()), "conv3d_wgrad")
                    gw = dw.view(3, 3, 3, cout, cin).permute(4, 3, 0, 1, 2).contiguous()
        return gx, gw, None, None, None


class RegressFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, depth_values):
        lg = logits.detach().contiguous()
        dv = depth_values.contiguous()
        depth, conf, _, _ = ops.regress(lg, dv)
        ctx.save_for_backward(lg, dv)
        ctx.mark_non_differentiable(conf)
        return depth, conf

    @staticmethod
    def backward(ctx, gdepth, _gconf):
        lg, dv = ctx.saved_tensors
        B, D, h, w = lg.shape
        out = torch.empty_like(lg)
        with torch.cuda.device(lg.device):
            check(_lib.load().casmvs_regress_bwd(_ptr(lg), _ptr(dv), int(dv.dim() == 1),
                                                 _ptr(gdepth.contiguous()), _ptr(out), B, D, h, w,
                                                 _stream()), "regress_bwd")
        return out, None


def conv3d(x, weight, kind=CONV, stride=1, precision=FP32):
    return Conv3dFn.apply(x, weight, kind, stride, precision)


def warp_cost(feats, proj_mats, depth_values, num_groups=1):
    return WarpCostFn.apply(feats, proj_mats, depth_values, num_groups)


def regress(logits, depth_values):
    return RegressFn.apply(logits, depth_values)
