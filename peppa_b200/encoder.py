"""Encoder tail: ``project`` + ``F.normalize`` of the reference encoders as one kernel.

Both reference encoders end with ``nn.Linear(n_features, 512)`` followed by
``nn.functional.normalize(x, p=2, dim=1)`` (pig/models.py:96-109 ``Wav2VecEncoder``, :130-150
``R3DEncoder``), and the loss then re-normalises the result (pig/util.py:11-12).  ``ProjectNormalize``
fuses the projection GEMM, the bias, the L2 normalisation and the bf16 cast into one tcgen05 launch
(``pb2_project_normalize``) and also emits the fp32 ``1/||row||`` of the rounded embeddings, i.e. exactly
what the scoring kernels take as ``rinv`` -- no fp32 round trip, no separate norm pass.

``ProjectNormalize`` keeps ``nn.Linear``'s parameter names (``weight`` [out, in], ``bias`` [out]) so a
checkpoint's ``project.*`` tensors load into it unchanged.  The forward GEMM is the hand-written kernel;
the backward (dX = dY W, dW = dY^T X) is two plain dense GEMMs with no epilogue to fuse and no operand the scoring
kernels share -- library-GEMM territory by this project's own rule (cuBLAS for plain library GEMMs) -- so it stays
torch.matmul; the scoring path never differentiates through it with anything but these two products.
"""
from __future__ import annotations

import math

import torch
import torch.nn

from . import ops


def _pad_cols(t: torch.Tensor, mult: int = 64) -> torch.Tensor:
    d = t.shape[1]
    return t if d % mult == 0 else torch.nn.functional.pad(t, (0, mult - d % mult))


class _ProjectNormalizeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
            raise RuntimeError(f"project_normalize expects x [N, in] and weight [out, in]; got {tuple(x.shape)}, {tuple(weight.shape)}")
        n_out = weight.shape[0]
        if n_out % 64 != 0 or n_out > 512:
            raise RuntimeError("project_normalize: out_features must be a multiple of 64 and at most 512")
        dev = ops.require_cuda(x.device)
        xb = _pad_cols(x.detach().to(device=dev, dtype=torch.bfloat16)).contiguous()
        wb = _pad_cols(weight.detach().to(device=dev, dtype=torch.bfloat16)).contiguous()
        b = bias.detach().to(device=dev, dtype=torch.float32).contiguous() if bias is not None else None
        out, rinv, norm = ops.project_normalize(xb, wb, b, eps)
        ctx.save_for_backward(xb, wb, out, norm)
        ctx.meta = (x.dtype, x.device, x.shape[1], weight.dtype, weight.device, bias is not None and bias.dtype, eps)
        if x.device != out.device:          # results follow the input's device, like the loss / metric drop-ins
            out, rinv = out.to(x.device), rinv.to(x.device)
        ctx.mark_non_differentiable(rinv)
        return out, rinv

    @staticmethod
    def backward(ctx, g_out, _g_rinv):
        xb, wb, out, norm = ctx.saved_tensors
        xd, xdev, n_in, wd, wdev, bd, eps = ctx.meta
        g = g_out.to(device=out.device, dtype=torch.float32)      # `out` was saved on the GPU
        e = out.float()
        # Jacobian of y -> y / max(||y||, eps):  (g - e <g, e>) / ||y||
        dy = (g - e * (g * e).sum(dim=1, keepdim=True)) / norm.clamp_min(eps).unsqueeze(1)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (dy @ wb.float())[:, :n_in].to(device=xdev, dtype=xd)
        if ctx.needs_input_grad[1]:
            gw = (dy.t() @ xb.float())[:, :n_in].to(device=wdev, dtype=wd)
        if ctx.needs_input_grad[2]:
            gb = dy.sum(dim=0).to(device=wdev, dtype=bd)
        return gx, gw, gb, None


def project_normalize(x, weight, bias=None, eps=1e-12, return_rinv=False):
    """``normalize(linear(x, weight, bias), p=2, dim=1)`` as bf16 (fp32 accumulate); optionally also the fp32
    ``1/||row||`` of the rounded rows (the ``rinv`` operand of the scoring kernels)."""
    out, rinv = _ProjectNormalizeFn.apply(x, weight, bias, float(eps))
    # the rows travel with their 1/||row||: TripletLoss / recall_* / GalleryStep take it from the tag instead of
    # recomputing the norms of embeddings this kernel just normalised (pig/util.py:11-12 after pig/models.py:109)
    ops.tag_rinv(out, rinv)
    return (out, rinv) if return_rinv else out


class ProjectNormalize(torch.nn.Module):
    """Drop-in for ``Compose([self.project, normalize])`` at the end of ``Wav2VecEncoder.forward`` /
    ``R3DEncoder.encode`` (pig/models.py:105-109, :141-150): same parameters as ``nn.Linear``."""

    def __init__(self, in_features: int, out_features: int = 512, bias: bool = True, eps: float = 1e-12):
        super().__init__()
        self.in_features, self.out_features, self.eps = in_features, out_features, eps
        self.weight = torch.nn.Parameter(torch.empty(out_features, in_features))
        self.bias = torch.nn.Parameter(torch.empty(out_features)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):          # nn.Linear's initialisation
        torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.in_features)
            torch.nn.init.uniform_(self.bias, -bound, bound)

    @classmethod
    def from_linear(cls, linear: torch.nn.Linear, eps: float = 1e-12):
        m = cls(linear.in_features, linear.out_features, linear.bias is not None, eps)
        m.load_state_dict(linear.state_dict())
        return m.to(linear.weight.device)

    def forward(self, x, return_rinv=False):
        return project_normalize(x, self.weight, self.bias, self.eps, return_rinv)
