"""CPU emulation of the kernel entry points used by peppa_b200.gallery.GalleryStep.  TEST ONLY.

Each function restates the contract of the C-ABI call of the same name (include/peppa_b200.h) with
plain torch-CPU arithmetic, so the multi-rank orchestration (sharding, collectives, merges) can be
exercised under gloo with world_size 2 on a box without GPUs.  Never imported by the product.
"""
import torch


def gmat_alloc(rows, cols, device):
    ld = ((cols + 63) // 64) * 64
    return torch.zeros(rows, ld, dtype=torch.float16, device=device), ld


def row_norms(x):
    n = torch.linalg.vector_norm(x.float(), dim=1)
    return 1.0 / n, n


def sim_diag(x, y, rinv_x=None, rinv_y=None):
    """(paired score, rank threshold).  The emulation keeps the threshold as the distance itself and
    sim_hinge below applies the original test fl32(1 - s) < dist."""
    s = (x.float() * y.float()).sum(dim=1) * rinv_x * rinv_y
    return s, 1.0 - s


def pair_dot(x, y, ix=None, iy=None, rinv_x=None, rinv_y=None, want_dist=False, want_thr=False):
    s = (x.float() * y.float()).sum(dim=1)
    if rinv_x is not None:
        s = s * rinv_x * rinv_y
    return s


def rows_scale_f16(x, rinv=None):
    xf = x.float()
    return (xf * rinv[:, None] if rinv is not None else xf).half()


def sim_hinge(x, y, rinv_x, rinv_y, diag_row, diag_col, margin, row_cnt, col_cnt, gmat=None, ld_g=0, row_offset=0,
              col_offset=0, pos_thr=None, rank=None):
    S = (x.float() @ y.float().T) * rinv_x[:, None] * rinv_y[None, :]
    r, c = S.shape
    off = (torch.arange(r)[:, None] + row_offset) != (torch.arange(c)[None, :] + col_offset)
    zc = margin + S - diag_col[None, :]
    zr = margin + S - diag_row[:, None]
    ic, ir = off & (zc >= 0), off & (zr >= 0)
    row_cnt += ir.sum(dim=1).to(torch.int32)
    col_cnt += ic.sum(dim=0).to(torch.int32)
    if gmat is not None:
        gmat[:r, :c] = (ic.float() + ir.float()).half()
    if rank is not None:
        rank += (off & ((1.0 - S) < pos_thr[:, None])).sum(dim=1).to(torch.int32)
    return ((ic.float() + ir.float()) * S).sum().reshape(1)      # completed by hinge_loss_terms


def hinge_loss_terms(out, partials=None, diag=None, cnt=None, margin=0.0, alpha=1.0, accumulate=True):
    t = torch.zeros((), dtype=torch.float64)
    if partials is not None:
        t = t + partials.double().sum()
    if diag is not None:
        t = t + ((margin - diag).double() * cnt.double()).sum()
    r = (t * alpha).float()
    if accumulate:
        out += r
    else:
        out.copy_(r)
    return out


def grad_gemm(gmat, g_rows, g_cols, ld_g, z, transpose, alpha=1.0, out=None, accumulate=False):
    G = gmat[:g_rows, :g_cols].float()
    res = alpha * ((G.T if transpose else G) @ z.float())
    if out is None:
        return res
    if accumulate:
        out += res
    else:
        out.copy_(res)
    return out


def hinge_finish(p, x, y, rinv_x, rinv_y, row_cnt, col_cnt, coef_host=1.0, coef_dev=None):
    g = p + (-(row_cnt + col_cnt).float() * rinv_y)[:, None] * y.float()
    xh = x.float() * rinv_x[:, None]
    return coef_host * rinv_x[:, None] * (g - xh * (g * xh).sum(dim=1, keepdim=True))


# ---- MIL-NCE entry points (natural-log statistics, fp16 gradient matrix scaled by 2^13) ---------------
def sim_lse_rows(x, y, rinv_x=None, rinv_y=None, scale=1.0, lse=None):
    s = torch.logsumexp((x.float() @ y.float().T) * scale, dim=1)
    return s if lse is None else torch.logaddexp(lse, s)


LSE_BOTH_MAX_BOUND = 60.0 / 1.4426950408889634


def logit_bound(x, y, scale=1.0):
    return float(torch.linalg.vector_norm(x.float(), dim=1).max() * torch.linalg.vector_norm(y.float(), dim=1).max()) \
        * abs(float(scale)) * 1.0001


def sim_lse_both(x, y, bound, rinv_x=None, rinv_y=None, scale=1.0, lse_row=None, lse_col=None, rank=None):
    s = (x.float() @ y.float().T) * scale
    assert float(s.abs().max()) <= bound
    if rank is not None:        # the fused ranking (pb2_sim_lse_both_rank); sim_diag above keeps the distance as threshold
        rrx, rry, pos_dist, row_off, col_off, counts = rank
        cos = (x.float() @ y.float().T) * rrx[:, None] * rry[None, :]
        closer = (1.0 - cos) < pos_dist[:, None]
        rel = torch.arange(x.shape[0]) + (row_off - col_off)
        ok = (rel >= 0) & (rel < y.shape[0])
        closer[torch.nonzero(ok).flatten(), rel[ok]] = False          # the positive never counts against itself
        counts += closer.sum(dim=1).to(counts.dtype)
    e = torch.exp2(s * 1.4426950408889634 - bound * 1.4426950408889634)      # the kernel's fixed-shift sums
    row = (bound * 1.4426950408889634 + torch.log2(e.sum(dim=1))) * 0.6931471805599453
    col = (bound * 1.4426950408889634 + torch.log2(e.sum(dim=0))) * 0.6931471805599453
    if lse_row is None:
        lse_row = row
    else:
        lse_row.copy_(torch.logaddexp(lse_row, row))
    if lse_col is None:
        lse_col = col
    else:
        lse_col.copy_(torch.logaddexp(lse_col, col))
    return lse_row, lse_col


def lse_combine(parts):
    return torch.logsumexp(parts, dim=0)


def milnce_loss(lse_row, lse_col, diag):
    den = torch.logaddexp(lse_row, lse_col)
    return (den - diag).mean(), den


def sim_lse_grad(x, y, den_row, den_col, gmat, ld_g, rinv_x=None, rinv_y=None, scale=1.0):
    s = (x.float() @ y.float().T) * scale
    r, c = s.shape
    gmat[:r, :c] = ((torch.exp(s - den_row[:, None]) + torch.exp(s - den_col[None, :])) * 8192.0).half()


def milnce_finish(p, y, coef_host=1.0, coef_dev=None):
    return coef_host * (p / 8192.0 - y.float())
