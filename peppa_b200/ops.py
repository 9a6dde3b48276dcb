"""Tensor-level wrappers over the C ABI (``include/peppa_b200.h``).

PyTorch is plumbing here: it owns device memory and the current stream; every function
below hands raw device pointers and sizes to the native library and returns freshly
allocated torch tensors.  No function in this module computes the hot path with torch ops.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from ._cabi import PB2_BF16, PB2_F16, PB2_F32, PB2_I8_PLANES, PB2_U8, check

_DTYPE_CODE = {torch.bfloat16: PB2_BF16, torch.float16: PB2_F16, torch.float32: PB2_F32, torch.uint8: PB2_U8}


# bench.py sets this to a list to time individual kernels with CUDA events on the launching stream:
# entries are (name, algorithmic flops of the launch, start event, end event).
EVENT_LOG = None


class _timed:
    def __init__(self, name, flops, device):
        self.name, self.flops, self.device = name, flops, device

    def __enter__(self):
        if EVENT_LOG is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream(self.device))

    def __exit__(self, *a):
        if EVENT_LOG is not None:
            self.e.record(torch.cuda.current_stream(self.device))
            EVENT_LOG.append((self.name, self.flops, self.s, self.e))


_CUDA_OK = False


def require_cuda(device=None) -> torch.device:
    global _CUDA_OK
    if not _CUDA_OK:        # asked once: torch.cuda.is_available() re-reads the environment on every call
        if not torch.cuda.is_available():
            raise RuntimeError("peppa_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
        _CUDA_OK = True
    if device is not None and torch.device(device).type == "cuda":
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    # a plain int (or None = NULL) for a c_void_p parameter: ctypes converts it without an intermediate object
    return t.data_ptr() if t is not None else None


def _dev_index(device):
    idx = getattr(device, "index", None)
    return idx if idx is not None else torch.cuda.current_device()


def _stream(device):
    # the raw handle of torch's current stream on that device (a few hundred ns; current_stream() builds a Stream object)
    return torch._C._cuda_getCurrentRawStream(_dev_index(device))


class _on_device:
    """``with _on_device(t.device):`` -- like torch.cuda.device(), but free when that device already is the current
    one (every launch of a one-GPU-per-process run): a batch-1k training step is launch bound, and the stock context
    manager costs several microseconds per call."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = _dev_index(device)
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *a):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


_ROW_DTYPES = (torch.bfloat16, torch.float16, torch.float32)


def as_rows(x: torch.Tensor, device=None, dtype=None) -> torch.Tensor:
    """The embedding matrix as the kernels read it: 2-D, contiguous, on the GPU, IN ITS OWN DTYPE (bf16, fp16 or
    fp32 -- nothing is rounded; other dtypes are taken as fp32), the feature dimension zero-padded to a multiple of
    64 (zero columns change neither dot products nor norms).  No copy when the input already is all of that."""
    if x.dim() != 2:
        raise ValueError(f"expected a 2-D [N, D] embedding matrix, got shape {tuple(x.shape)}")
    dev = require_cuda(device if device is not None else x.device)
    if dtype is None:
        dtype = x.dtype if x.dtype in _ROW_DTYPES else torch.float32
    y = x.detach()
    if y.device != dev or y.dtype != dtype:
        y = y.to(device=dev, dtype=dtype)
    d = y.shape[1]
    if d % 64 != 0:
        y = torch.nn.functional.pad(y, (0, 64 - d % 64))
    if not y.is_contiguous() or y.data_ptr() % 16 != 0:
        y = y.contiguous()
        if y.data_ptr() % 16 != 0:
            y = y.clone()
    return y


def as_row_pair(x: torch.Tensor, y: torch.Tensor):
    """Both matrices of a score computation through as_rows, on x's GPU and in ONE dtype (the kernels that read two
    matrices take one element type): equal dtypes are kept, mixed ones meet in fp32 (exact for bf16 and fp16).
    The feature sizes must agree like the reference's matmul demands (checked BEFORE the zero padding)."""
    if x.dim() == 2 and y.dim() == 2 and x.shape[1] != y.shape[1]:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({x.shape[0]}x{x.shape[1]} and "
                           f"{y.shape[1]}x{y.shape[0]})")
    same = x.dtype == y.dtype and x.dtype in _ROW_DTYPES
    dt = x.dtype if same else torch.float32
    xr = as_rows(x, dtype=dt)
    return xr, as_rows(y, device=xr.device, dtype=dt)


def split_f16(x_f32: torch.Tensor, rinv, side: int):
    """([n, 3 d] fp16 split operand, [n] fp32 power-of-two scales) of an fp32 matrix: the rows are multiplied by
    ``rinv`` (None = 1) and by 2^e, split into fp16 hi + lo, and laid out [hi | lo | hi] (side 0) or [hi | hi | lo]
    (side 1); the scales 2^-e take the place of ``rinv`` in the similarity kernels' epilogue."""
    n, d = x_f32.shape
    out = torch.empty(n, 3 * d, dtype=torch.float16, device=x_f32.device)
    scale = torch.empty(n, dtype=torch.float32, device=x_f32.device)
    with _on_device(x_f32.device):
        check(_cabi.lib().pb2_split_f16(_ptr(x_f32), _ptr(rinv), n, d, x_f32.stride(0), int(side), _ptr(out), 3 * d,
                                        _ptr(scale), _stream(x_f32.device)), "split_f16")
    return out, scale


def mma_pair(x: torch.Tensor, y: torch.Tensor, rinv_x=None, rinv_y=None):
    """Tensor-core operands of S = X Y^T for rows from as_row_pair, with the per-row factors the similarity kernels'
    epilogue applies: (x_op, y_op, fx, fy).  bf16 / fp16 rows are used as they are (tcgen05 kind::f16 takes both
    natively) with fx = rinv_x, fy = rinv_y; fp32 rows become the split-fp16 pair of their rinv-scaled values
    (contraction length 3 d) and fx / fy are the exact power-of-two scales of the split."""
    if x.dtype == torch.float32:
        xo, fx = split_f16(x, rinv_x, 0)
        yo, fy = split_f16(y, rinv_y, 1)
        return xo, yo, fx, fy
    return x, y, rinv_x, rinv_y


def row_norms(x: torch.Tensor):
    """(1/||x_i||, ||x_i||) in fp32 -- pig/util.py:11-12 without the divide (x bf16 / fp16 / fp32 rows)."""
    n, d = x.shape
    rinv = torch.empty(n, dtype=torch.float32, device=x.device)
    norm = torch.empty(n, dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        check(_cabi.lib().pb2_row_norms(_ptr(x), _DTYPE_CODE[x.dtype], n, d, x.stride(0), _ptr(rinv), _ptr(norm),
                                        _stream(x.device)), "row_norms")
    return rinv, norm


def _mm_code(x, y):
    if x.dtype != y.dtype:
        raise ValueError(f"operands must share one dtype, got {x.dtype} and {y.dtype} (see ops.as_row_pair)")
    return _DTYPE_CODE[x.dtype]


def pair_dot(x, y, ix=None, iy=None, rinv_x=None, rinv_y=None, want_dist=False, want_thr=False):
    """CUDA-core paired dot products; returns score [, dist = fl32(1 - score)] [, rank threshold]."""
    n = x.shape[0] if ix is None else ix.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    dist = torch.empty(n, dtype=torch.float32, device=x.device) if want_dist else None
    thr = torch.empty(n, dtype=torch.float32, device=x.device) if want_thr else None
    with _on_device(x.device):
        check(_cabi.lib().pb2_pair_dot(_ptr(x), _ptr(y), _mm_code(x, y), _ptr(ix), _ptr(iy), _ptr(rinv_x), _ptr(rinv_y), n, x.shape[1],
                                       x.stride(0), y.stride(0), _ptr(out), _ptr(dist), _ptr(thr), _stream(x.device)),
              "pair_dot")
    res = (out,) + ((dist,) if want_dist else ()) + ((thr,) if want_thr else ())
    return res if len(res) > 1 else out


def sim_diag(x, y, rinv_x=None, rinv_y=None):
    """Paired scores of row k of x with row k of y on the tensor-core pipeline (bit-identical to the
    entries the full passes compute).  Returns (score, rank threshold)."""
    n = x.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    thr = torch.empty(n, dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        check(_cabi.lib().pb2_sim_diag(_ptr(x), _ptr(y), _ptr(rinv_x), _ptr(rinv_y), n, x.shape[1], _mm_code(x, y), x.stride(0), y.stride(0),
                                       _ptr(out), _ptr(None), _ptr(thr), _stream(x.device)), "sim_diag")
    return out, thr


def sim_matrix(x, y, rinv_x=None, rinv_y=None, scale=1.0):
    """fp32 [r, c] scores; the storage row pitch is padded to a multiple of 4 floats (16-byte rows for
    the TMA stores), the returned tensor is the [:, :c] view."""
    r, c = x.shape[0], y.shape[0]
    ld = max(4, (c + 3) // 4 * 4)
    buf = torch.empty(r, ld, dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        check(_cabi.lib().pb2_sim_matrix(_ptr(x), _ptr(y), _ptr(rinv_x), _ptr(rinv_y), r, c, x.shape[1], _mm_code(x, y), x.stride(0),
                                         y.stride(0), float(scale), _ptr(buf), ld, _stream(x.device)), "sim_matrix")
    return buf[:, :c]


def sim_rank(q, g, rinv_q, rinv_g, pos_thr, pos_col, col_offset=0, rank=None):
    r, c = q.shape[0], g.shape[0]
    if rank is None:
        rank = torch.zeros(r, dtype=torch.int32, device=q.device)
    with _on_device(q.device), _timed("sim_rank", 2.0 * r * c * q.shape[1], q.device):
        check(_cabi.lib().pb2_sim_rank(_ptr(q), _ptr(g), _ptr(rinv_q), _ptr(rinv_g), _ptr(pos_thr), _ptr(pos_col), r, c,
                                       int(col_offset), q.shape[1], _mm_code(q, g), q.stride(0), g.stride(0), _ptr(rank),
                                       _stream(q.device)), "sim_rank")
    return rank


def subset_rank(scores, idx):
    """scores [G, G] fp32 (rows = queries), idx [n_samples, size] int64 on the same device -> int32 ranks."""
    n_samples, size = idx.shape
    rank = torch.empty(n_samples, size, dtype=torch.int32, device=scores.device)
    with _on_device(scores.device):
        check(_cabi.lib().pb2_subset_rank(_ptr(scores), scores.stride(0), _ptr(idx), n_samples, size, _ptr(rank),
                                          _stream(scores.device)), "subset_rank")
    return rank


def sim_grid(device) -> int:
    with _on_device(device):
        return int(_cabi.lib().pb2_sim_grid())


def gmat_alloc(rows, cols, device, dtype=torch.float16):
    """Gradient-matrix buffer with a leading dimension padded to whole TMA boxes: fp16 (any loss), or uint8 -- one
    byte per entry -- for the hinge loss, whose entries are exactly {0, 1, 2} (the kind::i8 gradient GEMMs)."""
    pad = 128 if dtype == torch.uint8 else 64
    ld = ((cols + pad - 1) // pad) * pad
    return torch.empty(rows, ld, dtype=dtype, device=device), ld


def byte_gmat_ok(dim):
    """The one-byte gradient matrix needs whole 256-column output tiles in the kind::i8 gradient GEMMs."""
    return dim % 256 == 0


def sim_hinge(x, y, rinv_x, rinv_y, diag_row, diag_col, margin, row_cnt, col_cnt, gmat=None, ld_g=0,
              row_offset=0, col_offset=0, pos_thr=None, rank=None, part=None):
    """Returns the per-CTA loss partials (fp32 [grid]).  ``part``: a caller buffer of at least sim_grid() floats that
    is ALREADY ZERO (e.g. one row of a per-step [blocks, grid] buffer cleared once): saves the per-call allocation and
    memset launch of a block-walking step."""
    r, c = x.shape[0], y.shape[0]
    n_part = sim_grid(x.device)
    if part is None:
        part = torch.empty(n_part, dtype=torch.float32, device=x.device)
        n_arg = n_part
    else:
        assert part.numel() >= n_part and part.dtype == torch.float32 and part.is_contiguous()
        n_arg = -n_part                                   # negative: "already cleared", no memset
    with _on_device(x.device), _timed("sim_hinge" + ("+rank" if rank is not None else ""), 2.0 * r * c * x.shape[1], x.device):
        check(_cabi.lib().pb2_sim_hinge(_ptr(x), _ptr(y), _ptr(rinv_x), _ptr(rinv_y), _ptr(diag_row), _ptr(diag_col), r, c,
                                        int(row_offset), int(col_offset), x.shape[1], _mm_code(x, y), x.stride(0), y.stride(0),
                                        float(margin), _ptr(part), n_arg, _ptr(row_cnt), _ptr(col_cnt), _ptr(gmat),
                                        _DTYPE_CODE[gmat.dtype] if gmat is not None else PB2_F16, int(ld_g), _ptr(pos_thr),
                                        _ptr(rank), _stream(x.device)), "sim_hinge")
    return part


def sim_lse_rows(x, y, rinv_x=None, rinv_y=None, scale=1.0, lse=None):
    """Row-wise log-sum-exp of s = scale * <x_i, y_j> over all j; optionally log-add into ``lse``."""
    r, c = x.shape[0], y.shape[0]
    lib = _cabi.lib()
    n_parts = int(lib.pb2_sim_lse_parts(c))
    pmax = torch.empty(n_parts, r, dtype=torch.float32, device=x.device)
    psum = torch.empty(n_parts, r, dtype=torch.float32, device=x.device)
    accumulate = lse is not None
    if lse is None:
        lse = torch.empty(r, dtype=torch.float32, device=x.device)
    with _on_device(x.device), _timed("sim_lse_rows", 2.0 * r * c * x.shape[1], x.device):
        st = _stream(x.device)
        check(lib.pb2_sim_lse_rows(_ptr(x), _ptr(y), _ptr(rinv_x), _ptr(rinv_y), r, c, x.shape[1], _mm_code(x, y), x.stride(0), y.stride(0),
                                   float(scale), _ptr(pmax), _ptr(psum), st), "sim_lse_rows")
        check(lib.pb2_lse_merge(_ptr(pmax), _ptr(psum), n_parts, r, _ptr(lse), int(accumulate), st), "lse_merge")
    return lse


LSE_BOTH_MAX_BOUND = 60.0 / 1.4426950408889634     # pb2_sim_lse_both refuses bound * log2(e) > 60


def sim_lse_both(x, y, bound, rinv_x=None, rinv_y=None, scale=1.0, lse_row=None, lse_col=None, rank=None):
    """Row AND column log-sum-exp of s = scale * <x_i, y_j> from one pass, for |s| <= ``bound``.

    Returns (lse_row [rows], lse_col [cols]); tensors passed in are log-added into (blocks of a larger matrix:
    ``lse_row[r0:r1]`` over the column blocks, ``lse_col[c0:c1]`` over the row blocks).

    ``rank=(rank_rinv_x, rank_rinv_y, pos_thr, row_offset, col_offset, counts)`` fuses the ranking of ``sim_rank`` into
    the same pass: ``counts[i]`` (int32, accumulated) += the columns other than the positive (global column
    ``row_offset + i``, this block starting at ``col_offset``) whose cosine reaches ``pos_thr[i]``."""
    r, c = x.shape[0], y.shape[0]
    lib = _cabi.lib()
    prow = torch.empty(int(lib.pb2_sim_lse_parts(c)), r, dtype=torch.float32, device=x.device)
    pcol = torch.empty(int(lib.pb2_sim_lse_col_parts(r)), c, dtype=torch.float32, device=x.device)
    acc_row, acc_col = lse_row is not None, lse_col is not None
    if lse_row is None:
        lse_row = torch.empty(r, dtype=torch.float32, device=x.device)
    if lse_col is None:
        lse_col = torch.empty(c, dtype=torch.float32, device=x.device)
    with _on_device(x.device), _timed("sim_lse_both", 2.0 * r * c * x.shape[1], x.device):
        st = _stream(x.device)
        if rank is None:
            check(lib.pb2_sim_lse_both(_ptr(x), _ptr(y), _ptr(rinv_x), _ptr(rinv_y), r, c, x.shape[1], _mm_code(x, y), x.stride(0),
                                       y.stride(0), float(scale), float(bound), _ptr(prow), _ptr(pcol), st), "sim_lse_both")
        else:
            rrx, rry, pos_thr, row_off, col_off, counts = rank
            check(lib.pb2_sim_lse_both_rank(_ptr(x), _ptr(y), _ptr(rinv_x), _ptr(rinv_y), r, c, x.shape[1], _mm_code(x, y),
                                            x.stride(0), y.stride(0), float(scale), float(bound), _ptr(prow), _ptr(pcol),
                                            _ptr(rrx), _ptr(rry), _ptr(pos_thr), int(row_off), int(col_off), _ptr(counts), st),
                  "sim_lse_both_rank")
        check(lib.pb2_lse_merge_const(_ptr(prow), prow.shape[0], r, float(bound), _ptr(lse_row), int(acc_row), st),
              "lse_merge_const")
        check(lib.pb2_lse_merge_const(_ptr(pcol), pcol.shape[0], c, float(bound), _ptr(lse_col), int(acc_col), st),
              "lse_merge_const")
    return lse_row, lse_col


def logit_bound(x_bf16, y_bf16, scale=1.0):
    """max_i ||x_i|| * max_j ||y_j|| * |scale| (Cauchy-Schwarz, with a 1e-4 margin for the fp32 accumulation), as a
    Python float: ONE device round trip.  NaN / inf norms give NaN / inf, which no caller accepts as a bound."""
    _, nx = row_norms(x_bf16)
    _, ny = row_norms(y_bf16)
    return float(nx.max() * ny.max()) * abs(float(scale)) * 1.0001


def lse_combine(parts):
    """parts [P, n] fp32 natural-log partial LSEs -> [n] log-sum-exp over P."""
    p_, n = parts.shape
    out = torch.empty(n, dtype=torch.float32, device=parts.device)
    with _on_device(parts.device):
        check(_cabi.lib().pb2_lse_combine(_ptr(parts), p_, n, _ptr(out), _stream(parts.device)), "lse_combine")
    return out


def sim_lse_grad(x, y, den_row, den_col, gmat, ld_g, rinv_x=None, rinv_y=None, scale=1.0):
    r, c = x.shape[0], y.shape[0]
    with _on_device(x.device), _timed("sim_lse_grad", 2.0 * r * c * x.shape[1], x.device):
        check(_cabi.lib().pb2_sim_lse_grad(_ptr(x), _ptr(y), _ptr(rinv_x), _ptr(rinv_y), _ptr(den_row), _ptr(den_col), r, c,
                                           x.shape[1], _mm_code(x, y), x.stride(0), y.stride(0), float(scale), _ptr(gmat), int(ld_g),
                                           _stream(x.device)), "sim_lse_grad")


_GG_WORKSPACE = {}        # (device, stream) -> zero-initialised stream-K workspace of pb2_grad_gemm_ws


def grad_gemm_workspace(device):
    """Stream-K workspace for the current stream of ``device`` (allocated and zeroed once; the kernel
    leaves its flag area zero, and reuse on one stream is stream-ordered)."""
    key = (device, torch._C._cuda_getCurrentRawStream(_dev_index(device)))
    ws = _GG_WORKSPACE.get(key)
    if ws is None:
        with _on_device(device):
            ws = torch.zeros(int(_cabi.lib().pb2_grad_gemm_workspace()), dtype=torch.uint8, device=device)
        _GG_WORKSPACE[key] = ws
    return ws


def grad_gemm(gmat, g_rows, g_cols, ld_g, z, transpose, alpha=1.0, out=None, accumulate=False, stream_k=True):
    m = g_cols if transpose else g_rows
    planes = gmat.dtype == torch.uint8        # one-byte G: z is the two-plane operand of rows_quant_i8, [K, 2 d] bytes
    if planes and z.dtype != torch.uint8:
        raise ValueError("a uint8 gradient matrix goes with the two-plane operand of ops.rows_quant_i8")
    d = z.shape[1] // 2 if planes else z.shape[1]
    if out is None:
        out = torch.empty(m, d, dtype=torch.float32, device=z.device)
        accumulate = False
    # small products are whole-tile anyway; skip the workspace (and its allocation under graph capture)
    ws = grad_gemm_workspace(z.device) if stream_k and m * d > 148 * 128 * 256 else None
    with _on_device(z.device), _timed("grad_gemm", 2.0 * g_rows * g_cols * d, z.device):
        check(_cabi.lib().pb2_grad_gemm_ws(_ptr(gmat), _DTYPE_CODE[gmat.dtype], g_rows, g_cols, int(ld_g),
                                           int(bool(transpose)), _ptr(z), PB2_I8_PLANES if planes else _DTYPE_CODE[z.dtype], d, z.stride(0),
                                           float(alpha), int(bool(accumulate)), _ptr(out), out.stride(0), _ptr(ws),
                                           ws.numel() if ws is not None else 0, _stream(z.device)), "grad_gemm")
    return out


def rows_scale_f16(x, rinv=None, out=None):
    """fp16(x * rinv): the embedding operand of the gradient GEMMs (x bf16 / fp16 / fp32 rows)."""
    n, d = x.shape
    if out is None:
        out = torch.empty(n, d, dtype=torch.float16, device=x.device)
    with _on_device(x.device):
        check(_cabi.lib().pb2_rows_scale_f16(_ptr(x), _DTYPE_CODE[x.dtype], _ptr(rinv), n, d, x.stride(0), _ptr(out),
                                             out.stride(0) if n else d, _stream(x.device)), "rows_scale_f16")
    return out


def rows_quant_i8(x, rinv=None, out=None):
    """Two-plane 8-bit operand [n, 2 d] uint8 = [hi (s8) | lo (u8)] of round(x * rinv * 32512): the embedding
    operand of the kind::i8 gradient GEMMs (x bf16 / fp16 / fp32 rows)."""
    n, d = x.shape
    if out is None:
        out = torch.empty(n, 2 * d, dtype=torch.uint8, device=x.device)
    with _on_device(x.device):
        check(_cabi.lib().pb2_rows_quant_i8(_ptr(x), _DTYPE_CODE[x.dtype], _ptr(rinv), n, d, x.stride(0), _ptr(out),
                                            out.stride(0) if n else 2 * d, _stream(x.device)), "rows_quant_i8")
    return out


def ipc_export(t: torch.Tensor):
    """(64-byte CUDA IPC handle of the allocation ``t`` lies in, byte offset of ``t`` inside it)."""
    handle = C.create_string_buffer(64)
    off = C.c_int64(0)
    with _on_device(t.device):
        check(_cabi.lib().pb2_ipc_export(_ptr(t), handle, C.byref(off)), "ipc_export")
    return handle.raw, int(off.value)


def ipc_open(handle: bytes, offset: int, device):
    """Map a peer rank's allocation on ``device``; returns (base, pointer) as ints (base is for ipc_close)."""
    base, ptr = C.c_void_p(0), C.c_void_p(0)
    with _on_device(device):
        check(_cabi.lib().pb2_ipc_open(C.c_char_p(handle), int(offset), C.byref(base), C.byref(ptr)), "ipc_open")
    return int(base.value), int(ptr.value)


def ipc_close(base: int, device):
    with _on_device(device):
        check(_cabi.lib().pb2_ipc_close(C.c_void_p(base)), "ipc_close")


def peer_reduce(ptrs, out: torch.Tensor):
    """out = sum over the device pointers ``ptrs`` (ints; local or peer memory, fp32, out.numel() elements each)."""
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    with _on_device(out.device), _timed("peer_reduce", float(len(ptrs) + 1) * out.numel() * 4, out.device):
        check(_cabi.lib().pb2_peer_reduce(arr, len(ptrs), out.numel(), _ptr(out), _stream(out.device)), "peer_reduce")
    return out


def hinge_finish(p, x, y, rinv_x, rinv_y, row_cnt, col_cnt, coef_host=1.0, coef_dev=None):
    rows, d = x.shape
    grad = torch.empty(rows, d, dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        check(_cabi.lib().pb2_hinge_finish(_ptr(p), p.stride(0), _ptr(x), _ptr(y), _mm_code(x, y), _ptr(rinv_x), _ptr(rinv_y),
                                           _ptr(row_cnt), _ptr(col_cnt), rows, d, x.stride(0), y.stride(0), float(coef_host),
                                           _ptr(coef_dev), _ptr(grad), grad.stride(0), _stream(x.device)), "hinge_finish")
    return grad


def milnce_finish(p, y, coef_host=1.0, coef_dev=None):
    rows, d = y.shape
    grad = torch.empty(rows, d, dtype=torch.float32, device=y.device)
    with _on_device(y.device):
        check(_cabi.lib().pb2_milnce_finish(_ptr(p), p.stride(0), _ptr(y), _DTYPE_CODE[y.dtype], rows, d, y.stride(0), float(coef_host),
                                            _ptr(coef_dev), _ptr(grad), grad.stride(0), _stream(y.device)), "milnce_finish")
    return grad


def milnce_finish_k(p, y, w, rows, group, y_div, coef_host=1.0):
    """MIL-NCE finish with K candidates per clip: see pb2_milnce_finish_k."""
    d = y.shape[1]
    grad = torch.empty(rows, d, dtype=torch.float32, device=y.device)
    with _on_device(y.device):
        check(_cabi.lib().pb2_milnce_finish_k(_ptr(p), p.stride(0), _ptr(y), _DTYPE_CODE[y.dtype], _ptr(w), rows, int(group), int(y_div), d,
                                              y.stride(0), float(coef_host), _ptr(None), _ptr(grad), grad.stride(0),
                                              _stream(y.device)), "milnce_finish_k")
    return grad


def project_normalize(x_bf16, w_bf16, bias=None, eps=1e-12):
    """bf16( normalize(x W^T + b) ) with the fp32 1/||row|| of the rounded rows and ||y||: the encoder tail
    (pig/models.py:96-109, :130-150) in one launch.  x [rows, n_in], W [n_out, n_in] (nn.Linear layout)."""
    rows, n_in = x_bf16.shape
    n_out = w_bf16.shape[0]
    dev = x_bf16.device
    out = torch.empty(rows, n_out, dtype=torch.bfloat16, device=dev)
    rinv = torch.empty(rows, dtype=torch.float32, device=dev)
    norm = torch.empty(rows, dtype=torch.float32, device=dev)
    with _on_device(dev), _timed("project_normalize", 2.0 * rows * n_in * n_out, dev):
        check(_cabi.lib().pb2_project_normalize(_ptr(x_bf16), _ptr(w_bf16), _ptr(bias), rows, n_in, n_out, x_bf16.stride(0),
                                                w_bf16.stride(0), float(eps), _ptr(out), out.stride(0) if rows else n_out,
                                                _ptr(rinv), _ptr(norm), _stream(dev)), "project_normalize")
    return out, rinv, norm


_STEP_WORKSPACE = {}      # (device, stream, n, d, dtype) -> uint8 workspace, reused across steps ON THAT STREAM


def tag_rinv(rows: torch.Tensor, rinv: torch.Tensor) -> torch.Tensor:
    """Remember on a tensor of embedding rows the fp32 1/||row|| its producer computed (the encoder tail emits it
    with the rows).  The tag names the tensor's version counter, so an in-place edit voids it."""
    rows._pb2_rinv = (rinv, rows._version)
    return rows


def known_rinv(original: torch.Tensor, rows: torch.Tensor):
    """The tagged 1/||row|| of ``original`` if ``rows`` (what as_rows made of it) still IS that tensor's memory --
    same storage, dtype and shape, untouched since the tag was set -- else None (the caller computes the norms)."""
    tag = getattr(original, "_pb2_rinv", None)
    if tag is None:
        return None
    rinv, version = tag
    if (original._version != version or rows.data_ptr() != original.data_ptr() or rows.dtype != original.dtype
            or rows.shape != original.shape or rinv.device != rows.device or rinv.shape != (rows.shape[0],)):
        return None
    return rinv


def rinv_of(original, rows):
    """1/||row|| of ``rows``: the producer's tagged vector when it is still valid, else pb2_row_norms."""
    r = known_rinv(original, rows)
    return r if r is not None else row_norms(rows)[0]


def hinge_step(vb, ab, margin, grad_dtype=torch.float32, rinv_v=None, rinv_a=None):
    """Whole TripletLoss forward + gradients for one gradient-matrix block (n <= 32768): one C call, four
    launches (prep, fused similarity/hinge pass, both gradient GEMMs, finish).  vb / ab: rows from as_row_pair
    (bf16, fp16 or fp32).  Returns (loss 0-d fp32, grads [2, n, d] in ``grad_dtype``: dV then dA)."""
    n, d = vb.shape
    dev = vb.device
    lib = _cabi.lib()
    code = _mm_code(vb, ab)
    capturing = torch.cuda.is_current_stream_capturing()
    # reuse is stream-ordered, so the cache is per stream: two steps of one shape on two streams never share
    # scratch (a workspace handed to a graph capture belongs to that graph and is not cached)
    key = (dev, torch._C._cuda_getCurrentRawStream(_dev_index(dev)), n, d, code)
    ws = None if capturing else _STEP_WORKSPACE.get(key)
    if ws is None:
        with _on_device(dev):
            ws = torch.empty(int(lib.pb2_hinge_step_workspace(n, d, code)), dtype=torch.uint8, device=dev)
        if not capturing:
            if len(_STEP_WORKSPACE) > 8:
                _STEP_WORKSPACE.clear()
            _STEP_WORKSPACE[key] = ws
    grads = torch.empty(2, n, d, dtype=grad_dtype, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    with _on_device(dev), _timed("hinge_step (4 kernels)", 6.0 * n * n * d, dev):
        check(lib.pb2_hinge_step(_ptr(vb), _ptr(ab), code, n, d, vb.stride(0), ab.stride(0), float(margin), _ptr(ws), ws.numel(),
                                 _ptr(loss), _ptr(grads[0]), _ptr(grads[1]), _DTYPE_CODE[grad_dtype], _ptr(rinv_v), _ptr(rinv_a),
                                 _stream(dev)), "hinge_step")
    return loss, grads


def hinge_forward(vb, ab, margin, rinv_v=None, rinv_a=None, want_state=True):
    """Forward half of the TripletLoss step for one gradient-matrix block (n <= 32768): one C call, three launches
    (prep, fused similarity / hinge pass, both gradient GEMMs with the scalar loss folded beside them).  Returns
    (loss 0-d fp32, state): ``state`` is a fresh uint8 tensor holding what ``hinge_backward`` needs (both products,
    1/||row||, the indicator counts); the scratch is the per-stream workspace of ``hinge_step``.
    ``want_state=False``: the loss alone (prep, the pass without a gradient matrix, the fold); state is None."""
    n, d = vb.shape
    dev = vb.device
    lib = _cabi.lib()
    code = _mm_code(vb, ab)
    capturing = torch.cuda.is_current_stream_capturing()
    key = (dev, torch._C._cuda_getCurrentRawStream(_dev_index(dev)), n, d, code)
    ws = None if capturing else _STEP_WORKSPACE.get(key)
    with _on_device(dev):
        if ws is None:
            ws = torch.empty(int(lib.pb2_hinge_step_workspace(n, d, code)), dtype=torch.uint8, device=dev)
            if not capturing:
                if len(_STEP_WORKSPACE) > 8:
                    _STEP_WORKSPACE.clear()
                _STEP_WORKSPACE[key] = ws
        state = torch.empty(int(lib.pb2_hinge_state_bytes(n, d)), dtype=torch.uint8, device=dev) if want_state else None
        loss = torch.empty((), dtype=torch.float32, device=dev)
        with _timed("hinge_forward (3 kernels)", (6.0 if want_state else 2.0) * n * n * d, dev):
            check(lib.pb2_hinge_forward(_ptr(vb), _ptr(ab), code, n, d, vb.stride(0), ab.stride(0), float(margin), _ptr(ws),
                                        ws.numel(), _ptr(state), state.numel() if want_state else 0, _ptr(loss), _ptr(rinv_v),
                                        _ptr(rinv_a), _stream(dev)), "hinge_forward")
    return loss, state


def hinge_backward(state, vb, ab, grad_out, out_dtype=torch.float32):
    """Backward half: (dV, dA) [n, d] in ``out_dtype`` = the mean hinge loss's gradients times the 0-d fp32 device tensor
    ``grad_out``, multiplied in fp32 before the rounding; one launch, two fresh contiguous tensors."""
    n, d = vb.shape
    dev = vb.device
    with _on_device(dev):
        g0 = torch.empty(n, d, dtype=out_dtype, device=dev)
        g1 = torch.empty(n, d, dtype=out_dtype, device=dev)
        check(_cabi.lib().pb2_hinge_backward(_ptr(state), state.numel(), _ptr(vb), _ptr(ab), _mm_code(vb, ab), n, d, vb.stride(0),
                                             ab.stride(0), _ptr(grad_out), _ptr(g0), _ptr(g1), _DTYPE_CODE[out_dtype],
                                             _stream(dev)), "hinge_backward")
    return g0, g1


def scale_pair(x0, x1, coef, out_dtype=torch.float32):
    """(x0 * coef, x1 * coef) rounded to ``out_dtype`` for fp32 ``x`` and a 0-d fp32 device tensor ``coef``: one
    launch, fresh contiguous results (the scale is applied in fp32, the rounding comes last)."""
    assert x0.dtype == x1.dtype == torch.float32 and x0.shape == x1.shape and x0.is_contiguous() and x1.is_contiguous()
    y0, y1 = torch.empty_like(x0, dtype=out_dtype), torch.empty_like(x1, dtype=out_dtype)
    with _on_device(x0.device):
        check(_cabi.lib().pb2_scale_pair(_ptr(x0), _ptr(x1), x0.numel(), _DTYPE_CODE[out_dtype], _ptr(coef), _ptr(y0), _ptr(y1),
                                         _stream(x0.device)), "scale_pair")
    return y0, y1


def sum_partials(part, alpha=1.0):
    out = torch.empty((), dtype=torch.float32, device=part.device)
    with _on_device(part.device):
        check(_cabi.lib().pb2_sum_partials(_ptr(part), part.numel(), float(alpha), _ptr(out), _stream(part.device)),
              "sum_partials")
    return out


def hinge_loss_terms(out, partials=None, diag=None, cnt=None, margin=0.0, alpha=1.0, accumulate=True):
    """out (=|+=) alpha * (sum partials + sum (margin - diag) * cnt): completes the hinge loss."""
    with _on_device(out.device):
        check(_cabi.lib().pb2_hinge_loss_terms(_ptr(partials), partials.numel() if partials is not None else 0, _ptr(diag),
                                               _ptr(cnt), cnt.numel() if cnt is not None else 0, float(margin), float(alpha),
                                               _ptr(out), int(bool(accumulate)), _stream(out.device)), "hinge_loss_terms")
    return out


def milnce_loss(lse_row, lse_col, diag):
    n = lse_row.shape[0]
    den = torch.empty(n, dtype=torch.float32, device=lse_row.device)
    out = torch.empty((), dtype=torch.float32, device=lse_row.device)
    with _on_device(lse_row.device):
        check(_cabi.lib().pb2_milnce_loss(_ptr(lse_row), _ptr(lse_col), _ptr(diag), n, _ptr(den), _ptr(out),
                                          _stream(lse_row.device)), "milnce_loss")
    return out, den


def contrastive_matrix(m, margin, want_grad, coef_dev=None):
    """pig/loss.py:41-48 on a materialised fp32 square matrix; returns (loss, dM or None)."""
    n = m.shape[0]
    n_part = sim_grid(m.device) * 4
    work = torch.empty(n_part + 2 * n, dtype=torch.float32, device=m.device)  # partials | int32 counts
    grad = torch.empty(n, n, dtype=torch.float32, device=m.device) if want_grad else None
    with _on_device(m.device):
        check(_cabi.lib().pb2_contrastive_matrix(_ptr(m), n, m.stride(0), float(margin), _ptr(work), n_part, _ptr(grad),
                                                 n if want_grad else 0, 1.0 / float(n) ** 2, _ptr(coef_dev),
                                                 _stream(m.device)), "contrastive_matrix")
        loss = sum_partials(work[:n_part], 1.0 / float(n) ** 2)
    return loss, grad


def triplet_score(anchor, positive, negative, ia=None, ip=None, in_=None, discrete=True):
    """anchor/positive/negative: 2-D, same dtype (bf16/f16/f32), same leading dimension."""
    t = anchor.shape[0] if ia is None else ia.shape[0]
    d = anchor.shape[1]
    out = torch.empty(t, dtype=torch.float32, device=anchor.device)
    code = _DTYPE_CODE[anchor.dtype]
    assert positive.dtype == anchor.dtype and negative.dtype == anchor.dtype
    ld = anchor.stride(0) if anchor.shape[0] > 1 else d
    for m in (positive, negative):
        if m.shape[0] > 1 and m.stride(0) != ld:
            raise ValueError("triplet_score: operands must share a leading dimension")
    with _on_device(anchor.device), _timed("triplet_score", float(t) * (3 * d * anchor.element_size() + 4), anchor.device):
        check(_cabi.lib().pb2_triplet_score(_ptr(anchor), _ptr(positive), _ptr(negative), _ptr(ia), _ptr(ip), _ptr(in_), t, d,
                                            ld, code, int(bool(discrete)), _ptr(out), _stream(anchor.device)),
              "triplet_score")
    return out
