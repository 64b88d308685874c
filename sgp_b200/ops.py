"""Thin torch-tensor wrappers over the C ABI (device pointers, strides, current stream).

All tensors must live on a CUDA device; nothing here computes on the host.  `view3` arguments are
[T, N, F] float32 views whose last dimension is contiguous (feature blocks of a wider buffer are
fine: their t / n strides are passed through).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import ACT_CODES, check, load


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _call(dev, name: str, *args) -> None:
    """One C-ABI launch with `dev` as the CURRENT CUDA device (kernel launches, cudaFuncSetAttribute
    and the cub calls inside the library all act on the current device, which need not be the
    tensors' device when the caller works on several GPUs)."""
    with torch.cuda.device(dev):
        check(getattr(load(), name)(*args), name)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.SgpError("sgp_b200 kernels need CUDA tensors (there is no CPU path)")


def _check_view3(t: torch.Tensor, name: str):
    if t.dim() != 3 or t.dtype != torch.float32 or (t.size(-1) > 1 and t.stride(-1) != 1):
        raise ValueError(f"{name}: expected a float32 [T, N, F] view with contiguous features, got "
                         f"{tuple(t.shape)} {t.dtype} strides {t.stride()}")


@dataclass
class Csr:
    """Normalised shift operator in CSR (device, int32 indices)."""
    rowptr: torch.Tensor
    col: torch.Tensor
    val: torch.Tensor
    num_nodes: int

    @property
    def nnz(self) -> int:
        return int(self.col.numel())


def csr_build(edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], num_nodes: int,
              flags: int) -> Csr:
    """Device CSR from a [2, E] int64 edge list in the reference convention ([0]=col, [1]=row)."""
    _require_cuda(edge_index, edge_weight)
    lib = load()
    dev = edge_index.device
    ei = edge_index.to(torch.int64).contiguous()
    E = int(ei.size(1))
    w = None if edge_weight is None else edge_weight.to(torch.float32).contiguous()
    if w is not None and w.numel() != E:
        raise ValueError(f"edge_weight has {w.numel()} entries for {E} edges")
    N = int(num_nodes)
    cap = (2 * E if flags & _lib.CSR_SYMMETRIZE else E) + (N if flags & _lib.CSR_SET_DIAG else 0)
    cap = max(cap, 1)
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    col = torch.empty(cap, dtype=torch.int32, device=dev)
    val = torch.empty(cap, dtype=torch.float32, device=dev)
    ws_bytes = int(lib.sgp_csr_build_workspace_bytes(E, N, flags))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    nnz = ctypes.c_int64(0)
    src, dst = (ei[0], ei[1]) if E else (None, None)
    _call(dev, "sgp_csr_build", _p(src), _p(dst), _p(w), E, N, flags, _p(rowptr), _p(col), _p(val), cap,
                            ctypes.byref(nnz), _p(ws), ws_bytes, _stream(dev))
    n = int(nnz.value)
    return Csr(rowptr, col[:n], val[:n], N)


def reservoir_pack(w_ih: torch.Tensor, w_hh: torch.Tensor) -> torch.Tensor:
    _require_cuda(w_ih, w_hh)
    H, Fin = w_ih.shape
    rows = int(load().sgp_reservoir_pack_rows(Fin, H))
    out = torch.empty(rows, H, dtype=torch.float32, device=w_ih.device)
    _call(w_ih.device, "sgp_reservoir_pack", _p(w_ih.contiguous()), _p(w_hh.contiguous()), Fin, H, _p(out),
                                    _stream(w_ih.device))
    return out


def reservoir_scan(x: torch.Tensor, wpack: torch.Tensor, bias: torch.Tensor, alpha: float,
                   activation: str, h_state: torch.Tensor, out: torch.Tensor) -> None:
    """One layer over a chunk: x [Tc,N,Fin] view, h_state [N,H] in/out, out [Tc,N,H] view."""
    _require_cuda(x, wpack, bias, h_state, out)
    _check_view3(x, "x")
    _check_view3(out, "out")
    Tc, N, Fin = x.shape
    H = int(bias.numel())
    assert out.shape == (Tc, N, H) and h_state.shape == (N, H) and h_state.is_contiguous()
    a = float(alpha)
    _call(x.device, "sgp_reservoir_scan", _p(x), x.stride(0), x.stride(1), Fin, _p(wpack), _p(bias),
                                    a, float(1.0 - a), ACT_CODES[activation], _p(h_state),
                                    _p(out), out.stride(0), out.stride(1), Tc, N, H,
                                    _stream(x.device))


def reservoir_scan_multi(x: torch.Tensor, w_ih: list, w_hh: list, bias: list, alphas: list, activation: str,
                         h_state: torch.Tensor, out: torch.Tensor) -> None:
    """All layers of a small reservoir (H in {16, 32, 64}) over a chunk in ONE launch: x [Tc,N,Fin],
    per-layer device weights as the reference holds them (w_ih[l] [H,Fin_l], w_hh[l] [H,H], bias[l]
    [H]), h_state [L,N,H] in/out, out [Tc,N,>=L*H] view."""
    _require_cuda(x, h_state, out, *w_ih, *w_hh, *bias)
    _check_view3(x, "x")
    _check_view3(out, "out")
    Tc, N, Fin = x.shape
    L, H = len(w_hh), int(w_hh[0].shape[0])
    assert h_state.shape == (L, N, H) and h_state.is_contiguous() and out.shape[:2] == (Tc, N) and out.shape[2] >= L * H
    assert all(t.is_contiguous() and t.dtype == torch.float32 for t in (*w_ih, *w_hh, *bias))
    ptrs = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts])      # noqa: E731
    al = (ctypes.c_float * L)(*[float(a) for a in alphas])
    _call(x.device, "sgp_reservoir_scan_multi", _p(x), x.stride(0), x.stride(1), Fin, ptrs(w_ih), ptrs(w_hh),
          ptrs(bias), al, ACT_CODES[activation], _p(h_state), _p(out), out.stride(0), out.stride(1), Tc, N, H, L,
          _stream(x.device))


def reservoir_tc_pack(w_hh: torch.Tensor) -> torch.Tensor:
    """W_hh [H, H] -> tf32 hi / lo images for the tensor-core scan (H in {128, 256})."""
    _require_cuda(w_hh)
    H = int(w_hh.shape[0])
    out = torch.empty(2 * H * H, dtype=torch.float32, device=w_hh.device)
    _call(w_hh.device, "sgp_reservoir_tc_pack", _p(w_hh.contiguous()), H, _p(out), _stream(w_hh.device))
    return out


def reservoir_scan_tc(x: torch.Tensor, wimg: torch.Tensor, w_ih: torch.Tensor, bias: torch.Tensor,
                      alpha: float, activation: str, h_state: torch.Tensor, out: torch.Tensor,
                      err: torch.Tensor, checksum: Optional[torch.Tensor] = None) -> None:
    """Tensor-core variant of reservoir_scan (same views); `err` is a device int32 flag, `checksum`
    an optional device float64 scalar that receives the sum of everything written to `out`."""
    _require_cuda(x, wimg, w_ih, bias, h_state, out, err)
    _check_view3(x, "x")
    _check_view3(out, "out")
    Tc, N, Fin = x.shape
    H = int(bias.numel())
    assert out.shape == (Tc, N, H) and h_state.shape == (N, H) and h_state.is_contiguous()
    a = float(alpha)
    _call(x.device, "sgp_reservoir_scan_tc", _p(x), x.stride(0), x.stride(1), Fin, _p(wimg), _p(w_ih), _p(bias),
                                       a, float(1.0 - a), ACT_CODES[activation], _p(h_state), _p(out),
                                       out.stride(0), out.stride(1), Tc, N, H, _p(err), _p(checksum), _stream(x.device))


def reservoir_tc16_pack(w_hh: torch.Tensor):
    """W_hh [H, H] -> (fp16 hi / lo images for the fp16x3 tensor-core scan, the power-of-two weight scale)."""
    _require_cuda(w_hh)
    H = int(w_hh.shape[0])
    wmax = float(w_hh.abs().max())
    scale = float(2.0 ** np.floor(np.log2(16384.0 / wmax))) if wmax > 0 else 1.0
    out = torch.empty(2 * H * H, dtype=torch.float16, device=w_hh.device)
    _call(w_hh.device, "sgp_reservoir_tc16_pack", _p(w_hh.contiguous()), H, scale, _p(out), _stream(w_hh.device))
    return out, scale


def reservoir_scan_tc16(x: torch.Tensor, wimg: torch.Tensor, w_scale: float, w_ih: torch.Tensor, bias: torch.Tensor,
                        alpha: float, h_state: torch.Tensor, out: torch.Tensor, err: torch.Tensor,
                        checksum: Optional[torch.Tensor] = None) -> None:
    """fp16x3 tensor-core scan (tanh, states within [-1, 1]); same views as reservoir_scan_tc."""
    _require_cuda(x, wimg, w_ih, bias, h_state, out, err)
    _check_view3(x, "x")
    _check_view3(out, "out")
    Tc, N, Fin = x.shape
    H = int(bias.numel())
    assert out.shape == (Tc, N, H) and h_state.shape == (N, H) and h_state.is_contiguous()
    a = float(alpha)
    _call(x.device, "sgp_reservoir_scan_tc16", _p(x), x.stride(0), x.stride(1), Fin, _p(wimg), float(w_scale), _p(w_ih),
          _p(bias), a, float(1.0 - a), _p(h_state), _p(out), out.stride(0), out.stride(1), Tc, N, H, _p(err),
          _p(checksum), _stream(x.device))


def spmm(csr: Csr, src: torch.Tensor, dst: torch.Tensor, row_order: Optional[torch.Tensor] = None,
         n_rows: Optional[int] = None, halo: Optional[torch.Tensor] = None, n_split: int = 0) -> None:
    _require_cuda(src, dst, csr.rowptr, halo)
    _check_view3(src, "src")
    _check_view3(dst, "dst")
    Tc, _, F = src.shape
    rows = int(csr.rowptr.numel() - 1) if n_rows is None else n_rows
    assert dst.shape[0] == Tc and dst.shape[2] == F and dst.shape[1] >= rows
    if halo is None:
        h_ptr, h_ts, h_ns = None, 0, 0
    else:
        _check_view3(halo, "halo")
        h_ptr, h_ts, h_ns = _p(halo), halo.stride(0), halo.stride(1)
    _call(src.device, "sgp_spmm_halo", _p(csr.rowptr), _p(csr.col), _p(csr.val), _p(row_order),
                               _p(src), src.stride(0), src.stride(1), h_ptr, h_ts, h_ns, n_split,
                               _p(dst), dst.stride(0), dst.stride(1), rows, F, Tc,
                               _stream(src.device))


def khop_spmm(csr: Csr, buf: torch.Tensor, block_in: int, block_out0: int, hops: int, F: int,
              row_order: Optional[torch.Tensor] = None) -> None:
    _require_cuda(buf)
    _check_view3(buf, "buf")
    Tc, N, _ = buf.shape
    _call(buf.device, "sgp_khop_spmm", _p(csr.rowptr), _p(csr.col), _p(csr.val), _p(row_order), _p(buf),
                               buf.stride(0), buf.stride(1), block_in, block_out0, hops, N, F, Tc,
                               _stream(buf.device))


@dataclass
class Rbu:
    """Row-block-union operator (see include/sgp_b200.h)."""
    grp_ptr: torch.Tensor
    grp_rows: torch.Tensor
    ucol: torch.Tensor
    uval: torch.Tensor
    R: int
    n_groups: int
    fill: float


def to_host(*tensors: torch.Tensor):
    """Device arrays -> numpy through pinned staging buffers (torch's host allocator caches them): the
    operator build copies the CSR (80 MB at C4) to the host for the row grouping, and a pageable
    ``.cpu()`` moves it at ~2 GB/s."""
    out = []
    for t in tensors:
        if t.is_cuda:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            out.append(h)
        else:
            out.append(t)
    if any(t.is_cuda for t in tensors):
        torch.cuda.current_stream(next(t.device for t in tensors if t.is_cuda)).synchronize()
    return tuple(h.numpy() for h in out)


def group_rows_host(rowptr: np.ndarray, col: np.ndarray, val: np.ndarray, N: int, R: int) -> np.ndarray:
    """Host greedy grouping (csrc/group_rows.cu): returns grp_rows [n_groups, R] int32."""
    lib = load()
    n_groups = (N + R - 1) // R
    out = np.empty(max(n_groups * R, 1), np.int32)
    got = ctypes.c_int32(0)
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    val = np.ascontiguousarray(val, np.float32)
    check(load().sgp_group_rows(rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, N, R,
                             out.ctypes.data, ctypes.byref(got)), "sgp_group_rows")
    return out[: n_groups * R].reshape(n_groups, R)


def partition_rows_host(rowptr: np.ndarray, col: np.ndarray, N: int, parts: int) -> np.ndarray:
    """Host recursive bisection (csrc/group_rows.cu::sgp_partition_rows): owner [N] int32."""
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    col = np.ascontiguousarray(col, np.int32)
    owner = np.empty(max(N, 1), np.int32)
    check(load().sgp_partition_rows(rowptr.ctypes.data, col.ctypes.data, N, parts, owner.ctypes.data),
          "sgp_partition_rows")
    return owner[:N]


def rbu_build(csr: Csr, R: int, grp_rows_h: Optional[np.ndarray] = None,
              n_cols: Optional[int] = None) -> Rbu:
    """Group rows on the host (needs the CSR there once) unless `grp_rows_h` [n_groups, R] is
    given, then assemble the union slabs on the device with sort/unique/scatter (one-off,
    O(nnz)).  `n_cols` > num rows for a rectangular (row-sharded, local + halo columns) operator."""
    dev = csr.rowptr.device
    N = csr.num_nodes
    if grp_rows_h is None:
        grp_rows_h = group_rows_host(*to_host(csr.rowptr, csr.col, csr.val), N, R)
    n_groups = grp_rows_h.shape[0]
    grp_rows = torch.from_numpy(grp_rows_h).to(dev)
    flat = grp_rows.reshape(-1).to(torch.int64)
    valid = flat >= 0
    slot_of = torch.empty(N, dtype=torch.int64, device=dev)
    slot_of[flat[valid]] = torch.arange(flat.numel(), device=dev)[valid]      # g*R + slot
    counts = (csr.rowptr[1:] - csr.rowptr[:-1]).to(torch.int64)
    row_of_e = torch.repeat_interleave(torch.arange(N, device=dev), counts)
    gs = slot_of[row_of_e]
    g_e, s_e = gs // R, gs % R
    NC = int(n_cols) if n_cols is not None else N
    key = g_e * NC + csr.col.to(torch.int64)
    ukey, inv = torch.unique(key, return_inverse=True)
    total_u = int(ukey.numel())
    ucol = (ukey % NC).to(torch.int32)
    ugrp = ukey // NC
    grp_ptr = torch.zeros(n_groups + 1, dtype=torch.int64, device=dev)
    grp_ptr[1:] = torch.cumsum(torch.bincount(ugrp, minlength=n_groups), 0)
    uval = torch.zeros(max(total_u, 1), R, dtype=torch.float32, device=dev)
    uval.index_put_((inv, s_e), csr.val, accumulate=True)
    fill = csr.nnz / max(total_u * R, 1)
    return Rbu(grp_ptr.to(torch.int32), grp_rows.contiguous(), ucol, uval, R, n_groups, fill)


def spmm_rbu(rbu: Rbu, src: torch.Tensor, dst: torch.Tensor, halo: Optional[torch.Tensor] = None,
             n_split: int = 0) -> None:
    """dst = S @ [src ; halo]: column ids < n_split index `src`, the rest index `halo` rows."""
    _require_cuda(src, dst, halo)
    _check_view3(src, "src")
    _check_view3(dst, "dst")
    Tc, _, F = src.shape
    if halo is None:
        h_ptr, h_ts, h_ns = None, 0, 0
    else:
        _check_view3(halo, "halo")
        h_ptr, h_ts, h_ns = _p(halo), halo.stride(0), halo.stride(1)
    _call(src.device, "sgp_spmm_rbu_halo", _p(rbu.grp_ptr), _p(rbu.grp_rows), _p(rbu.ucol), _p(rbu.uval), rbu.R,
                                   rbu.n_groups, _p(src), src.stride(0), src.stride(1), h_ptr, h_ts, h_ns,
                                   n_split, _p(dst), dst.stride(0), dst.stride(1), F, Tc,
                                   _stream(src.device))


@dataclass
class TcOp:
    """Tensor-core operator format (see sgp_spmm_rbu_tc in include/sgp_b200.h)."""
    chunk_ptr: torch.Tensor      # [n_groups+1] int32
    grp_rows: torch.Tensor       # [n_groups, 64] int32
    cols: torch.Tensor           # [total_chunks*32] int32
    bimg: torch.Tensor           # [total_chunks, 2048] float32 (pre-swizzled slab images)
    n_groups: int
    fill: float
    err: torch.Tensor            # device int32 flag


TC_R, TC_KC = 64, 32


def tc_build(csr: Csr, grp_rows_h: Optional[np.ndarray] = None, n_cols: Optional[int] = None) -> TcOp:
    """Build the tcgen05 operator format from a CSR: 64-row locality groups, per group the sorted
    union of columns padded to chunks of 32, and per chunk the [64 x 32] slab of values split into
    an fp32 image laid out exactly as the kernel's K-major SWIZZLE_128B shared-memory tile (the kernel splits it
    into tf32 hi / lo)."""
    dev = csr.rowptr.device
    N, R, KC = csr.num_nodes, TC_R, TC_KC
    if grp_rows_h is None:
        grp_rows_h = group_rows_host(*to_host(csr.rowptr, csr.col, csr.val), N, R)
    n_groups = grp_rows_h.shape[0]
    grp_rows = torch.from_numpy(np.ascontiguousarray(grp_rows_h)).to(dev)
    flat = grp_rows.reshape(-1).to(torch.int64)
    valid = flat >= 0
    slot_of = torch.empty(max(N, 1), dtype=torch.int64, device=dev)
    slot_of[flat[valid]] = torch.arange(flat.numel(), device=dev)[valid]
    counts = (csr.rowptr[1:] - csr.rowptr[:-1]).to(torch.int64)
    row_of_e = torch.repeat_interleave(torch.arange(N, device=dev), counts)
    gs = slot_of[row_of_e]
    g_e, s_e = gs // R, gs % R
    NC = int(n_cols) if n_cols is not None else N
    ukey, inv = torch.unique(g_e * NC + csr.col.to(torch.int64), return_inverse=True)
    ugrp, ucol = ukey // NC, (ukey % NC).to(torch.int32)
    cnt = torch.bincount(ugrp, minlength=n_groups)
    chunks = (cnt + KC - 1) // KC
    chunk_ptr = torch.zeros(n_groups + 1, dtype=torch.int64, device=dev)
    chunk_ptr[1:] = torch.cumsum(chunks, 0)
    total_chunks = int(chunk_ptr[-1])
    first_u = torch.zeros(n_groups + 1, dtype=torch.int64, device=dev)
    first_u[1:] = torch.cumsum(cnt, 0)
    # position of every union entry inside the padded column array
    pos_u = chunk_ptr[ugrp] * KC + (torch.arange(ukey.numel(), device=dev) - first_u[ugrp])
    cols = torch.zeros(max(total_chunks * KC, 1), dtype=torch.int32, device=dev)
    if ukey.numel():
        # padding slots repeat the group's first column (their slab values are zero)
        fill_col = ucol[first_u[:-1].clamp(max=ukey.numel() - 1)]
        chunk_grp = torch.repeat_interleave(torch.arange(n_groups, device=dev), chunks)
        cols[: total_chunks * KC] = fill_col[chunk_grp].repeat_interleave(KC)
        cols[pos_u] = ucol
    # slab images: entry (slot s, padded position p) -> chunk p // 32, k = p % 32
    p_e = pos_u[inv]
    chunk_e, k_e = p_e // KC, p_e % KC
    off = (s_e >> 3) * 256 + (s_e & 7) * 32 + (((k_e >> 2) ^ (s_e & 7)) << 2) + (k_e & 3)   # in floats
    img = torch.zeros(max(total_chunks, 1) * R * KC, dtype=torch.float32, device=dev)
    img.index_put_((chunk_e * (R * KC) + off,), csr.val, accumulate=True)
    bimg = img.view(-1, R * KC)          # fp32; the kernel splits it into tf32 hi / lo images
    fill = csr.nnz / max(int(cnt.sum()) * R, 1)
    return TcOp(chunk_ptr.to(torch.int32), grp_rows.contiguous(), cols, bimg, n_groups, fill,
                torch.zeros(1, dtype=torch.int32, device=dev))


def spmm_tc(tc: TcOp, src: torch.Tensor, dst: torch.Tensor, halo: Optional[torch.Tensor] = None,
            n_split: int = 0, checksum: Optional[torch.Tensor] = None) -> None:
    _require_cuda(src, dst, halo)
    _check_view3(src, "src")
    _check_view3(dst, "dst")
    Tc, _, F = src.shape
    if halo is None:
        h_ptr, h_ts, h_ns = None, 0, 0
    else:
        _check_view3(halo, "halo")
        h_ptr, h_ts, h_ns = _p(halo), halo.stride(0), halo.stride(1)
    _call(src.device, "sgp_spmm_rbu_tc", _p(tc.chunk_ptr), _p(tc.grp_rows), _p(tc.cols), _p(tc.bimg), tc.n_groups,
                                 _p(src), src.stride(0), src.stride(1), h_ptr, h_ts, h_ns, n_split,
                                 _p(dst), dst.stride(0), dst.stride(1), F, Tc, _p(tc.err), _p(checksum),
                                 _stream(src.device))


@dataclass
class Tc16Op:
    """fp16x3 tensor-core operator format, 96-row groups (see sgp_spmm_rbu_tc16 in include/sgp_b200.h)."""
    chunk_ptr: torch.Tensor      # [n_groups+1] int32
    grp_rows: torch.Tensor       # [n_groups, 96] int32
    cols: torch.Tensor           # [total_chunks*32] int32
    bimg: torch.Tensor           # [total_chunks, 96*64] float16 (hi | lo slab images, pre-swizzled)
    n_groups: int
    fill: float
    w_scale: float               # power of two applied to the operator values
    inf_norm: float              # max row sum of |values|: bound(S x) <= inf_norm * bound(x)
    err: torch.Tensor            # device int32 flag


TC16_R = 96


def _tc16_perm(device) -> torch.Tensor:
    """Position (in halves) of element (row r, kk) of a chunk's [96 x 64] hi | lo matrix inside its K-major
    SWIZZLE_128B image: 8-row atoms of 1 KB, 16-byte units XOR (row & 7)."""
    r = torch.arange(TC16_R, device=device)[:, None]
    kk = torch.arange(64, device=device)[None, :]
    return ((r >> 3) * 512 + (r & 7) * 64 + (((kk >> 3) ^ (r & 7)) << 3) + (kk & 7)).reshape(-1)


def tc16_build(csr: Csr, grp_rows_h: Optional[np.ndarray] = None, n_cols: Optional[int] = None) -> Tc16Op:
    """The fp16x3 operator format from a CSR: 96-row locality groups, per group the sorted union of
    columns padded to chunks of 32, per chunk the [96 x 32] slab of (scaled) values split into fp16
    hi | lo and laid out as the kernel's shared-memory image."""
    dev = csr.rowptr.device
    N, R, KC = csr.num_nodes, TC16_R, TC_KC
    if grp_rows_h is None:
        grp_rows_h = group_rows_host(*to_host(csr.rowptr, csr.col, csr.val), N, R)
    n_groups = grp_rows_h.shape[0]
    grp_rows = torch.from_numpy(np.ascontiguousarray(grp_rows_h)).to(dev)
    flat = grp_rows.reshape(-1).to(torch.int64)
    valid = flat >= 0
    slot_of = torch.empty(max(N, 1), dtype=torch.int64, device=dev)
    slot_of[flat[valid]] = torch.arange(flat.numel(), device=dev)[valid]
    counts = (csr.rowptr[1:] - csr.rowptr[:-1]).to(torch.int64)
    row_of_e = torch.repeat_interleave(torch.arange(N, device=dev), counts)
    gs = slot_of[row_of_e]
    g_e, s_e = gs // R, gs % R
    NC = int(n_cols) if n_cols is not None else N
    ukey, inv = torch.unique(g_e * NC + csr.col.to(torch.int64), return_inverse=True)
    ugrp, ucol = ukey // NC, (ukey % NC).to(torch.int32)
    cnt = torch.bincount(ugrp, minlength=n_groups)
    chunks = (cnt + KC - 1) // KC
    chunk_ptr = torch.zeros(n_groups + 1, dtype=torch.int64, device=dev)
    chunk_ptr[1:] = torch.cumsum(chunks, 0)
    total_chunks = int(chunk_ptr[-1])
    first_u = torch.zeros(n_groups + 1, dtype=torch.int64, device=dev)
    first_u[1:] = torch.cumsum(cnt, 0)
    pos_u = chunk_ptr[ugrp] * KC + (torch.arange(ukey.numel(), device=dev) - first_u[ugrp])
    cols = torch.zeros(max(total_chunks * KC, 1), dtype=torch.int32, device=dev)
    if ukey.numel():
        fill_col = ucol[first_u[:-1].clamp(max=ukey.numel() - 1)]
        chunk_grp = torch.repeat_interleave(torch.arange(n_groups, device=dev), chunks)
        cols[: total_chunks * KC] = fill_col[chunk_grp].repeat_interleave(KC)
        cols[pos_u] = ucol
    p_e = pos_u[inv]
    chunk_e, k_e = p_e // KC, p_e % KC
    C = max(total_chunks, 1)
    dense = torch.zeros(C * R * KC, dtype=torch.float32, device=dev)          # [chunk][row][k], duplicates summed
    dense.index_put_((chunk_e * (R * KC) + s_e * KC + k_e,), csr.val, accumulate=True)
    wmax = float(dense.abs().max()) if csr.nnz else 0.0
    w_scale = float(2.0 ** np.floor(np.log2(16384.0 / wmax))) if wmax > 0 else 1.0
    sc = dense.view(C, R, KC) * w_scale
    hi = sc.half()
    lo = (sc - hi.float()).half()
    del dense, sc
    bimg = torch.empty(C, R * 64, dtype=torch.float16, device=dev)
    bimg[:, _tc16_perm(dev)] = torch.cat([hi, lo], dim=2).view(C, R * 64)
    row_abs = torch.zeros(max(N, 1), device=dev).index_add_(0, row_of_e, csr.val.abs())
    inf_norm = float(row_abs.max()) if csr.nnz else 0.0
    fill = csr.nnz / max(int(cnt.sum()) * R, 1)
    return Tc16Op(chunk_ptr.to(torch.int32), grp_rows.contiguous(), cols, bimg, n_groups, fill, w_scale, inf_norm,
                  torch.zeros(1, dtype=torch.int32, device=dev))


def spmm_tc16(tc: Tc16Op, src: torch.Tensor, dst: torch.Tensor, bound: float, halo: Optional[torch.Tensor] = None,
              n_split: int = 0, checksum: Optional[torch.Tensor] = None) -> None:
    """fp16x3 tensor-core hop; `bound` >= max|src| (and |halo|): chooses the power-of-two panel scale."""
    _require_cuda(src, dst, halo)
    _check_view3(src, "src")
    _check_view3(dst, "dst")
    Tc, _, F = src.shape
    if not bound > 0:
        raise ValueError("spmm_tc16 needs a positive bound on |src|")
    x_scale = float(2.0 ** np.floor(np.log2(16384.0 / (float(bound) * (1 + 1e-6)))))
    if halo is None:
        h_ptr, h_ts, h_ns = None, 0, 0
    else:
        _check_view3(halo, "halo")
        h_ptr, h_ts, h_ns = _p(halo), halo.stride(0), halo.stride(1)
    _call(src.device, "sgp_spmm_rbu_tc16", _p(tc.chunk_ptr), _p(tc.grp_rows), _p(tc.cols), _p(tc.bimg), tc.n_groups,
          _p(src), src.stride(0), src.stride(1), h_ptr, h_ts, h_ns, n_split, _p(dst), dst.stride(0), dst.stride(1),
          F, Tc, x_scale, float(tc.w_scale), _p(tc.err), _p(checksum), _stream(src.device))


def tc_set_cta_limit(n_ctas: int) -> None:
    """Persistent CTAs per tensor-core hop launch (process-wide; 148 = one per SM)."""
    check(load().sgp_tc_set_cta_limit(int(n_ctas)), "sgp_tc_set_cta_limit")


def tc_check(tc) -> None:
    """Raise if a tensor-core launch reported an internal barrier timeout (synchronises)."""
    if int(tc.err.item()) != 0:
        raise _lib.SgpError("sgp_spmm_rbu_tc: internal barrier timed out (results invalid)")


def node_sum(src: torch.Tensor, sums: torch.Tensor) -> None:
    _check_view3(src, "src")
    Tc, N, F = src.shape
    assert sums.shape == (Tc, F) and sums.is_contiguous()
    _call(src.device, "sgp_node_sum", _p(src), src.stride(0), src.stride(1), _p(sums), N, F, Tc,
                              _stream(src.device))


def node_mean_broadcast(sums: torch.Tensor, n_total: int, dst: torch.Tensor) -> None:
    _check_view3(dst, "dst")
    Tc, N, F = dst.shape
    _call(dst.device, "sgp_node_mean_broadcast", _p(sums), n_total, _p(dst), dst.stride(0), dst.stride(1),
                                         N, F, Tc, _stream(dst.device))


def checksum(buf: torch.Tensor, acc: torch.Tensor) -> None:
    assert buf.is_contiguous() and buf.dtype == torch.float32 and acc.dtype == torch.float64
    _call(buf.device, "sgp_checksum", _p(buf), buf.numel(), _p(acc), _stream(buf.device))


def checksum_view(view: torch.Tensor, acc: torch.Tensor) -> None:
    """acc += sum(view) for a [T, N, F] float32 view with contiguous features (fp64 accumulation)."""
    _check_view3(view, "view")
    assert acc.dtype == torch.float64
    Tc, N, F = view.shape
    _call(view.device, "sgp_checksum_view", _p(view), view.stride(0), view.stride(1), N, F, Tc, _p(acc),
          _stream(view.device))


def gesn_update(x: torch.Tensor, w_ih: torch.Tensor, bias: Optional[torch.Tensor], prop: torch.Tensor,
                alpha: float, activation: str, h_state: torch.Tensor, out: torch.Tensor) -> None:
    """One DynGESN layer step: h' = (1-a) h + a act(x W_ih^T + b + prop); x [N, Fin], prop [N, H],
    h_state [N, H] contiguous in/out, out [N, H] view (row stride free)."""
    _require_cuda(x, w_ih, bias, prop, h_state, out)
    N, Fin = x.shape
    H = int(w_ih.shape[0])
    assert x.stride(1) == 1 and prop.stride(1) == 1 and out.stride(1) == 1 and h_state.is_contiguous()
    assert w_ih.is_contiguous() and prop.shape == (N, H) and out.shape == (N, H) and h_state.shape == (N, H)
    a = float(alpha)
    _call(x.device, "sgp_gesn_update", _p(x), x.stride(0), Fin, _p(w_ih), _p(bias), _p(prop), prop.stride(0),
          a, float(1.0 - a), ACT_CODES[activation], _p(h_state), _p(out), out.stride(0), N, H, _stream(x.device))


def grouped_linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], groups: int,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[r, g*Cout + o] = b + sum_i x[r, g*Cin + i] w[g*Cout + o, i]; x [rows, groups*Cin] (row stride
    free), weight [groups*Cout, Cin], returns [rows, groups*Cout]."""
    _require_cuda(x, weight, bias)
    rows, Din = x.shape
    Cin = Din // groups
    Cout = int(weight.shape[0]) // groups
    assert Din == groups * Cin and weight.shape == (groups * Cout, Cin) and weight.is_contiguous() and x.stride(1) == 1
    if out is None:
        out = torch.empty(rows, groups * Cout, device=x.device)
    _call(x.device, "sgp_grouped_linear", _p(x), x.stride(0), _p(weight), _p(bias), _p(out), out.stride(0), rows,
          groups, Cin, Cout, _stream(x.device))
    return out


def push_rows(src: torch.Tensor, index: torch.Tensor, dst_addr: torch.Tensor, dst_t_stride: int) -> None:
    """Rows index[k] of src [Tc, N, F] -> address dst_addr[k] + t*dst_t_stride (floats) for every t:
    the halo push into peer-mapped buffers (see sgp_push_rows)."""
    _check_view3(src, "src")
    Tc, _, F = src.shape
    assert index.dtype == torch.int32 and dst_addr.dtype == torch.int64 and index.numel() == dst_addr.numel()
    _call(src.device, "sgp_push_rows", _p(src), src.stride(0), src.stride(1), _p(index), _p(dst_addr),
          int(index.numel()), int(dst_t_stride), F, Tc, _stream(src.device))


def gather_tn(src: torch.Tensor, t_idx: torch.Tensor, n_idx: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[m, :] = src[t_idx[m], n_idx[m], :]; src a [T, N, F] view, indices device int64 [M]."""
    _require_cuda(src, t_idx, n_idx, dst)
    _check_view3(src, "src")
    T, N, F = src.shape
    M = int(t_idx.numel())
    assert t_idx.dtype == torch.int64 and n_idx.dtype == torch.int64 and n_idx.numel() == M
    assert t_idx.is_contiguous() and n_idx.is_contiguous()
    assert dst.dim() == 2 and dst.shape == (M, F) and dst.dtype == torch.float32 and dst.stride(1) == 1
    _call(src.device, "sgp_gather_tn", _p(src), src.stride(0), src.stride(1), T, N, F, _p(t_idx), _p(n_idx), M,
          _p(dst), dst.stride(0), _stream(src.device))


def gather_rows(src: torch.Tensor, index: torch.Tensor, dst: torch.Tensor) -> None:
    _check_view3(src, "src")
    _check_view3(dst, "dst")
    Tc, _, F = src.shape
    _call(src.device, "sgp_gather_rows", _p(src), src.stride(0), src.stride(1), _p(index), int(index.numel()),
                                 _p(dst), dst.stride(0), dst.stride(1), F, Tc, _stream(src.device))
