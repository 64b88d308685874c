"""ctypes binding of include/sgp_b200.h (libsgp_b200.so, built in-tree by sgp_b200/_build.py).

There is no fallback: if the shared library is missing and cannot be built, or a call returns a
non-zero code, this raises.  PyTorch only provides device memory and streams around these calls.
"""
from __future__ import annotations

import ctypes
import os
import re
import warnings
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

from . import _build

OK = 0
ACT_CODES = {"tanh": 0, "relu": 1, "self_norm": 2, "identity": 3}
CSR_SET_DIAG, CSR_REMOVE_DIAG, CSR_GCN_NORM, CSR_SYMMETRIZE, CSR_TRANSPOSE, CSR_NO_NORM = 1, 2, 4, 8, 16, 32

# name -> (restype, argtypes); must list every symbol include/sgp_b200.h declares
SIGNATURES = {
    "sgp_version": (c_int, []),
    "sgp_last_error": (c_char_p, []),
    "sgp_launch_count": (c_int64, []),
    "sgp_csr_build_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int]),
    "sgp_csr_build": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int, c_void_p, c_void_p,
                              c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sgp_reservoir_pack_rows": (c_int, [c_int, c_int]),
    "sgp_reservoir_pack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sgp_reservoir_scan": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_float,
                                   c_float, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                   c_int, c_void_p]),
    "sgp_reservoir_scan_multi": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int,
                                         c_void_p]),
    "sgp_reservoir_tc_pack": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "sgp_reservoir_scan_tc": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_float,
                                      c_float, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                      c_int, c_void_p, c_void_p, c_void_p]),
    "sgp_reservoir_tc16_pack": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "sgp_reservoir_scan_tc16": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_float, c_void_p, c_void_p,
                                        c_float, c_float, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p]),
    "sgp_spmm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                         c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    "sgp_spmm_halo": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                              c_int64, c_int64, c_int, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                              c_void_p]),
    "sgp_khop_spmm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                              c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "sgp_spmm_rbu": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64,
                             c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    "sgp_spmm_rbu_halo": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64,
                                  c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int64,
                                  c_int, c_int, c_void_p]),
    "sgp_spmm_rbu_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64,
                                c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int64, c_int, c_int,
                                c_void_p, c_void_p, c_void_p]),
    "sgp_spmm_rbu_tc16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64,
                                  c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int64, c_int, c_int,
                                  c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "sgp_tc_set_cta_limit": (c_int, [c_int]),
    "sgp_group_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "sgp_gesn_update": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float,
                                c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "sgp_grouped_linear": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                   c_int, c_void_p]),
    "sgp_partition_rows": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "sgp_node_sum": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_void_p]),
    "sgp_node_mean_broadcast": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int,
                                        c_int, c_void_p]),
    "sgp_checksum": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "sgp_checksum_view": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sgp_push_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p]),
    "sgp_gather_tn": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_int64,
                              c_void_p, c_int64, c_void_p]),
    "sgp_gather_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_int64, c_int64,
                                c_int, c_int, c_void_p]),
}


class SgpError(RuntimeError):
    pass


_lib = None


def header_abi_version() -> int:
    """SGP_B200_ABI_VERSION of include/sgp_b200.h (the header this binding was written against)."""
    hdr = os.path.join(os.path.dirname(_build.HERE), "include", "sgp_b200.h")
    with open(hdr) as f:
        m = re.search(r"#define\s+SGP_B200_ABI_VERSION\s+(\d+)", f.read())
    if not m:
        raise SgpError("include/sgp_b200.h does not define SGP_B200_ABI_VERSION")
    return int(m.group(1))


def load() -> ctypes.CDLL:
    """Load (building first if stale and nvcc is present) the C-ABI library; raise if impossible.
    A library that could not be rebuilt is accepted only if it is not older than its sources'
    ABI: sgp_version() must equal the header's SGP_B200_ABI_VERSION."""
    global _lib
    if _lib is not None:
        return _lib
    so = _build.SO
    try:
        so = os.environ.get("SGP_B200_SO") or _build.build()
    except Exception as e:  # noqa: BLE001 - no nvcc on the box: use the prebuilt file if present
        if not os.path.exists(so):
            raise SgpError(f"libsgp_b200.so is missing and could not be built: {e}") from e
        if _build.stale():
            warnings.warn(f"libsgp_b200.so is older than its sources and could not be rebuilt ({e}); "
                          "loading it only if its ABI version matches the header", RuntimeWarning)
    lib = ctypes.CDLL(so)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    have, want = int(lib.sgp_version()), header_abi_version()
    if have != want:
        raise SgpError(f"{so} implements ABI version {have}, include/sgp_b200.h declares {want}: rebuild "
                       "the library (python -m sgp_b200._build --force)")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = load().sgp_last_error()
        raise SgpError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().sgp_launch_count())
