"""ctypes binding of libhcir_b200.so (C ABI: include/hcir_b200.h).

The product path has no CPU fallback: if the CUDA library is missing and cannot be built,
or the device is not a B200 (sm_100), every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

HCIR_OK, HCIR_EINVAL, HCIR_EARCH, HCIR_ECUDA, HCIR_EWORKSPACE = 0, -1, -2, -3, -4


class Plan(C.Structure):
    """hcir_plan_t"""
    _fields_ = [("nsplit", C.c_int32), ("cap", C.c_int32), ("kc", C.c_int32), ("flags", C.c_int32),
                ("sample_rows", C.c_int32), ("sample_stride", C.c_int32), ("chunk_w", C.c_int32),
                ("num_chunks", C.c_int32), ("sample_nsplit", C.c_int32), ("nlists", C.c_int32),
                ("hint_rank", C.c_int32), ("q_rows", C.c_int32), ("thr_rank", C.c_int32), ("reserved_", C.c_int32),
                ("counts_off", C.c_uint64), ("thr_out_off", C.c_uint64), ("thr0_off", C.c_uint64),
                ("thr_hi_off", C.c_uint64), ("cmax_off", C.c_uint64), ("keys_off", C.c_uint64),
                ("bytes", C.c_uint64)]

    def kernels(self) -> int:
        """CUDA kernels one hcir_simtopk call enqueues with this plan."""
        return 3 if self.sample_rows > 0 else 1


class Tail(C.Structure):
    """hcir_tail_t: what every query's CTA of hcir_select_rescore does with its finished top-k
    (label gather, vote, stores into the peer regions).  A zero-initialised Tail does nothing."""
    _fields_ = [("labels", C.c_void_p), ("n_labels", C.c_int64), ("num_classes", C.c_int32), ("T", C.c_float),
                ("classes", C.c_void_p), ("pred", C.c_void_p), ("out_lab", C.c_void_p),
                ("world", C.c_int32), ("rank", C.c_int32), ("payload", C.c_int32), ("reserved_", C.c_int32),
                ("slot_bytes", C.c_uint64), ("regions", C.c_void_p * 16), ("step", C.c_void_p),
                ("qmap", C.c_void_p), ("active", C.c_void_p), ("commit_certified_only", C.c_int32),
                ("no_signal", C.c_int32), ("out_rows", C.c_int64)]


PAYLOAD_BLOCK, PAYLOAD_PRED = 1, 2

_P = C.c_void_p
_I64 = C.c_int64
_INT = C.c_int
_F = C.c_float

# name -> (restype, argtypes); every symbol include/hcir_b200.h declares
SIGNATURES = {
    "hcir_abi_version": (_INT, []),
    "hcir_last_error": (C.c_char_p, []),
    "hcir_device_supported": (_INT, []),
    "hcir_padded_dim": (_INT, [_INT]),
    "hcir_l2norm_cast": (_INT, [_P, _I64, _INT, _I64, _P, _P, _INT, _P, _P]),
    "hcir_simtopk_plan": (_INT, [_I64, _I64, _INT, _INT, _INT, C.POINTER(Plan)]),
    "hcir_simtopk": (_INT, [_P, _I64, _P, _I64, _INT, C.POINTER(Plan), _P, _P]),
    "hcir_simtopk_gated": (_INT, [_P, _I64, _P, _I64, _INT, C.POINTER(Plan), _P, _P, _P]),
    "hcir_retry_setup": (_INT, [_P, _P, _P, _INT, _I64, _INT, _P, _P, _P, _F, _F, _INT, _P, _P, _P, _P, _P, _P, _P,
                                _P, _P, _P]),
    "hcir_simtopk_debug": (_INT, [_P, _I64, _P, _I64, _INT, C.POINTER(Plan), _P, _P, _P]),
    "hcir_select_rescore": (_INT, [_P, _P, _INT, _I64, _I64, _INT, _I64, C.POINTER(Plan), _P, _P, _F, _F,
                                   _P, _P, _P, _P, C.POINTER(Tail), _P]),
    "hcir_exact_workspace_bytes": (C.c_size_t, [_I64, _I64, _INT, _INT]),
    "hcir_exact_topk": (_INT, [_P, _P, _INT, _I64, _INT, _I64, _P, _I64, _P, _P, _P, C.c_size_t, _INT, _P]),
    "hcir_gather_labels": (_INT, [_P, _I64, _P, _I64, _I64, _P, _P]),
    "hcir_vote": (_INT, [_P, _P, _I64, _INT, _INT, _F, _P, _P, _P]),
    "hcir_vote_idx": (_INT, [_P, _P, _P, _I64, _I64, _I64, _INT, _INT, _F, _P, _P, _P, _P]),
    "hcir_vote_classes": (_INT, [_P, _P, _I64, _INT, _INT, _F, _P, _P, _P]),
    "hcir_merge_topk": (_INT, [_P, _P, _P, _INT, _I64, _INT, _P, _P, _P, _P]),
    "hcir_packed_block_bytes": (C.c_size_t, [_I64, _INT, _INT]),
    "hcir_merge_topk_packed": (_INT, [_P, _INT, _I64, _INT, _INT, C.c_size_t, _P, _P, _P, _P]),
    "hcir_peer_region_bytes": (C.c_size_t, [_INT, C.c_size_t]),
    "hcir_peer_slot_offset": (C.c_size_t, [_INT, C.c_size_t, _INT, _INT]),
    "hcir_peer_push_ctas": (_INT, [C.c_size_t]),
    "hcir_peer_alloc": (_INT, [C.c_size_t, C.POINTER(_P), _P]),
    "hcir_peer_open": (_INT, [_P, C.POINTER(_P)]),
    "hcir_peer_close": (_INT, [_P]),
    "hcir_peer_free": (_INT, [_P]),
    "hcir_peer_push": (_INT, [_P, C.c_size_t, C.POINTER(_P), _INT, _INT, C.c_size_t, _P, _P, _P]),
    "hcir_peer_wait": (_INT, [_P, _INT, _P, _I64, _P]),
    "hcir_peer_merge_vote": (_INT, [_P, _INT, _I64, _INT, _INT, C.c_size_t, _P, _I64, _P, _P, _P, _INT, _F, _P, _P,
                                    _P]),
}

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if needed and possible) the in-tree shared library."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if build_if_missing and not _build.is_current():
        try:
            _build.build_library()   # serialised across processes by a file lock (torchrun: one rank builds)
        except Exception as e:  # no nvcc on this box: use the shipped .so if there is one
            if not os.path.exists(path):
                raise RuntimeError(
                    f"libhcir_b200.so is missing and could not be built ({e}); the hcir_b200 "
                    "product path has no CPU fallback") from e
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found; run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI/header drift
        fn.restype = res
        fn.argtypes = args
    if lib.hcir_abi_version() != 5:
        raise RuntimeError("hcir_b200: ABI version mismatch")
    _lib = lib
    return lib


def last_error() -> str:
    return load().hcir_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc == HCIR_OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == HCIR_EINVAL:
        raise ValueError(msg)
    raise RuntimeError(f"[hcir rc={rc}] {msg}")
