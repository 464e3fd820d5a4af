"""ctypes binding of liborr.so — the same C ABI the C# shim P/Invokes (include/orr.h).

There is no CPU fallback: if the library cannot be built/loaded this module raises, and if
there is no sm_100 device `orr_store_create` fails with ORR_E_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

ORR_ABI_VERSION = 1
ORR_OK, ORR_E_INVALID, ORR_E_CUDA, ORR_E_OOM, ORR_E_UNSUPPORTED, ORR_E_INTERNAL = 0, -1, -2, -3, -4, -5
ORR_MAX_QUERY_TERMS = 64
ORR_MAX_QUERY_PROBES = 128
TICKS_PER_DAY = 864_000_000_000

XCHG_HANDLE_BYTES = 64
STATUS_BOUND_FAILED, STATUS_XCHG_TIMEOUT = 1, 4
PATH_FUSED, PATH_EXACT, PATH_SUBSET, PATH_BATCH, PATH_TEXT, PATH_ESCALATED = 1, 2, 3, 4, 5, 0x100


class OrrConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("dim", C.c_int32), ("term_slots", C.c_int32),
        ("capacity_rows", C.c_int64), ("row_base", C.c_uint64),
        ("w_cos", C.c_double), ("w_kw", C.c_double), ("w_rec", C.c_double), ("recency_days", C.c_double),
    ]


class OrrHit(C.Structure):
    _fields_ = [("row", C.c_uint64), ("score", C.c_double), ("created_ticks", C.c_int64)]


class OrrTiming(C.Structure):
    _fields_ = [
        ("scan_ms", C.c_float), ("finalize_ms", C.c_float), ("total_device_ms", C.c_float),
        ("wall_ms", C.c_float), ("path", C.c_int32), ("n_survivors", C.c_int32), ("rows_scanned", C.c_int64),
    ]


class OrrSynthSpec(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("dim", C.c_int32), ("gen_dim", C.c_int32), ("terms_per_chunk", C.c_int32),
        ("vocab", C.c_int32), ("now_ticks", C.c_int64), ("zero_row_ppm", C.c_int32), ("dup_row_ppm", C.c_int32),
    ]


class OrrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"liborr error {code}: {message}")
        self.code = code


# every symbol include/orr.h declares, with its signature
_SIGNATURES = {
    "orr_config_default": (None, [C.POINTER(OrrConfig)]),
    "orr_store_create": (C.c_int, [C.POINTER(OrrConfig), C.POINTER(C.c_void_p)]),
    "orr_store_destroy": (None, [C.c_void_p]),
    "orr_store_upsert_document_chunks": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "orr_store_upsert_document_chunks_text": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p,
                                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p,
                                                        C.c_void_p]),
    "orr_store_upsert_document_texts": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_char_p, C.c_void_p, C.c_void_p]),
    "orr_store_upsert_documents_texts": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_char_p, C.c_void_p, C.c_void_p]),
    "orr_store_vocab_size": (C.c_int64, [C.c_void_p]),
    "orr_search_query": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "orr_expand_query": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32)]),
    "orr_store_delete_document": (C.c_int, [C.c_void_p, C.c_uint64]),
    "orr_store_compact": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "orr_store_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "orr_store_load": (C.c_int, [C.c_void_p, C.c_char_p]),
    "orr_store_count": (C.c_int64, [C.c_void_p]),
    "orr_store_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "orr_store_rows_used": (C.c_int64, [C.c_void_p]),
    "orr_hash_term": (C.c_uint64, [C.c_char_p, C.c_int32]),
    "orr_tokenize_query": (C.c_int, [C.c_char_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "orr_tokenize_content": (C.c_int, [C.c_char_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "orr_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                             C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "orr_search_text": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_void_p, C.c_int64,
                                  C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "orr_search_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "orr_search_device_timing": (C.c_int, [C.c_void_p, C.POINTER(OrrTiming)]),
    "orr_search_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "orr_search_batch_device": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "orr_debug_batch_scores": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                         C.c_void_p, C.c_int64]),
    "orr_debug_scan_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_int64, C.c_void_p, C.c_int64]),
    "orr_merge_hits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                 C.POINTER(C.c_int32)]),
    "orr_merge_hits_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "orr_merge_hits_batch_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                              C.c_void_p, C.c_void_p]),
    "orr_xchg_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "orr_xchg_destroy": (None, [C.c_void_p]),
    "orr_xchg_get_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "orr_xchg_open_peer": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "orr_xchg_attach_peer": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "orr_xchg_allgather_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p]),
    "orr_xchg_sequence": (C.c_uint32, [C.c_void_p]),
    "orr_xchg_resync": (C.c_int, [C.c_void_p, C.c_uint32]),
    "orr_xchg_set_timeout_ms": (C.c_int, [C.c_void_p, C.c_double]),
    "orr_cluster_create": (C.c_int, [C.POINTER(OrrConfig), C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "orr_cluster_destroy": (None, [C.c_void_p]),
    "orr_cluster_size": (C.c_int32, [C.c_void_p]),
    "orr_cluster_shard": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "orr_cluster_count": (C.c_int64, [C.c_void_p]),
    "orr_cluster_upsert_document_chunks": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]),
    "orr_cluster_delete_document": (C.c_int, [C.c_void_p, C.c_uint64]),
    "orr_cluster_fill_synthetic": (C.c_int, [C.c_void_p, C.POINTER(OrrSynthSpec), C.c_uint64, C.c_int64]),
    "orr_cluster_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_int64, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "orr_cluster_search_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "orr_cluster_search_many": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "orr_last_error": (C.c_char_p, []),
    "orr_last_timing": (C.c_int, [C.POINTER(OrrTiming)]),
    "orr_synth_spec_default": (None, [C.POINTER(OrrSynthSpec), C.c_int32]),
    "orr_synth_rows_host": (C.c_int, [C.POINTER(OrrSynthSpec), C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "orr_synth_query_host": (C.c_int, [C.POINTER(OrrSynthSpec), C.c_uint64, C.c_uint64, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p]),
    "orr_synth_term_text": (C.c_int, [C.c_uint32, C.c_char_p]),
    "orr_synth_row_info": (C.c_int, [C.POINTER(OrrSynthSpec), C.c_uint64, C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]),
    "orr_synth_row_text": (C.c_int64, [C.POINTER(OrrSynthSpec), C.c_uint64, C.c_char_p, C.c_int64]),
    "orr_store_fill_synthetic": (C.c_int, [C.c_void_p, C.POINTER(OrrSynthSpec), C.c_uint64, C.c_int64]),
}

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def lib():
    """Loads (building first if the .so is absent or stale) liborr.so and types its entry points."""
    global _lib
    if _lib is None:
        if os.environ.get("ORR_NO_BUILD") and os.path.exists(_build.LIB_PATH):
            path = _build.LIB_PATH
        else:
            path = _build.build_native()
        L = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def declared_symbols() -> list[str]:
    return sorted(_SIGNATURES)


def check(rc: int) -> None:
    if rc != ORR_OK:
        msg = lib().orr_last_error()
        raise OrrError(rc, msg.decode("utf-8", "replace") if msg else "")
