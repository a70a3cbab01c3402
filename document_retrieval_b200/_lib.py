"""ctypes binding of include/br_b200.h (the C-ABI shared library libbr_b200.so, built in-tree by
build.py).  There is no CPU fallback: importing works anywhere (so that host-side logic is testable),
but every compute entry point needs the CUDA library and a B200, and fails loudly otherwise."""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (initialises the CUDA primary context the library shares)

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("BR_B200_SO") or os.path.join(_HERE, "libbr_b200.so")   # BR_B200_SO: build-variant experiments

VARIANT_ID = {"notebook": 0, "okapi": 1, "okapi_no_plus1": 2}
BR_MAX_K = 1024

# name -> (restype, argtypes); mirrors include/br_b200.h one to one
_P = C.c_void_p
SIGNATURES = {
    "br_last_error": (C.c_char_p, []),
    "br_version": (C.c_char_p, []),
    "br_index_build": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int64, _P, C.POINTER(_P)]),
    "br_index_finalize": (C.c_int, [_P, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, _P, _P]),
    "br_index_destroy": (None, [_P]),
    "br_index_stats": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                 C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "br_index_stats_in_force": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "br_index_df_dev": (_P, [_P]),
    "br_index_export_df_idf": (C.c_int, [_P, _P, _P]),
    "br_index_export_csr": (C.c_int, [_P, _P, _P, _P, _P]),
    "br_index_import_csr": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64, _P, C.POINTER(_P)]),
    "br_index_export_csr_dev": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "br_index_import_csr_dev": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _P, C.POINTER(_P)]),
    "br_score_batch": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int, _P, _P]),
    "br_topk_batch": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int, _P, _P, _P, _P]),
    "br_topk_batch_records": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int, _P, _P, _P]),
    "br_topk_merge_records": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "br_rescore_docs": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int, _P, _P, _P, _P]),
    "br_index_enable_tfidf": (C.c_int, [_P, _P]),
    "br_tfidf_cosine_topk": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "br_rerank_v3_scores": (C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "br_topk_merge": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "br_row_inv_norms": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P]),
    "br_cosine_topk": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P]),
    "br_set_cosine_option": (C.c_int, [C.c_char_p, C.c_int]),
    "br_cosine_rerank": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, C.c_int32, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "br_dedupe_first_docs": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "br_set_thr_exchange": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "br_tile_launch_count": (C.c_int, [_P, C.c_int32]),
    "br_trim_scratch": (C.c_int, []),
    "br_last_query_stats": (C.c_int, [_P, _P]),
    "br_set_profiling": (C.c_int, [_P, C.c_int]),
    "br_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "br_tokenize_count": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, C.POINTER(C.c_int64), _P]),
    "br_vocab_build": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, C.c_int64, _P, _P, C.POINTER(_P)]),
    "br_vocab_lookup": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, _P, C.c_int64, _P, _P]),
    "br_vocab_stats": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "br_vocab_export": (C.c_int, [_P, _P, _P]),
    "br_vocab_import": (C.c_int, [_P, _P, C.c_int64, _P, C.POINTER(_P)]),
    "br_vocab_destroy": (None, [_P]),
}


THR_EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)


class QueryStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("queries_fused", C.c_int64), ("queries_dense", C.c_int64),
                ("candidates_rescored", C.c_int64), ("postings_bytes", C.c_int64), ("score_launches", C.c_int64),
                ("score_ms", C.c_double)]


class BRError(RuntimeError):
    pass


class VocabularyNotDistinct(BRError):
    """Vocabulary.from_terms: two terms have the same string form (e.g. the int 1 and the str "1")."""


_lib = None


def load():
    """Load libbr_b200.so (no compute happens here)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(SO_PATH):
            raise BRError(f"{SO_PATH} is missing - run `python -m document_retrieval_b200.build` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().br_last_error().decode("utf-8", "replace")
        raise BRError(f"{what or 'br_b200'} failed ({status}): {msg}")


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise BRError("document_retrieval_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise BRError(f"device must be a CUDA device, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
