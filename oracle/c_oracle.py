"""TEST INFRASTRUCTURE - ctypes wrapper over oracle/_ref/libbm25_oracle.so (plain-C restatement,
oracle/bm25_oracle.c).  Used by tests/ as a fast checker and by bench.py as the CPU baseline."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libbm25_oracle.so")
VARIANT_ID = {"notebook": 0, "okapi": 1, "okapi_no_plus1": 2}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bm25_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_build.restype = C.c_void_p
        L.orc_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double,
                                C.c_int, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_set_stats.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
        L.orc_nnz.restype = C.c_int64
        L.orc_nnz.argtypes = [C.c_void_p]
        L.orc_avgdl.restype = C.c_double
        L.orc_avgdl.argtypes = [C.c_void_p]
        L.orc_export.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orc_get_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int, C.c_void_p]
        L.orc_topk_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class COracle:
    def __init__(self, doc_offsets, token_ids, vocab, k1=1.5, b=0.75, variant="notebook", n_threads=None):
        self.doc_offsets = np.ascontiguousarray(doc_offsets, dtype=np.int64)
        self.token_ids = np.ascontiguousarray(token_ids, dtype=np.int32)
        self.n_docs = self.doc_offsets.size - 1
        self.vocab = int(vocab)
        self.n_threads = n_threads or os.cpu_count() or 1
        self._h = lib().orc_build(_p(self.token_ids), _p(self.doc_offsets), self.n_docs, self.vocab,
                                  float(k1), float(b), VARIANT_ID[variant], self.n_threads)
        if not self._h:
            raise ValueError("orc_build failed (empty corpus or token id outside [0, vocab))")
        self.nnz = int(lib().orc_nnz(self._h))
        self.avgdl = float(lib().orc_avgdl(self._h))

    def set_stats(self, n_stat, avgdl, df=None):
        df = None if df is None else np.ascontiguousarray(df, dtype=np.int64)
        lib().orc_set_stats(self._h, float(n_stat), float(avgdl), _p(df))

    def export(self):
        row_ptr = np.empty(self.vocab + 1, np.int64)
        doc = np.empty(self.nnz, np.int32)
        tf = np.empty(self.nnz, np.int32)
        dl = np.empty(self.n_docs, np.int32)
        df = np.empty(self.vocab, np.int64)
        idf = np.empty(self.vocab, np.float64)
        lib().orc_export(self._h, _p(row_ptr), _p(doc), _p(tf), _p(dl), _p(df), _p(idf))
        return dict(row_ptr=row_ptr, doc=doc, tf=tf, dl=dl, df=df, idf=idf)

    def get_scores(self, q_terms, dedup=True):
        q = np.ascontiguousarray(q_terms, dtype=np.int32)
        out = np.empty(self.n_docs, np.float64)
        lib().orc_get_scores(self._h, _p(q), q.size, int(dedup), _p(out))
        return out

    def topk_batch(self, q_terms, q_offsets, k, dedup=True, positive_only=False, n_threads=None):
        q = np.ascontiguousarray(q_terms, dtype=np.int32)
        o = np.ascontiguousarray(q_offsets, dtype=np.int32)
        nq = o.size - 1
        ids = np.empty((nq, k), np.int32)
        sc = np.empty((nq, k), np.float64)
        cnt = np.empty(nq, np.int32)
        lib().orc_topk_batch(self._h, _p(q), _p(o), nq, k, int(dedup), int(positive_only),
                             n_threads or self.n_threads, _p(ids), _p(sc), _p(cnt))
        return ids, sc, cnt

    def __del__(self):
        try:
            if self._h:
                lib().orc_free(self._h)
                self._h = None
        except Exception:
            pass
