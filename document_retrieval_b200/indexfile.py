"""Flat binary index file: what replaces ``joblib.dump(bm25_model, f'bm25_model_{lang}.joblib')``
(bm25_ranking.ipynb:312) and the sharded pickles of final_implementation.py:187-287, whose loading took the reference
longer than the retrieval itself (:222-251).

Layout: ``BRIX0001`` | uint64 header length | JSON header (UTF-8) | arrays, each at a 4096-byte aligned offset:
``row_ptr`` int64[V+1], ``doc`` int32[nnz], ``tf`` uint16[nnz], ``dl`` int32[N], optional ``df_stat`` int64[V] (a doc
shard's global df), optional vocabulary ``pool_off`` int64[V+1] + ``pool`` uint8.  Loading never unpickles anything:
every array is read into a pinned staging buffer and copied to the device chunk by chunk (read of chunk i+1 overlaps the
copy of chunk i), the CSR is validated on the device (br_index_import_csr_dev) and the weights / skip tables / rows are
rebuilt there (br_index_finalize)."""
from __future__ import annotations

import ctypes as C
import json
import os
import struct

import numpy as np
import torch

from . import _lib
from ._lib import BRError, check, ptr

MAGIC = b"BRIX0001"
ALIGN = 4096
CHUNK = 32 << 20
_DT = {"int64": np.int64, "int32": np.int32, "uint16": np.uint16, "uint8": np.uint8}
_staging = {}


def _pinned(dev):
    key = str(dev)
    if key not in _staging:
        _staging[key] = ([torch.empty(CHUNK, dtype=torch.uint8).pin_memory() for _ in range(2)],
                         [torch.cuda.Event() for _ in range(2)])
    return _staging[key]


def _to_device(f, offset, count, dtype, dev):
    """file[offset : offset + count*itemsize] -> device tensor, through two pinned chunks."""
    np_dt = np.dtype(_DT[dtype])
    nbytes = count * np_dt.itemsize
    out = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
    bufs, evs = _pinned(dev)
    f.seek(offset)
    done, i = 0, 0
    while done < nbytes:
        n = min(CHUNK, nbytes - done)
        b, e = bufs[i & 1], evs[i & 1]
        e.synchronize()                                           # the copy that last used this buffer has finished
        got = f.readinto(memoryview(b.numpy())[:n])
        if got != n:
            raise BRError("index file is truncated")
        out[done:done + n].copy_(b[:n], non_blocking=True)
        e.record(torch.cuda.current_stream(dev))
        done += n
        i += 1
    t_dt = {"int64": torch.int64, "int32": torch.int32, "uint16": torch.int16, "uint8": torch.uint8}[dtype]
    return out[:nbytes].view(t_dt)


def save(model, path):
    model._require()
    lib = _lib.load()
    dev = model._device
    st = model.stats()
    V, nnz, N = st["vocab"], st["nnz"], st["n_docs"]
    n_stat, sdl = C.c_double(), C.c_double()
    check(lib.br_index_stats_in_force(model._h, C.byref(n_stat), C.byref(sdl)), "br_index_stats_in_force")
    with torch.cuda.device(dev):
        row_ptr = torch.empty(V + 1, dtype=torch.int64, device=dev)
        doc = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        tf = torch.empty(max(nnz, 1), dtype=torch.int16, device=dev)
        dl = torch.empty(N, dtype=torch.int32, device=dev)
        check(lib.br_index_export_csr_dev(model._h, ptr(row_ptr), ptr(doc), ptr(tf), ptr(dl), _lib.stream_ptr(dev)),
              "br_index_export_csr_dev")
        arrays = [("row_ptr", "int64", row_ptr.cpu().numpy()), ("doc", "int32", doc[:nnz].cpu().numpy()),
                  ("tf", "uint16", tf[:nnz].cpu().numpy().view(np.uint16)), ("dl", "int32", dl.cpu().numpy())]
    df_stat, _ = model._export_df_idf()
    local_df = np.diff(arrays[0][2])
    shard = not np.array_equal(df_stat, local_df)
    if shard:
        arrays.append(("df_stat", "int64", np.ascontiguousarray(df_stat, np.int64)))
    has_terms = model.vocabulary is not None or model._terms is not None or model._term_pool is not None
    if has_terms:
        if model.vocabulary is not None:
            off, pool = model.vocabulary._export()
        elif model._term_pool is not None:
            off, pool = model._term_pool
        else:
            from .ingest import pack_texts
            pool, off = pack_texts([str(t) for t in model._terms])
        arrays += [("pool_off", "int64", np.ascontiguousarray(off, np.int64)), ("pool", "uint8", np.ascontiguousarray(pool, np.uint8))]
    header = dict(format=1, n_docs=N, vocab=V, nnz=nnz, doc_base=st["doc_base"], k1=model.k1, b=model.b, variant=model.variant,
                  dedup_query=bool(model.dedup_query), bigrams=bool(model.bigrams), n_stat=n_stat.value, sum_dl_stat=sdl.value,
                  shard=bool(shard), has_terms=bool(has_terms), arrays=[])
    # two passes: the header length decides the first offset
    def layout(hlen):
        pos = -(-(len(MAGIC) + 8 + hlen) // ALIGN) * ALIGN
        out = []
        for name, dt, arr in arrays:
            out.append(dict(name=name, dtype=dt, count=int(arr.size), offset=pos))
            pos = -(-(pos + arr.nbytes) // ALIGN) * ALIGN
        return out
    header["arrays"] = layout(0)
    blob = json.dumps(header).encode()
    header["arrays"] = layout(len(blob) + 64)
    room = len(blob) + 64
    blob = json.dumps(header).encode()
    assert len(blob) <= room
    blob = blob.ljust(room)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        for (name, dt, arr), meta in zip(arrays, header["arrays"]):
            f.seek(meta["offset"])
            arr.tofile(f)
    return path


def read_header(path):
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise BRError(f"{path} is not a br_b200 index file")
        (hlen,) = struct.unpack("<Q", f.read(8))
        if hlen > (1 << 24):
            raise BRError("index file header is corrupt")
        return json.loads(f.read(hlen).decode())


def load(cls, path, device=None):
    lib = _lib.load()
    dev = _lib.require_cuda(device)
    h = read_header(path)
    if h.get("format") != 1:
        raise BRError("unsupported index file format")
    meta = {a["name"]: a for a in h["arrays"]}
    size = os.path.getsize(path)
    for a in h["arrays"]:
        if a["offset"] + a["count"] * np.dtype(_DT[a["dtype"]]).itemsize > size:
            raise BRError("index file is truncated")
    V, nnz, N = int(h["vocab"]), int(h["nnz"]), int(h["n_docs"])
    want = {"row_ptr": V + 1, "doc": nnz, "tf": nnz, "dl": N}
    for k, n in want.items():
        if k not in meta or meta[k]["count"] != n:
            raise BRError(f"index file: array {k} has the wrong size")
    self = cls(None, h["k1"], h["b"], variant=h["variant"], dedup_query=h["dedup_query"], device=dev)
    self.bigrams = bool(h.get("bigrams", False))
    with torch.cuda.device(dev), open(path, "rb") as f:
        t = {k: _to_device(f, meta[k]["offset"], meta[k]["count"], meta[k]["dtype"], dev) for k in ("row_ptr", "doc", "tf", "dl")}
        hnd = C.c_void_p()
        check(lib.br_index_import_csr_dev(ptr(t["row_ptr"]), ptr(t["doc"]), ptr(t["tf"]), ptr(t["dl"]), N, V, nnz,
                                          int(h["doc_base"]), _lib.stream_ptr(dev), C.byref(hnd)), "br_index_import_csr_dev")
        self._h = hnd
        self.vocab_size, self.corpus_size, self.doc_base = V, N, int(h["doc_base"])
        del t
        if h.get("shard"):
            df_stat = np.fromfile(path, dtype=np.int64, count=meta["df_stat"]["count"], offset=meta["df_stat"]["offset"])
            self.finalize(h["n_stat"], h["sum_dl_stat"], df_stat)
        else:
            self.finalize()
        if h.get("has_terms"):
            off = np.fromfile(path, dtype=np.int64, count=meta["pool_off"]["count"], offset=meta["pool_off"]["offset"])
            pool = np.fromfile(path, dtype=np.uint8, count=meta["pool"]["count"], offset=meta["pool"]["offset"])
            if off.size != V + 1 or off[0] != 0 or np.any(np.diff(off) < 0) or off[-1] != pool.size:
                raise BRError("index file: vocabulary offsets are corrupt")
            self._term_pool = (off, pool)
    return self
