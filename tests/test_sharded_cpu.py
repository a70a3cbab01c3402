"""CPU (gloo, world_size 2): the host-side logic of the doc-sharded path - shard bounds, the
all-reduce of df / N / sum(dl), global-id mapping, the all-gather of [Q,k] candidates and the merge
order.  The per-shard scorer is a stand-in backed by the numpy oracle (test infrastructure only); the
product's scorer and merge are CUDA and are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bm25_oracle as orc
from document_retrieval_b200 import synth
from document_retrieval_b200.sharded import ShardedBM25, gather_candidates, reduce_stats, shard_bounds


def test_shard_bounds_cover_everything():
    for n, w in [(10, 1), (10, 3), (8_800_000, 8), (7, 8)]:
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def _merge_cpu(all_ids, all_sc, k):
    g, q, kk = all_ids.shape
    ids = all_ids.permute(1, 0, 2).reshape(q, g * kk).numpy()
    sc = all_sc.permute(1, 0, 2).reshape(q, g * kk).numpy()
    out_i = np.full((q, k), -1, np.int64)
    out_s = np.zeros((q, k))
    for i in range(q):
        ok = ids[i] >= 0
        o = np.lexsort((ids[i][ok], -sc[i][ok]))[:k]
        out_i[i, :o.size], out_s[i, :o.size] = ids[i][ok][o], sc[i][ok][o]
    return torch.from_numpy(out_i), torch.from_numpy(out_s)


class _OracleShard:
    """Stand-in for BM25 on one shard: same attributes ShardedBM25 touches."""

    def __init__(self, c, lo, hi, n_stat, sum_dl, df):
        do = c["doc_offsets"][lo:hi + 1] - c["doc_offsets"][lo]
        tk = c["token_ids"][c["doc_offsets"][lo]:c["doc_offsets"][hi]]
        self.ix = orc.build_index(do, tk, c["vocab"])
        self.doc_base, self.n_stat, self.avgdl, self.df = lo, n_stat, sum_dl / n_stat, df

    def retrieve_top_n_batch(self, queries, n):
        q_terms, q_off = queries
        ids = np.full((q_off.size - 1, n), -1, np.int32)
        sc = np.zeros((q_off.size - 1, n))
        for i in range(q_off.size - 1):
            s = orc.get_scores(self.ix, q_terms[q_off[i]:q_off[i + 1]], n_docs=self.n_stat, avgdl=self.avgdl, df=self.df)
            a, b = orc.topk_canonical(s, n)
            ids[i, :a.size], sc[i, :a.size] = a, b
        return torch.from_numpy(ids), torch.from_numpy(sc)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = synth.make_config("C1", scale=0.05)
        lo, hi = shard_bounds(c["n_docs"], world)[rank]
        do = c["doc_offsets"][lo:hi + 1] - c["doc_offsets"][lo]
        tk = c["token_ids"][c["doc_offsets"][lo]:c["doc_offsets"][hi]]
        local = orc.build_index(do, tk, c["vocab"])
        df, n_stat, sum_dl = reduce_stats(torch.from_numpy(local.df), hi - lo, int(local.dl.sum()))
        shard = _OracleShard(c, lo, hi, n_stat, sum_dl, df)
        sh = ShardedBM25(shard, merge=_merge_cpu)
        ids, sc = sh.retrieve_top_n_batch((c["q_terms"], c["q_offsets"]), 10)
        ret[rank] = (ids.numpy(), sc.numpy(), n_stat, sum_dl, df)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_equals_single_index():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    c = synth.make_config("C1", scale=0.05)
    ix = orc.build_index(c["doc_offsets"], c["token_ids"], c["vocab"])
    ids0, sc0, n_stat, sum_dl, df = ret[0]
    ids1, sc1, *_ = ret[1]
    assert n_stat == ix.n_docs and sum_dl == int(ix.dl.sum()) and np.array_equal(df, ix.df)
    assert np.array_equal(ids0, ids1) and np.array_equal(sc0, sc1)          # identical on every rank
    for i in range(c["q_offsets"].size - 1):
        a, b = orc.retrieve_top_n(ix, c["q_terms"][c["q_offsets"][i]:c["q_offsets"][i + 1]], 10)
        assert np.array_equal(ids0[i], a) and np.array_equal(sc0[i], b)     # bit-identical to one index


class _OracleCosShard:
    """Stand-in for CosineIndex on one row shard (oracle-backed, CPU)."""

    def __init__(self, rows, base):
        self.rows, self.base = rows, base

    def topk(self, q, k):
        ids, sims = orc.cosine_topk(self.rows, np.asarray(q, np.float32), k)
        return torch.from_numpy(ids + self.base), torch.from_numpy(sims)


def _cos_worker(rank, world, port, ret):
    from document_retrieval_b200.sharded import ShardedCosineIndex
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(4)
        docs = rng.standard_normal((3001, 32)).astype(np.float32)
        docs[2500] = docs[17]                         # exact tie across the two shards
        qs = np.concatenate([rng.standard_normal((40, 32)).astype(np.float32), docs[17:18]])
        lo, hi = shard_bounds(docs.shape[0], world)[rank]
        sh = ShardedCosineIndex(None, lo, local=_OracleCosShard(docs[lo:hi], lo), merge=_merge_cpu)
        ids, sims = sh.topk(qs, 10)
        ret[rank] = (ids.numpy(), sims.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_cosine_equals_single_index():
    """Row-sharded cosine top-k (config 5 layout): global row ids, all-gather, merge by (cosine desc, row asc)."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_cos_worker, args=(world, port, ret), nprocs=world, join=True)
    rng = np.random.default_rng(4)
    docs = rng.standard_normal((3001, 32)).astype(np.float32)
    docs[2500] = docs[17]
    qs = np.concatenate([rng.standard_normal((40, 32)).astype(np.float32), docs[17:18]])
    ids, sims = orc.cosine_topk(docs, qs, 10)
    assert np.array_equal(ret[0][0], ret[1][0]) and np.array_equal(ret[0][1], ret[1][1])
    assert np.array_equal(ret[0][0], ids) and np.array_equal(ret[0][1], sims.astype(np.float64))
    assert ret[0][0][-1, :2].tolist() == [17, 2500]


def test_abi_symbols_exported():
    """The C-ABI library loads and exports every symbol include/br_b200.h declares (no compute)."""
    import re
    from document_retrieval_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "br_b200.h")).read()
    declared = set(re.findall(r"\b(br_[a-z0-9_]+)\s*\(", hdr))
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.br_version()


def test_no_cuda_fails_loudly():
    from document_retrieval_b200 import BM25, _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BRError):
        BM25([["a", "b"], ["b"]])
