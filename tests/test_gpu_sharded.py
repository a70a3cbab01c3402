"""GPU: doc-sharding on one device ("fake multi-shard"): S shards built with global
statistics, per-shard top-k, br_topk_merge fed directly - must equal the single-index result bit for
bit (ids and float64 scores)."""
import numpy as np
import pytest
import torch

from document_retrieval_b200 import synth
from document_retrieval_b200.sharded import merge_topk_cuda, reduce_stats, shard_bounds

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 8])
def test_fake_shards_equal_single_index(world):
    from document_retrieval_b200 import BM25
    c = synth.make_config("C1", scale=0.5)
    q = (c["q_terms"], c["q_offsets"])
    single = BM25.from_token_ids(c["doc_offsets"], c["token_ids"], c["vocab"])
    ids1, sc1 = single.retrieve_top_n_batch(q, 10)
    shards = []
    for lo, hi in shard_bounds(c["n_docs"], world):
        do = c["doc_offsets"][lo:hi + 1] - c["doc_offsets"][lo]
        tk = c["token_ids"][c["doc_offsets"][lo]:c["doc_offsets"][hi]]
        shards.append(BM25.from_token_ids(do, tk, c["vocab"], doc_base=lo, finalize=False))
    df = sum(s.local_df_tensor().to(torch.int64) for s in shards).cpu().numpy()
    n_stat = sum(s.stats()["n_docs"] for s in shards)
    sum_dl = sum(s.stats()["sum_dl"] for s in shards)
    assert n_stat == c["n_docs"]
    all_ids, all_sc = [], []
    for s in shards:
        s.finalize(n_stat, sum_dl, df)
        assert s.avgdl == single.avgdl
        ids, sc = s.retrieve_top_n_batch(q, 10)
        all_ids.append(torch.where(ids >= 0, ids.to(torch.int64) + s.doc_base, torch.full_like(ids, -1, dtype=torch.int64)))
        all_sc.append(sc)
    ids_m, sc_m = merge_topk_cuda(torch.stack(all_ids), torch.stack(all_sc), 10)
    assert torch.equal(ids_m, ids1.to(torch.int64)) and torch.equal(sc_m, sc1)


def test_reduce_stats_single_process():
    df, n, s = reduce_stats(torch.tensor([1, 2, 3], dtype=torch.int32, device="cuda"), 5, 17)
    assert df.tolist() == [1, 2, 3] and (n, s) == (5, 17)


@pytest.mark.parametrize("world", [2, 5])
def test_fake_cosine_shards_equal_single_index(world):
    """Row-sharded brute-force cosine (config 5 layout): per-shard br_cosine_topk with doc_base + br_topk_merge
    equals the single-index call bit for bit."""
    from document_retrieval_b200.cosine import CosineIndex
    g = torch.Generator(device="cuda").manual_seed(9)
    docs = torch.randn(20_000, 256, generator=g, device="cuda").to(torch.bfloat16)
    docs[777] = docs[12_345]                      # an exact tie across shards -> lower global row first
    qs = torch.cat([torch.randn(300, 256, generator=g, device="cuda").to(torch.bfloat16), docs[777:778]])
    ids1, s1 = CosineIndex(docs).topk(qs, 10)
    parts = [CosineIndex(docs[lo:hi], doc_base=lo).topk(qs, 10) for lo, hi in shard_bounds(docs.shape[0], world)]
    ids_m, sc_m = merge_topk_cuda(torch.stack([p[0] for p in parts]), torch.stack([p[1].double() for p in parts]), 10)
    assert torch.equal(ids_m, ids1)
    assert torch.equal(sc_m, s1.double())
    assert ids1[-1, :2].tolist() == [777, 12_345]


def test_threshold_exchange_with_a_second_shard_equals_single_index():
    """br_set_thr_exchange (all-gather of the shards' k best scores, thresholds = k-th largest of the union) on one
    device: shard A runs with a callback that supplies, as "the other shard", shard B's final top-k scores (each the score
    of a distinct doc of B, scaled down by 1e-6 so that they are lower bounds of B's fp32 scores).  A may then return
    fewer than k docs; merged with B's own result it must still equal the single index bit for bit."""
    import ctypes as C
    from document_retrieval_b200 import BM25, _lib
    from document_retrieval_b200._lib import check
    k = 10
    c = synth.make_config("C1", scale=1.0)
    q = (c["q_terms"], c["q_offsets"])
    nq = c["q_offsets"].size - 1
    single = BM25.from_token_ids(c["doc_offsets"], c["token_ids"], c["vocab"])
    ids1, sc1 = single.retrieve_top_n_batch(q, k)
    shards = []
    for lo, hi in shard_bounds(c["n_docs"], 2):
        do = c["doc_offsets"][lo:hi + 1] - c["doc_offsets"][lo]
        tk = c["token_ids"][c["doc_offsets"][lo]:c["doc_offsets"][hi]]
        shards.append(BM25.from_token_ids(do, tk, c["vocab"], doc_base=lo, finalize=False))
    df = sum(s.local_df_tensor().to(torch.int64) for s in shards).cpu().numpy()
    n_stat = sum(s.stats()["n_docs"] for s in shards)
    sum_dl = sum(s.stats()["sum_dl"] for s in shards)
    for s in shards:
        s.finalize(n_stat, sum_dl, df)
    a, b = shards
    ids_b, sc_b = b.retrieve_top_n_batch(q, k)
    other = (sc_b * (1.0 - 1e-6)).to(torch.float32)
    other = torch.where(ids_b >= 0, other, torch.zeros_like(other)).contiguous()          # [nq, k], 0 = no doc
    other_thr = torch.where(ids_b[:, k - 1] >= 0, other[:, k - 1], torch.zeros_like(other[:, 0])).contiguous()
    calls = []

    def wrap(ptr, n):
        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_Arr(), device="cuda")

    def exchange(local_ptr, gathered_ptr, n_floats, stream, user):
        try:
            local, gathered = wrap(local_ptr, n_floats), wrap(gathered_ptr, 2 * n_floats)
            gathered[:n_floats].copy_(local)
            gathered[n_floats:].copy_((other_thr if n_floats == nq else other).flatten())
            calls.append(int(n_floats))
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return -1

    lib = _lib.load()
    cb = _lib.THR_EXCHANGE_FN(exchange)
    rounds = int(lib.br_tile_launch_count(a._h, k)) - 1
    assert rounds >= 0
    check(lib.br_set_thr_exchange(a._h, cb, None, rounds, 2), "br_set_thr_exchange")
    try:
        ids_a, sc_a = a.retrieve_top_n_batch(q, k)
    finally:
        check(lib.br_set_thr_exchange(a._h, None, None, -1, 1), "br_set_thr_exchange")
    assert calls and calls[0] == nq and all(n == nq * k for n in calls[1:]) and len(calls) == 1 + rounds
    ids_a0, _ = a.retrieve_top_n_batch(q, k)                     # without the exchange: full lists
    assert int((ids_a < 0).sum()) >= int((ids_a0 < 0).sum())    # shared thresholds can only shorten a shard's list
    glob = lambda s, ids: torch.where(ids >= 0, ids.to(torch.int64) + s.doc_base, torch.full_like(ids, -1, dtype=torch.int64))
    ids_m, sc_m = merge_topk_cuda(torch.stack([glob(a, ids_a), glob(b, ids_b)]), torch.stack([sc_a, sc_b]), k)
    assert torch.equal(ids_m, ids1.to(torch.int64)) and torch.equal(sc_m, sc1)
