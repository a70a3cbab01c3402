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
