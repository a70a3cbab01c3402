"""GPU: dense cosine top-k (tcgen05 GEMM, fused norm + top-k) and candidate re-rank against the
oracle restatement of team_run1.py:270-282 and its golden fixture.  Tolerance (north_star): cosine
scores within 1e-5 relative in fp32 (plus 2e-6 absolute for values near zero); ids identical except
where two reference scores are closer than that tolerance."""
import numpy as np
import pytest
import torch

from oracle import bm25_oracle as orc
from document_retrieval_b200 import synth

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 2e-6


def _check(ids, sims, docs_f32, q_f32, k):
    ref_ids, ref_sims = orc.cosine_topk(docs_f32, q_f32, k)
    d = torch.as_tensor(docs_f32)
    dn = d / (d.norm(dim=1, keepdim=True) + 1e-10)
    qn = torch.as_tensor(q_f32)
    qn = qn / (qn.norm(dim=1, keepdim=True) + 1e-10)
    full = (qn @ dn.T).numpy()
    np.testing.assert_allclose(sims, ref_sims, rtol=RTOL, atol=ATOL)
    for i in range(ids.shape[0]):
        if not np.array_equal(ids[i], ref_ids[i]):
            # any difference must be a near-tie in the reference
            np.testing.assert_allclose(full[i, ids[i]], ref_sims[i], rtol=RTOL, atol=ATOL)
            assert len(set(ids[i].tolist())) == k


def test_cosine_topk_golden(golden):
    from document_retrieval_b200.cosine import CosineIndex
    g = golden("cosine_small")
    docs = torch.from_numpy(g["docs_bf16"]).view(torch.bfloat16)
    qs = torch.from_numpy(g["queries_bf16"]).view(torch.bfloat16)
    ix = CosineIndex(docs)
    ids, sims = ix.topk(qs, 10)
    ids, sims = ids.cpu().numpy(), sims.cpu().numpy()
    np.testing.assert_allclose(sims, g["top_sims"], rtol=RTOL, atol=ATOL)
    assert np.array_equal(ids, g["top_ids"])


@pytest.mark.parametrize("n,d,nq,k", [(5000, 768, 300, 10), (777, 128, 5, 3), (130, 64, 257, 10), (20000, 384, 64, 32),
                                      (3000, 200, 500, 10),      # d not a multiple of the 64-wide k-block (TMA zero fill)
                                      (2000, 1024, 300, 10),     # d > 768: the query block does not fit -> multicast kernel
                                      (40000, 96, 1000, 5)])     # several L2 windows x several query blocks
def test_cosine_topk_random(n, d, nq, k):
    from document_retrieval_b200.cosine import CosineIndex
    g = torch.Generator().manual_seed(synth.ROOT_SEED + n)
    docs = torch.randn(n, d, generator=g).to(torch.bfloat16)
    qs = torch.randn(nq, d, generator=g).to(torch.bfloat16)
    docs[3] = 0                                              # zero-norm doc: e/(0+1e-10) = 0
    ix = CosineIndex(docs, doc_base=1000)
    ids, sims = ix.topk(qs, k)
    _check(ids.cpu().numpy() - 1000, sims.cpu().numpy(), docs.float().numpy(), qs.float().numpy(), k)


@pytest.mark.parametrize("env", [{"BR_COS_KERNEL": "mc"}, {"BR_COS_KERNEL": "plain"}, {"BR_COS_QS_BN": "128"},
                                 {"BR_COS_QS_BN": "160"}, {"BR_COS_QS_BN": "192"}, {"BR_COS_QS_WINDOW": "3"}])
def test_cosine_kernel_variants_agree(env, monkeypatch):
    """Every GEMM variant (query-stationary CTA pairs at each block width, 2-CTA multicast, one CTA per tile) returns
    the same ids and bit-identical similarities as the default: same K order of accumulation per element."""
    from document_retrieval_b200.cosine import CosineIndex
    g = torch.Generator().manual_seed(99)
    docs = torch.randn(9000, 320, generator=g).to(torch.bfloat16)
    qs = torch.randn(700, 320, generator=g).to(torch.bfloat16)
    ix = CosineIndex(docs)
    ids0, s0 = ix.topk(qs, 10)
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    ids1, s1 = ix.topk(qs, 10)
    assert torch.equal(ids0, ids1) and torch.equal(s0, s1)


def test_cosine_rerank_of_bm25_candidates():
    from document_retrieval_b200 import BM25
    from document_retrieval_b200.cosine import CosineIndex, rerank_bm25_with_cosine
    c = synth.make_config("C1", scale=0.3)
    m = BM25.from_token_ids(c["doc_offsets"], c["token_ids"], c["vocab"])
    g = torch.Generator().manual_seed(7)
    emb = torch.randn(c["n_docs"], 768, generator=g).to(torch.bfloat16)
    qe = torch.randn(c["q_offsets"].size - 1, 768, generator=g).to(torch.bfloat16)
    ix = CosineIndex(emb)
    ids, sims = rerank_bm25_with_cosine(m, ix, (c["q_terms"], c["q_offsets"]), qe, n_candidates=200, k=10)
    ids, sims = ids.cpu().numpy(), sims.cpu().numpy()
    cand, _ = m.retrieve_top_n_batch((c["q_terms"], c["q_offsets"]), 200)
    cand = cand.cpu().numpy()
    e32, q32 = emb.float().numpy(), qe.float().numpy()
    for i in range(0, ids.shape[0], 17):
        cc = cand[i][cand[i] >= 0]
        ri, rs = orc.cosine_topk(e32[cc], q32[i:i + 1], 10)
        np.testing.assert_allclose(sims[i], rs[0], rtol=RTOL, atol=ATOL)
        assert set(ids[i].tolist()) <= set(cc.tolist())
        if not np.array_equal(ids[i], cc[ri[0]]):
            assert np.allclose(np.sort(sims[i]), np.sort(rs[0]), rtol=RTOL, atol=ATOL)


def test_cosine_massive_ties_are_answered_exactly():
    """2,000 identical embedding rows: a query equal to them has 2,000 candidates tied at the top - more than the filter's
    1,024-slot lists hold - and must still get the exact answer (lowest row ids first), like every other query."""
    from document_retrieval_b200.cosine import CosineIndex
    g = torch.Generator().manual_seed(5)
    docs = torch.randn(6000, 128, generator=g).to(torch.bfloat16)
    dup = torch.arange(1000, 6000, 2)[:2000]
    docs[dup] = docs[999].clone()
    qs = torch.cat([docs[999:1000], torch.randn(40, 128, generator=g).to(torch.bfloat16)])
    ids, sims = CosineIndex(docs).topk(qs, 10)
    ids, sims = ids.cpu().numpy(), sims.cpu().numpy()
    want = np.sort(np.concatenate([[999], dup.numpy()]))[:10]
    assert np.array_equal(ids[0], want)
    np.testing.assert_allclose(sims[0], 1.0, rtol=1e-5)
    _check(ids[1:], sims[1:], docs.float().numpy(), qs[1:].float().numpy(), 10)
