"""GPU: the BASELINE.json configurations at (close to) full size against the plain-C oracle.
C2: 7 per-language indexes (268,022 docs), 2,000 mixed-language queries, Recall@10 - ids bit-exact.
C3: BM25 top-1000 -> cosine re-rank (768-d bf16) on the C2 'en' corpus, query subsample.
C4: size-independent properties on a 1M-doc slice (sharded == single, fused == dense)."""
import numpy as np
import pytest
import torch

from oracle import bm25_oracle as orc
from oracle.c_oracle import COracle
from document_retrieval_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    return synth.make_c2(scale=1.0)


def test_c2_full_recall_and_ids(c2):
    from document_retrieval_b200 import BM25, evaluate_recall_at_k, retrieve_test_queries
    langs, queries = c2
    assert sum(c["n_docs"] for c in langs.values()) == 268_022 and len(queries) == 2_000
    models, maps, oracles = {}, {}, {}
    for lang, c in langs.items():
        models[lang] = BM25.from_token_ids(c["doc_offsets"], c["token_ids"], c["vocab"])
        maps[lang] = np.arange(c["n_docs"]) + 1_000_000 * (1 + list(langs).index(lang))
        oracles[lang] = COracle(c["doc_offsets"], c["token_ids"], c["vocab"])
        assert models[lang].avgdl == oracles[lang].avgdl
    rows = [dict(query=q["terms"], lang=q["lang"], positive_docs=int(maps[q["lang"]][q["qrel"]]), query_id=i)
            for i, q in enumerate(queries)]
    want = [None] * len(rows)
    for lang in langs:
        idx = [i for i, r in enumerate(rows) if r["lang"] == lang]
        terms = np.concatenate([rows[i]["query"] for i in idx]).astype(np.int32)
        offs = np.cumsum([0] + [rows[i]["query"].size for i in idx]).astype(np.int32)
        oi, _, _ = oracles[lang].topk_batch(terms, offs, 10)
        for j, i in enumerate(idx):
            want[i] = [int(maps[lang][d]) for d in oi[j]]
    got = retrieve_test_queries(models, maps, rows, k=10)
    assert [[int(x) for x in g] for g in got] == want                 # every top-10 list identical
    rec = evaluate_recall_at_k(models, maps, rows, k=10)
    assert rec == orc.recall_at_k(want, [r["positive_docs"] for r in rows]) and rec > 0.9


def test_c3_bm25_top1000_then_cosine(c2):
    from document_retrieval_b200 import BM25
    from document_retrieval_b200.cosine import CosineIndex, rerank_bm25_with_cosine
    c = c2[0]["en"]
    m = BM25.from_token_ids(c["doc_offsets"], c["token_ids"], c["vocab"])
    qo, qt, _ = synth.make_queries(c["doc_offsets"], c["token_ids"], 512, c["vocab"], (3, 3))
    g = torch.Generator(device="cuda").manual_seed(synth.ROOT_SEED + 3)
    emb = torch.randn(c["n_docs"], 768, generator=g, device="cuda").to(torch.bfloat16)
    qe = torch.randn(512, 768, generator=g, device="cuda").to(torch.bfloat16)
    ix = CosineIndex(emb)
    ids, sims = rerank_bm25_with_cosine(m, ix, (qt, qo), qe, n_candidates=1000, k=10)
    # BM25 candidates: identical to the oracle's top-1000
    co = COracle(c["doc_offsets"], c["token_ids"], c["vocab"])
    oi, _, _ = co.topk_batch(qt, qo, 1000)
    cand, _ = m.retrieve_top_n_batch((qt, qo), 1000)
    assert np.array_equal(cand.cpu().numpy(), oi)
    # cosine over those candidates: fp32 torch restatement of team_run1.py:270-282
    e32, q32 = emb.float(), qe.float()
    for i in range(0, 512, 37):
        cc = torch.from_numpy(oi[i].astype(np.int64)).cuda()
        d = e32[cc]
        d = d / (d.norm(dim=1, keepdim=True) + 1e-10)
        q = q32[i] / (q32[i].norm() + 1e-10)
        s = d @ q
        top = torch.topk(s, 10)
        np.testing.assert_allclose(sims[i].cpu().numpy(), top.values.cpu().numpy(), rtol=1e-5, atol=2e-6)
        got = ids[i].cpu().numpy().astype(np.int64)
        want = cc[top.indices].cpu().numpy()           # canonical side: torch.topk over fp32 of the bf16 values
        if not np.array_equal(got, want):
            # any difference must be a near-tie of the reference cosines (same tolerance as tests/test_gpu_cosine.py)
            pos = {int(d): j for j, d in enumerate(oi[i].tolist())}
            np.testing.assert_allclose(s[[pos[int(d)] for d in got]].cpu().numpy(), top.values.cpu().numpy(), rtol=1e-5, atol=2e-6)
            assert len(set(got.tolist())) == 10 and set(got.tolist()) <= set(oi[i].tolist())


@pytest.fixture(scope="module")
def c4_slice():
    from document_retrieval_b200 import BM25
    do, tk = synth.make_corpus_torch(1_000_000, 1_000_000, 60, "cuda", seed=5)
    qo, qt, _ = synth.make_queries_torch(do, tk, 2_000, 1_000_000, seed=6)
    return do, tk, qo, qt, BM25.from_token_ids(do, tk, 1_000_000)


def test_c4_slice_against_the_oracle(c4_slice):
    """1M-doc slice of the C4 shape, 2,000 queries: top-10 ids and float64 scores identical to the plain-C oracle,
    whatever the deferral budget of the tiled scorer (0 = every term streamed, 1000 = the most aggressive plan)."""
    do, tk, qo, qt, m = c4_slice
    co = COracle(do.cpu().numpy(), tk.cpu().numpy(), 1_000_000)
    oi, osc, _ = co.topk_batch(qt, qo, 10)
    for pm in (800, 0, 1000, 400):
        m.set_option("defer_pm", pm)
        ids, sc = m.retrieve_top_n_batch((qt, qo), 10)
        assert m.query_stats()["queries_fused"] > 1900
        assert np.array_equal(ids.cpu().numpy(), oi), f"defer_pm={pm}"
        assert np.array_equal(sc.cpu().numpy(), osc), f"defer_pm={pm}"
    m.set_option("defer_pm", 800)
    # duplicates counted (team_run1.py:183) and the top-100 of score_documents_for_query through the large-k tiled path
    o2 = COracle(do.cpu().numpy(), tk.cpu().numpy(), 1_000_000, variant="okapi")
    from document_retrieval_b200 import BM25
    m2 = BM25.from_token_ids(do, tk, 1_000_000, variant="okapi", dedup_query=False)
    n300 = int(qo[300])
    oi2, osc2, oc2 = o2.topk_batch(qt[:n300], qo[:301], 100, dedup=False, positive_only=True)
    ids2, sc2, cnt2 = m2.retrieve_top_n_batch((qt[:n300], qo[:301]), 100, positive_only=True, return_counts=True)
    assert np.array_equal(cnt2.cpu().numpy(), oc2)
    assert np.array_equal(ids2.cpu().numpy(), oi2) and np.array_equal(sc2.cpu().numpy(), osc2)


def test_c4_slice_properties(c4_slice):
    """1M-doc slice of the C4 shape: fused == dense, 4 fake shards == single index, save/load round trip."""
    import os
    import tempfile
    from document_retrieval_b200 import BM25
    from document_retrieval_b200.sharded import merge_topk_cuda, shard_bounds
    do, tk, qo, qt, m = c4_slice
    q = (qt, qo)
    ids_f, sc_f = m.retrieve_top_n_batch(q, 10)
    assert m.query_stats()["queries_fused"] > 1900
    m.set_option("fused", 0)
    ids_d, sc_d = m.retrieve_top_n_batch(q, 10)
    m.set_option("fused", 1)
    assert torch.equal(ids_f, ids_d) and torch.equal(sc_f, sc_d)
    # the large-k tiled path (radix-select tighten) at this scale: top-100 and top-1000 of 300 queries
    q300 = (qt[:int(qo[300])], qo[:301])
    for kk in (100, 1000):
        a = m.retrieve_top_n_batch(q300, kk)
        assert m.query_stats()["queries_fused"] > 280
        m.set_option("fused", 0)
        b = m.retrieve_top_n_batch(q300, kk)
        m.set_option("fused", 1)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    shards = []
    for lo, hi in shard_bounds(1_000_000, 4):
        o = do[lo:hi + 1] - do[lo]
        t = tk[int(do[lo]):int(do[hi])]
        shards.append(BM25.from_token_ids(o, t, 1_000_000, doc_base=lo, finalize=False))
    df = sum(s.local_df_tensor().to(torch.int64) for s in shards).cpu().numpy()
    n = sum(s.stats()["n_docs"] for s in shards)
    sdl = sum(s.stats()["sum_dl"] for s in shards)
    parts = []
    for s in shards:
        s.finalize(n, sdl, df)
        i, sc = s.retrieve_top_n_batch(q, 10)
        parts.append((torch.where(i >= 0, i.long() + s.doc_base, torch.full_like(i, -1, dtype=torch.long)), sc))
    ids_m, sc_m = merge_topk_cuda(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), 10)
    assert torch.equal(ids_m, ids_f.long()) and torch.equal(sc_m, sc_f)
    del shards
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "index.brix")
        m.save(path)
        m2 = BM25.load(path)
        i2, s2 = m2.retrieve_top_n_batch(q, 10)
        assert torch.equal(i2, ids_f) and torch.equal(s2, sc_f)


def test_index_file_roundtrip_against_the_oracle(tmp_path):
    """BM25.save / BM25.load (flat binary file, no pickle): the loaded index answers like the plain-C oracle built
    from the original tokens; string vocabulary, a doc shard's statistics in force, and corrupt files."""
    import json
    from document_retrieval_b200 import BM25, indexfile
    from document_retrieval_b200._lib import BRError
    c = synth.make_config("C1", scale=0.3)
    co = COracle(c["doc_offsets"], c["token_ids"], c["vocab"])
    oi, osc, _ = co.topk_batch(c["q_terms"], c["q_offsets"], 10)
    m = BM25.from_token_ids(c["doc_offsets"], c["token_ids"], c["vocab"])
    path = str(tmp_path / "c1.brix")
    m.save(path)
    assert open(path, "rb").read(8) == b"BRIX0001" and b"pickle" not in open(path, "rb").read(4096)
    m2 = BM25.load(path)
    ids, sc = m2.retrieve_top_n_batch((c["q_terms"], c["q_offsets"]), 10)
    assert np.array_equal(ids.cpu().numpy(), oi) and np.array_equal(sc.cpu().numpy(), osc)
    assert m2.avgdl == co.avgdl and m2.corpus_size == co.n_docs
    # string surface: terms survive as a byte pool, decoded lazily
    docs = synth.to_strings(c["doc_offsets"][:201], c["token_ids"][:int(c["doc_offsets"][200])])
    ms = BM25(docs)
    p2 = str(tmp_path / "s.brix")
    ms.save(p2)
    ml = BM25.load(p2)
    assert ml._terms is None and ml.terms == ms.terms
    q = docs[7][:5]
    assert np.array_equal(ml.retrieve_top_n(q, 10), ms.retrieve_top_n(q, 10))
    assert ml.df == ms.df and ml.idf == ms.idf
    # a doc shard keeps the global statistics it was finalised with
    half = 1500
    do, tk = c["doc_offsets"], c["token_ids"]
    sh = BM25.from_token_ids(do[:half + 1], tk[:int(do[half])], c["vocab"], finalize=False)
    full_df = np.diff(m._export_csr()["row_ptr"]).astype(np.int64)
    sh.finalize(m.corpus_size, int(m.stats()["sum_dl"]), full_df)
    p3 = str(tmp_path / "shard.brix")
    sh.save(p3)
    assert indexfile.read_header(p3)["shard"] is True
    sl = BM25.load(p3)
    a = sh.retrieve_top_n_batch((c["q_terms"], c["q_offsets"]), 10)
    b = sl.retrieve_top_n_batch((c["q_terms"], c["q_offsets"]), 10)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and sl.avgdl == m.avgdl
    # corrupt files fail loudly instead of reading out of bounds
    raw = bytearray(open(path, "rb").read())
    open(str(tmp_path / "trunc.brix"), "wb").write(raw[:len(raw) // 2])
    with pytest.raises(BRError):
        BM25.load(str(tmp_path / "trunc.brix"))
    h = indexfile.read_header(path)
    off = next(x["offset"] for x in h["arrays"] if x["name"] == "doc")
    bad = bytearray(raw)
    bad[off:off + 4] = (10 ** 9).to_bytes(4, "little")                    # doc id far outside [0, n_docs)
    open(str(tmp_path / "bad.brix"), "wb").write(bad)
    with pytest.raises(BRError):
        BM25.load(str(tmp_path / "bad.brix"))
    bad = bytearray(raw)
    bad[off + 4:off + 8] = bad[off:off + 4]                               # doc ids not ascending inside a list
    open(str(tmp_path / "dup.brix"), "wb").write(bad)
    with pytest.raises(BRError):
        BM25.load(str(tmp_path / "dup.brix"))
    with pytest.raises(BRError):
        BM25.load(__file__)
