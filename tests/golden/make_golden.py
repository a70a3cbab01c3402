"""Generate the golden fixtures in this directory from the UNMODIFIED reference functions.

Run in the authoring container (where /root/reference exists):

    PYTHONHASHSEED=0 python tests/golden/make_golden.py

The reference has no golden vectors of its own, so the fixtures are outputs of the
reference's own code (oracle/ref_loader.py exec/ast-extracts it, source untouched) on synth-v1
inputs.  The inputs are stored next to the outputs, so tests never need the generator or the
reference at run time.  ``PYTHONHASHSEED=0`` pins the ``set(query)`` iteration order the
reference sums in (bm25_ranking.ipynb:193).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

if os.environ.get("PYTHONHASHSEED") != "0":
    os.environ["PYTHONHASHSEED"] = "0"
    os.execv(sys.executable, [sys.executable] + sys.argv)

from document_retrieval_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402


def canon(scores, k, positive_only=False):
    ids = np.arange(scores.size)
    if positive_only:
        ids = ids[scores != 0]
    o = np.lexsort((ids, -scores[ids]))[:k]
    return ids[o]


def vocab_arrays(model_df, model_idf, vocab, prefix="t"):
    df = np.zeros(vocab, np.int64)
    idf = np.full(vocab, np.nan)
    for w, v in model_df.items():
        df[int(w[len(prefix):])] = v
    for w, v in model_idf.items():
        idf[int(w[len(prefix):])] = v
    return df, idf


def notebook_fixture(name, n_docs, vocab, mean_len, n_q, key, keep_scores):
    BM25 = ref_loader.notebook_bm25_class()
    do, tk = synth.make_corpus(n_docs, vocab, mean_len, key)
    qo, qt, rel = synth.make_queries(do, tk, n_q, vocab, key, oov_every=10)
    # duplicate-heavy and degenerate queries appended by hand
    extra = [np.array([tk[0], tk[0], tk[1], tk[0]], np.int32), np.array([vocab], np.int32),
             np.array([], np.int32)]
    for e in extra:
        qt = np.concatenate([qt, e]).astype(np.int32)
        qo = np.concatenate([qo, [qo[-1] + e.size]]).astype(np.int32)
    docs = synth.to_strings(do, tk)
    qs = synth.queries_to_strings(qo, qt, vocab)
    m = BM25(docs, k1=1.5, b=0.75)
    df, idf = vocab_arrays(m.df, m.idf, vocab)
    nq = len(qs)
    scores = np.stack([m.get_scores(q) for q in qs])
    raw_top = np.stack([m.retrieve_top_n(q, n=10) for q in qs])
    top_ids = np.stack([canon(s, 10) for s in scores])
    top_scores = np.take_along_axis(scores, top_ids, 1)
    # the reference's own top-n must be the canonical one up to order among exact ties
    for i in range(nq):
        assert sorted(scores[i][raw_top[i]].tolist()) == sorted(top_scores[i].tolist())
    full_rank = np.stack([m.retrieve_top_n(qs[i], n=n_docs + 5) for i in range(3)])  # n >= N branch :208
    out = dict(doc_offsets=do, token_ids=tk, vocab=vocab, q_offsets=qo, q_terms=qt, qrels=rel,
               avgdl=m.avgdl, corpus_size=m.corpus_size, df=df, idf=idf, raw_top=raw_top,
               top_ids=top_ids, top_scores=top_scores, full_rank=full_rank)
    if keep_scores:
        out["scores"] = scores
    # final_implementation.py:91-154 - same math with precomputed doc_lengths / injected idf
    F = ref_loader.final_bm25_class()
    f = F()
    f.corpus_size = len(docs)                       # injected like final_implementation.ipynb:451-458
    f.avgdl = m.avgdl
    f.build(docs, "en")
    f.doc_lengths = f.precompute_doc_lengths()
    f.precomputed_idf = f.precompute_idf()
    fs = np.stack([f.calculate_scores(q) for q in qs])
    assert np.array_equal(fs, scores), "final_implementation.BM25 differs from the notebook BM25"
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "docs", n_docs, "queries", nq, "nnz", int(df.sum()))
    return do, tk, qo, qt, m


def team_run1_fixture(do, tk, vocab, qo, qt):
    """score_documents_for_query (team_run1.py:173-199) on integer 'sentence ids'."""
    n_docs = do.size - 1
    inverted_index, doc_lengths = {}, {}
    for d in range(n_docs):
        toks = tk[do[d]:do[d + 1]]
        doc_lengths[d] = len(toks)
        u, c = np.unique(toks, return_counts=True)
        for t, n in zip(u.tolist(), c.tolist()):
            inverted_index.setdefault(t, {})[d] = n
    N = len(doc_lengths)
    avg = sum(doc_lengths.values()) / N
    fn = ref_loader.score_documents_for_query_fn(inverted_index, doc_lengths, N, avg)
    nq = qo.size - 1
    top = np.full((nq, 100), -1, np.int64)
    cnt = np.zeros(nq, np.int64)
    for i in range(nq):
        toks = [int(t) for t in qt[qo[i]:qo[i + 1]]]
        qid, docs = fn((i, toks))
        assert qid == i
        cnt[i] = len(docs)
        top[i, :len(docs)] = docs
    np.savez_compressed(os.path.join(HERE, "team_run1_top100.npz"), top=top, cnt=cnt, avg_doc_length=avg)
    print("team_run1_top100", nq, "queries; candidates/query min", cnt.min(), "max", cnt.max())


def rerank_fixture(do, tk, vocab, qo, qt):
    """compute_tf_df_and_avgdl / compute_idf / bm25_score /
    rank_documents_with_cosine_similarity_and_bm25 (cosine_similarity_bm25_reranking.py:129-238,
    extracted from the identical copy in query_ranking_and_embedding.py)."""
    import pandas as pd
    n_docs = do.size - 1
    nq = qo.size - 1
    docs = synth.to_strings(do, tk)
    qs = synth.queries_to_strings(qo, qt, vocab)
    with tempfile.TemporaryDirectory() as tmp:
        path = tmp + "/"
        ns = ref_loader.rerank_functions(path)
        corpus = pd.DataFrame({"docid": [f"d{i}" for i in range(n_docs)],
                               "preprocessed_text": [" ".join(d) for d in docs]})
        queries = pd.DataFrame({"id": list(range(nq)), "preprocessed_query": [" ".join(q) for q in qs]})
        tf_dict, df_dict, avgdl, num_docs = ns["compute_tf_df_and_avgdl"](corpus, path)
        assert sorted(os.listdir(tmp)) == ["avgdl.pkl", "df_dict.pkl", "num_docs.pkl", "tf_dict.pkl"]
        idf_dict = ns["compute_idf"](df_dict, num_docs)
        df = np.zeros(vocab, np.int64)
        idf = np.full(vocab, np.nan)
        for w, v in df_dict.items():
            df[int(w[1:])] = v
        for w, v in idf_dict.items():
            idf[int(w[1:])] = v
        # term order of tf_dict.keys() defines the tf-idf columns (term_index, :199) - irrelevant to
        # the cosine value, kept for completeness
        term_order = np.array([int(w[1:]) for w in tf_dict.keys()], np.int64)
        # bm25_score on (query, doc) pairs: the source doc + 7 pseudo-random docs per query
        rng = np.random.default_rng(7)
        pair_docs = rng.integers(0, n_docs, size=(nq, 8))
        pair_scores = np.zeros((nq, 8))
        for i in range(nq):
            for j in range(8):
                pair_scores[i, j] = ns["bm25_score"](qs[i], f"d{int(pair_docs[i, j])}", tf_dict, idf_dict, avgdl)
        ranked = ns["rank_documents_with_cosine_similarity_and_bm25"](corpus, queries, tf_dict, idf_dict, avgdl,
                                                                      batch_size=16)
        # the candidate stage the ranker does not return: its own embedding functions (unmodified) + the four lines
        # of cosine_similarity_bm25_reranking.py:210-211,222-226 -> float64 cosines, then `argsort()[::-1][:200]` (:229)
        from scipy.sparse import csr_matrix, vstack
        from scipy.sparse.linalg import norm
        term_index = {term: idx for idx, term in enumerate(tf_dict.keys())}
        emb, emb_ids = ns["create_tfidf_embedding"](corpus, tf_dict, idf_dict, term_index)
        emb = emb.tocsr()
        assert emb_ids == list(corpus["docid"])
        doc_norms = norm(emb, axis=1).reshape(-1, 1)
        normalized = emb.multiply(1 / doc_norms)
        qe = []
        for text in queries["preprocessed_query"]:
            q = ns["generate_query_embedding"](text, tf_dict, idf_dict, term_index)
            qe.append(q.multiply(1 / norm(q)))
        cos = normalized.dot(csr_matrix(vstack(qe)).T).toarray()          # [N, nq] float64
        cos = np.nan_to_num(cos, nan=0.0)
        n_c = min(200, n_docs)
        cos_top = np.stack([np.argsort(cos[:, i])[::-1][:n_c] for i in range(nq)]).astype(np.int64)
        cos_all = cos.T.copy()                                            # [nq, N]
        top10 = np.full((nq, 10), -1, np.int64)
        for i in range(nq):
            ids = [int(d[1:]) for d in ranked[i]]
            top10[i, :len(ids)] = ids
    np.savez_compressed(os.path.join(HERE, "rerank_v3.npz"), df=df, idf=idf, avgdl=avgdl, num_docs=num_docs,
                        term_order=term_order, pair_docs=pair_docs, pair_scores=pair_scores, top10=top10,
                        cos_top200=cos_top, cos_all=cos_all)
    print("rerank_v3", nq, "queries")


def edge_fixture():
    """Hand-made corner cases: exact ties, n >= N, a doc with one repeated token, negative idf."""
    BM25 = ref_loader.notebook_bm25_class()
    docs = [["a", "b", "c"], ["a", "b", "c"], ["a", "a", "a", "a"], ["d"], ["b", "c", "e", "e"],
            ["a", "b", "c"], ["f", "g"], ["a"]]
    vocab = {w: i for i, w in enumerate("abcdefg")}
    m = BM25(docs)
    queries = [["a"], ["a", "b"], ["e", "d"], ["zzz"], ["c", "c", "b"], ["g", "a", "f"]]
    scores = np.stack([m.get_scores(q) for q in queries])
    toks = np.array([vocab[w] for d in docs for w in d], np.int32)
    do = np.cumsum([0] + [len(d) for d in docs]).astype(np.int64)
    qt = np.array([vocab.get(w, len(vocab)) for q in queries for w in q], np.int32)
    qo = np.cumsum([0] + [len(q) for q in queries]).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "edge_small.npz"), doc_offsets=do, token_ids=toks, vocab=len(vocab),
                        q_offsets=qo, q_terms=qt, scores=scores, avgdl=m.avgdl)
    print("edge_small", scores.shape)


def cosine_fixture():
    """team_run1.py:270-282 restated verbatim with torch (normalise with +1e-10, matmul, topk)."""
    import torch
    g = torch.Generator().manual_seed(synth.ROOT_SEED + 5)
    docs = torch.randn(512, 64, generator=g).to(torch.bfloat16)
    qs = torch.randn(24, 64, generator=g).to(torch.bfloat16)
    d32, q32 = docs.float(), qs.float()
    dn = torch.stack([e / (e.norm() + 1e-10) for e in d32])                # :270-271
    ids, sims = [], []
    for q in q32:
        qn = q / (q.norm() + 1e-10)                                        # :276
        s = torch.matmul(dn, qn)                                           # :281
        top = torch.topk(s, k=10)                                          # :282
        ids.append(top.indices.numpy())
        sims.append(top.values.numpy())
    np.savez_compressed(os.path.join(HERE, "cosine_small.npz"), docs_bf16=docs.view(torch.int16).numpy(),
                        queries_bf16=qs.view(torch.int16).numpy(), top_ids=np.stack(ids), top_sims=np.stack(sims))
    print("cosine_small")


def main():
    assert ref_loader.available(), "reference checkout not found"
    if "--rerank-only" in sys.argv:           # refresh rerank_v3.npz alone (inputs come from the stored nb_small fixture)
        g = np.load(os.path.join(HERE, "nb_small.npz"))
        rerank_fixture(g["doc_offsets"], g["token_ids"], 400, g["q_offsets"], g["q_terms"])
        return
    do, tk, qo, qt, _ = notebook_fixture("nb_small", 300, 400, 30, 40, (101,), keep_scores=True)
    team_run1_fixture(do, tk, 400, qo, qt)
    rerank_fixture(do, tk, 400, qo, qt)
    notebook_fixture("nb_c1slice", 1000, 30_000, 200, 100, (1,), keep_scores=False)
    edge_fixture()
    cosine_fixture()


if __name__ == "__main__":
    main()
