"""GPU: property tests on tiny random corpora - the cases the reference never tests but its
semantics define: out-of-vocabulary terms, duplicate query terms, empty docs and empty queries, k >= N, all-zero
scores, exact ties (identical docs -> doc id order), negative idf (okapi_no_plus1).  Every example is compared
bit-exactly (ids and float64 scores) with the numpy oracle restatement of the reference."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import bm25_oracle as orc

pytestmark = pytest.mark.gpu


@st.composite
def corpus_and_queries(draw):
    vocab = draw(st.integers(1, 12))
    n_docs = draw(st.integers(1, 40))
    term = st.integers(0, vocab - 1)
    docs = draw(st.lists(st.lists(term, min_size=0, max_size=8), min_size=n_docs, max_size=n_docs))
    if not any(docs):                                   # the reference divides by avgdl: keep one token in the corpus
        docs[0] = [0]
    if draw(st.booleans()) and n_docs > 2:              # identical docs -> exact ties
        docs[-1] = list(docs[0])
    qterm = st.integers(-1, vocab + 2)                  # -1 and >= vocab: out of vocabulary
    queries = draw(st.lists(st.lists(qterm, min_size=0, max_size=10), min_size=1, max_size=6))
    k = draw(st.integers(1, n_docs + 3))
    variant = draw(st.sampled_from(["notebook", "okapi", "okapi_no_plus1"]))
    dedup = draw(st.booleans())
    return vocab, docs, queries, k, variant, dedup


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(corpus_and_queries())
def test_topk_equals_oracle_on_tiny_corpora(case):
    from document_retrieval_b200 import BM25
    vocab, docs, queries, k, variant, dedup = case
    off = np.zeros(len(docs) + 1, np.int64)
    np.cumsum([len(d) for d in docs], out=off[1:])
    tok = np.asarray([t for d in docs for t in d], np.int32)
    ix = orc.build_index(off, tok, vocab)
    m = BM25.from_token_ids(off, tok, vocab, variant=variant, dedup_query=dedup)
    n = len(docs)
    kk = min(k, n)
    q_off = np.zeros(len(queries) + 1, np.int32)
    np.cumsum([len(q) for q in queries], out=q_off[1:])
    q_terms = np.asarray([t for q in queries for t in q], np.int32)
    ids, sc = m.retrieve_top_n_batch((q_terms, q_off), kk)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    for i, q in enumerate(queries):
        oi, os_ = orc.retrieve_top_n(ix, q, kk, variant=variant, dedup=dedup)
        assert np.array_equal(ids[i], oi), (case, i, ids[i], oi)
        assert np.array_equal(sc[i], os_), (case, i, sc[i], os_)
        # single-query surface, including n >= N -> full ranking (bm25_ranking.ipynb:208-209)
        if i == 0:
            full_i, _ = orc.retrieve_top_n(ix, q, min(k, n), variant=variant, dedup=dedup)
            assert np.array_equal(m.retrieve_top_n(np.asarray(q, np.int32), k), full_i)
