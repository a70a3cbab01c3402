"""GPU: the sentence-level index of implementation 3 (team_run1.py:80-124) and the sentence -> doc step (:286-294)."""
import numpy as np
import pytest

from oracle import bm25_oracle as orc
from document_retrieval_b200 import build_sentence_index, dedupe_sentences_to_docs

pytestmark = pytest.mark.gpu


def _loop(ranked_row, s2d, k):
    seen, want = set(), []
    for s in ranked_row:                      # team_run1.py:286-294
        if s < 0:
            continue
        d = int(s2d[s])
        if d not in seen:
            want.append(d)
            seen.add(d)
        if len(want) >= k:
            break
    return want


@pytest.mark.parametrize("k", [1, 10, 32])
def test_dedupe_kernel_matches_the_reference_loop(k):
    rng = np.random.default_rng(3)
    s2d = rng.integers(0, 40, size=500)
    ranked = np.stack([rng.permutation(500)[:100] for _ in range(64)])
    ranked[3, 50:] = -1
    ranked[4, :] = -1
    ranked[5, ::2] = -1                        # holes inside the list
    s2d_few = s2d.copy()
    got = dedupe_sentences_to_docs(ranked, s2d_few, k=k).cpu().numpy()
    for i in range(ranked.shape[0]):
        assert got[i][got[i] >= 0].tolist() == _loop(ranked[i], s2d_few, k), i
        assert np.all(got[i][(got[i] >= 0).sum():] == -1)
    # all sentences of one doc: a single result
    one = dedupe_sentences_to_docs(ranked[:2], np.zeros(500, np.int32), k=10).cpu().numpy()
    assert one[:, 0].tolist() == [0, 0] and np.all(one[:, 1:] == -1)
    with pytest.raises(Exception):
        dedupe_sentences_to_docs(np.array([[700]]), s2d, k=10)             # sentence id out of range


def test_sentence_index_build_and_query():
    """Sentence units, ids f"{docid}_{idx}", empty sentences skipped with their index kept (team_run1.py:88-98); the
    top sentences of a query equal the oracle's score_documents_for_query over the same units; docs de-duplicated."""
    rng = np.random.default_rng(5)
    words = [f"w{i}" for i in range(60)]
    docs = []
    for d in range(80):
        sents = [" ".join(rng.choice(words, size=rng.integers(0, 9))) for _ in range(rng.integers(1, 6))]
        docs.append({"docid": f"doc{d}", "text": ". ".join(sents) + ("." if d % 2 else "")})
    ix = build_sentence_index(docs)
    # the reference's own split / skip rule
    want_ids, want_tok, want_doc = [], [], []
    for di, doc in enumerate(docs):
        for idx, s in enumerate(doc["text"].split(".")):
            toks = s.split()
            if not toks:
                continue
            want_ids.append(f"{doc['docid']}_{idx}")
            want_tok.append(toks)
            want_doc.append(di)
    assert ix.sentence_ids == want_ids and ix.sentence_to_doc.cpu().tolist() == want_doc
    assert ix.model.corpus_size == len(want_ids)
    vocab = {w: i for i, w in enumerate(ix.model.terms)}
    do = np.cumsum([0] + [len(t) for t in want_tok]).astype(np.int64)
    tk = np.array([vocab[w] for t in want_tok for w in t], np.int32)
    oix = orc.build_index(do, tk, len(vocab))
    queries = [(qi, list(rng.choice(words, size=6))) for qi in range(20)]
    got = ix.score_documents_for_queries(queries, top=100)
    ranked = np.full((len(queries), 100), -1, np.int64)
    for (qid, toks), (gid, sids) in zip(queries, got):
        ids, sc = orc.score_documents_for_query(oix, [vocab[w] for w in toks], top=100)
        assert gid == qid and sids == [want_ids[i] for i in ids]
        ranked[qid, :len(ids)] = ids
    top_docs = ix.docs_of_ranked_sentences(ranked, k=10)
    for qi in range(len(queries)):
        assert top_docs[qi] == [docs[d]["docid"] for d in _loop(ranked[qi], want_doc, 10)]
