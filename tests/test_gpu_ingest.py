"""GPU text ingestion (csrc/br_ingest.cu) against Python's own ``str.split()`` and dict insertion order -
the reference's tokenisation step (bm25_ranking.ipynb:299) and vocabulary growth (:180-186), plus the
2-gram expansion of :105-107.  Bit-exact: same token boundaries, same term ids, same strings."""
import pickle

import numpy as np
import pytest

import document_retrieval_b200 as dr
from document_retrieval_b200 import ingest

pytestmark = pytest.mark.gpu

WS = [" ", "  ", "\t", "\n", "\r\n", "\x0b", "\x0c", "\x1c", "\x1d", "\x1e", "\x1f", "\x85", "\xa0", "\u1680", "\u2000",
      "\u2001", "\u2005", "\u200a", "\u2028", "\u2029", "\u202f", "\u205f", "\u3000"]
# not whitespace for str.split(): zero-width space, Mongolian vowel separator, word joiner, BOM, and characters whose
# UTF-8 bytes contain the byte values of the whitespace encodings (C3 85, C3 A0, E2 80 8B, E1 9A 81, F0 9A 80 80 ...)
NOT_WS = ["\u200b", "\u180e", "\u2060", "\ufeff", "\xc5", "\xe0", "\u1681", "\u200c", "\u2027", "\u205e", "\u3001",
          "\U0001a000", "\x00", "\x7f", "\x1b"]
WORDS = ["a", "b", "the", "Haus", "maison", "été", "naïve", "한국어", "문서", "日本", "😀", "x_y", "x", "y", "_", "a_b",
         "Zürich", "don't", "42", "🙂🙃"] + NOT_WS


def expected(texts, bigrams=False):
    toks = []
    for t in texts:
        w = t.split() if isinstance(t, str) else []
        if bigrams and len(w) >= 2:
            w = w + ["_".join(g) for g in zip(w, w[1:])]          # == nltk.ngrams(tokens, 2)
        toks.append(w)
    vocab = {}
    ids = [vocab.setdefault(w, len(vocab)) for d in toks for w in d]
    off = np.zeros(len(toks) + 1, np.int64)
    np.cumsum([len(d) for d in toks], out=off[1:])
    return toks, list(vocab), np.asarray(ids, np.int32), off


def random_texts(rng, n, max_tok=12):
    out = []
    for _ in range(n):
        m = int(rng.integers(0, max_tok + 1))
        s = rng.choice(WS) if rng.random() < 0.3 else ""
        for j in range(m):
            w = "".join(rng.choice(WORDS) for _ in range(int(rng.integers(1, 3))))
            s += w + (rng.choice(WS) if (j + 1 < m or rng.random() < 0.4) else "")
        out.append(s)
    return out


@pytest.mark.parametrize("bigrams", [False, True])
def test_tokens_and_vocabulary_equal_python(bigrams):
    rng = np.random.default_rng(7)
    texts = random_texts(rng, 3000) + ["", " ", "\u3000", "single", "a b", " a  b ", "a_b a b a_b", "x" * 5000 + " y"]
    texts[17] = None
    texts[18] = 3.5
    _, terms, ids, off = expected(texts, bigrams)
    voc, d_off, d_ids = ingest.Vocabulary.from_texts(texts, bigrams=bigrams)
    assert np.array_equal(d_off.cpu().numpy(), off)
    assert voc.terms == terms
    assert np.array_equal(d_ids.cpu().numpy(), ids)
    assert len(voc) == len(terms)


def test_every_whitespace_and_lookalike_character():
    texts = [f"l{w}r" for w in WS] + [f"l{c}r" for c in NOT_WS] + ["".join(WS), "".join(NOT_WS), "".join(WS) + "z" + "".join(WS)]
    _, terms, ids, off = expected(texts)
    voc, d_off, d_ids = ingest.Vocabulary.from_texts(texts)
    assert np.array_equal(d_off.cpu().numpy(), off)
    assert voc.terms == terms
    assert np.array_equal(d_ids.cpu().numpy(), ids)


@pytest.mark.parametrize("bigrams", [False, True])
def test_query_lookup_and_oov(bigrams):
    rng = np.random.default_rng(11)
    texts = random_texts(rng, 500)
    _, terms, _, _ = expected(texts, bigrams)
    vocab = {w: i for i, w in enumerate(terms)}
    voc, _, _ = ingest.Vocabulary.from_texts(texts, bigrams=bigrams)
    queries = random_texts(rng, 200, 6) + ["neverseen a", "a neverseen b", "", "  "]
    q_toks, _, _, q_off = expected(queries, bigrams)[0], None, None, expected(queries, bigrams)[3]
    want = np.asarray([vocab.get(w, -1) for q in q_toks for w in q], np.int32)
    ids, off = voc.encode_texts(queries)
    assert np.array_equal(off.cpu().numpy(), q_off)
    assert np.array_equal(ids.cpu().numpy(), want)
    assert (want == -1).any() and (want >= 0).any()
    # pickling keeps ids (the hash table is rebuilt from the term strings)
    voc2 = pickle.loads(pickle.dumps(voc))
    ids2, _ = voc2.encode_texts(queries)
    assert np.array_equal(ids2.cpu().numpy(), want)
    assert voc2.terms == terms


def test_empty_inputs():
    voc, off, ids = ingest.Vocabulary.from_texts(["", "   "])
    assert len(voc) == 0 and ids.numel() == 0 and off.cpu().tolist() == [0, 0, 0]
    assert voc.terms == []
    i2, o2 = voc.encode_texts(["a b"])
    assert i2.cpu().tolist() == [-1, -1] and o2.cpu().tolist() == [0, 2]
    voc0, off0, ids0 = ingest.Vocabulary.from_texts([])
    assert len(voc0) == 0 and off0.cpu().tolist() == [0]


def test_from_texts_model_equals_token_list_model():
    rng = np.random.default_rng(3)
    words = [f"w{i}" for i in range(400)]
    p = 1.0 / np.arange(1, 401)
    p /= p.sum()
    texts = [" ".join(rng.choice(words, size=int(rng.integers(3, 40)), p=p)) for _ in range(1500)]
    queries = [" ".join(rng.choice(words, size=int(rng.integers(2, 8)), p=p)) for _ in range(64)] + ["zzz w1", ""]
    a = dr.BM25([t.split() for t in texts])
    b = dr.BM25.from_texts(texts)
    assert b.corpus_size == a.corpus_size and b.avgdl == a.avgdl
    assert b.df == a.df and list(b.df) == list(a.df)
    assert b.idf == a.idf
    ids_a, sc_a = a.retrieve_top_n_batch([q.split() for q in queries], 10)
    ids_b, sc_b = b.retrieve_top_n_texts(queries, 10)
    assert np.array_equal(ids_a.cpu().numpy(), ids_b.cpu().numpy())
    assert np.array_equal(sc_a.cpu().numpy(), sc_b.cpu().numpy())
    # list[str] queries still work on a from_texts model (terms decoded lazily), and texts on a list model
    assert np.array_equal(b.retrieve_top_n(queries[0].split(), 5), a.retrieve_top_n(queries[0].split(), 5))
    ids_c, _ = a.retrieve_top_n_texts(queries, 10)
    assert np.array_equal(ids_a.cpu().numpy(), ids_c.cpu().numpy())


def test_bigram_model_equals_expanded_token_lists():
    rng = np.random.default_rng(5)
    words = [f"m{i}" for i in range(60)]
    texts = [" ".join(rng.choice(words, size=int(rng.integers(1, 25)))) for _ in range(800)]
    toks, _, _, _ = expected(texts, bigrams=True)
    a = dr.BM25(toks)
    b = dr.BM25.from_texts(texts, bigrams=True)
    assert b.df == a.df and list(b.df) == list(a.df)
    queries = texts[:40]
    ids_a, sc_a = a.retrieve_top_n_batch(expected(queries, True)[0], 10)
    ids_b, sc_b = b.retrieve_top_n_texts(queries, 10)
    assert np.array_equal(ids_a.cpu().numpy(), ids_b.cpu().numpy())
    assert np.array_equal(sc_a.cpu().numpy(), sc_b.cpu().numpy())


def test_large_corpus_matches_factorize():
    """2M tokens: ids against pandas.factorize (first-seen order) on the split tokens."""
    import pandas as pd
    rng = np.random.default_rng(1)
    V = 50000
    p = 1.0 / np.arange(1, V + 1)
    p /= p.sum()
    lens = rng.integers(20, 100, size=40000)
    flat = rng.choice(V, size=int(lens.sum()), p=p)
    words = np.asarray([f"t{i}" for i in range(V)], dtype=object)[flat]
    off = np.zeros(lens.size + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    texts = [" ".join(words[off[i]:off[i + 1]]) for i in range(lens.size)]
    codes, uniq = pd.factorize(words)
    voc, d_off, d_ids = ingest.Vocabulary.from_texts(texts)
    assert np.array_equal(d_off.cpu().numpy(), off)
    assert np.array_equal(d_ids.cpu().numpy(), codes.astype(np.int32))
    assert voc.terms == list(uniq)


def test_eval_loop_takes_query_strings_through_the_gpu_tokeniser():
    """evaluate_recall_at_k / retrieve_test_queries with preprocessed query *strings* (bm25_ranking.ipynb:341-347) give the
    same answers as with token lists; a pickled bigram model still expands query texts."""
    rng = np.random.default_rng(9)
    words = [f"k{i}" for i in range(200)]
    models, id_maps, rows_s, rows_t = dr.LanguageModels(), {}, [], []
    for lang, n in (("en", 600), ("fr", 300)):
        texts = [" ".join(rng.choice(words, size=int(rng.integers(4, 30)))) for _ in range(n)]
        models[lang] = dr.BM25.from_texts(texts) if lang == "en" else dr.BM25([t.split() for t in texts])
        id_maps[lang] = [f"{lang}-{i}" for i in range(n)]
        for j in rng.integers(0, n, size=40):
            q = " ".join(texts[j].split()[:6])
            rows_s.append({"query": "  " + q + "\t", "lang": lang, "positive_docs": f"{lang}-{j}"})
            rows_t.append({"query": q.split(), "lang": lang, "positive_docs": f"{lang}-{j}"})
    rows_s.append({"query": "k1 k2", "lang": "xx", "positive_docs": "none"})
    rows_t.append({"query": ["k1", "k2"], "lang": "xx", "positive_docs": "none"})
    assert dr.evaluate_recall_at_k(models, id_maps, rows_s, 10) == dr.evaluate_recall_at_k(models, id_maps, rows_t, 10)
    assert dr.retrieve_test_queries(models, id_maps, rows_s, 10) == dr.retrieve_test_queries(models, id_maps, rows_t, 10)
    texts = [" ".join(rng.choice(words[:30], size=int(rng.integers(2, 12)))) for _ in range(300)]
    b = dr.BM25.from_texts(texts, bigrams=True)
    b2 = pickle.loads(pickle.dumps(b))
    ids1, sc1 = b.retrieve_top_n_texts(texts[:50], 10)
    ids2, sc2 = b2.retrieve_top_n_texts(texts[:50], 10)
    assert b2.bigrams and np.array_equal(ids1.cpu().numpy(), ids2.cpu().numpy()) and np.array_equal(sc1.cpu().numpy(), sc2.cpu().numpy())


def test_build_language_models_equals_the_notebook_loop():
    """bm25_ranking.ipynb:276-316 (group by language, text.split(), BM25 per language) through one call."""
    rng = np.random.default_rng(12)
    langs_all = ["en", "fr", "ko"]
    corpus, langs, docids = [], [], []
    for i in range(900):
        lang = langs_all[int(rng.integers(0, 3))]
        corpus.append(" ".join(f"{lang}{int(t)}" for t in rng.zipf(1.3, size=int(rng.integers(3, 40))) if t < 500))
        langs.append(lang)
        docids.append(f"doc-{i}")
    models = dr.build_language_models(corpus, langs, docids)
    for lang in langs_all:
        idx = [i for i, x in enumerate(langs) if x == lang]
        ref = dr.BM25([corpus[i].split() for i in idx])
        assert models.doc_id_maps[lang] == [docids[i] for i in idx]
        assert models[lang].df == ref.df and models[lang].avgdl == ref.avgdl
        qs = [corpus[i].split()[:5] for i in idx[:30] if corpus[i].split()]
        a, sa = models[lang].retrieve_top_n_batch(qs, 10)
        b_, sb = ref.retrieve_top_n_batch(qs, 10)
        assert np.array_equal(a.cpu().numpy(), b_.cpu().numpy()) and np.array_equal(sa.cpu().numpy(), sb.cpu().numpy())
    rows = [{"query": " ".join(corpus[i].split()[:6]), "lang": langs[i], "positive_docs": docids[i]} for i in range(0, 900, 9)]
    r = dr.evaluate_recall_at_k(models, models.doc_id_maps, rows, 10)
    assert 0.5 < r <= 1.0
