"""Per-language routing + Recall@10 evaluation (the L4 layer of the reference):
``bm25_models[lang]`` / ``doc_id_maps[lang]`` (bm25_ranking.ipynb:282-316), ``evaluate_recall_at_k``
(:329-354), ``retrieve_test_queries`` (:368-389) and ``retrieve_top_n_batch``
(final_implementation.py:179-181).

The reference walks the DataFrame row by row; here queries are grouped by language and each
group goes through one batched GPU call, which changes nothing in the results (queries are
independent).  Text preprocessing (nltk / konlpy, bm25_ranking.ipynb:84-110) is out of scope:
pass ``preprocess`` (``(text, lang) -> list[str]``); by default a query that is already a token
list is used as is and a string is ``.split()``.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np

from ._lib import VocabularyNotDistinct


def _default_preprocess(query, lang):
    if isinstance(query, str):
        return query.split()
    return list(query)


def _rows(data):
    """DataFrame | list[dict] -> list of dict-like rows in order (DataFrame.iterrows order)."""
    if hasattr(data, "iterrows"):
        return [row for _, row in data.iterrows()]
    return list(data)


class LanguageModels(dict):
    """``bm25_models`` with the matching ``doc_id_maps`` attached (bm25_ranking.ipynb:282-316)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.doc_id_maps = {}

    def add(self, lang, model, doc_ids):
        self[lang] = model
        self.doc_id_maps[lang] = doc_ids
        return self


def build_language_models(preprocessed_corpus, document_langs, document_ids, k1=1.5, b=0.75, **kw):
    """The per-language build loop of bm25_ranking.ipynb:276-316 in one call: group the preprocessed texts by language,
    tokenise each group with ``text.split()`` (:299) and build its ``BM25`` (:306) - here through ``BM25.from_texts``, i.e.
    tokenisation, vocabulary and index build on the GPU.  -> ``LanguageModels`` (``bm25_models`` with ``doc_id_maps``
    attached, local doc index -> ``docid`` in corpus order like :288)."""
    from .bm25 import BM25
    lang_to_idx = defaultdict(list)
    for idx, lang in enumerate(document_langs):              # :276-280
        lang_to_idx[lang].append(idx)
    models = LanguageModels()
    for lang, idxs in lang_to_idx.items():
        texts = [preprocessed_corpus[i] for i in idxs]
        models.add(lang, BM25.from_texts(texts, k1, b, **kw), [document_ids[i] for i in idxs])
    return models


def _batched_top(bm25_models, rows, k, preprocess, batch_size):
    """-> list (aligned with rows) of np.ndarray local ids, or None when the language is unknown."""
    by_lang = defaultdict(list)
    for i, row in enumerate(rows):
        by_lang[row["lang"]].append(i)
    out = [None] * len(rows)
    for lang, idxs in by_lang.items():
        if lang not in bm25_models:              # bm25_ranking.ipynb:336-337 / :374-376
            continue
        model = bm25_models[lang]
        for s in range(0, len(idxs), batch_size):
            chunk = idxs[s:s + batch_size]
            kk = min(k, model.corpus_size)
            if kk < 1:
                continue
            qs = [rows[i]["query"] for i in chunk]
            has_vocab = getattr(model, "vocabulary", None) is not None or getattr(model, "_terms", None) is not None
            if preprocess is _default_preprocess and all(isinstance(q, str) for q in qs) and has_vocab and \
                    hasattr(model, "retrieve_top_n_texts"):
                # already-preprocessed query strings: `preprocessed_query.split()` (bm25_ranking.ipynb:341-347) and the
                # vocabulary lookup run on the GPU for the whole chunk
                try:
                    ids, _ = model.retrieve_top_n_texts(qs, kk)
                except VocabularyNotDistinct:        # term strings collide: look the tokens up on the host instead
                    ids, _ = model.retrieve_top_n_batch([preprocess(q, lang) for q in qs], kk)
            else:
                ids, _ = model.retrieve_top_n_batch([preprocess(q, lang) for q in qs], kk)
            ids = ids.cpu().numpy()
            for j, i in enumerate(chunk):
                out[i] = ids[j][ids[j] >= 0]
    return out


def evaluate_recall_at_k(bm25_models, doc_id_maps, val_data, k=10, preprocess=None, batch_size=4096):
    """bm25_ranking.ipynb:329-354: hits / len(val_data); queries of an unknown language are skipped
    but still counted in the denominator (:331,353)."""
    rows = _rows(val_data)
    total = len(rows)
    top = _batched_top(bm25_models, rows, k, preprocess or _default_preprocess, batch_size)
    recall_count = 0
    for row, ids in zip(rows, top):
        if ids is None:
            continue
        doc_ids = doc_id_maps[row["lang"]]
        retrieved = [doc_ids[int(i)] for i in ids]
        if row["positive_docs"] in retrieved:
            recall_count += 1
    return recall_count / total if total > 0 else 0


def retrieve_test_queries(bm25_models, doc_id_maps, test_df, k=10, preprocess=None, batch_size=4096):
    """bm25_ranking.ipynb:368-389 -> list[list[docid]]; ``[]`` for an unknown language."""
    rows = _rows(test_df)
    top = _batched_top(bm25_models, rows, k, preprocess or _default_preprocess, batch_size)
    out = []
    for row, ids in zip(rows, top):
        if ids is None:
            out.append([])
            continue
        doc_ids = doc_id_maps[row["lang"]]
        out.append([doc_ids[int(i)] for i in ids])
    return out


def retrieve_top_n_batch(args):
    """final_implementation.py:179-181: ``(bm25_model, tokenized_query_batch, k)`` ->
    ``[np.ndarray of local indices, ...]``."""
    bm25_model, tokenized_query_batch, k = args
    if len(tokenized_query_batch) == 0:
        return []
    if k >= bm25_model.corpus_size:
        return [bm25_model.retrieve_top_n(q, n=k) for q in tokenized_query_batch]
    ids, _ = bm25_model.retrieve_top_n_batch(tokenized_query_batch, k)
    return [r.astype(np.int64) for r in ids.cpu().numpy()]


def mrr_recall_at_k(ranked_docs, relevant_docs, k_values=(1, 5, 10)):
    """team_run1.py:297-325: mean MRR@k and Recall@k.  ``ranked_docs`` list of ranked id lists,
    ``relevant_docs`` list of (lists of) relevant ids -> {k: (mrr, recall)}."""
    out = {}
    for k in k_values:
        mrr, rec = [], []
        for ranked, rel in zip(ranked_docs, relevant_docs):
            rel = [rel] if isinstance(rel, (str, int, np.integer)) else list(rel)
            top = [d for d in list(ranked)[:k]]
            r = next((i + 1 for i, d in enumerate(top) if d in rel), None)
            mrr.append(1 / r if r else 0)
            rec.append(len(set(rel) & set(top)) / len(rel) if rel else 0)
        out[k] = (float(np.mean(mrr)) if mrr else 0.0, float(np.mean(rec)) if rec else 0.0)
    return out
