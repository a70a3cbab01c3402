"""Module-level functions of the reference's hot path with their original signatures:
``compute_tf_df_and_avgdl`` / ``compute_idf`` (cosine_similarity_bm25_reranking.py:129-182),
``score_documents_for_query`` (team_run1.py:173-199)."""
from __future__ import annotations

import pickle

import numpy as np

from .bm25 import BM25


# tf_dict (by identity) -> (GPU model, docid list): lets the dict-signature functions of the reference
# (bm25_score, rank_documents_with_cosine_similarity_and_bm25) find the index their dicts came from.
_models_by_tf_dict = {}


def _model_for(tf_dict, corpus=None):
    hit = _models_by_tf_dict.get(id(tf_dict))
    if hit is not None and hit[2] is tf_dict:
        return hit[0], hit[1]
    if corpus is None:
        raise RuntimeError("this tf_dict did not come from compute_tf_df_and_avgdl of this package; pass the corpus "
                           "DataFrame (rank_documents_with_cosine_similarity_and_bm25) or rebuild it with compute_tf_df_and_avgdl")
    docs = [str(t).split() for t in corpus["preprocessed_text"]]
    model = BM25(docs, variant="okapi_no_plus1", dedup_query=False)
    doc_ids = list(corpus["docid"])
    _models_by_tf_dict[id(tf_dict)] = (model, doc_ids, tf_dict)
    return model, doc_ids


def compute_tf_df_and_avgdl(corpus_df, path_to_saved_file=None, device=None, return_model=False):
    """cosine_similarity_bm25_reranking.py:129-172 -> ``(tf_dict, df_dict, avgdl, num_docs)`` with
    term-major ``tf_dict = {term: {docid: tf}}``; the four pickles are written as a side effect when
    ``path_to_saved_file`` is given (:163-170).  tf / df / avgdl come from the GPU index build."""
    doc_ids = list(corpus_df["docid"])
    docs = [str(t).split() for t in corpus_df["preprocessed_text"]]
    model = BM25(docs, variant="okapi_no_plus1", dedup_query=False, device=device)
    c = model._export_csr()
    rp = c["row_ptr"]
    tf_dict = {}
    for t, term in enumerate(model.terms):
        lo, hi = rp[t], rp[t + 1]
        if hi > lo:
            tf_dict[term] = {doc_ids[d]: f for d, f in zip(c["doc"][lo:hi].tolist(), c["tf"][lo:hi].tolist())}
    df_dict = {term: int(rp[t + 1] - rp[t]) for t, term in enumerate(model.terms) if rp[t + 1] > rp[t]}
    avgdl = model.avgdl
    num_docs = len(doc_ids)
    if path_to_saved_file is not None:
        for name, obj in (("tf_dict", tf_dict), ("df_dict", df_dict), ("avgdl", avgdl), ("num_docs", num_docs)):
            with open(path_to_saved_file + name + ".pkl", "wb") as f:
                pickle.dump(obj, f)
    _models_by_tf_dict.clear()                      # one live corpus at a time, like the reference script
    _models_by_tf_dict[id(tf_dict)] = (model, doc_ids, tf_dict)
    if return_model:
        return tf_dict, df_dict, avgdl, num_docs, model
    return tf_dict, df_dict, avgdl, num_docs


def bm25_score(query_terms, doc_id, tf_dict, idf_dict, avgdl, k1=1.5, b=0.75):
    """cosine_similarity_bm25_reranking.py:185-195 for one (query, doc) pair -> float.  ``doc_length``
    is the sum of the query terms' tf in the doc (:187); idf without +1; duplicates counted.  The pair
    is scored by the CUDA re-rank kernel on the index ``tf_dict`` was derived from (idf_dict / avgdl
    are that index's own statistics)."""
    model, doc_ids = _model_for(tf_dict)
    if (k1, b) != (model.k1, model.b):
        raise ValueError("k1/b differ from the index the dicts came from")
    pos = getattr(model, "_docid_pos", None)
    if pos is None:
        pos = model._docid_pos = {d: i for i, d in enumerate(doc_ids)}
    local = pos.get(doc_id, -1)
    if local < 0:
        return 0
    out = model.rerank_scores_v3([list(query_terms)], [[local]])
    return float(out[0, 0].item())


def rank_documents_with_cosine_similarity_and_bm25(corpus, train_query, tf_dict, idf_dict, avgdl, batch_size=400,
                                                   n_candidates=200, k=10):
    """cosine_similarity_bm25_reranking.py:198-238 -> ``{query_id: [docid] * 10}``: sparse TF-IDF cosine
    over the whole corpus -> top-200 (:229) -> bm25_score of the 200 (:232-233) -> stable sort
    descending, first 10 (:234; ties keep the cosine order).  All three stages run on the GPU; only the
    final [Q, 200] stable sort uses torch.sort."""
    import torch
    model, doc_ids = _model_for(tf_dict, corpus)
    q_ids = list(train_query["id"])
    queries = [str(t).split() for t in train_query["preprocessed_query"]]
    out = {}
    n_cand = min(n_candidates, model.corpus_size)
    for s in range(0, len(queries), batch_size):
        qs = queries[s:s + batch_size]
        cand, _ = model.tfidf_cosine_top_n_batch(qs, n_cand)
        v3 = model.rerank_scores_v3(qs, cand)
        v3 = torch.where(cand >= 0, v3, torch.full_like(v3, float("-inf")))
        order = torch.sort(v3, dim=1, descending=True, stable=True).indices[:, :k]
        top = torch.gather(cand, 1, order).cpu().numpy()
        for qid, row in zip(q_ids[s:s + batch_size], top):
            out[qid] = [doc_ids[int(d)] for d in row if d >= 0]
    return out


def _lang_models_for(tf_dict, corpus, corpus_lang):
    """Per-language sub-indexes of the corpus ``tf_dict`` came from, finalised with the GLOBAL statistics (N, sum of doc
    lengths, df): the TF-IDF cosine and the re-rank score of a doc depend only on its own tf and on global idf / avgdl,
    so ranking inside one of these equals ranking the whole corpus and dropping the other languages."""
    import torch
    model, doc_ids = _model_for(tf_dict, corpus)
    cache = getattr(model, "_lang_models", None)
    if cache is not None and cache[0] is corpus_lang:
        return model, doc_ids, cache[1]
    c = model._export_csr()
    df_global = np.diff(c["row_ptr"]).astype(np.int64)
    n_global, sum_dl = model.corpus_size, int(c["dl"].sum())
    vocab = model.vocab
    by_lang = {}
    for i, d in enumerate(doc_ids):
        by_lang.setdefault(corpus_lang[d], []).append(i)
    texts = list(corpus["preprocessed_text"])
    out = {}
    for lang, idxs in by_lang.items():
        toks = [[vocab[w] for w in str(texts[i]).split()] for i in idxs]
        off = np.zeros(len(idxs) + 1, np.int64)
        np.cumsum([len(t) for t in toks], out=off[1:])
        flat = np.fromiter((t for ts in toks for t in ts), dtype=np.int32, count=int(off[-1]))
        m = BM25.from_token_ids(off, flat, model.vocab_size, model.k1, model.b, variant="okapi_no_plus1",
                                dedup_query=False, device=model._device, finalize=False)
        m.finalize(n_global, sum_dl, df_global)
        m.terms = model.terms
        out[lang] = (m, [doc_ids[i] for i in idxs])
    model._lang_models = (corpus_lang, out)
    return model, doc_ids, out


def rank_documents_with_cosine_similarity_and_bm25_lang(corpus, train_query, tf_dict, idf_dict, avgdl, corpus_lang,
                                                        batch_size=400, n_candidates=1000, k=100):
    """The language-filtered variant, text_preprocessing_and_embedding_setup.py:264-389: candidates are the first
    ``n_candidates`` (1000, :342) docs OF THE QUERY'S LANGUAGE in descending TF-IDF cosine order (:333-343), re-scored
    with bm25_score (:345-346) and sorted descending, first ``k`` (100, :349-352).  ``corpus_lang`` maps docid ->
    language; ``train_query`` has columns id / preprocessed_query / lang.
    -> ``(ranked_documents_dict, query_lang_dict)`` like the reference.  Each language is a GPU sub-index finalised with
    the global statistics, so no other-language doc is ever scored."""
    import torch
    _, _, lang_models = _lang_models_for(tf_dict, corpus, corpus_lang)
    q_ids = list(train_query["id"])
    q_lang = list(train_query["lang"])
    queries = [str(t).split() for t in train_query["preprocessed_query"]]
    ranked, query_lang_dict = {}, {}
    by_lang = {}
    for i, lang in enumerate(q_lang):
        by_lang.setdefault(lang, []).append(i)
        query_lang_dict[q_ids[i]] = lang
    for lang, idxs in by_lang.items():
        if lang not in lang_models:
            for i in idxs:
                ranked[q_ids[i]] = []
            continue
        m, ids_l = lang_models[lang]
        n_cand = min(n_candidates, m.corpus_size)
        for s in range(0, len(idxs), batch_size):
            chunk = idxs[s:s + batch_size]
            qs = [queries[i] for i in chunk]
            cand, _ = m.tfidf_cosine_top_n_batch(qs, n_cand)
            v3 = m.rerank_scores_v3(qs, cand)
            v3 = torch.where(cand >= 0, v3, torch.full_like(v3, float("-inf")))
            order = torch.sort(v3, dim=1, descending=True, stable=True).indices[:, :k]
            top = torch.gather(cand, 1, order).cpu().numpy()
            for i, row in zip(chunk, top):
                ranked[q_ids[i]] = [ids_l[int(d)] for d in row if d >= 0]
    return {qid: ranked[qid] for qid in q_ids}, query_lang_dict


def per_language_recall(ranked_docs, positive_docs, query_lang_dict, query_ids=None):
    """text_preprocessing_and_embedding_setup.py:534-562: overall fraction of queries whose positive doc is in the
    ranked list and the same per language.  ``ranked_docs`` {query_id: [docid]}; ``positive_docs`` aligned with
    ``query_ids`` (default: the dict's order).  -> ``(overall, {lang: fraction})``"""
    query_ids = list(ranked_docs.keys()) if query_ids is None else list(query_ids)
    per_q, per_hit, hits = {}, {}, 0
    for qid, pos in zip(query_ids, positive_docs):
        lang = query_lang_dict[qid]
        per_q[lang] = per_q.get(lang, 0) + 1
        per_hit.setdefault(lang, 0)
        if pos in ranked_docs[qid]:
            hits += 1
            per_hit[lang] += 1
    n = len(query_ids)
    return (hits / n if n else 0), {lang: per_hit[lang] / per_q[lang] for lang in per_q}


def compute_idf(df_dict, num_docs):
    """cosine_similarity_bm25_reranking.py:176-182: ``{term: np.log((N - df + .5) / (df + .5))}`` - no
    +1, negative when df > N/2.  A dict-to-dict host function (V logs); the device-side idf of an
    index lives in ``BM25(variant="okapi_no_plus1").idf``."""
    terms = list(df_dict.keys())
    df = np.fromiter((df_dict[t] for t in terms), dtype=np.float64, count=len(terms))
    idf = np.log((num_docs - df + 0.5) / (df + 0.5))
    return {t: v for t, v in zip(terms, idf.tolist())}


class ScoreDocumentsContext:
    """What ``score_documents_for_query`` reads from module globals in the reference
    (``inverted_index``, ``doc_lengths``, ``N``, ``avg_doc_length``, team_run1.py:88-124): here one
    okapi-variant index with duplicate-counting queries."""

    def __init__(self, model: BM25, doc_ids=None, top=100):
        if model.variant != "okapi" or model.dedup_query:
            raise ValueError("score_documents_for_query needs BM25(variant='okapi', dedup_query=False)")
        self.model, self.doc_ids, self.top = model, doc_ids, top


_context = None


def set_context(ctx):
    global _context
    _context = ctx


def score_documents_for_queries(args_list, ctx=None):
    """Batched form: [(query_id, query_tokens), ...] -> [(query_id, [doc ids] <= 100), ...]."""
    ctx = ctx or _context
    if ctx is None:
        raise RuntimeError("score_documents_for_query: call set_context(ScoreDocumentsContext(...)) first")
    if not args_list:
        return []
    k = min(ctx.top, ctx.model.corpus_size)
    ids, _, cnt = ctx.model.retrieve_top_n_batch([a[1] for a in args_list], k, positive_only=True,
                                                 return_counts=True)
    ids, cnt = ids.cpu().numpy(), cnt.cpu().numpy()
    out = []
    for (qid, _), row, c in zip(args_list, ids, cnt):
        docs = row[:c].tolist()
        out.append((qid, [ctx.doc_ids[d] for d in docs] if ctx.doc_ids is not None else docs))
    return out


def score_documents_for_query(args):
    """team_run1.py:173-199: ``(query_id, query_tokens)`` -> ``(query_id, top_docs)``: okapi BM25,
    duplicates counted, only docs with a hit, ``heapq.nlargest(100, ...)`` order (ties by doc id
    here).  Do not fork after CUDA initialisation (the reference runs this under process_map,
    :202): use ``score_documents_for_queries`` for the whole list instead."""
    return score_documents_for_queries([args])[0]
