"""document_retrieval_b200 - B200-native BM25 scoring / top-k / cosine re-rank behind the Python
surface of Harkeerat2002/document-retrieval's retrieval hot path."""
from .bm25 import BM25  # noqa: F401
from .routing import (evaluate_recall_at_k, retrieve_test_queries, retrieve_top_n_batch,  # noqa: F401
                      LanguageModels, build_language_models, mrr_recall_at_k)
from .sentences import SentenceIndex, build_sentence_index, dedupe_sentences_to_docs, split_into_sentences  # noqa: F401
from .functional import (compute_tf_df_and_avgdl, compute_idf, bm25_score,  # noqa: F401
                         rank_documents_with_cosine_similarity_and_bm25,
                         rank_documents_with_cosine_similarity_and_bm25_lang, per_language_recall,
                         score_documents_for_query, score_documents_for_queries, ScoreDocumentsContext,
                         set_context)

__all__ = ["BM25", "evaluate_recall_at_k", "retrieve_test_queries", "retrieve_top_n_batch", "LanguageModels", "build_language_models", "dedupe_sentences_to_docs", "SentenceIndex", "build_sentence_index", "split_into_sentences", "mrr_recall_at_k",
           "compute_tf_df_and_avgdl", "compute_idf", "bm25_score", "rank_documents_with_cosine_similarity_and_bm25", "rank_documents_with_cosine_similarity_and_bm25_lang", "per_language_recall", "score_documents_for_query", "score_documents_for_queries", "set_context",
           "ScoreDocumentsContext"]
