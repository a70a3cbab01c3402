"""Sentence-level index of implementation 3 of the reference (team_run1.py): documents are split into sentences
(``text.split('.')``, :45-46), every non-empty sentence becomes a unit ``f"{docid}_{idx}"`` of an Okapi BM25 index with
duplicate-counting queries (``build_inverted_index`` :80-99, ``score_documents_for_query`` :173-199), and ranked
sentences are mapped back to their first-seen parent docs (:286-294).  Tokenisation of the sentences, the vocabulary,
the index and the sentence -> doc step all run on the GPU; language detection / stop words / punctuation stripping
(:49-77) are text preprocessing and stay with the caller (``preprocess``)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr
from .bm25 import BM25


def split_into_sentences(text):
    """team_run1.py:45-46."""
    return text.split('.')


class SentenceIndex:
    """``model``: BM25(variant="okapi", dedup_query=False) over the sentences; ``sentence_ids[i]`` = ``f"{docid}_{idx}"``;
    ``sentence_to_doc`` int32 tensor (sentence -> position in ``doc_ids``) on the device."""

    def __init__(self, model, sentence_ids, sentence_to_doc, doc_ids):
        self.model, self.sentence_ids, self.sentence_to_doc, self.doc_ids = model, sentence_ids, sentence_to_doc, doc_ids

    def score_documents_for_queries(self, args_list, top=100):
        """Batched score_documents_for_query (team_run1.py:173-199) over the sentence units:
        ``[(query_id, tokens), ...]`` -> ``[(query_id, [sentence_id, ...] <= top), ...]``."""
        from .functional import ScoreDocumentsContext, score_documents_for_queries
        return score_documents_for_queries(args_list, ScoreDocumentsContext(self.model, self.sentence_ids, top))

    def docs_of_ranked_sentences(self, ranked_sentences, k=10):
        """team_run1.py:286-294 for a batch: ranked sentence indices [Q, n] (-1 pads) -> list of docid lists (<= k)."""
        out = dedupe_sentences_to_docs(ranked_sentences, self.sentence_to_doc, k).cpu().numpy()
        return [[self.doc_ids[int(d)] for d in row if d >= 0] for row in out]


def build_sentence_index(docs, preprocess=None, k1=1.5, b=0.75, device=None):
    """build_inverted_index + the merge loop, team_run1.py:80-124, for the whole corpus at once.  ``docs`` is an
    iterable of ``{'docid', 'text'}``; ``preprocess(sentence) -> list[str]`` (default: the sentence is already
    preprocessed text and is tokenised with ``str.split()`` on the GPU).  Sentences without tokens are skipped but keep
    their index in the id, like :93-94."""
    sent_texts, sent_ids, s2d, doc_ids = [], [], [], []
    for di, doc in enumerate(docs):
        doc_ids.append(doc["docid"])
        for idx, sentence in enumerate(split_into_sentences(doc["text"])):
            if preprocess is not None:
                toks = preprocess(sentence)
                if not toks:
                    continue
                sentence = " ".join(toks)
            elif not sentence.strip():
                continue
            sent_texts.append(sentence)
            sent_ids.append(f"{doc['docid']}_{idx}")
            s2d.append(di)
    if not sent_texts:
        raise ZeroDivisionError("division by zero")            # avg_doc_length = sum / N with N == 0, :124
    model = BM25.from_texts(sent_texts, k1, b, variant="okapi", dedup_query=False, device=device)
    s2d_t = torch.tensor(s2d, dtype=torch.int32, device=model._device)
    return SentenceIndex(model, sent_ids, s2d_t, doc_ids)


def dedupe_sentences_to_docs(sentence_ids, sentence_to_doc, k=10):
    """team_run1.py:286-294: walk the ranked sentences, keep the first occurrence of every parent doc, stop at ``k``
    docs (k <= 32).  ``sentence_ids`` [Q, n] (best first, -1 pads; numpy or torch), ``sentence_to_doc`` maps a sentence
    index to its doc index -> int64[Q, k] on the GPU (-1 pads).  One warp per query in libbr_b200.so."""
    lib = _lib.load()
    m = torch.as_tensor(sentence_to_doc)
    dev = _lib.require_cuda(m.device if m.is_cuda else None)
    m = m.to(device=dev, dtype=torch.int32).contiguous()
    s = torch.as_tensor(sentence_ids).to(device=dev, dtype=torch.int64).contiguous()
    if s.dim() != 2:
        raise ValueError("sentence_ids must be [Q, n]")
    if not 1 <= int(k) <= 32:
        raise ValueError("k must be in [1, 32]")
    nq, n = s.shape
    with torch.cuda.device(dev):
        out = torch.empty((nq, int(k)), dtype=torch.int64, device=dev)
        check(lib.br_dedupe_first_docs(ptr(s), ptr(m), m.numel(), nq, n, int(k), ptr(out), _lib.stream_ptr(dev)),
              "br_dedupe_first_docs")
    return out
