"""Dense cosine similarity / re-rank on bf16 embeddings - the L3' layer of the reference
(team_run1.py:269-295: ``e / (e.norm() + 1e-10)`` on both sides, ``torch.matmul``, ``torch.topk``) on
the B200: a bf16 tcgen05 GEMM with the L2-normalisation and the top-k filter fused into its epilogue
(brute force over all docs), and a gather kernel for per-query candidate lists (BM25 top-1000 ->
cosine top-10).  Host code here only moves tensors; all arithmetic is in libbr_b200.so."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr


def _bf16(x, dev):
    t = torch.as_tensor(x)
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    return t.to(dev).contiguous()


def set_cosine_option(name, value):
    """br_set_cosine_option: process-wide kernel-variant / schedule switches of the brute-force path (tests, sweeps)."""
    check(_lib.load().br_set_cosine_option(name.encode(), int(value)), "br_set_cosine_option")


class CosineIndex:
    """Doc embeddings resident in HBM with their inverse norms (``1/(||e||+1e-10)``, computed once)."""

    def __init__(self, doc_embeddings, device=None, doc_base=0):
        lib = _lib.load()
        dev = _lib.require_cuda(device)
        self.device, self.doc_base = dev, int(doc_base)
        self.docs = _bf16(doc_embeddings, dev)
        if self.docs.dim() != 2 or self.docs.shape[1] % 8:
            raise ValueError("embeddings must be [n, d] with d a multiple of 8")
        self.n_docs, self.dim = self.docs.shape
        with torch.cuda.device(dev):
            self.inv_norm = torch.empty(self.n_docs, dtype=torch.float32, device=dev)
            check(lib.br_row_inv_norms(ptr(self.docs), self.n_docs, self.dim, ptr(self.inv_norm), _lib.stream_ptr(dev)),
                  "br_row_inv_norms")

    def topk(self, query_embeddings, k=10):
        """Brute-force cosine top-k -> (ids int64[Q, k] (doc_base + row), sims float32[Q, k]) on the device."""
        lib = _lib.load()
        q = _bf16(query_embeddings, self.device)
        nq = q.shape[0]
        with torch.cuda.device(self.device):
            ids = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            sims = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            check(lib.br_cosine_topk(ptr(self.docs), ptr(self.inv_norm), self.n_docs, self.dim, ptr(q), nq, int(k),
                                     self.doc_base, ptr(ids), ptr(sims), _lib.stream_ptr(self.device)), "br_cosine_topk")
        return ids, sims

    def rerank(self, query_embeddings, candidate_ids, k=10):
        """Cosine re-rank of per-query candidates [Q, c] (local rows, -1 = empty) -> (ids int32[Q, k], sims)."""
        lib = _lib.load()
        q = _bf16(query_embeddings, self.device)
        cand = torch.as_tensor(candidate_ids).to(device=self.device, dtype=torch.int32).contiguous()
        nq, c = cand.shape
        with torch.cuda.device(self.device):
            ids = torch.empty((nq, k), dtype=torch.int32, device=self.device)
            sims = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            check(lib.br_cosine_rerank(ptr(self.docs), ptr(self.inv_norm), self.n_docs, self.dim, ptr(q), nq, ptr(cand), c,
                                       int(k), ptr(ids), ptr(sims), _lib.stream_ptr(self.device)), "br_cosine_rerank")
        return ids, sims


def rerank_bm25_with_cosine(bm25_model, cosine_index, queries, query_embeddings, n_candidates=1000, k=10):
    """BASELINE config 3: BM25 top-``n_candidates`` then cosine re-rank to top-``k`` (the intended
    BM25 -> embedding re-rank of the reference, README.md:93-94; candidate cap 1000,
    text_preprocessing_and_embedding_setup.py:342)."""
    cand, _ = bm25_model.retrieve_top_n_batch(queries, min(n_candidates, bm25_model.corpus_size))
    return cosine_index.rerank(query_embeddings, cand, k)
