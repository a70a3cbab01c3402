#!/usr/bin/env python
"""BASELINE config 3 end to end: BM25 top-1000 candidates (C2 'en' corpus, 207,363 docs) then cosine re-rank with
768-d bf16 embeddings, 10k queries.  Prints one JSON line with the stage times."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from document_retrieval_b200 import BM25, synth  # noqa: E402
from document_retrieval_b200.cosine import CosineIndex  # noqa: E402


def main():
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
    dev = torch.device("cuda", 0)
    n, vocab = synth.C2_LANG_DOCS["en"], synth.C2_VOCAB["en"]
    do, tk = synth.make_corpus_torch(n, vocab, synth.C2_MEAN_LEN, dev, seed=synth.ROOT_SEED + 2)
    qo, qt, _ = synth.make_queries_torch(do, tk, nq, vocab, seed=synth.ROOT_SEED + 3)
    m = BM25.from_token_ids(do, tk, vocab)
    g = torch.Generator(device=dev).manual_seed(synth.ROOT_SEED + 3)
    emb = torch.randn(n, 768, generator=g, device=dev).to(torch.bfloat16)
    qe = torch.randn(nq, 768, generator=g, device=dev).to(torch.bfloat16)
    ix = CosineIndex(emb)
    d_t, d_o = torch.from_numpy(qt).to(dev), torch.from_numpy(qo).to(dev)

    def run():
        torch.cuda.synchronize()
        t0 = time.time()
        cand, _ = m.retrieve_top_n_batch((d_t, d_o), 1000)
        torch.cuda.synchronize()
        t1 = time.time()
        ids, sims = ix.rerank(qe, cand, 10)
        torch.cuda.synchronize()
        return t1 - t0, time.time() - t1, m.query_stats()

    run()
    best = min((run() for _ in range(3)), key=lambda r: r[0] + r[1])
    print(json.dumps({"metric": "C3 BM25 top-1000 -> cosine re-rank queries/sec", "value": nq / (best[0] + best[1]), "unit": "queries/s",
                      "bm25_top1000_ms": best[0] * 1e3, "cosine_rerank_ms": best[1] * 1e3,
                      "config": {"workload": f"{n} docs x ~200 tok (C2 en), {nq} queries, k=1000 -> 10, D=768 bf16"},
                      "bm25_path": {"fused": best[2]["queries_fused"], "dense": best[2]["queries_dense"]}}))


if __name__ == "__main__":
    main()
