"""Small end-to-end pass over the newer kernels (text ingestion, vocabulary lookup, query-stationary cosine GEMM,
re-rank, tiled BM25) - sized for `compute-sanitizer --tool memcheck`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import document_retrieval_b200 as dr  # noqa: E402
from document_retrieval_b200.cosine import CosineIndex  # noqa: E402

rng = np.random.default_rng(0)
words = [f"w{i}" for i in range(300)] + ["été", "한국어", "a_b", "😀"]
texts = [" 　".join(rng.choice(words, size=int(rng.integers(0, 30)))) + (" " if i % 3 else "") for i in range(2000)]
m = dr.BM25.from_texts(texts, bigrams=True)
ids, sc = m.retrieve_top_n_texts(texts[:300] + ["zzz", ""], 10)
ref = dr.BM25([t.split() + (["_".join(g) for g in zip(t.split(), t.split()[1:])] if len(t.split()) >= 2 else []) for t in texts])
ids2, sc2 = ref.retrieve_top_n_batch([t.split() + (["_".join(g) for g in zip(t.split(), t.split()[1:])] if len(t.split()) >= 2 else [])
                                      for t in texts[:300]] + [["zzz"], []], 10)
assert torch.equal(ids, ids2) and torch.equal(sc, sc2)
g = torch.Generator().manual_seed(1)
docs = torch.randn(3000, 200, generator=g).to(torch.bfloat16)
qs = torch.randn(500, 200, generator=g).to(torch.bfloat16)
ix = CosineIndex(docs)
i1, s1 = ix.topk(qs, 10)
full = (qs.float() / (qs.float().norm(dim=1, keepdim=True) + 1e-10)) @ (docs.float() / (docs.float().norm(dim=1, keepdim=True) + 1e-10)).T
assert (torch.topk(full, 10).values.cuda() - s1).abs().max().item() < 1e-5
cand = torch.randint(0, 3000, (500, 64), dtype=torch.int32)
i2, s2 = ix.rerank(qs, cand, 5)
torch.cuda.synchronize()
print("sanity ok")
