#!/usr/bin/env python
"""Wall-clock split of the C4 index build: br_index_build (keys, sort, run-length encode, df) vs br_index_finalize
(idf on the host, weights + upper bounds, skip tables, rows) - development aid."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from document_retrieval_b200 import BM25, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_800_000
do, tk = synth.make_corpus_torch(n, 1_000_000, 60, "cuda", seed=5)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.time()
    m = BM25.from_token_ids(do, tk, 1_000_000, finalize=False)
    torch.cuda.synchronize()
    t1 = time.time()
    m.finalize()
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"rep {rep}: build {1e3 * (t1 - t0):.1f} ms  finalize {1e3 * (t2 - t1):.1f} ms  tokens {tk.numel()} postings {m.stats()['nnz']}")
    del m
