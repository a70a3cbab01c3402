#!/usr/bin/env python
"""In-process A/B of br_cosine_topk's launch schedule / kernel variants on the config-5 shard (1.25 M x 768, 10k queries):
python benchmarks/sweep_cosine_schedule.py "chunk0=4" "chunk0=8 chunk_mult=4" ...  (options of br_set_cosine_option)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from document_retrieval_b200.cosine import CosineIndex, set_cosine_option  # noqa: E402

DEFAULTS = {"kernel": 0, "qs_bn": 224, "qs_window": 64, "chunk0": 1, "chunk_mult": 2, "tighten_threads": 64, "qs_epi": 1}


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(20241105 + 5)
    docs = torch.empty(1_250_000, 768, device=dev, dtype=torch.bfloat16)
    for a in range(0, docs.shape[0], 1 << 18):
        b = min(docs.shape[0], a + (1 << 18))
        docs[a:b] = torch.randn(b - a, 768, generator=g, device=dev).to(torch.bfloat16)
    qs = torch.randn(10_000, 768, generator=g, device=dev).to(torch.bfloat16)
    ix = CosineIndex(docs)
    ref = None
    for spec in [""] + sys.argv[1:]:
        for k, v in DEFAULTS.items():
            set_cosine_option(k, v)
        for kv in spec.split():
            k, v = kv.split("=")
            set_cosine_option(k, int(v))
        for _ in range(3):
            ids, sims = ix.topk(qs, 10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ids, sims = ix.topk(qs, 10)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if ref is None:
            ref = (ids.clone(), sims.clone())
        same = bool(torch.equal(ids, ref[0]) and torch.equal(sims, ref[1]))
        print(f"{spec or 'default':40s} {ms:7.3f} ms  {2 * 1.25e6 * 1e4 * 768 / ms / 1e9:7.1f} TFLOP/s  identical={same}", flush=True)


if __name__ == "__main__":
    main()
