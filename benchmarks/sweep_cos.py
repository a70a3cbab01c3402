"""A/B sweep of the cosine GEMM variants inside ONE process (variants interleaved round-robin, so clock / power drift
hits all of them alike).  Usage: python benchmarks/sweep_cos.py "BR_COS_QS_WINDOW=64" "BR_COS_KERNEL=mc" ..."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from document_retrieval_b200.cosine import CosineIndex  # noqa: E402

KEYS = ["BR_COS_KERNEL", "BR_COS_QS_BN", "BR_COS_QS_WINDOW", "BR_COS_DEBUG_NOEPI", "BR_COS_CHUNK0", "BR_COS_CHUNK_MULT", "BR_COS_TIGHTEN_T"]


def main():
    variants = sys.argv[1:] or [""]
    rounds = int(os.environ.get("SWEEP_ROUNDS", "5"))
    n_docs = int(os.environ.get("SWEEP_DOCS", "1250000"))
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(20241105 + 5)
    docs = torch.randn(n_docs, 768, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    qs = torch.randn(10_000, 768, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    ix = CosineIndex(docs)
    times = {v: [] for v in variants}
    ref = None
    for r in range(rounds + 1):
        for v in variants:
            for k in KEYS:
                os.environ.pop(k, None)
            for kv in v.split():
                k, x = kv.split("=")
                os.environ[k] = x
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ids, sims = ix.topk(qs, 10)
            e1.record()
            torch.cuda.synchronize()
            if r > 0:
                times[v].append(e0.elapsed_time(e1))
            if "NOEPI" not in v:
                if ref is None:
                    ref = ids.clone()
                assert torch.equal(ids, ref), f"variant {v!r} changed the result"
    flop = 2.0 * n_docs * 10_000 * 768
    for v in variants:
        t = np.asarray(times[v])
        print(f"{v or '(default)':60s} median {np.median(t):7.3f} ms  min {t.min():7.3f}  max {t.max():7.3f}   {flop / np.median(t) / 1e9:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
