"""Text ingestion throughput: C4-shaped synthetic text (docs of ~60 tokens "t{rank}",
Zipf vocabulary) -> term ids + vocabulary on the GPU, against the reference's own
``[text.split() for text in texts]`` + dict-insert vocabulary on one host core (sample).

    python benchmarks/bench_ingest.py [--docs 2000000] [--vocab 1000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from document_retrieval_b200 import ingest, synth  # noqa: E402


def synth_text(n_docs, vocab, mu, seed=5):
    """-> (uint8 text, int64 byte offsets, int32 token ranks, int64 token offsets); tokens separated by one space."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.rint(rng.normal(mu, mu / 4, size=n_docs)), 4, 4 * mu).astype(np.int64)
    tok_off = np.zeros(n_docs + 1, np.int64)
    np.cumsum(lens, out=tok_off[1:])
    T = int(tok_off[-1])
    cdf = np.cumsum(1.0 / np.arange(1, vocab + 1))
    cdf /= cdf[-1]
    table = np.zeros((vocab, 8), np.uint8)
    tl = np.zeros(vocab, np.int64)
    for r in range(vocab):                       # "t123 " padded to 8 bytes
        s = b"t%d" % r
        table[r, :len(s)] = np.frombuffer(s, np.uint8)
        tl[r] = len(s)
    chunks, ranks = [], np.empty(T, np.int32)
    byte_len = np.empty(T, np.int64)
    step = 8_000_000
    for a in range(0, T, step):
        r = np.searchsorted(cdf, rng.random(min(step, T - a))).astype(np.int32)
        ranks[a:a + r.size] = r
        L = tl[r] + 1
        byte_len[a:a + r.size] = L
        rows = table[r]
        rows[np.arange(r.size), tl[r]] = 0x20
        mask = np.arange(8)[None, :] < L[:, None]
        chunks.append(rows[mask])
    text = np.concatenate(chunks)
    bo = np.zeros(T + 1, np.int64)
    np.cumsum(byte_len, out=bo[1:])
    return text, bo[tok_off], ranks, tok_off


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=2_000_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--mu", type=int, default=60)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-docs", type=int, default=100_000)
    a = ap.parse_args()
    text, boff, ranks, tok_off = synth_text(a.docs, a.vocab, a.mu)
    T = int(tok_off[-1])
    dev = torch.device("cuda", 0)
    t_text = torch.from_numpy(text).pin_memory()
    t_off = torch.from_numpy(boff).pin_memory()
    best = None
    for it in range(a.steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        voc, d_off, d_ids = ingest.Vocabulary.from_texts((t_text.numpy(), t_off.numpy()), device=dev)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it > 0:
            best = dt if best is None else min(best, dt)
        if it < a.steps:
            del voc, d_off, d_ids
    # check against the generator: same token boundaries, and ids = first-seen relabelling of the ranks
    assert np.array_equal(d_off.cpu().numpy(), tok_off)
    ids = d_ids.cpu().numpy()
    first = np.full(a.vocab, -1, np.int64)
    uniq, idx = np.unique(ranks, return_index=True)
    order = np.argsort(idx, kind="stable")
    first[uniq[order]] = np.arange(uniq.size)
    assert np.array_equal(ids, first[ranks].astype(np.int32))
    assert len(voc) == uniq.size
    # CPU: the reference's own step on one core, first --cpu-docs docs
    n_cpu = min(a.cpu_docs, a.docs)
    raw = text.tobytes()
    texts = [raw[boff[i]:boff[i + 1]].decode() for i in range(n_cpu)]
    t0 = time.perf_counter()
    toks = [t.split() for t in texts]                       # bm25_ranking.ipynb:299
    vocab = {}
    for d in toks:                                          # vocabulary growth of BM25.build, :180-186
        for w in d:
            if w not in vocab:
                vocab[w] = len(vocab)
    cpu_dt = time.perf_counter() - t0
    cpu_tok = int(tok_off[n_cpu])
    print(json.dumps({
        "metric": "text ingestion tokens/sec (tokenise + vocabulary + term ids)", "value": T / best, "unit": "tokens/s",
        "seconds": best, "config": {"workload": f"{a.docs} docs x ~{a.mu} tok, {a.vocab}-term Zipf vocab, {text.size} bytes of text, "
                                                f"{uniq.size} distinct terms", "includes": "H2D copy of the text from pinned memory"},
        "text_gbs": text.size / best / 1e9,
        "cpu_baseline": {"value": cpu_tok / cpu_dt, "unit": "tokens/s", "cores": 1, "kind": "reference",
                         "sample": f"first {n_cpu} docs: str.split() + dict inserts ({cpu_dt:.2f} s)"},
        "check": "token offsets and term ids identical to the generator's first-seen relabelling"}))


if __name__ == "__main__":
    main()
