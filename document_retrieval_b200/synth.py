"""synth-v1: deterministic synthetic corpora / queries for the BM25 hot path.

The reference ships no dataset (`/data/` is git-ignored) so every parity and
throughput number in this repo is taken on inputs made here.  Shapes follow
BASELINE.json's configs:

* docs   : term rank r in [0, V) with p(r) ~ 1/(r+1) (Zipf s=1), tokens i.i.d.;
           length L = clip(round(N(mu, (mu/4)^2)), 4, 4*mu)
* queries: pick a source doc uniformly, m ~ U{5..15}, take m tokens of that doc
           without replacement (all of them when the doc is shorter) - duplicates
           are possible, which exercises ``set(query)`` de-duplication
           (bm25_ranking.ipynb:193) against the duplicate-counting path
           (team_run1.py:183); every 100th query gets one out-of-vocabulary
           token; the qrel is the source doc (one positive per query, like the
           reference's ``positive_docs`` column).

Two generators produce the same distribution: a numpy one (PCG64, seeds spawned
from SeedSequence(20241105)) for tests / small configs, and a torch one that
runs on the GPU for the 8.8M-doc config where 528M Zipf draws on the host would
take minutes.
"""
from __future__ import annotations

import numpy as np

ROOT_SEED = 20241105

# BASELINE.json configs (docs, vocab, mean doc length, queries)
CONFIGS = {
    "C1": dict(n_docs=10_000, vocab=30_000, mean_len=200, n_queries=1_000),
    "C4": dict(n_docs=8_800_000, vocab=1_000_000, mean_len=60, n_queries=10_000),
}
# per-language corpus sizes, final_implementation.py:310-318
C2_LANG_DOCS = {"ar": 8_829, "de": 10_992, "en": 207_363, "es": 11_019,
                "fr": 10_676, "it": 11_250, "ko": 7_893}
C2_VOCAB = {"en": 200_000}
C2_VOCAB_DEFAULT = 100_000
C2_MEAN_LEN = 200
C2_QUERIES = 2_000


def _rng(*key: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence([ROOT_SEED, *key])))


def zipf_cdf(vocab: int) -> np.ndarray:
    p = 1.0 / (np.arange(vocab, dtype=np.float64) + 1.0)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return cdf


def doc_lengths(n_docs: int, mean_len: int, rng: np.random.Generator) -> np.ndarray:
    L = np.rint(rng.normal(mean_len, mean_len / 4.0, size=n_docs))
    return np.clip(L, 4, 4 * mean_len).astype(np.int64)


def make_corpus(n_docs: int, vocab: int, mean_len: int, key=(0,)):
    """-> (doc_offsets int64[N+1], token_ids int32[T])"""
    rng = _rng(1, *key)
    L = doc_lengths(n_docs, mean_len, rng)
    doc_offsets = np.zeros(n_docs + 1, dtype=np.int64)
    np.cumsum(L, out=doc_offsets[1:])
    total = int(doc_offsets[-1])
    cdf = zipf_cdf(vocab)
    token_ids = np.empty(total, dtype=np.int32)
    step = 1 << 24
    for s in range(0, total, step):
        u = rng.random(min(step, total - s))
        token_ids[s:s + u.size] = np.minimum(np.searchsorted(cdf, u, side="right"), vocab - 1)
    return doc_offsets, token_ids


def make_queries(doc_offsets: np.ndarray, token_ids: np.ndarray, n_queries: int, vocab: int,
                 key=(0,), oov_every: int = 100):
    """-> (q_offsets int32[Q+1], q_terms int32[sum m], qrels int64[Q]).

    An OOV token is encoded as term id ``vocab`` (outside [0, vocab))."""
    rng = _rng(2, *key)
    n_docs = doc_offsets.size - 1
    src = rng.integers(0, n_docs, size=n_queries)
    m = rng.integers(5, 16, size=n_queries)
    terms, offs = [], [0]
    for qi in range(n_queries):
        lo, hi = int(doc_offsets[src[qi]]), int(doc_offsets[src[qi] + 1])
        take = min(int(m[qi]), hi - lo)
        pos = rng.choice(hi - lo, size=take, replace=False)
        t = token_ids[lo + pos].astype(np.int32)
        if oov_every and qi % oov_every == oov_every - 1:
            t = np.concatenate([t, np.array([vocab], dtype=np.int32)])
        terms.append(t)
        offs.append(offs[-1] + t.size)
    return (np.asarray(offs, dtype=np.int32), np.concatenate(terms).astype(np.int32),
            src.astype(np.int64))


def to_strings(doc_offsets, token_ids, prefix: str = "t"):
    """list[list[str]] form for the reference's string API (small configs only)."""
    toks = [f"{prefix}{int(t)}" for t in token_ids]
    return [toks[int(doc_offsets[i]):int(doc_offsets[i + 1])] for i in range(doc_offsets.size - 1)]


def queries_to_strings(q_offsets, q_terms, vocab: int, prefix: str = "t"):
    out = []
    for i in range(q_offsets.size - 1):
        seg = q_terms[int(q_offsets[i]):int(q_offsets[i + 1])]
        out.append([f"{prefix}{int(t)}" if t < vocab else f"oov{prefix}{i}" for t in seg])
    return out


def make_config(name: str, scale: float = 1.0):
    """Corpus + queries for C1 / C4 (optionally scaled down by ``scale`` in docs and queries)."""
    c = CONFIGS[name]
    n_docs = max(16, int(c["n_docs"] * scale))
    n_q = max(4, int(c["n_queries"] * scale)) if scale < 1.0 else c["n_queries"]
    key = (int(name[1:]),)
    doc_offsets, token_ids = make_corpus(n_docs, c["vocab"], c["mean_len"], key)
    q_offsets, q_terms, qrels = make_queries(doc_offsets, token_ids, n_q, c["vocab"], key)
    return dict(n_docs=n_docs, vocab=c["vocab"], doc_offsets=doc_offsets, token_ids=token_ids,
                q_offsets=q_offsets, q_terms=q_terms, qrels=qrels)


def make_c2(scale: float = 1.0):
    """Seven per-language corpora (sizes from final_implementation.py:310-318) and a mixed
    query stream with the language drawn uniformly."""
    langs = {}
    for li, (lang, n) in enumerate(C2_LANG_DOCS.items()):
        n_docs = max(16, int(n * scale))
        vocab = C2_VOCAB.get(lang, C2_VOCAB_DEFAULT)
        doc_offsets, token_ids = make_corpus(n_docs, vocab, C2_MEAN_LEN, (2, li))
        langs[lang] = dict(n_docs=n_docs, vocab=vocab, doc_offsets=doc_offsets, token_ids=token_ids)
    n_q = max(14, int(C2_QUERIES * scale))
    rng = _rng(3, 2)
    names = list(C2_LANG_DOCS)
    q_lang = [names[i] for i in rng.integers(0, len(names), size=n_q)]
    queries = []
    for qi, lang in enumerate(q_lang):
        c = langs[lang]
        qo, qt, rel = make_queries(c["doc_offsets"], c["token_ids"], 1, c["vocab"], (2, 1000 + qi),
                                   oov_every=0)
        queries.append(dict(lang=lang, terms=qt, qrel=int(rel[0])))
    return langs, queries


# ----------------------------------------------------------------------------------------------
# torch generator (device-side) for the 8.8M-doc config
# ----------------------------------------------------------------------------------------------
def make_corpus_torch(n_docs: int, vocab: int, mean_len: int, device, seed: int = ROOT_SEED):
    """Same distribution as make_corpus, drawn with torch on ``device``.
    -> (doc_offsets int64[N+1], token_ids int32[T]) as torch tensors on ``device``."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    L = torch.randn(n_docs, generator=g, device=device, dtype=torch.float32) * (mean_len / 4.0) + mean_len
    L = L.round().clamp_(4, 4 * mean_len).to(torch.int64)
    doc_offsets = torch.zeros(n_docs + 1, dtype=torch.int64, device=device)
    torch.cumsum(L, 0, out=doc_offsets[1:])
    total = int(doc_offsets[-1].item())
    cdf = torch.from_numpy(zipf_cdf(vocab)).to(device)
    token_ids = torch.empty(total, dtype=torch.int32, device=device)
    step = 1 << 26
    for s in range(0, total, step):
        n = min(step, total - s)
        u = torch.rand(n, generator=g, device=device, dtype=torch.float64)
        token_ids[s:s + n] = torch.searchsorted(cdf, u, right=True).clamp_(max=vocab - 1).to(torch.int32)
    return doc_offsets, token_ids


def make_queries_torch(doc_offsets, token_ids, n_queries: int, vocab: int, seed: int = ROOT_SEED + 1,
                       oov_every: int = 100):
    """Queries drawn from device-resident docs; only the few source docs are copied to the host.
    -> numpy (q_offsets int32[Q+1], q_terms int32, qrels int64[Q])"""
    import torch
    rng = np.random.Generator(np.random.PCG64(seed))
    n_docs = doc_offsets.numel() - 1
    src = rng.integers(0, n_docs, size=n_queries)
    m = rng.integers(5, 16, size=n_queries)
    src_t = torch.from_numpy(src).to(doc_offsets.device)
    lo = doc_offsets[src_t].cpu().numpy()
    hi = doc_offsets[src_t + 1].cpu().numpy()
    max_len = int((hi - lo).max())
    idx = torch.from_numpy(lo).to(doc_offsets.device)[:, None] + torch.arange(max_len, device=doc_offsets.device)[None, :]
    idx.clamp_(max=token_ids.numel() - 1)
    toks = token_ids[idx].cpu().numpy()
    terms, offs = [], [0]
    for qi in range(n_queries):
        ln = int(hi[qi] - lo[qi])
        take = min(int(m[qi]), ln)
        pos = rng.choice(ln, size=take, replace=False)
        t = toks[qi, pos].astype(np.int32)
        if oov_every and qi % oov_every == oov_every - 1:
            t = np.concatenate([t, np.array([vocab], dtype=np.int32)])
        terms.append(t)
        offs.append(offs[-1] + t.size)
    return (np.asarray(offs, dtype=np.int32), np.concatenate(terms).astype(np.int32),
            src.astype(np.int64))
