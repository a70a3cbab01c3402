#!/usr/bin/env python
"""CPU, authoring container only (needs /root/reference): how long the UNMODIFIED reference BM25 class
(bm25_ranking.ipynb:166-213) takes to joblib.dump / joblib.load (:312, :222-251) at config-1 scale, next to its build
time.  The number goes into BASELINE.md beside bench.py's `secondary.index_file` figure."""
import json
import os
import sys
import tempfile
import time

import joblib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from document_retrieval_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402


def main(scale=1.0):
    c = synth.make_config("C1", scale=scale)
    docs = synth.to_strings(c["doc_offsets"], c["token_ids"])
    BM25 = ref_loader.notebook_bm25_class()
    BM25.__module__ = "__main__"                     # the exec'd class must be importable for pickle, like in the notebook
    sys.modules["__main__"].BM25 = BM25
    t0 = time.time()
    m = BM25(docs)
    build_s = time.time() - t0
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "bm25_model_en.joblib")
        t0 = time.time()
        joblib.dump(m, p)
        dump_s = time.time() - t0
        size = os.path.getsize(p)
        t0 = time.time()
        m2 = joblib.load(p)
        load_s = time.time() - t0
    assert m2.corpus_size == m.corpus_size
    n_tok = int(c["doc_offsets"][-1])
    print(json.dumps({"docs": c["n_docs"], "tokens": n_tok, "postings": int(sum(len(d) for d in m.term_freqs)),
                      "build_s": build_s, "joblib_dump_s": dump_s, "joblib_load_s": load_s, "file_bytes": size,
                      "load_postings_per_s": sum(len(d) for d in m.term_freqs) / load_s}))


if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
