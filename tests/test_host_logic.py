"""CPU: host-side helpers that need no GPU (MRR/Recall, compute_idf, the C header)."""
import numpy as np

from oracle import bm25_oracle as orc
from document_retrieval_b200 import compute_idf, mrr_recall_at_k


def test_mrr_recall_matches_oracle():
    ranked = [[3, 1, 2], [9, 8, 7], [5]]
    rel = [1, [7, 6], 4]
    out = mrr_recall_at_k(ranked, rel, (1, 2, 3))
    for k in (1, 2, 3):
        m = [orc.mrr_recall_at_k(r, [x] if isinstance(x, int) else x, k) for r, x in zip(ranked, rel)]
        assert abs(out[k][0] - np.mean([a for a, _ in m])) < 1e-15 and abs(out[k][1] - np.mean([b for _, b in m])) < 1e-15


def test_compute_idf_is_the_reference_expression():
    df = {"a": 1, "b": 50, "c": 99}
    idf = compute_idf(df, 100)
    for t, d in df.items():
        assert idf[t] == float(np.log((100 - d + 0.5) / (d + 0.5)))     # cosine_similarity_bm25_reranking.py:179
    assert idf["c"] < 0


def test_pack_texts_arrow_layout():
    from document_retrieval_b200.ingest import pack_texts
    texts = ["a b", "", None, "été  x", 3.5, "한국어"]
    data, off = pack_texts(texts)
    clean = [t if isinstance(t, str) else "" for t in texts]        # bm25_ranking.ipynb:85-86
    enc = [t.encode("utf-8") for t in clean]
    assert data.tobytes() == b"".join(enc)
    assert off.tolist() == np.concatenate([[0], np.cumsum([len(e) for e in enc])]).tolist()
    d0, o0 = pack_texts([])
    assert d0.size == 0 and o0.tolist() == [0]


def test_c_abi_header_is_plain_c(tmp_path):
    """include/br_b200.h must be consumable by a C compiler (the boundary is a C ABI: no C++ or torch types): a C99
    translation unit that includes it and takes the address of every declared entry point compiles with -Wall -Werror."""
    import os
    import re
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    hdr = open(os.path.join(root, "include", "br_b200.h")).read()
    names = sorted(set(re.findall(r"\b(br_[a-z0-9_]+)\s*\(", hdr)))
    src = tmp_path / "abi.c"
    src.write_text('#include "br_b200.h"\n#include <stddef.h>\nconst void* const entry_points[] = {\n'
                   + "".join(f"    (const void*)&{n},\n" for n in names) + "};\n"
                   "int n_entry_points(void) { return (int)(sizeof(entry_points) / sizeof(entry_points[0])); }\n")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-Wno-pedantic", "-I", os.path.join(root, "include"),
                           "-c", str(src), "-o", str(tmp_path / "abi.o")])


class _StandInModel:
    """Oracle-backed stand-in with the attributes the routing layer touches (test infrastructure; the product's
    model is CUDA-only)."""

    def __init__(self, docs):
        off = np.zeros(len(docs) + 1, np.int64)
        np.cumsum([len(d) for d in docs], out=off[1:])
        self.words = sorted({w for d in docs for w in d})
        self.w2i = {w: i for i, w in enumerate(self.words)}
        tok = np.asarray([self.w2i[w] for d in docs for w in d], np.int32)
        self.ix = orc.build_index(off, tok, len(self.words))
        self.corpus_size = len(docs)

    def retrieve_top_n_batch(self, queries, n):
        import torch
        ids = np.full((len(queries), n), -1, np.int32)
        for i, q in enumerate(queries):
            a, _ = orc.retrieve_top_n(self.ix, [self.w2i.get(w, -1) for w in q], n)
            ids[i, :a.size] = a
        return torch.from_numpy(ids), None


def test_recall_and_test_retrieval_follow_the_notebook_semantics():
    """bm25_ranking.ipynb:329-354 / :368-389 on the CPU: unknown languages are skipped but stay in the Recall denominator
    (:331,353) and give [] in retrieve_test_queries (:374-376); k larger than a corpus is clipped."""
    from document_retrieval_b200 import evaluate_recall_at_k, retrieve_test_queries
    docs = {"en": [["apple", "pie"], ["banana", "split"], ["apple", "banana", "smoothie"]], "fr": [["tarte", "pomme"], ["banane"]]}
    models = {lang: _StandInModel(d) for lang, d in docs.items()}
    id_maps = {lang: [f"{lang}{i}" for i in range(len(d))] for lang, d in docs.items()}
    rows = [{"query": "apple pie", "lang": "en", "positive_docs": "en0"},
            {"query": ["banane"], "lang": "fr", "positive_docs": "fr1"},
            {"query": "smoothie", "lang": "en", "positive_docs": "en1"},          # a miss at k = 1
            {"query": "whatever", "lang": "ko", "positive_docs": "ko0"}]         # unknown language
    assert evaluate_recall_at_k(models, id_maps, rows, k=1) == 2 / 4
    assert evaluate_recall_at_k(models, id_maps, rows, k=10) == 3 / 4
    out = retrieve_test_queries(models, id_maps, rows, k=10)
    assert out[3] == [] and out[0][0] == "en0" and out[1][0] == "fr1" and len(out[1]) == 2 and len(out[0]) == 3
    assert evaluate_recall_at_k(models, id_maps, [], k=10) == 0


def test_index_file_header_and_packed_query_forms(tmp_path):
    """Host-side logic that needs no GPU: the index-file reader rejects foreign / corrupt files before touching the
    device, and pack_queries tells the packed (q_terms, q_offsets) ARRAY form from a batch of two token lists."""
    import struct
    import pytest
    import torch
    from document_retrieval_b200 import BM25, indexfile
    from document_retrieval_b200._lib import BRError
    p = tmp_path / "x.brix"
    p.write_bytes(b"not an index")
    with pytest.raises(BRError):
        indexfile.read_header(str(p))
    p.write_bytes(indexfile.MAGIC + struct.pack("<Q", 1 << 40))
    with pytest.raises(BRError):
        indexfile.read_header(str(p))
    hdr = b'{"format": 1, "arrays": []}'
    p.write_bytes(indexfile.MAGIC + struct.pack("<Q", len(hdr)) + hdr)
    assert indexfile.read_header(str(p))["format"] == 1
    m = BM25.__new__(BM25)
    m._vocab, m._terms, m._term_pool, m.vocabulary = None, None, None, None
    qt, qo = m.pack_queries((np.array([3, 4, 5], np.int32), np.array([0, 1, 3], np.int32)))
    assert qt.tolist() == [3, 4, 5] and qo.tolist() == [0, 1, 3]
    qt, qo = m.pack_queries(([3, 4, 5], [0, 1, 3]))                 # two queries given as lists: NOT the packed form
    assert qt.tolist() == [3, 4, 5, 0, 1, 3] and qo.tolist() == [0, 3, 6]
    with pytest.raises(ValueError):
        m.pack_queries((np.array([3, 4, 5], np.int32), np.array([0, 1, 2], np.int32)))   # offsets do not end at len(q_terms)
    assert torch.is_tensor(qt)


def test_per_language_recall_matches_the_oracle():
    from document_retrieval_b200 import per_language_recall
    ranked = {10: ["a", "b"], 11: ["c"], 12: [], 13: ["z"]}
    pos = ["b", "x", "q", "z"]
    langs = {10: "en", 11: "en", 12: "fr", 13: "ko"}
    got = per_language_recall(ranked, pos, langs)
    want = orc.per_language_recall([ranked[k] for k in ranked], pos, [langs[k] for k in ranked])
    assert got == want == (0.5, {"en": 0.5, "fr": 0.0, "ko": 1.0})


def test_ctypes_binding_matches_the_header():
    """_lib.SIGNATURES mirrors include/br_b200.h one to one: same entry points, same number of arguments (an ABI drift
    like a callback or an argument added on one side only would otherwise surface as a crash on the GPU box), and the
    built library exports every one of them (no compute call: loading and symbol lookup only)."""
    import ctypes as C
    import os
    import re
    from document_retrieval_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "br_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)                     # comments mention entry points too
    decl = {}
    for m in re.finditer(r"\b(br_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        if "(*" in m.group(0).split(name)[0][-12:]:                        # a function-pointer typedef, not an entry point
            continue
        decl[name] = 0 if args in ("", "void") else args.count(",") + 1
    typedefs = set(re.findall(r"typedef[^;]*\(\s*\*\s*(br_[a-z0-9_]+)\s*\)", hdr))
    for t in typedefs:
        decl.pop(t, None)
    assert set(decl) == set(_lib.SIGNATURES), (sorted(set(decl) ^ set(_lib.SIGNATURES)))
    for name, n_args in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == n_args, (name, n_args, len(_lib.SIGNATURES[name][1]))
    if os.path.isfile(_lib.SO_PATH):
        lib = C.CDLL(_lib.SO_PATH)
        for name in decl:
            assert hasattr(lib, name), name
