"""CPU simulation (numpy, analysis only): how much work does exact MaxScore-style pruning leave on the synth-v1
C4 workload?  Lists are processed rarest first; a doc is fully scored the first time it is seen (lookups into the
more frequent lists with early exit against the running k-th best score theta); processing stops as soon as the
summed upper bounds of the unprocessed lists fall below theta.  Counts postings read and lookups done."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from document_retrieval_b200 import synth  # noqa: E402


def main(n_docs=1_100_000, vocab=1_000_000, mu=60, nq=300, k=10, chunk=2048):
    t0 = time.time()
    doc_off, tok = synth.make_corpus(n_docs, vocab, mu)
    q_off, q_terms, _ = synth.make_queries(doc_off, tok, nq, vocab)
    dl = np.diff(doc_off)
    doc_of_tok = np.repeat(np.arange(n_docs, dtype=np.int64), dl)
    keys = tok.astype(np.int64) * n_docs + doc_of_tok
    uk, tf = np.unique(keys, return_counts=True)
    term = (uk // n_docs).astype(np.int32)
    doc = (uk % n_docs).astype(np.int32)
    df = np.bincount(term, minlength=vocab)
    row = np.zeros(vocab + 1, np.int64)
    np.cumsum(df, out=row[1:])
    avgdl = dl.mean()
    idf = np.log(1 + (n_docs - df + 0.5) / (df + 0.5))
    w = idf[term] * (tf * 2.5) / (tf + 1.5 * (1 - 0.75 + dl[doc] / avgdl))
    tmax = np.zeros(vocab)
    np.maximum.at(tmax, term, w)
    print(f"built {n_docs} docs, {uk.size} postings in {time.time() - t0:.1f}s", flush=True)
    tot_df, tot_read, tot_look, stops, maxlist = [], [], [], [], []
    for qi in range(nq):
        ts = np.unique(q_terms[q_off[qi]:q_off[qi + 1]])
        ts = ts[(ts < vocab)]
        ts = ts[df[ts] > 0]
        ts = ts[np.argsort(df[ts], kind="stable")]
        m = ts.size
        ub = tmax[ts]
        suffix = np.concatenate([np.cumsum(ub[::-1])[::-1], [0.0]])      # suffix[j] = sum ub[j:]
        dense = {}
        for j, t in enumerate(ts):
            v = np.zeros(n_docs, np.float32)
            v[doc[row[t]:row[t + 1]]] = w[row[t]:row[t + 1]]
            dense[j] = v
        seen = np.zeros(n_docs, bool)
        top = np.zeros(0)
        theta = 0.0
        read = look = 0
        stop = m
        for i in range(m):
            if i > 0 and suffix[i] < theta:
                stop = i
                break
            t = ts[i]
            d_all = doc[row[t]:row[t + 1]]
            w_all = w[row[t]:row[t + 1]]
            for a in range(0, d_all.size, chunk):
                d = d_all[a:a + chunk]
                s = w_all[a:a + chunk].astype(np.float64).copy()
                new = ~seen[d]
                seen[d] = True
                read += d.size
                alive = new.copy()
                for j in range(i + 1, m):
                    alive &= (s + suffix[j] >= theta)
                    look += int(alive.sum())
                    s[alive] += dense[j][d[alive]]
                alive &= s >= theta
                top = np.sort(np.concatenate([top, s[alive]]))[::-1][:k]
                if top.size >= k:
                    theta = top[k - 1]
        tot_df.append(int(df[ts].sum()))
        tot_read.append(read)
        tot_look.append(look)
        stops.append((stop, m))
        maxlist.append(int(df[ts[stop - 1]]) if stop > 0 else 0)
    tot_df, tot_read, tot_look, maxlist = map(np.asarray, (tot_df, tot_read, tot_look, maxlist))
    print("sum df per query      mean %.0f" % tot_df.mean())
    print("postings read         mean %.0f  median %.0f  p90 %.0f  p99 %.0f  max %.0f" % (
        tot_read.mean(), np.median(tot_read), np.percentile(tot_read, 90), np.percentile(tot_read, 99), tot_read.max()))
    print("lookups               mean %.0f  median %.0f  p90 %.0f  p99 %.0f  max %.0f" % (
        tot_look.mean(), np.median(tot_look), np.percentile(tot_look, 90), np.percentile(tot_look, 99), tot_look.max()))
    print("largest list processed  median %.0f p90 %.0f p99 %.0f max %.0f" % (
        np.median(maxlist), np.percentile(maxlist, 90), np.percentile(maxlist, 99), maxlist.max()))
    print("queries that processed every list:", sum(1 for s, m in stops if s == m), "of", nq)
    print("read / sum df = %.4f ; lookups / sum df = %.4f" % (tot_read.sum() / tot_df.sum(), tot_look.sum() / tot_df.sum()))
    for thr in (20_000, 50_000, 100_000, 200_000):
        print(f"queries with read <= {thr}: {(tot_read <= thr).mean():.3f}")


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
