"""Doc-sharded BM25 across the GPUs of one box (one process per GPU, torch.distributed).

The reference is single-process; sharding follows SURVEY 8e: contiguous doc ranges per rank, the
corpus statistics (N, sum of doc lengths, df histogram) all-reduced once at build time so that
every shard computes bit-identical idf / avgdl, and at query time only each rank's [Q, k]
(global id, float64 score) candidates cross NVLink (all-gather) before the k*G -> k merge.
Top-k over a doc partition is exactly decomposable, so sharded results equal single-GPU results
bit for bit.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr


def shard_bounds(n_docs: int, world: int):
    """Contiguous doc ranges: shard r = [r*ceil(N/G), min(N, (r+1)*ceil(N/G)))."""
    per = -(-n_docs // world)
    return [(min(n_docs, r * per), min(n_docs, (r + 1) * per)) for r in range(world)]


def reduce_stats(local_df: torch.Tensor, n_docs: int, sum_dl: int, group=None):
    """All-reduce (sum) of the df histogram and (N, sum dl) -> (df int64[V] on host, N, sum_dl)."""
    df = local_df.to(torch.int64)
    st = torch.tensor([n_docs, sum_dl], dtype=torch.int64, device=df.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(df, group=group)
        dist.all_reduce(st, group=group)
    st = st.cpu()
    return df.cpu().numpy(), int(st[0]), int(st[1])


def pack_records(ids: torch.Tensor, scores: torch.Tensor):
    """[Q, k] (int64 global id, float64 score) -> int64[Q, k, 2] records (the score's bits in [..., 1])."""
    return torch.stack([ids.to(torch.int64), scores.to(torch.float64).view(torch.int64)], dim=-1).contiguous()


def gather_records(rec: torch.Tensor, group=None):
    """ONE all-gather of the per-rank packed [Q, k, 2] records -> [G, Q, k, 2]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return rec[None]
    out = torch.empty((world * rec.shape[0],) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    return out.view((world,) + tuple(rec.shape))


def gather_candidates(ids: torch.Tensor, scores: torch.Tensor, group=None):
    """all-gather of per-rank [Q, k] candidates -> ([G, Q, k] ids, [G, Q, k] scores), one collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ids[None], scores[None]
    allr = gather_records(pack_records(ids, scores), group)
    return allr[..., 0].contiguous(), allr[..., 1].contiguous().view(torch.float64)


def merge_records_cuda(all_rec: torch.Tensor, k: int):
    """br_topk_merge_records: packed records [G, Q, k, 2] -> ([Q, k] ids, [Q, k] scores) by (score desc, id asc)."""
    lib = _lib.load()
    g, q, kk, _ = all_rec.shape
    dev = all_rec.device
    if kk != k:
        raise ValueError("candidate width must equal k")
    out_ids = torch.empty((q, k), dtype=torch.int64, device=dev)
    out_sc = torch.empty((q, k), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.br_topk_merge_records(ptr(all_rec.contiguous()), g, q, k, ptr(out_ids), ptr(out_sc), _lib.stream_ptr(dev)),
              "br_topk_merge_records")
    return out_ids, out_sc


def merge_topk_cuda(all_ids: torch.Tensor, all_sc: torch.Tensor, k: int):
    """br_topk_merge: [G, Q, k] -> [Q, k] by (score desc, id asc)."""
    lib = _lib.load()
    g, q, kk = all_ids.shape
    dev = all_ids.device
    out_ids = torch.empty((q, k), dtype=torch.int64, device=dev)
    out_sc = torch.empty((q, k), dtype=torch.float64, device=dev)
    if kk != k:
        raise ValueError("candidate width must equal k")
    with torch.cuda.device(dev):
        check(lib.br_topk_merge(ptr(all_ids.contiguous()), ptr(all_sc.contiguous()), g, q, k, ptr(out_ids), ptr(out_sc),
                                _lib.stream_ptr(dev)), "br_topk_merge")
    return out_ids, out_sc


class ShardedBM25:
    """One rank's view of a doc-sharded index.  ``local`` is this rank's ``BM25`` shard (built with
    ``doc_base`` = first global doc id of the shard and the all-reduced statistics)."""

    def __init__(self, local, group=None, merge=merge_topk_cuda, share_thresholds=True):
        self.local, self.group, self.merge = local, group, merge
        self.share_thresholds = share_thresholds
        self._xr = {}                    # k -> exchange rounds agreed between the ranks (-1: off)
        self._cb = None
        self._thr_views = {}

    # ---- threshold sharing (br_set_thr_exchange): every shard prunes with a bound learnt on the WHOLE corpus.  The library
    # hands over its thresholds (after seeding) or its current k best scores per query (after a tile launch); this callback
    # all-gathers them over the shards on the current stream and the library takes the maximum / the k-th largest of the
    # union - the threshold a single index would have after the same fraction of its docs.
    def _exchange(self, local_ptr, gathered_ptr, n_floats, stream, user):
        try:
            world = dist.get_world_size(self.group)
            key = (local_ptr, gathered_ptr, n_floats)
            view = self._thr_views.get(key)
            if view is None:
                def wrap(ptr, n):
                    class _Arr:
                        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3}
                    return torch.as_tensor(_Arr(), device=self.local._device)
                view = (wrap(local_ptr, n_floats), wrap(gathered_ptr, n_floats * world))
                self._thr_views[key] = view
            dist.all_gather_into_tensor(view[1], view[0], group=self.group)     # on the current stream
            return 0
        except Exception:                # never raise through the C frame
            import traceback
            traceback.print_exc()
            return -1

    def _setup_exchange(self, k):
        """Agree on the number of exchanges per batch (same on every rank, or none at all) for this k."""
        if k in self._xr:
            return self._xr[k]
        rounds = -1
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if (self.share_thresholds and world > 1 and hasattr(self.local, "_h") and self.local._h is not None
                and self.merge is merge_topk_cuda):
            lib = _lib.load()
            n = int(lib.br_tile_launch_count(self.local._h, int(k)))
            t = torch.tensor([n], dtype=torch.int64, device=self.local._device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            rounds = int(t.item()) - 1 if int(t.item()) > 0 else -1
        self._xr[k] = rounds
        return rounds

    def _set_exchange(self, rounds):
        lib = _lib.load()
        if rounds >= 0:
            if self._cb is None:
                self._cb = _lib.THR_EXCHANGE_FN(self._exchange)
            check(lib.br_set_thr_exchange(self.local._h, self._cb, None, int(rounds), dist.get_world_size(self.group)), "br_set_thr_exchange")
        else:
            check(lib.br_set_thr_exchange(self.local._h, None, None, -1, 1), "br_set_thr_exchange")

    @classmethod
    def from_local_token_ids(cls, doc_offsets, token_ids, vocab_size, doc_base, k1=1.5, b=0.75, *,
                             variant="notebook", dedup_query=None, device=None, group=None):
        from .bm25 import BM25
        m = BM25.from_token_ids(doc_offsets, token_ids, vocab_size, k1, b, variant=variant, dedup_query=dedup_query,
                                device=device, doc_base=doc_base, finalize=False)
        st = m.stats()
        df, n_stat, sum_dl = reduce_stats(m.local_df_tensor(), st["n_docs"], st["sum_dl"], group)
        m.finalize(n_stat, sum_dl, df)
        self = cls(m, group)
        self.n_docs_global = n_stat
        return self

    def retrieve_top_n_batch(self, queries, n=10):
        """-> (global ids int64[Q, n], float64 scores[Q, n]) identical on every rank."""
        if hasattr(self.local, "retrieve_records_batch") and self.merge is merge_topk_cuda:
            # CUDA shard: the library emits packed {global id, score} records - one all-gather, merged straight from
            # the gathered buffer
            rounds = self._setup_exchange(n)
            if rounds >= 0:
                self._set_exchange(rounds)
            try:
                rec = self.local.retrieve_records_batch(queries, n)
            finally:
                if rounds >= 0:
                    self._set_exchange(-1)
            allr = gather_records(rec, self.group)
            if allr.shape[0] == 1:
                return rec[..., 0].contiguous(), rec[..., 1].contiguous().view(torch.float64)
            ids, sc = merge_records_cuda(allr, n)
            if rounds >= 0:
                # with shared thresholds a query that has fewer than n matching docs in the WHOLE corpus comes back short:
                # those (rare) queries are repeated with local thresholds, whose zero-score fill is exact
                total = getattr(self, "n_docs_global", None)
                short = torch.nonzero(ids[:, min(n, total or n) - 1] < 0).flatten()
                if short.numel():
                    q_terms, q_off = self.local.pack_queries(queries)
                    q_terms, q_off = q_terms.cpu(), q_off.cpu()
                    rows = short.cpu().tolist()
                    sub_t = torch.cat([q_terms[int(q_off[r]):int(q_off[r + 1])] for r in rows]) if rows else q_terms[:0]
                    sub_o = torch.tensor([0] + [int(q_off[r + 1] - q_off[r]) for r in rows], dtype=torch.int32).cumsum(0).to(torch.int32)
                    rec2 = self.local.retrieve_records_batch((sub_t.to(torch.int32), sub_o), n)
                    i2, s2 = merge_records_cuda(gather_records(rec2, self.group), n)
                    ids[short], sc[short] = i2, s2
            return ids, sc
        ids, sc = self.local.retrieve_top_n_batch(queries, n)
        gids = torch.where(ids >= 0, ids.to(torch.int64) + self.local.doc_base, torch.full_like(ids, -1, dtype=torch.int64))
        all_ids, all_sc = gather_candidates(gids, sc, self.group)
        if all_ids.shape[0] == 1:
            return gids, sc
        return self.merge(all_ids, all_sc, n)


class ShardedCosineIndex:
    """Row-sharded brute-force cosine top-k (BASELINE config 5: 10 M x 768 bf16 over 8 GPUs; the
    reference's single-process path is team_run1.py:269-282).  Each rank holds the embedding rows
    ``[row_base, row_base + n_local)``, queries are replicated (10 k x 768 x 2 B = 15 MB), every rank
    runs the fused GEMM + top-k on its rows and only the [Q, k] (global row, cosine) candidates are
    all-gathered and merged - bit-identical to the single-GPU result because top-k over a row
    partition is exactly decomposable and ties break on the global row id."""

    def __init__(self, local_embeddings, row_base, device=None, group=None, merge=merge_topk_cuda, local=None):
        if local is None:
            from .cosine import CosineIndex
            local = CosineIndex(local_embeddings, device=device, doc_base=row_base)
        self.local, self.group, self.merge = local, group, merge

    def topk(self, query_embeddings, k=10):
        """-> (global rows int64[Q, k], cosine float64[Q, k] (fp32 values widened)) on every rank."""
        ids, sims = self.local.topk(query_embeddings, k)
        sims = sims.to(torch.float64)
        all_ids, all_sc = gather_candidates(ids, sims, self.group)
        if all_ids.shape[0] == 1:
            return ids, sims
        return self.merge(all_ids, all_sc, k)
