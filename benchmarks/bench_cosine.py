#!/usr/bin/env python
"""Secondary benchmark: the dense cosine kernels (BASELINE.json configs[4] per-GPU shard and configs[2]).
  C5 shard : brute-force cosine top-10, 1.25M x 768 bf16 docs (one of 8 shards of 10M), 10k queries
             -> TFLOP/s against MEASURED_PEAKS.json bf16_tflops (tensor roofline)
  C3       : cosine re-rank of 1000 candidates per query (BM25 top-1000 shape), 207,363 docs, 10k queries
             -> GB/s of gathered rows against hbm_gbs (gather / HBM roofline)
Prints one JSON line per config.  CPU baseline: torch.matmul + topk fp32 on the host (team_run1.py:280-282)
on a bounded sample."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from document_retrieval_b200.cosine import CosineIndex  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def main_sharded(args):
    """Config 5 proper: --total-docs rows sharded over the ranks of one box (torchrun), NCCL all-gather of the
    [Q, k] candidates + merge.  Timed on the device, max over ranks."""
    import torch.distributed as dist
    from document_retrieval_b200.sharded import ShardedCosineIndex, shard_bounds
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    hbm, tf_burst, tf_sus, src = peaks()
    lo, hi = shard_bounds(args.total_docs, world)[rank]
    g = torch.Generator(device=dev).manual_seed(20241105 + 5 + 1000 * rank)
    docs = torch.empty(hi - lo, args.dim, device=dev, dtype=torch.bfloat16)
    for a in range(0, hi - lo, 1 << 20):           # chunks: no fp32 copy of the whole shard
        b = min(hi - lo, a + (1 << 20))
        docs[a:b] = torch.randn(b - a, args.dim, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    gq = torch.Generator(device=dev).manual_seed(20241105 + 55)
    qs = torch.randn(args.queries, args.dim, generator=gq, device=dev, dtype=torch.float32).to(torch.bfloat16)
    ix = ShardedCosineIndex(docs, lo, device=dev)
    for _ in range(args.warmup):
        ids, sims = ix.topk(qs, 10)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ids, sims = ix.topk(qs, 10)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # every rank holds the same merged result; its local part must agree with a torch fp32 check on the local rows
    mine = (ids[:64] >= lo) & (ids[:64] < hi)
    dn = docs.float() if hi - lo <= 2_000_000 else None
    ok = None
    if dn is not None:
        dn = dn / (dn.norm(dim=1, keepdim=True) + 1e-10)
        qn = qs[:64].float()
        qn = qn / (qn.norm(dim=1, keepdim=True) + 1e-10)
        full = qn @ dn.T
        got = torch.where(mine, full.gather(1, (ids[:64] - lo).clamp(0, hi - lo - 1)), torch.zeros_like(sims[:64], dtype=torch.float32))
        ok = float(((got - sims[:64].float()).abs() * mine).max().item())
    if rank == 0:
        flop = 2.0 * args.total_docs * args.queries * args.dim
        tfs = flop / (ms * 1e-3) / 1e12
        print(json.dumps({
            "metric": "cosine top-10 queries/sec (brute force)", "value": args.queries / (ms * 1e-3), "unit": "queries/s",
            "n_gpus": world, "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup, "scaling": "strong",
            "config": {"workload": f"C5: {args.total_docs} x {args.dim} bf16 docs row-sharded over {world} GPU(s), "
                                   f"{args.queries} queries, k=10, NCCL all-gather of [Q,k] + merge"},
            "roofline": {"bound": "tensor", "achieved": tfs / world, "peak": tf_burst, "peak_sustained": tf_sus,
                         "unit": "TFLOP/s per GPU", "frac": tfs / world / tf_burst, "frac_of_sustained": tfs / world / tf_sus,
                         "peak_source": src},
            "check": {"max_abs_sim_err_vs_torch_fp32_local_rows": ok}}))
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-docs", type=int, default=10_000_000)
    ap.add_argument("--docs", type=int, default=1_250_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-docs", type=int, default=100_000)
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_sharded(args)
    dev = torch.device("cuda", 0)
    hbm, tf_burst, tf_sus, src = peaks()
    g = torch.Generator(device=dev).manual_seed(20241105 + 5)
    docs = torch.randn(args.docs, args.dim, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    qs = torch.randn(args.queries, args.dim, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    ix = CosineIndex(docs)
    ms, (ids, sims) = timed(lambda: ix.topk(qs, 10), args.steps, args.warmup)
    flop = 2.0 * args.docs * args.queries * args.dim
    tfs = flop / (ms * 1e-3) / 1e12
    # spot check against torch on the GPU (fp32 of the bf16 values), 64 queries
    dn = docs.float()
    dn = dn / (dn.norm(dim=1, keepdim=True) + 1e-10)
    qn = qs[:64].float()
    qn = qn / (qn.norm(dim=1, keepdim=True) + 1e-10)
    ref = torch.topk(qn @ dn.T, 10)
    ok_ids = float((ref.indices == ids[:64]).float().mean().item())
    max_err = float((ref.values - sims[:64]).abs().max().item())
    # CPU baseline: torch fp32 on the host over a doc subsample
    cd = docs[:args.cpu_docs].float().cpu()
    cq = qs.float().cpu()
    t0 = time.time()
    cdn = cd / (cd.norm(dim=1, keepdim=True) + 1e-10)
    cqn = cq / (cq.norm(dim=1, keepdim=True) + 1e-10)
    torch.topk(cqn @ cdn.T, 10)
    cpu_s = time.time() - t0
    cpu_qps_full = args.queries / (cpu_s * args.docs / args.cpu_docs)
    print(json.dumps({
        "metric": "cosine top-10 queries/sec (brute force)", "value": args.queries / (ms * 1e-3), "unit": "queries/s",
        "ms_per_step": ms, "config": {"workload": f"C5 shard: {args.docs} x {args.dim} bf16 docs, {args.queries} queries, k=10"},
        "roofline": {"bound": "tensor", "achieved": tfs, "peak": tf_burst, "peak_sustained": tf_sus, "unit": "TFLOP/s",
                     "frac": tfs / tf_burst, "frac_of_sustained": tfs / tf_sus, "peak_source": src,
                     "note": "whole call: GEMM chunks + tighten kernels + query norms"},
        "check": {"ids_equal_torch_fp32_frac": ok_ids, "max_abs_sim_err": max_err},
        "cpu_baseline": {"value": cpu_qps_full, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"torch fp32 matmul+topk, all queries x first {args.cpu_docs} docs ({cpu_s:.1f} s), scaled to {args.docs} docs"},
    }))
    # ---- C3: re-rank of 1000 candidates per query
    n3, c = 207_363, 1000
    docs3 = docs[:n3].contiguous()
    ix3 = CosineIndex(docs3)
    cand = torch.randint(0, n3, (args.queries, c), generator=g, device=dev, dtype=torch.int32)
    ms3, (i3, s3) = timed(lambda: ix3.rerank(qs, cand, 10), args.steps, args.warmup)
    bytes3 = args.queries * (c * args.dim * 2 + args.dim * 2 + 80)
    print(json.dumps({
        "metric": "cosine re-rank queries/sec (1000 candidates -> top-10)", "value": args.queries / (ms3 * 1e-3),
        "unit": "queries/s", "ms_per_step": ms3,
        "config": {"workload": f"C3: {n3} x {args.dim} bf16 docs, {args.queries} queries x {c} candidates, k=10"},
        "roofline": {"bound": "hbm", "achieved": bytes3 / (ms3 * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": bytes3 / (ms3 * 1e-3) / 1e9 / hbm, "peak_source": src,
                     "note": "algorithmic bytes = c*D*2 + D*2 + 8k per query (SURVEY 8d); the 318 MB table is L2-missing"},
    }))


if __name__ == "__main__":
    main()
