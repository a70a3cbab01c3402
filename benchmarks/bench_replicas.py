"""Query-sharded replicas: the trivial multi-GPU baseline SURVEY 8(e) asks to report beside doc-sharding.
Every rank holds the WHOLE C4 index (3.7 GB of postings) and answers its own 1/N of the 10k queries; nothing
crosses NVLink.  Same corpus / queries / timing rules as bench.py (barrier + device events, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        benchmarks/bench_replicas.py --gpus N
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    bench._quiet_stdout()
    args = bench.parse()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from document_retrieval_b200 import BM25
    doc_offsets, token_ids = bench.gen_blocks(args, range(bench.N_BLOCKS), dev)
    q_terms, q_off, _ = bench.gen_queries(args, doc_offsets, token_ids, 0, args.docs, 0, 1)
    model = BM25.from_token_ids(doc_offsets, token_ids, args.vocab, device=dev)
    del doc_offsets, token_ids
    per = -(-args.queries // world)
    lo, hi = min(args.queries, rank * per), min(args.queries, (rank + 1) * per)
    t = torch.from_numpy(q_terms[q_off[lo]:q_off[hi]]).to(dev)
    o = torch.from_numpy(q_off[lo:hi + 1] - q_off[lo]).to(dev)

    def step():
        return model.retrieve_top_n_batch((t, o), args.k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / args.steps
    if rank == 0:
        bench._emit({"metric": "BM25 top-10 queries/sec", "value": args.queries / (ms_step * 1e-3), "unit": "queries/s",
                     "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                     "higher_is_better": True, "scaling": "strong", "data": "synthetic",
                     "config": {"workload": f"C4: {args.docs} docs, {args.queries} queries, top-{args.k}",
                                "sharding": f"query-sharded replicas: full index on each of {world} GPU(s), "
                                            f"{hi - lo} queries per GPU, no collective"}})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
