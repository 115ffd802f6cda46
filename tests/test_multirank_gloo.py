"""world_size-2 host logic on CPU (gloo): shard assignment per rank, the per-rank partial-state wire format
and the merge at the N->1 exchange point (processors/processor_merge.rs:37-66 -> transform_aggregate_final.rs).
The kernel is stood in for by the oracle's per-shard aggregates (no GPU here); everything else is product code:
fuse_query_b200.shards and the C++ host mirror's AggregateFinalTransform merge."""
import os

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from fuse_query_b200 import _fuse_host as h
from fuse_query_b200.shards import generate_parts, shard_for_rank

NUM = "(col number)"


def _worker(rank, world, total, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as o
    begin, n = shard_for_rank(rank, world, total)
    # per-rank partial states of: sum(number)/count(number), max(number), min(number)  (leaf order sum,count | max | min)
    data = {"number": o.array(o.U64, range(begin, begin + n))}
    leaf = lambda e: o.run_query([e], table=data, is_aggregate=True, worker_threads=1, tail_quirk=False).columns[0].to_list()[0]
    U = h.DataType.UInt64
    rows = [h.DataValue(h.DataType.Struct, [h.DataValue(U, leaf(f"(sum {NUM})")), h.DataValue(U, n)]).to_json(),
            h.DataValue(h.DataType.Struct, [h.DataValue(U, leaf(f"(max {NUM})"))]).to_json(),
            h.DataValue(h.DataType.Struct, [h.DataValue(U, leaf(f"(min {NUM})"))]).to_json()]
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    if rank == 0:
        ctx = h.FuseQueryContext.create_ctx(1)
        plan = h.Planner().build_from_sql(ctx, f"select sum(number)/count(number), max(number), min(number) from system.numbers_mt({total})")
        agg = plan.children_to_plans()[1]
        funcs = [e.to_function() for e in agg.expr]
        for states in gathered:           # AggregateFinalTransform::execute's merge loop, one partial block per rank
            for f, js in zip(funcs, states):
                f.merge_state(h.DataValue.from_json(js).value)
        out.put([f.merge_result().value for f in funcs])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [16, 100_000])
def test_two_rank_partial_state_merge(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, total, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = total * (total - 1) // 2
    assert got == [s // total, total - 1, 0]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("total", [8, 10**7, 10**10, 10**10 + 5])
def test_shards_cover_the_table_like_the_reference_partitions(world, total):
    parts = generate_parts(total)
    assert len(parts) == 8 and parts[0][0] == 0 and parts[-1][1] == total - 1
    shards = [shard_for_rank(r, world, total) for r in range(world)]
    assert shards[0][0] == 0 and sum(n for _, n in shards) == total
    for (b0, n0), (b1, _) in zip(shards, shards[1:]):
        assert b0 + n0 == b1
    # every shard is a union of whole reference partitions
    starts = {p[0] for p in parts}
    assert all(b in starts for b, _ in shards)
