"""fq_graph: the launches of several pipes recorded once (CUDA stream capture) and replayed — every replay must give the
same results as direct launches, the fetch calls must work unchanged, and what cannot be replayed is refused.
(The reference rebuilds and re-runs its pipeline per query, interpreter_select.rs:27-40; this is the prepared-statement path.)"""
import pytest

from fuse_query_b200 import cabi
from oracle import binding as o
from oracle.groupby import run_group_by

pytestmark = pytest.mark.gpu
NUM = "(col number)"
README_PRED = f"(< (+ (+ (+ {NUM} (u64 1)) (/ {NUM} (u64 2))) (u64 1)) (u64 100))"
PROJ = [f"(alias c1 (+ {NUM} (u64 1)))", f"(alias c2 (/ {NUM} (u64 2)))"]


@pytest.fixture(scope="module")
def ctx():
    c = cabi.Context(0)
    yield c
    c.close()


def test_three_pipes_recorded_once_and_replayed(ctx):
    import torch
    stream = torch.cuda.Stream().cuda_stream
    n = 5_000_011
    col = ctx.numbers(0, n)
    src = cabi.make_source([col], n)
    agg = ctx.pipe([f"(sum {NUM})", f"(max {NUM})", f"(min {NUM})"], aggregate=True)
    sel = ctx.pipe(PROJ, predicate=README_PRED)
    outs = [ctx.column(cabi.U64, 100), ctx.column(cabi.U64, 100)]
    key = f"(- {NUM} (* (/ {NUM} (u64 13)) (u64 13)))"
    gb = ctx.pipe([f"(sum {NUM})", f"(count {NUM})"], keys=[key])
    gb.groupby_reserve(13)
    # direct launches first: the reference results of this file (themselves checked against the oracle below)
    agg.launch_aggregate(src, stream=stream)
    sel.launch_project(src, outs, 100, limit=3, stream=stream)
    gb.launch_groupby(src, stream=stream)
    want_states = agg.fetch_aggregate()
    want_sel = sel.fetch_project()
    want_rows = [c.to_numpy(3, stream=stream).tolist() for c in outs]
    assert gb.fetch_groupby() == 13
    ref = o.run_query(PROJ, total=n, predicate=README_PRED, limit=3, worker_threads=1)
    assert list(zip(*want_rows)) == ref.rows() and want_sel == (66, 3)
    assert want_states[0][0][1] == n * (n - 1) // 2 and want_states[1] == n

    launches0 = ctx.launch_count
    ctx.graph_begin(stream)
    agg.launch_aggregate(src, stream=stream)
    sel.launch_project(src, outs, 100, limit=3, stream=stream)
    gb.launch_groupby(src, stream=stream)
    graph = ctx.graph_end(stream)
    assert ctx.launch_count == launches0          # recording runs nothing
    for c in outs:                                # the replays must rewrite these
        ctx.fill_numbers(c, 777, 100, stream)
    per_replay = None
    for _ in range(4):
        before = ctx.launch_count
        graph.launch(stream)
        assert agg.fetch_aggregate() == want_states
        assert sel.fetch_project() == want_sel
        assert [c.to_numpy(3, stream=stream).tolist() for c in outs] == want_rows
        assert gb.fetch_groupby() == 13
        per_replay = ctx.launch_count - before
    assert per_replay >= 4                         # aggregate + select + group-by kernels are counted per replay
    keys, _, leaves, _ = gb.export_groups(13, stream=stream)
    got = sorted(zip(keys[0].to_numpy(13).tolist(), leaves[0].to_numpy(13).tolist(), leaves[1].to_numpy(13).tolist()))
    assert got == run_group_by([key], [f"(sum {NUM})", f"(count {NUM})"], total=n)[1]   # a replay restarts the table: no doubling
    graph.destroy()
    for p in (agg, sel, gb):
        p.destroy()
    for c in outs + keys + leaves + [col]:
        c.free()


def test_what_cannot_be_recorded_is_refused(ctx):
    import torch
    stream = torch.cuda.Stream().cuda_stream
    with pytest.raises(cabi.FuseGpuError) as e:
        ctx.graph_begin(0)
    assert "explicit stream" in str(e.value)
    with pytest.raises(cabi.FuseGpuError) as e:
        ctx.graph_end(stream)
    assert "without a matching fq_graph_begin" in str(e.value)
    n = 100_000
    col = ctx.numbers(0, n)
    pipe = ctx.pipe([f"(sum {NUM})"], aggregate=True)
    group = ctx.group(0, 1)
    pipe.set_group(group)
    ctx.graph_begin(stream)
    with pytest.raises(cabi.FuseGpuError) as e:
        ctx.graph_begin(stream)
    assert "already being recorded" in str(e.value)
    with pytest.raises(cabi.FuseGpuError) as e:
        pipe.launch_aggregate(cabi.make_source([col], n), stream=stream)
    assert "cannot be recorded" in str(e.value)
    pipe.set_group(None)
    pipe.launch_aggregate(cabi.make_source([col], n), stream=stream)
    graph = ctx.graph_end(stream)
    graph.launch(stream)
    states, rows = pipe.fetch_aggregate()
    assert rows == n and states[0][1] == n * (n - 1) // 2
    graph.destroy()
    pipe.destroy()
    group.destroy()
    col.free()
