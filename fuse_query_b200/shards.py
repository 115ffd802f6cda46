"""Partitioning of system.numbers_mt(total) — host-side restatement of NumbersTable::generate_parts
(datasources/system/numbers_table.rs:29-55) used to shard the table over GPUs."""
from typing import List, Tuple


def generate_parts(total: int) -> List[Tuple[int, int]]:
    """8 contiguous [start, end] (inclusive) ranges; the last takes total % 8; total < 8 -> one part."""
    workers = 8
    chunk = total // workers
    if chunk == 0:
        return [(0, total - 1)]
    parts = [(p * chunk, (p + 1) * chunk - 1) for p in range(workers)]
    rem = total % workers
    if rem:
        parts[-1] = (parts[-1][0], parts[-1][1] + rem)
    return parts


def shard_for_rank(rank: int, world: int, total: int) -> Tuple[int, int]:
    """(begin, n_rows) of the consecutive partitions rank `rank` of `world` owns, chunked like
    processors/pipeline_builder.rs:73-95 chunks partitions over workers."""
    parts = generate_parts(total)
    per = max(1, len(parts) // world)
    mine = parts[rank * per:(rank + 1) * per] if rank < world - 1 else parts[rank * per:]
    if not mine:
        return parts[-1][1] + 1, 0
    return mine[0][0], mine[-1][1] - mine[0][0] + 1
