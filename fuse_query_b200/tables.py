"""Registering host data (numpy / pyarrow) as HBM-resident tables of a FuseQueryContext."""
from __future__ import annotations

from typing import Mapping

import numpy as np

from . import _fuse_host as h

_TAG = {np.dtype(np.bool_): h.DataType.Boolean, np.dtype(np.int8): h.DataType.Int8, np.dtype(np.int16): h.DataType.Int16,
        np.dtype(np.int32): h.DataType.Int32, np.dtype(np.int64): h.DataType.Int64, np.dtype(np.uint8): h.DataType.UInt8,
        np.dtype(np.uint16): h.DataType.UInt16, np.dtype(np.uint32): h.DataType.UInt32, np.dtype(np.uint64): h.DataType.UInt64,
        np.dtype(np.float32): h.DataType.Float32, np.dtype(np.float64): h.DataType.Float64}


def _arrow_column(col):
    """pyarrow ChunkedArray/Array -> (values, validity or None).  The validity comes back as ("bitmap", buffer, bit
    offset): the arrow null bitmap itself (1 bit per row, LSB first), which the device expands to its own layout (one
    byte per row) — no host-side unpacking."""
    import pyarrow as pa
    arr = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    if not arr.null_count:
        return arr.to_numpy(zero_copy_only=False), None
    zero = False if pa.types.is_boolean(arr.type) else 0
    values = arr.fill_null(zero).to_numpy(zero_copy_only=False)   # the payload of a NULL slot is never read as a value
    return values, ("bitmap", arr.buffers()[0], arr.offset)


def register_table(ctx, gpu, db: str, name: str, columns: Mapping[str, "np.ndarray"]):
    """Upload `columns` to HBM and register them as table `db.name`.  Returns the MemoryTable.

    A column is a numpy array (NOT NULL), a numpy masked array, a `(values, valid)` pair (valid: 1 = not NULL), or a
    pyarrow array; a pyarrow.Table is taken column by column.  Columns with NULLs become nullable fields whose validity
    rides beside the values as one byte per row."""
    if hasattr(columns, "column_names"):  # pyarrow.Table
        tbl = columns
        columns = {n: _arrow_column(tbl.column(n)) for n in tbl.column_names}
    fields, arrays = [], []
    for n, a in columns.items():
        valid = None
        if isinstance(a, tuple):
            a, valid = a
        elif isinstance(a, np.ma.MaskedArray):
            a, valid = a.filled(0), (None if a.mask is np.ma.nomask else (~np.ma.getmaskarray(a)).astype(np.uint8))
        elif hasattr(a, "null_count"):          # pyarrow Array / ChunkedArray
            a, valid = _arrow_column(a)
        a = np.ascontiguousarray(np.asarray(a))
        if a.dtype not in _TAG:
            raise h.FuseQueryError(f"Internal Error: Unsupported on the device path: column {n} of dtype {a.dtype}")
        fields.append(h.DataField(n, _TAG[a.dtype], valid is not None))
        if valid is None:
            arrays.append(h.DataArray.from_numpy(gpu, a))
        elif isinstance(valid, tuple) and valid[0] == "bitmap":
            arr = h.DataArray.from_numpy(gpu, a)
            arr.set_validity(h.DataArray.from_arrow_bitmap(gpu, valid[1].address, valid[2], len(a)))
            arrays.append(arr)
        else:
            arrays.append(h.DataArray.from_numpy_masked(gpu, a, np.asarray(valid)))
    table = h.MemoryTable(db, name, h.DataSchema(fields), arrays)
    ds = ctx.datasource()
    ds.add_database(db) if not _has_db(ds, db) else None
    ds.add_table(db, table)
    return table


def register_parquet(ctx, gpu, db: str, name: str, path, columns=None):
    """Read a Parquet file (or a directory of them) column by column into HBM and register it as table `db.name`:
    the reference's "table storage engine" slot (`datasources/table.rs:13-22`, README "Remote (S3 or other table
    storage engine)") filled with the one format this image can decode (pyarrow).  Decoding is host work done once at
    registration; queries then run against the resident columns.  `columns` restricts what is loaded."""
    import pyarrow.parquet as pq
    return register_table(ctx, gpu, db, name, pq.read_table(str(path), columns=list(columns) if columns else None))


def register_arrow_ipc(ctx, gpu, db: str, name: str, path, columns=None):
    """Same for an Arrow IPC (Feather v2) file, memory-mapped: the uploads read straight from the page cache."""
    import pyarrow as pa
    import pyarrow.ipc as ipc
    with pa.memory_map(str(path), "r") as f:
        tbl = ipc.open_file(f).read_all()
        if columns:
            tbl = tbl.select(list(columns))
        return register_table(ctx, gpu, db, name, tbl)


def _has_db(ds, db: str) -> bool:
    try:
        ds.get_table(db, "\0")
    except h.FuseQueryError as e:
        return "Cannot find the database" not in str(e)
    return True
