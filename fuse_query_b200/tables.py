"""Registering host data (numpy / pyarrow) as HBM-resident tables of a FuseQueryContext."""
from __future__ import annotations

from typing import Mapping

import numpy as np

from . import _fuse_host as h

_TAG = {np.dtype(np.bool_): h.DataType.Boolean, np.dtype(np.int8): h.DataType.Int8, np.dtype(np.int16): h.DataType.Int16,
        np.dtype(np.int32): h.DataType.Int32, np.dtype(np.int64): h.DataType.Int64, np.dtype(np.uint8): h.DataType.UInt8,
        np.dtype(np.uint16): h.DataType.UInt16, np.dtype(np.uint32): h.DataType.UInt32, np.dtype(np.uint64): h.DataType.UInt64,
        np.dtype(np.float32): h.DataType.Float32, np.dtype(np.float64): h.DataType.Float64}


def register_table(ctx, gpu, db: str, name: str, columns: Mapping[str, "np.ndarray"]):
    """Upload `columns` (numpy arrays, or a pyarrow.Table / dict of pyarrow arrays without nulls) to HBM and register
    them as table `db.name`.  Returns the MemoryTable."""
    if hasattr(columns, "column_names"):  # pyarrow.Table
        tbl = columns
        columns = {}
        for n in tbl.column_names:
            col = tbl.column(n)
            if col.null_count:
                raise h.FuseQueryError(f"Internal Error: Unsupported on the device path: column {n} has NULLs")
            columns[n] = col.to_numpy()
    fields, arrays = [], []
    for n, a in columns.items():
        a = np.ascontiguousarray(np.asarray(a))
        if a.dtype not in _TAG:
            raise h.FuseQueryError(f"Internal Error: Unsupported on the device path: column {n} of dtype {a.dtype}")
        fields.append(h.DataField(n, _TAG[a.dtype], False))
        arrays.append(h.DataArray.from_numpy(gpu, a))
    table = h.MemoryTable(db, name, h.DataSchema(fields), arrays)
    ds = ctx.datasource()
    ds.add_database(db) if not _has_db(ds, db) else None
    ds.add_table(db, table)
    return table


def _has_db(ds, db: str) -> bool:
    try:
        ds.get_table(db, "\0")
    except h.FuseQueryError as e:
        return "Cannot find the database" not in str(e)
    return True
