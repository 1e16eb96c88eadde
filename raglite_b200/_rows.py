"""Readers for the rows RAGLite stores in its ``chunk_embedding`` table.

The reference declares the ``embedding`` column as ``Embedding(dim)`` (``_typing.py:211-230``), which
becomes, per dialect:

* DuckDB ``FLOAT[d]`` (``DuckDBSingleVec``, ``_typing.py:178-208``): bound as ``np.ravel(v).tolist()``,
  read back as a Python ``list[float]`` -> ``np.asarray(value, dtype=np.float32)``.  The VALUES are
  float16-rounded (``_embed.py:140`` casts every embedding to float16 before it is inserted); DuckDB only
  widens them.
* PostgreSQL ``halfvec(d)`` (``PostgresHalfVec``, ``_typing.py:145-175``): bound as the text
  ``"[x0,x1,...]"`` with ``str()`` of each float16, read back with
  ``np.fromstring(value.strip("[]"), sep=",", dtype=np.float16)``.
* any other dialect: a ``np.save`` blob (``NumpyArray``, ``_typing.py:57-78``).

This module restates those processors without SQLAlchemy (the drivers -- ``duckdb``, ``pg8000`` -- are
what a deployment brings; they hand back exactly these Python values) and turns a result set read in
insertion order (``SELECT chunk_id, embedding FROM chunk_embedding ORDER BY id``, ``_database.py:403-430``)
into the matrix + CSR the device index is built from -- float16 when the values are float16-representable
(always true for data RAGLite wrote), so the lossless 2-byte layout is the default for real tables.
"""

from __future__ import annotations

import io
from collections.abc import Iterable, Sequence
from typing import Any

import numpy as np

from ._typing import ChunkId


# ---- value processors (one row) ----------------------------------------------------------------------
def duckdb_list_to_vector(value: Sequence[float] | None) -> np.ndarray | None:
    """``DuckDBSingleVec.result_processor`` (``_typing.py:203-206``)."""
    return np.asarray(value, dtype=np.float32) if value is not None else None


def vector_to_duckdb_list(value: np.ndarray | None) -> list[float] | None:
    """``DuckDBSingleVec.bind_processor`` (``_typing.py:192-195``)."""
    return np.ravel(value).tolist() if value is not None else None


def halfvec_text_to_vector(value: str | None) -> np.ndarray | None:
    """``PostgresHalfVec.result_processor`` (``_typing.py:168-173``): decimal text -> float16.
    (The reference parses with ``np.fromstring(..., sep=",", dtype=np.float16)``: each field goes through
    a double and is then rounded to float16; ``np.array(fields, float64).astype(float16)`` is that.)"""
    if value is None:
        return None
    body = value.strip("[]")
    if not body:
        return np.zeros(0, np.float16)
    return np.array(body.split(","), dtype=np.float64).astype(np.float16)


def vector_to_halfvec_text(value: np.ndarray | None) -> str | None:
    """``PostgresHalfVec.bind_processor`` (``_typing.py:159-162``)."""
    return f"[{','.join(str(x) for x in np.ravel(value))}]" if value is not None else None


def numpy_blob_to_vector(value: bytes | None) -> np.ndarray | None:
    """``NumpyArray.process_result_value`` (``_typing.py:71-78``)."""
    return np.load(io.BytesIO(value), allow_pickle=False) if value is not None else None


def vector_to_numpy_blob(value: np.ndarray | None) -> bytes | None:
    """``NumpyArray.process_bind_param`` (``_typing.py:61-69``)."""
    if value is None:
        return None
    buf = io.BytesIO()
    np.save(buf, value, allow_pickle=False)
    return buf.getvalue()


# ---- result sets -> matrix --------------------------------------------------------------------------------
def _as_matrix(vectors: list[np.ndarray], what: str) -> np.ndarray:
    if not vectors:
        return np.zeros((0, 0), np.float16)
    d = len(vectors[0])
    for i, v in enumerate(vectors):
        if v.ndim != 1 or len(v) != d:
            raise ValueError(f"{what}: row {i} has {v.shape} values, expected ({d},)")
    return np.stack(vectors)


def halfvec_rows_to_matrix(texts: Iterable[str]) -> np.ndarray:
    """Many ``halfvec`` texts -> float16 ``[n, d]`` in ONE parse: the fields of all rows are split once and
    converted by a single ``np.array(..., float64)`` call (the per-row reference spelling costs a NumPy
    call per row; a 10M-row table needs the batched one)."""
    bodies = [t.strip("[]") for t in texts]
    if not bodies:
        return np.zeros((0, 0), np.float16)
    d = bodies[0].count(",") + 1 if bodies[0] else 0
    fields = ",".join(bodies).split(",") if d else []
    if len(fields) != d * len(bodies):
        raise ValueError("halfvec rows have different dimensions")
    return np.array(fields, dtype=np.float64).astype(np.float16).reshape(len(bodies), d)


def duckdb_rows_to_matrix(lists: Iterable[Sequence[float]]) -> np.ndarray:
    """Many DuckDB ``FLOAT[d]`` values (Python lists, or the 2-D array ``fetchnumpy`` stacks) -> float32 ``[n, d]``."""
    if isinstance(lists, np.ndarray) and lists.ndim == 2:
        return lists.astype(np.float32, copy=False)
    vecs = [np.asarray(v, dtype=np.float32) for v in lists]
    return _as_matrix(vecs, "chunk_embedding.embedding").astype(np.float32, copy=False)


def lossless_float16(E: np.ndarray) -> np.ndarray | None:
    """``E`` as float16 if every value survives the round trip (true for tables RAGLite wrote), else None."""
    if E.dtype == np.float16:
        return E
    with np.errstate(over="ignore"):
        h = E.astype(np.float16)
    return h if np.array_equal(h.astype(E.dtype), E) else None


def table_rows(rows: Iterable[tuple[Any, ...]], dialect: str) -> tuple[list[ChunkId], np.ndarray]:
    """``(chunk_id, embedding)`` tuples of a ``chunk_embedding`` result set in insertion order ->
    ``(row chunk ids, matrix)``.  ``dialect``: ``"duckdb"`` (lists), ``"postgresql"`` (halfvec text),
    ``"numpy"`` (``np.save`` blobs)."""
    ids, vals = [], []
    for cid, emb in rows:
        ids.append(str(cid))
        vals.append(emb)
    if dialect == "duckdb":
        return ids, duckdb_rows_to_matrix(vals)
    if dialect == "postgresql":
        return ids, halfvec_rows_to_matrix(vals)
    if dialect == "numpy":
        return ids, _as_matrix([np.ravel(numpy_blob_to_vector(v)) for v in vals], "chunk_embedding.embedding")
    raise ValueError(f"Unsupported dialect: {dialect}")
