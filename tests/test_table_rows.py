"""f-1: readers for the rows RAGLite stores in ``chunk_embedding`` (DuckDB ``FLOAT[d]`` lists, PostgreSQL
``halfvec`` text, ``np.save`` blobs) against fixtures produced by the REFERENCE's own column processors
(``raglite/_typing.py:57-78, 145-208``, loaded by ``tools/make_golden_from_reference.py``)."""

from __future__ import annotations

import json

import numpy as np
import pytest


@pytest.fixture(scope="module")
def fx(golden_dir):
    z = np.load(golden_dir / "table_rows.npz")
    blobs, off = [], 0
    raw = bytes(z["blobs"])
    for n in z["blob_len"]:
        blobs.append(raw[off:off + int(n)])
        off += int(n)
    return dict(E16=z["E16"], from_text=z["from_text"], from_list=z["from_list"], from_blob=z["from_blob"],
                texts=json.loads(bytes(z["texts"]).decode()), lists=json.loads(bytes(z["lists"]).decode()), blobs=blobs,
                row_chunk=z["row_chunk"])


def test_value_processors_match_the_reference_bit_for_bit(fx):
    from raglite_b200 import _rows

    for i, row in enumerate(fx["E16"]):
        assert _rows.vector_to_halfvec_text(row) == fx["texts"][i]                    # PostgresHalfVec.bind_processor
        assert _rows.vector_to_duckdb_list(row) == fx["lists"][i]                     # DuckDBSingleVec.bind_processor
        assert _rows.vector_to_numpy_blob(row) == fx["blobs"][i]                      # NumpyArray.process_bind_param
        got = _rows.halfvec_text_to_vector(fx["texts"][i])
        assert got.dtype == np.float16 and np.array_equal(got.view(np.uint16), fx["from_text"][i].view(np.uint16))
        got = _rows.duckdb_list_to_vector(fx["lists"][i])
        assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), fx["from_list"][i].view(np.uint32))
        got = _rows.numpy_blob_to_vector(fx["blobs"][i])
        assert got.dtype == np.float16 and np.array_equal(got.view(np.uint16), fx["from_blob"][i].view(np.uint16))
    assert _rows.halfvec_text_to_vector(None) is None and _rows.duckdb_list_to_vector(None) is None


def test_batched_readers_and_storage_choice(fx):
    from raglite_b200 import _rows
    from raglite_b200._index import CorpusIndex

    M = _rows.halfvec_rows_to_matrix(fx["texts"])
    assert M.dtype == np.float16 and np.array_equal(M.view(np.uint16), fx["from_text"].view(np.uint16))
    D = _rows.duckdb_rows_to_matrix(fx["lists"])
    assert D.dtype == np.float32 and np.array_equal(D, fx["from_list"])
    ids = [f"chunk-{c}" for c in fx["row_chunk"]]
    got_ids, E = _rows.table_rows(zip(ids, fx["blobs"]), "numpy")
    assert got_ids == ids and np.array_equal(E.view(np.uint16), fx["from_blob"].view(np.uint16))
    with pytest.raises(ValueError):
        _rows.halfvec_rows_to_matrix(["[1,2,3]", "[1,2]"])
    with pytest.raises(ValueError):
        _rows.table_rows([("a", [1.0])], "sqlite")
    # FLOAT[d] values that RAGLite wrote are float16-representable: the lossless 2-byte layout is chosen ...
    unit = fx["from_list"][2:]                      # (rows 0-1 hold the edge values: 65504 is too large for the fp16 scan)
    E, storage = CorpusIndex._pick_storage(unit, "auto")
    assert storage == "fp16" and E.dtype == np.float16 and np.array_equal(E.astype(np.float32), unit)
    # ... anything else stays float32
    E, storage = CorpusIndex._pick_storage(unit + np.float32(1e-5), "auto")
    assert storage == "fp32" and E.dtype == np.float32
    E, storage = CorpusIndex._pick_storage(fx["from_list"], "auto")
    assert storage == "fp32"


@pytest.mark.gpu
@pytest.mark.parametrize("dialect", ["duckdb", "postgresql", "numpy"])
def test_index_from_driver_rows_equals_index_from_matrix(fx, dialect):
    import raglite_b200 as rl
    from parity import check_sql_semantics
    from synth import make_queries

    rows_of = {"duckdb": fx["lists"], "postgresql": fx["texts"], "numpy": fx["blobs"]}[dialect]
    ids = [f"chunk-{c}" for c in fx["row_chunk"]]
    keep = slice(2, None)   # rows 0-1 carry edge values (65504) that are not embeddings
    idx = rl.CorpusIndex.from_table_rows(list(zip(ids[keep], rows_of[keep])), dialect)
    assert idx.storage == "fp16" and idx.chunk_ids[0] == ids[2]
    E = fx["from_text"][keep].astype(np.float32)
    off = idx.chunk_off
    Q = make_queries(E, 4, seed=5)
    out_ids, sims, counts = rl.vector_search_batch(Q, num_results=5, config=rl.RAGLiteConfig(reranker=None), index=idx)
    for b in range(len(Q)):
        check_sql_semantics(E, off, Q[b], out_ids[b, :counts[b]], sims[b, :counts[b]], k=5)
    half = len(ids) // 2
    while ids[half] == ids[half - 1]:   # cut between two chunks (a flush never splits a chunk's vectors)
        half += 1
    idx2 = rl.CorpusIndex.from_table_rows(list(zip(ids[2:half], rows_of[2:half])), dialect)
    idx2.append_table_rows(list(zip(ids[half:], rows_of[half:])), dialect)
    assert np.array_equal(idx2.chunk_off, idx.chunk_off) and idx2.chunk_ids == idx.chunk_ids
    assert np.array_equal(idx2.E.cpu().numpy().view(np.uint16), idx.E.cpu().numpy().view(np.uint16))
