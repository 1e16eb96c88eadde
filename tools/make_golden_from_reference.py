"""Generate golden vectors by running the REFERENCE's own code (dev container only).

Usage (needs /root/reference, which does not exist on the GPU box):

    python tools/make_golden_from_reference.py

``import raglite`` fails offline (duckdb/sqlalchemy/litellm/rerankers/llama_cpp are not
installed), but two pieces of the hot path are pure NumPy/SciPy once their imports resolve:

* ``raglite/_embed.py``  -- ``embed_strings_with_late_chunking`` / ``_embed_string_batch``
  (token counting, segmenting, largest-remainder split, mean pool, normalise, fp16 cast);
* ``raglite/_query_adapter.py`` -- ``_optimize_query_target``;
* ``raglite/_typing.py`` -- the column processors of the ``chunk_embedding.embedding`` column
  (DuckDB ``FLOAT[d]`` lists, PostgreSQL ``halfvec`` text, ``np.save`` blobs).

This script loads those two files *unmodified from where they lie* under a synthetic ``raglite``
package whose heavy dependencies are replaced by stubs, and whose embedder is
``tests/fake_llama.FakeLlama``.  Outputs go to ``tests/golden/*.npz`` (committed).
"""

from __future__ import annotations

import importlib.util
import json
import sys
import types
from dataclasses import dataclass
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/src/raglite")
sys.path.insert(0, str(REPO / "tests"))
from fake_llama import FakeLlama, make_sentences  # noqa: E402

GOLDEN = REPO / "tests" / "golden"


class _Anything:
    """Attribute sink: any attribute / call returns another sink (for unused ORM symbols)."""

    def __getattr__(self, name):  # noqa: ANN001, ANN204
        return _Anything()

    def __call__(self, *a, **k):  # noqa: ANN002, ANN003, ANN204
        return _Anything()


def _stub(name: str, **attrs) -> types.ModuleType:  # noqa: ANN003
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__getattr__ = lambda attr: _Anything()  # type: ignore[method-assign]
    sys.modules[name] = mod
    return mod


def _load(name: str, path: Path) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@dataclass(frozen=True)
class _Config:
    embedder: str = "llama-cpp-python/fake/fake.gguf@64"
    embedder_normalize: bool = True
    vector_search_distance_metric: str = "cosine"


_CURRENT: dict[str, FakeLlama] = {}


class _LlamaCppPythonLLM:
    @staticmethod
    def llm(model: str, **kwargs):  # noqa: ANN003, ANN205, ARG004
        return _CURRENT["llm"]


def install_reference_stubs() -> tuple[types.ModuleType, types.ModuleType]:
    pkg = types.ModuleType("raglite")
    pkg.__path__ = []  # a package, but never executes the reference's __init__ (it imports the world)
    sys.modules["raglite"] = pkg
    _stub("litellm", embedding=_Anything())
    _stub("sqlalchemy", text=_Anything())
    _stub("sqlalchemy.orm")
    _stub("sqlalchemy.orm.attributes", flag_modified=_Anything())
    _stub("sqlmodel", Session=_Anything(), col=_Anything(), select=_Anything())
    _stub("raglite._config", RAGLiteConfig=_Config)
    _stub("raglite._lazy_llama", LLAMA_POOLING_TYPE_NONE=0, Llama=FakeLlama)
    _stub("raglite._litellm", LlamaCppPythonLLM=_LlamaCppPythonLLM)
    _stub("raglite._typing", FloatMatrix=np.ndarray, FloatVector=np.ndarray, IntVector=np.ndarray)
    _stub("raglite._database")
    _stub("raglite._search", vector_search=_Anything())
    embed = _load("raglite._embed", REF / "_embed.py")
    qa = _load("raglite._query_adapter", REF / "_query_adapter.py")
    return embed, qa


def golden_pool(embed_mod: types.ModuleType) -> None:
    cases = [
        dict(name="pool_small", n_sent=9, n_ctx=64, dim=32, seed=1, normalize=True),
        dict(name="pool_multi", n_sent=70, n_ctx=96, dim=48, seed=2, normalize=True),
        dict(name="pool_nonorm", n_sent=33, n_ctx=80, dim=40, seed=3, normalize=False),
        dict(name="pool_wide", n_sent=120, n_ctx=512, dim=384, seed=4, normalize=True),
    ]
    for c in cases:
        llm = FakeLlama(n_ctx=c["n_ctx"], dim=c["dim"], seed=c["seed"])
        _CURRENT["llm"] = llm
        sentences = make_sentences(c["n_sent"], seed=c["seed"])
        cfg = _Config(embedder_normalize=c["normalize"])
        out = embed_mod.embed_strings_with_late_chunking(sentences, config=cfg)
        simple = embed_mod._embed_string_batch(sentences[:7], config=cfg)  # noqa: SLF001
        np.savez_compressed(
            GOLDEN / f"{c['name']}.npz",
            meta=np.frombuffer(json.dumps({**c, "sentences": sentences}).encode(), dtype=np.uint8),
            late_chunking=out,
            simple=simple,
        )
        print(c["name"], out.shape, out.dtype, simple.shape)


def golden_adapter(qa_mod: types.ModuleType) -> None:
    rng = np.random.default_rng(7)
    arrays = {}
    for i, (d, npos, nneg, dtype) in enumerate([(16, 2, 5, np.float16), (48, 3, 9, np.float32), (96, 1, 12, np.float16)]):
        def unit(n):  # noqa: ANN001, ANN202
            x = rng.standard_normal((n, d))
            return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(dtype)
        q, P, N = unit(1)[0], unit(npos), unit(nneg)
        t = qa_mod._optimize_query_target(q, P, N, α=0.05)  # noqa: SLF001
        arrays.update({f"q{i}": q, f"P{i}": P, f"N{i}": N, f"t{i}": t})
    np.savez_compressed(GOLDEN / "adapter_target.npz", **arrays)
    print("adapter_target", sorted(arrays))


class _Generic:
    """Stand-in for SQLAlchemy's generic base classes (``TypeDecorator[...]``, ``UserDefinedType[...]``)."""

    Comparator: type

    def __class_getitem__(cls, item):  # noqa: ANN001, ANN206
        return cls

    def __init__(self, *a, **k) -> None:  # noqa: ANN002, ANN003
        pass


_Generic.Comparator = _Generic


def load_reference_typing() -> types.ModuleType:
    """``raglite/_typing.py`` unmodified, with SQLAlchemy reduced to the few names its class bodies touch:
    the column processors (``bind_processor`` / ``result_processor`` / ``process_*``) are pure NumPy."""
    for name in ("sqlalchemy", "sqlalchemy.engine", "sqlalchemy.ext", "sqlalchemy.ext.compiler", "sqlalchemy.sql",
                 "sqlalchemy.sql.functions", "sqlalchemy.sql.operators", "sqlalchemy.types"):
        _stub(name)
    sys.modules["sqlalchemy.ext.compiler"].compiles = lambda *a, **k: (lambda fn: fn)
    sys.modules["sqlalchemy.sql.functions"].FunctionElement = _Generic
    sys.modules["sqlalchemy.engine"].Dialect = object
    sys.modules["sqlalchemy.sql.operators"].Operators = object
    t = sys.modules["sqlalchemy.types"]
    t.Float, t.LargeBinary, t.TypeDecorator, t.TypeEngine, t.UserDefinedType = _Generic, _Generic, _Generic, _Generic, _Generic
    return _load("raglite._typing_ref", REF / "_typing.py")


def golden_table_rows(typing_mod: types.ModuleType) -> None:
    """What the reference's own column processors write to / read from ``chunk_embedding.embedding``
    (``_typing.py:57-78, 145-208``) for float16 embeddings as ``_embed.py:140`` produces them."""
    rng = np.random.default_rng(11)
    d, n = 24, 40
    E = rng.standard_normal((n, d)).astype(np.float32)
    E /= np.linalg.norm(E, axis=1, keepdims=True)
    E16 = E.astype(np.float16)
    # edge values a halfvec / FLOAT[] column can hold: signed zero, subnormals, the largest half, tiny and integral values
    E16[0, :8] = np.array([0.0, -0.0, 5.96e-8, -5.96e-8, 6.1e-5, 65504.0, -65504.0, 1.0], np.float16)
    E16[1, :4] = np.array([0.333251953125, 1e-3, 123.0, -0.5], np.float16)
    duck, pg, npy = typing_mod.DuckDBSingleVec(d), typing_mod.PostgresHalfVec(d), typing_mod.NumpyArray()
    duck_bind, duck_res = duck.bind_processor(None), duck.result_processor(None, None)
    pg_bind, pg_res = pg.bind_processor(None), pg.result_processor(None, None)
    lists = [duck_bind(r) for r in E16]                     # what RAGLite hands DuckDB for FLOAT[d]
    texts = [pg_bind(r) for r in E16]                       # ... and PostgreSQL for halfvec(d)
    blobs = [npy.process_bind_param(r, None) for r in E16]  # ... and any other dialect
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        from_text = np.stack([pg_res(t) for t in texts])
    from_list = np.stack([duck_res(v) for v in lists])
    from_blob = np.stack([npy.process_result_value(b, None) for b in blobs])
    counts = rng.integers(1, 5, size=64)
    row_chunk = np.repeat(np.arange(len(counts)), counts)[:n]
    np.savez_compressed(
        GOLDEN / "table_rows.npz", E16=E16, from_text=from_text, from_list=from_list, from_blob=from_blob, row_chunk=row_chunk,
        texts=np.frombuffer(json.dumps(texts).encode(), dtype=np.uint8),
        lists=np.frombuffer(json.dumps(lists).encode(), dtype=np.uint8),
        blobs=np.frombuffer(b"".join(blobs), dtype=np.uint8), blob_len=np.array([len(b) for b in blobs]))
    print("table_rows", from_text.dtype, from_list.dtype, from_blob.dtype, texts[0][:48])


def golden_rrf() -> None:
    """``reciprocal_rank_fusion`` (``_search.py:233-254``): the reference's own function text, compiled on its own
    (the module around it imports SQLAlchemy / sqlmodel), run on seeded rankings with ties and duplicates."""
    import ast
    from collections import defaultdict

    src = (REF / "_search.py").read_text()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "reciprocal_rank_fusion")
    ns: dict = {"defaultdict": defaultdict, "ChunkId": str}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(REF / "_search.py"), "exec"), ns)  # noqa: S102
    rrf = ns["reciprocal_rank_fusion"]
    rng = np.random.default_rng(21)
    cases = []
    for _ in range(12):
        R, L, universe = int(rng.integers(1, 4)), int(rng.integers(1, 40)), int(rng.integers(5, 60))
        rankings = [[int(x) for x in rng.permutation(universe)[: int(rng.integers(0, L + 1))]] for _ in range(R)]
        weights = [float(w) for w in rng.choice([1.0, 0.75, 0.25, 0.5], size=R)]
        k = int(rng.choice([60, 1, 10]))
        ids, scores = rrf([[str(c) for c in r] for r in rankings], k=k, weights=weights)
        cases.append({"rankings": rankings, "weights": weights, "k": k, "ids": [int(c) for c in ids], "scores": [float(x) for x in scores]})
    # the hybrid_search call shape: two rankings, weights 0.75 / 0.25 (_search.py:271-275)
    a, b = [int(x) for x in rng.permutation(50)[:20]], [int(x) for x in rng.permutation(50)[:20]]
    ids, scores = rrf([[str(c) for c in a], [str(c) for c in b]], weights=[0.75, 0.25])
    cases.append({"rankings": [a, b], "weights": [0.75, 0.25], "k": 60, "ids": [int(c) for c in ids], "scores": [float(x) for x in scores]})
    np.savez_compressed(GOLDEN / "rrf.npz", cases=np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8))
    print("rrf", len(cases), "cases")


if __name__ == "__main__":
    GOLDEN.mkdir(parents=True, exist_ok=True)
    golden_rrf()
    embed_mod, qa_mod = install_reference_stubs()
    golden_pool(embed_mod)
    golden_adapter(qa_mod)
    golden_table_rows(load_reference_typing())
