"""Pin ``oracle.pool`` / ``oracle.adapter`` against golden vectors produced by the reference's own
code (``tools/make_golden_from_reference.py``)."""

from __future__ import annotations

import json

import numpy as np
import pytest
from fake_llama import FakeLlama

from oracle import adapter as oadapter
from oracle import pool as opool

CASES = ["pool_small", "pool_multi", "pool_nonorm", "pool_wide"]


def load_case(golden_dir, name):
    z = np.load(golden_dir / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z["late_chunking"], z["simple"]


@pytest.mark.parametrize("name", CASES)
def test_late_chunking_matches_reference(golden_dir, name):
    meta, want, _ = load_case(golden_dir, name)
    llm = FakeLlama(n_ctx=meta["n_ctx"], dim=meta["dim"], seed=meta["seed"])
    got = opool.embed_with_llama(meta["sentences"], llm, normalize=meta["normalize"])
    assert got.dtype == np.float16 and got.shape == want.shape
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16))
    if meta["normalize"]:  # reference tests/test_embed.py:26
        assert np.allclose(np.linalg.norm(got.astype(np.float64), axis=1), 1.0, rtol=1e-3)


@pytest.mark.parametrize("name", CASES)
def test_simple_pool_matches_reference(golden_dir, name):
    meta, _, want = load_case(golden_dir, name)
    llm = FakeLlama(n_ctx=meta["n_ctx"], dim=meta["dim"], seed=meta["seed"])
    got = opool.simple_pool(llm.embed(meta["sentences"][:7]), normalize=meta["normalize"])
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16))


def test_largest_remainder_sums_and_bounds():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(1, 20))
        toks = rng.integers(1, 40, size=n)
        rows = int(rng.integers(n, 600))
        sizes = opool.largest_remainder_sizes(rows, toks)
        assert sizes.sum() == rows
        frac = rows * toks / toks.sum()
        assert np.all(sizes >= np.floor(frac)) and np.all(sizes <= np.floor(frac) + 1)


def test_optimize_query_target_matches_reference(golden_dir):
    z = np.load(golden_dir / "adapter_target.npz")
    for i in range(3):
        t = oadapter.optimize_query_target(z[f"q{i}"], z[f"P{i}"], z[f"N{i}"], alpha=0.05)
        assert t.dtype == z[f"t{i}"].dtype
        assert np.array_equal(t, z[f"t{i}"])


def test_fit_query_adapter_is_orthogonal():
    rng = np.random.default_rng(3)
    Q = rng.standard_normal((20, 8))
    T = Q + 0.1 * rng.standard_normal((20, 8))
    A = oadapter.fit_query_adapter(Q, T, "cosine")
    assert np.allclose(A @ A.T, np.eye(8), atol=1e-10)
    Ad = oadapter.fit_query_adapter(Q, T, "dot")
    assert np.isclose(np.linalg.norm(Ad, "fro"), np.sqrt(8))
