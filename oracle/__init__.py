"""CPU oracle for the RAGLite retrieval hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy, the arithmetic of the reference hot path
(superlinear-ai/raglite @ 2069f8d) so that the CUDA product path in ``raglite_b200``
can be checked against it.  It is *not* part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  ``raglite_b200`` never imports ``oracle``.

Pinning status (see DESIGN.md "Oracle"):

* ``oracle.pool``          -- PINNED: checked against golden vectors produced by running the
  reference's own ``raglite._embed`` code (imported from /root/reference with its llama.cpp /
  LiteLLM imports stubbed by a deterministic fake token embedder; generator committed as
  ``tools/make_golden_from_reference.py``, fixtures in ``tests/golden/pool_*.npz``).
* ``oracle.adapter``       -- fit helper ``optimize_query_target`` PINNED the same way
  (``raglite._query_adapter._optimize_query_target`` is pure NumPy/SciPy).  The *apply*
  (``_search.py:58-62``) is a single NumPy expression; restated verbatim.
* ``oracle.vector_search`` -- PARITY UNPINNED: the per-vector distance, ORDER BY/LIMIT and
  GROUP BY/max run inside DuckDB (>=1.1.3, un-vendored, not installable here: no network).
  The restatement follows the reference's SQL call sites (``_search.py:65-79,143-153``,
  ``_typing.py:123-134``) and DuckDB's published semantics for ``array_cosine_distance`` /
  ``array_negative_inner_product`` / ``array_distance``; the reference's own tests hold no
  numeric golden vector for it.
* ``oracle.rerank``        -- PARITY UNPINNED vs FlashRank's ONNX model (weights unavailable
  offline); architectural parity against ``transformers.BertForSequenceClassification`` fp32.
"""
