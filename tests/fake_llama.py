"""A deterministic stand-in for ``llama_cpp.Llama`` in embedding mode (test infrastructure).

The reference's late-chunking code (``raglite/_embed.py:16-141``) only needs ``n_ctx()``,
``n_batch``, ``tokenize``, ``detokenize`` and ``embed`` from the embedder.  This fake implements those
with a regex tokenizer and hash-seeded token embeddings, so that the *reference's own* pooling code
can be run offline to produce golden vectors (``tools/make_golden_from_reference.py``) and the
product path can be fed the very same token embeddings.
"""

from __future__ import annotations

import re
import zlib

import numpy as np

_TOKEN_RE = re.compile(r"\n?⊕| ?⊕|\s?\w+|\s|[^\w\s]", re.UNICODE)


class FakeLlama:
    def __init__(self, n_ctx: int = 64, dim: int = 32, extra_tokens: int = 2, seed: int = 0):
        self._n_ctx = n_ctx
        self.n_batch = n_ctx
        self.dim = dim
        self.extra_tokens = extra_tokens  # BOS/EOS-like rows that llama.cpp adds to the output.
        self.seed = seed
        self._vocab: dict[int, str] = {}

    def n_ctx(self) -> int:
        return self._n_ctx

    def n_embd(self) -> int:
        return self.dim

    def tokenize(self, text: bytes, add_bos: bool = False, special: bool = False) -> list[int]:  # noqa: ARG002
        out = []
        for piece in _TOKEN_RE.findall(text.decode()):
            tok = 10 + zlib.crc32(piece.encode()) % 100_000
            self._vocab[tok] = piece
            out.append(tok)
        return out

    def detokenize(self, tokens: list[int]) -> bytes:
        return "".join(self._vocab.get(t, "") for t in tokens).encode()

    def _embed_one(self, text: str) -> list[list[float]]:
        toks = self.tokenize(text.encode())
        n = min(len(toks), self.n_batch) + self.extra_tokens
        rows = np.empty((n, self.dim), dtype=np.float32)
        for i in range(n):
            tok = toks[i - 1] if 0 < i <= len(toks) else 1
            rng = np.random.default_rng([self.seed, tok, i])
            rows[i] = rng.standard_normal(self.dim).astype(np.float32) + np.float32(0.25)
        return rows.tolist()  # llama-cpp-python returns nested Python lists of floats.

    def embed(self, text):  # noqa: ANN001, ANN201
        if isinstance(text, str):
            return self._embed_one(text)
        return [self._embed_one(t) for t in text]


def make_sentences(n: int, seed: int = 0) -> list[str]:
    """Synthetic 'sentences' of varying length; they concatenate to the document (as in RAGLite)."""
    rng = np.random.default_rng(seed)
    words = ["alpha", "beta", "gamma", "delta", "light", "clock", "rod", "frame", "event", "time",
             "x", "of", "the", "simultaneous", "observer", "velocity", "é", "naïve"]
    out = []
    for _ in range(n):
        k = int(rng.integers(1, 14))
        s = " ".join(rng.choice(words, size=k)) + rng.choice([". ", "? ", ".\n\n", "; "])
        out.append(str(s))
    return out
