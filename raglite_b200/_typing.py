"""Types mirroring ``raglite._typing`` (reference ``_typing.py:20-54``) without its SQLAlchemy baggage."""

from __future__ import annotations

from collections.abc import Mapping
from typing import TYPE_CHECKING, Any, Literal, Protocol

import numpy as np

if TYPE_CHECKING:
    from ._config import RAGLiteConfig
    from ._index import Chunk
    from ._search import ChunkSpan

ChunkId = str
DocumentId = str
IndexId = str

DistanceMetric = Literal["cosine", "dot", "l1", "l2"]

MetadataValue = str | int | float | bool
MetadataFilter = Mapping[str, list[MetadataValue] | MetadataValue]

FloatMatrix = np.ndarray[tuple[int, int], np.dtype[np.floating[Any]]]
FloatVector = np.ndarray[tuple[int], np.dtype[np.floating[Any]]]
IntVector = np.ndarray[tuple[int], np.dtype[np.intp]]


class BasicSearchMethod(Protocol):
    """``_typing.py:35-43``."""

    def __call__(
        self, query: str, *, num_results: int, metadata_filter: MetadataFilter | None = None,
        config: "RAGLiteConfig | None" = None,
    ) -> tuple[list[ChunkId], list[float]]: ...


class SearchMethod(Protocol):
    """``_typing.py:46-54``."""

    def __call__(
        self, query: str, *, num_results: int, metadata_filter: MetadataFilter | None = None,
        config: "RAGLiteConfig | None" = None,
    ) -> "tuple[list[ChunkId], list[float]] | list[Chunk] | list[ChunkSpan]": ...
