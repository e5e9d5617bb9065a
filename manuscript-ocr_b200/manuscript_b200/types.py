"""Result types at the drop-in boundary.

Field-compatible with the reference's `Word` / `Block` / `Page` (detectors/_types.py:5-33): same names, same value
types, the same 0..1 validation of the two confidences -- objects built here can be handed to the reference's
`visualize_page` / `Pipeline.get_text` and vice versa.
"""
from typing import Annotated, List, Optional, Tuple

from pydantic import BaseModel, Field

Point = Tuple[float, float]                                # (x, y) in pixels of the original image
Confidence = Annotated[float, Field(ge=0.0, le=1.0)]       # _types.py:9-11, 15-17


class Word(BaseModel):
    """One detected text region; `text` and `recognition_confidence` are filled in by the pipeline."""

    polygon: List[Point]
    detection_confidence: Confidence
    text: Optional[str] = None
    recognition_confidence: Optional[Confidence] = None


class Block(BaseModel):
    """A group of words (the detector emits one block per page)."""

    words: List[Word]


class Page(BaseModel):
    """All blocks of one page image."""

    blocks: List[Block]
