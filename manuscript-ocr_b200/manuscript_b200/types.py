"""Word / Block / Page: the result types at the drop-in boundary (reference detectors/_types.py:5-33).
Same field names, types and validation, so objects built here are interchangeable with the reference's."""
from typing import List, Optional, Tuple

from pydantic import BaseModel, Field


class Word(BaseModel):
    polygon: List[Tuple[float, float]] = Field(..., description="vertices (x, y) of the region")
    detection_confidence: float = Field(..., ge=0.0, le=1.0)
    text: Optional[str] = Field(None)
    recognition_confidence: Optional[float] = Field(None, ge=0.0, le=1.0)


class Block(BaseModel):
    words: List[Word]


class Page(BaseModel):
    blocks: List[Block]
