"""snake_case / camelCase identifier splitting (dpu_utils.codeutils behaviour, restated).

Used by buglab/representations/data.py:113 (HasSubtoken vocabulary nodes) and the subtoken embedder."""
import re
from functools import lru_cache
from typing import List

# runs of: UPPER followed by lowers (Title), all-UPPER acronym (not followed by a lower), lowers, digits, other
_CAMEL = re.compile(r"[A-Z]+(?![a-z])|[A-Z][a-z]+|[a-z]+|[0-9]+|[^A-Za-z0-9]+")


def split_camelcase(part: str) -> List[str]:
    return _CAMEL.findall(part)


@lru_cache(maxsize=200000)
def _split(identifier: str):
    parts: List[str] = []
    for piece in identifier.split("_"):
        if piece:
            parts.extend(s.lower() for s in split_camelcase(piece))
    if not parts:
        return (identifier,)
    return tuple(parts)


def split_identifier_into_parts(identifier: str) -> List[str]:
    """``fooBar_baz2`` -> ``['foo', 'bar', 'baz', '2']``; an identifier made only of underscores is returned as is."""
    return list(_split(identifier))
