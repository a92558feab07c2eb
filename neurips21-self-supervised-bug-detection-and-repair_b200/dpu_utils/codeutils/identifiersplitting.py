"""snake_case / camelCase identifier splitting (dpu_utils.codeutils behaviour, restated).

Used by buglab/representations/data.py:113 (HasSubtoken vocabulary nodes) and the subtoken embedder."""
import re
from functools import lru_cache
from typing import List

# runs of: UPPER followed by lowers (Title), all-UPPER acronym (not followed by a lower), lowers, digits, other
_CAMEL = re.compile(r"[A-Z]+(?![a-z])|[A-Z][a-z]+|[a-z]+|[0-9]+|[^A-Za-z0-9]+")


def split_camelcase_unicode(part: str) -> List[str]:
    """The upstream character-class state machine (``str.isupper`` / ``isdigit`` / ``isalnum``), which also decides
    non-ASCII letters and digits (``naïve`` stays one part, ``Größe`` splits like an ASCII word would)."""
    if not part:
        return []
    result: List[str] = []
    current = part[0]
    prev_upper, prev_digit, prev_special = part[0].isupper(), part[0].isdigit(), not part[0].isalnum()
    for c in part[1:]:
        upper, digit, special = c.isupper(), c.isdigit(), not c.isalnum()
        if (digit and not prev_digit) or (upper and not prev_upper) or (special and not prev_special):
            result.append(current)
            current = c
        elif not upper and prev_upper and len(current) > 1:
            result.append(current[:-1])
            current = current[-1] + c
        elif (not digit and prev_digit) or (not special and prev_special):
            result.append(current)
            current = c
        else:
            current += c
        prev_upper, prev_digit, prev_special = upper, digit, special
    result.append(current)
    return result


def split_camelcase(part: str) -> List[str]:
    # the regex is the same state machine specialised to ASCII classes (checked against it in tests/test_host_cpu.py)
    return _CAMEL.findall(part) if part.isascii() else split_camelcase_unicode(part)


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
