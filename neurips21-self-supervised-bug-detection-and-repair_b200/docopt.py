"""A small docopt: parses ``Usage:`` + ``Options:`` docstrings of the shape used by
buglab/models/train.py:2-19 and evaluate.py:2-18 (long options with optional ``=<value>`` and
``[default: x]``, positional arguments, ``-h --help``).  The real ``docopt`` package is an unpinned
dependency of the reference (requirements.txt) that is not installable offline."""
import re
import sys
from typing import Dict, List, Optional

__all__ = ["docopt", "DocoptExit"]


class DocoptExit(SystemExit):
    pass


def _parse_options(doc: str):
    options: Dict[str, dict] = {}
    in_options = False
    for line in doc.splitlines():
        if re.match(r"^\s*options:\s*$", line, re.I):
            in_options = True
            continue
        if not in_options:
            continue
        stripped = line.strip()
        if not stripped.startswith("-"):
            continue
        spec, _, description = stripped.partition("  ")
        names = [n for n in re.split(r"[ ,]+", spec) if n.startswith("-")]
        takes_value = "=" in spec
        long_name = next((n.split("=")[0] for n in names if n.startswith("--")), names[0].split("=")[0])
        default = None
        m = re.search(r"\[default:\s*(.*?)\]", description)
        if m and takes_value:
            default = m.group(1)
        entry = dict(name=long_name, takes_value=takes_value, default=default if takes_value else False)
        for n in names:
            options[n.split("=")[0]] = entry
    return options


def _parse_positionals(doc: str) -> List[str]:
    m = re.search(r"usage:\s*(.*?)(?:\n\s*\n|\Z)", doc, re.I | re.S)
    if not m:
        return []
    first = m.group(1).strip().splitlines()[0]
    return [tok for tok in first.split()[1:] if re.fullmatch(r"[A-Z][A-Z0-9_]*|<[^>]+>", tok)]


def docopt(doc: str, argv: Optional[List[str]] = None, help: bool = True, version=None) -> Dict[str, object]:
    argv = list(sys.argv[1:] if argv is None else argv)
    options = _parse_options(doc)
    positionals = _parse_positionals(doc)
    result: Dict[str, object] = {}
    for entry in options.values():
        result[entry["name"]] = entry["default"]
    values: List[str] = []
    i = 0
    while i < len(argv):
        arg = argv[i]
        if arg == "--":
            values.extend(argv[i + 1:])
            break
        if arg.startswith("-") and arg != "-":
            name, eq, val = arg.partition("=")
            entry = options.get(name)
            if entry is None:
                raise DocoptExit(f"unknown option {name}\n{doc}")
            if entry["name"] == "--help" and help:
                print(doc.strip())
                raise SystemExit(0)
            if entry["takes_value"]:
                if not eq:
                    i += 1
                    if i >= len(argv):
                        raise DocoptExit(f"{name} requires a value\n{doc}")
                    val = argv[i]
                result[entry["name"]] = val
            else:
                result[entry["name"]] = True
        else:
            values.append(arg)
        i += 1
    if len(values) != len(positionals):
        raise DocoptExit(f"expected arguments {positionals}, got {values}\n{doc}")
    for name, val in zip(positionals, values):
        result[name] = val
    return result
