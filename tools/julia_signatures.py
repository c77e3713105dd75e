"""A small parser for Julia `function name(positional...; keywords...) [where T]` headers — enough for the reference's
solver files and for julia/gab1pde_dropin.jl.  Used by tests/golden/make_reference_signatures.py (extracts the reference's
call surface into a committed fixture) and by tests/test_julia_surface.py (checks the drop-in file against it)."""
from __future__ import annotations

import re


def strip_comments(src: str) -> str:
    out = []
    for line in src.splitlines():
        buf, in_str, i = [], False, 0
        while i < len(line):
            c = line[i]
            if c == '"':
                in_str = not in_str
            if c == "#" and not in_str:
                break
            buf.append(c)
            i += 1
        out.append("".join(buf))
    return "\n".join(out)


def split_top(s: str, sep: str):
    parts, depth, cur = [], 0, []
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(c)
    parts.append("".join(cur))
    return [p.strip() for p in parts if p.strip()]


def norm(expr: str) -> str:
    return re.sub(r"\s+", "", expr)


def parse_arg(a: str):
    """`name::Type=default` -> (name, type or None, default or None), white space removed."""
    default = None
    depth = 0
    for i, c in enumerate(a):
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        elif c == "=" and depth == 0 and a[i + 1:i + 2] != "=" and a[i - 1:i] not in ("=", "<", ">", "!"):
            default = norm(a[i + 1:])
            a = a[:i]
            break
    name, _, typ = a.partition("::")
    return norm(name), (norm(typ) or None), default


def signatures(src: str) -> dict:
    """{function name: [{"positional": [...], "keywords": [...], "where": bool}, ...]} for every `function f(...)`."""
    src = strip_comments(src)
    found = {}
    for m in re.finditer(r"\bfunction\s+([A-Za-z_][A-Za-z_0-9!]*)\s*\(", src):
        name, i = m.group(1), m.end()
        depth, j = 1, i
        while depth and j < len(src):
            depth += src[j] in "([{"
            depth -= src[j] in ")]}"
            j += 1
        inner = src[i:j - 1]
        tail = src[j:j + 20]
        halves = split_top(inner, ";")
        pos = split_top(halves[0], ",") if halves else []
        kws = split_top(halves[1], ",") if len(halves) > 1 else []
        found.setdefault(name, []).append({
            "positional": [list(parse_arg(a)) for a in pos],
            "keywords": [list(parse_arg(a)) for a in kws],
            "where": bool(re.match(r"\s*where\b", tail)),
        })
    return found
