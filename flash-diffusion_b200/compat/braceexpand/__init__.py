"""`braceexpand.braceexpand` for the shard URLs of the example yamls ("/data/{000000..000010}.tar", "a{b,c}d")."""
import re

_RANGE = re.compile(r"^(-?\d+)\.\.(-?\d+)(?:\.\.(-?\d+))?$")


def braceexpand(pattern):
    i = pattern.find("{")
    if i < 0:
        yield pattern
        return
    depth, j = 0, i
    for j in range(i, len(pattern)):
        depth += pattern[j] == "{"
        depth -= pattern[j] == "}"
        if depth == 0:
            break
    else:
        yield pattern
        return
    head, body, tail = pattern[:i], pattern[i + 1:j], pattern[j + 1:]
    m = _RANGE.match(body)
    if m:
        a, b = int(m.group(1)), int(m.group(2))
        step = abs(int(m.group(3))) if m.group(3) else 1
        width = max(len(m.group(1)), len(m.group(2))) if (m.group(1).startswith("0") or m.group(2).startswith("0")) else 0
        seq = range(a, b + 1, step) if a <= b else range(a, b - 1, -step)
        parts = [str(v).zfill(width) for v in seq]
    else:
        parts, depth, cur = [], 0, ""
        for ch in body:
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
                continue
            depth += ch == "{"
            depth -= ch == "}"
            cur += ch
        parts.append(cur)
        if len(parts) == 1:
            parts = ["{" + body + "}"]
    for p in parts:
        for mid in braceexpand(p):
            for rest in braceexpand(tail):
                yield head + mid + rest
