"""Stand-in for omegaconf (not installed here): yaml load + `${a.b}` interpolation
+ `${name:x,y}` resolvers, enough for agent_torch/core/helpers/general.py:101-129."""
import re

import yaml

_RESOLVERS = {}
_PAT = re.compile(r"\$\{([^${}]+)\}")


def _lookup(root, dotted):
    cur = root
    for part in dotted.split("."):
        cur = cur[int(part)] if isinstance(cur, list) else cur[part]
    return cur


def _coerce(tok):
    try:
        return yaml.safe_load(tok)
    except Exception:
        return tok


def _resolve_str(root, s):
    for _ in range(64):
        m = _PAT.search(s) if isinstance(s, str) else None
        if m is None:
            return s
        expr = m.group(1).strip()
        if ":" in expr and expr.split(":", 1)[0] in _RESOLVERS:
            name, args = expr.split(":", 1)
            val = _RESOLVERS[name](*[_coerce(_resolve_str(root, a.strip())) if isinstance(a, str) else a
                                     for a in args.split(",")])
        else:
            val = _resolve_str(root, _lookup(root, expr))
        if m.span() == (0, len(s)):
            s = val
        else:
            s = s[: m.start()] + str(val) + s[m.end():]
    raise ValueError("interpolation did not converge")


def _walk(root, node):
    if isinstance(node, dict):
        return {k: _walk(root, v) for k, v in node.items()}
    if isinstance(node, list):
        return [_walk(root, v) for v in node]
    if isinstance(node, str):
        return _resolve_str(root, node)
    return node


class OmegaConf:
    @staticmethod
    def register_new_resolver(name, fn):
        assert name not in _RESOLVERS, f"resolver '{name}' is already registered"
        _RESOLVERS[name] = fn

    @staticmethod
    def load(path):
        with open(path, "r", encoding="utf-8") as f:
            return yaml.safe_load(f)

    @staticmethod
    def to_object(cfg):
        return _walk(cfg, cfg)
