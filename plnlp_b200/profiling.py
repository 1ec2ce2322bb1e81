"""Opt-in per-kernel timing with CUDA events on the launching stream (used by bench.py's
instrumented pass; zero cost when disabled)."""
from __future__ import annotations

import contextlib

import torch

_enabled = False
_spans = []          # (name, start_event, end_event, bytes, flops)
_NULL = contextlib.nullcontext()


class _Span:
    __slots__ = ("name", "nbytes", "flops", "e0")

    def __init__(self, name, nbytes, flops):
        self.name, self.nbytes, self.flops = name, nbytes, flops

    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e0.record()
        return self

    def __exit__(self, *a):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        _spans.append((self.name, self.e0, e1, self.nbytes, self.flops))


def span(name, nbytes=0, flops=0):
    return _Span(name, nbytes, flops) if _enabled else _NULL


def enabled():
    return _enabled


def enable():
    global _enabled
    _spans.clear()
    _enabled = True


def disable():
    """-> {name: {n, ms, bytes, flops}} (synchronises)"""
    global _enabled
    _enabled = False
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, nbytes, flops in _spans:
        s = out.setdefault(name, {"n": 0, "ms": 0.0, "bytes": 0, "flops": 0})
        s["n"] += 1
        s["ms"] += e0.elapsed_time(e1)
        s["bytes"] += nbytes
        s["flops"] += flops
    _spans.clear()
    return out
