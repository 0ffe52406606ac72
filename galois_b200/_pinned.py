"""
Page-locked host arrays for the `*_host` pipelines.

`cudaHostAlloc` costs milliseconds per call, so blocks are recycled through a size-bucketed free
list; an ndarray handed out by ``empty`` returns its block to the pool when it is garbage-collected.
This is memory plumbing only — no arithmetic happens here.
"""

from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _lib

_MIN_PINNED = 1 << 16  # smaller arrays use ordinary NumPy memory
_MAX_CACHED = 16 << 30  # bytes kept in the free list

_lock = threading.Lock()
_free: dict[int, list[int]] = {}
_cached_bytes = 0


class _Block:
    __slots__ = ("ptr", "size")

    def __init__(self, ptr: int, size: int):
        self.ptr, self.size = ptr, size

    def __del__(self):
        global _cached_bytes
        try:
            with _lock:
                if _cached_bytes + self.size <= _MAX_CACHED:
                    _free.setdefault(self.size, []).append(self.ptr)
                    _cached_bytes += self.size
                    return
            _lib.load().gfb_host_free(ctypes.c_void_p(self.ptr))
        except Exception:  # interpreter shutdown
            pass


def _bucket(nbytes: int) -> int:
    size = 1 << 16
    while size < nbytes:
        size <<= 1
    return size


def empty(shape, dtype) -> np.ndarray:
    """An uninitialised C-contiguous ndarray, in pinned memory when it is large enough."""
    global _cached_bytes
    dtype = np.dtype(dtype)
    shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(shape)
    count = 1
    for s in shape:
        count *= int(s)
    nbytes = count * dtype.itemsize
    if nbytes < _MIN_PINNED or dtype == np.object_:
        return np.empty(shape, dtype=dtype)
    size = _bucket(nbytes)
    ptr = None
    with _lock:
        lst = _free.get(size)
        if lst:
            ptr = lst.pop()
            _cached_bytes -= size
    if ptr is None:
        ptr = _lib.lib().gfb_host_alloc(size)
        if not ptr:
            return np.empty(shape, dtype=dtype)  # pinning failed (ulimit); pageable memory still works
    buf = (ctypes.c_char * nbytes).from_address(ptr)
    buf._gfb_block = _Block(ptr, size)  # ties the block's lifetime to the buffer (the ndarray's base)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def trim() -> None:
    """Release every cached block."""
    global _cached_bytes
    with _lock:
        blocks = [(p, s) for s, lst in _free.items() for p in lst]
        _free.clear()
        _cached_bytes = 0
    for p, _ in blocks:
        _lib.load().gfb_host_free(ctypes.c_void_p(p))
