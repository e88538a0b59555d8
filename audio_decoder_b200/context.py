"""Context / buffer plumbing over the C ABI (include/blast_cuda.h, lifecycle + memory sections)."""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from .errors import check


class DevBuf:
    """A device allocation owned by a Context (blast_dev_alloc)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(ctx.lib.blast_dev_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value
        ctx._bufs.add(self)               # weakly: a context that is closed first frees what is still allocated

    def free(self):
        if self.ptr is not None and self.ctx.h:
            check(self.ctx.lib.blast_dev_free(self.ctx.h, self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def upload(self, arr: np.ndarray, offset: int = 0):
        a = np.ascontiguousarray(arr)
        assert offset + a.nbytes <= self.nbytes
        check(self.ctx.lib.blast_memcpy_h2d(self.ctx.h, self.ptr + offset, a.ctypes.data, a.nbytes))
        self.ctx.sync()          # `a` may be a temporary
        return self

    def download(self, dtype, count: int, offset: int = 0) -> np.ndarray:
        out = np.empty(count, dtype=dtype)
        assert offset + out.nbytes <= self.nbytes
        check(self.ctx.lib.blast_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr + offset, out.nbytes))
        self.ctx.sync()
        return out

    def zero(self):
        check(self.ctx.lib.blast_memset_dev(self.ctx.h, self.ptr, 0, self.nbytes))


class HostBuf:
    """Pinned host memory (blast_host_alloc) exposed as a numpy array."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(ctx.lib.blast_host_alloc(ctx.h, self.nbytes, C.byref(p)))
        self.ptr = p.value
        self.u8 = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(max(1, self.nbytes),))[:self.nbytes]

    def view(self, dtype, count=None, offset=0) -> np.ndarray:
        item = np.dtype(dtype).itemsize
        if count is None:
            count = (self.nbytes - offset) // item
        return self.u8[offset:offset + count * item].view(dtype)

    def free(self):
        if self.ptr is not None and self.ctx.h:
            self.u8 = None
            check(self.ctx.lib.blast_host_free(self.ctx.h, self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Event:
    def __init__(self, ctx: "Context"):
        self.ctx = ctx
        p = C.c_void_p()
        check(ctx.lib.blast_event_create(ctx.h, C.byref(p)))
        self.h = p.value

    def record(self):
        check(self.ctx.lib.blast_event_record(self.ctx.h, self.h))
        return self

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        check(self.ctx.lib.blast_event_elapsed_ms(self.h, stop.h, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.blast_event_destroy(self.h)
        except Exception:
            pass
        self.h = None


class Context:
    """One GPU, one stream (blast_ctx).  Raises errors.BlastError(ERR_NO_DEVICE) without a B200."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = _lib.load()
        self._bufs = weakref.WeakSet()
        p = C.c_void_p()
        check(self.lib.blast_ctx_create(C.byref(p), device))
        self.h = p.value
        if stream is not None:
            check(self.lib.blast_ctx_set_stream(self.h, stream))

    def close(self):
        if getattr(self, "h", None):
            for b in list(getattr(self, "_bufs", ())):      # device buffers that outlive the context would never be freed
                try:
                    b.free()
                except Exception:
                    pass
            self.lib.blast_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def sync(self):
        check(self.lib.blast_ctx_sync(self.h))

    def trim(self):
        """release the context's grow-only device scratch"""
        check(self.lib.blast_ctx_trim(self.h))

    @property
    def device(self) -> int:
        return self.lib.blast_ctx_device(self.h)

    @property
    def sm_count(self) -> int:
        return self.lib.blast_ctx_sm_count(self.h)

    @property
    def launch_count(self) -> int:
        return self.lib.blast_ctx_launch_count(self.h)

    @property
    def stream(self) -> int:
        return self.lib.blast_ctx_stream(self.h) or 0

    def alloc(self, nbytes: int) -> DevBuf:
        return DevBuf(self, nbytes)

    def pinned(self, nbytes: int) -> HostBuf:
        return HostBuf(self, nbytes)

    def event(self) -> Event:
        return Event(self)

    def to_device(self, arr: np.ndarray) -> DevBuf:
        a = np.ascontiguousarray(arr)
        return self.alloc(max(a.nbytes, 1)).upload(a)


class BorrowedContext(Context):
    """A blast_ctx owned by somebody else (a blast_group member): same methods, never destroyed from here."""

    def __init__(self, lib, handle: int):
        self.lib = lib
        self.h = handle
        self._bufs = weakref.WeakSet()

    def close(self):
        self.h = None
