"""DecodeError / DecodeResult mirror (blast/src/file_parsing/decode_helpers.rs:1-15)."""
from __future__ import annotations

from . import _lib


class BlastError(RuntimeError):
    """Any non-zero status from libblast_cuda."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[blast status {code}] {msg}")
        self.code = code
        self.msg = msg


class DecodeError(BlastError):
    """DecodeError enum of the reference; `variant` is the Rust variant name."""
    variant = "DecodeError"


class Io(DecodeError):
    variant = "Io"


class UnsupportedFormat(DecodeError):
    variant = "UnsupportedFormat"


class UnexpectedEof(DecodeError):
    variant = "UnexpectedEof"


class InvalidData(DecodeError):
    variant = "InvalidData"


class ReferencePanic(BlastError):
    """Inputs on which the reference panics (index out of bounds, usize underflow)."""


_BY_CODE = {
    _lib.ERR_IO: Io, _lib.ERR_UNSUPPORTED_FORMAT: UnsupportedFormat, _lib.ERR_UNEXPECTED_EOF: UnexpectedEof,
    _lib.ERR_INVALID_DATA: InvalidData, _lib.ERR_REF_PANIC: ReferencePanic,
}


def check(rc: int):
    if rc == _lib.OK:
        return
    msg = _lib.load().blast_last_error().decode(errors="replace")
    raise _BY_CODE.get(rc, BlastError)(rc, msg)
