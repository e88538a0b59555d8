"""audio_decoder_b200 — B200-native (sm_100a) implementation of BLAST's data-parallel hot path.

Host-side mirror of the reference's Rust surface over the C ABI in include/blast_cuda.h:
  file_parsing.{wav,aiff}.parse, AudioFile, DecodeError   (blast/src/file_parsing/)
All arithmetic runs in hand-written CUDA kernels inside libblast_cuda.so; importing a
submodule that needs the library raises ImportError if it has not been built.
"""
from . import _lib  # noqa: F401
from .context import Context, DevBuf, HostBuf, Event  # noqa: F401
from .errors import (BlastError, DecodeError, Io, UnsupportedFormat, UnexpectedEof, InvalidData,  # noqa: F401
                     ReferencePanic)
from . import file_parsing  # noqa: F401
from .file_parsing import AudioFile, DeviceTrack  # noqa: F401
from . import audio_processing  # noqa: F401
from . import blast_rand  # noqa: F401

__all__ = ["Context", "DevBuf", "HostBuf", "Event", "file_parsing", "AudioFile", "DeviceTrack", "BlastError",
           "DecodeError", "Io", "UnsupportedFormat", "UnexpectedEof", "InvalidData", "ReferencePanic"]
