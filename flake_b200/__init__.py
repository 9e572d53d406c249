"""flake_b200 -- B200-native (sm_100a) implementation of Flake's FLAC encoding hot path
behind libflake's C API.  See DESIGN.md.

Python here is only the ctypes mirror of the C API (api.py), the synthetic PCM
generator (synth.py) and the build recipes (build.py); the product is
flake_b200/lib/libflake.so.
"""
from .api import (DEFAULT_LIB, PCM_S16LE, PCM_S24LE, PCM_S32, PCM_S8, EncodedStream, Encoder,  # noqa: F401
                  FlakeLibraryError, encode_batch, encode_per_block, load_library)

__version__ = "0.1.0"
