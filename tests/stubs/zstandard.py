"""Stand-in for the `zstandard` package (absent from this image; game_runner.py:21 and neural_net.py:9 import it at
module top): ZstdCompressor / ZstdDecompressor / ZstdError over the system libzstd through ctypes, zlib when that is
missing too. Test infrastructure only — it lets the UNMODIFIED reference Python run against this repo's `alphazero`
module (last bullet)."""
import ctypes
import ctypes.util
import zlib


class ZstdError(Exception):
    pass


_lib = None
for _name in ("libzstd.so.1", ctypes.util.find_library("zstd")):
    if not _name:
        continue
    try:
        _lib = ctypes.CDLL(_name)
        break
    except OSError:
        continue
if _lib is not None:
    _lib.ZSTD_compressBound.restype = ctypes.c_size_t
    _lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    _lib.ZSTD_compress.restype = ctypes.c_size_t
    _lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    _lib.ZSTD_decompress.restype = ctypes.c_size_t
    _lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    _lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
    _lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    _lib.ZSTD_isError.restype = ctypes.c_uint
    _lib.ZSTD_isError.argtypes = [ctypes.c_size_t]


class ZstdCompressor:
    def __init__(self, level=3, threads=0, **kw):
        self.level = level

    def compress(self, data):
        data = bytes(data)
        if _lib is None:
            return b"ZLIB" + zlib.compress(data, 1)
        cap = _lib.ZSTD_compressBound(len(data))
        out = ctypes.create_string_buffer(cap)
        n = _lib.ZSTD_compress(out, cap, data, len(data), int(self.level))
        if _lib.ZSTD_isError(n):
            raise ZstdError("ZSTD_compress failed")
        return out.raw[:n]


class ZstdDecompressor:
    def __init__(self, **kw):
        pass

    def decompress(self, data, max_output_size=0):
        data = bytes(data)
        if data[:4] == b"ZLIB":
            return zlib.decompress(data[4:])
        if _lib is None:
            raise ZstdError("no libzstd")
        size = _lib.ZSTD_getFrameContentSize(data, len(data))
        if size in (2 ** 64 - 1, 2 ** 64 - 2):
            size = max_output_size or 64 * len(data)
        out = ctypes.create_string_buffer(max(1, size))
        n = _lib.ZSTD_decompress(out, max(1, size), data, len(data))
        if _lib.ZSTD_isError(n):
            raise ZstdError("ZSTD_decompress failed")
        return out.raw[:n]
