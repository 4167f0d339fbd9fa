"""ctypes binding for the CPU oracle (oracle/libzpaq_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product (zpaq_v_b200 / libzpaqgpu) never imports this module.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("ZPAQ_ORACLE_SO")   # another build of the same sources (tests/test_oracle_ubsan.py)
    if not path:
        path = os.path.join(ORACLE_DIR, "libzpaq_oracle.so")
        src = os.path.join(ORACLE_DIR, "zpaq_oracle.c")
        if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
            build()
    L = C.CDLL(path)
    u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    L.zo_squash_table.restype = C.POINTER(C.c_int32)
    L.zo_stretch_table.restype = C.POINTER(C.c_int32)
    L.zo_dt_table.restype = C.POINTER(C.c_int32)
    L.zo_dt2k_table.restype = C.POINTER(C.c_int32)
    L.zo_state_table.restype = u8p
    L.zo_level_header.argtypes = [C.c_int, C.c_char_p, C.c_int]
    L.zo_sha1.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    L.zo_raw_encode.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.c_int, C.POINTER(u8p)]
    L.zo_raw_encode.restype = C.c_size_t
    L.zo_raw_decode.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(u8p)]
    L.zo_raw_decode.restype = C.c_size_t
    L.zo_compress_block.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t, C.c_char_p,
                                    C.c_char_p, C.POINTER(u8p)]
    L.zo_compress_block.restype = C.c_size_t
    L.zo_decompress_archive.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(u8p), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int)]
    L.zo_decompress_archive.restype = C.c_size_t
    for f in (L.zo_compress_blocks_mt,):
        f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p,
                      u64p, C.c_int]
    L.zo_decompress_blocks_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64,
                                          C.c_void_p, u64p, C.c_int]
    # object API
    L.zo_compressor_new.restype = C.c_void_p
    L.zo_compressor_free.argtypes = [C.c_void_p]
    L.zo_compressor_set_input.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.zo_compressor_output.argtypes = [C.c_void_p]
    L.zo_compressor_output.restype = C.POINTER(_Buf)
    L.zo_compressor_start_block.argtypes = [C.c_void_p, C.c_int]
    L.zo_compressor_start_block_header.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.zo_compressor_start_segment.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.zo_compressor_compress.argtypes = [C.c_void_p, C.c_int]
    L.zo_compressor_end_segment.argtypes = [C.c_void_p]
    L.zo_compressor_end_block.argtypes = [C.c_void_p]
    L.zo_decompresser_new.restype = C.c_void_p
    L.zo_decompresser_free.argtypes = [C.c_void_p]
    L.zo_decompresser_set_input.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.zo_decompresser_input_pos.argtypes = [C.c_void_p]
    L.zo_decompresser_input_pos.restype = C.c_size_t
    L.zo_decompresser_output.argtypes = [C.c_void_p]
    L.zo_decompresser_output.restype = C.POINTER(_Buf)
    for name in ("find_block", "find_filename", "last_sha1_ok"):
        getattr(L, "zo_decompresser_" + name).argtypes = [C.c_void_p]
    L.zo_decompresser_filename.argtypes = [C.c_void_p]
    L.zo_decompresser_filename.restype = C.c_char_p
    L.zo_decompresser_comment.argtypes = [C.c_void_p]
    L.zo_decompresser_comment.restype = C.c_char_p
    L.zo_decompresser_decompress.argtypes = [C.c_void_p, C.c_int]
    L.zo_decompresser_read_segment_end.argtypes = [C.c_void_p]
    L.zo_fragment.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
    L.zo_fragment.restype = C.c_size_t
    L.zo_jidac_fragment.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long,
                                    C.POINTER(C.c_long)]
    L.zo_jidac_fragment.restype = C.c_long
    L.zo_jidac_add.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_char_p,
                               C.c_void_p, C.c_int, C.POINTER(u8p)]
    L.zo_jidac_add.restype = C.c_size_t
    _LIB = L
    return L


class _Buf(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_size_t), ("cap", C.c_size_t)]


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _take(ptr, n):
    out = C.string_at(ptr, n) if n else b""
    _libc.free(ptr)
    return out


def squash_table():
    return list(lib().zo_squash_table()[:4096])


def stretch_table():
    return list(lib().zo_stretch_table()[:32768])


def state_table():
    return bytes(lib().zo_state_table()[:1024])


def level_header(level):
    b = C.create_string_buffer(128)
    n = lib().zo_level_header(level, b, 128)
    return b.raw[:n]


def sha1(data):
    out = C.create_string_buffer(20)
    lib().zo_sha1(bytes(data), len(data), out)
    return out.raw


def raw_encode(hdr, data, with_pp=False):
    p = C.POINTER(C.c_uint8)()
    n = lib().zo_raw_encode(bytes(hdr), len(hdr), bytes(data), len(data), int(with_pp), C.byref(p))
    return _take(p, n)


def raw_decode(hdr, code):
    p = C.POINTER(C.c_uint8)()
    n = lib().zo_raw_decode(bytes(hdr), len(hdr), bytes(code), len(code), C.byref(p))
    return _take(p, n)


def compress_block(level, data, filename="", comment="", header=None):
    p = C.POINTER(C.c_uint8)()
    hdr = bytes(header) if header is not None else None
    n = lib().zo_compress_block(level, hdr, len(hdr) if hdr else 0, bytes(data), len(data),
                                filename.encode(), comment.encode(), C.byref(p))
    return _take(p, n)


def decompress_archive(arc):
    p = C.POINTER(C.c_uint8)()
    segs, bad = C.c_int(0), C.c_int(0)
    n = lib().zo_decompress_archive(bytes(arc), len(arc), C.byref(p), C.byref(segs), C.byref(bad))
    return _take(p, n), segs.value, bad.value


class Compressor:
    """Object API of the oracle, method for method the V Compressor (compressor.v:33-413)."""

    def __init__(self):
        self._h = lib().zo_compressor_new()
        self._keep = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().zo_compressor_free(self._h)
            self._h = None

    def set_input(self, data):
        self._keep = bytes(data)
        lib().zo_compressor_set_input(self._h, self._keep, len(self._keep))

    def start_block(self, level):
        lib().zo_compressor_start_block(self._h, level)

    def start_block_header(self, hdr):
        lib().zo_compressor_start_block_header(self._h, bytes(hdr), len(hdr))

    def start_segment(self, filename="", comment=""):
        lib().zo_compressor_start_segment(self._h, filename.encode(), comment.encode())

    def compress(self, n):
        return bool(lib().zo_compressor_compress(self._h, n))

    def end_segment(self):
        lib().zo_compressor_end_segment(self._h)

    def end_block(self):
        lib().zo_compressor_end_block(self._h)

    def output(self):
        b = lib().zo_compressor_output(self._h).contents
        return C.string_at(b.data, b.len) if b.len else b""


class Decompresser:
    """Object API of the oracle, method for method the V Decompresser (decompressor.v:187-640)."""

    def __init__(self):
        self._h = lib().zo_decompresser_new()
        self._keep = None

    def __del__(self):
        if getattr(self, "_h", None):
            lib().zo_decompresser_free(self._h)
            self._h = None

    def set_input(self, data):
        self._keep = bytes(data)
        lib().zo_decompresser_set_input(self._h, self._keep, len(self._keep))

    def find_block(self):
        return bool(lib().zo_decompresser_find_block(self._h))

    def find_filename(self):
        return bool(lib().zo_decompresser_find_filename(self._h))

    def get_filename(self):
        return lib().zo_decompresser_filename(self._h).decode("latin1")

    def get_comment(self):
        return lib().zo_decompresser_comment(self._h).decode("latin1")

    def decompress(self, n=-1):
        return bool(lib().zo_decompresser_decompress(self._h, n))

    def read_segment_end(self):
        lib().zo_decompresser_read_segment_end(self._h)

    def last_sha1_ok(self):
        return lib().zo_decompresser_last_sha1_ok(self._h)

    def input_pos(self):
        return lib().zo_decompresser_input_pos(self._h)

    def output(self):
        b = lib().zo_decompresser_output(self._h).contents
        return C.string_at(b.data, b.len) if b.len else b""


# ---- jidac front end (oracle/jidac_oracle.c) ----
class _Frag(C.Structure):
    _fields_ = [("off", C.c_uint64), ("len", C.c_uint64), ("file", C.c_uint32), ("id", C.c_uint32),
                ("stored", C.c_uint32), ("sha1", C.c_uint8 * 20)]


def fragment_ends(data, fragment):
    data = bytes(data)
    n = lib().zo_fragment(data, len(data), fragment, None, 0)
    ends = (C.c_uint64 * max(n, 1))()
    lib().zo_fragment(data, len(data), fragment, ends, n)
    return list(ends[:n])


def _ranges(files):
    n = len(files)
    off = (C.c_uint64 * (n + 1))()
    pos = 0
    for i, b in enumerate(files):
        off[i] = pos
        pos += len(b)
    off[n] = pos
    return n, b"".join(bytes(b) for b in files), off


def jidac_fragment(files, fragment, dedup):
    n, data, off = _ranges(files)
    ns = C.c_long(0)
    cnt = lib().zo_jidac_fragment(data, off, n, fragment, int(dedup), None, 0, C.byref(ns))
    frs = (_Frag * max(cnt, 1))()
    lib().zo_jidac_fragment(data, off, n, fragment, int(dedup), frs, cnt, C.byref(ns))
    return [dict(off=f.off, len=f.len, file=f.file, id=f.id, stored=f.stored, sha1=bytes(f.sha1))
            for f in frs[:cnt]], ns.value


def jidac_add(names, files, date, level=0, fragment=-1, dedup=False, block_bytes=0):
    n, data, off = _ranges(files)
    arr = (C.c_char_p * max(n, 1))()
    for i, x in enumerate(names):
        arr[i] = x.encode() if isinstance(x, str) else x
    p = C.POINTER(C.c_uint8)()
    ln = lib().zo_jidac_add(date, level, fragment, int(dedup), block_bytes, arr, data, off, n, C.byref(p))
    return _take(p, ln)
