"""ctypes binding of libzpaqgpu (include/zpaqgpu.h).

The library is the product; this module only marshals pointers and sizes.  There is no Python or
CPU implementation behind it: if the shared library or a CUDA device is missing, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzpaqgpu.so")

OK, E_NODEVICE, E_CUDA, E_NOSPACE, E_ARG, E_FORMAT, E_UNSUPPORTED, E_STATE, E_NOMEM = 0, -1, -2, -3, -4, -5, -6, -7, -8
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_CHAIN = 0, 1, 2
TABLES_AUTO, TABLES_DENSE, TABLES_PAGED = 0, 1, 2

EXPORTS = [
    "zpaqgpu_init", "zpaqgpu_destroy", "zpaqgpu_strerror", "zpaqgpu_last_error", "zpaqgpu_set_kernel",
    "zpaqgpu_set_table_mode",
    "zpaqgpu_set_workspace_limit", "zpaqgpu_set_stream", "zpaqgpu_level_header", "zpaqgpu_tables",
    "zpaqgpu_compress_blocks", "zpaqgpu_compress_blocks_header", "zpaqgpu_compress_blocks_dev",
    "zpaqgpu_find_blocks", "zpaqgpu_decompress_archive", "zpaqgpu_decompress_blocks_dev",
    "zpaqgpu_block_begin", "zpaqgpu_block_begin_header", "zpaqgpu_segment_begin", "zpaqgpu_segment_write",
    "zpaqgpu_segment_end", "zpaqgpu_block_end", "zpaqgpu_last_stats", "zpaqgpu_describe_model",
    "zpaqgpu_jidac_fragment", "zpaqgpu_jidac_add", "zpaqgpu_jidac_extract", "zpaqgpu_jidac_last_stats",
    "zpaqgpu_stream_batch", "zpaqgpu_block_end_queue", "zpaqgpu_queued", "zpaqgpu_flush",
    "zpaqgpu_multi_init", "zpaqgpu_multi_destroy", "zpaqgpu_multi_device_count", "zpaqgpu_multi_ctx",
    "zpaqgpu_multi_last_error", "zpaqgpu_multi_compress_blocks", "zpaqgpu_multi_decompress_archive",
    "zpaqgpu_multi_last_stats", "zpaqgpu_multi_jidac_add",
]


class ModelInfo(C.Structure):
    _fields_ = [("n", C.c_int32), ("cend", C.c_int32), ("hbegin", C.c_int32), ("hend", C.c_int32),
                ("hsize", C.c_int32), ("is_chain", C.c_int32), ("n_isse", C.c_int32), ("has_mix2", C.c_int32),
                ("ctx_mode", C.c_int32), ("n_hash", C.c_int32), ("workspace_bytes", C.c_uint64),
                ("hash_table_bytes", C.c_uint64)]


class Segment(C.Structure):
    _fields_ = [("block_start", C.c_uint64), ("block_end", C.c_uint64), ("name_off", C.c_uint64),
                ("comment_off", C.c_uint64), ("out_off", C.c_uint64), ("out_len", C.c_uint64),
                ("block_index", C.c_int32), ("sha1_ok", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("init_ms", C.c_float), ("codec_ms", C.c_float), ("sha1_ms", C.c_float), ("pack_ms", C.c_float),
                ("h2d_ms", C.c_float), ("d2h_ms", C.c_float), ("launches", C.c_int32),
                ("codec_launches", C.c_int32), ("waves", C.c_int32), ("retries", C.c_int32), ("kernel", C.c_int32),
                ("warps_per_cta", C.c_int32), ("workspace_bytes_per_block", C.c_uint64),
                ("pool_bytes_used", C.c_uint64), ("paged", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MultiStats(C.Structure):
    _fields_ = [("device", C.c_int32), ("first_unit", C.c_int32), ("n_units", C.c_int32),
                ("fallback_single", C.c_int32), ("stage_ms", C.c_float), ("fetch_ms", C.c_float), ("stats", Stats)]


class JidacOpts(C.Structure):
    _fields_ = [("date", C.c_int64), ("level", C.c_int32), ("fragment", C.c_int32), ("dedup", C.c_int32),
                ("reserved", C.c_int32), ("block_bytes", C.c_uint64)]


class Fragment(C.Structure):
    _fields_ = [("off", C.c_uint64), ("len", C.c_uint64), ("file", C.c_uint32), ("id", C.c_uint32),
                ("stored", C.c_uint32), ("sha1", C.c_uint8 * 20)]


class JidacFile(C.Structure):
    _fields_ = [("name_off", C.c_uint64), ("out_off", C.c_uint64), ("out_len", C.c_uint64), ("date", C.c_int64),
                ("n_fragments", C.c_int32), ("sha1_ok", C.c_int32)]


class JidacStats(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("fragment_ms", C.c_float), ("sha1_ms", C.c_float),
                ("dedup_ms", C.c_float), ("gather_ms", C.c_float), ("codec_ms", C.c_float),
                ("pack_ms", C.c_float), ("d2h_ms", C.c_float), ("launches", C.c_int32), ("n_files", C.c_int32),
                ("n_fragments", C.c_int32), ("n_stored", C.c_int32), ("n_dblocks", C.c_int32),
                ("reserved", C.c_int32), ("input_bytes", C.c_uint64), ("stored_bytes", C.c_uint64),
                ("archive_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class ZpaqGpuError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("libzpaqgpu: %s (%d)" % (text, code))
        self.code = code


_lib = None


def lib():
    """Load libzpaqgpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZpaqGpuError(E_NODEVICE, "libzpaqgpu.so is not built (run zpaq-v_b200/build.sh); no CPU fallback exists")
    L = C.CDLL(LIB_PATH)
    vp, u64p, i32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int)
    L.zpaqgpu_init.argtypes = [C.POINTER(vp), C.c_int]
    L.zpaqgpu_destroy.argtypes = [vp]
    L.zpaqgpu_destroy.restype = None
    L.zpaqgpu_strerror.argtypes = [C.c_int]
    L.zpaqgpu_strerror.restype = C.c_char_p
    L.zpaqgpu_last_error.argtypes = [vp]
    L.zpaqgpu_last_error.restype = C.c_char_p
    L.zpaqgpu_set_kernel.argtypes = [vp, C.c_int]
    L.zpaqgpu_set_table_mode.argtypes = [vp, C.c_int]
    L.zpaqgpu_set_workspace_limit.argtypes = [vp, C.c_uint64]
    L.zpaqgpu_set_stream.argtypes = [vp, vp]
    L.zpaqgpu_level_header.argtypes = [C.c_int, C.c_char_p, C.c_int]
    L.zpaqgpu_tables.argtypes = [vp, vp, vp]
    L.zpaqgpu_compress_blocks.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_uint64, vp, u64p]
    L.zpaqgpu_compress_blocks_header.argtypes = [vp, C.c_char_p, C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_uint64, vp, u64p]
    L.zpaqgpu_compress_blocks_dev.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, C.c_uint64, vp, u64p]
    L.zpaqgpu_find_blocks.argtypes = [vp, vp, C.c_uint64, vp, C.c_int, i32p]
    L.zpaqgpu_decompress_archive.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, u64p, vp, C.c_int, i32p]
    L.zpaqgpu_decompress_blocks_dev.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, i32p]
    L.zpaqgpu_block_begin.argtypes = [vp, C.c_int]
    L.zpaqgpu_block_begin_header.argtypes = [vp, C.c_char_p, C.c_int]
    L.zpaqgpu_segment_begin.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.zpaqgpu_segment_write.argtypes = [vp, vp, C.c_uint64]
    L.zpaqgpu_segment_end.argtypes = [vp]
    L.zpaqgpu_block_end.argtypes = [vp, vp, C.c_uint64, u64p]
    L.zpaqgpu_block_end.restype = C.c_int64
    L.zpaqgpu_last_stats.argtypes = [vp, C.POINTER(Stats)]
    L.zpaqgpu_describe_model.argtypes = [C.c_char_p, C.c_int, C.POINTER(ModelInfo)]
    L.zpaqgpu_jidac_fragment.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, i32p, i32p]
    L.zpaqgpu_jidac_add.argtypes = [vp, C.POINTER(JidacOpts), vp, vp, vp, C.c_int, vp, C.c_uint64, u64p, u64p]
    L.zpaqgpu_jidac_last_stats.argtypes = [vp, C.POINTER(JidacStats)]
    L.zpaqgpu_jidac_extract.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, u64p, vp, C.c_int, i32p, vp, C.c_uint64,
                                        u64p]
    L.zpaqgpu_stream_batch.argtypes = [vp, C.c_int, C.c_uint64]
    L.zpaqgpu_block_end_queue.argtypes = [vp]
    L.zpaqgpu_queued.argtypes = [vp, i32p, u64p]
    L.zpaqgpu_flush.argtypes = [vp, vp, C.c_uint64, u64p]
    L.zpaqgpu_flush.restype = C.c_int64
    L.zpaqgpu_multi_init.argtypes = [C.POINTER(vp), vp, C.c_int]
    L.zpaqgpu_multi_destroy.argtypes = [vp]
    L.zpaqgpu_multi_destroy.restype = None
    L.zpaqgpu_multi_device_count.argtypes = [vp]
    L.zpaqgpu_multi_ctx.argtypes = [vp, C.c_int]
    L.zpaqgpu_multi_ctx.restype = vp
    L.zpaqgpu_multi_last_error.argtypes = [vp]
    L.zpaqgpu_multi_last_error.restype = C.c_char_p
    L.zpaqgpu_multi_compress_blocks.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_uint64, vp, u64p]
    L.zpaqgpu_multi_decompress_archive.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, u64p, vp, C.c_int, i32p]
    L.zpaqgpu_multi_last_stats.argtypes = [vp, C.c_int, C.POINTER(MultiStats)]
    L.zpaqgpu_multi_jidac_add.argtypes = [vp, C.POINTER(JidacOpts), vp, vp, vp, C.c_int, vp, C.c_uint64, u64p, u64p]
    _lib = L
    return L


def level_header(level):
    buf = C.create_string_buffer(128)
    n = lib().zpaqgpu_level_header(level, buf, 128)
    if n < 0:
        raise ZpaqGpuError(n, "level_header")
    return buf.raw[:n]


def describe_model(header):
    info = ModelInfo()
    rc = lib().zpaqgpu_describe_model(bytes(header), len(header), C.byref(info))
    if rc < 0:
        raise ZpaqGpuError(rc, lib().zpaqgpu_strerror(rc).decode())
    return {k: getattr(info, k) for k, _ in info._fields_}


def tables():
    sq = (C.c_int32 * 4096)()
    st = (C.c_int32 * 32768)()
    ns = (C.c_uint8 * 1024)()
    lib().zpaqgpu_tables(sq, st, ns)
    return list(sq), list(st), bytes(ns)


class Context:
    """One zpaqgpu_ctx: bound to one GPU, single-threaded."""

    def __init__(self, device=-1):
        self._h = C.c_void_p()
        rc = lib().zpaqgpu_init(C.byref(self._h), device)
        if rc != OK:
            self._h = None
            raise ZpaqGpuError(rc, lib().zpaqgpu_strerror(rc).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().zpaqgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            detail = lib().zpaqgpu_last_error(self._h).decode(errors="replace")
            text = lib().zpaqgpu_strerror(int(rc)).decode()
            raise ZpaqGpuError(int(rc), text + (": " + detail if detail else ""))
        return rc

    def set_kernel(self, kernel):
        self._check(lib().zpaqgpu_set_kernel(self._h, kernel))

    def set_table_mode(self, mode):
        """0 auto, 1 dense, 2 paged (include/zpaqgpu.h ZPAQGPU_TABLES_*)."""
        self._check(lib().zpaqgpu_set_table_mode(self._h, mode))

    def set_workspace_limit(self, nbytes):
        self._check(lib().zpaqgpu_set_workspace_limit(self._h, nbytes))

    def set_stream(self, stream_ptr):
        self._check(lib().zpaqgpu_set_stream(self._h, stream_ptr))

    def stats(self):
        s = Stats()
        lib().zpaqgpu_last_stats(self._h, C.byref(s))
        return s.as_dict()

    # ---- batch, host buffers ----
    def compress_blocks(self, level, blocks, names=None, comments=None, header=None):
        """blocks: list of bytes-like.  Returns list of bytes, one finished ZPAQ block each."""
        n = len(blocks)
        data = b"".join(bytes(b) for b in blocks)
        off = (C.c_uint64 * (n + 1))()
        pos = 0
        for i, b in enumerate(blocks):
            off[i] = pos
            pos += len(b)
        off[n] = pos
        src = C.create_string_buffer(data, len(data)) if data else C.create_string_buffer(1)

        def strs(v):
            if v is None:
                return None
            arr = (C.c_char_p * n)()
            for i, x in enumerate(v):
                arr[i] = x.encode() if isinstance(x, str) else x
            return arr

        a_names, a_comments = strs(names), strs(comments)
        cap = len(data) + len(data) // 4 + 4096 * max(n, 1)
        out_off = (C.c_uint64 * (n + 1))()
        need = C.c_uint64(0)
        for _ in range(2):
            out = C.create_string_buffer(cap)
            if header is not None:
                rc = lib().zpaqgpu_compress_blocks_header(self._h, bytes(header), len(header), src, off, n,
                                                          a_names, a_comments, out, cap, out_off, C.byref(need))
            else:
                rc = lib().zpaqgpu_compress_blocks(self._h, level, src, off, n, a_names, a_comments, out, cap,
                                                   out_off, C.byref(need))
            if rc == E_NOSPACE:
                cap = need.value + 16
                continue
            self._check(rc)
            raw = out.raw
            return [raw[out_off[i]:out_off[i + 1]] for i in range(n)]
        raise ZpaqGpuError(E_NOSPACE, "output sizing did not converge")

    def find_blocks(self, arc):
        arc = bytes(arc)
        cap = 1024
        while True:
            starts = (C.c_uint64 * cap)()
            found = C.c_int(0)
            rc = lib().zpaqgpu_find_blocks(self._h, arc, len(arc), starts, cap, C.byref(found))
            if rc == E_NOSPACE:
                cap = found.value + 16
                continue
            self._check(rc)
            return [starts[i] for i in range(found.value)]

    def decompress_archive(self, arc):
        """Returns (plaintext bytes, [segment dict...]).  Raises on device errors; format trouble is
        reported through the returned status key of the last segment list entry."""
        arc = bytes(arc)
        cap = max(4 * len(arc), 1 << 16)
        seg_cap = 256
        need = C.c_uint64(0)
        nseg = C.c_int(0)
        for _ in range(4):
            out = C.create_string_buffer(cap)
            segs = (Segment * seg_cap)()
            rc = lib().zpaqgpu_decompress_archive(self._h, arc, len(arc), out, cap, C.byref(need), segs, seg_cap,
                                                  C.byref(nseg))
            if rc == E_NOSPACE:
                cap = max(cap, need.value + 16)
                seg_cap = max(seg_cap, nseg.value + 16)
                continue
            status = rc
            if rc not in (OK, E_FORMAT, E_UNSUPPORTED):
                self._check(rc)
            res = []
            for i in range(nseg.value):
                s = segs[i]
                name = arc[s.name_off:arc.index(b"\0", s.name_off)]
                comment = arc[s.comment_off:arc.index(b"\0", s.comment_off)]
                res.append(dict(filename=name.decode("latin1"), comment=comment.decode("latin1"),
                                out_off=s.out_off, out_len=s.out_len, block_index=s.block_index,
                                block_start=s.block_start, block_end=s.block_end, sha1_ok=s.sha1_ok))
            return out.raw[:need.value], res, status
        raise ZpaqGpuError(E_NOSPACE, "output sizing did not converge")

    # ---- jidac front end ----
    @staticmethod
    def _file_ranges(files):
        n = len(files)
        data = b"".join(bytes(b) for b in files)
        off = (C.c_uint64 * (n + 1))()
        pos = 0
        for i, b in enumerate(files):
            off[i] = pos
            pos += len(b)
        off[n] = pos
        src = C.create_string_buffer(data, len(data)) if data else C.create_string_buffer(1)
        return n, src, off, len(data)

    def jidac_fragment(self, files, fragment=6, dedup=True):
        """Fragment table of the files (list of bytes): list of dicts off/len/file/id/stored/sha1."""
        n, src, off, total = self._file_ranges(files)
        cap = 1024
        while True:
            frs = (Fragment * cap)()
            nf, ns = C.c_int(0), C.c_int(0)
            rc = lib().zpaqgpu_jidac_fragment(self._h, src, off, n, fragment, int(dedup), frs, cap, C.byref(nf),
                                              C.byref(ns))
            if rc == E_NOSPACE:
                cap = nf.value + 16
                continue
            self._check(rc)
            return [dict(off=f.off, len=f.len, file=f.file, id=f.id, stored=f.stored, sha1=bytes(f.sha1))
                    for f in frs[:nf.value]], ns.value

    def jidac_add(self, names, files, date, level=0, fragment=-1, dedup=False, block_bytes=0):
        """A journaling archive (c, d.., h.., i blocks) of the files; see include/zpaqgpu.h."""
        n, src, off, total = self._file_ranges(files)
        arr = (C.c_char_p * max(n, 1))()
        for i, x in enumerate(names):
            arr[i] = x.encode() if isinstance(x, str) else x
        opts = JidacOpts(date, level, fragment, int(dedup), 0, block_bytes)
        cap = total + total // 4 + 4096 * (n + 4)
        ln, need = C.c_uint64(0), C.c_uint64(0)
        for _ in range(2):
            out = C.create_string_buffer(cap)
            rc = lib().zpaqgpu_jidac_add(self._h, C.byref(opts), arr, src, off, n, out, cap, C.byref(ln),
                                         C.byref(need))
            if rc == E_NOSPACE:
                cap = need.value + 16
                continue
            self._check(rc)
            return out.raw[:ln.value]
        raise ZpaqGpuError(E_NOSPACE, "output sizing did not converge")

    def jidac_extract(self, arc):
        """Files of a journaling archive: list of dicts name/data/date/n_fragments/sha1_ok, index order."""
        arc = bytes(arc)
        out_cap, files_cap, names_cap = max(4 * len(arc), 1 << 16), 1024, 1 << 16
        need, nf, nn = C.c_uint64(0), C.c_int(0), C.c_uint64(0)
        for _ in range(3):
            out = C.create_string_buffer(out_cap)
            files = (JidacFile * files_cap)()
            names = C.create_string_buffer(names_cap)
            rc = lib().zpaqgpu_jidac_extract(self._h, arc, len(arc), out, out_cap, C.byref(need), files, files_cap,
                                             C.byref(nf), names, names_cap, C.byref(nn))
            if rc == E_NOSPACE:
                out_cap, files_cap, names_cap = need.value + 16, nf.value + 16, nn.value + 16
                continue
            self._check(rc)
            raw, nraw = out.raw, names.raw
            res = []
            for f in files[:nf.value]:
                name = nraw[f.name_off:nraw.index(b"\0", f.name_off)].decode("latin1")
                res.append(dict(name=name, data=raw[f.out_off:f.out_off + f.out_len], date=f.date,
                                n_fragments=f.n_fragments, sha1_ok=f.sha1_ok))
            return res
        raise ZpaqGpuError(E_NOSPACE, "output sizing did not converge")

    def jidac_stats(self):
        s = JidacStats()
        lib().zpaqgpu_jidac_last_stats(self._h, C.byref(s))
        return s.as_dict()

    # ---- streaming-shaped ----
    def block_begin(self, level=None, header=None):
        if header is not None:
            return lib().zpaqgpu_block_begin_header(self._h, bytes(header), len(header))
        return lib().zpaqgpu_block_begin(self._h, level)

    def segment_begin(self, filename, comment):
        return lib().zpaqgpu_segment_begin(self._h, filename.encode("latin1"), comment.encode("latin1"))

    def segment_write(self, data):
        data = bytes(data)
        return lib().zpaqgpu_segment_write(self._h, data, len(data))

    def segment_end(self):
        return lib().zpaqgpu_segment_end(self._h)

    def block_end_queue(self):
        """Compressor.end_block(), deferred: 0 queued, 1 queued and the queue is full, < 0 error."""
        return lib().zpaqgpu_block_end_queue(self._h)

    def stream_batch(self, max_blocks=0, max_bytes=0):
        self._check(lib().zpaqgpu_stream_batch(self._h, max_blocks, max_bytes))

    def queued(self):
        n, b = C.c_int(0), C.c_uint64(0)
        self._check(lib().zpaqgpu_queued(self._h, C.byref(n), C.byref(b)))
        return n.value, b.value

    def flush(self):
        """Codes every queued block in one batch; their bytes in queue order."""
        need = C.c_uint64(0)
        n = lib().zpaqgpu_flush(self._h, None, 0, C.byref(need))
        if n == E_NOSPACE:
            buf = C.create_string_buffer(max(need.value, 1))
            n = lib().zpaqgpu_flush(self._h, buf, need.value, C.byref(need))
            if n >= 0:
                return buf.raw[:n]
        self._check(n)
        return b""

    def block_end(self):
        need = C.c_uint64(0)
        n = lib().zpaqgpu_block_end(self._h, None, 0, C.byref(need))
        if n == E_NOSPACE:
            buf = C.create_string_buffer(max(need.value, 1))
            n = lib().zpaqgpu_block_end(self._h, buf, need.value, C.byref(need))
            if n >= 0:
                return buf.raw[:n]
        if n == E_STATE:
            return None
        self._check(n)
        return b""


class Multi:
    """zpaqgpu_multi: several GPUs of one box behind one handle (one host thread + stream per device,
    contiguous block ranges, results in block order)."""

    def __init__(self, devices=None):
        self._h = C.c_void_p()
        if devices is None:
            rc = lib().zpaqgpu_multi_init(C.byref(self._h), None, 0)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = lib().zpaqgpu_multi_init(C.byref(self._h), arr, len(devices))
        if rc != OK:
            self._h = None
            raise ZpaqGpuError(rc, lib().zpaqgpu_strerror(rc).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().zpaqgpu_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            detail = lib().zpaqgpu_multi_last_error(self._h).decode(errors="replace")
            raise ZpaqGpuError(int(rc), lib().zpaqgpu_strerror(int(rc)).decode() + (": " + detail if detail else ""))
        return rc

    def device_count(self):
        return lib().zpaqgpu_multi_device_count(self._h)

    def stats(self):
        out = []
        for k in range(self.device_count()):
            s = MultiStats()
            lib().zpaqgpu_multi_last_stats(self._h, k, C.byref(s))
            d = {f: getattr(s, f) for f, _ in s._fields_ if f != "stats"}
            d["stats"] = s.stats.as_dict()
            out.append(d)
        return out

    def compress_blocks(self, level, blocks, names=None, comments=None):
        n = len(blocks)
        data = b"".join(bytes(b) for b in blocks)
        off = (C.c_uint64 * (n + 1))()
        pos = 0
        for i, b in enumerate(blocks):
            off[i] = pos
            pos += len(b)
        off[n] = pos
        src = C.create_string_buffer(data, len(data)) if data else C.create_string_buffer(1)

        def strs(v):
            if v is None:
                return None
            arr = (C.c_char_p * max(n, 1))()
            for i, x in enumerate(v):
                arr[i] = x.encode() if isinstance(x, str) else x
            return arr

        a_names, a_comments = strs(names), strs(comments)
        cap = len(data) + len(data) // 4 + 4096 * max(n, 1)
        out_off = (C.c_uint64 * (n + 1))()
        need = C.c_uint64(0)
        for _ in range(2):
            out = C.create_string_buffer(cap)
            rc = lib().zpaqgpu_multi_compress_blocks(self._h, level, src, off, n, a_names, a_comments, out, cap,
                                                     out_off, C.byref(need))
            if rc == E_NOSPACE:
                cap = need.value + 16
                continue
            self._check(rc)
            raw = out.raw
            return [raw[out_off[i]:out_off[i + 1]] for i in range(n)]
        raise ZpaqGpuError(E_NOSPACE, "output sizing did not converge")

    def jidac_add(self, names, files, date, level=0, fragment=-1, dedup=False, block_bytes=0):
        """`jidac add` over all devices of the handle (zpaqgpu_multi_jidac_add)."""
        n, src, off, total = Context._file_ranges(files)
        arr = (C.c_char_p * max(n, 1))()
        for i, x in enumerate(names):
            arr[i] = x.encode() if isinstance(x, str) else x
        opts = JidacOpts(date, level, fragment, int(dedup), 0, block_bytes)
        cap = total + total // 4 + 4096 * (n + 4)
        ln, need = C.c_uint64(0), C.c_uint64(0)
        for _ in range(2):
            out = C.create_string_buffer(cap)
            rc = lib().zpaqgpu_multi_jidac_add(self._h, C.byref(opts), arr, src, off, n, out, cap, C.byref(ln),
                                               C.byref(need))
            if rc == E_NOSPACE:
                cap = need.value + 16
                continue
            self._check(rc)
            return out.raw[:ln.value]
        raise ZpaqGpuError(E_NOSPACE, "output sizing did not converge")

    def decompress_archive(self, arc):
        """(plaintext, [segment dict...], status), as Context.decompress_archive."""
        arc = bytes(arc)
        cap = max(4 * len(arc), 1 << 16)
        seg_cap = 256
        need, nseg = C.c_uint64(0), C.c_int(0)
        for _ in range(4):
            out = C.create_string_buffer(cap)
            segs = (Segment * seg_cap)()
            rc = lib().zpaqgpu_multi_decompress_archive(self._h, arc, len(arc), out, cap, C.byref(need), segs, seg_cap,
                                                        C.byref(nseg))
            if rc == E_NOSPACE:
                cap = max(cap, need.value + 16)
                seg_cap = max(seg_cap, nseg.value + 16)
                continue
            status = rc
            if rc not in (OK, E_FORMAT, E_UNSUPPORTED):
                self._check(rc)
            res = []
            for i in range(nseg.value):
                s = segs[i]
                name = arc[s.name_off:arc.index(b"\0", s.name_off)]
                comment = arc[s.comment_off:arc.index(b"\0", s.comment_off)]
                res.append(dict(filename=name.decode("latin1"), comment=comment.decode("latin1"),
                                out_off=s.out_off, out_len=s.out_len, block_index=s.block_index,
                                block_start=s.block_start, block_end=s.block_end, sha1_ok=s.sha1_ok))
            return out.raw[:need.value], res, status
        raise ZpaqGpuError(E_NOSPACE, "output sizing did not converge")
