"""Host-side mirror of the reference's JidacArchive (/root/reference/zpaq/jidac.v:120-296) over
libzpaqgpu, plus the reader the reference lacks (it only writes journaling archives).

All hashing, fragmentation, dedup and block coding happens on the GPU through the C ABI
(zpaqgpu_jidac_add / zpaqgpu_decompress_archive); this module only marshals and, for extraction,
follows the h and i tables to put fragments back into files.
"""
import struct

from . import binding
from .codec import default_context


def make_jidac_filename(date, block_type, num):
    """jidac.v:47-49"""
    return "jDC%s%s%s" % (str(date).rjust(14, "0"), block_type, str(num).rjust(10, "0"))


class JidacArchive:
    """jidac.v:120-296.  create_archive(files, method) writes c, d.., h.., i blocks to the Writer;
    `date` defaults to the current time like get_jidac_date() (jidac.v:31-35) and can be fixed for
    reproducible bytes."""

    def __init__(self, ctx=None, date=None):
        self._ctx = ctx or default_context()
        if date is None:
            import time
            t = time.localtime()
            date = (t.tm_year * 10000000000 + t.tm_mon * 100000000 + t.tm_mday * 1000000 + t.tm_hour * 10000 +
                    t.tm_min * 100 + t.tm_sec)
        self.date = date
        self.output = None

    @classmethod
    def new(cls, ctx=None, date=None):
        return cls(ctx, date)

    def set_output(self, writer):
        self.output = writer

    def create_archive(self, files, method=0):
        """files: dict name -> bytes (insertion order is archive order, as with a V map).  The
        reference ignores `method` and always stores (jidac.v:94-118); so does this call."""
        if self.output is None:
            return                      # jidac.v:182-184
        del method
        names = list(files.keys())
        arc = self._ctx.jidac_add(names, [files[k] for k in names], self.date, level=0, fragment=-1, dedup=False,
                                  block_bytes=0)
        self.output.write(arc)

    def add(self, files, level=1, fragment=6, dedup=True, block_bytes=1 << 24):
        """What the task calls `jidac add`: rolling-hash fragments, SHA-1 dedup, packed d blocks
        coded at `level`."""
        if self.output is None:
            return
        names = list(files.keys())
        arc = self._ctx.jidac_add(names, [files[k] for k in names], self.date, level=level, fragment=fragment,
                                  dedup=dedup, block_bytes=block_bytes)
        self.output.write(arc)


def parse_index(segments, plain):
    """Fragment and file tables out of the decoded h and i blocks.  segments/plain as returned by
    Context.decompress_archive.  Returns (frag: id -> (d_block_first_id, offset_in_block, size),
    files: name -> [ids] in archive order, dblocks: first_id -> plaintext range)."""
    frag, files, dblocks = {}, {}, {}
    for s in segments:
        nm = s["filename"]
        if len(nm) != 28 or not nm.startswith("jDC"):
            continue
        kind, num = nm[17], int(nm[18:])
        body = plain[s["out_off"]:s["out_off"] + s["out_len"]]
        if kind == "d":
            dblocks[num] = (s["out_off"], s["out_len"])
        elif kind == "h":
            pos, fid, at = 4, num, 0
            while pos + 24 <= len(body):
                size = struct.unpack_from("<I", body, pos + 20)[0]
                frag[fid] = (num, at, size, bytes(body[pos:pos + 20]))
                at += size
                fid += 1
                pos += 24
        elif kind == "i":
            pos = 0
            while pos + 8 <= len(body):
                date = struct.unpack_from("<q", body, pos)[0]
                end = body.index(b"\0", pos + 8)
                name = body[pos + 8:end].decode("latin1")
                pos = end + 1
                ids = []
                if date != 0:
                    na = struct.unpack_from("<I", body, pos)[0]
                    pos += 4 + na
                    ni = struct.unpack_from("<I", body, pos)[0]
                    pos += 4
                    ids = list(struct.unpack_from("<%dI" % ni, body, pos))
                    pos += 4 * ni
                    files[name] = ids
                else:
                    files.pop(name, None)
    return frag, files, dblocks


def extract(archive, ctx=None):
    """Decode a journaling archive on the GPU and reassemble its files there (zpaqgpu_jidac_extract):
    dict name -> bytes.  Raises when a fragment does not hash to the SHA-1 its table records."""
    ctx = ctx or default_context()
    out = {}
    for f in ctx.jidac_extract(archive):
        if not f["sha1_ok"]:
            raise binding.ZpaqGpuError(binding.E_FORMAT, "SHA-1 mismatch in file " + f["name"])
        out[f["name"]] = f["data"]
    return out


def extract_on_host(archive, ctx=None):
    """The same result with the tables followed on the host (parse_index): kept as a cross-check of
    the C ABI call in the tests."""
    ctx = ctx or default_context()
    plain, segs, status = ctx.decompress_archive(archive)
    if status != binding.OK:
        raise binding.ZpaqGpuError(status, "archive did not decode")
    if any(s["sha1_ok"] == 0 for s in segs):
        raise binding.ZpaqGpuError(binding.E_FORMAT, "SHA-1 mismatch in a block")
    frag, files, dblocks = parse_index(segs, plain)
    out = {}
    for name, ids in files.items():
        parts = []
        for fid in ids:
            first, at, size, _ = frag[fid]
            lo = dblocks[first][0] + at
            parts.append(plain[lo:lo + size])
        out[name] = b"".join(parts)
    return out
