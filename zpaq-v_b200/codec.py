"""Host-side mirror of the reference's Compressor / Decompresser method surface over libzpaqgpu.

Same method names, argument meaning and (silent) error behaviour as the V types
(/root/reference/zpaq/compressor.v:33-413, decompressor.v:187-640, io.v:6-21), so that a user of
zpaq.Compressor / zpaq.Decompresser can switch by changing the import.  All coding happens on the
GPU through the C ABI; nothing here compresses or decompresses on the host.
"""
from . import binding


class FileReader:
    """io.v:41-78 -- reads from a byte buffer; get() returns -1 at EOF."""

    def __init__(self, data=b""):
        self.data = bytes(data)
        self.pos = 0

    def get(self):
        if self.pos >= len(self.data):
            return -1
        c = self.data[self.pos]
        self.pos += 1
        return c

    def read(self, n):
        chunk = self.data[self.pos:self.pos + n]
        self.pos += len(chunk)
        return chunk

    def position(self):
        return self.pos


class FileWriter:
    """io.v:80-110 -- collects bytes."""

    def __init__(self):
        self.buf = bytearray()

    def put(self, c):
        self.buf.append(c & 255)

    def write(self, data):
        self.buf += bytes(data)

    def bytes(self):
        return bytes(self.buf)


_shared_ctx = None


def default_context():
    global _shared_ctx
    if _shared_ctx is None:
        _shared_ctx = binding.Context()
    return _shared_ctx


class Compressor:
    """compressor.v:16-413.  Bytes are gathered per segment and the block is coded on the device
    at end_block(); the bytes written to the Writer are identical to the reference's.

    batch=True is for callers that write many blocks one after the other (cmd/main.v:288-317, one block
    per file): end_block() queues the block, the queue is coded in ONE launch when it is full or at
    flush() -- which the caller adds before it closes the output (cmd/main.v:320).  Same bytes, same order."""

    def __init__(self, ctx=None, batch=False):
        self._ctx = ctx or default_context()
        self._input = None
        self._output = None
        self._state = "start"
        self._batch = batch

    def set_input(self, reader):
        self._input = reader

    def set_output(self, writer):
        self._output = writer

    def start_block(self, level):
        if self._state != "start":
            return                      # compressor.v:80-82: silently ignored
        self._ctx._check(self._ctx.block_begin(level=level))
        self._state = "block"

    def start_block_header(self, header):
        """Extension mirroring the oracle's: a model header in the levels.v layout."""
        if self._state != "start":
            return
        self._ctx._check(self._ctx.block_begin(header=header))
        self._state = "block"

    def start_segment(self, filename="", comment=""):
        if self._state != "block":
            return                      # compressor.v:213-215
        self._ctx._check(self._ctx.segment_begin(filename, comment))
        self._state = "segment"

    def compress(self, n):
        """Pull up to n bytes from the Reader; True when n were consumed (compressor.v:259-293)."""
        if self._state != "segment" or self._input is None:
            return False
        if hasattr(self._input, "read"):
            chunk = self._input.read(n)
        else:
            got = bytearray()
            while len(got) < n:
                c = self._input.get()
                if c < 0:
                    break
                got.append(c)
            chunk = bytes(got)
        self._ctx._check(self._ctx.segment_write(chunk))
        return len(chunk) == n

    def end_segment(self):
        if self._state != "segment":
            return
        self._ctx._check(self._ctx.segment_end())
        self._state = "block"

    def end_block(self):
        if self._state != "block":
            return
        if self._batch:
            full = self._ctx._check(self._ctx.block_end_queue())
            self._state = "start"
            if full:
                self.flush()
            return
        data = self._ctx.block_end()
        if data and self._output is not None:
            self._output.write(data)
        self._state = "start"

    def flush(self):
        """Deliver the queued blocks (batch mode); a no-op otherwise."""
        if self._state != "start":
            return
        data = self._ctx.flush()
        if data and self._output is not None:
            self._output.write(data)


class Decompresser:
    """decompressor.v:170-640.  The archive is decoded on the device in one batch when the first
    block is asked for; the calls below then walk the results in the reference's order."""

    def __init__(self, ctx=None):
        self._ctx = ctx or default_context()
        self._input = None
        self._output = None
        self._segs = None
        self._plain = b""
        self._block = -1        # index of the current block
        self._cursor = 0        # next segment record
        self._cur = None        # current segment
        self._served = 0
        self._state = "start"
        self.status = binding.OK

    def set_input(self, reader):
        self._input = reader
        self._segs = None

    def set_output(self, writer):
        self._output = writer

    def _load(self):
        if self._segs is not None:
            return
        if hasattr(self._input, "data"):
            arc = self._input.data[self._input.pos:]
        else:
            arc = bytearray()
            while True:
                c = self._input.get()
                if c < 0:
                    break
                arc.append(c)
            arc = bytes(arc)
        self._plain, self._segs, self.status = self._ctx.decompress_archive(arc)
        self._blocks = sorted({s["block_index"] for s in self._segs})
        self._block_pos = 0

    def find_block(self):
        if self._input is None:
            return False
        self._load()
        if self._block_pos >= len(self._blocks):
            return False
        self._block = self._blocks[self._block_pos]
        self._block_pos += 1
        while self._cursor < len(self._segs) and self._segs[self._cursor]["block_index"] < self._block:
            self._cursor += 1
        self._state = "block"
        return True

    def find_filename(self):
        if self._state != "block":
            return False
        if self._cursor < len(self._segs) and self._segs[self._cursor]["block_index"] == self._block:
            self._cur = self._segs[self._cursor]
            self._cursor += 1
            self._served = 0
            self._state = "segment"
            return True
        self._state = "start"
        return False

    def get_filename(self):
        return self._cur["filename"] if self._cur else ""

    def get_comment(self):
        return self._cur["comment"] if self._cur else ""

    def decompress(self, n=-1):
        """Write up to n bytes (all when n < 0) to the Writer; True while more remain."""
        if self._state != "segment":
            return False
        left = self._cur["out_len"] - self._served
        take = left if n < 0 else min(n, left)
        lo = self._cur["out_off"] + self._served
        if take and self._output is not None:
            self._output.write(self._plain[lo:lo + take])
        self._served += take
        # true when n bytes were produced; false once the EOF marker is reached (decompressor.v:484-514)
        return n >= 0 and take == n

    def read_segment_end(self):
        if self._state != "segment":
            return
        self._state = "block"

    def last_sha1_ok(self):
        return self._cur["sha1_ok"] if self._cur else -1
