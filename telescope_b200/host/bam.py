# -*- coding: utf-8 -*-
"""Minimal streaming reader for name-collated BAM / SAM files (host side; stdlib only).

The reference reads alignments through pysam (telescope/utils/alignment.py:115-161, calignment.pyx); pysam is not
part of this image, and the only fields Telescope needs from a record are the ones below.  BGZF is a series of gzip
members, which the stdlib `gzip` module streams transparently.
"""
import gzip
import struct

FPAIRED, FPROPER, FUNMAP, FREVERSE, FREAD1, FSECONDARY = 0x1, 0x2, 0x4, 0x10, 0x40, 0x100
_FIXED = struct.Struct("<iiBBHHHIiii")
_TAG_FMT = {b"c": "<b", b"C": "<B", b"s": "<h", b"S": "<H", b"i": "<i", b"I": "<I", b"f": "<f"}
_TAG_SIZE = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}
_CIGAR_REF = (True, False, True, True, False, False, False, True, True)   # M I D N S H P = X consume reference
_CIGAR_BLOCK = (True, False, False, False, False, False, False, True, True)  # M = X produce an aligned block


class Segment(object):
    """One alignment record: just what fragment pairing, overlap and scoring need."""
    __slots__ = ("name", "flag", "ref_id", "pos", "next_ref_id", "next_pos", "tlen", "blocks", "score", "tags",
                 "raw", "aux_off", "_ed")   # raw BAM record, offset of its aux block (kept on request), its editor

    @property
    def is_paired(self):
        return bool(self.flag & FPAIRED)

    @property
    def is_proper_pair(self):
        return bool(self.flag & FPROPER)

    @property
    def is_unmapped(self):
        return bool(self.flag & FUNMAP)

    @property
    def is_reverse(self):
        return bool(self.flag & FREVERSE)

    @property
    def is_read1(self):
        return bool(self.flag & FREAD1)


def _blocks_from_cigar(pos, ops):
    """Aligned reference blocks [start, end) as pysam's get_blocks() reports them."""
    out = []
    for op, ln in ops:
        if _CIGAR_BLOCK[op]:
            out.append((pos, pos + ln))
            pos += ln
        elif _CIGAR_REF[op]:
            pos += ln
    return out


def _scan_tags(buf, want=b"AS"):
    """Walk a BAM aux block; returns {tag: value} for integer/char/string tags (enough for AS, ZF, ZT, CB)."""
    tags, p, n = {}, 0, len(buf)
    while p + 3 <= n:
        tag, typ = buf[p:p + 2], buf[p + 2:p + 3]
        p += 3
        if typ in _TAG_FMT:
            size = _TAG_SIZE[typ]
            tags[tag] = struct.unpack_from(_TAG_FMT[typ], buf, p)[0]
            p += size
        elif typ == b"A":
            tags[tag] = buf[p:p + 1].decode()
            p += 1
        elif typ in (b"Z", b"H"):
            e = buf.index(b"\0", p)
            tags[tag] = buf[p:e].decode()
            p = e + 1
        elif typ == b"B":
            sub = buf[p:p + 1]
            cnt = struct.unpack_from("<I", buf, p + 1)[0]
            p += 5 + cnt * _TAG_SIZE[sub]
        else:
            break
    return tags


class AlignmentReader(object):
    """Iterate Segments of a BAM (binary, BGZF) or SAM (text) file in file order."""

    def __init__(self, path, keep_raw=False):
        self.path = path
        self.keep_raw = keep_raw
        fh = open(path, "rb")
        magic = fh.read(2)
        fh.close()
        self.is_bam = magic == b"\x1f\x8b"
        self.references, self.lengths, self.header_text = [], [], ""
        self._fh = gzip.open(path, "rb") if self.is_bam else open(path, "rt")
        if self.is_bam:
            self._read_bam_header()
        else:
            self._read_sam_header()

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- BAM
    def _read_bam_header(self):
        fh = self._fh
        if fh.read(4) != b"BAM\1":
            raise ValueError("%s: not a BAM file" % self.path)
        l_text, = struct.unpack("<i", fh.read(4))
        self.header_text = fh.read(l_text).rstrip(b"\0").decode()
        n_ref, = struct.unpack("<i", fh.read(4))
        for _ in range(n_ref):
            l_name, = struct.unpack("<i", fh.read(4))
            self.references.append(fh.read(l_name)[:-1].decode())
            self.lengths.append(struct.unpack("<i", fh.read(4))[0])

    def _iter_bam(self):
        fh = self._fh
        while True:
            head = fh.read(4)
            if len(head) < 4:
                return
            size, = struct.unpack("<i", head)
            rec = fh.read(size)
            ref_id, pos, l_name, _mapq, _bin, n_cig, flag, l_seq, nref, npos, tlen = _FIXED.unpack_from(rec, 0)
            s = Segment()
            s._ed = None
            s.flag, s.ref_id, s.pos, s.next_ref_id, s.next_pos, s.tlen = flag, ref_id, pos, nref, npos, tlen
            p = 32
            s.name = rec[p:p + l_name - 1].decode()
            p += l_name
            cig = struct.unpack_from("<%dI" % n_cig, rec, p) if n_cig else ()
            p += 4 * n_cig
            s.blocks = _blocks_from_cigar(pos, [(c & 0xF, c >> 4) for c in cig])
            p += (l_seq + 1) // 2 + l_seq
            s.tags = _scan_tags(rec[p:])
            s.score = s.tags.get(b"AS")
            if self.keep_raw:
                s.raw, s.aux_off = rec, p
            yield s

    # ---- SAM
    def _read_sam_header(self):
        self._pending = None
        lines = []
        for line in self._fh:
            if not line.startswith("@"):
                self._pending = line
                break
            lines.append(line)
            if line.startswith("@SQ"):
                f = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                self.references.append(f["SN"])
                self.lengths.append(int(f["LN"]))
        self.header_text = "".join(lines)

    def _iter_sam(self):
        import itertools
        import re
        ref_index = {n: i for i, n in enumerate(self.references)}
        ops = "MIDNSHP=X"
        first = [self._pending] if self._pending else []
        for line in itertools.chain(first, self._fh):
            f = line.rstrip("\n").split("\t")
            s = Segment()
            s._ed = None
            s.name, s.flag = f[0], int(f[1])
            s.ref_id = ref_index.get(f[2], -1)
            s.pos = int(f[3]) - 1
            s.next_ref_id = s.ref_id if f[6] == "=" else ref_index.get(f[6], -1)
            s.next_pos = int(f[7]) - 1
            s.tlen = int(f[8])
            cig = [] if f[5] == "*" else [(ops.index(o), int(n)) for n, o in re.findall(r"(\d+)([MIDNSHP=X])", f[5])]
            s.blocks = _blocks_from_cigar(s.pos, cig)
            s.tags = {}
            for t in f[11:]:
                k, typ, v = t.split(":", 2)
                s.tags[k.encode()] = int(v) if typ == "i" else v
            s.score = s.tags.get(b"AS")
            yield s

    def __iter__(self):
        return self._iter_bam() if self.is_bam else self._iter_sam()


def bundles(segments):
    """Consecutive records that share a query name (the file must be name-collated, as the reference requires)."""
    group = []
    for s in segments:
        if group and s.name != group[0].name:
            yield group
            group = []
        group.append(s)
    if group:
        yield group


# --------------------------------------------------------------------------------------------------- writing
def _aux_without(aux, tag):
    """The aux block with every occurrence of `tag` removed."""
    out, p, n = [], 0, len(aux)
    while p + 3 <= n:
        t, typ = aux[p:p + 2], aux[p + 2:p + 3]
        q = p + 3
        if typ in _TAG_SIZE:
            q += _TAG_SIZE[typ]
        elif typ in (b"Z", b"H"):
            q = aux.index(b"\0", q) + 1
        elif typ == b"B":
            sub = aux[q:q + 1]
            q += 5 + struct.unpack_from("<I", aux, q + 1)[0] * _TAG_SIZE[sub]
        else:
            break
        if t != tag:
            out.append(aux[p:q])
        p = q
    return b"".join(out)


def _encode_tag(tag, value):
    if isinstance(value, str):
        return tag + b"Z" + value.encode() + b"\0"
    v = int(value)                       # smallest integer type that holds the value, as pysam chooses it
    if v < 0:
        for typ, lo in ((b"c", -128), (b"s", -32768), (b"i", -2 ** 31)):
            if v >= lo:
                return tag + typ + struct.pack(_TAG_FMT[typ], v)
    for typ, hi in ((b"C", 255), (b"S", 65535), (b"I", 2 ** 32 - 1)):
        if v <= hi:
            return tag + typ + struct.pack(_TAG_FMT[typ], v)
    raise ValueError("integer tag out of range")


class RecordEditor(object):
    """Mutable view of one raw BAM record: flag, mapping quality and aux tags (what Telescope edits)."""

    def __init__(self, seg):
        self.seg = seg
        self.head = bytearray(seg.raw[:seg.aux_off])
        self.aux = bytes(seg.raw[seg.aux_off:])

    flag = property(lambda self: struct.unpack_from("<H", self.head, 14)[0])

    def set_flag(self, bits):
        struct.pack_into("<H", self.head, 14, self.flag | bits)

    def unset_flag(self, bits):
        struct.pack_into("<H", self.head, 14, self.flag & ~bits & 0xFFFF)

    def set_mapq(self, q):
        self.head[9] = max(0, min(255, int(q)))

    def set_tag(self, tag, value):
        tag = tag.encode() if isinstance(tag, str) else tag
        self.aux = _aux_without(self.aux, tag) + _encode_tag(tag, value)
        self.seg.tags[tag] = value

    def tobytes(self):
        body = bytes(self.head) + self.aux
        return struct.pack("<i", len(body)) + body


_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


class BamWriter(object):
    """Write a BAM file: BGZF blocks of at most 64 KiB of payload, raw-deflate compressed (stdlib zlib)."""

    def __init__(self, path, header_text, references, lengths, level=6):
        import zlib
        self._zlib, self._level = zlib, level
        self._fh = open(path, "wb")
        self._buf = bytearray()
        text = header_text.encode()
        self.write_bytes(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(references)))
        for name, ln in zip(references, lengths):
            nm = name.encode() + b"\0"
            self.write_bytes(struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln))

    def write_bytes(self, data):
        self._buf += data
        while len(self._buf) >= 0xFF00:
            self._flush_block(self._buf[:0xFF00])
            del self._buf[:0xFF00]

    def write(self, record):
        """record: RecordEditor, Segment (with raw kept) or raw bytes of one record body."""
        if isinstance(record, RecordEditor):
            self.write_bytes(record.tobytes())
        elif isinstance(record, Segment):
            self.write_bytes(struct.pack("<i", len(record.raw)) + record.raw)
        else:
            self.write_bytes(struct.pack("<i", len(record)) + record)

    def _flush_block(self, payload):
        z = self._zlib
        comp = z.compressobj(self._level, z.DEFLATED, -15)
        data = comp.compress(bytes(payload)) + comp.flush()
        bsize = len(data) + 25                                  # total block size - 1
        self._fh.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + data +
                       struct.pack("<II", z.crc32(bytes(payload)) & 0xFFFFFFFF, len(payload)))

    def close(self):
        if self._fh is None:
            return
        if self._buf:
            self._flush_block(self._buf)
            self._buf = bytearray()
        self._fh.write(_BGZF_EOF)
        self._fh.close()
        self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
