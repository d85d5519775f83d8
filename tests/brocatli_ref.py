"""Test infrastructure: a whole-buffer restatement of the reference's stream stitcher BroCatli (src/concat/mod.rs:38-120 header
parsing, :296-330 stripping the final empty metablock, :331-604 splicing a new stream in).  Not used by the product."""


class NotCraftedForConcatenation(Exception):
    pass


def _bits(buf, pos, n):
    v = 0
    for i in range(n):
        v |= ((buf[(pos + i) >> 3] >> ((pos + i) & 7)) & 1) << i
    return v


def window_bits(buf):
    """parse_window_size (mod.rs:38-72): (lgwin, header length in bits)."""
    b0 = buf[0]
    if b0 & 1 == 0:
        return 16, 1
    if (b0 & 15) in (3, 5, 7, 9, 11, 13, 15):
        return 18 + ((b0 & 15) - 3) // 2, 4
    m = {0x71: 15, 0x61: 14, 0x51: 13, 0x41: 12, 0x31: 11, 0x21: 10, 0x01: 17}
    if (b0 & 127) in m:
        return m[b0 & 127], 7
    raise NotCraftedForConcatenation("large-window header")


def first_metablock_aligned_offset(buf, start=None):
    """detect_varlen_offset (mod.rs:74-120): bit offset behind the header of the first metablock, which must be an empty last
    metablock, a metadata metablock or an uncompressed one -- after it the stream is byte aligned."""
    pos = window_bits(buf)[1] if start is None else start
    if _bits(buf, pos, 1):  # ISLAST
        if _bits(buf, pos + 1, 1):
            return pos + 2
        pos += 1  # ISLAST with data: falls through to MNIBBLES like the reference does
        pos += 1
    else:
        pos += 1
    mn = _bits(buf, pos, 2)
    pos += 2
    if mn == 3:
        if _bits(buf, pos, 1):
            raise NotCraftedForConcatenation("reserved bit")
        skipbytes = _bits(buf, pos + 1, 2)
        return pos + 3 + 8 * skipbytes
    pos += 4 * (mn + 4)
    if not _bits(buf, pos, 1):
        raise NotCraftedForConcatenation("first metablock is compressed")
    return pos + 1


def strip_final_empty_metablock(bits):
    """flush_previous_stream (mod.rs:296-330): the two highest set bits of the stream must be ISLAST, ISLASTEMPTY."""
    n = len(bits)
    while n and not bits[n - 1]:
        n -= 1
    if n < 2 or not bits[n - 2]:
        raise NotCraftedForConcatenation("stream does not end with an empty last metablock")
    return bits[:n - 2]


def _to_bits(buf):
    return [(byte >> i) & 1 for byte in buf for i in range(8)]


def concat(streams):
    """Stitches complete brotli streams into one.  Every stream but the first must start (behind its window bits) with a metablock
    whose payload is byte aligned; its header bits are re-emitted at the current bit offset, then both sides are byte aligned."""
    out = []
    lgwin0 = None
    for k, s in enumerate(streams):
        lgwin, hdr = window_bits(s)
        if k == 0:
            lgwin0 = lgwin
            out = strip_final_empty_metablock(_to_bits(s))
            continue
        if lgwin > lgwin0:
            raise NotCraftedForConcatenation("window larger than the first file's")
        off = first_metablock_aligned_offset(s)
        bits = _to_bits(s)
        head = bits[hdr:off]
        if len(head) == 2 and head == [1, 1]:  # an empty stream contributes nothing
            continue
        out += head
        out += [0] * (-len(out) % 8)
        body = strip_final_empty_metablock(bits[(off + 7) // 8 * 8:])
        out += body
    out += [1, 1]
    out += [0] * (-len(out) % 8)
    return bytes(sum(out[i + j] << j for j in range(8)) for i in range(0, len(out), 8))
