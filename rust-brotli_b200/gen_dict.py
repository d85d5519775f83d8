"""Build step: writes csrc/bro_dict_data.inc -- the RFC 7932 static dictionary (appendix A) and this library's own
hash table over the 4-byte prefixes of its words.

The 122 784 dictionary bytes are a constant of the brotli format; they are read at build time from the system's
libbrotlicommon (BrotliGetDictionary), so nothing is copied from the reference tree.  The hash table is generated here
(it does not have to equal the reference's kStaticDictionaryHash: any table only changes which words are found):
bucket = Hash14(first 4 bytes), two slots per bucket, entry = word_index << 5 | length, 0 = empty.
Slot policy: per bucket keep the two words with the lowest index inside their length class (most frequent first),
longer word first on ties."""
import ctypes, os, sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "csrc", "bro_dict_data.inc")


class BrotliDictionary(ctypes.Structure):
    _fields_ = [("size_bits_by_length", ctypes.c_uint8 * 32), ("offsets_by_length", ctypes.c_uint32 * 32),
                ("data_size", ctypes.c_size_t), ("data", ctypes.POINTER(ctypes.c_uint8))]


def load_dictionary():
    lib = ctypes.CDLL("libbrotlicommon.so.1")
    lib.BrotliGetDictionary.restype = ctypes.POINTER(BrotliDictionary)
    d = lib.BrotliGetDictionary().contents
    data = bytes(d.data[:d.data_size])
    assert len(data) == 122784
    return data, list(d.size_bits_by_length), list(d.offsets_by_length)


def hash14(b4):
    return ((int.from_bytes(b4, "little") * 0x1E35A7BD) & 0xFFFFFFFF) >> 18


def build_hash(data, bits, offs, policy="index"):
    buckets = {}
    for ln in range(4, 25):
        for idx in range(1 << bits[ln]):
            w = data[offs[ln] + ln * idx: offs[ln] + ln * (idx + 1)]
            buckets.setdefault(hash14(w[:4]), []).append((ln, idx))
    table = [0] * 32768
    for h, items in buckets.items():
        if policy == "index":      # most frequent (lowest index) first, longer first on ties
            items.sort(key=lambda t: (t[1], -t[0]))
        elif policy == "long":     # longest first
            items.sort(key=lambda t: (-t[0], t[1]))
        elif policy == "short":
            items.sort(key=lambda t: (t[0], t[1]))
        elif policy == "relindex":  # rank relative to the size of the length class
            items.sort(key=lambda t: (t[1] / float(1 << bits[t[0]]), -t[0]))
        for s, (ln, idx) in enumerate(items[:2]):
            table[2 * h + s] = (idx << 5) | ln
    return table


# ---- quality 10 / 11: every word (plain, first letter uppercased, all uppercased) by the hash of its first four bytes, and the
# RFC 7932 appendix B transforms grouped by prefix, so that bro_dict.cuh can enumerate all (word, transform) pairs that match
# at a position (the job of BrotliFindAllStaticDictionaryMatches, static_dict.rs:309; the reference hard-codes a decision tree
# over the transforms, here the transform list itself is the table).
class BrotliTransforms(ctypes.Structure):
    _fields_ = [("prefix_suffix_size", ctypes.c_uint16), ("prefix_suffix", ctypes.POINTER(ctypes.c_uint8)),
                ("prefix_suffix_map", ctypes.POINTER(ctypes.c_uint16)), ("num_transforms", ctypes.c_uint32),
                ("transforms", ctypes.POINTER(ctypes.c_uint8)), ("params", ctypes.POINTER(ctypes.c_uint8)),
                ("cutoff", ctypes.c_int16 * 10)]


def load_transforms():
    lib = ctypes.CDLL("libbrotlicommon.so.1")
    lib.BrotliGetTransforms.restype = ctypes.POINTER(BrotliTransforms)
    t = lib.BrotliGetTransforms().contents
    ps = bytes(t.prefix_suffix[:t.prefix_suffix_size])

    def s(i):
        o = t.prefix_suffix_map[i]
        return ps[o + 1:o + 1 + ps[o]]
    tr = [(s(t.transforms[3 * i]), t.transforms[3 * i + 1], s(t.transforms[3 * i + 2])) for i in range(t.num_transforms)]
    assert len(tr) == 121 and tr[0] == (b"", 0, b"") and tr[9] == (b"", 10, b"") and tr[44] == (b"", 11, b"")
    return tr


def to_upper(word, all_):
    """RFC 7932 section 8: uppercase-first / uppercase-all on (pseudo) UTF-8"""
    w = bytearray(word)
    i = 0
    while i < len(w):
        c = w[i]
        if c < 0xC0:
            if 97 <= c <= 122:
                w[i] ^= 32
            step = 1
        elif c < 0xE0:
            if i + 1 < len(w):
                w[i + 1] ^= 32
            step = 2
        else:
            if i + 2 < len(w):
                w[i + 2] ^= 5
            step = 3
        i += step
        if not all_:
            break
    return bytes(w)


def build_lut(data, bits, offs):
    buckets = {}
    for ln in range(4, 25):
        for idx in range(1 << bits[ln]):
            w = data[offs[ln] + ln * idx: offs[ln] + ln * (idx + 1)]
            seen = {w: 0}
            first, all_ = to_upper(w, False), to_upper(w, True)
            if first not in seen:
                seen[first] = 1
            if all_ not in seen:
                seen[all_] = 2
            for v, kind in seen.items():
                buckets.setdefault(hash15(v[:4]), []).append((ln, idx, kind))
    table = [0] * 32768
    entries = []
    for h in sorted(buckets):
        table[h] = len(entries) + 1
        items = sorted(buckets[h])
        for k, (ln, idx, kind) in enumerate(items):
            entries.append(idx | (ln << 16) | (kind << 21) | ((1 << 31) if k + 1 == len(items) else 0))
    assert len(entries) < 65535
    return table, entries


def hash15(b4):
    return ((int.from_bytes(b4, "little") * 0x1E35A7BD) & 0xFFFFFFFF) >> 17


def build_transform_groups(tr):
    groups = {}
    for tid, (pre, typ, suf) in enumerate(tr):
        if typ >= 12:   # omit-first-N: not searched (the reference does not either)
            continue
        assert len(pre) <= 8 and len(suf) <= 8
        groups.setdefault(pre, []).append((tid, typ, suf))
    order = sorted(groups, key=lambda p: (len(p), p))
    return [(p, groups[p]) for p in order]


def main(policy="index"):
    data, bits, offs = load_dictionary()
    table = build_hash(data, bits, offs, policy)
    with open(OUT, "w") as f:
        f.write("// generated by rust-brotli_b200/gen_dict.py (policy %s) -- do not edit, do not commit\n" % policy)
        f.write("static const unsigned char kDictData[122784] = {\n")
        for i in range(0, len(data), 32):
            f.write(",".join(str(b) for b in data[i:i + 32]) + ",\n")
        f.write("};\nstatic const unsigned short kDictHash[32768] = {\n")
        for i in range(0, 32768, 16):
            f.write(",".join(str(v) for v in table[i:i + 16]) + ",\n")
        f.write("};\n")
        lut, entries = build_lut(data, bits, offs)
        f.write("static const unsigned short kDictLutBuckets[32768] = {\n")
        for i in range(0, 32768, 16):
            f.write(",".join(str(v) for v in lut[i:i + 16]) + ",\n")
        f.write("};\nstatic const unsigned int kDictLutEntries[%d] = {\n" % len(entries))
        for i in range(0, len(entries), 8):
            f.write(",".join("%uu" % v for v in entries[i:i + 8]) + ",\n")
        f.write("};\n")
        groups = build_transform_groups(load_transforms())
        # group table: prefix_len, prefix[8], first, count ; transform table: id, type, suffix_len, suffix[8]
        f.write("static const unsigned char kDictTrGroups[%d][11] = {\n" % len(groups))
        first = 0
        for pre, lst in groups:
            f.write("{%d,%s,%d,%d},\n" % (len(pre), ",".join(str(b) for b in pre.ljust(8, b"\0")), first, len(lst)))
            first += len(lst)
        f.write("};\nstatic const unsigned char kDictTransforms[%d][11] = {\n" % first)
        for pre, lst in groups:
            for tid, typ, suf in lst:
                f.write("{%d,%d,%d,%s},\n" % (tid, typ, len(suf), ",".join(str(b) for b in suf.ljust(8, b"\0"))))
        f.write("};\n#define BRO_DICT_NUM_TR_GROUPS %d\n#define BRO_DICT_NUM_TR %d\n#define BRO_DICT_NUM_LUT_ENTRIES %d\n" % (len(groups), first, len(entries)))
    return OUT


if __name__ == "__main__":
    print(main(sys.argv[1] if len(sys.argv) > 1 else "index"))
