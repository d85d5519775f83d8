"""Development tool: a small RFC 7932 decoder that accounts where the bits of a brotli stream go (per metablock: header
bits, block-switch bits, literal / command / distance symbol and extra bits, number of commands / literals, block types,
context-map clusters, NPOSTFIX / NDIRECT).  Used to compare this library's streams with libbrotlienc's stage by stage.
Not part of the product; window-only (no static dictionary text reconstruction: dictionary copies are counted, their bytes
are skipped)."""
import sys

K_CL_ORDER = [1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15]
INS_BASE = [0, 1, 2, 3, 4, 5, 6, 8, 10, 14, 18, 26, 34, 50, 66, 98, 130, 194, 322, 578, 1090, 2114, 6210, 22594]
INS_EXTRA = [0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 12, 14, 24]
COPY_BASE = [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 18, 22, 30, 38, 54, 70, 102, 134, 198, 326, 582, 1094, 2118]
COPY_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 24]
BL_OFFSET = [1, 5, 9, 13, 17, 25, 33, 41, 49, 65, 81, 97, 113, 145, 177, 209, 241, 305, 369, 497, 753, 1265, 2289, 4337, 8433, 16625]
BL_NBITS = [2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12, 13, 24]


class BR:
    def __init__(self, data):
        self.d = data
        self.pos = 0

    def bits(self, n):
        v = 0
        for i in range(n):
            v |= ((self.d[self.pos >> 3] >> (self.pos & 7)) & 1) << i
            self.pos += 1
        return v


class Code:
    """canonical prefix code from code lengths; decode bit by bit (LSB-first stream, codes MSB-first)"""

    def __init__(self, lengths):
        self.table = {}
        code = 0
        nz = [(l, s) for s, l in enumerate(lengths) if l]
        if len(nz) == 1:
            self.single = nz[0][1]
            return
        self.single = None
        bl = [0] * 17
        for l, _ in nz:
            bl[l] += 1
        nxt = [0] * 17
        for i in range(1, 17):
            code = (code + bl[i - 1]) << 1
            nxt[i] = code
        for s, l in enumerate(lengths):
            if l:
                self.table[(l, nxt[l])] = s
                nxt[l] += 1

    def read(self, br):
        if self.single is not None:
            return self.single
        code, l = 0, 0
        while True:
            code = (code << 1) | br.bits(1)
            l += 1
            s = self.table.get((l, code))
            if s is not None:
                return s
            if l > 15:
                raise ValueError("bad prefix code")


def read_prefix_code(br, alphabet):
    hskip = br.bits(2)
    if hskip == 1:  # simple
        nsym = br.bits(2) + 1
        nb = max(1, (alphabet - 1).bit_length())
        syms = [br.bits(nb) for _ in range(nsym)]
        lengths = [0] * alphabet
        if nsym == 1:
            lengths[syms[0]] = 1  # zero-length code: Code() treats a single symbol specially
            c = Code(lengths)
            c.single = syms[0]
            return c
        if nsym == 2:
            ls = [1, 1]
        elif nsym == 3:
            ls = [1, 2, 2]
        else:
            ls = [1, 2, 3, 3] if br.bits(1) else [2, 2, 2, 2]
        for s, l in zip(syms, ls):
            lengths[s] = l
        return Code(lengths)
    cl = [0] * 18
    space, num = 32, 0
    fixed = Code([2, 4, 3, 2, 2, 4])  # symbols 0..5 with lengths 2,4,3,2,2,4
    for i in range(hskip, 18):
        v = fixed.read(br)
        cl[K_CL_ORDER[i]] = v
        if v:
            space -= 32 >> v
            num += 1
            if space <= 0:
                break
    clc = Code(cl)
    if num == 1:
        clc.single = [i for i, v in enumerate(cl) if v][0]
    lengths = [0] * alphabet
    i, prev, rep, rep_len, space = 0, 8, 0, 0, 32768
    while i < alphabet and space > 0:
        s = clc.read(br)
        if s < 16:
            lengths[i] = s
            i += 1
            rep = 0
            if s:
                prev = s
                space -= 32768 >> s
        else:
            extra = br.bits(2 if s == 16 else 3)
            new_len = prev if s == 16 else 0
            if rep_len != new_len:
                rep, rep_len = 0, new_len
            old = rep
            if rep > 0:
                rep = (rep - 2) << (2 if s == 16 else 3)
            rep += extra + 3
            delta = rep - old
            for _ in range(delta):
                lengths[i] = rep_len
                i += 1
            if rep_len:
                space -= delta * (32768 >> rep_len)
    return Code(lengths)


def read_block_len(br, code):
    c = code.read(br)
    return BL_OFFSET[c] + br.bits(BL_NBITS[c])


def read_context_map(br, size, ntrees_out):
    n = br.bits(1)
    if n:
        nb = br.bits(3)
        n = (1 << nb) + br.bits(nb) + 1
    else:
        n = 1
    ntrees_out.append(n)
    if n == 1:
        return [0] * size
    rlemax = br.bits(4) + 1 if br.bits(1) else 0
    code = read_prefix_code(br, n + rlemax)
    cm = []
    while len(cm) < size:
        s = code.read(br)
        if s == 0:
            cm.append(0)
        elif s <= rlemax:
            cm.extend([0] * ((1 << s) + br.bits(s)))
        else:
            cm.append(s - rlemax)
    cm = cm[:size]
    if br.bits(1):  # inverse MTF
        mtf = list(range(256))
        for i, v in enumerate(cm):
            x = mtf[v]
            cm[i] = x
            del mtf[v]
            mtf.insert(0, x)
    return cm


def analyze(data):
    br = BR(data)
    if br.bits(1) == 0:
        lgwin = 16
    else:
        n = br.bits(3)
        if n:
            lgwin = 17 + n
        else:
            n = br.bits(3)
            lgwin = 17 if n == 0 else 8 + n
    out_len = 0
    mbs = []
    while True:
        st = {"hdr": 0, "lit_bits": 0, "cmd_bits": 0, "cmd_extra": 0, "dist_bits": 0, "dist_extra": 0, "switch_bits": 0, "ncmd": 0,
              "nlit": 0, "ndist_sym": 0, "ndict": 0}
        p0 = br.pos
        islast = br.bits(1)
        if islast and br.bits(1):
            break
        mn = br.bits(2)
        if mn == 3:
            br.bits(1)
            sk = br.bits(2)
            ln = br.bits(8 * sk) + 1 if sk else 0
            br.pos = (br.pos + 7) & ~7
            br.pos += 8 * ln
            continue
        mlen = br.bits(4 * (mn + 4)) + 1
        if not islast and br.bits(1):
            br.pos = (br.pos + 7) & ~7
            br.pos += 8 * mlen
            out_len += mlen
            mbs.append({"raw": mlen})
            continue
        ntypes, btcode, blcode, blen, btype, prev_types = [], [], [], [], [], []
        for cat in range(3):
            n = br.bits(1)
            if n:
                nb = br.bits(3)
                n = (1 << nb) + br.bits(nb) + 1
            else:
                n = 1
            ntypes.append(n)
            if n > 1:
                btcode.append(read_prefix_code(br, n + 2))
                blcode.append(read_prefix_code(br, 26))
                blen.append(read_block_len(br, blcode[cat]))
            else:
                btcode.append(None)
                blcode.append(None)
                blen.append(1 << 28)
            btype.append(0)
            prev_types.append([0, 1])
        npostfix = br.bits(2)
        ndirect = br.bits(4) << npostfix
        cmodes = [br.bits(2) for _ in range(ntypes[0])]
        nl, nd = [], []
        lcm = read_context_map(br, 64 * ntypes[0], nl)
        dcm = read_context_map(br, 4 * ntypes[2], nd)
        lcodes = [read_prefix_code(br, 256) for _ in range(nl[0])]
        ccodes = [read_prefix_code(br, 704) for _ in range(ntypes[1])]
        dalpha = 16 + ndirect + (48 << npostfix)
        dcodes = [read_prefix_code(br, dalpha) for _ in range(nd[0])]
        st["hdr"] = br.pos - p0
        st.update({"mlen": mlen, "ntypes": ntypes, "lit_trees": nl[0], "dist_trees": nd[0], "npostfix": npostfix, "ndirect": ndirect,
                   "cmodes": sorted(set(cmodes))})

        def switch(cat):
            p = br.pos
            c = btcode[cat].read(br)
            pt = prev_types[cat]
            t = pt[1] if c == 0 else ((pt[0] + 1) % ntypes[cat] if c == 1 else c - 2)
            if t >= ntypes[cat]:
                t -= ntypes[cat]
            prev_types[cat] = [t, pt[0]]
            btype[cat] = t
            blen[cat] = read_block_len(br, blcode[cat])
            st["switch_bits"] += br.pos - p

        produced = 0
        while produced < mlen:
            if blen[1] == 0:
                switch(1)
            blen[1] -= 1
            p = br.pos
            cs = ccodes[btype[1]].read(br)
            st["cmd_bits"] += br.pos - p
            cell = cs >> 6
            ic_hi = [0, 0, 0, 0, 8, 8, 0, 16, 8, 16, 16][cell]  # insert code offset
            cc_hi = [0, 8, 0, 8, 0, 8, 16, 0, 16, 8, 16][cell]  # copy code offset
            icode = ic_hi + ((cs >> 3) & 7)
            ccode = cc_hi + (cs & 7)
            p = br.pos
            ins = INS_BASE[icode] + br.bits(INS_EXTRA[icode])
            cpy = COPY_BASE[ccode] + br.bits(COPY_EXTRA[ccode])
            st["cmd_extra"] += br.pos - p
            st["ncmd"] += 1
            for _ in range(ins):
                if blen[0] == 0:
                    switch(0)
                blen[0] -= 1
                p = br.pos
                # the context needs previous bytes: we do not reconstruct the text, so use the cost of decoding with the right
                # tree only when the map is trivial; otherwise reconstruct text lazily
                raise_needed = False
                lit_ctx = CTX.get_context(cmodes[btype[0]])
                tree = lcm[64 * btype[0] + lit_ctx]
                b = lcodes[tree].read(br)
                CTX.push(b)
                st["lit_bits"] += br.pos - p
                st["nlit"] += 1
            produced += ins
            if produced >= mlen:
                break
            if cs < 128:
                dist_code = 0
            else:
                if blen[2] == 0:
                    switch(2)
                blen[2] -= 1
                p = br.pos
                dctx = 3 if cpy > 4 else cpy - 2
                ds = dcodes[dcm[4 * btype[2] + dctx]].read(br)
                st["dist_bits"] += br.pos - p
                st["ndist_sym"] += 1
                dist_code = ds
            # resolve distance
            if dist_code < 16:
                ring = CTX.ring
                idx = [0, 1, 2, 3, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1][dist_code]
                off = [0, 0, 0, 0, -1, 1, -2, 2, -3, 3, -1, 1, -2, 2, -3, 3][dist_code]
                dist = ring[idx] + off
                if dist_code != 0:
                    pass
            elif dist_code < 16 + ndirect:
                dist = dist_code - 15
            else:
                p = br.pos
                x = dist_code - ndirect - 16
                nbits = 1 + (x >> (npostfix + 1))
                hcode = x >> npostfix
                lcode = x & ((1 << npostfix) - 1)
                offset = ((2 + (hcode & 1)) << nbits) - 4
                dist = ((offset + br.bits(nbits)) << npostfix) + lcode + ndirect + 1
                st["dist_extra"] += br.pos - p
            max_dist = min(CTX.total, (1 << lgwin) - 16)
            if dist > max_dist:
                st["ndict"] += 1
                CTX.dict_copy(cpy, dist - max_dist - 1)
                produced += CTX.last_dict_len
            else:
                if dist_code != 0:
                    CTX.ring = [dist] + CTX.ring[:3]
                CTX.copy(dist, cpy)
                produced += cpy
        out_len += mlen
        st["total_bits"] = br.pos - p0
        mbs.append(st)
        if islast:
            break
    return lgwin, out_len, mbs


class Ctx:
    """text reconstruction for literal contexts"""

    def __init__(self):
        self.buf = bytearray()
        self.ring = [4, 11, 15, 16]
        self.total = 0
        self.last_dict_len = 0
        self._dict = None

    LUT0 = None

    def get_context(self, mode):
        p1 = self.buf[-1] if len(self.buf) >= 1 else 0
        p2 = self.buf[-2] if len(self.buf) >= 2 else 0
        if mode == 0:
            return p1 & 0x3f
        if mode == 1:
            return p1 >> 2
        if mode == 2:
            return utf8_lut0(p1) | utf8_lut1(p2)

        def s(c):
            return 0 if c == 0 else 1 if c < 16 else 2 if c < 64 else 3 if c < 128 else 4 if c < 192 else 5 if c < 240 else 6 if c < 255 else 7
        return (s(p1) << 3) | s(p2)

    def push(self, b):
        self.buf.append(b)
        self.total += 1

    def copy(self, dist, n):
        for _ in range(n):
            self.buf.append(self.buf[-dist])
        self.total += n

    def dict_copy(self, copy_len, word_id):
        import ctypes
        if self._dict is None:
            class D(ctypes.Structure):
                _fields_ = [("size_bits_by_length", ctypes.c_uint8 * 32), ("offsets_by_length", ctypes.c_uint32 * 32),
                            ("data_size", ctypes.c_size_t), ("data", ctypes.POINTER(ctypes.c_uint8))]
            lib = ctypes.CDLL("libbrotlicommon.so.1")
            lib.BrotliGetDictionary.restype = ctypes.POINTER(D)
            d = lib.BrotliGetDictionary().contents
            self._dict = (bytes(d.data[:d.data_size]), list(d.size_bits_by_length), list(d.offsets_by_length))
            self._lib = lib
        data, bits, offs = self._dict
        nb = bits[copy_len]
        idx = word_id & ((1 << nb) - 1)
        tr = word_id >> nb
        w = data[offs[copy_len] + idx * copy_len: offs[copy_len] + (idx + 1) * copy_len]
        out = ctypes.create_string_buffer(64)
        lib = ctypes.CDLL("libbrotlicommon.so.1")
        # BrotliTransformDictionaryWord(dst, word, len, transforms, transform_idx)
        lib.BrotliGetTransforms.restype = ctypes.c_void_p
        lib.BrotliTransformDictionaryWord.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        n = lib.BrotliTransformDictionaryWord(out, w, copy_len, lib.BrotliGetTransforms(), tr)
        self.buf += out.raw[:n]
        self.total += n
        self.last_dict_len = n


def utf8_lut0(c):
    a = [0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 0, 0, 4, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 12, 16, 12, 12, 20, 12, 16, 24, 28, 12, 12, 32,
         12, 36, 12, 44, 44, 44, 44, 44, 44, 44, 44, 44, 44, 32, 32, 24, 40, 28, 12, 12, 48, 52, 52, 52, 48, 52, 52, 52, 48, 52, 52, 52, 52, 52, 48,
         52, 52, 52, 52, 52, 48, 52, 52, 52, 52, 52, 24, 12, 28, 12, 12, 12, 56, 60, 60, 60, 56, 60, 60, 60, 56, 60, 60, 60, 60, 60, 56, 60, 60, 60,
         60, 60, 56, 60, 60, 60, 60, 60, 24, 12, 28, 12, 0]
    if c < 128:
        return a[c]
    if c < 192:
        return c & 1
    return 2 + (c & 1)


def utf8_lut1(c):
    if c < 32:
        return 0
    if c < 128:
        if c in (32, 127):
            return 0
        if 48 <= c <= 57 or 65 <= c <= 90:
            return 2
        if 97 <= c <= 122:
            return 3
        return 1
    if c < 224:
        return 0
    return 2


CTX = None


def stats(data):
    global CTX
    CTX = Ctx()
    lgwin, n, mbs = analyze(data)
    return lgwin, n, mbs, bytes(CTX.buf)


def summarize(data):
    lgwin, n, mbs, text = stats(data)
    tot = {}
    for m in mbs:
        for k, v in m.items():
            if isinstance(v, int):
                tot[k] = tot.get(k, 0) + v
    return lgwin, n, mbs, tot, text


if __name__ == "__main__":
    d = open(sys.argv[1], "rb").read()
    lgwin, n, mbs, tot, text = summarize(d)
    print("lgwin", lgwin, "decoded", n, "metablocks", len(mbs))
    for m in mbs[:4]:
        print({k: v for k, v in m.items()})
    print({k: (v // 8 if "bits" in k or k in ("hdr", "cmd_extra", "dist_extra") else v) for k, v in tot.items()})
