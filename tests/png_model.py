"""CPU model (test infrastructure) of the byte stream ascent_b200/csrc/png.cu produces, written against RFC 1950 /
RFC 1951 / the PNG specification: Sub-filtered scanlines; per scanline ONE fixed-Huffman deflate block whose tokens
are found independently in 256 equal segments (a literal, then distance-1 matches of up to 258 for every run of equal
bytes), closed by an empty stored block; a final empty stored block; Adler-32; chunk CRCs.  The GPU test demands the
library's file to equal this model byte for byte, and both to decode (zlib, PIL) to the input pixels."""
import struct
import zlib

import numpy as np

THREADS = 256
LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]


def _rev(v, n):
    r = 0
    for i in range(n):
        r = (r << 1) | ((v >> i) & 1)
    return r


def literal_code(v):
    return (_rev(0x30 + v, 8), 8) if v < 144 else (_rev(0x190 + (v - 144), 9), 9)


def length_symbol(L):
    """(index into LEN_BASE, extra bits, base) by the closed form the kernel uses; checked against the RFC's table"""
    d = L - 3
    if L == 258:
        k, eb, base = 28, 0, 258
    elif d < 8:
        k, eb, base = d, 0, L
    else:
        eb = d.bit_length() - 1 - 2
        k = 4 * eb + 4 + ((d >> eb) & 3)
        base = 3 + ((4 + (k & 3)) << eb)
    return k, eb, base


for _L in range(3, 259):
    _k, _eb, _base = length_symbol(_L)
    assert LEN_BASE[_k] == _base and LEN_EXTRA[_k] == _eb and _base <= _L and (_k == 28 or _L < LEN_BASE[_k + 1])


def match_code(L):
    k, eb, base = length_symbol(L)
    sym = 257 + k
    c, n = (_rev(sym - 256, 7), 7) if sym < 280 else (_rev(0xC0 + (sym - 280), 8), 8)
    c |= (L - base) << n
    return c, n + eb + 5  # + distance code 0 (distance 1)


def segment_tokens(seg):
    out = []
    i, n = 0, len(seg)
    while i < n:
        b, L = seg[i], 1
        while i + L < n and seg[i + L] == b:
            L += 1
        out.append(literal_code(b))
        rem = L - 1
        while rem >= 3:
            m = min(rem, 258)
            out.append(match_code(m))
            rem -= m
        out += [literal_code(b)] * rem
        i += L
    return out


def encode_row(f):
    n = len(f)
    seg = (n + THREADS - 1) // THREADS
    bits, nb = 2, 3  # BFINAL = 0, BTYPE = 01
    for t in range(THREADS):
        for c, k in segment_tokens([int(x) for x in f[t * seg:min((t + 1) * seg, n)]]):
            bits |= c << nb
            nb += k
    nb += 7 + 3  # end of block, header of the empty stored block
    return bits.to_bytes((nb + 7) // 8, "little") + b"\x00\x00\xff\xff"


def encode(rgba):
    """rgba: (H, W, 4) uint8, row 0 = first scanline of the file -> (png bytes, filtered scanlines, zlib stream)"""
    H, W, _ = rgba.shape
    rows, raw = [], b""
    for r in range(H):
        line = rgba[r].reshape(-1).astype(np.int32)
        prev = np.concatenate([np.zeros(4, np.int32), line[:-4]])
        f = np.concatenate([[1], (line - prev) & 255]).astype(np.uint8)
        raw += f.tobytes()
        rows.append(encode_row(f))
    z = b"\x78\x01" + b"".join(rows) + b"\x01\x00\x00\xff\xff" + struct.pack(">I", zlib.adler32(raw))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))

    png = (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 6, 0, 0, 0)) + chunk(b"IDAT", z) +
           chunk(b"IEND", b""))
    return png, raw, z
