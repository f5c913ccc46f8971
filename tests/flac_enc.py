"""A small FLAC *encoder* used only by the tests of the native decoder (csrc/flac_decode.cpp): written from RFC 9639 with a
different structure than the decoder (string-free bit accumulator, numpy residuals) so that the two do not share mistakes;
the decoder itself is anchored on the RFC's published example files (tests/test_host.py).  Not a product component: it makes
no attempt at good compression, it only has to reach every syntax element the decoder implements."""
import hashlib
import struct

import numpy as np


class BitWriter:
    def __init__(self):
        self.acc, self.nbits, self.out = 0, 0, bytearray()

    def put(self, value, bits):
        if bits == 0:
            return
        self.acc = (self.acc << bits) | (int(value) & ((1 << bits) - 1))
        self.nbits += bits
        while self.nbits >= 8:
            self.nbits -= 8
            self.out.append((self.acc >> self.nbits) & 0xFF)
        self.acc &= (1 << self.nbits) - 1

    def unary(self, zeros):
        while zeros >= 32:
            self.put(0, 32)
            zeros -= 32
        self.put(1, zeros + 1)

    def pad(self):
        if self.nbits:
            self.put(0, 8 - self.nbits)

    def bytes(self):
        assert self.nbits == 0
        return bytes(self.out)


def crc8(data):
    c = 0
    for x in data:
        c ^= x
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def crc16(data):
    c = 0
    for x in data:
        c ^= x << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


def _utf8_number(n):
    """The "UTF-8-like" coded number of the frame header (RFC 9639 section 9.1.5): 1 byte below 2^7, else L = 2..7 bytes holding
    5 (L - 1) + 6 ... bits: lead byte = L one-bits, a zero, then the top bits; every following byte 10xxxxxx."""
    if n < 0x80:
        return bytes([n])
    for length, bits in ((2, 11), (3, 16), (4, 21), (5, 26), (6, 31), (7, 36)):
        if n < (1 << bits):
            break
    lead = ((0xFF << (8 - length)) & 0xFF) | (n >> (6 * (length - 1)))
    return bytes([lead] + [0x80 | ((n >> (6 * i)) & 0x3F) for i in range(length - 2, -1, -1)])


FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _residual(x, coefs, shift):
    order = len(coefs)
    res = np.zeros(len(x), dtype=np.int64)
    for i in range(order, len(x)):
        acc = sum(int(coefs[j]) * int(x[i - 1 - j]) for j in range(order))
        res[i] = int(x[i]) - (acc >> shift)
    return res[order:]


def _write_residual(bw, res, blocksize, order, porder, method, escape_partitions=()):
    bw.put(method, 2)
    bw.put(porder, 4)
    pbits, esc = (4, 15) if method == 0 else (5, 31)
    pos = 0
    for p in range(1 << porder):
        count = (blocksize >> porder) - (order if p == 0 else 0)
        part = res[pos:pos + count]
        pos += count
        u = np.where(part >= 0, 2 * part, -2 * part - 1).astype(np.int64)
        if p in escape_partitions:
            raw = 0 if not np.any(part) else 1 + max(int(v).bit_length() if v >= 0 else int(-v - 1).bit_length() for v in part)
            bw.put(esc, pbits)
            bw.put(raw, 5)
            for v in part:
                bw.put(int(v), raw)
            continue
        best_k, best_bits = 0, None
        for k in range(esc):
            bits = int(np.sum((u >> k) + 1 + k)) if count else 0
            if best_bits is None or bits < best_bits:
                best_k, best_bits = k, bits
        bw.put(best_k, pbits)
        for v in u:
            bw.unary(int(v) >> best_k)
            bw.put(int(v) & ((1 << best_k) - 1), best_k)
    assert pos == len(res)


def _write_subframe(bw, x, bps, kind, porder, method, escape_partitions, allow_wasted):
    x = np.asarray(x, dtype=np.int64)
    wasted = 0
    if allow_wasted and np.any(x):
        while not np.any(x & ((1 << (wasted + 1)) - 1)) and wasted + 1 < bps:
            wasted += 1
    if wasted:
        x = x >> wasted
        bps -= wasted

    def header(type_code):
        bw.put(0, 1)
        bw.put(type_code, 6)
        bw.put(1 if wasted else 0, 1)
        if wasted:
            bw.unary(wasted - 1)

    if kind == "constant":
        assert np.all(x == x[0])
        header(0)
        bw.put(int(x[0]), bps)
    elif kind == "verbatim":
        header(1)
        for v in x:
            bw.put(int(v), bps)
    elif kind.startswith("fixed"):
        order = int(kind[5:])
        header(8 + order)
        for v in x[:order]:
            bw.put(int(v), bps)
        _write_residual(bw, _residual(x, FIXED[order], 0), len(x), order, porder, method, escape_partitions)
    elif kind.startswith("lpc"):
        order = int(kind[3:])
        precision, shift = 12, 9
        # least-squares predictor on the block itself, quantised to `precision` bits at 2^-shift
        xf = x.astype(np.float64)
        rows = np.stack([xf[order - 1 - j:len(x) - 1 - j] for j in range(order)], axis=1)
        sol = np.linalg.lstsq(rows, xf[order:], rcond=None)[0] if len(x) > 2 * order else np.zeros(order)
        coefs = np.clip(np.round(sol * (1 << shift)), -(1 << (precision - 1)), (1 << (precision - 1)) - 1).astype(np.int64)
        header(32 + order - 1)
        for v in x[:order]:
            bw.put(int(v), bps)
        bw.put(precision - 1, 4)
        bw.put(shift, 5)
        for c in coefs:
            bw.put(int(c), precision)
        _write_residual(bw, _residual(x, coefs, shift), len(x), order, porder, method, escape_partitions)
    else:
        raise ValueError(kind)


def encode(samples, bps=16, rate=16000, blocksize=4096, kind="fixed2", stereo=None, porder=0, method=0, escape_partitions=(),
           allow_wasted=True, with_md5=True, extra_metadata=True):
    """samples: int array [n] or [n, channels].  stereo: None (independent) | 8 (left/side) | 9 (side/right) | 10 (mid/side)."""
    x = np.asarray(samples, dtype=np.int64)
    x = x[:, None] if x.ndim == 1 else x
    n, ch = x.shape
    frames, sizes = [], []
    for fno, start in enumerate(range(0, n, blocksize)):
        blk = x[start:start + blocksize]
        bs = len(blk)
        po = porder if (bs % (1 << porder) == 0 and (bs >> porder) > 4 + 32 * kind.startswith("lpc")) else 0
        hdr = BitWriter()
        hdr.put(0xFFF8 >> 1, 15)
        hdr.put(0, 1)
        table = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13, 16384: 14, 32768: 15}
        bs_code = table.get(bs, 6 if bs <= 256 else 7)
        hdr.put(bs_code, 4)
        hdr.put(0, 4)                                                   # sample rate: see STREAMINFO
        hdr.put((ch - 1) if stereo is None else stereo, 4)
        hdr.put({8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}.get(bps, 0), 3)
        hdr.put(0, 1)
        for byte in _utf8_number(fno):
            hdr.put(byte, 8)
        if bs_code == 6:
            hdr.put(bs - 1, 8)
        elif bs_code == 7:
            hdr.put(bs - 1, 16)
        head = hdr.bytes()
        bw = BitWriter()
        k = kind if bs > (int(kind[3:]) if kind.startswith("lpc") else 4) else "verbatim"
        if stereo is None:
            chans = [(blk[:, c], bps) for c in range(ch)]
        else:
            assert ch == 2
            left, right = blk[:, 0], blk[:, 1]
            side = left - right
            chans = {8: [(left, bps), (side, bps + 1)], 9: [(side, bps + 1), (right, bps)], 10: [((left + right) >> 1, bps), (side, bps + 1)]}[stereo]
        for data, b in chans:
            kk = "constant" if (k == "constant" and np.all(data == data[0])) else ("verbatim" if k == "constant" else k)
            _write_subframe(bw, data, b, kk, po, method, escape_partitions, allow_wasted)
        bw.pad()
        body = head + bytes([crc8(head)]) + bw.bytes()
        frame = body + struct.pack(">H", crc16(body))
        frames.append(frame)
        sizes.append(len(frame))
    nbytes = (bps + 7) // 8
    pcm = b"".join(int(v).to_bytes(nbytes, "little", signed=True) for v in x.reshape(-1))
    md5 = hashlib.md5(pcm).digest() if with_md5 else bytes(16)
    si = BitWriter()
    si.put(blocksize, 16); si.put(blocksize, 16)
    si.put(min(sizes) if sizes else 0, 24); si.put(max(sizes) if sizes else 0, 24)
    si.put(rate, 20); si.put(ch - 1, 3); si.put(bps - 1, 5); si.put(n, 36)
    out = b"fLaC" + bytes([0x00 if extra_metadata else 0x80]) + (34).to_bytes(3, "big") + si.bytes() + md5
    if extra_metadata:
        vendor = b"slsb200 test encoder"
        vc = struct.pack("<I", len(vendor)) + vendor + struct.pack("<I", 0)
        out += bytes([0x04]) + len(vc).to_bytes(3, "big") + vc
        out += bytes([0x81]) + (10).to_bytes(3, "big") + bytes(10)
    return out + b"".join(frames)
