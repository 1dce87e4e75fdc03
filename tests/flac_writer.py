"""A small FLAC ENCODER for the tests of the decoder in libwb_audio.so (test infrastructure, not shipped).

Written independently of the decoder, from the format description: it can produce every construct the decoder handles —
CONSTANT / VERBATIM / FIXED (order 0-4) / LPC subframes, partitioned Rice residuals with 4- or 5-bit parameters and escaped
(raw) partitions, wasted bits, the three stereo decorrelation modes, fixed and variable block-size streams, every block-size
and sample-rate header code, 8 / 12 / 16 / 20 / 24 / 32-bit samples, STREAMINFO with the MD5 of the PCM (hashlib), extra
metadata blocks and an ID3v2 tag in front.  No attempt at good compression: the caller picks the constructs.
"""
import hashlib
from typing import List, Optional, Sequence

import numpy as np


class BitWriter:
    def __init__(self):
        self.acc = 0
        self.n = 0

    def write(self, value: int, bits: int):
        if bits == 0:
            return
        assert 0 <= value < (1 << bits), (value, bits)
        self.acc = (self.acc << bits) | value
        self.n += bits

    def write_signed(self, value: int, bits: int):
        assert -(1 << (bits - 1)) <= value < (1 << (bits - 1)), (value, bits)
        self.write(value & ((1 << bits) - 1), bits)

    def write_unary(self, zeros: int):
        self.write(1, zeros + 1)

    def align(self):
        if self.n % 8:
            self.write(0, 8 - self.n % 8)

    def tobytes(self) -> bytes:
        assert self.n % 8 == 0
        return self.acc.to_bytes(self.n // 8, "big")


def crc(data: bytes, poly: int, width: int) -> int:
    c, top, mask = 0, 1 << (width - 1), (1 << width) - 1
    for byte in data:
        c ^= byte << (width - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
    return c


def utf8_number(v: int) -> bytes:
    if v < 0x80:
        return bytes([v])
    for n in range(2, 8):                       # n bytes carry 5n + 1 payload bits (n = 7: 36)
        if v < (1 << (5 * n + 1)) or n == 7:
            lead = (0xff << (8 - n)) & 0xff
            first = lead | (v >> (6 * (n - 1)))
            return bytes([first] + [0x80 | ((v >> (6 * k)) & 0x3f) for k in range(n - 2, -1, -1)])
    raise ValueError(v)


BLOCKSIZE_CODES = {192: 1, 576: 2, 1152: 3, 2304: 4, 4608: 5, 256: 8, 512: 9, 1024: 10, 2048: 11, 4096: 12, 8192: 13, 16384: 14,
                   32768: 15}
RATE_CODES = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}
BITS_CODES = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}
FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _rice_cost(res: Sequence[int], k: int) -> int:
    return sum((((r << 1) ^ (r >> 63)) >> k) + 1 + k for r in res)


def write_residual(w: BitWriter, res: List[int], blocksize: int, order: int, partition_order: int, five_bit: bool, escape: bool):
    while partition_order and (blocksize % (1 << partition_order) or (blocksize >> partition_order) < order):
        partition_order -= 1                 # a short last block: fall back to the largest order that divides it
    w.write(1 if five_bit else 0, 2)
    w.write(partition_order, 4)
    pbits, esc = (5, 31) if five_bit else (4, 15)
    parts = 1 << partition_order
    assert blocksize % parts == 0 and blocksize // parts >= order
    pos = 0
    for p in range(parts):
        count = blocksize // parts - (order if p == 0 else 0)
        chunk = res[pos:pos + count]
        pos += count
        if escape:
            raw = max([0] + [(r if r >= 0 else ~r).bit_length() + 1 for r in chunk])
            if all(r == 0 for r in chunk):
                raw = 0
            w.write(esc, pbits)
            w.write(raw, 5)
            for r in chunk:
                if raw:
                    w.write_signed(r, raw)
        else:
            k = min(range(esc), key=lambda kk: _rice_cost(chunk, kk))
            w.write(k, pbits)
            for r in chunk:
                u = (r << 1) if r >= 0 else ((-r) << 1) - 1
                w.write_unary(u >> k)
                w.write(u & ((1 << k) - 1), k)
    assert pos == len(res)


def write_subframe(w: BitWriter, x: List[int], bits: int, kind, partition_order=0, five_bit=False, escape=False, wasted="auto"):
    """kind: 'constant' | 'verbatim' | ('fixed', order) | ('lpc', coefficients, precision, shift)."""
    n = len(x)
    k = 0
    if wasted == "auto":
        if any(x):
            while all((v >> k) & 1 == 0 for v in x):
                k += 1
        k = min(k, bits - 1)
    elif wasted:
        k = int(wasted)
        assert all(v % (1 << k) == 0 for v in x)
    x = [v >> k for v in x]
    bits -= k
    if kind == "constant":
        assert all(v == x[0] for v in x)
        code = 0
    elif kind == "verbatim":
        code = 1
    elif kind[0] == "fixed":
        code = 8 + kind[1]
    else:
        code = 32 + len(kind[1]) - 1
    w.write(0, 1)
    w.write(code, 6)
    if k:
        w.write(1, 1)
        w.write_unary(k - 1)
    else:
        w.write(0, 1)
    if kind == "constant":
        w.write_signed(x[0], bits)
    elif kind == "verbatim":
        for v in x:
            w.write_signed(v, bits)
    else:
        if kind[0] == "fixed":
            coefs, shift, order = FIXED[kind[1]], 0, kind[1]
        else:
            coefs, precision, shift = list(kind[1]), kind[2], kind[3]
            order = len(coefs)
        for v in x[:order]:
            w.write_signed(v, bits)
        if kind[0] == "lpc":
            w.write(precision - 1, 4)
            w.write_signed(shift, 5)
            for c in coefs:
                w.write_signed(c, precision)
        res = [x[i] - (sum(c * x[i - 1 - j] for j, c in enumerate(coefs)) >> shift) for i in range(order, n)]
        write_residual(w, res, n, order, partition_order, five_bit, escape)


def encode(pcm: np.ndarray, bps: int = 16, rate: int = 16000, blocksize: int = 4096, kind=("fixed", 2), stereo: str = "independent",
           partition_order: int = 0, five_bit: bool = False, escape: bool = False, wasted="auto", variable: bool = False,
           explicit_blocksize: bool = False, rate_mode: str = "code", bits_from_streaminfo: bool = False, md5: bool = True,
           id3: bool = False, padding_block: bool = True, total_samples_known: bool = True, trailer: bytes = b"",
           first_frame_number: int = 0, per_frame: Optional[dict] = None) -> bytes:
    """pcm int [n, C] -> FLAC bytes.  `kind` may be a list with one entry per channel.  `per_frame[i]` overrides keyword
    arguments (kind, partition_order, escape, five_bit, stereo, blocksize) for frame i."""
    pcm = np.asarray(pcm, dtype=np.int64)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    n, C = pcm.shape
    out = bytearray()
    if id3:
        out += b"ID3\x04\x00\x00" + bytes([0, 0, 0, 12]) + b"\x00" * 12
    out += b"fLaC"
    nbytes = (bps + 7) // 8
    raw = b"".join(int(v).to_bytes(nbytes, "little", signed=True) for v in pcm.reshape(-1))
    info = BitWriter()
    info.write(min(blocksize, 65535), 16)
    info.write(min(blocksize, 65535), 16)
    info.write(0, 24)
    info.write(0, 24)
    info.write(rate, 20)
    info.write(C - 1, 3)
    info.write(bps - 1, 5)
    info.write(n if total_samples_known else 0, 36)
    info_bytes = info.tobytes() + (hashlib.md5(raw).digest() if md5 else b"\x00" * 16)
    out += bytes([0x00 if padding_block else 0x80]) + len(info_bytes).to_bytes(3, "big") + info_bytes
    if padding_block:      # a VORBIS_COMMENT-like block and a PADDING block the decoder has to skip
        out += bytes([4]) + (8).to_bytes(3, "big") + b"\x00" * 8
        out += bytes([0x81]) + (5).to_bytes(3, "big") + b"\x00" * 5
    pos, frame = 0, 0
    while pos < n:
        o = dict(kind=kind, partition_order=partition_order, escape=escape, five_bit=five_bit, stereo=stereo, blocksize=blocksize)
        o.update((per_frame or {}).get(frame, {}))
        bs = min(o["blocksize"], n - pos)
        block = pcm[pos:pos + bs]
        w = BitWriter()
        w.write(0x7ffc, 15)
        w.write(1 if variable else 0, 1)
        if not explicit_blocksize and bs in BLOCKSIZE_CODES:
            bs_code = BLOCKSIZE_CODES[bs]
        else:
            bs_code = 6 if bs <= 256 else 7
        w.write(bs_code, 4)
        if rate_mode == "streaminfo":
            rate_code = 0
        elif rate_mode == "code" and rate in RATE_CODES:
            rate_code = RATE_CODES[rate]
        elif rate_mode == "khz" or (rate_mode == "code" and rate % 1000 == 0 and rate // 1000 < 256):
            rate_code = 12
        elif rate_mode == "tens" or (rate % 10 == 0 and rate // 10 < 65536 and rate >= 65536):
            rate_code = 14
        else:
            rate_code = 13
        w.write(rate_code, 4)
        assign = {"independent": C - 1, "left_side": 8, "side_right": 9, "mid_side": 10}[o["stereo"]]
        w.write(assign, 4)
        w.write(0 if bits_from_streaminfo else BITS_CODES[bps], 3)
        w.write(0, 1)
        for byte in utf8_number(pos if variable else frame + first_frame_number):
            w.write(byte, 8)
        if bs_code == 6:
            w.write(bs - 1, 8)
        elif bs_code == 7:
            w.write(bs - 1, 16)
        if rate_code == 12:
            w.write(rate // 1000, 8)
        elif rate_code == 13:
            w.write(rate, 16)
        elif rate_code == 14:
            w.write(rate // 10, 16)
        w.write(crc(w.tobytes(), 0x07, 8), 8)
        chans = [[int(v) for v in block[:, c]] for c in range(C)]
        widths = [bps] * C
        if o["stereo"] != "independent":
            assert C == 2
            left, right = chans
            side = [a - b for a, b in zip(left, right)]
            if o["stereo"] == "left_side":
                chans, widths = [left, side], [bps, bps + 1]
            elif o["stereo"] == "side_right":
                chans, widths = [side, right], [bps + 1, bps]
            else:
                chans, widths = [[(a + b) >> 1 for a, b in zip(left, right)], side], [bps, bps + 1]
        kinds = o["kind"] if isinstance(o["kind"], list) else [o["kind"]] * C
        for c in range(C):
            write_subframe(w, chans[c], widths[c], kinds[c], o["partition_order"], o["five_bit"], o["escape"], wasted)
        w.align()
        body = w.tobytes()
        out += body + crc(body, 0x8005, 16).to_bytes(2, "big")
        pos += bs
        frame += 1
    return bytes(out) + trailer
