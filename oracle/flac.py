"""Minimal pure-Python FLAC decoder (test infrastructure, oracle side only).

The build image has no FLAC decoder (no soundfile / librosa / libFLAC), and the
reference's only real-audio fixture is ``tests/test.flac`` (16 kHz, mono, 16-bit,
fixed block size).  This decoder implements the subset of the FLAC format needed
for such files (CONSTANT / VERBATIM / FIXED / LPC subframes, Rice partitions,
mono or independent channels) and verifies the STREAMINFO MD5 of the decoded PCM,
which pins the decode bit-exactly.

Used only by ``tests/golden/make_golden.py`` to mint the committed log-mel golden.
"""
from __future__ import annotations

import hashlib
import struct

import numpy as np


class _Bits:
    def __init__(self, data: bytes, pos: int = 0):
        self.d = data
        self.p = pos * 8

    def read(self, n: int) -> int:
        v = 0
        p = self.p
        d = self.d
        while n > 0:
            byte = d[p >> 3]
            avail = 8 - (p & 7)
            take = avail if avail < n else n
            v = (v << take) | ((byte >> (avail - take)) & ((1 << take) - 1))
            p += take
            n -= take
        self.p = p
        return v

    def read_signed(self, n: int) -> int:
        v = self.read(n)
        return v - (1 << n) if v >> (n - 1) else v

    def unary(self) -> int:
        n = 0
        while self.read(1) == 0:
            n += 1
        return n

    def align(self):
        self.p = (self.p + 7) & ~7

    @property
    def byte_pos(self) -> int:
        return self.p >> 3


_FIXED = [[], [1], [2, -1], [3, -3, 1], [4, -6, 4, -1]]


def _residual(br: _Bits, blocksize: int, order: int) -> list:
    method = br.read(2)
    assert method in (0, 1), "reserved residual coding method"
    pbits = 4 if method == 0 else 5
    esc = (1 << pbits) - 1
    porder = br.read(4)
    nparts = 1 << porder
    out = []
    for part in range(nparts):
        n = (blocksize >> porder) - (order if part == 0 else 0) if porder else blocksize - order
        k = br.read(pbits)
        if k == esc:
            nb = br.read(5)
            for _ in range(n):
                out.append(br.read_signed(nb) if nb else 0)
        else:
            for _ in range(n):
                q = br.unary()
                r = br.read(k) if k else 0
                u = (q << k) | r
                out.append((u >> 1) ^ -(u & 1))
    return out


def _subframe(br: _Bits, blocksize: int, bps: int) -> list:
    assert br.read(1) == 0
    typ = br.read(6)
    wasted = 0
    if br.read(1):
        wasted = br.unary() + 1
        bps -= wasted
    if typ == 0:
        samples = [br.read_signed(bps)] * blocksize
    elif typ == 1:
        samples = [br.read_signed(bps) for _ in range(blocksize)]
    elif 8 <= typ <= 12:
        order = typ - 8
        samples = [br.read_signed(bps) for _ in range(order)]
        res = _residual(br, blocksize, order)
        coefs = _FIXED[order]
        for r in res:
            pred = 0
            for j, c in enumerate(coefs):
                pred += c * samples[-1 - j]
            samples.append(pred + r)
    elif typ >= 32:
        order = (typ & 31) + 1
        samples = [br.read_signed(bps) for _ in range(order)]
        prec = br.read(4) + 1
        shift = br.read_signed(5)
        coefs = [br.read_signed(prec) for _ in range(order)]
        res = _residual(br, blocksize, order)
        for r in res:
            pred = 0
            for j, c in enumerate(coefs):
                pred += c * samples[-1 - j]
            samples.append((pred >> shift) + r)
    else:
        raise ValueError(f"reserved subframe type {typ}")
    if wasted:
        samples = [s << wasted for s in samples]
    return samples


def decode_flac(path: str):
    """Returns (pcm int array [n_samples, channels], sample_rate, bits_per_sample)."""
    data = open(path, "rb").read()
    assert data[:4] == b"fLaC"
    pos = 4
    info = None
    while True:
        hdr = data[pos]
        length = int.from_bytes(data[pos + 1 : pos + 4], "big")
        if hdr & 0x7F == 0:
            info = data[pos + 4 : pos + 4 + length]
        pos += 4 + length
        if hdr & 0x80:
            break
    x = int.from_bytes(info[10:18], "big")
    sample_rate = x >> 44
    channels = ((x >> 41) & 7) + 1
    bps = ((x >> 36) & 31) + 1
    total = x & ((1 << 36) - 1)
    md5 = info[18:34]

    chans = [[] for _ in range(channels)]
    br = _Bits(data, pos)
    ndec = 0
    while ndec < total:
        assert br.read(14) == 0x3FFE, "lost frame sync"
        br.read(1)
        br.read(1)  # blocking strategy
        bs_code = br.read(4)
        sr_code = br.read(4)
        ch_code = br.read(4)
        ss_code = br.read(3)
        br.read(1)
        # UTF-8 coded frame/sample number
        first = br.read(8)
        nfollow = 0
        while first & (0x80 >> nfollow):
            nfollow += 1
        for _ in range(max(0, nfollow - 1)):
            br.read(8)
        if bs_code == 1:
            blocksize = 192
        elif 2 <= bs_code <= 5:
            blocksize = 576 << (bs_code - 2)
        elif bs_code == 6:
            blocksize = br.read(8) + 1
        elif bs_code == 7:
            blocksize = br.read(16) + 1
        else:
            blocksize = 256 << (bs_code - 8)
        if sr_code == 12:
            br.read(8)
        elif sr_code in (13, 14):
            br.read(16)
        br.read(8)  # header CRC-8
        fbps = bps if ss_code == 0 else {1: 8, 2: 12, 4: 16, 5: 20, 6: 24}[ss_code]
        subs = []
        if ch_code < 8:
            for _ in range(ch_code + 1):
                subs.append(_subframe(br, blocksize, fbps))
        else:
            # stereo decorrelation
            side_first = ch_code == 9
            a = _subframe(br, blocksize, fbps + (1 if side_first else 0))
            b = _subframe(br, blocksize, fbps + (0 if side_first else 1))
            if ch_code == 8:  # left/side
                subs = [a, [l - s for l, s in zip(a, b)]]
            elif ch_code == 9:  # side/right
                subs = [[s + r for s, r in zip(a, b)], b]
            else:  # mid/side
                left, right = [], []
                for m, s in zip(a, b):
                    m = (m << 1) | (s & 1)
                    left.append((m + s) >> 1)
                    right.append((m - s) >> 1)
                subs = [left, right]
        br.align()
        br.read(16)  # frame CRC-16
        for c in range(channels):
            chans[c].extend(subs[c])
        ndec += blocksize
    pcm = np.array(chans, dtype=np.int64).T[:total]
    if bps == 16:
        raw = pcm.astype("<i2").tobytes()
    elif bps == 24:
        raw = b"".join(struct.pack("<i", int(v))[:3] for v in pcm.reshape(-1))
    else:
        raw = pcm.astype("<i1").tobytes()
    assert hashlib.md5(raw).digest() == md5, "FLAC MD5 mismatch: decode is wrong"
    return pcm, sample_rate, bps
