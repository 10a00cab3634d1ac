"""Symbols <-> codes and packed batches.

Code = index in 'AGCUYRWSKMDVHBN' (IRMethods.py:13; the order of costs.json).  Code 15 is a spare
symbol that only matches itself (the reference never looks equal characters up, SED:79-81).
Packed batch layout: see include/rsd.h ("Symbols and packing")."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

SYMBOLS = "AGCUYRWSKMDVHBN"
_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate(SYMBOLS):
    _LUT[ord(_ch)] = _i
_DEC = np.frombuffer((SYMBOLS + "?").encode(), dtype=np.uint8)


def encode(s: str) -> np.ndarray:
    """str of table symbols -> uint8 codes.  KeyError(symbol) for anything else, like the dict
    lookup the reference would perform (SED:87)."""
    raw = np.frombuffer(s.encode("latin-1", "replace"), dtype=np.uint8)
    codes = _LUT[raw]
    bad = np.nonzero(codes == 255)[0]
    if bad.size:
        raise KeyError(s[int(bad[0])])
    return codes


def decode(codes) -> str:
    return _DEC[np.asarray(codes, dtype=np.uint8)].tobytes().decode()


def concat(seqs):
    """list[str] -> (codes uint8, offsets int64[n+1]) with one LUT pass over the joined text."""
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    joined = "".join(seqs)
    raw = np.frombuffer(joined.encode("latin-1", "replace"), dtype=np.uint8)
    codes = _LUT[raw]
    bad = np.nonzero(codes == 255)[0]
    if bad.size:
        raise KeyError(joined[int(bad[0])])
    return codes, off


@dataclass
class PackedSeqs:
    words: np.ndarray    # uint32, rsd_pack layout (+4 padding words)
    start: np.ndarray    # int64 [n]  word index
    len: np.ndarray      # int32 [n]
    bits: int            # 2 or 4
    symmask: int         # bit c set = code c present
    canonical: bool = False   # rsd_pack's layout (start[p] = words of the sequences before p): start[] can stay on the host

    @property
    def n(self) -> int:
        return int(self.len.shape[0])

    @property
    def max_len(self) -> int:
        return int(self.len.max()) if self.len.size else 0

    def repack(self, bits: int) -> "PackedSeqs":
        if bits == self.bits:
            return self
        codes, off = unpack(self)
        return pack((codes, off), bits=bits)


def pack(seqs, bits: int | None = None) -> PackedSeqs:
    """seqs: list[str] or (codes uint8, offsets int64[n+1]).  bits=None picks 2 when only ACGU occur."""
    lib = _lib.load_library()
    codes, off = seqs if isinstance(seqs, tuple) else concat(seqs)
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = off.shape[0] - 1
    if bits is None:
        bits = 2 if (codes.size == 0 or int(codes.max(initial=0)) < 4) else 4
    lens = np.diff(off).astype(np.int32)
    nw = lib.rsd_pack_words(_lib.ptr(lens, C.c_int32), n, bits)
    if nw < 0:
        raise _lib.RsdError(_lib.RSD_EINVAL, "rsd_pack_words failed")
    words = np.zeros(nw, dtype=np.uint32)
    start = np.zeros(max(n, 1), dtype=np.int64)[:n]
    out_len = np.zeros(max(n, 1), dtype=np.int32)[:n]
    mask = C.c_uint32(0)
    if codes.size == 0:
        codes = np.zeros(1, dtype=np.uint8)
    _lib.check(lib.rsd_pack(_lib.ptr(codes, C.c_uint8), _lib.ptr(off, C.c_int64), n, bits,
                            _lib.ptr(words, C.c_uint32), _lib.ptr(start, C.c_int64),
                            _lib.ptr(out_len, C.c_int32), C.byref(mask)))
    return PackedSeqs(words, start, out_len, bits, int(mask.value), canonical=True)


def unpack(p: PackedSeqs):
    """PackedSeqs -> (codes uint8, offsets int64[n+1]) (numpy, for tests and repacking)."""
    per = 32 // p.bits
    off = np.zeros(p.n + 1, dtype=np.int64)
    np.cumsum(p.len, out=off[1:])
    codes = np.zeros(int(off[-1]), dtype=np.uint8)
    shifts = (np.arange(per, dtype=np.uint32) * p.bits)
    maskv = np.uint32((1 << p.bits) - 1)
    for i in range(p.n):
        L = int(p.len[i])
        if not L:
            continue
        nw = (L + per - 1) // per
        w = p.words[int(p.start[i]): int(p.start[i]) + nw]
        sym = ((w[:, None] >> shifts[None, :]) & maskv).astype(np.uint8).reshape(-1)[:L]
        codes[off[i]:off[i + 1]] = sym
    return codes, off
