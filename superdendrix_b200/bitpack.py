"""Bit-packed row layout shared by the host code and the C ABI (include/sdx.h).

A packed matrix is a C-contiguous ``uint32`` array of shape (n_rows, words_per_row);
bit ``j`` of word ``i`` of a row is element ``32*i + j`` of that row; padding bits
are zero and must stay zero.  Little-endian hosts only.
"""
from __future__ import annotations

import numpy as np


def words_for(n_bits: int) -> int:
    return (int(n_bits) + 31) // 32


def pack_rows(dense) -> np.ndarray:
    """(R, L) 0/1 or bool array -> (R, ceil(L/32)) uint32."""
    d = np.ascontiguousarray(np.asarray(dense) != 0)
    if d.ndim != 2:
        raise ValueError("pack_rows expects a 2-D array")
    R, L = d.shape
    wpr = words_for(L)
    padded = np.zeros((R, wpr * 32), dtype=bool)
    padded[:, :L] = d
    return np.packbits(padded, axis=1, bitorder="little").view("<u4").reshape(R, wpr).copy()


def unpack_rows(packed, n_bits: int) -> np.ndarray:
    """(…, wpr) uint32 -> (…, n_bits) uint8 of 0/1."""
    p = np.ascontiguousarray(packed, dtype="<u4")
    bits = np.unpackbits(p.view(np.uint8).reshape(*p.shape[:-1], -1), axis=-1, bitorder="little")
    return bits[..., :n_bits]


def pack_mask(indices, n_bits: int) -> np.ndarray:
    m = np.zeros(n_bits, dtype=bool)
    m[np.asarray(indices, dtype=np.int64)] = True
    return pack_rows(m[None, :])[0]


def popcount_rows(packed) -> np.ndarray:
    p = np.ascontiguousarray(packed, dtype="<u4")
    return np.unpackbits(p.view(np.uint8).reshape(*p.shape[:-1], -1), axis=-1).sum(axis=-1).astype(np.int64)
