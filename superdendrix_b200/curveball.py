"""Drop-in for the reference's ``src/curveball.py``: same function names, arguments
and return types; the trades run in the sm_100a kernels of libsdx.so.

    find_presences(input_matrix)                      src/curveball.py:7-14
    curve_ball(input_matrix, r_hp, num_iterations=-1) src/curveball.py:17-40

As in the reference, ``curve_ball`` continues from the state held in ``r_hp``
and leaves the new state there (the null driver's Markov chain,
src/generate_null_matrices.py:71-74).  Random draws come from the library's
counter-based streams (seed set with :func:`seed`, one null index per call), not
from CPython's Mersenne Twister, so outputs agree with the reference in
distribution, and bit for bit under a replayed schedule (``Engine.curveball_replay``).

This per-call shim uploads the state on every call; bulk generation should use
``Engine.curveball`` / ``Engine.null_pvalue`` (one launch for thousands of nulls).
"""
from __future__ import annotations

import numpy as np

from . import bitpack
from .engine import Engine

_state = {"engine": None, "device": 0, "seed": 1764, "next_null": 0}


def seed(value: int, first_null: int = 0) -> None:
    """Plays the role of ``random.seed(args.random_seed)`` (src/generate_null_matrices.py:48)."""
    _state["seed"] = int(value)
    _state["next_null"] = int(first_null)


def set_device(device: int) -> None:
    if _state["engine"] is not None:
        _state["engine"].close()
        _state["engine"] = None
    _state["device"] = int(device)


def _engine() -> Engine:
    if _state["engine"] is None:
        _state["engine"] = Engine(_state["device"])        # raises without a GPU: no CPU fallback
    return _state["engine"]


def find_presences(input_matrix):
    """Adjacency lists over the shorter dimension (src/curveball.py:7-14)."""
    m = np.asarray(input_matrix)
    n_rows, n_cols = m.shape
    view = m if n_cols >= n_rows else m.T
    return [list(np.flatnonzero(view[r] == 1)) for r in range(view.shape[0])]


def _pack_lists(r_hp, n_bits: int) -> np.ndarray:
    dense = np.zeros((len(r_hp), n_bits), dtype=bool)
    for r, cols in enumerate(r_hp):
        dense[r, np.asarray(cols, dtype=np.int64)] = True
    return bitpack.pack_rows(dense)


def curve_ball(input_matrix, r_hp, num_iterations=-1):
    """5*min(R, C) pair trades (or ``num_iterations``) starting from ``r_hp``; returns the
    int8 matrix in the orientation of ``input_matrix`` and stores the new lists in ``r_hp``."""
    n_rows, n_cols = np.asarray(input_matrix).shape
    short, long_ = min(n_rows, n_cols), max(n_rows, n_cols)
    if len(r_hp) != short:
        raise ValueError("r_hp must hold one list per row of the shorter dimension")
    eng = _engine()
    state = _pack_lists(r_hp, long_)                        # rows over the longer dimension
    # the library wants feature rows over samples; give it a matrix whose trade orientation is `state`
    if n_cols >= n_rows:       # rows of input are the trade rows: act as samples x features with F >= S
        dense = bitpack.unpack_rows(state, long_)           # (short, long) = (S, F)
        eng.upload_matrix(bitpack.pack_rows(dense.T), short)
    else:                      # columns of input are the trade rows: features over samples, F < S
        eng.upload_matrix(state, long_)
    R, L, wpr, _ = eng.trade_shape()
    assert (R, L) == (short, long_)
    null_id = _state["next_null"]
    _state["next_null"] += 1
    rows, _ = eng.curveball(_state["seed"], null_id, 1, n_trades=int(num_iterations), chain_len=1, want_applied=False)
    bits = bitpack.unpack_rows(rows[0], long_)              # (short, long)
    for r in range(short):
        r_hp[r] = list(np.flatnonzero(bits[r]))
    out = bits.astype(np.int8)
    return out if n_cols >= n_rows else out.T
