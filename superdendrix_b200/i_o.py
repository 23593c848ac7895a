"""Loaders with the reference's names, arguments and return shapes (src/i_o.py),
plus the packers that turn their dictionaries into the bit-packed rows the
C ABI consumes.  Host-side boundary code: nothing here runs on the GPU.
"""
from __future__ import annotations

import logging
from collections import defaultdict

import numpy as np

from . import bitpack

FORMAT = "%(asctime)s %(filename)-15s %(levelname)-10s: %(message)s"
logging.basicConfig(format=FORMAT)


def getLogger(verbosity=logging.INFO):
    """src/i_o.py:11-14."""
    logger = logging.getLogger(__name__)
    logger.setLevel(verbosity)
    return logger


def _rows(path):
    with open(path) as fh:
        for line in fh:
            if not line.startswith("#"):
                yield line.rstrip("\n")


def load_mutation_data(filename, patients=None, geneFile=None, mut_only=False, min_freq=0, max_freq=float("inf"),
                       verbose=False):
    """src/i_o.py:18-88.  ``sample<TAB>event<TAB>event...`` ->
    (m, n, genes, patients, geneToCases, patientToGenes).

    Same filters in the same order: sample whitelist, event whitelist, ``_MUT``
    only, then carrier-count window.  The reference's event order comes from
    iterating Python sets (hash-seed dependent); here events are listed in
    order of first appearance in the file, which is one of the orders the
    reference can produce.
    """
    whitelist = None
    if geneFile:
        whitelist = {ln.split()[0] for ln in _rows(geneFile) if ln.strip()}
    gene_cases: dict = {}
    patient_genes: dict = {}
    for ln in _rows(filename):
        fields = ln.rstrip().split("\t")
        who = fields[0]
        if patients and who not in patients:
            continue
        events = dict.fromkeys(fields[1:])             # ordered de-duplication
        if whitelist:
            events = {e: None for e in events if e in whitelist}
        patient_genes[who] = set(events)
        for e in events:
            gene_cases.setdefault(e, set()).add(who)

    def discard(unwanted):
        for e in unwanted:
            for who in gene_cases.pop(e):
                patient_genes[who].discard(e)

    if mut_only:
        gone = [e for e in gene_cases if not e.endswith("MUT")]
        if verbose:
            print("Warning:", len(gone), "events removed because no mutations")
        discard(gone)
    gone = [e for e, who in gene_cases.items() if len(who) < min_freq or len(who) > max_freq]
    if verbose:
        print("Warning:", len(gone), "events removed because of too low/high frequency")
    discard(gone)
    genes, who_all = list(gene_cases), list(patient_genes)
    g2c, p2g = defaultdict(set, gene_cases), defaultdict(set, patient_genes)
    return len(genes), len(who_all), genes, who_all, g2c, p2g


def load_events(mutation_file, verbose=0):
    """src/i_o.py:91-105 -> (eventToCases, mutation_samples)."""
    event_cases, seen = defaultdict(set), set()
    for ln in _rows(mutation_file):
        fields = ln.split("\t")
        seen.add(fields[0])
        for e in fields[1:]:
            event_cases[e].add(fields[0])
    if verbose > 0:
        print("* Loading mutations...")
        print("\t- Events: %s" % len(event_cases))
    return event_cases, seen


def _delim(path):
    return "\t" if path.endswith(".tsv") else ","


def load_profiles(profile_file, sample_whitelist=None, verbose=0):
    """src/i_o.py:109-140 -> ({profile name: fp64 vector}, samples)."""
    rows = [ln.split(_delim(profile_file)) for ln in _rows(profile_file)]
    names = rows[0][1:]
    body = rows[1:]
    ids = [r[0] for r in body]
    mat = np.array([[float(x) if x != "" and x.lower() != "na" else np.nan for x in r[1:]] for r in body], dtype=np.float64)
    if mat.size == 0:
        mat = mat.reshape(len(body), len(names))
    if sample_whitelist is not None:
        keep = [i for i, s in enumerate(ids) if s in sample_whitelist]
        ids = [ids[i] for i in keep]
        mat = mat[keep]
    profiles = {name: mat[:, j] for j, name in enumerate(names)}
    if verbose > 0:
        print("* Loading profiles...")
        print("\t- Number of profiles:", len(profiles))
        print("\t- Samples in common:", len(ids))
    return profiles, ids


def load_single_profile(profile_fn, gene):
    """src/i_o.py:143-163: one column (matched on the part before ``_``), blanks -> 0.0."""
    d = _delim(profile_fn)
    with open(profile_fn) as fh:
        head = [c.split("_")[0] for c in fh.readline().rstrip().split(d)]
        col = head.index(gene)
        out = {}
        for ln in fh:
            if ln.startswith("#"):
                continue
            f = ln.rstrip().split(d)
            out[f[0]] = float(f[col]) if f[col] != "" else 0.0
    return out


# ---------------------------------------------------------------------------
# packers (the layout of include/sdx.h)
# ---------------------------------------------------------------------------
def pack_features(genes, samples, geneToCases) -> np.ndarray:
    """(F, ceil(S/32)) uint32 feature rows over ``samples`` (in the given orders)."""
    pos = {s: i for i, s in enumerate(samples)}
    dense = np.zeros((len(genes), len(samples)), dtype=bool)
    for g, e in enumerate(genes):
        idx = [pos[s] for s in geneToCases[e] if s in pos]
        dense[g, idx] = True
    return bitpack.pack_rows(dense)


def dense_matrix(genes, samples, patientToGenes) -> np.ndarray:
    """The S x F 0/1 matrix of src/generate_null_matrices.py:62-68 (uint8)."""
    col = {e: j for j, e in enumerate(genes)}
    mat = np.zeros((len(samples), len(genes)), dtype=np.uint8)
    for i, s in enumerate(samples):
        js = [col[e] for e in patientToGenes[s] if e in col]
        mat[i, js] = 1
    return mat


def write_sample_events(path, samples, genes, dense) -> None:
    """One null / feature file in the reference's format (src/generate_null_matrices.py:82-86):
    header ``#Samples\\tEvents`` then ``sample\\tevent...`` with events in ``genes`` order."""
    names = np.asarray(genes, dtype=object)
    r, c = np.nonzero(dense)
    cuts = np.searchsorted(r, np.arange(1, len(samples)))
    parts = np.split(names[c], cuts)
    with open(path, "w") as fh:
        fh.write("#Samples\tEvents\n")
        fh.write("".join("\t".join([s, *p]) + "\n" for s, p in zip(samples, parts)))


# ---------------------------------------------------------------------------
# native (C++) writer / reader of the same files: the host-side format path of libsdx.so
# ---------------------------------------------------------------------------
def _names(seq):
    import ctypes as C
    arr = (C.c_char_p * len(seq))()
    arr[:] = [str(x).encode() for x in seq]
    return arr


def write_nulls_tsv(path_prefix, first_index, feature_rows, samples, genes, n_threads=0) -> None:
    """feature_rows: (n, F, ceil(S/32)) uint32.  Writes ``<path_prefix><first_index + i>.tsv`` for every null, in the
    reference's format (src/generate_null_matrices.py:82-86), with the library's multi-threaded writer."""
    import ctypes as C

    from . import _capi
    rows = np.ascontiguousarray(feature_rows, dtype=np.uint32)
    if rows.ndim == 2:
        rows = rows[None]
    n, F, wpr = rows.shape
    lib = _capi.load()
    rc = lib.sdx_write_nulls_tsv(rows.ctypes.data_as(C.POINTER(C.c_uint32)), n, F, len(samples), wpr, _names(samples), _names(genes),
                                 str(path_prefix).encode(), int(first_index), int(n_threads))
    if rc != 0:
        raise _capi.SdxError(rc, (lib.sdx_io_last_error() or b"").decode())


def read_nulls_tsv(path_prefix, first_index, n_nulls, samples, genes, n_threads=0):
    """Reads ``<path_prefix><first_index + i>.tsv`` into packed feature rows (n, F, ceil(S/32)) over the given sample and
    event tables; returns (rows, number of events that were not in ``genes``)."""
    import ctypes as C

    from . import _capi
    F, S = len(genes), len(samples)
    rows = np.zeros((int(n_nulls), F, bitpack.words_for(S)), dtype=np.uint32)
    unknown = C.c_int64(0)
    lib = _capi.load()
    rc = lib.sdx_read_nulls_tsv(str(path_prefix).encode(), int(first_index), int(n_nulls), F, S, rows.shape[2], _names(samples),
                                _names(genes), rows.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(unknown), int(n_threads))
    if rc != 0:
        raise _capi.SdxError(rc, (lib.sdx_io_last_error() or b"").decode())
    return rows, unknown.value
