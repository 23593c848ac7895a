"""oracle/refport.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure-Python port of the reference's CPU path, keeping its data structures
(adjacency lists, Python sets/dicts) and its use of CPython's ``random`` module,
so that (a) on the same interpreter and seed it reproduces the reference's
outputs exactly — pinned by ``tests/test_oracle_golden.py`` against fixtures
made from the reference itself — and (b) timing it on the GPU box's host cores
is a fair stand-in for timing the reference, which cannot travel there
(``bench.py`` ``cpu_baseline`` / ``--impl reference``, kind = "port").

Citations are relative to /root/reference/.
"""
from __future__ import annotations

import random
from collections import defaultdict

import numpy as np


# -- src/curveball.py:7-14 ---------------------------------------------------
def presences(matrix: np.ndarray) -> list:
    """Adjacency lists over the shorter dimension of a 0/1 matrix."""
    n_rows, n_cols = matrix.shape
    view = matrix if n_cols >= n_rows else matrix.T
    return [list(np.where(view[r] == 1)[0]) for r in range(min(n_rows, n_cols))]


# -- src/curveball.py:17-40 --------------------------------------------------
def curveball(matrix: np.ndarray, lists: list, n_trades: int = -1, recorder=None) -> np.ndarray:
    """5*min(R,C) pair trades on ``lists`` (mutated in place), then densify.

    ``recorder(a, b, to_a_or_None)`` is called once per trade when given; it is
    how the golden schedules are captured (None = skipped trade)."""
    n_rows, n_cols = matrix.shape
    short = min(n_rows, n_cols)
    index_range = range(len(lists))
    if n_trades == -1:
        n_trades = 5 * short
    for _ in range(n_trades):
        a, b = random.sample(index_range, 2)
        common = set(lists[a]) & set(lists[b])
        n_common, n_a, n_b = len(common), len(lists[a]), len(lists[b])
        if n_common == n_a or n_common == n_b:
            if recorder:
                recorder(a, b, None)
            continue
        pool = list(set(lists[a] + lists[b]) - common)
        keep = list(common)
        random.shuffle(pool)
        cut = n_a - n_common
        lists[a] = keep + pool[:cut]
        lists[b] = keep + pool[cut:]
        if recorder:
            recorder(a, b, pool[:cut])
    wide = n_cols >= n_rows
    out = np.zeros(matrix.shape if wide else matrix.T.shape, dtype="int8")
    for r in range(short):
        out[r, lists[r]] = 1
    return out if wide else out.T


# -- src/superdendrix.py:387-391 ---------------------------------------------
def super_w(cases_by_event, profile) -> float:
    covered = set(s for cases in cases_by_event for s in cases)
    pos = sum(profile[s] for s in covered if profile[s] > 0)
    neg = sum(abs(profile[s]) for cases in cases_by_event for s in cases)
    return 2 * pos - neg


# -- src/superdendrix.py:361-384 ---------------------------------------------
def conditional_permutation(module, event_to_cases, profile, samples, n_perm):
    cases_by_event = [event_to_cases[e] for e in module]
    real = super_w(cases_by_event, profile)
    worst_count, worst_event = 0, None
    for i, (event, cases) in enumerate(zip(module, cases_by_event)):
        hits = 0
        for _ in range(n_perm):
            cases_by_event[i] = random.sample(samples, len(cases))
            hits += super_w(cases_by_event, profile) >= real
        cases_by_event[i] = cases
        if hits > worst_count:
            worst_count, worst_event = hits, event
    return worst_event, worst_count


# -- src/i_o.py:18-88 ---------------------------------------------------------
def load_mutation_data(path, patients=None, gene_file=None, mut_only=False, min_freq=0, max_freq=float("inf")):
    """sample<TAB>event... TSV -> (m, n, genes, patients, geneToCases, patientToGenes)."""
    allowed = set()
    if gene_file:
        with open(gene_file) as fh:
            allowed = set(line.rstrip().split()[0] for line in fh if not line.startswith("#"))
    gene_cases, patient_genes = defaultdict(set), defaultdict(set)
    with open(path) as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            fields = line.rstrip().split("\t")
            who, events = fields[0], set(fields[1:])
            if patients and who not in patients:
                continue
            if allowed:
                events &= allowed
            patient_genes[who] = events
            for ev in events:
                gene_cases[ev].add(who)
    genes, who_all = list(gene_cases.keys()), list(patient_genes.keys())

    def drop(ev):
        for p in gene_cases[ev]:
            patient_genes[p].remove(ev)
        del gene_cases[ev]
        genes.remove(ev)

    if mut_only:
        for ev in [g for g in genes if not g.endswith("MUT")]:
            drop(ev)
    for ev in [g for g in genes if len(gene_cases[g]) < min_freq or len(gene_cases[g]) > max_freq]:
        drop(ev)
    return len(genes), len(who_all), genes, who_all, gene_cases, patient_genes


# -- src/generate_null_matrices.py:62-86 (in-memory part) ---------------------
def null_chain(matrix: np.ndarray, n_nulls: int, seed: int):
    """Yield n_nulls matrices of one Markov chain, as the reference driver does."""
    random.seed(seed)
    np.random.seed(seed)
    lists = presences(matrix)
    for _ in range(n_nulls):
        yield curveball(matrix, lists)


def null_rows(permuted: np.ndarray, samples, genes):
    """Rows of a null TSV: [sample, events carried...] (generate_null_matrices.py:75-80)."""
    return [[samples[r]] + [g for bit, g in zip(permuted[r], genes) if bit == 1] for r in range(len(permuted))]
