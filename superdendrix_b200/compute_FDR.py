#!/usr/bin/env python
"""Drop-in for ``utils/compute_FDR.py`` without rpy2: appends a Benjamini-Hochberg ``FDR`` column to a concatenated
SuperDendrix results table (what the reference obtains from R's ``p.adjust(method='BH')``, utils/compute_FDR.py:35).
Same flags (-i, -p, -o), same handling of ``na`` p-values (counted as 1.0, :30-33), same output layout (:37-44)."""
from __future__ import annotations

import argparse
import sys

import numpy as np

from .screen import benjamini_hochberg


def get_parser():
    parser = argparse.ArgumentParser(fromfile_prefix_chars="@")
    parser.add_argument("-i", "--input_file", required=True, help="input concat SD results")
    parser.add_argument("-p", "--pval_index", required=False, default=4, type=int, help="pvalue index in input file")
    parser.add_argument("-o", "--output_file", required=True, help="name of output file")
    return parser


def run(args):
    header, rows, ps = None, [], []
    with open(args.input_file) as f:
        for r in f:
            row = r.strip("\n").split("\t")
            if header is None:
                header = row + ["FDR"]
                continue
            rows.append(row)
            ps.append(1.0 if row[args.pval_index] == "na" else float(row[args.pval_index]))
    adj = benjamini_hochberg(np.array(ps, dtype=np.float64)) if ps else np.zeros(0)
    with open(args.output_file, "w") as of:
        of.write("\t".join(header))
        for row, a in zip(rows, adj):
            of.write("\n" + "\t".join(row) + "\t" + str(float(a)))
    return adj


def main(argv=None):
    run(get_parser().parse_args(argv))
    return 0


if __name__ == "__main__":
    sys.exit(main())
