#!/usr/bin/env python
"""Drop-in for ``utils/parse_SD_results.py``: from a results table with an FDR column, the significant profiles
(FDR <= 0.20, :52), the (profile, feature) pairs with the feature's carrier count, and tallies of mutated genes / features
(:45-68); written as ``<out>pairs.tsv``, ``<out>profiles.tsv``, ``<out>mutated_genes.tsv``, ``<out>mutated_features.tsv``
in the reference's layout (:80-112)."""
from __future__ import annotations

import argparse
import sys

from .i_o import load_mutation_data


def get_parser():
    parser = argparse.ArgumentParser(description="Parse SuperDendrix results", fromfile_prefix_chars="@")
    parser.add_argument("-i", "--results", required=True, help="File name for SD results")
    parser.add_argument("-ef", "--mutation_matrix", required=False, help="File name for mutation data")
    parser.add_argument("-o", "--output_file", required=False, help="Prefix of the output files")
    parser.add_argument("-ct", "--CT", default=False, action="store_true")
    parser.add_argument("-find", "--fdr_index", default=13, type=int)
    parser.add_argument("--fdr", type=float, default=0.20, help="significance threshold [0.20]")
    return parser


def run(args):
    features_to_samples = {}
    if args.mutation_matrix:
        features_to_samples = load_mutation_data(args.mutation_matrix)[4]
    delim = "\t" if args.results.endswith("tsv") else ","
    pairs, profile_genes, mutated_genes, mutated_features = [], [], {}, {}
    suffix = "_Cancer" if args.CT else "_MUT"
    with open(args.results) as f:
        for line in f:
            if line.startswith("profile") or not line.strip():
                continue
            arr = line.rstrip("\n").split(delim)
            if len(arr) <= args.fdr_index or arr[args.fdr_index] in ("na", "nan", ""):
                continue
            profile, features, fdr = arr[0], arr[1].split(","), float(arr[args.fdr_index])
            if fdr > args.fdr:
                continue
            profile_genes.append([profile, str(fdr)])
            for feat in features:
                if feat.endswith(suffix):
                    gene = feat[:-len(suffix)]
                    mutated_genes[gene] = mutated_genes.get(gene, 0) + 1
                mutated_features[feat] = mutated_features.get(feat, 0) + 1
                pairs.append([profile, feat, str(len(features_to_samples.get(feat, ()))), str(fdr)])
    if args.output_file:
        with open(args.output_file + "pairs.tsv", "w") as of:
            of.write("profile\tfeature\tcoverage\tFDR\n")
            for row in pairs:
                of.write("\t".join(row) + "\n")
        with open(args.output_file + "profiles.tsv", "w") as of:
            of.write("profile\tfdr\n")
            for row in profile_genes:
                of.write("\t".join(row) + "\n")
        with open(args.output_file + "mutated_genes.tsv", "w") as of:
            of.write("mutated_gene\tcount\n")
            for k, v in mutated_genes.items():
                of.write(k + "\t" + str(v) + "\n")
        with open(args.output_file + "mutated_features.tsv", "w") as of:
            of.write("feature\tcount\ttype\tgene\n")
            for k, v in mutated_features.items():
                if k.endswith("_Cancer"):
                    of.write(k + "\t" + str(v) + "\t\t" + k[:-len("_Cancer")] + "\n")
    return {"pairs": pairs, "profiles": profile_genes, "mutated_genes": mutated_genes, "mutated_features": mutated_features}


def main(argv=None):
    run(get_parser().parse_args(argv))
    return 0


if __name__ == "__main__":
    sys.exit(main())
