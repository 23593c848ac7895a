"""Host-side logic: loaders (against the fixtures made from the reference), the TSV writer,
profile parsing, bit packing and the sharding / merging helpers.  CPU only."""
import json
import os

import numpy as np
import pytest

from superdendrix_b200 import bitpack, dist, i_o, synth
from superdendrix_b200 import superdendrix as sd

from conftest import GOLDEN


def test_loader_matches_reference_fixtures():
    with open(os.path.join(GOLDEN, "loader.json")) as fh:
        cases = json.load(fh)
    d = os.path.join(GOLDEN, "loader")
    for name, c in cases.items():
        kw = dict(c["kwargs"])
        m, n, genes, patients, g2c, p2g = i_o.load_mutation_data(
            os.path.join(d, "features.tsv"), kw.get("patients"),
            os.path.join(d, kw["geneFile"]) if "geneFile" in kw else None,
            kw.get("mut_only", False), kw.get("min_freq", 0), kw.get("max_freq", float("inf")))
        assert (m, n) == (c["m"], c["n"]), name
        assert sorted(genes) == c["genes"] and list(patients) == c["patients"]
        assert {g: sorted(v) for g, v in g2c.items()} == c["geneToCases"]
        assert {p: sorted(v) for p, v in p2g.items()} == c["patientToGenes"]


def test_reference_null_files_round_trip(tmp_path):
    """Files written by the reference generator parse, pack, unpack and re-write to the same content."""
    gdir = os.path.join(GOLDEN, "generator")
    m, n, genes, samples, g2c, p2g = i_o.load_mutation_data(os.path.join(gdir, "features.tsv"))
    dense = i_o.dense_matrix(genes, samples, p2g)
    rows = i_o.pack_features(genes, samples, g2c)
    assert np.array_equal(bitpack.unpack_rows(rows, len(samples)), dense.T)
    for f in sorted(os.listdir(gdir)):
        if not f.startswith("ref_null_"):
            continue
        pm, pn, pgenes, psamples, pg2c, pp2g = i_o.load_mutation_data(os.path.join(gdir, f))
        nd = i_o.dense_matrix(genes, samples, pp2g)
        assert np.array_equal(nd.sum(0), dense.sum(0)) and np.array_equal(nd.sum(1), dense.sum(1))   # Curveball margins
        out = tmp_path / f
        i_o.write_sample_events(str(out), samples, genes, nd)
        again = i_o.load_mutation_data(str(out))
        assert {g: sorted(v) for g, v in again[4].items()} == {g: sorted(v) for g, v in pg2c.items()}
        assert again[3] == psamples                      # samples with no events are kept, in order
        assert open(out).readline() == "#Samples\tEvents\n"


def test_parse_profile_semantics(tmp_path):
    p = tmp_path / "t.tsv"
    p.write_text("id\tA\tB\ns1\t1.5\t2\ns2\tnan\t3\ns3\tinf\t4\ns4\t\t5\ns5\t-2\tx\n# c\ts\ns6\t0\t6\n")
    w, bad, rows = sd.parse_profile(str(p), "A")                      # increased_dependency: sign flip
    assert w == {"s1": -1.5, "s3": -100.0, "s5": 2.0, "s6": -0.0} and bad == 3 and rows == 6
    w, bad, _ = sd.parse_profile(str(p), "A", direction="positive", offset=0.5)
    assert w == {"s1": 2.0, "s3": 100.0, "s5": -1.5, "s6": 0.5}
    w, _, _ = sd.parse_profile(str(p), "A", direction="positive", unit_weights=True)
    assert w == {"s1": 1, "s3": 100.0, "s5": -1, "s6": -1}
    w, bad, _ = sd.parse_profile(str(p), "B", direction="positive")
    assert bad == 1 and "s5" not in w and w["s6"] == 6.0


def test_profile_loaders(tmp_path):
    p = tmp_path / "prof.csv"
    p.write_text("id,G1_x,G2_y\na,1,NA\nb,,2.5\n")
    profs, samples = i_o.load_profiles(str(p))
    assert samples == ["a", "b"] and np.isnan(profs["G1_x"][1]) and np.isnan(profs["G2_y"][0]) and profs["G2_y"][1] == 2.5
    profs, samples = i_o.load_profiles(str(p), sample_whitelist={"b"})
    assert samples == ["b"] and profs["G2_y"].tolist() == [2.5]
    assert i_o.load_single_profile(str(p), "G1") == {"a": 1.0, "b": 0.0}
    ev, ms = i_o.load_events(os.path.join(GOLDEN, "loader", "features.tsv"))
    assert ms and all(isinstance(v, set) for v in ev.values())


def test_bitpack_round_trip_and_padding():
    rng = np.random.default_rng(0)
    for L in (1, 31, 32, 33, 770):
        d = rng.random((5, L)) < 0.3
        p = bitpack.pack_rows(d)
        assert p.shape == (5, bitpack.words_for(L)) and p.dtype == np.uint32
        assert np.array_equal(bitpack.unpack_rows(p, L), d.astype(np.uint8))
        assert np.array_equal(bitpack.popcount_rows(p), d.sum(1))
        if L % 32:
            assert (p[:, -1] >> (L % 32) == 0).all()
    assert bitpack.pack_mask([0, 33], 40).tolist() == [1, 2]


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 3, 8):
            r = [dist.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_shard_chains_keeps_chains_whole():
    for n, L, world in ((100, 32, 3), (100000, 32, 8), (5, 32, 4), (64, 1, 2), (0, 16, 2)):
        r = [dist.shard_chains(n, L, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(a % L == 0 or a == n for a, _ in r)


def test_scan_feature_ranges_balance_work():
    F = 2000
    for world in (1, 2, 4, 8):
        r = dist.scan_feature_ranges(F, 3, world)
        assert r[0][0] == 0 and r[-1][1] == F and all(r[i][1] == r[i + 1][0] for i in range(world - 1))

        def work(lo, hi):
            g = np.arange(lo, hi, dtype=np.float64)
            q = F - 1 - g
            return (1 + q + q * (q - 1) / 2).sum()
        w = [work(*x) for x in r]
        assert sum(w) == F + F * (F - 1) // 2 + F * (F - 1) * (F - 2) // 6
        assert max(w) / (sum(w) / world) < 1.02
    assert dist.scan_feature_ranges(3, 3, 8)[-1][1] == 3


def test_merge_hits_orders_like_a_single_scan():
    a = {"W": np.array([5.0, 3.0, -np.inf]), "features": np.array([[1, 2, 3], [4, -1, -1], [-1, -1, -1]]), "coverage": np.array([9, 2, 0])}
    b = {"W": np.array([5.0, 4.0, 3.0]), "features": np.array([[0, 9, -1], [7, 8, 9], [2, -1, -1]]), "coverage": np.array([7, 5, 1])}
    m = dist.merge_hits([a, b], 4)
    assert m["W"].tolist() == [5.0, 5.0, 4.0, 3.0]
    assert m["features"].tolist() == [[0, 9, -1], [1, 2, 3], [7, 8, 9], [2, -1, -1]]      # equal W: pairs before triples
    assert m["coverage"].tolist() == [7, 9, 5, 1]
    assert dist.merge_hits([a], 5)["features"][2:].tolist() == [[-1, -1, -1]] * 3


def test_synthetic_shapes_are_stable():
    dense, samples, feats = synth.feature_matrix(770, 400)
    assert dense.shape[0] == 770 and 380 <= dense.shape[1] <= 400 and dense.sum(0).min() >= 1
    assert dense.sum() == synth.feature_matrix(770, 400)[0].sum()
    w = synth.weights(770, 2)
    assert w.shape == (2, 770) and (w == 0).mean() > 0.005 and np.abs(w).max() <= 20


def test_native_tsv_writer_and_reader_match_the_python_path(tmp_path):
    """The C++ writer emits byte-identical files to the Python writer (reference format) and the C++ reader
    reproduces the packed rows, applies the sample whitelist and skips unknown events like the gene whitelist."""
    rng = np.random.default_rng(3)
    S, F, n = 70, 45, 9
    samples = ["ACH-%06d" % i for i in range(S)]
    genes = ["G%d_%s_MUT" % (j, "AIO"[j % 3]) for j in range(F)]
    dense = (rng.random((n, S, F)) < 0.08).astype(np.uint8)
    dense[0, 5] = 0                                              # a sample without events: bare name line
    rows = np.stack([bitpack.pack_rows(d.T) for d in dense])
    out = tmp_path / "nulls"
    out.mkdir()
    i_o.write_nulls_tsv(str(out) + "/", 100, rows, samples, genes, n_threads=3)
    for i in range(n):
        ref = tmp_path / ("py%d.tsv" % i)
        i_o.write_sample_events(str(ref), samples, genes, dense[i])
        assert open(out / ("%d.tsv" % (100 + i))).read() == open(ref).read()
    back, unknown = i_o.read_nulls_tsv(str(out) + "/", 100, n, samples, genes)
    assert unknown == 0 and np.array_equal(back, rows)
    # whitelist semantics: fewer samples (in another order) and fewer events
    sub_s = [samples[i] for i in (9, 3, 50, 5)]
    sub_g = genes[::2]
    back, unknown = i_o.read_nulls_tsv(str(out) + "/", 100, n, sub_s, sub_g)
    for i in range(n):
        want = dense[i][[9, 3, 50, 5]][:, ::2]
        assert np.array_equal(bitpack.unpack_rows(back[i], len(sub_s)), want.T)
    assert unknown == int(sum(d[[9, 3, 50, 5]][:, 1::2].sum() for d in dense))
    # files written by the reference generator parse to the same matrices as the Python loader
    gdir = os.path.join(GOLDEN, "generator")
    _, _, g, s, _, p2g = i_o.load_mutation_data(os.path.join(gdir, "features.tsv"))
    rows, unknown = i_o.read_nulls_tsv(os.path.join(gdir, "ref_null_"), 5, 3, s, g)
    for i in range(3):
        pp2g = i_o.load_mutation_data(os.path.join(gdir, "ref_null_%d.tsv" % (5 + i)))[5]
        assert np.array_equal(bitpack.unpack_rows(rows[i], len(s)).T, i_o.dense_matrix(g, s, pp2g))
    assert unknown == 0
    with pytest.raises(Exception):
        i_o.read_nulls_tsv(str(out) + "/", 0, 1, samples, genes)        # missing file: error, not silence


def test_benjamini_hochberg_matches_definition():
    from superdendrix_b200.screen import benjamini_hochberg
    p = np.array([0.01, 0.04, 0.03, np.nan, 0.005, 0.5])
    adj = benjamini_hochberg(p)
    # R: p.adjust(c(0.01, 0.04, 0.03, 0.005, 0.5), "BH") = 0.025 0.05 0.05 0.025 0.5
    assert np.allclose(adj[[0, 1, 2, 4, 5]], [0.025, 0.05, 0.05, 0.025, 0.5]) and np.isnan(adj[3])
    assert np.isnan(benjamini_hochberg([np.nan])).all() and benjamini_hochberg([0.2]).tolist() == [0.2]


# ---------------------------------------------------------------- property tests (hypothesis)
from hypothesis import given, settings, strategies as st


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 40), st.integers(1, 200), st.integers(0, 2 ** 31 - 1))
def test_property_bitpack_round_trip(rows, bits, seed):
    d = np.random.default_rng(seed).random((rows, bits)) < 0.3
    p = bitpack.pack_rows(d)
    assert np.array_equal(bitpack.unpack_rows(p, bits), d.astype(np.uint8))
    assert np.array_equal(bitpack.popcount_rows(p), d.sum(1))
    if bits % 32:
        assert ((p[:, -1] >> (bits % 32)) == 0).all()          # padding bits stay zero (contract of include/sdx.h)


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 25), st.integers(1, 40), st.floats(0.0, 0.6), st.integers(0, 2 ** 31 - 1))
def test_property_tsv_round_trip_native_equals_python(S, F, density, seed):
    import tempfile
    rng = np.random.default_rng(seed)
    dense = (rng.random((S, F)) < density).astype(np.uint8)
    samples = ["s%d" % i for i in range(S)]
    genes = ["g%d_MUT" % j for j in range(F)]
    with tempfile.TemporaryDirectory() as d:
        rows = bitpack.pack_rows(dense.T)
        i_o.write_nulls_tsv(d + "/", 3, rows, samples, genes, n_threads=1)
        i_o.write_sample_events(d + "/py.tsv", samples, genes, dense)
        assert open(d + "/3.tsv").read() == open(d + "/py.tsv").read()
        back, unknown = i_o.read_nulls_tsv(d + "/", 3, 1, samples, genes, n_threads=1)
        assert unknown == 0 and np.array_equal(back[0], rows)
        m, n, g2, s2, g2c, p2g = i_o.load_mutation_data(d + "/3.tsv")
        assert s2 == samples                                        # every sample is listed, even without events
        assert np.array_equal(i_o.dense_matrix(genes, samples, p2g), dense)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 10 ** 6), st.integers(1, 16))
def test_property_shard_ranges_partition(n, world):
    r = [dist.shard_range(n, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_fdr_and_result_parsing_cli_like_the_reference_utils(tmp_path):
    """utils/compute_FDR.py (BH column; 'na' counts as 1.0) and utils/parse_SD_results.py (profiles with FDR <= 0.20, pairs, tallies)"""
    import scipy.stats

    from superdendrix_b200 import compute_FDR, parse_SD_results, synth
    res = tmp_path / "all.tsv"
    rows = [("P%d" % i, "G%d_A_MUT,G%d_I_MUT" % (i % 3 + 1, i % 2 + 1), "3,4", "1.0,0.5", p) for i, p in enumerate(["0.001", "0.2", "na", "0.01", "0.04", "0.9"])]
    with open(res, "w") as fh:
        fh.write("profile\tfeatures\t#samples\tscores\tP_value\n")
        fh.write("\n".join("\t".join(r) for r in rows))
    out = tmp_path / "fdr.tsv"
    adj = compute_FDR.run(compute_FDR.get_parser().parse_args(["-i", str(res), "-p", "4", "-o", str(out)]))
    want = scipy.stats.false_discovery_control([0.001, 0.2, 1.0, 0.01, 0.04, 0.9], method="bh")
    assert np.allclose(adj, want)
    lines = open(out).read().split("\n")
    assert lines[0].split("\t")[-1] == "FDR" and len(lines) == 7 and float(lines[1].split("\t")[-1]) == adj[0]
    dense, samples, features = synth.feature_matrix(30, 12, seed=1)
    mut = tmp_path / "f.tsv"
    synth.write_features_tsv(str(mut), dense, samples, features)
    r = parse_SD_results.run(parse_SD_results.get_parser().parse_args(["-i", str(out), "-ef", str(mut), "-o", str(tmp_path) + "/x_", "-find", "5"]))
    sig = [i for i in range(6) if adj[i] <= 0.20]
    assert [p[0] for p in r["profiles"]] == ["P%d" % i for i in sig]
    assert len(r["pairs"]) == 2 * len(sig) and sum(r["mutated_genes"].values()) == 2 * len(sig)
    assert open(str(tmp_path) + "/x_profiles.tsv").read().splitlines()[0] == "profile\tfdr"
