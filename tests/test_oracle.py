"""CPU-only: pins the oracles.  The C restatement (oracle/cx_oracle.c) and, where it could be built,
the reference's own compiled sources (oracle/_ref) must reproduce (a) the constants the reference's
RUnit tests assert, (b) the committed golden vectors, (c) the BAM-free known-answer tests."""
import os

import numpy as np
import pytest

from helpers import (BETAS, CASES, KAT_C1_BETA, KAT_C1_NOTHR_ZXH, KAT_C1_PASS, KAT_C1_ROWS, KAT_C1_THR_ZXH, KAT_C2_ROWS,
                     KAT_C2_Z, KAT_C3_NOTHR_ZXH, KAT_C3_PASS, KAT_C3_ROWS, KAT_C3_THR_ZX, assert_report_equal, bam_from_xm,
                     golden_names, kat_c4_rows, load_golden, report_tuples)
from oracle.oracle import REF_SO, PortCpu, RefCpu

ORACLES = [PortCpu] + ([RefCpu] if os.path.exists(REF_SO) or os.path.isdir("/root/reference/src") else [])


@pytest.mark.parametrize("oracle", ORACLES)
@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden_vectors(oracle, name):
    bam, z = load_golden(name)
    o = oracle(bam)
    for cname, thr, tctx, ms, mb, mo, rctx in CASES:
        p, rep = o.generateCytosineReport(thr, tctx, ms, mb, mo, rctx)
        assert np.array_equal(p.astype(np.uint8), z[cname + ".pass"]), (name, cname)
        assert_report_equal(rep, {k: z[cname + "." + k] for k in rep}, "%s/%s" % (name, cname))
    for bname, m, u in BETAS:
        assert np.array_equal(o.rcpp_get_xm_beta(m, u), z[bname])


@pytest.mark.parametrize("oracle", ORACLES)
def test_reference_runit_constants_capture(oracle):
    # inst/unitTests/test_generateCytosineReport.R:7-60
    bam, _ = load_golden("capture")
    o = oracle(bam)
    _, cg = o.generateCytosineReport()
    _, cx = o.generateCytosineReport(threshold_reads=False, report_context="CX")
    assert cg["pos"].shape[0] == 15408 and cx["pos"].shape[0] == 97186
    assert (cg["meth"].sum(), cg["unmeth"].sum()) == (4974, 15245)
    assert (cx["meth"].sum(), cx["unmeth"].sum()) == (6051, 125903)
    assert np.bincount(cx["strand"]).tolist()[1:] == [48517, 48669]
    assert [int((cx["context"] == c).sum()) for c in (2, 6, 7)] == [58292, 23486, 15408]
    key = cx["rname"].astype(np.int64) << 32 | cx["pos"]
    for s in (1, 2):  # no duplicated (rname,pos) within a strand
        k = key[cx["strand"] == s]
        assert np.unique(k).shape[0] == k.shape[0]
    # :117-156 (min.mapq=30, min.baseq=20)
    bamq, _ = load_golden("capture_q30_bq20")
    oq = oracle(bamq)
    _, cgq = oq.generateCytosineReport()
    _, cxq = oq.generateCytosineReport(threshold_reads=False, report_context="CX")
    assert (cgq["pos"].shape[0], cgq["meth"].sum(), cgq["unmeth"].sum()) == (15197, 4830, 15062)
    assert (cxq["pos"].shape[0], cxq["meth"].sum(), cxq["unmeth"].sum()) == (96151, 5873, 124333)


@pytest.mark.parametrize("oracle", ORACLES)
def test_reference_runit_constants_single_end_and_amplicon(oracle):
    # inst/unitTests/test_generateCytosineReport.R:215-233
    bam, _ = load_golden("dragen_se")
    _, cx = oracle(bam).generateCytosineReport(threshold_reads=False, report_context="CX")
    assert cx["pos"].shape[0] == 3236
    assert [int((cx["context"] == c).sum()) for c in (2, 6, 7)] == [2165, 802, 269]
    assert (cx["meth"].sum(), cx["unmeth"].sum()) == (355, 3599)
    # config 1: amplicon010meth, defaults; pass count pinned by test_generateBedReport.R:22-35 (48 = 13+7+5+9+14)
    amp, _ = load_golden("amplicon010meth")
    p, cg = oracle(amp).generateCytosineReport()
    assert (amp.nrow, int(amp.blob()[1][-1]), int(p.sum())) == (500, 190455, 48)
    assert (cg["pos"].shape[0], cg["meth"].sum(), cg["unmeth"].sum()) == (478, 632, 6449)


@pytest.mark.parametrize("oracle", ORACLES)
def test_known_answer_tests(oracle):
    cg = ("Z", "z", "XH", "xh", 2, 0.5, 0.1)
    o = oracle(bam_from_xm(KAT_C1_ROWS))
    p = o.rcpp_threshold_reads(*cg)
    assert p.astype(int).tolist() == KAT_C1_PASS
    assert o.rcpp_get_xm_beta("Z", "z").tolist() == KAT_C1_BETA
    assert report_tuples(o.rcpp_cx_report(p, "ZXH")) == KAT_C1_THR_ZXH
    assert report_tuples(o.rcpp_cx_report(None, "ZXH")) == KAT_C1_NOTHR_ZXH

    o = oracle(bam_from_xm(KAT_C2_ROWS))
    p = o.rcpp_threshold_reads(*cg)
    assert p.all()
    rep = o.rcpp_cx_report(p, "Z")
    assert report_tuples(rep) == KAT_C2_Z and report_tuples(o.rcpp_cx_report(None, "Z")) == KAT_C2_Z
    assert (rep["pos"].shape[0], rep["meth"].sum(), rep["unmeth"].sum()) == (24, 30, 18)

    o = oracle(bam_from_xm(KAT_C3_ROWS))
    p = o.rcpp_threshold_reads(*cg)
    assert p.astype(int).tolist() == KAT_C3_PASS
    assert report_tuples(o.rcpp_cx_report(None, "ZXH")) == KAT_C3_NOTHR_ZXH
    assert report_tuples(o.rcpp_cx_report(p, "ZX")) == KAT_C3_THR_ZX

    o = oracle(bam_from_xm(kat_c4_rows()))
    p = o.rcpp_threshold_reads(*cg)
    assert p.sum() == 1 and p[0]
    rep = o.rcpp_cx_report(p, "Z")
    assert (rep["pos"].shape[0], rep["meth"].sum(), rep["unmeth"].sum()) == (10, 10, 9990)


def test_port_matches_reference_on_synthetic_reads():
    if RefCpu not in ORACLES:
        pytest.skip("reference sources not available here")
    from epialleler_b200 import synth_config, synth_host
    for cfg in (synth_config(7, 4000, [60000], 150), synth_config(8, 3000, [50000, 30000], 40, 300, n_islands=6, island_len=900)):
        bam = synth_host(cfg)
        a, b = PortCpu(bam), RefCpu(bam)
        for thr, tctx, rctx in ((True, "CG", "CG"), (False, "CG", "CX"), (True, "CHH", "CxG")):
            pa, ra = a.generateCytosineReport(thr, tctx, 2, 0.5, 0.1, rctx)
            pb, rb = b.generateCytosineReport(thr, tctx, 2, 0.5, 0.1, rctx)
            assert np.array_equal(pa, pb)
            assert_report_equal(ra, rb)
            assert ra["pos"].shape[0] > 0


# ---- BED-assisted report -------------------------------------------------
from helpers import BED_CASES  # noqa: E402


def golden_bed(z):
    return z["bed.chr"], z["bed.start"], z["bed.end"]


@pytest.mark.parametrize("oracle", ORACLES)
def test_reference_runit_constants_bed_report(oracle):
    # inst/unitTests/test_generateBedReport.R:12-45 (dims, nreads sums, per-amplicon VEF)
    bam, z = load_golden("amplicon010meth")
    o = oracle(bam)
    _, rep = o.generateBedReport(golden_bed(z), "amplicon")
    assert rep["VEF"].shape[0] == 5  # four amplicons + the NA row
    assert np.nansum(rep["nreads-"]) == 440 and np.nansum(rep["nreads+"]) + np.nansum(rep["nreads-"]) == 500
    assert np.allclose(rep["VEF"], [0.08333333333, 0.11475409836, 0.05376344086, 0.10714285714, 0.13207547170], rtol=0, atol=1e-10)
    _, nothr = o.generateBedReport(golden_bed(z), "amplicon", threshold_reads=False)
    assert nothr["VEF"].shape[0] == 5 and np.isnan(nothr["VEF"]).all()  # :47-54
    bam, z = load_golden("capture")
    _, rep = oracle(bam).generateBedReport(golden_bed(z), "capture")
    assert rep["VEF"].shape[0] == 565
    assert np.nansum(rep["nreads-"]) == 1472 and np.nansum(rep["nreads+"]) + np.nansum(rep["nreads-"]) == 2968


@pytest.mark.parametrize("oracle", ORACLES)
def test_reference_runit_constants_bed_ecdf(oracle):
    # inst/unitTests/test_generateBedEcdf.R:8-12: eCDF(0.5) of the in-context / out-of-context betas of amplicons 1 and 2
    bam, z = load_golden("amplicon010meth")
    o = oracle(bam)
    match = o.rcpp_match_amplicon(golden_bed(z), 1)
    ctx, oo = o.rcpp_get_xm_beta("Z", "z"), o.rcpp_get_xm_beta("XH", "xh")
    got = []
    for row in (1, 2):
        sel = match == row
        got += [float((ctx[sel] <= 0.5).mean()), float((oo[sel] <= 0.5).mean())]
    assert np.allclose(got, [0.916666666667, 1, 0.885245901639, 1], rtol=0, atol=1e-8)


@pytest.mark.parametrize("oracle", ORACLES)
@pytest.mark.parametrize("name", sorted(BED_CASES))
def test_oracle_reproduces_golden_bed_vectors(oracle, name):
    bam, z = load_golden(name)
    o = oracle(bam)
    for cname, typ, par in BED_CASES[name]:
        match, rep = o.generateBedReport(golden_bed(z), typ, match_tolerance=par, match_min_overlap=par)
        assert np.array_equal(match, z["bed." + cname + ".match"]), (name, cname)
        assert np.array_equal(rep["counts"], z["bed." + cname + ".counts"]), (name, cname)
        assert np.array_equal(rep["VEF"], z["bed." + cname + ".vef"], equal_nan=True), (name, cname)


def random_bed(rng, bam, nbed, overlapping=True):
    """Unsorted random targets around the reads' own coordinates, with overlaps, duplicates and chromosomes the BAM lacks."""
    i = rng.integers(0, bam.nrow, nbed)
    _, off = bam.blob()
    ln = np.diff(off)[bam.templid[i]]
    st = bam.start[i] + rng.integers(-30, 30, nbed)
    en = st + np.maximum(1, (ln * rng.uniform(0.2, 1.5, nbed)).astype(np.int64) + (rng.integers(0, 400, nbed) if overlapping else 0))
    chr_ = bam.rname[i].astype(np.int32).copy()
    chr_[rng.random(nbed) < 0.1] = -2147483648
    return chr_, st.astype(np.int32), en.astype(np.int32)


def test_port_matches_reference_on_random_beds():
    if RefCpu not in ORACLES:
        pytest.skip("compiled reference not available")
    rng = np.random.default_rng(7)
    bam, _ = load_golden("capture")
    ref, port = RefCpu(bam), PortCpu(bam)
    for nbed in (0, 1, 37, 400):
        bed = random_bed(rng, bam, nbed)
        for par in (-5, 0, 1, 40):
            assert np.array_equal(ref.rcpp_match_amplicon(bed, par), port.rcpp_match_amplicon(bed, par))
            assert np.array_equal(ref.rcpp_match_capture(bed, par), port.rcpp_match_capture(bed, par))


# ---- lMHL report -------------------------------------------------------------
from helpers import MHL_CASES, assert_mhl_equal  # noqa: E402


@pytest.mark.parametrize("oracle", ORACLES)
def test_reference_runit_constants_mhl(oracle):
    # inst/unitTests/test_generateMhlReport.R:7-30,48-83
    bam, _ = load_golden("capture")
    o = oracle(bam)
    m1 = o.generateMhlReport(max_haplotype_window=1)
    _, cg = o.generateCytosineReport(threshold_reads=False)
    assert np.array_equal(m1["lmhl"], cg["meth"] / (cg["meth"] + cg["unmeth"]))  # :7-12, identical()
    m = o.generateMhlReport()
    sp, sm = m["strand"] == 1, m["strand"] == 2
    assert (int(m["coverage"].sum()), int(m["coverage"][sp].sum()), int(m["coverage"][sm].sum())) == (20219, 10188, 10031)
    assert np.allclose([m["length"].sum(), m["lmhl"].sum()], [229119.960, 2666.456], rtol=1e-6, atol=0)  # RUnit::checkEquals tolerance
    assert np.allclose([m["length"][sp].sum(), m["lmhl"][sp].sum(), m["length"][sm].sum(), m["lmhl"][sm].sum()],
                       [119605.010, 1281.342, 109514.950, 1385.114], rtol=1e-6, atol=0)
    bam, _ = load_golden("amplicon010meth")
    o = oracle(bam)
    for kw, want in ((dict(max_outofcontext_beta=1), (7081, 342, 6739, 6060.46765, 45.78637, 2380.83333, 34.03206, 3679.63432, 11.75431)),
                     ({}, (7070, 339, 6731, 6051.54262, 43.53694, 2375.83333, 33.78206, 3675.709286, 9.754883))):
        m = o.generateMhlReport(**kw)
        sp, sm = m["strand"] == 1, m["strand"] == 2
        assert (int(m["coverage"].sum()), int(m["coverage"][sp].sum()), int(m["coverage"][sm].sum())) == want[:3]
        assert np.allclose([m["length"].sum(), m["lmhl"].sum(), m["length"][sp].sum(), m["lmhl"][sp].sum(), m["length"][sm].sum(), m["lmhl"][sm].sum()],
                           want[3:], rtol=1e-6, atol=0)


@pytest.mark.parametrize("oracle", ORACLES)
@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden_mhl_vectors(oracle, name):
    bam, z = load_golden(name)
    o = oracle(bam)
    for cname, hctx, hmax, hmin, moo in MHL_CASES:
        rep = o.generateMhlReport(hctx, hmax, hmin, moo)
        assert_mhl_equal(rep, {k: z[cname + "." + k] for k in rep}, "%s/%s" % (name, cname))


# ---- VCF-assisted report -------------------------------------------------------
from helpers import VCF_CASES, VCF_COLS, golden_vcf, random_vcf  # noqa: E402


@pytest.mark.parametrize("oracle", ORACLES)
def test_reference_runit_constants_vcf_report(oracle):
    # inst/unitTests/test_generateVcfReport.R:33-125 (amplicon010meth + amplicon.bed + amplicon.vcf.gz, vcf.style NCBI)
    bam, z = load_golden("amplicon010meth")
    vcf = golden_vcf(z)
    assert len(vcf["REF"]) == 56                                                           # dim(amplicon.report)[1]
    _, rep = oracle(bam).generateVcfReport(vcf)
    assert np.nansum(rep["FEp+"]) == 40                                                    # :56-59
    assert abs(np.nansum(rep["FEp-"]) - 40.15024191) < 1e-8                                # :60-64
    assert np.nansum(rep["SumRef"]) == 5282 and np.nansum(rep["SumAlt"]) == 14            # :66-74
    groups = sorted(set(zip(vcf["REF"], vcf["ALT"])))
    by = lambda col: [int(np.nansum([rep[col][i] for i in range(56) if (vcf["REF"][i], vcf["ALT"][i]) == g])) for g in groups]
    assert [sum(1 for i in range(56) if (vcf["REF"][i], vcf["ALT"][i]) == g) for g in groups] == [3, 4, 1, 7, 2, 13, 11, 5, 4, 2, 3, 1]
    assert by("M+Ref") == [0] * 12 and by("U+Ref") == [0] * 12 and by("M+Alt") == [0] * 12 and by("U+Alt") == [0] * 12
    assert by("M-Ref") == [25, 0, 13, 72, 20, 149, 0, 54, 53, 26, 39, 12]
    assert by("U-Ref") == [201, 0, 142, 722, 196, 1617, 0, 555, 534, 285, 427, 140]
    assert by("M-Alt") == [0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0]
    assert by("U-Alt") == [0, 0, 0, 3, 1, 4, 0, 1, 1, 0, 1, 2]
    assert by("SumRef") == [226, 0, 155, 794, 216, 1766, 0, 609, 587, 311, 466, 152]
    assert by("SumAlt") == [0, 0, 0, 4, 1, 4, 0, 1, 1, 0, 1, 2]
    bam, z = load_golden("capture")
    assert z["vcf.chr"].shape[0] == 26292                                                  # dim(capture.report)[1], :42-45


@pytest.mark.parametrize("oracle", ORACLES)
@pytest.mark.parametrize("name", sorted(VCF_CASES))
def test_oracle_reproduces_golden_vcf_vectors(oracle, name):
    bam, z = load_golden(name)
    o = oracle(bam)
    for cname, thr in VCF_CASES[name]:
        freqs, rep = o.generateVcfReport(golden_vcf(z), threshold_reads=thr)
        assert np.array_equal(freqs, z["vcf." + cname + ".freqs"]), (name, cname)
        for k in VCF_COLS:
            assert np.array_equal(rep[k], z["vcf." + cname + "." + k], equal_nan=True), (name, cname, k)
        for k in ("FEp+", "FEp-"):  # the port sums the tables by definition, HTSlib's walk steps from table to table
            assert np.allclose(rep[k], z["vcf." + cname + "." + k], rtol=0, atol=1e-9, equal_nan=True), (name, cname, k)


def test_port_matches_reference_on_random_vcfs():
    if RefCpu not in ORACLES:
        pytest.skip("compiled reference not available")
    rng = np.random.default_rng(11)
    for name in ("capture", "amplicon010meth"):
        bam, _ = load_golden(name)
        ref, port = RefCpu(bam), PortCpu(bam)
        for n in (0, 1, 50, 3000):
            vcf = random_vcf(rng, bam, n)
            pass_ = rng.random(bam.nrow) < 0.4
            a = ref.rcpp_get_base_freqs(pass_, (vcf["seqnames"], vcf["start"]))
            b = port.rcpp_get_base_freqs(pass_, (vcf["seqnames"], vcf["start"]))
            assert a.shape == (n, 20) and np.array_equal(a, b), (name, n)


def test_fisher_exact_host_function_matches_the_oracles():
    """ea_fisher_exact is host arithmetic inside the product library (no GPU needed): same walk as the reference build's
    kt_fisher_exact stand-in -> bit-identical; the port sums every table by definition -> 1e-9."""
    import epialleler_b200 as ea
    rng = np.random.default_rng(3)
    n = 5000
    cols = [rng.integers(0, m, n).astype(np.int32) for m in (300, 3000, 20, 200)]
    cols[1][::97] = -2147483648
    cols.append(None)
    bam = bam_from_xm([(1, 1, 1, "Z")])
    got = ea.rcpp_fep(*cols[:4])
    want = PortCpu(bam).rcpp_fep(*cols[:4])
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got).sum() == len(range(0, n, 97))
    assert np.allclose(got, want, rtol=0, atol=1e-9, equal_nan=True)
    if RefCpu in ORACLES:
        assert np.array_equal(got, RefCpu(bam).rcpp_fep(*cols[:4]), equal_nan=True)
    # textbook values: Fisher's tea-tasting table and a balanced table
    p = ea.rcpp_fep([3, 5, 0], [1, 5, 0], [1, 5, 0], [3, 5, 0])
    assert abs(p[0] - 34.0 / 70.0) < 1e-12 and p[1] == 1.0 and p[2] == 1.0


def test_oracle_side_generator_equals_the_library_generator():
    """oracle/_build/libsynth.so (the generator header built for the host with g++; what bench.py's reference arm uses so
    that it never loads the product library) produces the rows ea_synth_host produces, bit for bit, sub-ranges included."""
    import epialleler_b200 as ea
    from oracle.oracle import synth_rows
    for kw, lo, hi in ((dict(seed=20261018, n_reads=5000, chr_len=[300000], len_min=150), 0, None),
                       (dict(seed=20261019, n_reads=8000, chr_len=[90000, 50000, 20000], len_min=100, len_max=417, n_islands=7, island_len=900), 1234, 6000),
                       (dict(seed=20261020, n_reads=60, chr_len=[500000, 200000], len_min=10000, len_max=30000), 10, 50)):
        a = ea.synth_host(ea.synth_config(**kw), lo, hi)
        b = synth_rows(row_begin=lo, row_end=hi, **kw)
        for k in ("rname", "strand", "start"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), k
        assert np.array_equal(a.blob()[0], b.blob()[0]) and np.array_equal(a.blob()[1], b.blob()[1])
