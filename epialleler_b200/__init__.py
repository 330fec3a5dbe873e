"""epialleler_b200 -- B200-native cytosine-report hot path of BBCG/epialleleR.

Host-side mirror of the reference's interface for this path (same names and argument meaning):
``preprocessBam`` (pass-through of already preprocessed reads), ``generateCytosineReport`` and the
three Rcpp routines ``rcpp_threshold_reads``, ``rcpp_get_xm_beta``, ``rcpp_cx_report``; all compute
runs in hand-written sm_100a kernels behind the C ABI of include/epialleler_b200.h.
"""
from .api import (BamData, CONTEXT_TO_BASES, generateCytosineReport, preprocessBam, rcpp_cx_report,  # noqa: F401
                  rcpp_get_xm_beta, rcpp_threshold_reads, synth_config, synth_device, synth_host)
from .api import (generateBedReport, generateAmpliconReport, generateCaptureReport, generateBedEcdf, ecdf_at,  # noqa: F401
                  rcpp_match_amplicon, rcpp_match_capture, rcpp_mhl_report, generateMhlReport)
from .api import generateVcfReport, rcpp_get_base_freqs, rcpp_fep, read_vcf, report_text  # noqa: F401
from .api import Ingest, MultiBam  # noqa: F401
from ._native import EaError  # noqa: F401
