// Reference-side binding of the B200 library (SOURCE ONLY: R, Rcpp and a GPU-enabled R build are not
// available in the build image, so this file has never been compiled; the same C ABI is exercised
// from Python/ctypes by the test-suite).
//
// Drop this file into the reference's src/ in place of rcpp_threshold_reads.cpp, rcpp_get_xm_beta.cpp
// and rcpp_cx_report.cpp (and, for the siblings at the end of the file, rcpp_mhl_report.cpp, rcpp_match_target.cpp,
// rcpp_get_base_freqs.cpp, rcpp_fep.cpp), add `PKG_LIBS += -L$(EA_B200_LIB) -lepialleler_b200` and
// `PKG_CPPFLAGS += -I$(EA_B200_INCLUDE)` to src/Makevars, and re-run Rcpp::compileAttributes():
// the three exported R closures (R/RcppExports.R:12-14,28-30,64-66) and the registered native
// symbols (src/RcppExports.cpp:253-271) keep their names, arity and argument meaning.
#include <Rcpp.h>
#include <algorithm>
#include <cstdio>
#include <initializer_list>
#include <memory>
#include <thread>
#include <epialleler_b200.h>

namespace {

void ea_check(int rc) {
  if (rc != EA_OK) Rcpp::stop("epialleler_b200: %s", ea_last_error());
}

int poll_interrupt() {           // Rcpp::checkUserInterrupt() throws; the C ABI wants a flag
  try { Rcpp::checkUserInterrupt(); } catch (...) { return 1; }
  return 0;
}

struct Handle {
  ea_handle* h = nullptr;
  const void* seqxm = nullptr;   // identity of the strings the handle was packed from
  R_xlen_t nrow = 0;
  const int* rname0 = nullptr;   // identity of the columns (a subset / reordered bam.data has other vectors)
  const int* start0 = nullptr;
  ~Handle() { ea_free(h); }
};

// Packs bam.data.  need_bytes = false (thresholds, betas, cytosine reports, BED matching, lMHL): call planes only
// (ea_pack_calls_records / ea_pack_calls - the host transposes the reader's strings where they lie, half the bytes cross PCIe);
// need_bytes = true (lMHL, VCF base frequencies): byte rows too (ea_pack).  The pointer/length tables over the
// std::vector<std::string> are filled by a few threads (no R API inside them) - integration/harness/shim_harness.cpp is this
// very code without Rcpp, and what bench.py's end-to-end leg and tests/test_gpu_widen.py run.
// Where this standard library's std::string keeps its data pointer and its length (probe strings; libstdc++: 0 and 8).
bool string_layout(int64_t* ptr_off, int64_t* len_off) {
  int64_t po = -1, lo = -1;
  for (size_t n : {(size_t)100, (size_t)3, (size_t)7777}) {
    std::string probe(n, 'x');
    const unsigned char* raw = reinterpret_cast<const unsigned char*>(&probe);
    int64_t p = -1, l = -1;
    for (size_t o = 0; o + 8 <= sizeof(std::string); o += 8) {
      if (p < 0 && *reinterpret_cast<const char* const*>(raw + o) == probe.data()) p = (int64_t)o;
      else if (l < 0 && *reinterpret_cast<const size_t*>(raw + o) == n) l = (int64_t)o;
    }
    if (p < 0 || l < 0 || (po >= 0 && (p != po || l != lo))) return false;
    po = p; lo = l;
  }
  *ptr_off = po; *len_off = lo;
  return true;
}

Handle* pack(Rcpp::DataFrame& df, bool need_bytes) {
  Rcpp::XPtr<std::vector<std::string>> seqxm((SEXP)df.attr("seqxm_xptr"));
  Rcpp::IntegerVector rname = df["rname"], strand = df["strand"], start = df["start"], templid = df["templid"];
  const size_t n = seqxm->size();
  Handle* hd = new Handle;
  hd->seqxm = seqxm.get(); hd->nrow = df.nrows(); hd->rname0 = rname.begin(); hd->start0 = start.begin();
  int rc;
  int64_t po, lo;
  if (!need_bytes && sizeof(std::string) % 8 == 0 && string_layout(&po, &lo)) {
    // the vector's std::string objects are handed over as they are: no pointer tables (55 ms per 50 M strings)
    rc = ea_pack_calls_records(rname.begin(), strand.begin(), start.begin(), templid.begin(), df.nrows(), seqxm->data(),
                               static_cast<int64_t>(n), static_cast<int64_t>(sizeof(std::string)), po, lo, /*device=*/0, &hd->h);
  } else {
    std::unique_ptr<const uint8_t*[]> str(new const uint8_t*[n ? n : 1]);
    std::unique_ptr<int64_t[]> len(new int64_t[n ? n : 1]);
    const unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
      th.emplace_back([&, t]() {
        for (size_t i = n * t / nt; i < n * (t + 1) / nt; i++) {
          str[i] = reinterpret_cast<const uint8_t*>((*seqxm)[i].data());
          len[i] = static_cast<int64_t>((*seqxm)[i].size());
        }
      });
    for (auto& x : th) x.join();
    rc = (need_bytes ? ea_pack : ea_pack_calls)(rname.begin(), strand.begin(), start.begin(), templid.begin(), df.nrows(), str.get(),
                                                len.get(), static_cast<int64_t>(n), /*device=*/0, &hd->h);
  }
  if (rc != EA_OK) { delete hd; ea_check(rc); }
  ea_set_handle_poll(hd->h, poll_interrupt);   // per handle: other users of the library in the process keep theirs
  return hd;
}

// The device-resident packed copy hangs on the data.frame as a second external pointer, next to the
// reference's "seqxm_xptr" (src/rcpp_read_bam.cpp:185): pack once, reuse across reports.  It is reused only for the very
// same strings AND column vectors (a subset or reordered copy of bam.data has the same seqxm_xptr but other vectors).
ea_handle* handle_of(Rcpp::DataFrame& df, bool need_bytes = false) {
  Rcpp::XPtr<std::vector<std::string>> seqxm((SEXP)df.attr("seqxm_xptr"));
  Rcpp::IntegerVector rname = df["rname"], start = df["start"];
  if (df.hasAttribute("b200_xptr")) {
    Rcpp::XPtr<Handle> p((SEXP)df.attr("b200_xptr"));
    if (p->h && p->seqxm == seqxm.get() && p->nrow == df.nrows() && p->rname0 == rname.begin() && p->start0 == start.begin() &&
        (!need_bytes || ea_has_bytes(p->h)))
      return p->h;
  }
  Handle* hd = pack(df, need_bytes);
  df.attr("b200_xptr") = Rcpp::XPtr<Handle>(hd, true);
  return hd->h;
}

// a call that the call planes cannot serve answers EA_ERR_NEEDS_BYTES: pack again with the byte rows and repeat
template <class F>
void with_handle(Rcpp::DataFrame& df, F f) {
  int rc = f(handle_of(df, false));
  if (rc == EA_ERR_NEEDS_BYTES) rc = f(handle_of(df, true));
  ea_check(rc);
}

Rcpp::DataFrame wrap_report(ea_handle* h, int64_t n, Rcpp::DataFrame& df) {
  Rcpp::IntegerVector rname(n), strand(n), pos(n), context(n), meth(n), unmeth(n);
  ea_check(ea_cx_report_fetch(h, rname.begin(), strand.begin(), pos.begin(), context.begin(), meth.begin(), unmeth.begin()));
  Rcpp::IntegerVector in_rname = df["rname"], in_strand = df["strand"];
  rname.attr("class") = "factor";   rname.attr("levels") = in_rname.attr("levels");     // src/rcpp_cx_report.cpp:142-148
  strand.attr("class") = "factor";  strand.attr("levels") = in_strand.attr("levels");
  context.attr("class") = "factor";
  context.attr("levels") = Rcpp::CharacterVector::create("NA1", "CHH", "NA3", "NA4", "NA5", "CHG", "CG");  // :150-155
  return Rcpp::DataFrame::create(Rcpp::Named("rname") = rname, Rcpp::Named("strand") = strand, Rcpp::Named("pos") = pos,
                                 Rcpp::Named("context") = context, Rcpp::Named("meth") = meth, Rcpp::Named("unmeth") = unmeth);
}

}  // namespace

// [[Rcpp::export("rcpp_threshold_reads")]]
std::vector<bool> rcpp_threshold_reads(Rcpp::DataFrame& df, const std::string ctx_meth, const std::string ctx_unmeth,
                                       const std::string ooctx_meth, const std::string ooctx_unmeth, const unsigned int min_n_ctx,
                                       const double min_ctx_meth_frac, const double max_ooctx_meth_frac) {
  std::vector<uint8_t> flags(static_cast<size_t>(df.nrows()) + 1);
  with_handle(df, [&](ea_handle* h) {
    return ea_threshold_reads(h, ctx_meth.c_str(), ctx_unmeth.c_str(), ooctx_meth.c_str(), ooctx_unmeth.c_str(), min_n_ctx,
                              min_ctx_meth_frac, max_ooctx_meth_frac, flags.data());
  });
  return std::vector<bool>(flags.begin(), flags.end() - 1);
}

// [[Rcpp::export("rcpp_get_xm_beta")]]
std::vector<double> rcpp_get_xm_beta(Rcpp::DataFrame& df, const std::string ctx_meth, const std::string ctx_unmeth) {
  std::vector<double> beta(static_cast<size_t>(df.nrows()));
  with_handle(df, [&](ea_handle* h) { return ea_get_xm_beta(h, ctx_meth.c_str(), ctx_unmeth.c_str(), beta.data()); });
  return beta;
}

// [[Rcpp::export("rcpp_cx_report")]]
Rcpp::DataFrame rcpp_cx_report(Rcpp::DataFrame& df, Rcpp::LogicalVector& pass, const std::string ctx) {
  std::vector<uint8_t> flags(pass.size());
  for (R_xlen_t i = 0; i < pass.size(); i++) flags[i] = pass[i] != 0;  // NA_LOGICAL counts as TRUE (src/rcpp_cx_report.cpp:118)
  int64_t n = 0;
  with_handle(df, [&](ea_handle* h) { return ea_cx_report(h, flags.data(), ctx.c_str(), &n); });
  return wrap_report(handle_of(df), n, df);
}

// Optional fourth entry: the fused body of generateCytosineReport (R/generateCytosineReport.R:181-201),
// one pass over the XM bytes instead of two.  R side:
//   cx.report <- rcpp_cytosine_report_fused(bam, threshold.reads, ctx.meth, ctx.unmeth, ooctx.meth, ooctx.unmeth,
//                                            min.context.sites, min.context.beta, max.outofcontext.beta, report.ctx)
// [[Rcpp::export("rcpp_cytosine_report_fused")]]
Rcpp::DataFrame rcpp_cytosine_report_fused(Rcpp::DataFrame& df, const bool threshold_reads, const std::string ctx_meth,
                                           const std::string ctx_unmeth, const std::string ooctx_meth, const std::string ooctx_unmeth,
                                           const unsigned int min_n_ctx, const double min_ctx_meth_frac,
                                           const double max_ooctx_meth_frac, const std::string report_ctx) {
  ea_report_params p;
  p.threshold_reads = threshold_reads; p.ctx_meth = ctx_meth.c_str(); p.ctx_unmeth = ctx_unmeth.c_str();
  p.ooctx_meth = ooctx_meth.c_str(); p.ooctx_unmeth = ooctx_unmeth.c_str(); p.min_n_ctx = min_n_ctx;
  p.min_ctx_meth_frac = min_ctx_meth_frac; p.max_ooctx_meth_frac = max_ooctx_meth_frac; p.report_ctx = report_ctx.c_str();
  int64_t n = 0;
  with_handle(df, [&](ea_handle* h) { return ea_cytosine_report(h, &p, nullptr, &n); });
  return wrap_report(handle_of(df), n, df);
}

// Optional fifth entry: the same on several GPUs of the box from the one R process.  The rows are cut into contiguous genomic
// ranges with halo rows (ea_pack_multi_records plans them), every device packs and reports its range at once, the owned rows
// come back concatenated in genomic order - identical to the single-device report.  R side:
//   cx.report <- rcpp_cytosine_report_fused_multi(bam, devices = 0:7, threshold.reads, ..., report.ctx)
// [[Rcpp::export("rcpp_cytosine_report_fused_multi")]]
Rcpp::DataFrame rcpp_cytosine_report_fused_multi(Rcpp::DataFrame& df, Rcpp::IntegerVector devices, const bool threshold_reads,
                                                 const std::string ctx_meth, const std::string ctx_unmeth, const std::string ooctx_meth,
                                                 const std::string ooctx_unmeth, const unsigned int min_n_ctx, const double min_ctx_meth_frac,
                                                 const double max_ooctx_meth_frac, const std::string report_ctx) {
  Rcpp::XPtr<std::vector<std::string>> seqxm((SEXP)df.attr("seqxm_xptr"));
  Rcpp::IntegerVector rname = df["rname"], strand = df["strand"], start = df["start"], templid = df["templid"];
  int64_t po, lo;
  if (!string_layout(&po, &lo)) Rcpp::stop("epialleler_b200: std::string layout not understood");
  std::vector<int> dev(devices.begin(), devices.end());
  ea_multi* m = nullptr;
  ea_check(ea_pack_multi_records(rname.begin(), strand.begin(), start.begin(), templid.begin(), df.nrows(), seqxm->data(),
                                 static_cast<int64_t>(seqxm->size()), static_cast<int64_t>(sizeof(std::string)), po, lo, dev.data(),
                                 static_cast<int>(dev.size()), &m));
  ea_report_params p;
  p.threshold_reads = threshold_reads; p.ctx_meth = ctx_meth.c_str(); p.ctx_unmeth = ctx_unmeth.c_str();
  p.ooctx_meth = ooctx_meth.c_str(); p.ooctx_unmeth = ooctx_unmeth.c_str(); p.min_n_ctx = min_n_ctx;
  p.min_ctx_meth_frac = min_ctx_meth_frac; p.max_ooctx_meth_frac = max_ooctx_meth_frac; p.report_ctx = report_ctx.c_str();
  int64_t n = 0;
  int rc = ea_multi_cytosine_report(m, &p, nullptr, &n);
  Rcpp::IntegerVector o_rname(n), o_strand(n), pos(n), context(n), meth(n), unmeth(n);
  if (rc == EA_OK) rc = ea_multi_report_fetch(m, o_rname.begin(), o_strand.begin(), pos.begin(), context.begin(), meth.begin(), unmeth.begin());
  ea_multi_free(m);
  ea_check(rc);
  o_rname.attr("class") = "factor";   o_rname.attr("levels") = rname.attr("levels");
  o_strand.attr("class") = "factor";  o_strand.attr("levels") = strand.attr("levels");
  context.attr("class") = "factor";
  context.attr("levels") = Rcpp::CharacterVector::create("NA1", "CHH", "NA3", "NA4", "NA5", "CHG", "CG");
  return Rcpp::DataFrame::create(Rcpp::Named("rname") = o_rname, Rcpp::Named("strand") = o_strand, Rcpp::Named("pos") = pos,
                                 Rcpp::Named("context") = context, Rcpp::Named("meth") = meth, Rcpp::Named("unmeth") = unmeth);
}

// ---- the siblings that reuse the packed reads; same exported names and signatures as the reference ----

// R/RcppExports.R:40-42, src/rcpp_mhl_report.cpp:46
// [[Rcpp::export("rcpp_mhl_report")]]
Rcpp::DataFrame rcpp_mhl_report(Rcpp::DataFrame& df, const std::string ctx, int hmax, const int hmin, const double max_ooctx_meth_frac) {
  ea_handle* h = handle_of(df, /*need_bytes=*/true);
  int64_t n = 0;
  ea_check(ea_mhl_report(h, ctx.c_str(), hmax, hmin, max_ooctx_meth_frac, &n));
  Rcpp::IntegerVector rname(n), strand(n), pos(n), context(n), coverage(n);
  Rcpp::NumericVector length(n), lmhl(n);
  ea_check(ea_mhl_report_fetch(h, rname.begin(), strand.begin(), pos.begin(), context.begin(), coverage.begin(), length.begin(), lmhl.begin()));
  Rcpp::IntegerVector in_rname = df["rname"], in_strand = df["strand"];
  rname.attr("class") = "factor";   rname.attr("levels") = in_rname.attr("levels");
  strand.attr("class") = "factor";  strand.attr("levels") = in_strand.attr("levels");
  context.attr("class") = "factor";
  context.attr("levels") = Rcpp::CharacterVector::create("NA1", "CHH", "NA3", "NA4", "NA5", "CHG", "CG");
  return Rcpp::DataFrame::create(Rcpp::Named("rname") = rname, Rcpp::Named("strand") = strand, Rcpp::Named("pos") = pos,
                                 Rcpp::Named("context") = context, Rcpp::Named("coverage") = coverage,
                                 Rcpp::Named("length") = length, Rcpp::Named("lmhl") = lmhl);
}

namespace {
std::vector<int> match_target(Rcpp::DataFrame& df, Rcpp::DataFrame& bed, int mode, int param) {
  ea_handle* h = handle_of(df);
  Rcpp::IntegerVector chr = bed["seqnames"], start = bed["start"], end = bed["end"];   // factor codes on levels(bam$rname), NA_INTEGER = INT32_MIN
  std::vector<int> out(static_cast<size_t>(ea_nrow(h)));
  ea_check(ea_match_target(h, mode, chr.begin(), start.begin(), end.begin(), chr.size(), param, out.data()));
  return out;                                                                           // 1-based BED row or NA_INTEGER
}
}  // namespace

// R/RcppExports.R:32-38, src/rcpp_match_target.cpp:16,51
// [[Rcpp::export("rcpp_match_amplicon")]]
std::vector<int> rcpp_match_amplicon(Rcpp::DataFrame& df, Rcpp::DataFrame& bed, const int tolerance) { return match_target(df, bed, 0, tolerance); }
// [[Rcpp::export("rcpp_match_capture")]]
std::vector<int> rcpp_match_capture(Rcpp::DataFrame& df, Rcpp::DataFrame& bed, const signed int min_overlap) { return match_target(df, bed, 1, min_overlap); }

// R/RcppExports.R:24-26, src/rcpp_get_base_freqs.cpp:15
// [[Rcpp::export("rcpp_get_base_freqs")]]
Rcpp::NumericMatrix rcpp_get_base_freqs(Rcpp::DataFrame& df, std::vector<bool> pass, Rcpp::DataFrame& vcf) {
  ea_handle* h = handle_of(df, /*need_bytes=*/true);
  Rcpp::IntegerVector chr = vcf["seqnames"], pos = vcf["start"];
  std::vector<uint8_t> flags(pass.begin(), pass.end());
  Rcpp::NumericMatrix res(pos.size(), 20);                                              // column-major, like the library's output
  ea_check(ea_base_freqs(h, flags.data(), chr.begin(), pos.begin(), pos.size(), res.begin()));
  return res;
}

// R/RcppExports.R:20-22, src/rcpp_fep.cpp:9
// [[Rcpp::export]]
std::vector<double> rcpp_fep(Rcpp::DataFrame& df, std::vector<std::string> colnames) {
  Rcpp::IntegerVector A = df[colnames[0]], B = df[colnames[1]], C = df[colnames[2]], D = df[colnames[3]];
  std::vector<double> p(A.size(), NA_REAL);
  ea_check(ea_fisher_exact(A.begin(), B.begin(), C.begin(), D.begin(), A.size(), p.data()));
  for (auto& x : p) if (x != x) x = NA_REAL;                                            // the library reports NA as NaN
  return p;
}

// .writeReport for the cytosine report (R/internal.R:273-287): the file's bytes come formatted from the device.
// R side: rcpp_write_cx_report(bam, report.file) after rcpp_cx_report / rcpp_cytosine_report_fused on the same bam.
// [[Rcpp::export("rcpp_write_cx_report")]]
void rcpp_write_cx_report(Rcpp::DataFrame& df, const std::string path) {
  ea_handle* h = handle_of(df);
  Rcpp::IntegerVector in_rname = df["rname"];
  Rcpp::CharacterVector lv = in_rname.attr("levels");
  std::vector<std::string> names(lv.size());
  std::vector<const char*> ptr(lv.size());
  for (R_xlen_t i = 0; i < lv.size(); i++) { names[i] = Rcpp::as<std::string>(lv[i]); ptr[i] = names[i].c_str(); }
  int64_t nbytes = 0;
  ea_check(ea_cx_report_text(h, ptr.data(), static_cast<int32_t>(ptr.size()), /*header=*/1, &nbytes));
  char* buf = static_cast<char*>(ea_host_alloc(static_cast<size_t>(nbytes)));           // page-locked, cached by the library
  if (!buf) Rcpp::stop("epialleler_b200: %s", ea_last_error());
  int rc = ea_cx_report_text_fetch(h, buf);
  if (rc == EA_OK) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (f) { std::fwrite(buf, 1, static_cast<size_t>(nbytes), f); std::fclose(f); } else rc = -1000;
  }
  ea_host_free(buf);
  if (rc == -1000) Rcpp::stop("cannot open %s", path.c_str());
  ea_check(rc);
}
