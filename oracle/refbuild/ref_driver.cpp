// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-callable driver around the reference's own, UNMODIFIED hot-path functions
//   rcpp_threshold_reads  (/root/reference/src/rcpp_threshold_reads.cpp:15)
//   rcpp_get_xm_beta      (/root/reference/src/rcpp_get_xm_beta.cpp:10)
//   rcpp_cx_report        (/root/reference/src/rcpp_cx_report.cpp:34)
// which oracle/Makefile compiles from where they lie into oracle/_ref/.
// Used (a) to validate oracle/cx_oracle.c and the CUDA path, and (b) as the
// CPU baseline ("kind":"reference") in bench.py.  Timings returned here cover
// only the reference function call, never the marshalling around it.
#include <Rcpp.h>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

std::vector<bool> rcpp_threshold_reads(Rcpp::DataFrame& df, const std::string ctx_meth,
                                       const std::string ctx_unmeth, const std::string ooctx_meth,
                                       const std::string ooctx_unmeth, const unsigned int min_n_ctx,
                                       const double min_ctx_meth_frac, const double max_ooctx_meth_frac);
std::vector<double> rcpp_get_xm_beta(Rcpp::DataFrame& df, const std::string ctx_meth,
                                     const std::string ctx_unmeth);
Rcpp::DataFrame rcpp_cx_report(Rcpp::DataFrame& df, Rcpp::LogicalVector& pass, const std::string ctx);
// /root/reference/src/rcpp_match_target.cpp:16,51
std::vector<int> rcpp_match_amplicon(Rcpp::DataFrame& df, Rcpp::DataFrame& bed, const int tolerance);
std::vector<int> rcpp_match_capture(Rcpp::DataFrame& df, Rcpp::DataFrame& bed, const signed int min_overlap);
// /root/reference/src/rcpp_mhl_report.cpp:46
Rcpp::DataFrame rcpp_mhl_report(Rcpp::DataFrame& df, const std::string ctx, int hmax, const int hmin, const double max_ooctx_meth_frac);

// /root/reference/src/rcpp_get_base_freqs.cpp:15, /root/reference/src/rcpp_fep.cpp:9
Rcpp::NumericMatrix rcpp_get_base_freqs(Rcpp::DataFrame& df, std::vector<bool> pass, Rcpp::DataFrame& vcf);
std::vector<double> rcpp_fep(Rcpp::DataFrame& df, std::vector<std::string> colnames);

// HTSlib's base-code table (hts.c: seq_nt16_int): 4-bit BAM base code "=ACMGRSVTWYHKDBN" -> A,C,G,T = 0..3, else 4
extern "C" const int seq_nt16_int[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};

// Fisher's exact test as HTSlib's kfunc.c publishes it (restated, HTSlib is absent here): probabilities of the 2x2
// tables with the observed margins, accumulated from the smallest and from the largest n11 while they stay below the
// observed table's probability (relative guards 0.99999999 / 1.00000001); two-sided p = left tail + right tail.
namespace {
double lbinom(int n, int k) { return (n == 0 || k == 0 || n == k) ? 0.0 : std::lgamma(n + 1.0) - std::lgamma(k + 1.0) - std::lgamma(n - k + 1.0); }
double hypergeo(int n11, int n1_, int n_1, int n) { return std::exp(lbinom(n1_, n11) + lbinom(n - n1_, n_1 - n11) - lbinom(n, n_1)); }
struct HgAcc { int n11, n1_, n_1, n; double p; };
double hypergeo_acc(int n11, int n1_, int n_1, int n, HgAcc* a) {
  if (n1_ || n_1 || n) {
    a->n11 = n11; a->n1_ = n1_; a->n_1 = n_1; a->n = n;
  } else {  // step from the previous table
    if (n11 % 11 && n11 + a->n - a->n1_ - a->n_1) {
      if (n11 == a->n11 + 1) {
        a->p *= (double)(a->n1_ - a->n11) / n11 * (a->n_1 - a->n11) / (n11 + a->n - a->n1_ - a->n_1);
        a->n11 = n11;
        return a->p;
      }
      if (n11 == a->n11 - 1) {
        a->p *= (double)a->n11 / (a->n1_ - n11) * (a->n11 + a->n - a->n1_ - a->n_1) / (a->n_1 - n11);
        a->n11 = n11;
        return a->p;
      }
    }
    a->n11 = n11;
  }
  a->p = hypergeo(a->n11, a->n1_, a->n_1, a->n);
  return a->p;
}
}  // namespace
extern "C" double kt_fisher_exact(int n11, int n12, int n21, int n22, double* _left, double* _right, double* two) {
  const int n1_ = n11 + n12, n_1 = n11 + n21, n = n11 + n12 + n21 + n22;
  const int max = n_1 < n1_ ? n_1 : n1_;
  int min = n1_ + n_1 - n;
  if (min < 0) min = 0;
  *two = *_left = *_right = 1.0;
  if (min == max) return 1.0;
  HgAcc aux;
  const double q = hypergeo_acc(n11, n1_, n_1, n, &aux);
  if (q == 0.0) {  // the observed table is (numerically) impossible: the tails cannot be compared with it
    if (n11 * (n + 2) < (n_1 + 1) * (n1_ + 1)) { *_left = 0.0; *_right = 1.0; *two = 0.0; }
    else { *_left = 1.0; *_right = 0.0; *two = 0.0; }
    return 0.0;
  }
  double p = hypergeo_acc(min, 0, 0, 0, &aux), left, right;
  int i, j;
  for (left = 0.0, i = min + 1; p < 0.99999999 * q && i <= max; ++i) { left += p; p = hypergeo_acc(i, 0, 0, 0, &aux); }
  --i;
  if (p < 1.00000001 * q) left += p; else --i;
  p = hypergeo_acc(max, 0, 0, 0, &aux);
  for (right = 0.0, j = max - 1; p < 0.99999999 * q && j >= 0; --j) { right += p; p = hypergeo_acc(j, 0, 0, 0, &aux); }
  ++j;
  if (p < 1.00000001 * q) right += p; else ++j;
  *two = left + right;
  if (*two > 1.0) *two = 1.0;
  if (std::abs(i - n11) < std::abs(j - n11)) right = 1.0 - left + q; else left = 1.0 - right + q;
  *_left = left; *_right = right;
  return q;
}

namespace {
struct RefBam {
  Rcpp::DataFrame df;
  std::vector<std::string>* seqxm = nullptr;
  long n = 0;
  // last report
  std::shared_ptr<std::vector<int>> cols[6];
  // last lMHL report: rname, strand, pos, context, coverage + length, lmhl
  std::shared_ptr<std::vector<int>> mcols[5];
  std::shared_ptr<std::vector<double>> mdcols[2];
};
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" {

// Build the reference's bam.data: four int columns + the seqxm string vector
// (strings given as one byte blob + nstr+1 offsets, in FILE order = templid).
void* ref_bam_new(long n, const int* rname, const int* strand, const int* start, const int* templid,
                  long nstr, const uint8_t* blob, const int64_t* off) {
  RefBam* b = new RefBam;
  b->n = n;
  b->df.set_column("rname", std::make_shared<std::vector<int>>(rname, rname + n));
  b->df.set_column("strand", std::make_shared<std::vector<int>>(strand, strand + n));
  b->df.set_column("start", std::make_shared<std::vector<int>>(start, start + n));
  b->df.set_column("templid", std::make_shared<std::vector<int>>(templid, templid + n));
  b->seqxm = new std::vector<std::string>;
  b->seqxm->reserve(static_cast<size_t>(nstr));
  for (long i = 0; i < nstr; i++)
    b->seqxm->emplace_back(reinterpret_cast<const char*>(blob) + off[i], static_cast<size_t>(off[i + 1] - off[i]));
  b->df.attr("seqxm_xptr") = static_cast<SEXP>(b->seqxm);
  return b;
}

void ref_bam_free(void* h) {
  RefBam* b = static_cast<RefBam*>(h);
  if (!b) return;
  delete b->seqxm;
  delete b;
}

// pass_out: nstr bytes (the reference sizes its result by the number of strings).
double ref_threshold_reads(void* h, const char* ctx_meth, const char* ctx_unmeth, const char* ooctx_meth,
                           const char* ooctx_unmeth, unsigned int min_n_ctx, double min_beta, double max_oo_beta,
                           uint8_t* pass_out) {
  RefBam* b = static_cast<RefBam*>(h);
  double t0 = now_s();
  std::vector<bool> r = rcpp_threshold_reads(b->df, ctx_meth, ctx_unmeth, ooctx_meth, ooctx_unmeth, min_n_ctx,
                                             min_beta, max_oo_beta);
  double t1 = now_s();
  for (size_t i = 0; i < r.size(); i++) pass_out[i] = r[i] ? 1 : 0;
  return t1 - t0;
}

double ref_get_xm_beta(void* h, const char* ctx_meth, const char* ctx_unmeth, double* beta_out) {
  RefBam* b = static_cast<RefBam*>(h);
  double t0 = now_s();
  std::vector<double> r = rcpp_get_xm_beta(b->df, ctx_meth, ctx_unmeth);
  double t1 = now_s();
  std::memcpy(beta_out, r.data(), r.size() * sizeof(double));
  return t1 - t0;
}

// pass: n ints (R logical encoding: 0 FALSE, non-zero TRUE, INT_MIN = NA).
// Returns the number of report rows; fetch them with ref_cx_fetch.
long ref_cx_report(void* h, const int* pass, const char* ctx, double* seconds) {
  RefBam* b = static_cast<RefBam*>(h);
  Rcpp::LogicalVector pv(std::make_shared<std::vector<int>>(pass, pass + b->n));
  double t0 = now_s();
  Rcpp::DataFrame res = rcpp_cx_report(b->df, pv, ctx);
  double t1 = now_s();
  if (seconds) *seconds = t1 - t0;
  static const char* names[6] = {"rname", "strand", "pos", "context", "meth", "unmeth"};
  for (int c = 0; c < 6; c++) b->cols[c] = res[names[c]].shared();
  return static_cast<long>(b->cols[0]->size());
}

void ref_cx_fetch(void* h, int* rname, int* strand, int* pos, int* context, int* meth, int* unmeth) {
  RefBam* b = static_cast<RefBam*>(h);
  int* dst[6] = {rname, strand, pos, context, meth, unmeth};
  for (int c = 0; c < 6; c++)
    if (b->cols[c] && !b->cols[c]->empty())
      std::memcpy(dst[c], b->cols[c]->data(), b->cols[c]->size() * sizeof(int));
}

// mode 0: rcpp_match_amplicon(df, bed, tolerance = param); mode 1: rcpp_match_capture(df, bed, min_overlap = param).
// bed columns: seqnames (factor code on the BAM's rname levels, NA_INTEGER if absent), start, end.  out: n ints
// (1-based BED row or NA_INTEGER).
double ref_match_target(void* h, int mode, long nbed, const int* chr, const int* start, const int* end, int param, int* out) {
  RefBam* b = static_cast<RefBam*>(h);
  Rcpp::DataFrame bed;
  bed.set_column("seqnames", std::make_shared<std::vector<int>>(chr, chr + nbed));
  bed.set_column("start", std::make_shared<std::vector<int>>(start, start + nbed));
  bed.set_column("end", std::make_shared<std::vector<int>>(end, end + nbed));
  double t0 = now_s();
  std::vector<int> r = mode == 0 ? rcpp_match_amplicon(b->df, bed, param) : rcpp_match_capture(b->df, bed, param);
  double t1 = now_s();
  std::memcpy(out, r.data(), r.size() * sizeof(int));
  return t1 - t0;
}

// rcpp_mhl_report(df, ctx, hmax, hmin, max_ooctx_meth_frac); returns the row count, rows fetched with ref_mhl_fetch.
long ref_mhl_report(void* h, const char* ctx, int hmax, int hmin, double max_ooctx, double* seconds) {
  RefBam* b = static_cast<RefBam*>(h);
  double t0 = now_s();
  Rcpp::DataFrame res = rcpp_mhl_report(b->df, ctx, hmax, hmin, max_ooctx);
  double t1 = now_s();
  if (seconds) *seconds = t1 - t0;
  static const char* names[5] = {"rname", "strand", "pos", "context", "coverage"};
  for (int c = 0; c < 5; c++) b->mcols[c] = res[names[c]].shared();
  b->mdcols[0] = res.double_column("length");
  b->mdcols[1] = res.double_column("lmhl");
  return static_cast<long>(b->mcols[0]->size());
}

void ref_mhl_fetch(void* h, int* rname, int* strand, int* pos, int* context, int* coverage, double* length, double* lmhl) {
  RefBam* b = static_cast<RefBam*>(h);
  int* dst[5] = {rname, strand, pos, context, coverage};
  for (int c = 0; c < 5; c++)
    if (b->mcols[c] && !b->mcols[c]->empty()) std::memcpy(dst[c], b->mcols[c]->data(), b->mcols[c]->size() * sizeof(int));
  double* dd[2] = {length, lmhl};
  for (int c = 0; c < 2; c++)
    if (b->mdcols[c] && !b->mdcols[c]->empty()) std::memcpy(dd[c], b->mdcols[c]->data(), b->mdcols[c]->size() * sizeof(double));
}

// rcpp_get_base_freqs(df, pass, vcf): pass = n bytes (0/1); vcf columns seqnames (factor code on the BAM's rname levels,
// NA_INTEGER if absent) and start; out = nvcf x 20 doubles, column-major (R's NumericMatrix).
double ref_base_freqs(void* h, const uint8_t* pass, long nvcf, const int* chr, const int* pos, double* out) {
  RefBam* b = static_cast<RefBam*>(h);
  Rcpp::DataFrame vcf;
  vcf.set_column("seqnames", std::make_shared<std::vector<int>>(chr, chr + nvcf));
  vcf.set_column("start", std::make_shared<std::vector<int>>(pos, pos + nvcf));
  std::vector<bool> pv(static_cast<size_t>(b->n));
  for (long i = 0; i < b->n; i++) pv[static_cast<size_t>(i)] = pass[i] != 0;
  double t0 = now_s();
  Rcpp::NumericMatrix res = rcpp_get_base_freqs(b->df, pv, vcf);
  double t1 = now_s();
  if (nvcf > 0) std::memcpy(out, res.shared()->data(), static_cast<size_t>(nvcf) * 20 * sizeof(double));
  return t1 - t0;
}

// rcpp_fep(df, colnames) over four int columns (NA_INTEGER = NA); out = n doubles, NaN where the reference returns NA.
void ref_fep(long n, const int* a, const int* b_, const int* c, const int* d, double* out) {
  Rcpp::DataFrame df;
  df.set_column("A", std::make_shared<std::vector<int>>(a, a + n));
  df.set_column("B", std::make_shared<std::vector<int>>(b_, b_ + n));
  df.set_column("C", std::make_shared<std::vector<int>>(c, c + n));
  df.set_column("D", std::make_shared<std::vector<int>>(d, d + n));
  std::vector<double> r = rcpp_fep(df, {"A", "B", "C", "D"});
  if (n > 0) std::memcpy(out, r.data(), static_cast<size_t>(n) * sizeof(double));
}

}  // extern "C"
