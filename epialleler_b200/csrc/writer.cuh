// Report writer: the text of the Bismark-like cytosine report as data.table::fwrite writes it
// (/root/reference/R/internal.R:273-287: tab separated, no quotes, header line, factor labels for rname / strand /
// context, plain decimal integers, "\n" line ends), formatted on the device from the report's six integer columns:
// one thread per row computes the line's length, an exclusive scan places the lines, one thread per row writes it.
#pragma once
#include "kernels.cuh"

namespace ea {

struct TextLevels {
  const char* chars;        // concatenated rname labels
  const int32_t* off;       // n_levels + 1 offsets into chars
  int32_t n_levels;
};

__device__ __forceinline__ int dec_digits(uint32_t v) {
  int d = 1;
  while (v >= 10u) { v /= 10u; d++; }
  return d;
}
__device__ __forceinline__ int int_chars(int32_t v) { return v < 0 ? 1 + dec_digits(0u - (uint32_t)v) : dec_digits((uint32_t)v); }
__device__ __forceinline__ char* put_int(char* p, int32_t v) {
  uint32_t u = (uint32_t)v;
  if (v < 0) { *p++ = '-'; u = 0u - u; }
  const int d = dec_digits(u);
  for (int i = d - 1; i >= 0; i--) { p[i] = (char)('0' + u % 10u); u /= 10u; }
  return p + d;
}
// context factor labels (src/rcpp_cx_report.cpp:150-152): code 2 CHH, 6 CHG, 7 CG; other codes never occur in a report
__device__ __forceinline__ int ctx_chars(int code) { return code == 7 ? 2 : 3; }

__global__ void k_text_len(ReportCols cols, int64_t n, TextLevels lv, int64_t* __restrict__ len_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int rn = cols.rname[i];
  const int name = (rn >= 1 && rn <= lv.n_levels) ? lv.off[rn] - lv.off[rn - 1] : int_chars(rn);
  len_out[i] = name + 1 + 1 + 1 + int_chars(cols.pos[i]) + 1 + ctx_chars(cols.ctx[i]) + 1 + int_chars(cols.meth[i]) + 1 + int_chars(cols.unmeth[i]) + 1;
}

__global__ void k_text_write(ReportCols cols, int64_t n, TextLevels lv, const int64_t* __restrict__ off, int64_t base, char* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  char* p = out + base + off[i];
  const int rn = cols.rname[i];
  if (rn >= 1 && rn <= lv.n_levels) {
    for (int k = lv.off[rn - 1]; k < lv.off[rn]; k++) *p++ = lv.chars[k];
  } else {
    p = put_int(p, rn);
  }
  *p++ = '\t';
  *p++ = cols.strand[i] == 2 ? '-' : '+';
  *p++ = '\t';
  p = put_int(p, cols.pos[i]);
  *p++ = '\t';
  const int code = cols.ctx[i];
  *p++ = 'C';
  if (code == 7) { *p++ = 'G'; } else { *p++ = 'H'; *p++ = code == 6 ? 'G' : 'H'; }
  *p++ = '\t';
  p = put_int(p, cols.meth[i]);
  *p++ = '\t';
  p = put_int(p, cols.unmeth[i]);
  *p++ = '\n';
}

}  // namespace ea
