// Deterministic, counter-based synthetic read generator shared by host and device code.
//
// Every quantity is a pure function of (seed, stream, indices), so the GPU can generate a data
// set directly in HBM while the CPU baseline regenerates any sub-range of the same rows bit for
// bit.  The byte format is the reference's packed SEQ|XM string: high
// nibble = a base code, low nibble = methylation-call code ((c+2)>>2)&15 of the XM character
// (/root/reference/src/epialleleR.h:15-28, src/rcpp_read_bam.cpp:129).
#pragma once
#include <stdint.h>

#include "../../include/epialleler_b200.h"

#if defined(__CUDACC__)
#define EA_HD __host__ __device__ __forceinline__
#else
#define EA_HD inline
#endif

namespace easynth {

// row partition of the data set over chromosomes, derived on the host from ea_synth_cfg
struct Plan {
  uint64_t seed;
  int64_t n_reads;
  int32_t n_chr;
  int32_t len_min, len_max;
  int32_t n_islands, island_len;
  int64_t chr_len[EA_SYNTH_MAX_CHR];
  int64_t chr_row0[EA_SYNTH_MAX_CHR + 1];  // first row of each chromosome
  int32_t chr_islands[EA_SYNTH_MAX_CHR];   // islands per chromosome (island mode)
};

inline int make_plan(const ea_synth_cfg* c, Plan* p) {
  if (!c || c->n_chr < 1 || c->n_chr > EA_SYNTH_MAX_CHR || c->n_reads < 0) return -1;
  if (c->len_min < 1 || c->len_max < c->len_min) return -1;
  p->seed = c->seed; p->n_reads = c->n_reads; p->n_chr = c->n_chr;
  p->len_min = c->len_min; p->len_max = c->len_max;
  p->n_islands = c->n_islands; p->island_len = c->island_len;
  long double tot = 0;
  for (int k = 0; k < c->n_chr; k++) {
    if (c->chr_len[k] < (int64_t)c->len_max + 2 || c->chr_len[k] > 2000000000ll) return -1;
    p->chr_len[k] = c->chr_len[k];
    tot += (long double)c->chr_len[k];
  }
  long double acc = 0;
  p->chr_row0[0] = 0;
  for (int k = 0; k < c->n_chr; k++) {
    acc += (long double)c->chr_len[k];
    p->chr_row0[k + 1] = (k + 1 == c->n_chr) ? c->n_reads : (int64_t)((long double)c->n_reads * (acc / tot));
    int64_t isl = 0;
    if (c->n_islands > 0) {
      if (c->island_len < c->len_max + 1) return -1;
      isl = (int64_t)((long double)c->n_islands * ((long double)c->chr_len[k] / tot));
      if (isl < 1) isl = 1;
      if (isl * (int64_t)(c->island_len + c->len_max) > c->chr_len[k]) isl = c->chr_len[k] / (c->island_len + c->len_max);
      if (isl < 1) return -1;
    }
    p->chr_islands[k] = (int32_t)isl;
  }
  return 0;
}

EA_HD uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
EA_HD uint64_t hash3(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
  return mix64(mix64(seed ^ (stream * 0xD6E8FEB86659FD93ull)) ^ mix64(a * 0xA24BAED4963EE407ull + b));
}

enum { ST_GENOME = 1, ST_READ = 2, ST_BYTE = 3 };

struct RowInfo {
  int32_t rname;   // 1-based chromosome code
  int32_t strand;  // 1 '+', 2 '-'
  int32_t start;   // 1-based
  int32_t len;
  uint32_t flags;  // bit0 hyper-methylated epiallele, bit1 unconverted read
};

EA_HD RowInfo row_info(const Plan& p, int64_t row) {
  RowInfo r;
  int k = 0;
  while (k + 1 < p.n_chr && row >= p.chr_row0[k + 1]) k++;
  const int64_t idx = row - p.chr_row0[k];
  const int64_t nk = p.chr_row0[k + 1] - p.chr_row0[k];
  const uint64_t h = hash3(p.seed, ST_READ, (uint64_t)row, 0);
  r.rname = k + 1;
  r.len = p.len_min + (int32_t)((h >> 8) % (uint64_t)(p.len_max - p.len_min + 1));
  r.strand = 1 + (int32_t)(h & 1);
  r.flags = (((h >> 40) % 100) < 10 ? 1u : 0u) | ((((h >> 48) % 100) < 2) ? 2u : 0u);
  if (p.n_islands <= 0) {
    const int64_t span = p.chr_len[k] - p.len_max;  // > 0 by make_plan
    r.start = 1 + (int32_t)((idx * span) / nk);  // idx*span < 2^63 (n_reads <= 2^31, span < 2^31)
  } else {
    const int64_t ni = p.chr_islands[k];
    const int64_t t = (idx * ni) / nk;                       // island of this row
    const int64_t first = (t * nk + ni - 1) / ni;            // first row of island t
    const int64_t next = ((t + 1) * nk + ni - 1) / ni;
    const int64_t m = next - first > 0 ? next - first : 1;
    const int64_t pitch = p.chr_len[k] / ni;
    const int64_t origin = 1 + t * pitch;
    const int64_t room = p.island_len - p.len_min;
    r.start = (int32_t)(origin + ((idx - first) * room) / m);
  }
  return r;
}

// genome track: which cytosine context (0 none, 2 CHH, 6 CHG, 7 CG) lives at (chr,pos) and on which strand
EA_HD uint32_t site_of(const Plan& p, int32_t rname, int64_t pos) {
  const uint64_t g = hash3(p.seed, ST_GENOME, (uint64_t)rname, (uint64_t)pos);
  const uint32_t u = (uint32_t)(g % 10000u);
  if (u < 200u) return 7u | (1u << 4);  // CG anchor: '+' strand here, '-' strand at pos+1
  const uint64_t gp = hash3(p.seed, ST_GENOME, (uint64_t)rname, (uint64_t)(pos - 1));
  if ((uint32_t)(gp % 10000u) < 200u) return 7u | (2u << 4);
  const uint32_t strand = 1u + (uint32_t)((g >> 32) & 1u);
  if (u < 700u) return 6u | (strand << 4);
  if (u < 2100u) return 2u | (strand << 4);
  return 0u;
}

EA_HD uint8_t read_byte(const Plan& p, int64_t row, const RowInfo& r, int32_t j) {
  const uint64_t b = hash3(p.seed, ST_BYTE, (uint64_t)row, (uint64_t)j);
  const uint32_t hi = (1u << ((b >> 60) & 3u)) << 4;  // base code 1,2,4,8
  if ((b & 0xFFFFu) < 655u) return 0xFBu;             // ~1%: unsequenced / low quality ('N','-')
  const uint32_t site = site_of(p, r.rname, (int64_t)r.start + j);
  uint32_t ctx = site & 15u;
  if (ctx == 0u || (site >> 4) != (uint32_t)r.strand) return (uint8_t)(hi | 12u);  // '.'
  const uint32_t e = (uint32_t)((b >> 16) & 0xFFFFu);
  if (e < 328u) return (uint8_t)(hi | 12u);           // 0.5%: mismatch -> '.'
  if (e < 459u) ctx = (ctx == 7u) ? 6u : (ctx == 6u ? 2u : 7u);  // 0.2%: wrong context
  const uint32_t d = (uint32_t)((b >> 32) & 0xFFFFu);
  uint32_t thr;
  if (ctx == 7u) thr = (r.flags & 1u) ? 58982u : 3277u;  // 0.90 / 0.05
  else thr = (r.flags & 2u) ? 32768u : 655u;             // 0.50 / 0.01
  const uint32_t meth = d < thr ? 0u : 8u;               // lower-case (unmethylated) = code | 8
  return (uint8_t)(hi | ctx | meth);
}

}  // namespace easynth
