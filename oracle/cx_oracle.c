/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Plain-C restatement ("port") of the reference's cytosine-report hot path.  It is the
 * checker the GPU tests compare against on the GPU box (where /root/reference is absent);
 * tests/test_oracle.py pins it against the reference's own golden constants and against
 * the reference's unmodified sources compiled into oracle/_ref (see oracle/Makefile).
 * Parity status: PINNED (goldens of inst/unitTests/test_generateCytosineReport.R:32-60,
 * :117-156, :215-233, test_generateBedReport.R:22-35, test_simulateBam.R:55-89 all hold).
 *
 * What each function follows:
 *   code_of()               ctx_to_idx            /root/reference/src/epialleleR.h:28
 *   nibble_hist()           per-read 16-bin count /root/reference/src/rcpp_threshold_reads.cpp:32-37
 *   port_threshold_reads()  pass/fail chain       /root/reference/src/rcpp_threshold_reads.cpp:27-71
 *   port_get_xm_beta()      per-read beta         /root/reference/src/rcpp_get_xm_beta.cpp:18-40
 *   port_cx_report()        aggregation + flush   /root/reference/src/rcpp_cx_report.cpp:58-131
 *   port_base_freqs()       base counts at SNVs   /root/reference/src/rcpp_get_base_freqs.cpp:15-61
 *   port_fep()              Fisher's exact test   /root/reference/src/rcpp_fep.cpp:9-36 (HTSlib kt_fisher_exact, by definition)
 *
 * The data structure is deliberately NOT the reference's (sorted flat_map of 32-int arrays):
 * an open-addressing hash keyed by position plus a key sort at flush time.  Semantics kept:
 * flush when start > max_pos or rname changes, stale last-position carry-over after an
 * all-'-' read, NA pass == TRUE, rows emitted in ascending position, '+' before '-'.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
  long n, nstr;
  int *rname, *strand, *start, *templid;
  uint8_t* blob;
  int64_t* off;
  /* last report */
  long nrow, cap;
  int* col[6];
} PortBam;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static unsigned code_of(unsigned c) { return ((c + 2u) >> 2) & 15u; }

static int* dup_i(const int* p, long n) {
  int* r = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (n > 0) memcpy(r, p, sizeof(int) * (size_t)n);
  return r;
}

void* port_bam_new(long n, const int* rname, const int* strand, const int* start, const int* templid, long nstr,
                   const uint8_t* blob, const int64_t* off) {
  PortBam* b = (PortBam*)calloc(1, sizeof(PortBam));
  b->n = n;
  b->nstr = nstr;
  b->rname = dup_i(rname, n);
  b->strand = dup_i(strand, n);
  b->start = dup_i(start, n);
  b->templid = dup_i(templid, n);
  int64_t total = off[nstr];
  b->blob = (uint8_t*)malloc((size_t)(total > 0 ? total : 1));
  if (total > 0) memcpy(b->blob, blob, (size_t)total);
  b->off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nstr + 1));
  memcpy(b->off, off, sizeof(int64_t) * (size_t)(nstr + 1));
  return b;
}

void port_bam_free(void* h) {
  PortBam* b = (PortBam*)h;
  if (!b) return;
  free(b->rname); free(b->strand); free(b->start); free(b->templid); free(b->blob); free(b->off);
  for (int c = 0; c < 6; c++) free(b->col[c]);
  free(b);
}

static void nibble_hist(const PortBam* b, long x, unsigned hist[16]) {
  memset(hist, 0, 16 * sizeof(unsigned));
  long t = b->templid[x];
  const uint8_t* s = b->blob + b->off[t];
  int64_t len = b->off[t + 1] - b->off[t];
  for (int64_t i = 0; i < len; i++) hist[s[i] & 15u]++;
}

static unsigned letters_sum(const unsigned hist[16], const char* letters) {
  unsigned s = 0; /* one term per LETTER, so repeated letters count twice, as in the reference */
  for (const unsigned char* p = (const unsigned char*)letters; *p; p++) s += hist[code_of(*p)];
  return s;
}

double port_threshold_reads(void* h, const char* ctx_meth, const char* ctx_unmeth, const char* ooctx_meth,
                            const char* ooctx_unmeth, unsigned min_n_ctx, double min_beta, double max_oo_beta,
                            uint8_t* pass_out) {
  PortBam* b = (PortBam*)h;
  double t0 = now_s();
  unsigned hist[16];
  for (long x = 0; x < b->nstr; x++) {
    pass_out[x] = 0;
    nibble_hist(b, x, hist);
    unsigned m = letters_sum(hist, ctx_meth);
    if (m == 0) continue;
    unsigned all = m + letters_sum(hist, ctx_unmeth);
    if (all < min_n_ctx) continue;
    if ((double)m / all < min_beta) continue;
    unsigned om = letters_sum(hist, ooctx_meth);
    if (om > 0) {
      unsigned oall = om + letters_sum(hist, ooctx_unmeth);
      if ((double)om / oall > max_oo_beta) continue;
    }
    pass_out[x] = 1;
  }
  return now_s() - t0;
}

double port_get_xm_beta(void* h, const char* ctx_meth, const char* ctx_unmeth, double* beta_out) {
  PortBam* b = (PortBam*)h;
  double t0 = now_s();
  unsigned hist[16];
  for (long x = 0; x < b->nstr; x++) {
    nibble_hist(b, x, hist);
    unsigned m = letters_sum(hist, ctx_meth);
    unsigned all = m + letters_sum(hist, ctx_unmeth);
    if (all == 0) all = 1;
    beta_out[x] = (double)m / all;
  }
  return now_s() - t0;
}

/* ---- position-keyed counters: open addressing, linear probing ---- */
typedef struct { int64_t key; int cnt[32]; } Slot;
typedef struct { Slot* s; size_t cap, used; } PosMap;

static void pm_init(PosMap* m, size_t cap) {
  m->cap = cap; m->used = 0;
  m->s = (Slot*)malloc(sizeof(Slot) * cap);
  for (size_t i = 0; i < cap; i++) m->s[i].key = -1;
}
static size_t pm_hash(int64_t k, size_t cap) { return (size_t)((uint64_t)k * 0x9E3779B97F4A7C15ull >> 20) & (cap - 1); }
static Slot* pm_get(PosMap* m, int64_t key);
static void pm_grow(PosMap* m) {
  PosMap n; pm_init(&n, m->cap * 2);
  for (size_t i = 0; i < m->cap; i++) if (m->s[i].key >= 0) { Slot* d = pm_get(&n, m->s[i].key); memcpy(d->cnt, m->s[i].cnt, sizeof d->cnt); }
  free(m->s); *m = n;
}
static Slot* pm_get(PosMap* m, int64_t key) {
  if ((m->used + 1) * 4 > m->cap * 3) pm_grow(m);
  size_t i = pm_hash(key, m->cap);
  for (;;) {
    if (m->s[i].key == key) return &m->s[i];
    if (m->s[i].key < 0) { m->s[i].key = key; memset(m->s[i].cnt, 0, sizeof m->s[i].cnt); m->used++; return &m->s[i]; }
    i = (i + 1) & (m->cap - 1);
  }
}
static int cmp_i64(const void* a, const void* b) { int64_t x = *(const int64_t*)a, y = *(const int64_t*)b; return (x > y) - (x < y); }

static void emit(PortBam* b, int rname, int strand, int pos, int ctx, int meth, int unmeth) {
  if (b->nrow == b->cap) {
    b->cap = b->cap ? b->cap * 2 : 1024;
    for (int c = 0; c < 6; c++) b->col[c] = (int*)realloc(b->col[c], sizeof(int) * (size_t)b->cap);
  }
  int v[6] = {rname, strand, pos, ctx, meth, unmeth};
  for (int c = 0; c < 6; c++) b->col[c][b->nrow] = v[c];
  b->nrow++;
}

static void flush(PortBam* b, PosMap* m, int rname, const unsigned want[16]) {
  if (m->used) {
    int64_t* keys = (int64_t*)malloc(sizeof(int64_t) * m->used);
    size_t k = 0;
    for (size_t i = 0; i < m->cap; i++) if (m->s[i].key >= 0) keys[k++] = m->s[i].key;
    qsort(keys, k, sizeof(int64_t), cmp_i64);
    for (size_t i = 0; i < k; i++) {
      Slot* s = pm_get(m, keys[i]);
      for (int st = 0; st < 2; st++) {
        const int* v = s->cnt + 16 * st;
        int cov = v[9];
        if (cov == 0) continue;
        int half = cov / 2, ctx;
        if (v[12] > half) continue;
        else if (v[2] + v[10] > half) ctx = 2;
        else if (v[6] + v[14] > half) ctx = 6;
        else if (v[7] + v[15] > half) ctx = 7;
        else continue;
        if (want[ctx]) emit(b, rname, st + 1, (int)keys[i], ctx, v[ctx], v[ctx | 8]);
      }
    }
    free(keys);
    for (size_t i = 0; i < m->cap; i++) m->s[i].key = -1;
    m->used = 0;
  }
}

long port_cx_report(void* h, const int* pass, const char* ctx, double* seconds) {
  PortBam* b = (PortBam*)h;
  double t0 = now_s();
  unsigned want[16] = {0};
  for (const unsigned char* p = (const unsigned char*)ctx; *p; p++) want[code_of(*p)] = 1;
  b->nrow = 0;
  PosMap m; pm_init(&m, 1 << 12);
  int cur_rname = 0, max_pos = 0, last_pos = 0;
  for (long x = 0; x < b->n; x++) {
    int st = b->start[x];
    if (st > max_pos || b->rname[x] != cur_rname) {
      flush(b, &m, cur_rname, want);
      max_pos = 0;
      cur_rname = b->rname[x];
    }
    int shift = (b->strand[x] - 1) << 4;
    unsigned lower = (!pass[x]) ? 8u : 0u;
    long t = b->templid[x];
    const uint8_t* s = b->blob + b->off[t];
    int64_t len = b->off[t + 1] - b->off[t];
    for (int64_t i = 0; i < len; i++) {
      unsigned idx = (s[i] & 15u) | lower;
      if (idx == 11) continue;
      last_pos = st + (int)i;
      Slot* sl = pm_get(&m, (int64_t)(uint64_t)(int64_t)last_pos & 0x7FFFFFFFFFFFFFFFll);
      sl->cnt[idx + shift]++;
      sl->cnt[9 + shift]++;
    }
    if (max_pos < last_pos) max_pos = last_pos; /* last_pos may be stale after an all-'-' read */
  }
  flush(b, &m, cur_rname, want);
  free(m.s);
  if (seconds) *seconds = now_s() - t0;
  return b->nrow;
}

void port_cx_fetch(void* h, int* rname, int* strand, int* pos, int* context, int* meth, int* unmeth) {
  PortBam* b = (PortBam*)h;
  int* dst[6] = {rname, strand, pos, context, meth, unmeth};
  for (int c = 0; c < 6; c++) if (b->nrow) memcpy(dst[c], b->col[c], sizeof(int) * (size_t)b->nrow);
}

/* ------------------------------------------------------------------------------------------------
 * BED target matching, restating /root/reference/src/rcpp_match_target.cpp:
 *   mode 0, rcpp_match_amplicon (:16-45): first BED row i (file order) on the read's chromosome whose start is within
 *           `param` of the read's start OR whose end is within `param` of the read's end;
 *   mode 1, rcpp_match_capture (:51-81): first BED row on the read's chromosome with
 *           min(read_end, end_i) - max(read_start, start_i) + 1 >= param.
 * read_end = start + length(string) - 1 (:33,:68).  Result: 1-based BED row, or INT_MIN (R's NA_integer_).
 * ------------------------------------------------------------------------------------------------ */
double port_match_target(void* h, int mode, long nbed, const int* chr, const int* start, const int* end, int param, int* out) {
  PortBam* b = (PortBam*)h;
  double t0 = now_s();
  for (long x = 0; x < b->n; x++) {
    const int t = b->templid[x];
    const int rs = b->start[x];
    const int re = rs + (int)(b->off[t + 1] - b->off[t]) - 1;
    int res = INT32_MIN;
    for (long i = 0; i < nbed; i++) {
      if (b->rname[x] != chr[i]) continue;
      if (mode == 0) {
        if (abs(rs - start[i]) <= param || abs(re - end[i]) <= param) { res = (int)i + 1; break; }
      } else {
        const int lo = rs > start[i] ? rs : start[i], hi = re < end[i] ? re : end[i];
        if (hi - lo + 1 >= param) { res = (int)i + 1; break; }
      }
    }
    out[x] = res;
  }
  return now_s() - t0;
}

/* ------------------------------------------------------------------------------------------------
 * Linearised MHL report, restating /root/reference/src/rcpp_mhl_report.cpp:46-228.
 * Independent formulation: since flush boundaries are unobservable, positions are accumulated in a dense
 * per-chromosome-run array (rows are sorted by start) instead of a map.  Per read (:150-182): h = number of bytes
 * whose code is in `ctx`; methylated stretches (codes < 8 of `ctx`, uninterrupted by an unmethylated `ctx` byte) give
 * every byte between the first and the last methylated byte of the stretch the numerator S(min(size, hmax));
 * out-of-context methylated / unmethylated bytes decide the read's fate (skip if h < hmin or the fraction >
 * max_ooctx; NaN compares false).  Per kept read and non-'-' byte (:184-195): counters of the code, coverage, sum of h,
 * numerator, denominator S(min(h, hmax)).  Row rule as in the cytosine report (:72-101):
 * coverage = C + c, length = sum(h) / coverage, lmhl = numerator / denominator (doubles).
 * ------------------------------------------------------------------------------------------------ */
typedef struct { uint64_t cnt[16]; uint64_t hsum, num, den; } MhlPos;  /* per (position, strand) */
typedef struct { int* icol[5]; double* dcol[2]; long nrow, cap; } MhlOut;
static MhlOut g_mhl;  /* last report (single-threaded test helper) */

static uint64_t S_of(uint64_t n) { return n < 2 ? n : (n * (n + 1) * (n + 2)) / 6; }

static void mhl_emit(int rname, int strand, int pos, int ctx, int cov, double len, double lmhl) {
  MhlOut* o = &g_mhl;
  if (o->nrow == o->cap) {
    o->cap = o->cap ? o->cap * 2 : 1024;
    for (int c = 0; c < 5; c++) o->icol[c] = (int*)realloc(o->icol[c], sizeof(int) * (size_t)o->cap);
    for (int c = 0; c < 2; c++) o->dcol[c] = (double*)realloc(o->dcol[c], sizeof(double) * (size_t)o->cap);
  }
  int v[5] = {rname, strand, pos, ctx, cov};
  for (int c = 0; c < 5; c++) o->icol[c][o->nrow] = v[c];
  o->dcol[0][o->nrow] = len; o->dcol[1][o->nrow] = lmhl;
  o->nrow++;
}

static void mhl_flush(MhlPos* a, int base, int span, int rname, const unsigned inctx[16]) {
  for (int p = 0; p < span; p++)
    for (int st = 0; st < 2; st++) {
      const MhlPos* v = &a[2 * (size_t)p + st];
      uint64_t cov = 0;
      for (int k = 0; k < 16; k++) if (k != 11) cov += v->cnt[k];
      if (!cov) continue;
      const uint64_t half = cov / 2;
      int ctx;
      if (v->cnt[12] > half) continue;
      else if (v->cnt[2] + v->cnt[10] > half) ctx = 2;
      else if (v->cnt[6] + v->cnt[14] > half) ctx = 6;
      else if (v->cnt[7] + v->cnt[15] > half) ctx = 7;
      else continue;
      if (!inctx[ctx]) continue;
      const int c = (int)(v->cnt[ctx] + v->cnt[ctx | 8]);
      mhl_emit(rname, st + 1, base + p, ctx, c, (double)v->hsum / c, (double)v->num / (double)v->den);
    }
}

long port_mhl_report(void* h, const char* ctx, int hmax, int hmin, double max_ooctx, double* seconds) {
  PortBam* b = (PortBam*)h;
  double t0 = now_s();
  unsigned inctx[16] = {0};
  for (const unsigned char* p = (const unsigned char*)ctx; *p; p++) inctx[code_of(*p)] = 1;
  const uint64_t cap = hmax > 0 ? (uint64_t)(hmax < 65536 ? hmax : 65536) : 65536;
  g_mhl.nrow = 0;
  MhlPos* a = NULL; size_t acap = 0;
  uint64_t* nb = NULL; size_t nbcap = 0;
  int cur_rname = 0, base = 0, span = 0;  /* dense array covers positions [base, base + span) */
  for (long x = 0; x < b->n; x++) {
    const int st = b->start[x];
    if (b->rname[x] != cur_rname || st >= base + span) {  /* new chromosome or coverage gap: everything before is final */
      if (span) mhl_flush(a, base, span, cur_rname, inctx);
      cur_rname = b->rname[x]; base = st; span = 0;
    }
    const long t = b->templid[x];
    const uint8_t* s = b->blob + b->off[t];
    const int64_t len = b->off[t + 1] - b->off[t];
    if ((size_t)len > nbcap) { nbcap = (size_t)len * 2; nb = (uint64_t*)realloc(nb, sizeof(uint64_t) * nbcap); }
    memset(nb, 0, sizeof(uint64_t) * (size_t)len);
    uint64_t hsz = 0, msz = 0, oo[16] = {0};
    int64_t ms = 0, me = 0;
    for (int64_t i = 0; i < len; i++) {
      const unsigned code = s[i] & 15u;
      if (inctx[code]) {
        hsz++;
        if (code < 8) { if (!msz) ms = i; me = i; msz++; }
        else if (msz) { for (int64_t k = ms; k <= me; k++) nb[k] = S_of(msz < cap ? msz : cap); msz = 0; }
      } else oo[code]++;
    }
    const uint64_t om = oo[2] + oo[5] + oo[6] + oo[7], ou = oo[10] + oo[13] + oo[14] + oo[15];
    const double frac = (double)om / (double)(om + ou);
    if ((int)hsz < hmin || frac > max_ooctx) continue;
    if (msz) for (int64_t k = ms; k <= me; k++) nb[k] = S_of(msz < cap ? msz : cap);
    const int need = (int)(st - base + len);
    if (need > span) {
      if ((size_t)need * 2 > acap) { acap = (size_t)need * 4; a = (MhlPos*)realloc(a, sizeof(MhlPos) * acap); }
      memset(a + 2 * (size_t)span, 0, sizeof(MhlPos) * 2 * (size_t)(need - span));
      span = need;
    }
    const uint64_t den = S_of(hsz < cap ? hsz : cap);
    for (int64_t i = 0; i < len; i++) {
      const unsigned code = s[i] & 15u;
      if (code == 11) continue;
      MhlPos* v = &a[2 * (size_t)(st - base + i) + (b->strand[x] - 1)];
      v->cnt[code]++; v->hsum += hsz; v->num += nb[i]; v->den += den;
    }
  }
  if (span) mhl_flush(a, base, span, cur_rname, inctx);
  free(a); free(nb);
  if (seconds) *seconds = now_s() - t0;
  return g_mhl.nrow;
}

void port_mhl_fetch(void* h, int* rname, int* strand, int* pos, int* context, int* coverage, double* length, double* lmhl) {
  (void)h;
  int* dst[5] = {rname, strand, pos, context, coverage};
  for (int c = 0; c < 5; c++) if (g_mhl.nrow) memcpy(dst[c], g_mhl.icol[c], sizeof(int) * (size_t)g_mhl.nrow);
  if (g_mhl.nrow) { memcpy(length, g_mhl.dcol[0], sizeof(double) * (size_t)g_mhl.nrow); memcpy(lmhl, g_mhl.dcol[1], sizeof(double) * (size_t)g_mhl.nrow); }
}


/* ---- base frequencies at VCF positions ---------------------------------------------------
 * /root/reference/src/rcpp_get_base_freqs.cpp:15-61 walks reads and VCF entries with a shared cursor and ASSUMES both
 * are sorted.  Restated from the other side, without a cursor: per read, binary-search the sorted valid entries for the
 * ones inside [start, start+len-1] on the read's chromosome.  Entries whose chromosome is NA (INT_MIN) never match, but
 * they are not inert in the reference: its cursor treats them as "before the read" and moves onto them as soon as a
 * read's scan reaches them (rcpp_get_base_freqs.cpp:44-47), which hides every earlier entry from all LATER reads.  A
 * scan reaches an NA entry exactly when the read's end is not before the last valid entry in front of it, so:
 *   limit(i) = the first read (in row order) whose (rname, end) >= the key of the last valid entry before the first NA
 *              entry that follows entry i;   entry i is visible to reads 0..limit(i) only.
 * out: nvcf x 20 doubles, column-major; column = base (A,C,G,T = 0..3, anything else 4; HTSlib's seq_nt16_int of the
 * high nibble) + 5*(strand-1) + 10*pass. */
static const int kNt16Int[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4};

double port_base_freqs(void* h, const uint8_t* pass, long nvcf, const int* chr, const int* pos, double* out) {
  PortBam* b = (PortBam*)h;
  double t0 = now_s();
  for (long i = 0; i < nvcf * 20; i++) out[i] = 0.0;
  const size_t cap = (size_t)(nvcf > 0 ? nvcf : 1);
  long nv = 0;
  long* idx = (long*)malloc(sizeof(long) * cap);
  long* limit = (long*)malloc(sizeof(long) * cap);
  int64_t* key = (int64_t*)malloc(sizeof(int64_t) * cap);
  /* running maximum of the reads' (rname, end) keys: reads are sorted by (rname, start), ends are not */
  int64_t* endmax = (int64_t*)malloc(sizeof(int64_t) * (size_t)(b->n > 0 ? b->n : 1));
  int64_t run = INT64_MIN;
  for (long x = 0; x < b->n; x++) {
    const long t = b->templid[x];
    const int64_t len = b->off[t + 1] - b->off[t];
    const int64_t e = (((int64_t)b->rname[x] << 32) | (uint32_t)b->start[x]) + len - 1;
    if (e > run) run = e;
    endmax[x] = run;
  }
  for (long i = 0; i < nvcf; i++)
    if (chr[i] != INT32_MIN) { idx[nv] = i; key[nv] = ((int64_t)chr[i] << 32) | (uint32_t)pos[i]; nv++; }
  {  /* visibility limits, from the back: the nearest following NA entry decides */
    long cur_limit = LONG_MAX;   /* no NA entry behind: visible to every read */
    long v = nv;                 /* valid entries in front of position i */
    for (long i = nvcf - 1; i >= 0; i--) {
      if (chr[i] == INT32_MIN) {
        if (v == 0) { cur_limit = -1; continue; }  /* nothing valid in front of it: irrelevant for the entries behind */
        const int64_t K = key[v - 1];
        long a = 0, z = b->n;      /* first read whose running end reaches K */
        while (a < z) { const long m = (a + z) >> 1; if (endmax[m] < K) a = m + 1; else z = m; }
        cur_limit = a;             /* == n: no read reaches it */
      } else {
        v--;
        limit[v] = cur_limit;
      }
    }
  }
  for (long x = 0; x < b->n; x++) {
    const long t = b->templid[x];
    const int64_t len = b->off[t + 1] - b->off[t];
    if (len <= 0) continue;
    const int64_t lo_key = ((int64_t)b->rname[x] << 32) | (uint32_t)b->start[x];
    const int64_t hi_key = lo_key + (len - 1);
    long a = 0, z = nv;
    while (a < z) { const long m = (a + z) >> 1; if (key[m] < lo_key) a = m + 1; else z = m; }
    for (long k = a; k < nv && key[k] <= hi_key; k++) {
      if (x > limit[k]) continue;
      const uint8_t c = b->blob[b->off[t] + (key[k] - lo_key)];
      const int col = kNt16Int[c >> 4] + (b->strand[x] - 1) * 5 + (pass[x] ? 10 : 0);
      out[(long)col * nvcf + idx[k]] += 1.0;
    }
  }
  free(idx); free(key); free(limit); free(endmax);
  return now_s() - t0;
}

/* Two-sided Fisher's exact test by definition: the sum of the probabilities of all 2x2 tables with the observed margins
 * that are not more probable than the observed one (relative guard 1e-8, as HTSlib's kt_fisher_exact uses), every
 * probability computed directly from log-gamma.  NA (INT_MIN) in any cell -> NaN. */
static double lchoose_(int n, int k) { return lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(n - k + 1.0); }
void port_fep(long n, const int* a, const int* b_, const int* c, const int* d, double* out) {
  for (long x = 0; x < n; x++) {
    if (a[x] == INT32_MIN || b_[x] == INT32_MIN || c[x] == INT32_MIN || d[x] == INT32_MIN) { out[x] = NAN; continue; }
    const int n11 = a[x], n1_ = a[x] + b_[x], n_1 = a[x] + c[x], nn = a[x] + b_[x] + c[x] + d[x];
    const int mx = n_1 < n1_ ? n_1 : n1_;
    int mn = n1_ + n_1 - nn;
    if (mn < 0) mn = 0;
    if (mn == mx) { out[x] = 1.0; continue; }
    const double lden = lchoose_(nn, n_1);
    const double q = exp(lchoose_(n1_, n11) + lchoose_(nn - n1_, n_1 - n11) - lden);
    double two = 0.0;
    for (int i = mn; i <= mx; i++) {
      const double pi = exp(lchoose_(n1_, i) + lchoose_(nn - n1_, n_1 - i) - lden);
      if (pi < 1.00000001 * q) two += pi;
    }
    out[x] = two > 1.0 ? 1.0 : two;
  }
}
