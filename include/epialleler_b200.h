/*
 * epialleler_b200 -- C ABI of the B200-native cytosine-report hot path.
 *
 * This is the drop-in boundary: what the reference's three `.Call` routines for this path
 * would bind instead of their CPU loops.  Plain C, plain pointers and sizes; no C++/torch/R
 * types cross it.  All functions return EA_OK (0) or a negative EA_ERR_* code;
 * ea_last_error() gives the message (thread-local).  The library must be called from one
 * host thread per handle; it never calls back into the caller except through `ea_poll_fn`.
 *
 * Reference interface each entry point replaces (paths relative to BBCG/epialleleR):
 *
 *   ea_pack / ea_pack_blob   the `bam.data` object: data.table {rname,strand,start,templid} + attr
 *                            "seqxm_xptr" -> std::vector<std::string>
 *                            (src/rcpp_read_bam.cpp:45,61-69,166-189; R/internal.R:193-195)
 *   ea_threshold_reads       rcpp_threshold_reads   src/rcpp_threshold_reads.cpp:15-74
 *                                                   src/RcppExports.cpp:234-251, R/RcppExports.R:64-66
 *   ea_get_xm_beta           rcpp_get_xm_beta       src/rcpp_get_xm_beta.cpp:10-43
 *                                                   src/RcppExports.cpp:97-109,  R/RcppExports.R:28-30
 *   ea_cx_report             rcpp_cx_report         src/rcpp_cx_report.cpp:34-159
 *                                                   src/RcppExports.cpp:39-51,   R/RcppExports.R:12-14
 *   ea_cytosine_report       the body of generateCytosineReport after preprocessBam:
 *                            .thresholdReads + .getCytosineReport fused into one pass over XM
 *                            (R/generateCytosineReport.R:181-201; R/internal.R:439-456,486-500)
 *
 * Data contract (identical to the reference's):
 *   - rows are sorted by (rname, start)  [R/internal.R:195]; unsorted input is rejected
 *     (the reference documents "PRE-SORTED DATASET IS A REQUIREMENT", src/rcpp_cx_report.cpp:19);
 *   - strings are one byte per reference position, low nibble = methylation-call code
 *     ((c+2)>>2)&15 of the XM character (src/epialleleR.h:15-28), high nibble = base (ignored here);
 *   - templid[x] indexes the strings (file order); nrow must equal the number of strings for the
 *     per-read calls, as in the reference (src/rcpp_threshold_reads.cpp:27-28);
 *   - low-nibble codes outside {2,5,6,7,10,11,12,13,14,15} cannot be produced by the reference's
 *     readers; the packer rejects them (EA_ERR_DOMAIN) instead of emulating the reference's
 *     undefined counter aliasing for them.
 */
#ifndef EPIALLELER_B200_H
#define EPIALLELER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EA_OK 0
#define EA_ERR_ARG (-1)      /* bad argument (null pointer, negative size, templid out of range ...) */
#define EA_ERR_CUDA (-2)     /* CUDA runtime failure (message has the cudaError string) */
#define EA_ERR_UNSORTED (-3) /* rows not sorted by (rname, start) */
#define EA_ERR_DOMAIN (-4)   /* a string holds a low-nibble code the reference's readers never emit */
#define EA_ERR_NODEVICE (-5) /* no CUDA device: there is NO CPU fallback */
#define EA_ERR_INTERRUPT (-6)/* the poll callback asked to stop */
#define EA_ERR_STATE (-7)    /* call order (e.g. fetch before report) */
#define EA_ERR_NEEDS_BYTES (-8) /* the handle was packed by ea_pack_blob_calls (call planes only): this call needs ea_pack_blob */

typedef struct ea_handle ea_handle;

/* Optional interrupt poll (the reference polls Rcpp::checkUserInterrupt every 2^16..2^20 rows,
 * src/rcpp_cx_report.cpp:110).  Return non-zero to abort; checked between kernel launches. */
typedef int (*ea_poll_fn)(void);

int ea_version(void);
const char* ea_last_error(void);
int ea_device_count(void);
void ea_set_poll(ea_poll_fn fn);                         /* process-wide */
void ea_set_handle_poll(ea_handle* h, ea_poll_fn fn);    /* this handle only; wins over the process-wide callback */
/* Kernel choice for the report: 0 automatic (default: the row-per-lane kernel for one-context reports over rows of at
 * most 256 bytes that are multiples of 32 bytes, else the chunk-per-lane kernel for reads up to 1 KB, else the long-read
 * kernel (flags from the stand-alone threshold pass), else - more than 32767 rows over one window - the general
 * lane-group kernel), 1 always the lane-group kernel, 2 prefer the thread-per-row TMA kernel (reads up to 512 bytes),
 * 3 prefer the chunk-per-lane kernel, 4 prefer the long-read kernel (any read length; thresholding runs as its own pass).
 * 5 automatic over the byte rows only, 6 prefer the vertical-counter kernel; 0 (default) prefers the kernels on the planar
 * call planes.  A negative value only queries.  Returns the previous policy.  Results are identical either way; used by the
 * parity tests and A/B timing. */
int ea_set_kernel_policy(int policy);

/* ---- packer: host bam.data -> device-resident CSR (row order) ------------------------------ */

/* str[i]/len[i]: the i-th string of the reference's std::vector<std::string> (file order). */
int ea_pack(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid,
            int64_t nrow, const uint8_t* const* str, const int64_t* len, int64_t nstr, int device,
            ea_handle** out);

/* Same, strings given as one blob + nstr+1 offsets (file order).  When templid is the identity
 * the blob is copied to the device as is (no host gather). */
int ea_pack_blob(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid,
                 int64_t nrow, const uint8_t* blob, const int64_t* off, int64_t nstr, int device,
                 ea_handle** out);

/* Calls-only variant of ea_pack_blob for the cytosine-report path (thresholds, betas, reports, BED matching): the host
 * transposes the strings into the library's bit-sliced call planes with all its cores while the previous chunk crosses
 * PCIe, so 4 bits per position travel instead of 8 (the base code in the high nibble, which this path never reads -
 * `& 15`, src/rcpp_cx_report.cpp:117 - is dropped).  Needs rows == strings in file order (templid NULL or the identity, what
 * preprocessBam produces); anything else is handed to ea_pack_blob.  Calls that need the byte rows (ea_base_freqs,
 * ea_mhl_report, ea_export_host, non-default kernel policies, unusual letter sets) return EA_ERR_NEEDS_BYTES on such a
 * handle.  EA_HOST_PACK_THREADS overrides the number of host threads (default: all, at most 32). */
int ea_pack_blob_calls(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid,
                       int64_t nrow, const uint8_t* blob, const int64_t* off, int64_t nstr, int device,
                       ea_handle** out);
/* Scattered strings (one heap allocation per template in pageable memory: the reference's std::vector<std::string>,
 * src/rcpp_read_bam.cpp:185) -> call planes only.  All host threads transpose the strings straight from where they lie
 * into a small ring of page-locked blocks while the previous block crosses PCIe: nothing is gathered, half the bytes
 * travel, and a first call pins ~100 MB instead of the whole data set.  Same arguments as ea_pack; same handle kind as
 * ea_pack_blob_calls.  What the R shim calls for generateCytosineReport. */
int ea_pack_calls(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid,
                  int64_t nrow, const uint8_t* const* str, const int64_t* len, int64_t nstr, int device, ea_handle** out);
/* The same without pointer tables: `records` = nstr fixed-size records (the std::string OBJECTS of the reader's vector) with
 * the data pointer at ptr_offset and the length (size_t) at len_offset inside each record of `stride` bytes.  The binding
 * finds the two offsets of its standard library with a probe string (integration/harness/shim_harness.cpp). */
int ea_pack_calls_records(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid,
                          int64_t nrow, const void* records, int64_t nstr, int64_t stride, int64_t ptr_offset, int64_t len_offset,
                          int device, ea_handle** out);
int ea_has_bytes(const ea_handle* h);            /* 1: byte rows present (ea_pack, ea_pack_blob, ...), 0: call planes only */
int64_t ea_pack_h2d_bytes(const ea_handle* h);   /* bytes the packer copied host -> device for this handle */

/* Adopt CSR arrays that already live on `device` in ROW order (d_off: nrow+1 int64, d_strand: 1|2 as
 * uint8).  The handle takes ownership iff take_ownership != 0 (pointers must come from cudaMalloc). */
int ea_attach_device(int32_t* d_rname, uint8_t* d_strand, int32_t* d_start, uint8_t* d_data, int64_t* d_off,
                     int64_t nrow, int device, int take_ownership, ea_handle** out);

void ea_free(ea_handle* h);
/* Device memory of freed handles stays cached in the device's stream-ordered pool (re-packing tens of GB per call
 * would otherwise pay cudaMalloc/cudaFree each time); ea_trim returns the cached blocks to the driver. */
int ea_trim(int device);

/* Page-locked host memory for result buffers (device->host copies into it run at full PCIe speed).  Freed blocks are
 * cached and reused by later ea_host_alloc calls; ea_trim releases the cache.  NULL on failure (see ea_last_error). */
void* ea_host_alloc(size_t bytes);
void ea_host_free(void* p);

int64_t ea_nrow(const ea_handle* h);
int64_t ea_nbytes(const ea_handle* h); /* sum of string lengths */
int ea_device(const ea_handle* h);

/* ---- the three reference routines ----------------------------------------------------------- */

/* pass_out: nrow bytes, 1 = read passes.  Letter strings exactly as the reference takes them
 * (ooctx_* may be empty; a letter repeated in one string counts twice, as in the reference). */
int ea_threshold_reads(ea_handle* h, const char* ctx_meth, const char* ctx_unmeth, const char* ooctx_meth,
                       const char* ooctx_unmeth, uint32_t min_n_ctx, double min_ctx_meth_frac,
                       double max_ooctx_meth_frac, uint8_t* pass_out);

int ea_get_xm_beta(ea_handle* h, const char* ctx_meth, const char* ctx_unmeth, double* beta_out);

/* pass: nrow bytes (0 = FALSE, anything else incl. R's NA = TRUE), or NULL for all TRUE.
 * Two-phase: this call computes the report on the device and returns the row count; the caller
 * allocates six int32 columns of that length and calls ea_cx_report_fetch (any column may be NULL). */
int ea_cx_report(ea_handle* h, const uint8_t* pass, const char* ctx, int64_t* nrow_out);
int ea_cx_report_fetch(ea_handle* h, int32_t* rname, int32_t* strand, int32_t* pos, int32_t* context,
                       int32_t* meth, int32_t* unmeth);

/* Report writer: the last report of this handle as the text .writeReport produces through
 * data.table::fwrite (R/internal.R:273-287): tab separated, header line if header != 0, factor labels for rname (taken
 * from rname_levels, the levels of bam$rname), strand (+ -) and context (CHH CHG CG), decimal integers, "\n" line ends.
 * Formatted on the device; two-phase like the row fetch: byte count first, then the caller's buffer is filled. */
int ea_cx_report_text(ea_handle* h, const char* const* rname_levels, int32_t n_levels, int header, int64_t* nbytes_out);
int ea_cx_report_text_fetch(ea_handle* h, char* out);

/* ---- fused generateCytosineReport body ------------------------------------------------------- */

typedef struct ea_report_params {
  int32_t threshold_reads;      /* 0: pass = all TRUE (R/generateCytosineReport.R:194) */
  const char* ctx_meth;         /* the four threshold letter sets (R/internal.R:54-65) */
  const char* ctx_unmeth;
  const char* ooctx_meth;
  const char* ooctx_unmeth;
  uint32_t min_n_ctx;           /* min.context.sites */
  double min_ctx_meth_frac;     /* min.context.beta */
  double max_ooctx_meth_frac;   /* max.outofcontext.beta */
  const char* report_ctx;       /* ctx.meth letters of report.context, e.g. "Z" or "ZXH" */
} ea_report_params;

/* One pass over the XM bytes: per-read pass flags and the report.  pass_out may be NULL.
 * Rows are fetched with ea_cx_report_fetch. */
int ea_cytosine_report(ea_handle* h, const ea_report_params* p, uint8_t* pass_out, int64_t* nrow_out);

/* ---- BED-assisted report ------------------------------------------------ *
 * ea_match_target replaces rcpp_match_amplicon (mode 0, param = tolerance) and rcpp_match_capture (mode 1, param =
 * min overlap): src/rcpp_match_target.cpp:16-45,51-81; R/RcppExports.R (rcpp_match_amplicon / rcpp_match_capture).
 * bed_chr: factor code of the BED row's chromosome on the BAM's rname levels (INT32_MIN = NA, as R/internal.R:467-468
 * produces), bed_start / bed_end 1-based closed.  match_out: nrow int32, 1-based BED row in FILE order (first match) or
 * INT32_MIN (NA); may be NULL (the result stays on the device for ea_bed_report). */
int ea_match_target(ea_handle* h, int mode, const int32_t* bed_chr, const int32_t* bed_start, const int32_t* bed_end, int64_t nbed,
                    int32_t param, int32_t* match_out);
/* Body of generateBedReport after preprocessBam (R/generateBedReport.R:243-262, R/internal.R:529-561) in one call: match,
 * per-read thresholds (thr->threshold_reads != 0; else the caller's `pass` flags, NULL = all TRUE) and the per-target
 * counts.  counts_out: (nbed + 1) * 4 int64, row i (row nbed = unmatched reads) holds reads by
 * {FALSE+, FALSE-, TRUE+, TRUE-} (pass, strand).  match_out / pass_out may be NULL. */
int ea_bed_report(ea_handle* h, int mode, const int32_t* bed_chr, const int32_t* bed_start, const int32_t* bed_end, int64_t nbed,
                  int32_t param, const ea_report_params* thr, const uint8_t* pass, int32_t* match_out, uint8_t* pass_out,
                  int64_t* counts_out);

/* ---- VCF-assisted report ------------------------------------------------- *
 * ea_base_freqs replaces rcpp_get_base_freqs (src/rcpp_get_base_freqs.cpp:15-61; R/RcppExports.R rcpp_get_base_freqs):
 * for every VCF entry the number of reads covering it, by base (high nibble of the packed byte through HTSlib's
 * seq_nt16_int: A,C,G,T,other), strand and pass flag.  pass: nrow bytes (NULL = all TRUE).  vcf_chr: factor code of the
 * entry's chromosome on the BAM's rname levels (INT32_MIN = NA: never matches), vcf_pos 1-based; the entries must be
 * sorted by (vcf_chr, vcf_pos) like the reference assumes (EA_ERR_UNSORTED otherwise).  freqs_out: n_vcf x 20 doubles,
 * COLUMN-major like R's NumericMatrix: U+A U+C U+G U+T U+N, U-..., M+..., M-... */
int ea_base_freqs(ea_handle* h, const uint8_t* pass, const int32_t* vcf_chr, const int32_t* vcf_pos, int64_t n_vcf, double* freqs_out);
/* rcpp_fep (src/rcpp_fep.cpp:9-36): two-sided Fisher's exact p of the 2x2 tables {a b; c d}, INT32_MIN in any cell ->
 * NaN (R's NA).  Host arithmetic on the report's few rows (the reference calls HTSlib's kt_fisher_exact here). */
int ea_fisher_exact(const int32_t* a, const int32_t* b, const int32_t* c, const int32_t* d, int64_t n, double* p_out);


/* ---- lMHL report ------------------------------------------------------- *
 * ea_mhl_report replaces rcpp_mhl_report (src/rcpp_mhl_report.cpp:46-228; R/RcppExports.R rcpp_mhl_report;
 * R/generateMhlReport.R:170-197): ctx = the haplotype context's methylated + unmethylated letters (e.g. "Zz"),
 * hmax = max.haplotype.window (<= 0: unlimited), hmin = min.haplotype.length, max_ooctx_meth_frac =
 * max.outofcontext.beta.  Two-phase like ea_cx_report: the call returns the row count, the caller allocates five int32
 * and two double columns of that length and fetches (any column may be NULL). */
int ea_mhl_report(ea_handle* h, const char* ctx, int32_t hmax, int32_t hmin, double max_ooctx_meth_frac, int64_t* nrow_out);
int ea_mhl_report_fetch(ea_handle* h, int32_t* rname, int32_t* strand, int32_t* pos, int32_t* context, int32_t* coverage,
                        double* length, double* lmhl);

/* ---- measurement hooks (device-resident, no host copies; used by bench.py) ------------------- */

typedef struct ea_timing {
  float total_ms;        /* CUDA-event time of the whole device pipeline of the call */
  float kernel_ms[8];    /* per-stage times, see ea_stage_name() */
  int32_t n_launches;    /* kernels launched by the call */
  int64_t n_rows;        /* report rows produced */
  int64_t n_pass;        /* reads that passed */
  int64_t alg_bytes;     /* algorithmic bytes of the call: sum(L) + 14 N + 24 R (13 N unthresholded) */
} ea_timing;

/* mode bit0: 0 fused kernel, 1 threshold kernel + report kernel (two passes over XM);
 * mode bit1: 0 automatic kernel choice, 1 force the general lane-group window kernel (A/B measurements). */
int ea_cytosine_report_device(ea_handle* h, const ea_report_params* p, int mode, ea_timing* t);
const char* ea_stage_name(int i);
/* Name of the window kernel the last report on this handle ran (diagnostics / bench provenance). */
const char* ea_report_kernel(const ea_handle* h);
/* Write `bytes` bytes of scratch on the handle's device (L2 flush between timed iterations). */
int ea_flush_l2(ea_handle* h, int64_t bytes);

/* ---- deterministic synthetic reads (same generator on host and device) ----------------------- */

/* ---- streaming ingestion: what push_template (src/rcpp_read_bam.cpp:61-69) would call instead of
 * seqxm->emplace_back.  Templates arrive one at a time in FILE order; their bytes go through a ring of page-locked blocks
 * to the device while the reader keeps decoding.  ea_ingest_finish orders the rows by (rname,start) with a stable sort on
 * the device - what R/internal.R:195 (setorder) does to the reader's table - packs them, and hands the sorted columns
 * back (nullable arrays of ea_ingest count entries; templid = template index of every row).  The ingest object is
 * released by finish or abort. */
typedef struct ea_ingest ea_ingest;
int ea_ingest_begin(int device, int64_t expected_templates, ea_ingest** out);
/* returns the template's index (>= 0) or a negative EA_ERR_* code */
int64_t ea_ingest_append(ea_ingest* g, int32_t rname, int32_t strand, int32_t start, const uint8_t* bytes, int64_t len);
int ea_ingest_finish(ea_ingest* g, int32_t* rname_out, int32_t* strand_out, int32_t* start_out, int32_t* templid_out,
                     ea_handle** out);
void ea_ingest_abort(ea_ingest* g);

/* ---- range sharding: contiguous genomic ranges, halo rows, no data-path collective ----
 * key = (int64)rname << 32 | pos.  Shard r packs rows [row_lo, own_hi) (rows [row_lo, own_lo) are the halo: they start
 * before the cut but reach past it), reports pass flags for rows [own_lo, own_hi) and the report rows with
 * key_lo <= key < key_hi. */
typedef struct ea_shard { int64_t row_lo, own_lo, own_hi, key_lo, key_hi; } ea_shard;
/* Cuts the sorted rows into `world` ranges of about equal bytes, each cut snapped to the first row of its (rname,start)
 * run.  off/templid/nstr as in ea_pack_blob (templid may be NULL = identity). */
int ea_plan_shards(const int32_t* rname, const int32_t* start, const int32_t* templid, int64_t nrow, const int64_t* off,
                   int64_t nstr, int world, ea_shard* out);
/* Ownership filter ON THE DEVICE for a handle that holds one shard's rows: later ea_cytosine_report / ea_cx_report calls
 * keep the report rows with key in [key_lo, key_hi) (one contiguous run of the ordered report, found by two binary
 * searches on the device; only those rows are fetched) and ea_cytosine_report writes the pass flags of rows
 * [own_row_lo, nrow) to pass_out[0 ..].  key_lo < 0 clears the filter. */
int ea_set_owned_range(ea_handle* h, int64_t key_lo, int64_t key_hi, int64_t own_row_lo);

/* One process, several devices (what the single-threaded R process calls to use every GPU of the box): plans the shards,
 * packs shard i on devices[i] (one host thread per device; zero-copy slices of the blob when rows are the strings in file
 * order), runs the fused report on all devices at once and concatenates the owned rows in genomic order.
 * calls_only != 0 packs call planes only (ea_pack_blob_calls).  nrow must equal nstr. */
typedef struct ea_multi ea_multi;
int ea_pack_multi(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid, int64_t nrow,
                  const uint8_t* blob, const int64_t* off, int64_t nstr, const int* devices, int ndev, int calls_only,
                  ea_multi** out);
/* The same from the reader's scattered strings, described as ea_pack_calls_records describes them (call planes only). */
int ea_pack_multi_records(const int32_t* rname, const int32_t* strand, const int32_t* start, const int32_t* templid, int64_t nrow,
                          const void* records, int64_t nstr, int64_t stride, int64_t ptr_offset, int64_t len_offset,
                          const int* devices, int ndev, ea_multi** out);
int ea_multi_ndev(const ea_multi* m);
ea_handle* ea_multi_handle(ea_multi* m, int i);           /* shard i's handle (owned by m) */
int ea_multi_shard(const ea_multi* m, int i, ea_shard* out);
/* pass_out: nrow flags in row order (nullable); nrow_out: rows of the concatenated report */
int ea_multi_cytosine_report(ea_multi* m, const ea_report_params* p, uint8_t* pass_out, int64_t* nrow_out);
int ea_multi_report_fetch(ea_multi* m, int32_t* rname, int32_t* strand, int32_t* pos, int32_t* context, int32_t* meth,
                          int32_t* unmeth);
void ea_multi_free(ea_multi* m);

#define EA_SYNTH_MAX_CHR 32
typedef struct ea_synth_cfg {
  uint64_t seed;
  int64_t n_reads;          /* rows of the WHOLE data set */
  int32_t n_chr;
  int32_t len_min, len_max; /* read length uniform in [len_min, len_max] */
  int32_t n_islands;        /* 0: reads tile each chromosome uniformly; >0: target islands */
  int32_t island_len;
  int32_t reserved;
  int64_t chr_len[EA_SYNTH_MAX_CHR];
} ea_synth_cfg;

/* Generate rows [row_begin, row_end) of the data set directly in device memory and pack them. */
int ea_synth_device(const ea_synth_cfg* cfg, int64_t row_begin, int64_t row_end, int device, ea_handle** out);
/* Same rows on the host (CPU baseline / parity inputs): caller-allocated arrays; blob sized by a first
 * call with blob == NULL that only fills off[] (row_end-row_begin+1 entries). */
int ea_synth_host(const ea_synth_cfg* cfg, int64_t row_begin, int64_t row_end, int32_t* rname, int32_t* strand,
                  int32_t* start, int64_t* off, uint8_t* blob);
/* Copy a handle's CSR back to the host (used to build host inputs for the end-to-end timing). */
int ea_export_host(ea_handle* h, int32_t* rname, int32_t* strand, int32_t* start, int64_t* off, uint8_t* blob);

#ifdef __cplusplus
}
#endif
#endif /* EPIALLELER_B200_H */
