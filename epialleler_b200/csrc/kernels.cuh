// Hand-written sm_100a kernels of the cytosine-report hot path.
//
// Device data layout ("padded CSR", built by the packer in pack.cu):
//   data   : every read's packed SEQ|XM string, rows in (rname,start) order, each row starting on a
//            16-byte boundary and padded to a multiple of 16 bytes with code 3 (a low nibble the
//            reference's readers never emit; it belongs to no class below, so no edge masking is ever
//            needed).  The LAST pad byte carries the pad count in its high nibble (0x03 | n<<4), which
//            gives the true row end without reading len[].  One 16-byte "chunk" is the unit a lane
//            loads (LDG.128 / one TMA bulk copy per tile of rows).
//   off    : int64[n+1] byte offsets of the padded rows;  len : int32[n] true lengths
//   start  : int32[n] 1-based reference position of byte 0;  strand : uint8[n] 1 '+', 2 '-'
//   sp     : uint8[n] strand '-' flag (bit 0) | pad bytes of the row << 4 (true length without len[])
//   rname  : int32[n] chromosome factor code
//   windows: the covered part of every chromosome cut into windows of S positions; per window the
//            contiguous row range [lo,hi) that can overlap it (rows are sorted, so it is contiguous).
//
// Classification works on nibble-packed registers: the low nibbles of the 16 bytes of a chunk are
// packed into two 32-bit registers P0 (words 0,1) and P1 (words 2,3); the four bit planes of the
// call code are then aligned by shifts and every class membership is one LOP3 over 8 codes at a
// time.  Code table (src/epialleleR.h:15-28): H 2, U 5, X 6, Z 7, h 10, +- 11, . 12, u 13, x 14, z 15;
// bit3 = "unmethylated", (code&7) in {2,6,7} = CHH/CHG/CG.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ea {

constexpr int kThreads = 512;        // window kernel CTA size
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kNib0 = 0x11111111u;  // indicator positions of P0 results
constexpr uint32_t kNib1 = 0x22222222u;  // indicator positions of P1 results

// threshold contexts (R/internal.R:54-65)
enum { TC_CG = 0, TC_CHG = 1, TC_CHH = 2, TC_CXG = 3, TC_CX = 4, TC_NONE = 5 };
// report-context bit set: bit0 CHH(code 2), bit1 CHG(code 6), bit2 CG(code 7)
enum { W_CHH = 1, W_CHG = 2, W_CG = 4 };

struct Csr {
  const uint8_t* data;
  const int64_t* off;
  const int32_t* len;
  const int32_t* start;
  const uint8_t* strand;
  const uint8_t* sp;  // per row: strand '-' (bit 0) | pad bytes of the padded row << 4
  const int32_t* rname;
  int64_t n;
};

struct ThrParams {
  uint32_t min_n;
  double min_beta;
  double max_oo;
};

__device__ __forceinline__ uint4 ldg_chunk(const uint8_t* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

constexpr uint32_t kPadByte = 0x03u;

struct Planes {
  uint32_t last;            // raw bytes 12..15 of the chunk (byte 15 of a row's last chunk holds the pad count)
  uint32_t P0, P1;          // nibble-packed codes
  uint32_t A0, B0, C0, D0;  // bit planes of P0 aligned at bit 0 of every nibble (B0 masked)
  uint32_t A1, B1, C1, D1;  // bit planes of P1 aligned at bit 1 of every nibble (B1 masked)
};

// Merged indicator layout: bit i <-> nibble t=i>>2 of P(i&1); chunk byte = 8*(i&1) + 4*(t&1) + (t>>1).
__device__ __forceinline__ void make_planes(const uint4& v, Planes& p) {
  const uint32_t M = 0x0F0F0F0Fu;
  p.last = v.w;
  p.P0 = (v.x & M) | ((v.y << 4) & ~M);
  p.P1 = (v.z & M) | ((v.w << 4) & ~M);
  p.A0 = p.P0;       p.B0 = (p.P0 >> 1) & kNib0; p.C0 = p.P0 >> 2; p.D0 = p.P0 >> 3;
  p.A1 = p.P1 << 1;  p.B1 = p.P1 & kNib1;        p.C1 = p.P1 >> 1; p.D1 = p.P1 >> 2;
}

// number of pad bytes at the end of a row, read from its last chunk
__device__ __forceinline__ int pad_count(const Planes& p) { return ((p.last >> 24) & 15u) == kPadByte ? (int)(p.last >> 28) : 0; }

// every class below contains the (masked) B plane positively, so results are clean indicator masks
__device__ __forceinline__ uint32_t cls_cg(uint32_t A, uint32_t B, uint32_t C) { return A & B & C; }
__device__ __forceinline__ uint32_t cls_chg(uint32_t A, uint32_t B, uint32_t C) { return ~A & B & C; }
__device__ __forceinline__ uint32_t cls_chh(uint32_t A, uint32_t B, uint32_t C) { return ~A & B & ~C; }
__device__ __forceinline__ uint32_t cls_cyt(uint32_t A, uint32_t B, uint32_t C) { return B & (C | ~A); }

template <int TCTX>
__device__ __forceinline__ uint32_t cls_in(uint32_t A, uint32_t B, uint32_t C) {
  if (TCTX == TC_CG) return cls_cg(A, B, C);
  if (TCTX == TC_CHG) return cls_chg(A, B, C);
  if (TCTX == TC_CHH) return cls_chh(A, B, C);
  if (TCTX == TC_CXG) return B & C;
  if (TCTX == TC_CX) return cls_cyt(A, B, C);
  return 0u;
}
template <int TCTX>
__device__ __forceinline__ uint32_t cls_out(uint32_t A, uint32_t B, uint32_t C) {
  if (TCTX == TC_CG) return ~A & B;
  if (TCTX == TC_CHG) return B & ~(A ^ C);
  if (TCTX == TC_CHH) return B & C;
  if (TCTX == TC_CXG) return cls_chh(A, B, C);
  return 0u;
}
template <int WANT>
__device__ __forceinline__ uint32_t cls_want(uint32_t A, uint32_t B, uint32_t C) {
  uint32_t r = 0u;
  if (WANT == 7) return cls_cyt(A, B, C);
  if (WANT & W_CG) r |= cls_cg(A, B, C);
  if (WANT & W_CHG) r |= cls_chg(A, B, C);
  if (WANT & W_CHH) r |= cls_chh(A, B, C);
  return r;
}

// per-chunk class counts for thresholding: c0 = n_in | n_in_meth<<16, c1 = n_out | n_out_meth<<16
template <int TCTX>
__device__ __forceinline__ void count_chunk(const Planes& p, uint32_t& c0, uint32_t& c1) {
  const uint32_t t0 = cls_in<TCTX>(p.A0, p.B0, p.C0), t1 = cls_in<TCTX>(p.A1, p.B1, p.C1);
  c0 += __popc(t0 | t1) + (__popc((t0 & ~p.D0) | (t1 & ~p.D1)) << 16);
  if (TCTX != TC_CX) {
    const uint32_t o0 = cls_out<TCTX>(p.A0, p.B0, p.C0), o1 = cls_out<TCTX>(p.A1, p.B1, p.C1);
    c1 += __popc(o0 | o1) + (__popc((o0 & ~p.D0) | (o1 & ~p.D1)) << 16);
  }
}

// The reference's pass/fail chain, same order, same IEEE double divisions
// (src/rcpp_threshold_reads.cpp:39-70).
__device__ __forceinline__ bool eval_pass(uint32_t m, uint32_t all, uint32_t om, uint32_t oall, const ThrParams& t) {
  if (m == 0u) return false;
  if (all < t.min_n) return false;
  if ((double)m / (double)all < t.min_beta) return false;
  if (om > 0u) {
    if ((double)om / (double)oall > t.max_oo) return false;
  }
  return true;
}

// Sum v over the lanes of one row group (g consecutive lanes starting at lane-cib); all lanes of the
// warp must call.  Result valid in every lane of the group.
__device__ __forceinline__ uint32_t group_sum(uint32_t v, int g, int cib, int lane) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const uint32_t o = __shfl_down_sync(0xFFFFFFFFu, v, d);
    if (d < g && cib + d < g) v += o;
  }
  return __shfl_sync(0xFFFFFFFFu, v, lane - cib);
}

template <int WANT>
__device__ __forceinline__ int slot_of(uint32_t code7) {
  // position of the context among the requested ones, order CHH, CHG, CG
  if (code7 == 2u) return 0;
  if (code7 == 6u) return (WANT & W_CHH) ? 1 : 0;
  return ((WANT & W_CHH) ? 1 : 0) + ((WANT & W_CHG) ? 1 : 0);
}
template <int WANT>
__host__ __device__ constexpr int n_ctx() { return ((WANT >> 0) & 1) + ((WANT >> 1) & 1) + ((WANT >> 2) & 1); }

template <typename CT> struct CtTraits;
template <> struct CtTraits<uint32_t> {
  static constexpr uint32_t kMethOne = 1u << 16;
  __device__ static __forceinline__ uint32_t tot(uint32_t v) { return v & 0xFFFFu; }
  __device__ static __forceinline__ uint32_t meth(uint32_t v) { return v >> 16; }
};
template <> struct CtTraits<unsigned long long> {
  static constexpr unsigned long long kMethOne = 1ull << 32;
  __device__ static __forceinline__ uint32_t tot(unsigned long long v) { return (uint32_t)v; }
  __device__ static __forceinline__ uint32_t meth(unsigned long long v) { return (uint32_t)(v >> 32); }
};

// Shared-memory window state: covd = coverage difference array per strand (prefix-summed at the end),
// tm = per requested context and strand, packed {total calls of that context, methylated calls of
// passing reads}.  By SURVEY.md 8a invariants 3-4 this is all the reference's 32 counters per
// position amount to.
template <typename CT>
struct WinSmem {
  int32_t* covd;  // [2][Sp]
  CT* tm;         // [NC][2][S]
  int S, Sp;
};

// Turn one chunk into shared-memory events.  rel = window-relative position of the chunk's byte 0.
template <int WANT, typename CT>
__device__ __forceinline__ void push_chunk(const Planes& p, const WinSmem<CT>& w, int strand0, int rel, bool pass,
                                           bool first_in_win, bool is_last) {
  const int S = w.S;
  int32_t* covd = w.covd + strand0 * w.Sp;
  const uint32_t cand = cls_want<WANT>(p.A0, p.B0, p.C0) | cls_want<WANT>(p.A1, p.B1, p.C1);
  const uint32_t E = (p.A0 & p.B0 & ~p.C0 & p.D0) | (p.A1 & p.B1 & ~p.C1 & p.D1);  // code 11 bytes
  const uint32_t prev = (E << 8) | ((E >> 20) & 0x30u) | ((E >> 27) & 0x2u);        // "byte before me is 11"
  const uint32_t starts = E & ~prev, ends = prev & ~E;
  const bool end16 = (E >> 29) & 1u;
  if (first_in_win) atomicAdd(&covd[max(rel, 0)], 1);
  uint32_t ev = cand | starts | ends;
  while (ev) {
    const int i = __ffs(ev) - 1;
    const uint32_t bit = 1u << i;
    ev &= ev - 1u;
    const int q = rel + ((i & 1) << 3) + (i & 4) + (i >> 3);
    if (q < S) {
      const int qc = max(q, 0);
      if (starts & bit) atomicAdd(&covd[qc], -1);
      if (ends & bit) atomicAdd(&covd[qc], 1);
      if ((cand & bit) && q >= 0) {
        const uint32_t code = (((i & 1) ? p.P1 : p.P0) >> (i & 28)) & 15u;
        const int slot = slot_of<WANT>(code & 7u);
        CT inc = (CT)1;
        if (pass && code < 8u) inc += CtTraits<CT>::kMethOne;
        atomicAdd(&w.tm[(slot * 2 + strand0) * S + q], inc);
      }
    }
  }
  if (!is_last) {
    const int qe = rel + 16;
    if (end16 && qe < S) atomicAdd(&covd[max(qe, 0)], 1);
  } else if (!end16) {  // row end = true end (a '-' run reaching the end has already closed the coverage)
    const int qe = rel + 16 - pad_count(p);
    if (qe < S) atomicAdd(&covd[max(qe, 0)], -1);
  }
}

// exclusive scan of one int per thread over the CTA; returns the exclusive prefix, *total = CTA sum.
__device__ __forceinline__ int block_exscan(int v, int* total, int* s_warp /* kWarps+1 ints */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();  // protect s_warp reuse
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int x = lane < kWarps ? s_warp[lane] : 0;
    int xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, xi, d);
      if (lane >= d) xi += o;
    }
    if (lane < kWarps) s_warp[lane] = xi - x;
    if (lane == kWarps - 1) s_warp[kWarps] = xi;
  }
  __syncthreads();
  *total = s_warp[kWarps];
  return s_warp[warp] + inc - v;
}

struct WinTable {
  const int32_t* rname;  // chromosome code of the window
  const int32_t* pos0;   // first position of the window
  const int32_t* lo;     // candidate row range
  const int32_t* hi;
  int32_t n;
  int32_t S;
};

struct ReportOut {
  int4* scratch;                // {pos, strand | ctx<<8, meth, unmeth} in window-local order
  long long cap;                // rows the scratch can hold
  unsigned long long* cursor;   // bump allocator over scratch (total rows after the kernel)
  long long* w_base;            // per window: first scratch row
  int32_t* w_cnt;               // per window: rows
  unsigned long long* n_pass;   // reads that passed (fused mode)
};

// ------------------------------------------------------------------------------------------------
// Window kernel.  One CTA per genomic window: (1) every warp takes 32/g rows at a time, one 16-byte
// chunk per lane; FUSED: the row's class counts are reduced over its lane group, the reference's
// pass/fail chain is evaluated, owned rows store their flag; (2) cytosine calls of the requested
// contexts and '-' run boundaries become shared-memory atomics; (3) coverage is prefix-summed and
// every (pos,strand) with total_ctx > floor(cov/2) becomes a report row (src/rcpp_cx_report.cpp:58-85).
// ------------------------------------------------------------------------------------------------
template <int TCTX, int WANT, typename CT, bool FUSED>
__global__ void __launch_bounds__(kThreads)
k_window(Csr csr, WinTable wt, ThrParams thr, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
         int g, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_scan[kWarps + 1];
  __shared__ long long s_base;
  __shared__ int s_npass;

  const int S = wt.S, Sp = S + 4;
  WinSmem<CT> w;
  w.S = S; w.Sp = Sp;
  w.covd = reinterpret_cast<int32_t*>(smem_raw);
  w.tm = reinterpret_cast<CT*>(smem_raw + sizeof(int32_t) * 2 * Sp);
  const int n_words = (int)((sizeof(int32_t) * 2 * Sp + sizeof(CT) * NC * 2 * S) / 4);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / g;          // rows per warp step
  const int rib = lane / g;        // row in batch
  const int cib = lane - rib * g;  // chunk in batch
  const bool lane_on = rib < rpw;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_words / 4; i += kThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) s_npass = 0;
    __syncthreads();

    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];

    for (int base = lo + warp * rpw; base < hi; base += kWarps * rpw) {
      const int r = base + rib;
      const bool valid = lane_on && r < hi;
      int start = 0, strand0 = 0, nch = 0;
      const uint8_t* row = csr.data;
      if (valid) {
        const int64_t o0 = csr.off[r], o1 = csr.off[r + 1];
        start = csr.start[r];
        strand0 = csr.strand[r] - 1;
        nch = (int)((o1 - o0) >> 4);
        row = csr.data + o0;
      }
      const int rel0 = start - w0;                              // window-relative position of byte 0
      const bool overlaps = valid && rel0 + nch * 16 > 0;        // rel0 < S holds by construction of hi
      const int c_first = rel0 >= 0 ? 0 : ((-rel0) >> 4);        // first chunk reaching into the window
      int c_end = nch;                                           // chunks at or beyond the window end are not pushed
      if (rel0 + nch * 16 > S) c_end = (S - rel0 + 15) >> 4;

      if (FUSED) {
        // ---- whole-row class counts -> pass flag ----
        uint32_t c0 = 0, c1 = 0;
        Planes p0;           // planes of the lane's first chunk, kept for the push
        bool have0 = false;
        if (overlaps) {
          for (int c = cib; c < nch; c += g) {
            Planes p;
            make_planes(ldg_chunk(row + (int64_t)c * 16), p);
            count_chunk<TCTX>(p, c0, c1);
            if (c == cib) { p0 = p; have0 = true; }
          }
        }
        bool pass = true;
        if (TCTX != TC_NONE) {
          c0 = group_sum(c0, g, cib, lane);
          c1 = group_sum(c1, g, cib, lane);
          pass = eval_pass(c0 >> 16, c0 & 0xFFFFu, c1 >> 16, c1 & 0xFFFFu, thr);
          const bool owner = overlaps && cib == 0 && rel0 >= 0;
          if (owner) pass_out[r] = pass ? 1 : 0;
          const uint32_t np = __popc(__ballot_sync(0xFFFFFFFFu, owner && pass));
          if (lane == 0 && np) atomicAdd(&s_npass, (int)np);
        }
        // ---- push ----
        if (overlaps) {
          if (have0 && cib >= c_first && cib < c_end)
            push_chunk<WANT, CT>(p0, w, strand0, rel0 + cib * 16, pass, cib == c_first, cib == nch - 1);
          for (int c = cib + g; c < c_end; c += g) {
            if (c < c_first) continue;
            Planes p;
            make_planes(ldg_chunk(row + (int64_t)c * 16), p);
            push_chunk<WANT, CT>(p, w, strand0, rel0 + c * 16, pass, c == c_first, c == nch - 1);
          }
        }
      } else {
        if (overlaps) {
          const bool pass = pass_in ? (pass_in[r] != 0) : true;
          int c = c_first + cib;
          for (; c < c_end; c += g) {
            Planes p;
            make_planes(ldg_chunk(row + (int64_t)c * 16), p);
            push_chunk<WANT, CT>(p, w, strand0, rel0 + c * 16, pass, c == c_first, c == nch - 1);
          }
        }
      }
    }
    __syncthreads();

    // ---- coverage prefix sum + row detection; thread t owns positions [t*per, (t+1)*per) ----
    const int per = S / kThreads;  // S is a multiple of kThreads
    const int q0 = threadIdx.x * per;
    int sum0 = 0, sum1 = 0;
    for (int k = 0; k < per; k++) { sum0 += w.covd[q0 + k]; sum1 += w.covd[Sp + q0 + k]; }
    int tot0, tot1;
    int run0 = block_exscan(sum0, &tot0, s_scan);
    int run1 = block_exscan(sum1, &tot1, s_scan);
    // pass 1: count rows, remember which (pos,strand) emit
    uint32_t flags = 0;  // bit 2k+s
    int cov0 = run0, cov1 = run1;
    for (int k = 0; k < per; k++) {
      cov0 += w.covd[q0 + k];
      cov1 += w.covd[Sp + q0 + k];
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const uint32_t half = (uint32_t)(s ? cov1 : cov0) >> 1;
        bool emit = false;
#pragma unroll
        for (int c = 0; c < NC; c++) emit |= CtTraits<CT>::tot(w.tm[(c * 2 + s) * S + q0 + k]) > half;
        if (emit) flags |= 1u << (2 * k + s);
      }
    }
    int nrows_total;
    const int my_off = block_exscan(__popc(flags), &nrows_total, s_scan);
    if (threadIdx.x == 0) {
      const unsigned long long b = nrows_total ? atomicAdd(out.cursor, (unsigned long long)nrows_total) : 0ull;
      s_base = (long long)b;
      out.w_base[win] = (long long)b;
      out.w_cnt[win] = nrows_total;
      if (FUSED && TCTX != TC_NONE && s_npass) atomicAdd(out.n_pass, (unsigned long long)s_npass);
    }
    __syncthreads();
    const long long wbase = s_base;
    if (flags && wbase + nrows_total <= out.cap) {
      long long dst = wbase + my_off;
      cov0 = run0; cov1 = run1;
      for (int k = 0; k < per; k++) {
        cov0 += w.covd[q0 + k];
        cov1 += w.covd[Sp + q0 + k];
#pragma unroll
        for (int s = 0; s < 2; s++) {
          if (!((flags >> (2 * k + s)) & 1u)) continue;
          const uint32_t half = (uint32_t)(s ? cov1 : cov0) >> 1;
          int ctx_code = 0; uint32_t mt = 0, tt = 0;
          int slot = 0;
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            if (!((WANT >> cc) & 1)) continue;
            const CT v = w.tm[(slot * 2 + s) * S + q0 + k];
            if (CtTraits<CT>::tot(v) > half) { ctx_code = cc == 0 ? 2 : (cc == 1 ? 6 : 7); mt = CtTraits<CT>::meth(v); tt = CtTraits<CT>::tot(v); }
            slot++;
          }
          out.scratch[dst++] = make_int4(w0 + q0 + k, (s + 1) | (ctx_code << 8), (int)mt, (int)(tt - mt));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stand-alone per-read kernel: rcpp_threshold_reads (pass flags) and rcpp_get_xm_beta (betas).
// Same row-group scheme as the window kernel, no shared memory.
// ------------------------------------------------------------------------------------------------
template <int TCTX>
__global__ void __launch_bounds__(256)
k_threshold(Csr csr, ThrParams thr, int g, uint8_t* __restrict__ pass_out, double* __restrict__ beta_out,
            unsigned long long* n_pass) {
  const int lane = threadIdx.x & 31;
  const int rpw = 32 / g;
  const int rib = lane / g, cib = lane - rib * g;
  const bool lane_on = rib < rpw;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t my_pass = 0;
  for (int64_t base = warp_global * rpw; base < csr.n; base += n_warps * rpw) {
    const int64_t r = base + rib;
    const bool valid = lane_on && r < csr.n;
    uint32_t c0 = 0, c1 = 0;
    if (valid) {
      const int64_t o0 = csr.off[r], o1 = csr.off[r + 1];
      const int nch = (int)((o1 - o0) >> 4);
      const uint8_t* row = csr.data + o0;
      for (int c = cib; c < nch; c += g) {
        Planes p;
        make_planes(ldg_chunk(row + (int64_t)c * 16), p);
        count_chunk<TCTX>(p, c0, c1);
      }
    }
    c0 = group_sum(c0, g, cib, lane);
    c1 = group_sum(c1, g, cib, lane);
    if (valid && cib == 0) {
      const uint32_t m = c0 >> 16, all = c0 & 0xFFFFu;
      if (pass_out) {
        const bool pass = eval_pass(m, all, c1 >> 16, c1 & 0xFFFFu, thr);
        pass_out[r] = pass ? 1 : 0;
        my_pass += pass;
      }
      if (beta_out) beta_out[r] = (double)m / (double)(all ? all : 1u);
    }
  }
  if (n_pass) {
    for (int d = 16; d >= 1; d >>= 1) my_pass += __shfl_down_sync(0xFFFFFFFFu, my_pass, d);
    if (lane == 0 && my_pass) atomicAdd(n_pass, (unsigned long long)my_pass);
  }
}

// Arbitrary letter sets (anything the named contexts do not cover, repeated letters, rows longer than
// 65535 bytes): one thread per read, the reference's own 16-bin histogram and weighted sums
// (src/rcpp_threshold_reads.cpp:32-61).  wt[k][code] = how many times a letter of set k maps to code.
struct LetterWeights { uint8_t w[4][16]; };
__global__ void __launch_bounds__(256)
k_threshold_generic(Csr csr, LetterWeights lw, ThrParams thr, uint8_t* __restrict__ pass_out, double* __restrict__ beta_out,
                    unsigned long long* n_pass) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= csr.n) return;
  uint32_t hist[16];
#pragma unroll
  for (int i = 0; i < 16; i++) hist[i] = 0;
  const uint8_t* row = csr.data + csr.off[r];
  const int len = csr.len[r];
  for (int i = 0; i < len; i++) {
    const uint32_t code = row[i] & 15u;
#pragma unroll
    for (int k = 0; k < 16; k++) hist[k] += (code == (uint32_t)k);
  }
  uint32_t s[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int i = 0; i < 16; i++) s[k] += hist[i] * lw.w[k][i];
  if (pass_out) {
    const bool pass = eval_pass(s[0], s[0] + s[1], s[2], s[2] + s[3], thr);
    pass_out[r] = pass ? 1 : 0;
    if (pass && n_pass) atomicAdd(n_pass, 1ull);
  }
  if (beta_out) {
    const uint32_t all = s[0] + s[1];
    beta_out[r] = (double)s[0] / (double)(all ? all : 1u);
  }
}

// ------------------------------------------------------------------------------------------------
// Ordered output: windows are already in (rname,pos) order; w_dst = exclusive scan of w_cnt.
// ------------------------------------------------------------------------------------------------
// One CTA; thread t owns the contiguous chunk [t*per, (t+1)*per) of the windows: chunk sums -> one block-wide scan ->
// exclusive prefixes written chunk by chunk (two passes over 4 bytes per window, three barriers in all; the earlier
// 1024-windows-per-iteration loop took 106 us for 123 k windows, four barriers per iteration).
__global__ void __launch_bounds__(1024) k_scan_windows(const int32_t* __restrict__ cnt, long long* __restrict__ dst, int n,
                                                       unsigned long long* total_out = nullptr) {
  __shared__ long long s_part[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (((n + 1023) / 1024) + 3) & ~3;   // a multiple of 4: chunks start 16-byte aligned
  const int i0 = min(threadIdx.x * per, n), i1 = min(i0 + per, n);
  long long sum = 0;
  {
    int i = i0;
    for (; i + 4 <= i1; i += 4) {
      const int4 v = *reinterpret_cast<const int4*>(cnt + i);
      sum += (long long)v.x + v.y + v.z + v.w;
    }
    for (; i < i1; i++) sum += cnt[i];
  }
  long long inc = sum;
  for (int d = 1; d < 32; d <<= 1) {
    const long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_part[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    long long x = s_part[lane], xi = x;
    for (int d = 1; d < 32; d <<= 1) {
      const long long o = __shfl_up_sync(0xFFFFFFFFu, xi, d);
      if (lane >= d) xi += o;
    }
    s_part[lane] = xi - x;
    if (total_out && lane == 31) *total_out = (unsigned long long)xi;
  }
  __syncthreads();
  long long run = s_part[warp] + inc - sum;
  for (int i = i0; i < i1; i++) { dst[i] = run; run += cnt[i]; }
}

struct ReportCols { int32_t *rname, *strand, *pos, *ctx, *meth, *unmeth; };

// A warp takes 32 consecutive windows at a time: their (count, scratch base, destination) are one coalesced load per
// array, the rows of the 32 windows are then copied as ONE flattened run (a lane finds the window of its row with a
// five-step shuffle search over the warp's prefix sums) - the destination rows of consecutive windows are contiguous, so
// every store instruction writes 32 consecutive elements however few rows a window holds.
__global__ void __launch_bounds__(128)
k_gather_rows(const int4* __restrict__ scratch, const long long* __restrict__ w_base, const long long* __restrict__ w_dst,
              const int32_t* __restrict__ w_cnt, const int32_t* __restrict__ w_rname, int n_win, long long cap,
              ReportCols out, int group = 32 /* windows per warp step, 1..32 (fewer when there are few windows) */) {
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int w0 = warp * group; w0 < n_win; w0 += n_warps * group) {
    const int w = w0 + lane;
    int cnt = 0, rn = 0;
    long long src = 0, dst = 0;
    if (lane < group && w < n_win) {
      cnt = w_cnt[w]; src = w_base[w]; dst = w_dst[w]; rn = w_rname[w];
      if (src + cnt > cap || dst + cnt > cap) cnt = 0;  // capacity overflow: the host grows the buffers and reruns
    }
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += o;
    }
    const int ex = inc - cnt;
    const int total = __shfl_sync(FULL, inc, 31);
    for (int c0 = 0; c0 < total; c0 += 32) {
      const int c = c0 + lane;
      int j = 0;  // the window of row c: the last lane whose exclusive count is <= c (empty windows share their successor's)
#pragma unroll
      for (int step = 16; step; step >>= 1) {
        const int e = __shfl_sync(FULL, ex, j + step);
        if (e <= c) j += step;
      }
      const long long sj = __shfl_sync(FULL, src, j), dj = __shfl_sync(FULL, dst, j);
      const int ej = __shfl_sync(FULL, ex, j), rj = __shfl_sync(FULL, rn, j);
      if (c < total) {
        const int4 v = scratch[sj + (c - ej)];
        const long long d = dj + (c - ej);
        out.rname[d] = rj;
        out.strand[d] = v.y & 0xFF;
        out.pos[d] = v.x;
        out.ctx[d] = v.y >> 8;
        out.meth[d] = v.z;
        out.unmeth[d] = v.w;
      }
    }
  }
}

}  // namespace ea
