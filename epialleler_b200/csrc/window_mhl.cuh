// Linearised MHL report: replaces rcpp_mhl_report (/root/reference/src/rcpp_mhl_report.cpp:46-228).
//
// Same windowing as the cytosine report (one CTA per genomic window, the window's sorted candidate rows, ordered
// compaction, scan + gather), one THREAD PER READ.  What the reference accumulates per (position, strand) with a map
// update per byte - coverage, sum of haplotype sizes, lMHL numerator and denominator - is piecewise constant along a
// read, so it is accumulated here as DIFFERENCE ARRAYS in shared memory (a few events per read: where its non-'-'
// coverage starts and stops, and around every methylated stretch) and prefix-summed once per window; only the
// cytosine calls themselves are per-position counts.  Per read:
//   1. nibble-plane counts over the whole row (like the threshold kernels): h = bytes whose code is in `ctx`,
//      out-of-context methylated / unmethylated bytes -> the read is dropped if h < hmin or the out-of-context beta
//      exceeds the limit (IEEE double division, NaN compares false; rcpp_mhl_report.cpp:179-182);
//   2. one byte-serial sweep: coverage transitions push (+-1, +-h, +-S(min(h,hmax))) into dW / dH / dDe, cytosine calls
//      count into T, methylated stretches of `ctx` bytes are tracked (:163-176); when a stretch closes, its span is
//      walked once more to push +-S(min(size,hmax)) into dNu around the span's non-'-' pieces.
// Row rule and outputs as the reference's spit_results (:72-101): a row for (pos, strand) iff one context c of `ctx`
// has C + c > floor(cov/2); coverage = C + c, length = sum(h) / coverage, lmhl = numerator / denominator (doubles).
// Preconditions (dispatcher): rows of at most 65535 bytes, at most 2^31 rows per window.
#pragma once
#include "kernels.cuh"
#include "win_common.cuh"

namespace ea {

constexpr int kMhlThreads = 256;
constexpr int kMhlWarps = kMhlThreads / 32;
constexpr int kMhlS = 1024;

struct MhlRow {  // one report row in window-local order
  int32_t pos, strand_ctx, coverage, pad;
  double length, lmhl;
};
struct MhlOut {
  MhlRow* scratch;
  long long cap;
  unsigned long long* cursor;
  long long* w_base;
  int32_t* w_cnt;
};
struct MhlParams {
  uint32_t ctx_mask;     // bit k: code k is in `ctx` (ctx_map of the reference)
  int32_t n_cls;         // cytosine classes (CHH, CHG, CG) that get a counter plane in shared memory: those with a letter in `ctx`
  int32_t slot[3];       // class -> plane (or -1: the class cannot produce a row, its calls are not counted)
  uint32_t cap;          // min(hmax, 65536), or 65536 when hmax <= 0 (:115)
  int32_t hmin;
  double max_ooctx;
};

__device__ __forceinline__ unsigned long long mhl_S(unsigned long long n) { return n < 2 ? n : (n * (n + 1) * (n + 2)) / 6; }  // nrS, :39-43

template <int NW>
__device__ __forceinline__ long long block_exscan_i64(long long v, long long* s_warp /* NW+1 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const long long x = lane < NW ? s_warp[lane] : 0;
    long long xi = x;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
      const long long o = __shfl_up_sync(0xFFFFFFFFu, xi, d);
      if (lane >= d) xi += o;
    }
    if (lane < NW) s_warp[lane] = xi - x;
  }
  __syncthreads();
  return s_warp[warp] + inc - v;
}

// exclusive block scan of four values at once (two barriers); s_warp: [4][NW] of long long.  Every warp scans the NW
// per-warp totals itself with shuffles (one shared load per value instead of NW).
template <int NW, typename T>
__device__ __forceinline__ void block_exscan4(T (&v)[4], long long* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc[4];
#pragma unroll
  for (int j = 0; j < 4; j++) inc[j] = v[j];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const T o = __shfl_up_sync(0xFFFFFFFFu, inc[j], d);
      if (lane >= d) inc[j] += o;
    }
  __syncthreads();  // protects s_warp against the previous call's readers
  if (lane == 31) {
#pragma unroll
    for (int j = 0; j < 4; j++) s_warp[j * NW + warp] = (long long)inc[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const T x = lane < NW ? (T)s_warp[j * NW + lane] : (T)0;
    T xi = x;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
      const T o = __shfl_up_sync(0xFFFFFFFFu, xi, d);
      if (lane >= d) xi += o;
    }
    const T before = __shfl_sync(0xFFFFFFFFu, xi - x, warp);   // sum of the warps in front of this one
    v[j] = before + inc[j] - v[j];
  }
}
template <int NW>
__device__ __forceinline__ void block_exscan4_i64(long long (&v)[4], long long* s_warp) { block_exscan4<NW, long long>(v, s_warp); }

// bit 0 of every byte of t (t & 0x01010101 == t) gathered into a nibble: bit k = byte k
__device__ __forceinline__ uint32_t gather4(uint32_t t) { return (t * 0x01020408u) >> 24; }

// shared-memory add on the accumulator type of the difference arrays: int (one native ATOMS.ADD) when the dispatcher has
// shown that every sum of the window fits 31 bits, else 64-bit
__device__ __forceinline__ void acc_add(long long* p, long long v) { atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v); }
__device__ __forceinline__ void acc_add(int* p, long long v) { atomicAdd(p, (int)v); }

// difference-array event: value v becomes active at window-relative position q (clamped to the window start),
// nothing to do at or beyond the window end
template <typename ACC>
__device__ __forceinline__ void diff_event(int q, int S, int* dW, ACC* dH, ACC* dDe, int w, long long hv, long long dv) {
  if (q >= S) return;
  q = max(q, 0);
  atomicAdd(&dW[q], w);
  acc_add(&dH[q], hv);
  acc_add(&dDe[q], dv);
}

// The window's difference arrays -> prefix sums in place, row test, ordered compaction into the scratch rows (shared by the
// byte-row kernel below and the planar kernel of window_mhl_pl.cuh).  Called by the whole CTA after a barrier.
template <typename ACC>
__device__ __forceinline__ void mhl_window_finalize(unsigned char* smem_raw, const MhlParams& prm, const MhlOut& out, int win, int w0,
                                                    long long* s_scan, uint32_t* s_scan32, long long* s_base_p) {
  constexpr int S = kMhlS;
  constexpr int PER = S / kMhlThreads;
  const int NCLS = prm.n_cls;
  ACC* const L64 = reinterpret_cast<ACC*>(smem_raw);
  int* const dWb = reinterpret_cast<int*>(L64 + 2 * 3 * S);
  uint32_t* const Tb = reinterpret_cast<uint32_t*>(dWb + 2 * S);
  long long& s_base = *s_base_p;
  // ---- finalize: the four difference arrays of each strand become their prefix sums IN PLACE (every thread owns
  //      PER consecutive positions), then the row test and the ordered compaction read them back ----
  const int q0 = threadIdx.x * PER;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    int* const pw = dWb + s * S + q0;
    ACC* const ph = L64 + (s * 3 + 0) * S + q0;
    ACC* const pd = L64 + (s * 3 + 1) * S + q0;
    ACC* const pn = L64 + (s * 3 + 2) * S + q0;
    ACC rw = 0, rh = 0, rd = 0, rn = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { rw += pw[k]; rh += ph[k]; rd += pd[k]; rn += pn[k]; }
    ACC ex[4] = {rw, rh, rd, rn};
    block_exscan4<kMhlWarps, ACC>(ex, s_scan);
    rw = ex[0]; rh = ex[1]; rd = ex[2]; rn = ex[3];
#pragma unroll
    for (int k = 0; k < PER; k++) {
      rw += pw[k]; pw[k] = (int)rw;
      rh += ph[k]; ph[k] = rh;
      rd += pd[k]; pd[k] = rd;
      rn += pn[k]; pn[k] = rn;
    }
  }
  uint32_t flags = 0;   // bit 2k+s
  uint32_t which = 0;   // 2 bits per (k,s): class of the row
#pragma unroll
  for (int k = 0; k < PER; k++)
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const int cv = dWb[s * S + q0 + k];
      if (cv <= 0) continue;
      const uint32_t half = (uint32_t)cv >> 1;
#pragma unroll
      for (int cls = 0; cls < 3; cls++) {
        const uint32_t code = cls == 0 ? 2u : (cls == 1 ? 6u : 7u);
        if (!((prm.ctx_mask >> code) & 1u)) continue;
        if (Tb[(s * NCLS + prm.slot[cls]) * S + q0 + k] > half) { flags |= 1u << (2 * k + s); which |= (uint32_t)cls << (2 * (2 * k + s)); }
      }
    }
  uint32_t nrows_u;
  const int my_off = (int)block_exscan_u32<kMhlWarps>((uint32_t)__popc(flags), &nrows_u, s_scan32);
  const int nrows_total = (int)nrows_u;
  if (threadIdx.x == 0) {
    const unsigned long long bb = nrows_total ? atomicAdd(out.cursor, (unsigned long long)nrows_total) : 0ull;
    s_base = (long long)bb;
    out.w_base[win] = (long long)bb;
    out.w_cnt[win] = nrows_total;
  }
  __syncthreads();
  const long long wbase = s_base;
  if (flags && wbase + nrows_total <= out.cap) {
    MhlRow* dst = out.scratch + wbase + my_off;
#pragma unroll
    for (int k = 0; k < PER; k++)
#pragma unroll
      for (int s = 0; s < 2; s++) {
        if (!((flags >> (2 * k + s)) & 1u)) continue;
        const int cls = (which >> (2 * (2 * k + s))) & 3;
        const int code = cls == 0 ? 2 : (cls == 1 ? 6 : 7);
        const int c = (int)Tb[(s * NCLS + prm.slot[cls]) * S + q0 + k];
        MhlRow rrow;
        rrow.pos = w0 + q0 + k; rrow.strand_ctx = (s + 1) | (code << 8); rrow.coverage = c; rrow.pad = 0;
        rrow.length = (double)(unsigned long long)L64[(s * 3 + 0) * S + q0 + k] / (double)c;                                            // (double)uint64 / int, :97
        rrow.lmhl = (double)(unsigned long long)L64[(s * 3 + 2) * S + q0 + k] / (double)(unsigned long long)L64[(s * 3 + 1) * S + q0 + k];  // :98
        *dst++ = rrow;
      }
  }
}

// PAIRS: `ctx` is a union of the letter pairs Hh / Xx / Zz (what generateMhlReport passes, R/generateMhlReport.R:178-180):
// the row is swept 16 bytes at a time with per-word indicator planes gathered into byte-order masks, and only the set bits
// (coverage transitions, calls of the haplotype context, stretch events) are visited.  Any other letter set takes the
// byte-serial sweep.
template <int CLS /* 0: any letter set (byte-serial sweep); else bit 0 Hh, bit 1 Xx, bit 2 Zz */>
__global__ void __launch_bounds__(kMhlThreads, 3)
k_window_mhl(Csr csr, WinTable wt, MhlParams prm, MhlOut out) {
  constexpr bool PAIRS = CLS != 0;
  constexpr int S = kMhlS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ long long s_scan[4 * kMhlWarps];
  __shared__ uint32_t s_scan32[kMhlWarps + 1];
  __shared__ long long s_base;
  // per strand: dH, dDe, dNu (int64) then dW (int32) and T[n_cls] (uint32)
  const int NCLS = prm.n_cls;
  long long* const L64 = reinterpret_cast<long long*>(smem_raw);               // [2][3][S]
  int* const dWb = reinterpret_cast<int*>(L64 + 2 * 3 * S);                    // [2][S]
  uint32_t* const Tb = reinterpret_cast<uint32_t*>(dWb + 2 * S);               // [2][NCLS][S]
  const int kZero4 = (2 * 3 * S * 8 + 2 * S * 4 + 2 * NCLS * S * 4) / 16;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < kZero4; i += kMhlThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];

    for (int r = lo + (int)threadIdx.x; r < hi; r += kMhlThreads) {
      const long long o0 = csr.off[r], o1 = csr.off[r + 1];
      const int len = csr.len[r];
      const int rel0 = csr.start[r] - w0;
      if (len <= 0 || rel0 + len <= 0) continue;  // rel0 < S by construction of hi
      const int s = csr.strand[r] - 1;
      const uint8_t* row = csr.data + o0;
      const int nch = (int)((o1 - o0) >> 4);
      // ---- 1. whole-row counts: h, out-of-context methylated / unmethylated ----
      uint32_t h = 0, om = 0, ou = 0;
      constexpr uint32_t sel2 = (CLS & 1) ? 0xFFFFFFFFu : 0u, sel6 = (CLS & 2) ? 0xFFFFFFFFu : 0u, sel7 = (CLS & 4) ? 0xFFFFFFFFu : 0u;
      if (PAIRS) {
        uint4 vn = nch > 0 ? __ldg(reinterpret_cast<const uint4*>(row)) : make_uint4(0, 0, 0, 0);
        for (int c = 0; c < nch; c++) {
          const uint4 v = vn;
          if (c + 1 < nch) vn = __ldg(reinterpret_cast<const uint4*>(row) + c + 1);  // next chunk in flight while this one is counted
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; k++) {  // bit 0 of every byte; pad bytes (code 3) are in none of the classes
            const uint32_t b0 = w[k], b1 = w[k] >> 1, b2 = w[k] >> 2, b3 = w[k] >> 3;
            const uint32_t in = (((~b0 & b1 & ~b2) & sel2) | ((~b0 & b1 & b2) & sel6) | ((b0 & b1 & b2) & sel7)) & 0x01010101u;
            const uint32_t ooc = ((b1 & ~b0) | (b0 & b2)) & 0x01010101u & ~in;  // low three bits 2, 6 | 5, 7: a call letter
            h += __popc(in);
            om += __popc(ooc & ~b3);
            ou += __popc(ooc & b3);
          }
        }
      } else {
      for (int c = 0; c < nch; c++) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(row) + c);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const uint32_t code = (w[k >> 2] >> (8 * (k & 3))) & 15u;  // pad bytes are code 3: in no set below unless ctx holds it (it cannot: 3 is no call letter)
          const uint32_t in = (prm.ctx_mask >> code) & 1u;
          h += in;
          om += (~in) & (0x00E4u >> code) & 1u;   // H U X Z: 2 5 6 7
          ou += (~in) & (0xE400u >> code) & 1u;   // h u x z: 10 13 14 15
        }
      }
      }
      if ((int)h < prm.hmin || (double)om / (double)(om + ou) > prm.max_ooctx) continue;
      const long long hv = (long long)h;
      const long long dv = (long long)mhl_S(h < prm.cap ? h : prm.cap);
      int* const dW = dWb + s * S;
      long long* const dH = L64 + (s * 3 + 0) * S;
      long long* const dDe = L64 + (s * 3 + 1) * S;
      long long* const dNu = L64 + (s * 3 + 2) * S;
      uint32_t* const T = Tb + s * NCLS * S;
      // ---- 2. byte-serial sweep ----
      bool cov = false;
      uint32_t msz = 0;
      int ms = 0, me = 0;
      // '-' bytes seen since the open stretch began: up to four 16-bit row positions packed into 64 bits (rows are at most
      // 65535 bytes); more than four, or the byte-serial sweep (which does not track them): walk the span byte by byte
      unsigned long long dlist = 0;
      int dcnt = PAIRS ? 0 : 5;
      auto close_stretch = [&]() {  // push +-S(min(size, hmax)) around the non-'-' pieces of [ms, me]
        const long long nv = (long long)mhl_S(msz < prm.cap ? msz : prm.cap);
        if (dcnt <= 4) {
          // the whole span, then every '-' byte strictly inside it switched off for one position (adjacent ones compose)
          if (rel0 + ms < S) atomicAdd(reinterpret_cast<unsigned long long*>(&dNu[max(rel0 + ms, 0)]), (unsigned long long)nv);
          if (rel0 + me + 1 < S) atomicAdd(reinterpret_cast<unsigned long long*>(&dNu[max(rel0 + me + 1, 0)]), (unsigned long long)(-nv));
          for (int j = 0; j < dcnt; j++) {
            const int p = (int)((dlist >> (16 * j)) & 0xFFFFull);
            if (p < me) {
              if (rel0 + p < S) atomicAdd(reinterpret_cast<unsigned long long*>(&dNu[max(rel0 + p, 0)]), (unsigned long long)(-nv));
              if (rel0 + p + 1 < S) atomicAdd(reinterpret_cast<unsigned long long*>(&dNu[max(rel0 + p + 1, 0)]), (unsigned long long)nv);
            }
          }
        } else {
          bool on = false;
          for (int i = ms; i <= me; i++) {
            const bool c2 = (__ldg(row + i) & 15u) != 11u;
            if (c2 != on) {
              const int q = rel0 + i;
              if (q < S) atomicAdd(reinterpret_cast<unsigned long long*>(&dNu[max(q, 0)]), (unsigned long long)(c2 ? nv : -nv));
              on = c2;
            }
          }
          if (on && rel0 + me + 1 < S) atomicAdd(reinterpret_cast<unsigned long long*>(&dNu[max(rel0 + me + 1, 0)]), (unsigned long long)(-nv));
        }
        msz = 0;
      };
      if (PAIRS) {
        uint32_t covb = 0;  // coverage state as a bit
        uint4 vn = __ldg(reinterpret_cast<const uint4*>(row));
        for (int c = 0; c < nch; c++) {
          const uint4 v = vn;
          if (c + 1 < nch) vn = __ldg(reinterpret_cast<const uint4*>(row) + c + 1);  // next chunk in flight while this one is swept
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          const int kmax = min(16, len - 16 * c);
          const uint32_t valid = kmax >= 16 ? 0xFFFFu : ((1u << kmax) - 1u);
          uint32_t dash16 = 0, meth16 = 0, t16[3] = {0, 0, 0};
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint32_t b0 = w[k], b1 = w[k] >> 1, b2 = w[k] >> 2, b3 = w[k] >> 3;
            dash16 |= gather4(b0 & b1 & ~b2 & b3 & 0x01010101u) << (4 * k);
            uint32_t in = 0;
            if (sel2) { const uint32_t c2 = ~b0 & b1 & ~b2 & 0x01010101u; t16[0] |= gather4(c2) << (4 * k); in |= c2; }
            if (sel6) { const uint32_t c6 = ~b0 & b1 & b2 & 0x01010101u; t16[1] |= gather4(c6) << (4 * k); in |= c6; }
            if (sel7) { const uint32_t c7 = b0 & b1 & b2 & 0x01010101u; t16[2] |= gather4(c7) << (4 * k); in |= c7; }
            meth16 |= gather4(in & ~b3) << (4 * k);
          }
          const uint32_t nd16 = ~dash16 & valid;
          const uint32_t call16 = (t16[0] | t16[1] | t16[2]) & valid;
          const int base = rel0 + 16 * c;
          for (uint32_t tr = (nd16 ^ ((nd16 << 1) | covb)) & valid; tr; tr &= tr - 1u) {  // coverage transitions
            const int k = __ffs(tr) - 1;
            const bool on = (nd16 >> k) & 1u;
            diff_event(base + k, S, dW, dH, dDe, on ? 1 : -1, on ? hv : -hv, on ? dv : -dv);
          }
          covb = (nd16 >> (kmax - 1)) & 1u;
#pragma unroll
          for (int cls = 0; cls < 3; cls++)  // calls of the haplotype context (the other classes cannot produce rows)
            for (uint32_t m = t16[cls] & valid; m; m &= m - 1u) {
              const int q = base + __ffs(m) - 1;
              if ((unsigned)q < (unsigned)S) atomicAdd(&T[prm.slot[cls] * S + q], 1u);
            }
          // methylated stretches: context calls and '-' bytes in position order
          for (uint32_t m = call16 | (msz || (call16 & meth16) ? dash16 & valid : 0u); m; m &= m - 1u) {
            const int k = __ffs(m) - 1;
            const int i = 16 * c + k;
            if (!((call16 >> k) & 1u)) {  // a '-' byte
              if (msz) { dlist = (dlist << 16) | (unsigned long long)i; dcnt++; }
              continue;
            }
            if ((meth16 >> k) & 1u) {
              if (!msz) { ms = i; dcnt = 0; }
              me = i; msz++;
            } else if (msz) close_stretch();
          }
        }
        cov = covb != 0u;
      } else {
      for (int c = 0; c < nch; c++) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(row) + c);
        const int kmax = min(16, len - 16 * c);
        for (int k = 0; k < kmax; k++) {
          const uint32_t word = k < 8 ? (k < 4 ? v.x : v.y) : (k < 12 ? v.z : v.w);
          const uint32_t code = (word >> (8 * (k & 3))) & 15u;
          const int i = 16 * c + k;
          const bool c2 = code != 11u;
          if (c2 != cov) {
            diff_event(rel0 + i, S, dW, dH, dDe, c2 ? 1 : -1, c2 ? hv : -hv, c2 ? dv : -dv);
            cov = c2;
          }
          const uint32_t c7 = code & 7u;
          if (c7 == 2u || c7 == 6u || c7 == 7u) {  // a cytosine call: count it at its position
            const int q = rel0 + i;
            const int sl = prm.slot[c7 == 2u ? 0 : (c7 == 6u ? 1 : 2)];
            if (sl >= 0 && (unsigned)q < (unsigned)S) atomicAdd(&T[sl * S + q], 1u);
          }
          if ((prm.ctx_mask >> code) & 1u) {
            if (code < 8u) { if (!msz) ms = i; me = i; msz++; }
            else if (msz) close_stretch();
          }
        }
      }
      }
      if (cov) diff_event(rel0 + len, S, dW, dH, dDe, -1, -hv, -dv);
      if (msz) close_stretch();
    }
    __syncthreads();

    mhl_window_finalize<long long>(smem_raw, prm, out, win, w0, s_scan, s_scan32, &s_base);
  }
}

struct MhlCols { int32_t *rname, *strand, *pos, *ctx, *coverage; double *length, *lmhl; };

__global__ void __launch_bounds__(128)
k_gather_mhl(const MhlRow* __restrict__ scratch, const long long* __restrict__ w_base, const long long* __restrict__ w_dst,
             const int32_t* __restrict__ w_cnt, const int32_t* __restrict__ w_rname, int n_win, long long cap, MhlCols out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int w = warp; w < n_win; w += n_warps) {
    const int cnt = w_cnt[w];
    if (!cnt) continue;
    const long long src = w_base[w], dst = w_dst[w];
    if (src + cnt > cap || dst + cnt > cap) continue;
    const int rn = w_rname[w];
    for (int i = lane; i < cnt; i += 32) {
      const MhlRow v = scratch[src + i];
      out.rname[dst + i] = rn;
      out.strand[dst + i] = v.strand_ctx & 0xFF;
      out.pos[dst + i] = v.pos;
      out.ctx[dst + i] = v.strand_ctx >> 8;
      out.coverage[dst + i] = v.coverage;
      out.length[dst + i] = v.length;
      out.lmhl[dst + i] = v.lmhl;
    }
  }
}

}  // namespace ea
