// Short-read window kernel, "chunk per lane" formulation (the default for reads up to 1 KB).
//
// One CTA per genomic window of S positions; each of its 8 warps takes tiles of up to 32 consecutive
// rows (one contiguous, 16-byte aligned byte range of the padded CSR).  A tile goes through three
// phases that alternate between the two natural decompositions of the work:
//
//   rows    one LANE PER ROW: offsets/start/strand are read coalesced, the row's coverage start/end
//           deltas are pushed, and a 32-bit descriptor (shared-memory index of the chunk's first
//           position, or 0 when the chunk lies outside the window) is written for every 16-byte
//           chunk of the row;
//   chunks  one LANE PER CHUNK: coalesced LDG.128 straight from HBM (several loads in flight per
//           lane, no staging buffer), nibble-plane classification, per-chunk class counts for the
//           thresholds, and everything that does not depend on the pass flag - {total calls} of the
//           requested contexts and '-' bytes - as shared-memory reductions.  Count word and the
//           methylated-call masks of the chunk replace its descriptor in shared memory;
//   rows    one LANE PER ROW again: the row's chunk counts are summed, the reference's pass/fail
//           chain is evaluated for 32 rows at once (src/rcpp_threshold_reads.cpp:39-70; through exact
//           integer tables built from the same IEEE double divisions, see k_thr_tables), flags are
//           stored, and the few passing rows push their methylated calls (SURVEY.md 8a invariant 4:
//           only the methylated count depends on `pass`).
//
// No lane ever walks a row serially over data bytes, every global load is coalesced, and the only
// warp-divergent code is the event loop over the handful of calls in a chunk.
// The finalize (coverage prefix sum, row test total_ctx > floor(cov/2), ordered compaction;
// src/rcpp_cx_report.cpp:58-85) touches shared memory only.
//
// Shared-memory window state, SG = 16 + S + 32 words per array (guard bands absorb the out-of-window
// bytes of chunks that straddle a window edge, so events need no bounds checks):
//   W       running-coverage DELTA of both strands: '+' in the low, '-' in the high 16 bits (+1 row
//           start, -1 row end, -1/+1 around long '-' runs; packed two's complement arithmetic is
//           exact because every true prefix sum is a coverage, 0..32767)
//   D       number of isolated '-' bytes at the position, '+' | '-' << 16;
//           cov(p) = prefix_sum(W)(p) - D(p), both strands in one subtraction
//   T[c][s] per requested context and strand: total calls | methylated calls of passing reads << 16
// Preconditions (checked by the dispatcher): padded rows <= 1024 bytes, rpt * max chunks per row <= tc,
// at most 32767 rows overlap a window.
#pragma once
#include "kernels.cuh"
#include "window_tpr.cuh"  // block_exscan_n, bit_pos, smem_u32

namespace ea {

constexpr int kCplThreads = 256;
constexpr int kCplWarps = kCplThreads / 32;
constexpr int kGuardLo = 16;
constexpr int kGuardHi = 32;
constexpr int kCplBatch = 5;      // chunk loads in flight per lane
constexpr int kThrTable = 1025;   // class counts 0..1024 (rows are at most 1024 bytes here)

// Integer form of the two IEEE-double comparisons of the pass/fail chain (src/rcpp_threshold_reads.cpp:52-53,66-67).
// For a fixed denominator the rounded quotient is monotone in the numerator, so each comparison is a threshold on the
// numerator; the thresholds are found with the very same double divisions:
//   mmin[all]  smallest m >= 1 with !((double)m/all < min_beta)    (0xFFFF: none)
//   omax[oall] largest om >= 1 with !((double)om/oall > max_oo)    (0: none)
struct ThrTables { const uint16_t* mmin; const uint16_t* omax; };

__global__ void k_thr_tables(ThrParams t, uint16_t* __restrict__ mmin, uint16_t* __restrict__ omax) {
  const int all = blockIdx.x * blockDim.x + threadIdx.x;
  if (all >= kThrTable) return;
  // both predicates are monotone in m (the rounded quotient is), so bisect
  uint32_t lo = 0xFFFFu, hi = 0u;
  if (all >= 1) {
    int a = 1, b = all + 1;  // first m in [1, all] with !(m/all < min_beta), all+1 if none
    while (a < b) {
      const int m = (a + b) >> 1;
      if ((double)m / (double)all < t.min_beta) a = m + 1; else b = m;
    }
    if (a <= all) lo = (uint32_t)a;
    a = 0; b = all;          // last m in [1, all] with !(m/all > max_oo), 0 if none
    while (a < b) {
      const int m = (a + b + 1) >> 1;
      if ((double)m / (double)all > t.max_oo) b = m - 1; else a = m;
    }
    hi = (uint32_t)a;
  }
  mmin[all] = (uint16_t)lo;
  omax[all] = (uint16_t)hi;
}

__device__ __forceinline__ bool eval_pass_tables(uint32_t m, uint32_t all, uint32_t om, uint32_t oall, const ThrParams& t, const ThrTables& tb) {
  if (all >= (uint32_t)kThrTable || oall >= (uint32_t)kThrTable) return eval_pass(m, all, om, oall, t);
  return m != 0u && all >= t.min_n && m >= (uint32_t)__ldg(tb.mmin + all) && (om == 0u || om <= (uint32_t)__ldg(tb.omax + oall));
}

__device__ __forceinline__ void red_shared(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// exclusive scan of one word per thread over the CTA (wrapping unsigned arithmetic); *total = CTA sum
template <int NWARPS>
__device__ __forceinline__ uint32_t block_exscan_u32(uint32_t v, uint32_t* total, uint32_t* s_warp /* NWARPS+1 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t x = lane < NWARPS ? s_warp[lane] : 0u;
    uint32_t xi = x;
#pragma unroll
    for (int d = 1; d < NWARPS; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, xi, d);
      if (lane >= d) xi += o;
    }
    if (lane < NWARPS) s_warp[lane] = xi - x;
    if (lane == NWARPS - 1) s_warp[NWARPS] = xi;
  }
  __syncthreads();
  *total = s_warp[NWARPS];
  return s_warp[warp] + inc - v;
}

// class counts of one chunk, one byte each: in-context calls | methylated << 8 | out-of-context << 16 | methylated << 24
template <int TCTX>
__device__ __forceinline__ uint32_t count4(const Planes& p) {
  const uint32_t t0 = cls_in<TCTX>(p.A0, p.B0, p.C0), t1 = cls_in<TCTX>(p.A1, p.B1, p.C1);
  uint32_t c = __popc(t0 | t1) + (__popc((t0 & ~p.D0) | (t1 & ~p.D1)) << 8);
  if (TCTX != TC_CX) {
    const uint32_t o0 = cls_out<TCTX>(p.A0, p.B0, p.C0), o1 = cls_out<TCTX>(p.A1, p.B1, p.C1);
    c += (__popc(o0 | o1) << 16) + (__popc((o0 & ~p.D0) | (o1 & ~p.D1)) << 24);
  }
  return c;
}

// calls of the requested contexts (slot order CHH, CHG, CG among the requested) and their methylated subsets
template <int WANT>
__device__ __forceinline__ void want_masks(const Planes& p, uint32_t (&m)[3], uint32_t (&meth)[3]) {
  int slot = 0;
  m[0] = m[1] = m[2] = 0u;
  meth[0] = meth[1] = meth[2] = 0u;
  if (WANT & W_CHH) { const uint32_t a = cls_chh(p.A0, p.B0, p.C0), b = cls_chh(p.A1, p.B1, p.C1); m[slot] = a | b; meth[slot] = (a & ~p.D0) | (b & ~p.D1); slot++; }
  if (WANT & W_CHG) { const uint32_t a = cls_chg(p.A0, p.B0, p.C0), b = cls_chg(p.A1, p.B1, p.C1); m[slot] = a | b; meth[slot] = (a & ~p.D0) | (b & ~p.D1); slot++; }
  if (WANT & W_CG) { const uint32_t a = cls_cg(p.A0, p.B0, p.C0), b = cls_cg(p.A1, p.B1, p.C1); m[slot] = a | b; meth[slot] = (a & ~p.D0) | (b & ~p.D1); slot++; }
}

// Shared-memory events of one in-window chunk.  All addresses are 32-bit shared-space BYTE addresses of the chunk's
// byte 0: w_addr in W, d_addr in D, t_addr in T[0][strand]; sinc = 1 << 16*strand.  methpass = methylated calls to be
// counted now (0 in the fused kernel); bp_tbl = table of 4*bit_pos(i).
template <int NC, int SG>
__device__ __forceinline__ void chunk_events(const Planes& p, const uint32_t (&m)[3], uint32_t methpass, uint32_t w_addr, uint32_t d_addr,
                                             uint32_t t_addr, uint32_t sinc, uint32_t w_floor, uint32_t bp_tbl) {
  const uint32_t e0 = p.A0 & p.B0 & ~p.C0, e1 = p.A1 & p.B1 & ~p.C1;
  uint32_t E = (e0 & p.D0) | (e1 & p.D1);  // code 11 bytes
  if (__popc(E) > 2) {  // runs (mate gaps): -1 where a run starts, +1 right after it ends; clamped to the window start
    const uint32_t prev = (E << 8) | ((E >> 20) & 0x30u) | ((E >> 27) & 0x2u);  // "byte before me is 11"
    const uint32_t starts = E & ~prev;
    for (uint32_t mm = starts | (prev & ~E); mm;) {
      const int i = 31 - __clz(mm);
      const uint32_t bit = 1u << i;
      mm ^= bit;
      red_shared(max(w_addr + lds_u8(bp_tbl + i), w_floor), (starts & bit) ? 0u - sinc : sinc);
    }
    if ((E >> 29) & 1u) red_shared(max(w_addr + 64u, w_floor), sinc);  // the run reaches the chunk end
    E = 0u;
  }
  uint32_t ev = m[0] | m[1] | m[2] | E;
  while (ev) {
    const int i = 31 - __clz(ev);
    const uint32_t bit = 1u << i;
    ev ^= bit;
    const uint32_t bp4 = lds_u8(bp_tbl + i);
    uint32_t a = t_addr;
    uint32_t v = (methpass & bit) ? 0x10001u : 1u;
    if (NC > 1 && (m[1] & bit)) a = t_addr + 2 * SG * 4;
    if (NC > 2 && (m[2] & bit)) a = t_addr + 4 * SG * 4;
    if (E & bit) { a = d_addr; v = sinc; }
    red_shared(a + bp4, v);
  }
}

// methylated calls of a passing row's chunk (fused kernel, after the flag is known)
__device__ __forceinline__ void push_meth(uint32_t mm, uint32_t t_addr, uint32_t bp_tbl) {
  while (mm) {
    const int i = 31 - __clz(mm);
    mm ^= 1u << i;
    red_shared(t_addr + lds_u8(bp_tbl + i), 0x10000u);
  }
}

template <int TCTX, int WANT, bool FUSED, int S>
__global__ void __launch_bounds__(kCplThreads, 4)
k_window_cpl(Csr csr, WinTable wt, ThrParams thr, ThrTables tbl, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
             int rpt, int tc, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr int SG = kGuardLo + S + kGuardHi;
  constexpr int MW = FUSED ? (NC == 3 ? 2 : 1) : 0;  // mask words per chunk
  constexpr int PER = S / kCplThreads;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[kCplWarps + 1];
  __shared__ long long s_base;
  __shared__ int s_npass;
  __shared__ __align__(4) uint8_t s_bp4[32];

  uint32_t* const W = reinterpret_cast<uint32_t*>(smem_raw);  // [SG]
  uint32_t* const D = W + SG;                                 // [SG]
  uint32_t* const T = D + SG;                                 // [NC][2][SG]
  constexpr int kZero4 = (NC + 1) * 2 * SG / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* const aux_d = T + NC * 2 * SG + warp * (1 + MW) * tc;  // descriptor, then count word
  uint32_t* const aux_m = aux_d + tc;                              // methylated-call masks (slots 0,1)
  uint32_t* const aux_m2 = aux_m + tc;                             // slot 2 (CX reports)
  const uint32_t w_base = smem_u32(W), d_base = smem_u32(D), t_base = smem_u32(T), bp_tbl = smem_u32(s_bp4);
  const uint32_t w_floor = w_base + 4u * kGuardLo;

  if (threadIdx.x < 32) s_bp4[threadIdx.x] = (uint8_t)(4 * bit_pos(threadIdx.x));

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < kZero4; i += kCplThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) s_npass = 0;
    __syncthreads();

    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    int npass_w = 0;

    for (int base = lo + warp * rpt; base < hi; base += kCplWarps * rpt) {
      // ------------------------------ rows: metadata, coverage deltas, chunk descriptors ------------------------------
      const int tend = min(base + rpt, hi);
      const int r = base + lane;
      const bool valid = r < tend;
      const long long o0 = csr.off[valid ? r : tend];
      long long o1 = __shfl_down_sync(FULL, o0, 1);
      if (lane == 31) o1 = csr.off[tend];
      const long long tile_o0 = __shfl_sync(FULL, o0, 0);
      const int n_chunks = (int)((__shfl_sync(FULL, o1, 31) - tile_o0) >> 4);
      int start = 0, spv = 0;
      if (valid) { start = csr.start[r]; spv = csr.sp[r]; }
      const int nch = (int)((o1 - o0) >> 4);
      const int cbeg = (int)((o0 - tile_o0) >> 4);
      const int strand0 = spv & 1;
      const int len = nch * 16 - (spv >> 4);
      const int rel0 = start - w0;
      const bool overlaps = valid && len > 0 && rel0 + len > 0;  // rel0 < S by construction of hi
      {
        // descriptor of chunk c: T-relative word index of its byte 0 (strand0*SG + rel + kGuardLo) when the chunk
        // overlaps the window, else 0; bit 31 = pass flag in the two-pass kernel
        int rel16 = rel0 + kGuardLo;
        uint32_t flag = 0u;
        if (overlaps) {
          const uint32_t sinc = strand0 ? 0x10000u : 1u;
          red_shared(w_floor + 4u * (uint32_t)max(rel0, 0), sinc);
          const int e = rel0 + len;
          if (e < S) red_shared(w_floor + 4u * (uint32_t)e, 0u - sinc);
          if (!FUSED && (pass_in == nullptr || pass_in[r] != 0)) flag = 0x80000000u;
        } else {
          rel16 = -0x40000000;  // never in the window
        }
        uint32_t* dst = aux_d + cbeg;
        const uint32_t sbase = (uint32_t)(strand0 * SG) | flag;
        for (int c = 0; c < nch; c++) {
          dst[c] = ((unsigned)(rel16 - 1) < (unsigned)(S + 15)) ? sbase + (uint32_t)rel16 : 0u;
          rel16 += 16;
        }
      }
      __syncwarp();

      // ------------------------------ chunks: classification, counts, pass-independent events ------------------------------
      const uint8_t* ptr = csr.data + tile_o0 + (size_t)lane * 16;
      for (int j0 = lane; j0 < n_chunks; j0 += 32 * kCplBatch, ptr += 512 * kCplBatch) {
        uint4 v[kCplBatch];
#pragma unroll
        for (int u = 0; u < kCplBatch; u++)
          if (j0 + 32 * u < n_chunks) v[u] = ldg_chunk(ptr + 512 * u);
#pragma unroll
        for (int u = 0; u < kCplBatch; u++) {
          const int j = j0 + 32 * u;
          if (j >= n_chunks) break;
          const uint32_t d = aux_d[j];
          Planes p;
          make_planes(v[u], p);
          uint32_t m[3], meth[3];
          want_masks<WANT>(p, m, meth);
          const uint32_t ix = d & 0x7FFFFFFFu;  // strand*SG + rel16, 0 = not in the window
          if (ix) {
            const uint32_t s = ix >= (uint32_t)SG ? 1u : 0u;
            const uint32_t rel4 = 4u * (ix - s * (uint32_t)SG);
            const uint32_t methpass = FUSED ? 0u : ((d >> 31) ? (meth[0] | meth[1] | meth[2]) : 0u);
            chunk_events<NC, SG>(p, m, methpass, w_base + rel4, d_base + rel4, t_base + 4u * ix, s ? 0x10000u : 1u, w_floor, bp_tbl);
          }
          if (FUSED) {
            aux_d[j] = count4<TCTX>(p);
            aux_m[j] = ix ? (meth[0] | (meth[1] << 2)) : 0u;
            if (NC == 3) aux_m2[j] = ix ? meth[2] : 0u;
          }
        }
      }

      // ------------------------------ rows: class counts -> pass flag -> methylated calls of passing rows ------------------------------
      if (FUSED) {
        __syncwarp();
        if (overlaps) {
          uint32_t part = 0, acc_in = 0, acc_out = 0;
          const uint32_t* src = aux_d + cbeg;
          for (int c = 0; c < nch; c++) {
            part += src[c];
            if ((c & 7) == 7) {
              acc_in += (part & 0xFFu) | ((part & 0xFF00u) << 8);
              acc_out += ((part >> 16) & 0xFFu) | ((part >> 24) << 16);
              part = 0;
            }
          }
          acc_in += (part & 0xFFu) | ((part & 0xFF00u) << 8);
          acc_out += ((part >> 16) & 0xFFu) | ((part >> 24) << 16);
          const bool pass = eval_pass_tables(acc_in >> 16, acc_in & 0xFFFFu, acc_out >> 16, acc_out & 0xFFFFu, thr, tbl);
          if (rel0 >= 0) {  // the window that holds the row's start owns its flag
            pass_out[r] = pass ? 1 : 0;
            npass_w += pass;
          }
          if (pass) {
            const uint32_t ta = t_base + 4u * (uint32_t)(strand0 * SG + rel0 + kGuardLo);
            for (int c = 0; c < nch; c++) {
              const uint32_t mm = aux_m[cbeg + c];
              if (mm) {
                push_meth(mm & 0x33333333u, ta + 64u * c, bp_tbl);
                if (NC > 1) push_meth((mm >> 2) & 0x33333333u, ta + 64u * c + 2 * SG * 4, bp_tbl);
              }
              if (NC > 2) {
                const uint32_t m2 = aux_m2[cbeg + c];
                if (m2) push_meth(m2, ta + 64u * c + 4 * SG * 4, bp_tbl);
              }
            }
          }
        }
      }
      __syncwarp();  // the next tile's descriptors overwrite aux
    }
    if (FUSED) {
      for (int d = 16; d >= 1; d >>= 1) npass_w += __shfl_down_sync(FULL, npass_w, d);
      if (lane == 0 && npass_w) atomicAdd(&s_npass, npass_w);
    }
    __syncthreads();

    // ------------------------------ finalize: coverage prefix sum, row detection, ordered compaction ------------------------------
    const int q0 = threadIdx.x * PER;
    uint32_t cov[PER];  // coverage of both strands: '+' | '-' << 16
    {
      uint32_t run = 0;
#pragma unroll
      for (int v4 = 0; v4 < PER / 4; v4++) {
        const uint4 w = *reinterpret_cast<const uint4*>(&W[kGuardLo + q0 + v4 * 4]);
        const uint4 d = *reinterpret_cast<const uint4*>(&D[kGuardLo + q0 + v4 * 4]);
        run += w.x; cov[v4 * 4 + 0] = run - d.x;
        run += w.y; cov[v4 * 4 + 1] = run - d.y;
        run += w.z; cov[v4 * 4 + 2] = run - d.z;
        run += w.w; cov[v4 * 4 + 3] = run - d.w;
      }
      uint32_t tot;
      const uint32_t ex = block_exscan_u32<kCplWarps>(run, &tot, s_scan);
#pragma unroll
      for (int k = 0; k < PER; k++) cov[k] += ex;
    }
    uint32_t flags = 0;  // bit 2k+s: (q0+k, strand s) is a report row
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
      for (int c = 0; c < NC; c++)
#pragma unroll
        for (int v4 = 0; v4 < PER / 4; v4++) {
          const uint4 v = *reinterpret_cast<const uint4*>(&T[(c * 2 + s) * SG + kGuardLo + q0 + v4 * 4]);
          if (v.x | v.y | v.z | v.w) {
            const uint32_t t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const uint32_t cv = cov[v4 * 4 + j];
              const uint32_t half = s ? (cv >> 17) : ((cv & 0xFFFFu) >> 1);
              if ((t[j] & 0xFFFFu) > half) flags |= 1u << (2 * (v4 * 4 + j) + s);
            }
          }
        }
    uint32_t nrows_u;
    const int my_off = (int)block_exscan_u32<kCplWarps>((uint32_t)__popc(flags), &nrows_u, s_scan);
    const int nrows_total = (int)nrows_u;
    if (threadIdx.x == 0) {
      const unsigned long long bb = nrows_total ? atomicAdd(out.cursor, (unsigned long long)nrows_total) : 0ull;
      s_base = (long long)bb;
      out.w_base[win] = (long long)bb;
      out.w_cnt[win] = nrows_total;
      if (FUSED && s_npass) atomicAdd(out.n_pass, (unsigned long long)s_npass);
    }
    __syncthreads();
    const long long wbase = s_base;
    if (flags && wbase + nrows_total <= out.cap) {
      int4* dst = out.scratch + wbase + my_off;
      while (flags) {  // ascending (position, strand) order
        const int b = __ffs(flags) - 1;
        flags &= flags - 1u;
        const int k = b >> 1, s = b & 1;
        int code = (WANT & W_CHH) ? 2 : ((WANT & W_CHG) ? 6 : 7);
        uint32_t val = T[s * SG + kGuardLo + q0 + k];
        if (NC > 1) {  // which of the requested contexts holds the majority
          uint32_t cv = 0;
#pragma unroll
          for (int kk = 0; kk < PER; kk++) cv = (kk == k) ? cov[kk] : cv;
          const uint32_t half = s ? (cv >> 17) : ((cv & 0xFFFFu) >> 1);
          int slot = 0;
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            if (!((WANT >> cc) & 1)) continue;
            const uint32_t v = T[(slot * 2 + s) * SG + kGuardLo + q0 + k];
            if ((v & 0xFFFFu) > half) { code = cc == 0 ? 2 : (cc == 1 ? 6 : 7); val = v; }
            slot++;
          }
        }
        *dst++ = make_int4(w0 + q0 + k, (s + 1) | (code << 8), (int)(val >> 16), (int)((val & 0xFFFFu) - (val >> 16)));
      }
    }
  }
}

}  // namespace ea
