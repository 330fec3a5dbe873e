// Pieces shared by the short-read window kernels (window_cpl.cuh, window_tpr.cuh): shared-memory window state
// layout, event emission, integer threshold tables, finalize.
//
// Shared-memory window state, SG = kGuardLo + S + kGuardHi words per array (guard bands absorb the out-of-window
// bytes of chunks that straddle a window edge, so events need no bounds checks):
//   W       running-coverage DELTA of both strands: '+' in the low, '-' in the high 16 bits (+1 row start, -1 row
//           end, -1/+1 around long '-' runs; packed two's complement arithmetic is exact because every true prefix
//           sum is a coverage, 0..32767)
//   D       number of isolated '-' bytes at the position, '+' | '-' << 16;
//           cov(p) = prefix_sum(W)(p) - D(p), both strands in one subtraction
//   T[c][s] per requested context and strand: total calls | methylated calls of passing reads << 16
#pragma once
#include "kernels.cuh"

namespace ea {

constexpr int kGuardLo = 32;   // a unit (up to 32 bytes) may start this far before the window
constexpr int kGuardHi = 48;   // ... and end this far after it (+1 for the delta after a run)
constexpr int kThrTable = 1025;   // class counts 0..1024 (rows are at most 1024 bytes in these kernels)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// position (byte index within the chunk) of merged-indicator bit i
__device__ __forceinline__ int bit_pos(int i) { return ((i & 1) << 3) + (i & 4) + (i >> 3); }
// same for a 32-byte unit: bits 4t, 4t+1 belong to the first chunk, 4t+2, 4t+3 (masks shifted left by 2) to the second
__device__ __forceinline__ int bit_pos32(int i) { return ((i & 2) << 3) + bit_pos(i & ~2); }

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

// The same chain with the tables staged in SHARED memory (entries 0..limit): the two dependent table loads sit at the end
// of every row, and as global loads they stalled the row's warp for a full memory round trip - staging them took the
// fused row kernel from 3.34 to 2.62 ms per 50 M reads.
__device__ __forceinline__ void stage_thr_tables(const ThrTables& g, uint16_t* s_mmin, uint16_t* s_omax, int limit) {
  for (int i = threadIdx.x; i <= limit; i += blockDim.x) { s_mmin[i] = g.mmin[i]; s_omax[i] = g.omax[i]; }
}
__device__ __forceinline__ bool eval_pass_staged(uint32_t m, uint32_t all, uint32_t om, uint32_t oall, const ThrParams& t,
                                                 const uint16_t* s_mmin, const uint16_t* s_omax, uint32_t limit) {
  if (all > limit || oall > limit) return eval_pass(m, all, om, oall, t);
  return m != 0u && all >= t.min_n && m >= (uint32_t)s_mmin[all] && (om == 0u || om <= (uint32_t)s_omax[oall]);
}

__device__ __forceinline__ void red_shared(uint32_t addr, uint32_t v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// 4 * position (within a 32-byte unit) of merged-indicator bit i: the shared table (one LDS) or pure ALU (EA_POS_ALU)
__device__ __forceinline__ uint32_t pos4_of(uint32_t bp_tbl, int i);
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

// out-of-context calls of a chunk and their methylated subset (merged-indicator layout)
template <int TCTX>
__device__ __forceinline__ void out_masks(const Planes& p, uint32_t& o, uint32_t& om) {
  const uint32_t o0 = cls_out<TCTX>(p.A0, p.B0, p.C0), o1 = cls_out<TCTX>(p.A1, p.B1, p.C1);
  o = o0 | o1;
  om = (o0 & ~p.D0) | (o1 & ~p.D1);
}

// When every class of the threshold context is also reported, the in-context counts of the pass/fail chain are the
// popcounts of the event masks want_masks() produces anyway (the default: CG thresholds, CG report), and only the
// out-of-context classification is extra work.
template <int TCTX, int WANT>
__host__ __device__ constexpr bool in_from_want() {
  return (TCTX == TC_CG && (WANT & W_CG)) || (TCTX == TC_CHG && (WANT & W_CHG)) || (TCTX == TC_CHH && (WANT & W_CHH)) ||
         (TCTX == TC_CXG && (WANT & W_CG) && (WANT & W_CHG)) || (TCTX == TC_CX && WANT == 7);
}
// packed byte-wide counts like count4(): in-context total | methylated << 8 | out-of-context total << 16 | methylated << 24;
// m / meth: want_masks() of the chunk or unit, oc / ocm: out_masks() (merged the same way)
template <int TCTX, int WANT>
__device__ __forceinline__ uint32_t count_from_masks(const uint32_t (&m)[3], const uint32_t (&meth)[3], uint32_t oc, uint32_t ocm) {
  constexpr int s_chg = (WANT & W_CHH) ? 1 : 0, s_cg = ((WANT & W_CHH) ? 1 : 0) + ((WANT & W_CHG) ? 1 : 0);
  uint32_t in, inm;
  if (TCTX == TC_CG) { in = m[s_cg]; inm = meth[s_cg]; }
  else if (TCTX == TC_CHG) { in = m[s_chg]; inm = meth[s_chg]; }
  else if (TCTX == TC_CHH) { in = m[0]; inm = meth[0]; }
  else if (TCTX == TC_CXG) { in = m[s_chg] | m[s_cg]; inm = meth[s_chg] | meth[s_cg]; }
  else { in = m[0] | m[1] | m[2]; inm = meth[0] | meth[1] | meth[2]; }
  uint32_t c = __popc(in) + (__popc(inm) << 8);
  if (TCTX != TC_CX) c += (__popc(oc) << 16) + (__popc(ocm) << 24);
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

__device__ __forceinline__ uint32_t pos4_of(uint32_t bp_tbl, int i) {
#ifdef EA_POS_ALU
  const uint32_t u = (uint32_t)i;
  return ((u & 3u) << 5) | ((u & 4u) << 2) | ((u >> 1) & 0xCu);
#else
  return lds_u8(bp_tbl + i);
#endif
}

// '-' bytes (code 11) of one 16-byte chunk.  Isolated ones (at most two per chunk) are returned as a merged-layout mask
// and counted in D by the caller; longer runs (mate gaps) become coverage deltas in W right here: -1 where a run starts,
// +1 right after it ends, clamped to the window start.  w_addr = shared BYTE address of the chunk's byte 0 in W,
// sinc = 1 << 16*strand, bp_tbl = table of 4*position of merged-indicator bit i.
__device__ __forceinline__ uint32_t dash_events(const Planes& p, uint32_t w_addr, uint32_t sinc, uint32_t w_floor, uint32_t bp_tbl) {
  const uint32_t e0 = p.A0 & p.B0 & ~p.C0, e1 = p.A1 & p.B1 & ~p.C1;
  uint32_t E = (e0 & p.D0) | (e1 & p.D1);
  if (__popc(E) > 2) {
    const uint32_t prev = (E << 8) | ((E >> 20) & 0x30u) | ((E >> 27) & 0x2u);  // "byte before me is 11"
    const uint32_t starts = E & ~prev;
    for (uint32_t mm = starts | (prev & ~E); mm;) {
      const int i = 31 - __clz(mm);
      const uint32_t bit = 1u << i;
      mm ^= bit;
      red_shared(max(w_addr + pos4_of(bp_tbl, i), w_floor), (starts & bit) ? 0u - sinc : sinc);
    }
    if ((E >> 29) & 1u) red_shared(max(w_addr + 64u, w_floor), sinc);  // the run reaches the chunk end
    E = 0u;
  }
  return E;
}

// Shared-memory events of one in-window unit (one chunk, or two chunks with the second one's masks shifted left by 2):
// m[k] = calls of requested context k, E = isolated '-' bytes, methpass = methylated calls to be counted now (0 in the
// fused kernels).  d_addr / t_addr = shared BYTE addresses of the unit's byte 0 in D and in T[0][strand].
template <int NC, int SG>
__device__ __forceinline__ void unit_events(const uint32_t (&m)[3], uint32_t E, uint32_t methpass, uint32_t d_addr, uint32_t t_addr,
                                            uint32_t sinc, uint32_t bp_tbl) {
  uint32_t ev = m[0] | m[1] | m[2] | E;
  while (ev) {
    const int i = 31 - __clz(ev);
    const uint32_t bit = 1u << i;
    ev ^= bit;
    const uint32_t bp4 = pos4_of(bp_tbl, i);
    uint32_t a = t_addr;
    uint32_t v = (methpass & bit) ? 0x10001u : 1u;
    if (NC > 1 && (m[1] & bit)) a = t_addr + 2 * SG * 4;
    if (NC > 2 && (m[2] & bit)) a = t_addr + 4 * SG * 4;
    if (E & bit) { a = d_addr; v = sinc; }
    red_shared(a + bp4, v);
  }
}

template <int NC, int SG>
__device__ __forceinline__ void chunk_events(const Planes& p, const uint32_t (&m)[3], uint32_t methpass, uint32_t w_addr, uint32_t d_addr,
                                             uint32_t t_addr, uint32_t sinc, uint32_t w_floor, uint32_t bp_tbl) {
  const uint32_t E = dash_events(p, w_addr, sinc, w_floor, bp_tbl);
  unit_events<NC, SG>(m, E, methpass, d_addr, t_addr, sinc, bp_tbl);
}

// methylated calls of a passing row's chunk (fused kernel, after the flag is known)
__device__ __forceinline__ void push_meth(uint32_t mm, uint32_t t_addr, uint32_t bp_tbl) {
  while (mm) {
    const int i = 31 - __clz(mm);
    mm ^= 1u << i;
    red_shared(t_addr + pos4_of(bp_tbl, i), 0x10000u);
  }
}


// Finalize one window: coverage prefix sum, row test total_ctx > floor(cov/2) (src/rcpp_cx_report.cpp:58-85), ordered
// compaction into the scratch rows.  All NT threads of the CTA call it after a __syncthreads().
template <int WANT, int PER, int NT, bool FUSED>
__device__ __forceinline__ void window_finalize_rt(uint32_t* __restrict__ W, uint32_t* __restrict__ D, uint32_t* __restrict__ T,
                                                   int S /* positions, multiple of 4, <= PER * NT */, int w0, int win, const ReportOut& out,
                                                   uint32_t* s_scan /* 2 * NT/32 words */, long long* s_base, int* s_npass) {
  // Finalizes the window AND leaves its state zeroed for the next one (the kernels zero only once, before their window
  // loop).  Three CTA barriers: per-warp coverage totals, per-warp row counts, the window's base in the scratch rows.
  constexpr int NC = n_ctx<WANT>();
  constexpr int NW = NT / 32;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int SG = kGuardLo + S + kGuardHi;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q0 = threadIdx.x * PER;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  uint32_t cov[PER];  // coverage of both strands: '+' | '-' << 16
  uint32_t run = 0;
#pragma unroll
  for (int v4 = 0; v4 < PER / 4; v4++) {
    uint4 w = zero4, d = zero4;
    if (q0 + v4 * 4 < S) {
      uint4* pw = reinterpret_cast<uint4*>(&W[kGuardLo + q0 + v4 * 4]);
      uint4* pd = reinterpret_cast<uint4*>(&D[kGuardLo + q0 + v4 * 4]);
      w = *pw; d = *pd;
      *pw = zero4; *pd = zero4;
    }
    run += w.x; cov[v4 * 4 + 0] = run - d.x;
    run += w.y; cov[v4 * 4 + 1] = run - d.y;
    run += w.z; cov[v4 * 4 + 2] = run - d.z;
    run += w.w; cov[v4 * 4 + 3] = run - d.w;
  }
  uint32_t inc = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(FULL, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_scan[warp] = inc;
  __syncthreads();
  uint32_t ex = inc - run;
#pragma unroll
  for (int w = 0; w < NW; w++) ex += (w < warp) ? s_scan[w] : 0u;
#pragma unroll
  for (int k = 0; k < PER; k++) cov[k] += ex;

  uint32_t flags = 0;  // bit 2k+s: (q0+k, strand s) is a report row
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
      for (int v4 = 0; v4 < PER / 4; v4++) {
        uint4 v = zero4;
        if (q0 + v4 * 4 < S) v = *reinterpret_cast<const uint4*>(&T[(c * 2 + s) * SG + kGuardLo + q0 + v4 * 4]);
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
  const int mine = __popc(flags);
  int rinc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(FULL, rinc, d);
    if (lane >= d) rinc += o;
  }
  if (lane == 31) s_scan[NW + warp] = (uint32_t)rinc;
  __syncthreads();
  int my_off = rinc - mine, nrows_total = 0;
#pragma unroll
  for (int w = 0; w < NW; w++) {
    const int c = (int)s_scan[NW + w];
    nrows_total += c;
    my_off += (w < warp) ? c : 0;
  }
  if (threadIdx.x == 0) {
    const unsigned long long bb = nrows_total ? atomicAdd(out.cursor, (unsigned long long)nrows_total) : 0ull;
    *s_base = (long long)bb;
    out.w_base[win] = (long long)bb;
    out.w_cnt[win] = nrows_total;
    if (FUSED) {
      if (*s_npass) atomicAdd(out.n_pass, (unsigned long long)*s_npass);
      *s_npass = 0;
    }
  }
  __syncthreads();
  const long long wbase = *s_base;
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
  // leave T zeroed (the caller's barrier at the top of its window loop orders this before the next sweep)
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
      for (int v4 = 0; v4 < PER / 4; v4++)
        if (q0 + v4 * 4 < S) *reinterpret_cast<uint4*>(&T[(c * 2 + s) * SG + kGuardLo + q0 + v4 * 4]) = zero4;
}

// compile-time window size: PER = S / NT positions per thread
template <int WANT, int S, int NT, bool FUSED>
__device__ __forceinline__ void window_finalize(uint32_t* __restrict__ W, uint32_t* __restrict__ D, uint32_t* __restrict__ T,
                                                int w0, int win, const ReportOut& out, uint32_t* s_scan, long long* s_base, int* s_npass) {
  window_finalize_rt<WANT, S / NT, NT, FUSED>(W, D, T, S, w0, win, out, s_scan, s_base, s_npass);
}

}  // namespace ea
