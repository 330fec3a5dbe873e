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
// When every padded row is a multiple of 32 bytes (always true for 150 bp reads) the unit a lane handles is a PAIR
// of chunks (UB = 32: one 256-bit load, the second chunk's indicator masks shifted left by 2 so that one event loop,
// one count word and one descriptor serve 32 positions); otherwise UB = 16.
// No lane ever walks a row serially over data bytes, every global load is coalesced, and the only
// warp-divergent code is the event loop over the handful of calls in a chunk.
// The finalize (coverage prefix sum, row test total_ctx > floor(cov/2), ordered compaction;
// src/rcpp_cx_report.cpp:58-85) touches shared memory only.
//
// Window state layout, events and finalize: win_common.cuh.
// Preconditions (checked by the dispatcher): padded rows <= 1024 bytes, rpt * max chunks per row <= tc,
// at most 32767 rows overlap a window.
#pragma once
#include "win_common.cuh"

namespace ea {

constexpr int kCplThreads = 256;
constexpr int kCplWarps = kCplThreads / 32;
// tuning knobs (overridable for A/B builds)
#ifndef EA_CPL_MINB32
#define EA_CPL_MINB32 5   // CTAs per SM the 32-byte-unit variant is compiled for (measured: 5 x 8 warps beats 4 and 3)
#endif
#ifndef EA_CPL_MINB16
#define EA_CPL_MINB16 4
#endif
#ifndef EA_CPL_BATCH32
#define EA_CPL_BATCH32 2
#endif
#ifndef EA_CPL_BATCH16
#define EA_CPL_BATCH16 5
#endif

__device__ __forceinline__ void ldg_unit(const uint8_t* p, uint4& a) { a = ldg_chunk(p); }
__device__ __forceinline__ void ldg_unit(const uint8_t* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}

template <int TCTX, int WANT, bool FUSED, int S, int UB>
__global__ void __launch_bounds__(kCplThreads, UB == 32 ? EA_CPL_MINB32 : EA_CPL_MINB16)
k_window_cpl(Csr csr, WinTable wt, ThrParams thr, ThrTables tbl, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
             int rpt, int tc, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr int SG = kGuardLo + S + kGuardHi;
  constexpr int MW = FUSED ? NC : 0;              // mask words per unit
  constexpr int USH = UB == 32 ? 5 : 4;           // log2(unit bytes)
  constexpr int BATCH = UB == 32 ? EA_CPL_BATCH32 : EA_CPL_BATCH16;  // unit loads in flight per lane
  constexpr int FLUSH = UB == 32 ? 4 : 8;         // units whose byte-wide class counts can be summed without overflow
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[2 * kCplWarps];
  __shared__ long long s_base;
  __shared__ int s_npass;
  __shared__ __align__(4) uint8_t s_bp4[32];

  uint32_t* const W = reinterpret_cast<uint32_t*>(smem_raw);  // [SG]
  uint32_t* const D = W + SG;                                 // [SG]
  uint32_t* const T = D + SG;                                 // [NC][2][SG]
  constexpr int kZero4 = (NC + 1) * 2 * SG / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* const aux_d = T + NC * 2 * SG + warp * (1 + MW) * tc;  // per unit: descriptor, then count word
  uint32_t* const aux_m = aux_d + tc;                              // [MW][tc] methylated-call masks per requested context
  const uint32_t w_base = smem_u32(W), d_base = smem_u32(D), t_base = smem_u32(T), bp_tbl = smem_u32(s_bp4);
  const uint32_t w_floor = w_base + 4u * kGuardLo;

  if (threadIdx.x < 32) s_bp4[threadIdx.x] = (uint8_t)(4 * bit_pos32(threadIdx.x));

  // the window state is zeroed once; every finalize leaves it zeroed again (guard bands are never read)
  for (int i = threadIdx.x; i < kZero4; i += kCplThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) s_npass = 0;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();

    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    int npass_w = 0;
    // the window's rows are split evenly over the warps, each warp's share into equal tiles of at most rpt rows
    // Tiles: full rounds of kCplWarps tiles of rpt rows, then the remaining rows split evenly over the warps, so that
    // all warps reach the finalize barrier together (n_full rounds + at most one shorter tile each).
    const int n_rows = hi - lo;
    const int n_full = n_rows / (kCplWarps * rpt);
    const int rem_lo = lo + n_full * kCplWarps * rpt;
    const int rem_tile = (hi - rem_lo + kCplWarps - 1) / kCplWarps;  // <= rpt

    for (int t = 0; t <= n_full; t++) {
      int base, tend;
      if (t < n_full) { base = lo + (t * kCplWarps + warp) * rpt; tend = base + rpt; }
      else { base = rem_lo + warp * rem_tile; tend = min(base + rem_tile, hi); }
      if (base >= tend) break;
      // ------------------------------ rows: metadata, coverage deltas, unit descriptors ------------------------------
      const int r = base + lane;
      const bool valid = r < tend;
      const long long o0 = csr.off[valid ? r : tend];
      long long o1 = __shfl_down_sync(FULL, o0, 1);
      if (lane == 31) o1 = csr.off[tend];
      const long long tile_o0 = __shfl_sync(FULL, o0, 0);
      const int n_units = (int)((__shfl_sync(FULL, o1, 31) - tile_o0) >> USH);
      int start = 0, spv = 0;
      if (valid) { start = csr.start[r]; spv = csr.sp[r]; }
      const int nu = (int)((o1 - o0) >> USH);        // units of this row
      const int ubeg = (int)((o0 - tile_o0) >> USH);  // its first unit in the tile
      const int strand0 = spv & 1;
      const int len = (int)(o1 - o0) - (spv >> 4);
      const int rel0 = start - w0;
      const bool overlaps = valid && len > 0 && rel0 + len > 0;  // rel0 < S by construction of hi
      {
        // descriptor of unit c: T-relative word index of its byte 0 (strand0*SG + rel + kGuardLo) when the unit
        // overlaps the window, else 0; bit 31 = pass flag in the two-pass kernel
        int relg = rel0 + kGuardLo;
        uint32_t flag = 0u;
        if (overlaps) {
          const uint32_t sinc = strand0 ? 0x10000u : 1u;
          red_shared(w_floor + 4u * (uint32_t)max(rel0, 0), sinc);
          const int e = rel0 + len;
          if (e < S) red_shared(w_floor + 4u * (uint32_t)e, 0u - sinc);
          if (!FUSED && (pass_in == nullptr || pass_in[r] != 0)) flag = 0x80000000u;
        } else {
          relg = -0x40000000;  // never in the window
        }
        uint32_t* dst = aux_d + ubeg;
        const uint32_t sbase = (uint32_t)(strand0 * SG) | flag;
        for (int c = 0; c < nu; c++) {  // in the window: -UB < rel < S
          dst[c] = ((unsigned)(relg - (kGuardLo - UB) - 1) < (unsigned)(S + UB - 1)) ? sbase + (uint32_t)relg : 0u;
          relg += UB;
        }
      }
      __syncwarp();

      // ------------------------------ units: classification, counts, pass-independent events ------------------------------
      const uint8_t* ptr = csr.data + tile_o0 + (size_t)lane * UB;
      for (int j0 = lane; j0 < n_units; j0 += 32 * BATCH, ptr += 32 * UB * BATCH) {
        uint4 va[BATCH], vb[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++)
          if (j0 + 32 * u < n_units) {
            if (UB == 32) ldg_unit(ptr + 32 * UB * u, va[u], vb[u]);
            else ldg_unit(ptr + 32 * UB * u, va[u]);
          }
#pragma unroll
        for (int u = 0; u < BATCH; u++) {
          const int j = j0 + 32 * u;
          if (j >= n_units) break;
          const uint32_t d = aux_d[j];
          const uint32_t ix = d & 0x7FFFFFFFu;  // strand*SG + rel + kGuardLo, 0 = not in the window
          const uint32_t s = ix >= (uint32_t)SG ? 1u : 0u;
          const uint32_t rel4 = 4u * (ix - s * (uint32_t)SG);
          const uint32_t sinc = s ? 0x10000u : 1u;
          uint32_t m[3], meth[3], E = 0u, cnt = 0u;
          {
            Planes p;
            make_planes(va[u], p);
            want_masks<WANT>(p, m, meth);
            if (FUSED) cnt = count4<TCTX>(p);
            if (ix) E = dash_events(p, w_base + rel4, sinc, w_floor, bp_tbl);
          }
          if (UB == 32) {
            Planes p;
            make_planes(vb[u], p);
            uint32_t m2[3], meth2[3];
            want_masks<WANT>(p, m2, meth2);
            if (FUSED) cnt += count4<TCTX>(p);
            if (ix) E |= dash_events(p, w_base + rel4 + 64u, sinc, w_floor, bp_tbl) << 2;
#pragma unroll
            for (int k = 0; k < 3; k++) { m[k] |= m2[k] << 2; meth[k] |= meth2[k] << 2; }
          }
          if (ix) {
            const uint32_t methpass = FUSED ? 0u : ((d >> 31) ? (meth[0] | meth[1] | meth[2]) : 0u);
            unit_events<NC, SG>(m, E, methpass, d_base + rel4, t_base + 4u * ix, sinc, bp_tbl);
          }
          if (FUSED) {
            aux_d[j] = cnt;
#pragma unroll
            for (int k = 0; k < NC; k++) aux_m[k * tc + j] = ix ? meth[k] : 0u;
          }
        }
      }

      // ------------------------------ rows: class counts -> pass flag -> methylated calls of passing rows ------------------------------
      if (FUSED) {
        __syncwarp();
        bool pass = false;
        if (overlaps) {
          uint32_t acc_in = 0, acc_out = 0;
          const uint32_t* src = aux_d + ubeg;
          for (int c0 = 0; c0 < nu; c0 += FLUSH) {  // byte-wide fields hold the counts of FLUSH units
            uint32_t part = 0;
            const int c1 = min(c0 + FLUSH, nu);
            for (int c = c0; c < c1; c++) part += src[c];
            acc_in += (part & 0xFFu) | ((part & 0xFF00u) << 8);
            acc_out += ((part >> 16) & 0xFFu) | ((part >> 24) << 16);
          }
          pass = eval_pass_tables(acc_in >> 16, acc_in & 0xFFFFu, acc_out >> 16, acc_out & 0xFFFFu, thr, tbl);
          if (rel0 >= 0) {  // the window that holds the row's start owns its flag
            pass_out[r] = pass ? 1 : 0;
            npass_w += pass;
          }
        }
        // methylated calls of the passing rows, one row at a time with a lane per unit
        const uint32_t ta = t_base + 4u * (uint32_t)(strand0 * SG + rel0 + kGuardLo);
        for (uint32_t pm = __ballot_sync(FULL, pass); pm; pm &= pm - 1u) {
          const int src_lane = __ffs(pm) - 1;
          const int ub = __shfl_sync(FULL, ubeg, src_lane), n = __shfl_sync(FULL, nu, src_lane);
          const uint32_t ta_s = __shfl_sync(FULL, ta, src_lane);
          for (int c = lane; c < n; c += 32) {
#pragma unroll
            for (int k = 0; k < NC; k++) {
              const uint32_t mm = aux_m[k * tc + ub + c];
              if (mm) push_meth(mm, ta_s + 4u * UB * c + k * 2 * SG * 4, bp_tbl);
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

    window_finalize<WANT, S, kCplThreads, FUSED>(W, D, T, w0, win, out, s_scan, &s_base, &s_npass);
  }
}

}  // namespace ea
