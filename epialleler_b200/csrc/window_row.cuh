// Short-read window kernel, "row per lane, straight from HBM" formulation: the fast path for CpG/CHG/CHH
// reports (one requested context) over reads of up to MAXU*32 bytes whose padded rows are multiples of 32 bytes
// (always true for 150 bp reads: 160-byte rows).
//
// One CTA per genomic window of S positions, 8 warps, a tile = 32 consecutive rows, ONE LANE PER ROW.  A lane
// walks its own row in 32-byte units with 256-bit loads (LDG.E.ENL2.256): every 32-byte sector that crosses the
// L2->SM link is used completely, so nothing is staged in shared memory, the only shared-memory state is the
// window itself (win_common.cuh) and five CTAs of 8 warps fit one SM.  Per row, everything lives in registers:
//   * the class counts of the pass/fail chain (src/rcpp_threshold_reads.cpp:39-70; exact integer tables);
//   * the methylated-call masks of its units, so that a row that passes adds its methylated calls after the
//     flag is known without touching the data again (SURVEY.md 8a invariant 4);
// {total calls} of the requested context, '-' bytes and the row's coverage deltas are pushed while the row streams
// by.  There are no per-tile staging phases, no cross-lane reductions and no descriptors: row metadata loads are
// coalesced over 32 rows and the pass chain runs for 32 rows per instruction.
// Rows of different lengths leave lanes idle towards the end of a tile; reads of widely varying length, CX reports
// and rows that are not multiples of 32 bytes take the chunk-per-lane kernel (window_cpl.cuh).
// Preconditions (checked by the dispatcher): one requested context, every padded row a multiple of 32 bytes and
// at most MAXU*32 bytes, at most 32767 rows overlap a window.
#pragma once
#include "win_common.cuh"
#include "window_cpl.cuh"  // ldg_unit

namespace ea {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;
#ifndef EA_ROW_MINB
#define EA_ROW_MINB 5   // measured on B200: 5 CTAs/SM (48 registers) beats 4 and 6
#endif

constexpr int kRowPer = 12;                        // positions per thread in the finalize: windows of up to 3072 positions
constexpr int kRowMaxS = kRowPer * kRowThreads;

template <int TCTX, int WANT, bool FUSED, int MAXU>
__global__ void __launch_bounds__(kRowThreads, EA_ROW_MINB)
k_window_row(Csr csr, WinTable wt, ThrParams thr, ThrTables tbl, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
             ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  static_assert(NC == 1, "k_window_row handles one requested context");
  // thresholds on the reported context (the default: CG / CG): no separate in-context classification for the counts
  constexpr bool SAME = FUSED && in_from_want<TCTX, WANT>();
  const int S = wt.S;                                 // window size: runtime, chosen by the dispatcher (multiple of 32)
  const int SG = kGuardLo + S + kGuardHi;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[2 * kRowWarps];
  __shared__ long long s_base;
  __shared__ int s_npass;
  __shared__ __align__(4) uint8_t s_bp4[32];
  // the pass chain's integer tables (win_common.cuh) for counts up to the longest row this instance handles
  __shared__ uint16_t s_mmin[FUSED ? MAXU * 32 + 1 : 1], s_omax[FUSED ? MAXU * 32 + 1 : 1];

  uint32_t* const W = reinterpret_cast<uint32_t*>(smem_raw);  // [SG]
  uint32_t* const D = W + SG;                                 // [SG]
  uint32_t* const T = D + SG;                                 // [2][SG]
  const int kZero4 = (NC + 1) * 2 * SG / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t w_base = smem_u32(W), d_base = smem_u32(D), t_base = smem_u32(T), bp_tbl = smem_u32(s_bp4);
  const uint32_t w_floor = w_base + 4u * kGuardLo;

  if (threadIdx.x < 32) s_bp4[threadIdx.x] = (uint8_t)(4 * bit_pos32(threadIdx.x));
  if (FUSED) stage_thr_tables(tbl, s_mmin, s_omax, MAXU * 32);

  // the window state is zeroed once; every finalize leaves it zeroed again (guard bands are never read)
  for (int i = threadIdx.x; i < kZero4; i += kRowThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) s_npass = 0;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();

    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    int npass_w = 0;
    // Tiles of 32 consecutive rows, dealt round robin to the warps.  The kernel is bound by instruction issue, so a full
    // tile on some warps (the others wait at the barrier and leave their issue slots to the co-resident CTAs) beats
    // spreading the window's last rows over all warps as partly filled tiles (measured).
    for (int base = lo + warp * 32; base < hi; base += kRowWarps * 32) {
      const int tend = min(base + 32, hi);
      const int r = base + lane;
      const bool valid = r < tend;
      const long long o0 = csr.off[valid ? r : tend];
      long long o1 = __shfl_down_sync(FULL, o0, 1);
      if (lane == 31) o1 = csr.off[tend];
      int start = 0, spv = 0;
      if (valid) { start = csr.start[r]; spv = csr.sp[r]; }
      const int nu = (int)((o1 - o0) >> 5);
      const int strand0 = spv & 1;
      const int len = (int)(o1 - o0) - (spv >> 4);
      const int rel0 = start - w0;
      const bool overlaps = valid && len > 0 && rel0 + len > 0;  // rel0 < S by construction of hi
      if (overlaps) {
        const uint32_t sinc = strand0 ? 0x10000u : 1u;
        red_shared(w_floor + 4u * (uint32_t)max(rel0, 0), sinc);
        const int e = rel0 + len;
        if (e < S) red_shared(w_floor + 4u * (uint32_t)e, 0u - sinc);
        // byte addresses of the row's position 0 in W/D (strand independent) and in T[strand0]
        const uint32_t rel4 = 4u * (uint32_t)(rel0 + kGuardLo);
        const uint32_t wq = w_base + rel4, dq = d_base + rel4, tq = t_base + 4u * (uint32_t)(strand0 * SG) + rel4;
        const bool pass_known = FUSED ? false : (pass_in == nullptr || pass_in[r] != 0);
        const uint8_t* row = csr.data + o0;
        uint32_t cnt = 0;        // byte-wide class counts of the row (<= 8 units * 32 bytes = 256 > 255 only if every byte is one class)
        uint32_t acc_in = 0, acc_out = 0;
        uint32_t methm[MAXU];
        uint4 va, vb, na, nb;
        ldg_unit(row, va, vb);
#pragma unroll
        for (int u = 0; u < MAXU; u++) {
          methm[u] = 0u;
          if (u < nu) {
            if (u + 1 < MAXU && u + 1 < nu) ldg_unit(row + 32 * (u + 1), na, nb);
            const int rel = rel0 + 32 * u;
            const bool inwin = (unsigned)(rel + 31) < (unsigned)(S + 31);  // -32 < rel < S
            uint32_t m[3], meth[3], E = 0u;
            uint32_t oc = 0u, ocm = 0u;  // SAME: out-of-context calls of the unit and their methylated subset
            {
              Planes p;
              make_planes(va, p);
              want_masks<WANT>(p, m, meth);
              if (FUSED && !SAME) cnt += count4<TCTX>(p);
              if (FUSED && SAME) out_masks<TCTX>(p, oc, ocm);
              if (inwin) E = dash_events(p, wq + 128u * u, sinc, w_floor, bp_tbl);
            }
            {
              Planes p;
              make_planes(vb, p);
              uint32_t m2[3], meth2[3];
              want_masks<WANT>(p, m2, meth2);
              if (FUSED && !SAME) cnt += count4<TCTX>(p);
              if (FUSED && SAME) {
                uint32_t o2, om2;
                out_masks<TCTX>(p, o2, om2);
                oc += o2 << 2; ocm += om2 << 2;  // disjoint bits: the second chunk's indicators sit two bits higher
              }
              if (inwin) E |= dash_events(p, wq + 128u * u + 64u, sinc, w_floor, bp_tbl) << 2;
              m[0] |= m2[0] << 2; meth[0] |= meth2[0] << 2;
            }
            // thresholds and report on the same context: the in-context counts are the popcounts of the event masks
            if (FUSED && SAME) cnt += count_from_masks<TCTX, WANT>(m, meth, oc, ocm);
            m[1] = m[2] = 0u;
            if (inwin) {
              unit_events<NC, 0>(m, E, (!FUSED && pass_known) ? meth[0] : 0u, dq + 128u * u, tq + 128u * u, sinc, bp_tbl);  // NC == 1: the SG stride is unused
              methm[u] = meth[0];
            }
            if (FUSED && (u & 3) == 3) {  // byte-wide fields hold the counts of four units
              acc_in += (cnt & 0xFFu) | ((cnt & 0xFF00u) << 8);
              acc_out += ((cnt >> 16) & 0xFFu) | ((cnt >> 24) << 16);
              cnt = 0;
            }
            va = na; vb = nb;
          }
        }
        if (FUSED) {
          acc_in += (cnt & 0xFFu) | ((cnt & 0xFF00u) << 8);
          acc_out += ((cnt >> 16) & 0xFFu) | ((cnt >> 24) << 16);
          // a row of at most MAXU*32 bytes cannot count more than that: always inside the shared tables
          const uint32_t pm = acc_in >> 16, pall = acc_in & 0xFFFFu, pom = acc_out >> 16, poall = acc_out & 0xFFFFu;
          const bool pass = pm != 0u && pall >= thr.min_n && pm >= (uint32_t)s_mmin[pall] && (pom == 0u || pom <= (uint32_t)s_omax[poall]);
          if (rel0 >= 0) {  // the window that holds the row's start owns its flag
            pass_out[r] = pass ? 1 : 0;
            npass_w += pass;
          }
          if (pass) {
#pragma unroll
            for (int u = 0; u < MAXU; u++)
              if (methm[u]) push_meth(methm[u], tq + 128u * u, bp_tbl);
          }
        }
      }
    }
    if (FUSED) {
      for (int d = 16; d >= 1; d >>= 1) npass_w += __shfl_down_sync(FULL, npass_w, d);
      if (lane == 0 && npass_w) atomicAdd(&s_npass, npass_w);
    }
    __syncthreads();

    window_finalize_rt<WANT, kRowPer, kRowThreads, FUSED>(W, D, T, S, w0, win, out, s_scan, &s_base, &s_npass);
  }
}

}  // namespace ea
