// Long-read window kernel (reads of kilobases: nanopore / PacBio methylation calls, BASELINE.json configs[3]).
//
// With reads far longer than a window, a window is crossed by only ~coverage rows and each of them contributes a long
// contiguous byte range, so the work item is the in-window part of a row: a warp takes rows round robin, its lanes read
// 32 CONSECUTIVE 32-byte units of the row per step (a fully coalesced kilobyte per warp instruction, the next step's
// units requested before the current ones are processed, nothing staged in shared memory), classify them with the
// nibble planes and push their events into the window state of win_common.cuh ('-' bytes into D, long '-' runs as coverage deltas, calls into T).  The pass flags are
// an input: thresholding a long read needs the counts of the whole read, which is the stand-alone k_threshold pass
// (it runs at ~75 % of the HBM roofline on long rows) - the reference advises threshold.reads=FALSE for long reads
// anyway (R/generateCytosineReport.R:97-98).  Finalize: window_finalize_rt (three barriers, leaves the state zeroed).
// Preconditions (dispatcher): at most 32767 rows overlap a window, S a multiple of 32 and <= kLongMaxS.
#pragma once
#include "win_common.cuh"
#include "window_cpl.cuh"  // ldg_unit
#include "window_row.cuh"  // kRowThreads, kRowWarps, kRowPer, kRowMaxS

namespace ea {

#ifndef EA_LONG_MINB
#define EA_LONG_MINB 4
#endif
#ifndef EA_LONG_S
#define EA_LONG_S 3072  // window size for one requested context (more contexts: 2048)
#endif
constexpr int kLongPer = (EA_LONG_S + kRowThreads - 1) / kRowThreads / 4 * 4 < (EA_LONG_S + kRowThreads - 1) / kRowThreads
                             ? ((EA_LONG_S + kRowThreads - 1) / kRowThreads / 4 + 1) * 4 : (EA_LONG_S + kRowThreads - 1) / kRowThreads;
constexpr int kLongMaxS = kLongPer * kRowThreads;

// one 32-byte unit of a row: a 256-bit load when rows are 32-byte aligned, two 128-bit loads otherwise; when the row
// has an odd number of 16-byte chunks the second half of its last unit is replaced by pad bytes (code 3)
template <bool A32>
__device__ __forceinline__ void load_unit(const uint8_t* p, bool second_half, uint4& a, uint4& b) {
  if (A32) {
    ldg_unit(p, a, b);
  } else {
    ldg_unit(p, a);
    b = make_uint4(0x03030303u, 0x03030303u, 0x03030303u, 0x03030303u);
    if (second_half) ldg_unit(p + 16, b);
  }
}

// unit_events with a run-time distance between the context planes of T
template <int NC>
__device__ __forceinline__ void unit_events_rt(const uint32_t (&m)[3], uint32_t E, uint32_t methpass, uint32_t d_addr, uint32_t t_addr,
                                               uint32_t ctx_stride, uint32_t sinc, uint32_t bp_tbl) {
  uint32_t ev = m[0] | m[1] | m[2] | E;
  while (ev) {
    const int i = 31 - __clz(ev);
    const uint32_t bit = 1u << i;
    ev ^= bit;
    const uint32_t bp4 = pos4_of(bp_tbl, i);
    uint32_t a = t_addr;
    uint32_t v = (methpass & bit) ? 0x10001u : 1u;
    if (NC > 1 && (m[1] & bit)) a = t_addr + ctx_stride;
    if (NC > 2 && (m[2] & bit)) a = t_addr + 2u * ctx_stride;
    if (E & bit) { a = d_addr; v = sinc; }
    red_shared(a + bp4, v);
  }
}

template <int WANT, bool A32>
__global__ void __launch_bounds__(kRowThreads, EA_LONG_MINB)
k_window_long(Csr csr, WinTable wt, const uint8_t* __restrict__ pass_in, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  const int S = wt.S;
  const int SG = kGuardLo + S + kGuardHi;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[2 * kRowWarps];
  __shared__ long long s_base;
  __shared__ int s_npass;  // unused (no fused thresholding); window_finalize_rt takes the pointer
  __shared__ __align__(4) uint8_t s_bp4[32];

  uint32_t* const W = reinterpret_cast<uint32_t*>(smem_raw);  // [SG]
  uint32_t* const D = W + SG;                                 // [SG]
  uint32_t* const T = D + SG;                                 // [NC][2][SG]
  const int kZero4 = (NC + 1) * 2 * SG / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t w_base = smem_u32(W), d_base = smem_u32(D), t_base = smem_u32(T), bp_tbl = smem_u32(s_bp4);
  const uint32_t w_floor = w_base + 4u * kGuardLo;
  const uint32_t ctx_stride = 8u * (uint32_t)SG;  // bytes between T[c][s] and T[c+1][s]

  if (threadIdx.x < 32) s_bp4[threadIdx.x] = (uint8_t)(4 * bit_pos32(threadIdx.x));
  for (int i = threadIdx.x; i < kZero4; i += kRowThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) s_npass = 0;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    for (int r = lo + warp; r < hi; r += kRowWarps) {  // a warp per row, rows dealt round robin
      const long long o0 = csr.off[r], o1 = csr.off[r + 1];  // warp-uniform loads
      const int start = csr.start[r], spv = csr.sp[r];
      const int nch = (int)((o1 - o0) >> 4);
      const int nu = (nch + 1) >> 1;
      const int len = (int)(o1 - o0) - (spv >> 4);
      const int rel0 = start - w0;
      if (len <= 0 || rel0 + len <= 0) continue;  // rel0 < S by construction of hi
      const int u_lo = rel0 <= -32 ? (-rel0) >> 5 : 0;   // first unit with rel0 + 32u > -32
      const int u_hi = min(nu, (S - rel0 + 31) >> 5);    // first unit with rel0 + 32u >= S
      const uint32_t strand = (uint32_t)(spv & 1);
      const uint32_t sinc = strand ? 0x10000u : 1u;
      if (lane == 0) {  // the row's coverage deltas
        red_shared(w_floor + 4u * (uint32_t)max(rel0, 0), sinc);
        const int e = rel0 + len;
        if (e < S) red_shared(w_floor + 4u * (uint32_t)e, 0u - sinc);
      }
      const bool pass = pass_in == nullptr || pass_in[r] != 0;
      const uint8_t* row = csr.data + o0;
      const uint32_t t_row = t_base + 4u * strand * (uint32_t)SG;
      int u = u_lo + lane;
      uint4 va, vb, na, nb;
      if (u < u_hi) load_unit<A32>(row + 32ll * u, 2 * u + 1 < nch, va, vb);
      while (u < u_hi) {  // the lanes' trip counts differ by at most one
        const int un = u + 32;
        if (un < u_hi) load_unit<A32>(row + 32ll * un, 2 * un + 1 < nch, na, nb);
        const uint32_t rel4 = 4u * (uint32_t)(rel0 + 32 * u + kGuardLo);  // >= 4: rel0 + 32u > -32
        const uint32_t wq = w_base + rel4, dq = d_base + rel4, tq = t_row + rel4;
        uint32_t m[3], meth[3], E;
        {
          Planes p;
          make_planes(va, p);
          want_masks<WANT>(p, m, meth);
          E = dash_events(p, wq, sinc, w_floor, bp_tbl);
        }
        {
          Planes p;
          make_planes(vb, p);
          uint32_t m2[3], meth2[3];
          want_masks<WANT>(p, m2, meth2);
          E |= dash_events(p, wq + 64u, sinc, w_floor, bp_tbl) << 2;
#pragma unroll
          for (int c = 0; c < 3; c++) { m[c] |= m2[c] << 2; meth[c] |= meth2[c] << 2; }
        }
        unit_events_rt<NC>(m, E, pass ? (meth[0] | meth[1] | meth[2]) : 0u, dq, tq, ctx_stride, sinc, bp_tbl);
        va = na; vb = nb; u = un;
      }
    }
    __syncthreads();
    window_finalize_rt<WANT, kLongPer, kRowThreads, false>(W, D, T, S, w0, win, out, s_scan, &s_base, &s_npass);
  }
}

}  // namespace ea
