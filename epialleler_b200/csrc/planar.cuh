// Bit-sliced ("planar") device format of the methylation-call strings and the kernels that read it.
//
// The hot path only ever looks at the LOW nibble of a packed byte (`& 15`, src/rcpp_cx_report.cpp:117,
// src/rcpp_threshold_reads.cpp:34, src/rcpp_get_xm_beta.cpp:24): the 4-bit call code of src/epialleleR.h:15-28.  The
// packer therefore keeps, next to the byte rows that the base-level paths (VCF, lMHL) read, a second copy of the codes
// laid out for the B200's integer pipes and memory system:
//
//   unit   = 32 consecutive positions of one row as FOUR 32-bit bit planes {A,B,C,D} (bit k of plane b = bit b of the
//            code at position 32*j + k): 16 bytes, one 128-bit load.  A class test over 32 positions is ONE LOP3
//            (CG = A&B&C, CHG = ~A&B&C, CHH = ~A&B&~C, methylated = &~D, '-' = A&B&~C&D), a class count one POPC,
//            and the bit index of an event IS its position - no nibble merge, no plane alignment, no position table.
//            Positions beyond the row's end hold code 0 (in no class, not a '-').
//   tile   = 32 consecutive rows (rows are sorted by (rname,start)); the units of a tile are INTERLEAVED:
//            unit j of row l sits at 16 * ((toff[tile] + j) * 32 + l), so when lane l of a warp walks row l the warp
//            reads 512 contiguous bytes per instruction.  A tile has max(ceil(len/32)) unit rows; toff[] is their
//            exclusive prefix sum.
//   per row: start (int32) and `ls` = len | (strand=='-') << 31 (uint32).  No per-row byte offsets.
//
// 150 bp reads: 80 bytes of call planes + 8 bytes of metadata per read (the byte rows: 160 + 13).
//
// The window state in shared memory, the event model and the finalize are those of win_common.cuh.
#pragma once
#include "win_common.cuh"
#include "window_row.cuh"  // kRowThreads, kRowWarps, kRowPer, kRowMaxS

namespace ea {

struct Planar {
  const uint4* units;
  const int64_t* toff;   // [n_tiles + 1] unit rows in front of the tile
  const uint32_t* ls;    // [n] len | strand0 << 31
  const int32_t* start;  // [n]
  int64_t n;
};

__device__ __forceinline__ uint4 ldg_plunit(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
// pack time: byte rows -> planar tiles
// ---------------------------------------------------------------------------------------------------------------

// unit rows of every tile of 32 rows: max over its rows of ceil(len / 32); also ls[]
__global__ void k_pl_tile_units(const int32_t* __restrict__ len, const uint8_t* __restrict__ strand, int64_t n,
                                int64_t* __restrict__ tile_nu, uint32_t* __restrict__ ls) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // blockDim.x is a multiple of 32
  int nu = 0;
  if (r < n) {
    const int l = len[r];
    nu = (l + 31) >> 5;
    ls[r] = (uint32_t)l | (strand[r] == 2 ? 0x80000000u : 0u);
  }
  for (int d = 16; d >= 1; d >>= 1) nu = max(nu, __shfl_xor_sync(0xFFFFFFFFu, nu, d));
  if ((threadIdx.x & 31) == 0 && (r >> 5) <= ((n - 1) >> 5)) tile_nu[r >> 5] = nu;
}

// bit b of each of the four bytes of w, gathered into bits 0..3 (byte k -> bit k)
__device__ __forceinline__ uint32_t gather_bit(uint32_t w, int b) {
  return ((((w >> b) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu;
}

// Tiles [tile0, tile1): one warp per tile, lane l transposes row l.  `data`/`off`: the padded byte rows (kernels.cuh).
__global__ void __launch_bounds__(256)
k_planarize(const uint8_t* __restrict__ data, const int64_t* __restrict__ off, const int32_t* __restrict__ len, int64_t n,
            const int64_t* __restrict__ toff, int64_t tile0, int64_t tile1, uint4* __restrict__ units) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = tile0 + warp; t < tile1; t += n_warps) {
    const int64_t r = t * 32 + lane;
    const int64_t u0 = toff[t];
    const int nu_t = (int)(toff[t + 1] - u0);
    int l = 0;
    long long o = 0;
    if (r < n) { l = len[r]; o = off[r]; }
    const int padded = (l + 15) & ~15;
    uint4* dst = units + u0 * 32 + lane;
    for (int j = 0; j < nu_t; j++) {
      uint32_t P[4] = {0u, 0u, 0u, 0u};
      const int b0 = 32 * j;
      if (b0 < l) {
        const uint4 lo4 = *reinterpret_cast<const uint4*>(data + o + b0);
        uint4 hi4 = make_uint4(0, 0, 0, 0);
        if (b0 + 16 < padded) hi4 = *reinterpret_cast<const uint4*>(data + o + b0 + 16);
        const uint32_t w[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
        for (int k = 0; k < 8; k++)
#pragma unroll
          for (int b = 0; b < 4; b++) P[b] |= gather_bit(w[k], b) << (4 * k);
        const int nvalid = l - b0;  // pad bytes (code 3) and bytes beyond the padded row become code 0
        if (nvalid < 32) {
          const uint32_t keep = (1u << nvalid) - 1u;
#pragma unroll
          for (int b = 0; b < 4; b++) P[b] &= keep;
        }
      }
      dst[(int64_t)j * 32] = make_uint4(P[0], P[1], P[2], P[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// events of one unit (32 positions) on the shared-memory window state
// ---------------------------------------------------------------------------------------------------------------

// calls of the requested contexts (slot order CHH, CHG, CG among the requested) of one unit
template <int WANT>
__device__ __forceinline__ void pl_want(const uint4& v, uint32_t (&m)[3]) {
  int slot = 0;
  m[0] = m[1] = m[2] = 0u;
  if (WANT & W_CHH) m[slot++] = cls_chh(v.x, v.y, v.z);
  if (WANT & W_CHG) m[slot++] = cls_chg(v.x, v.y, v.z);
  if (WANT & W_CG) m[slot++] = cls_cg(v.x, v.y, v.z);
}

// '-' bytes of a unit: E = the unit's '-' positions, Eprev = those of the row's previous unit.  Isolated ones (returned)
// are counted in D by the caller; the edges of longer runs become coverage deltas in W right here: -1 where a run starts,
// +1 right after it ends, clamped to the window start.  w_addr = shared BYTE address of the unit's position 0 in W.
__device__ __forceinline__ uint32_t pl_dash(uint32_t E, uint32_t Eprev, uint32_t w_addr, uint32_t sinc, uint32_t w_floor) {
  const uint32_t prev = __funnelshift_l(Eprev, E, 1);                    // "the position before me is a '-'"
  const uint32_t iso = E & ~prev & ~(E >> 1) & 0x7FFFFFFFu;              // bit 31 may continue in the next unit: an edge
  uint32_t ed = (E ^ prev) & ~(iso | (iso << 1));
  while (ed) {
    const int i = 31 - __clz(ed);
    const uint32_t bit = 1u << i;
    ed ^= bit;
    red_shared(max(w_addr + 4u * (uint32_t)i, w_floor), (E & bit) ? 0u - sinc : sinc);
  }
  return iso;
}

// m[k] = calls of requested context k, iso = isolated '-' positions, methnow = methylated calls to be counted right now
// (0 in the fused kernels).  d_addr / t_addr = shared BYTE addresses of the unit's position 0 in D and in T[0][strand].
template <int NC>
__device__ __forceinline__ void pl_events(const uint32_t (&m)[3], uint32_t iso, uint32_t methnow, uint32_t d_addr, uint32_t t_addr,
                                          uint32_t ctx_stride, uint32_t sinc) {
  uint32_t ev = m[0] | m[1] | m[2] | iso;
  while (ev) {
    const int i = 31 - __clz(ev);
    const uint32_t bit = 1u << i;
    ev ^= bit;
    uint32_t a = t_addr;
    uint32_t v = (methnow & bit) ? 0x10001u : 1u;
    if (NC > 1 && (m[1] & bit)) a = t_addr + ctx_stride;
    if (NC > 2 && (m[2] & bit)) a = t_addr + 2u * ctx_stride;
    if (iso & bit) { a = d_addr; v = sinc; }
    red_shared(a + 4u * (uint32_t)i, v);
  }
}

// methylated calls of a passing row's unit (fused kernels, after the flag is known)
__device__ __forceinline__ void pl_push_meth(uint32_t mm, uint32_t t_addr) {
  while (mm) {
    const int i = 31 - __clz(mm);
    mm ^= 1u << i;
    red_shared(t_addr + 4u * (uint32_t)i, 0x10000u);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Window kernel on planar tiles: one CTA per genomic window, a warp per tile, ONE LANE PER ROW.
//   MAXU > 0 : rows of at most MAXU units; the methylated-call masks of a row's units stay in registers until its pass
//              flag is known (SURVEY.md 8a invariant 4), the unit loop is unrolled, all loads of a row are in flight at once;
//   MAXU == 0: any row length up to 1024 positions; passing rows read their in-window units a second time (L2 hits).
// The pass/fail chain (src/rcpp_threshold_reads.cpp:39-70) runs for 32 rows per instruction through the exact integer
// tables of win_common.cuh, staged in shared memory.
// Preconditions (dispatcher): rows <= 1024 positions, at most 32767 rows overlap a window, S a multiple of 32, <= kRowMaxS.
// ---------------------------------------------------------------------------------------------------------------
#ifndef EA_PL_MINB
#define EA_PL_MINB 5
#endif

template <int TCTX, int WANT, bool FUSED, int MAXU>
__global__ void __launch_bounds__(kRowThreads, EA_PL_MINB)
k_window_pl(Planar pl, WinTable wt, ThrParams thr, ThrTables tbl, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
            ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr bool KEEP = MAXU > 0;
  constexpr int KU = KEEP ? MAXU : 1;
  constexpr int TBL = KEEP ? MAXU * 32 : kThrTable - 1;   // largest class count a row of this instance can reach
  const int S = wt.S;
  const int SG = kGuardLo + S + kGuardHi;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[2 * kRowWarps];
  __shared__ long long s_base;
  __shared__ int s_npass;
  __shared__ uint16_t s_mmin[FUSED ? TBL + 1 : 1], s_omax[FUSED ? TBL + 1 : 1];

  uint32_t* const W = reinterpret_cast<uint32_t*>(smem_raw);  // [SG]
  uint32_t* const D = W + SG;                                 // [SG]
  uint32_t* const T = D + SG;                                 // [NC][2][SG]
  const int kZero4 = (NC + 1) * 2 * SG / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t w_base = smem_u32(W), d_base = smem_u32(D), t_base = smem_u32(T);
  const uint32_t w_floor = w_base + 4u * kGuardLo;
  const uint32_t ctx_stride = 8u * (uint32_t)SG;  // bytes between T[c][s] and T[c+1][s]

  if (FUSED) stage_thr_tables(tbl, s_mmin, s_omax, TBL);
  for (int i = threadIdx.x; i < kZero4; i += kRowThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) s_npass = 0;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    int npass_w = 0;
    const int t_hi = (hi + 31) >> 5;
    for (int t = (lo >> 5) + warp; t < t_hi; t += kRowWarps) {
      const int r = t * 32 + lane;
      const bool valid = r >= lo && r < hi;
      const long long tu0 = pl.toff[t];
      int start = 0;
      uint32_t lsv = 0;
      if (valid) { start = pl.start[r]; lsv = pl.ls[r]; }
      const int len = (int)(lsv & 0x7FFFFFFFu);
      const uint32_t strand0 = lsv >> 31;
      const int rel0 = start - w0;
      const bool overlaps = valid && len > 0 && rel0 + len > 0;  // rel0 < S by construction of hi
      const int nu = overlaps ? (len + 31) >> 5 : 0;
      const uint4* src = pl.units + tu0 * 32 + lane;
      const uint32_t sinc = strand0 ? 0x10000u : 1u;
      const uint32_t rel4 = 4u * (uint32_t)(rel0 + kGuardLo);
      const uint32_t wq = w_base + rel4, dq = d_base + rel4, tq = t_base + 4u * strand0 * (uint32_t)SG + rel4;
      const bool pass_known = FUSED ? false : (valid && (pass_in == nullptr || pass_in[r] != 0));
      if (overlaps) {
        red_shared(w_floor + 4u * (uint32_t)max(rel0, 0), sinc);
        const int e = rel0 + len;
        if (e < S) red_shared(w_floor + 4u * (uint32_t)e, 0u - sinc);
      }
      uint32_t n_in = 0, n_inm = 0, n_out = 0, n_outm = 0;
      uint32_t Eprev = 0;
      bool entered = false;
      uint32_t methm[KU][NC];

      auto do_unit = [&](const uint4& v, int u, uint32_t (&mk)[NC]) {
        const int rel = rel0 + 32 * u;
        const bool inwin = (unsigned)(rel + 31) < (unsigned)(S + 31);  // -32 < rel < S
        if (FUSED) {
          const uint32_t ci = cls_in<TCTX>(v.x, v.y, v.z), co = cls_out<TCTX>(v.x, v.y, v.z);
          n_in += __popc(ci); n_inm += __popc(ci & ~v.w);
          if (TCTX != TC_CX) { n_out += __popc(co); n_outm += __popc(co & ~v.w); }
        }
        const uint32_t E = v.x & v.y & ~v.z & v.w;
#pragma unroll
        for (int c = 0; c < NC; c++) mk[c] = 0u;
        if (inwin) {
          const uint32_t off4 = 128u * (uint32_t)u;
          if (!entered && (Eprev >> 31)) red_shared(w_floor, 0u - sinc);  // a '-' run is open where the row enters the window
          entered = true;
          uint32_t m[3];
          pl_want<WANT>(v, m);
          const uint32_t iso = pl_dash(E, Eprev, wq + off4, sinc, w_floor);
          uint32_t methnow = 0u;
#pragma unroll
          for (int c = 0; c < NC; c++) { mk[c] = m[c] & ~v.w; methnow |= mk[c]; }
          pl_events<NC>(m, iso, pass_known ? methnow : 0u, dq + off4, tq + off4, ctx_stride, sinc);
        }
        Eprev = E;
      };

      if (KEEP) {
        uint4 v[KU];
#pragma unroll
        for (int u = 0; u < KU; u++)
          if (u < nu) v[u] = ldg_plunit(src + u * 32);
#pragma unroll
        for (int u = 0; u < KU; u++) {
#pragma unroll
          for (int c = 0; c < NC; c++) methm[u][c] = 0u;
          if (u < nu) do_unit(v[u], u, methm[u]);
        }
      } else {
        uint4 cur, nxt;
        if (nu > 0) cur = ldg_plunit(src);
        for (int u = 0; u < nu; u++) {
          if (u + 1 < nu) nxt = ldg_plunit(src + (u + 1) * 32);
          uint32_t mk[NC];
          do_unit(cur, u, mk);
          cur = nxt;
        }
      }
      if (overlaps) {
        const int e = rel0 + len;
        if ((len & 31) == 0 && (Eprev >> 31) && e < S) red_shared(max(wq + 4u * (uint32_t)len, w_floor), sinc);  // a run that reaches the row end
      }
      if (FUSED && overlaps) {
        const uint32_t all = n_in, oall = n_out;
        bool pass;
        if (KEEP) pass = n_inm != 0u && all >= thr.min_n && n_inm >= (uint32_t)s_mmin[all] && (n_outm == 0u || n_outm <= (uint32_t)s_omax[oall]);
        else pass = eval_pass_staged(n_inm, all, n_outm, oall, thr, s_mmin, s_omax, (uint32_t)TBL);
        if (rel0 >= 0) {  // the window that holds the row's start owns its flag
          pass_out[r] = pass ? 1 : 0;
          npass_w += pass;
        }
        if (pass) {
          if (KEEP) {
#pragma unroll
            for (int u = 0; u < KU; u++)
#pragma unroll
              for (int c = 0; c < NC; c++)
                if (methm[u][c]) pl_push_meth(methm[u][c], tq + 128u * u + (uint32_t)c * ctx_stride);
          } else {
            const int u_lo = rel0 <= -32 ? (-rel0) >> 5 : 0;
            const int u_hi = min(nu, (S - rel0 + 31) >> 5);
            for (int u = u_lo; u < u_hi; u++) {
              const uint4 v = ldg_plunit(src + u * 32);
              uint32_t m[3];
              pl_want<WANT>(v, m);
#pragma unroll
              for (int c = 0; c < NC; c++) {
                const uint32_t mm = m[c] & ~v.w;
                if (mm) pl_push_meth(mm, tq + 128u * u + (uint32_t)c * ctx_stride);
              }
            }
          }
        }
      }
    }
    if (FUSED) {
      for (int d = 16; d >= 1; d >>= 1) npass_w += __shfl_down_sync(0xFFFFFFFFu, npass_w, d);
      if (lane == 0 && npass_w) atomicAdd(&s_npass, npass_w);
    }
    __syncthreads();
    window_finalize_rt<WANT, kRowPer, kRowThreads, FUSED>(W, D, T, S, w0, win, out, s_scan, &s_base, &s_npass);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Stand-alone per-read kernel on planar tiles: rcpp_threshold_reads (pass flags) and rcpp_get_xm_beta (betas).
// A warp per tile, one lane per row, 512 contiguous bytes per warp load, no shared memory but the staged tables.
// ---------------------------------------------------------------------------------------------------------------
template <int TCTX>
__global__ void __launch_bounds__(256)
k_threshold_pl(Planar pl, int64_t n_tiles, ThrParams thr, ThrTables tbl, uint8_t* __restrict__ pass_out, double* __restrict__ beta_out,
               unsigned long long* n_pass) {
  __shared__ uint16_t s_mmin[kThrTable], s_omax[kThrTable];
  if (pass_out) {
    stage_thr_tables(tbl, s_mmin, s_omax, kThrTable - 1);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t my_pass = 0;
  for (int64_t t = warp; t < n_tiles; t += n_warps) {
    const int64_t r = t * 32 + lane;
    const int64_t tu0 = pl.toff[t];
    uint32_t lsv = 0;
    if (r < pl.n) lsv = pl.ls[r];
    const int nu = (int)(((lsv & 0x7FFFFFFFu) + 31u) >> 5);
    const uint4* src = pl.units + tu0 * 32 + lane;
    uint32_t n_in = 0, n_inm = 0, n_out = 0, n_outm = 0;
    uint4 cur, nxt, nx2;
    if (nu > 0) cur = ldg_plunit(src);
    if (nu > 1) nxt = ldg_plunit(src + 32);
    for (int u = 0; u < nu; u++) {
      if (u + 2 < nu) nx2 = ldg_plunit(src + (u + 2) * 32);
      const uint32_t ci = cls_in<TCTX>(cur.x, cur.y, cur.z), co = cls_out<TCTX>(cur.x, cur.y, cur.z);
      n_in += __popc(ci); n_inm += __popc(ci & ~cur.w);
      if (TCTX != TC_CX) { n_out += __popc(co); n_outm += __popc(co & ~cur.w); }
      cur = nxt; nxt = nx2;
    }
    if (r < pl.n) {
      if (pass_out) {
        const bool pass = eval_pass_staged(n_inm, n_in, n_outm, n_out, thr, s_mmin, s_omax, kThrTable - 1);
        pass_out[r] = pass ? 1 : 0;
        my_pass += pass;
      }
      if (beta_out) beta_out[r] = (double)n_inm / (double)(n_in ? n_in : 1u);
    }
  }
  if (n_pass) {
    for (int d = 16; d >= 1; d >>= 1) my_pass += __shfl_down_sync(0xFFFFFFFFu, my_pass, d);
    if (lane == 0 && my_pass) atomicAdd(n_pass, (unsigned long long)my_pass);
  }
}

}  // namespace ea
