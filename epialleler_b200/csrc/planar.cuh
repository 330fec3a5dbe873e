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
//   long reads (data sets with rows of more than 1024 positions): tiles of ONE row, i.e. the units of a row are contiguous
//            (toff[] has one entry per row) and a WARP walks a row, 32 consecutive units = 512 contiguous bytes per load.
//
// 150 bp reads: 80 bytes of call planes + 8 bytes of metadata per read (the byte rows: 160 + 13).
//
// The window state in shared memory, the event model and the finalize are those of win_common.cuh.
#pragma once
#include <type_traits>
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

// long-read layout: units per row (a "tile" is one row) and ls[]
__global__ void k_pl_row_units(const int32_t* __restrict__ len, const uint8_t* __restrict__ strand, int64_t n,
                               int64_t* __restrict__ row_nu, uint32_t* __restrict__ ls) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int l = len[r];
  row_nu[r] = (l + 31) >> 5;
  ls[r] = (uint32_t)l | (strand[r] == 2 ? 0x80000000u : 0u);
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

// 32 bytes of a padded byte row (kernels.cuh) -> one unit; b0 = first byte, l = row length, padded = padded row length
__device__ __forceinline__ uint4 pl_unit_from_bytes(const uint8_t* __restrict__ row, int b0, int l, int padded) {
  uint32_t P[4] = {0u, 0u, 0u, 0u};
  if (b0 < l) {
    const uint4 lo4 = *reinterpret_cast<const uint4*>(row + b0);
    uint4 hi4 = make_uint4(0, 0, 0, 0);
    if (b0 + 16 < padded) hi4 = *reinterpret_cast<const uint4*>(row + b0 + 16);
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
  return make_uint4(P[0], P[1], P[2], P[3]);
}

// long-read layout: rows [row0, row1), one warp per row, lane l transposes units l, l + 32, ... (coalesced both ways)
__global__ void __launch_bounds__(256)
k_planarize_rows(const uint8_t* __restrict__ data, const int64_t* __restrict__ off, const int32_t* __restrict__ len,
                 const int64_t* __restrict__ toff, int64_t row0, int64_t row1, uint4* __restrict__ units) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = row0 + warp; r < row1; r += n_warps) {
    const int l = len[r];
    const int padded = (l + 15) & ~15;
    const int nu = (l + 31) >> 5;
    const uint8_t* row = data + off[r];
    uint4* dst = units + toff[r];
    for (int j = lane; j < nu; j += 32) dst[j] = pl_unit_from_bytes(row, 32 * j, l, padded);
  }
}

// 32 code bytes given as eight 32-bit words -> one unit, the first nvalid positions kept
__device__ __forceinline__ uint4 pl_unit_from_words(const uint32_t (&w)[8], int nvalid) {
  uint32_t P[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int k = 0; k < 8; k++)
#pragma unroll
    for (int b = 0; b < 4; b++) P[b] |= gather_bit(w[k], b) << (4 * k);
  if (nvalid < 32) {
    const uint32_t keep = nvalid <= 0 ? 0u : (1u << nvalid) - 1u;
#pragma unroll
    for (int b = 0; b < 4; b++) P[b] &= keep;
  }
  return make_uint4(P[0], P[1], P[2], P[3]);
}

// Tiles [tile0, tile1) straight from TIGHT strings (a chunk of the caller's blob as it arrived, no padded copy): `raw` holds
// the bytes from string offset `raw_base` on, soff[r - row0] is the offset of row r's string.  One warp per tile, lane l
// transposes row l; the unaligned source is read as aligned 32-bit words and assembled with funnel shifts.  The buffer
// behind `raw` must be readable 8 bytes past the chunk's last byte.
__global__ void __launch_bounds__(256)
k_planarize_tight(const uint8_t* __restrict__ raw, int64_t raw_base, const int64_t* __restrict__ soff, int64_t row0, int64_t n,
                  const int64_t* __restrict__ toff, int64_t tile0, int64_t tile1, int th, uint4* __restrict__ units) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = tile0 + warp; t < tile1; t += n_warps) {
    const int64_t u0 = toff[t];
    const int nu_t = (int)(toff[t + 1] - u0);
    if (th == 32) {
      const int64_t r = t * 32 + lane;
      long long so = 0;
      int l = 0;
      if (r < n) { so = soff[r - row0]; l = (int)(soff[r - row0 + 1] - so); }
      const uint8_t* src = raw + (so - raw_base);
      uint4* dst = units + u0 * 32 + lane;
      for (int j = 0; j < nu_t; j++) {
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int nv = l - 32 * j;
        if (nv > 0) {
          const uintptr_t p = reinterpret_cast<uintptr_t>(src + 32 * j);
          const uint32_t sh = (uint32_t)(p & 3) * 8u;
          const uint32_t* a = reinterpret_cast<const uint32_t*>(p & ~(uintptr_t)3);
          const int need = (int)(p & 3) + min(nv, 32);   // bytes of the aligned words that belong to this unit
          uint32_t x[9];
#pragma unroll
          for (int k = 0; k < 9; k++) x[k] = 4 * k < need ? a[k] : 0u;
#pragma unroll
          for (int k = 0; k < 8; k++) w[k] = __funnelshift_r(x[k], x[k + 1], sh);
        }
        dst[(int64_t)j * 32] = pl_unit_from_words(w, nv);
      }
    } else {  // long reads: a tile is a row, the lanes take its units in turn
      const long long so = soff[t - row0];
      const int l = (int)(soff[t - row0 + 1] - so);
      const uint8_t* src = raw + (so - raw_base);
      for (int j = lane; j < nu_t; j += 32) {
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int nv = l - 32 * j;
        if (nv > 0) {
          const uintptr_t p = reinterpret_cast<uintptr_t>(src + 32ll * j);
          const uint32_t sh = (uint32_t)(p & 3) * 8u;
          const uint32_t* a = reinterpret_cast<const uint32_t*>(p & ~(uintptr_t)3);
          const int need = (int)(p & 3) + min(nv, 32);
          uint32_t x[9];
#pragma unroll
          for (int k = 0; k < 9; k++) x[k] = 4 * k < need ? a[k] : 0u;
#pragma unroll
          for (int k = 0; k < 8; k++) w[k] = __funnelshift_r(x[k], x[k + 1], sh);
        }
        units[u0 + j] = pl_unit_from_words(w, nv);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// events of one unit (32 positions) on the shared-memory window state
// ---------------------------------------------------------------------------------------------------------------

// red.shared.add.u32 executed only where cond != 0 (a predicated instruction: no branch, no reconvergence point)
__device__ __forceinline__ void red_shared_if(uint32_t cond, uint32_t addr, uint32_t v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p red.shared.add.u32 [%1], %2;\n\t}" ::"r"(cond), "r"(addr), "r"(v) : "memory");
}
// highest set bit of x as a mask (0 for x == 0; PTX shifts clamp at 32)
__device__ __forceinline__ uint32_t top_bit(uint32_t x, int& idx) {
  uint32_t lz, b;
  asm("clz.b32 %0, %1;" : "=r"(lz) : "r"(x));
  asm("shr.u32 %0, %1, %2;" : "=r"(b) : "r"(0x80000000u), "r"(lz));
  idx = 31 - (int)lz;
  return b;
}

// calls of the requested contexts (slot order CHH, CHG, CG among the requested) of one unit
template <int WANT>
__device__ __forceinline__ void pl_want(const uint4& v, uint32_t (&m)[3]) {
  int slot = 0;
  m[0] = m[1] = m[2] = 0u;
  if (WANT & W_CHH) m[slot++] = cls_chh(v.x, v.y, v.z);
  if (WANT & W_CHG) m[slot++] = cls_chg(v.x, v.y, v.z);
  if (WANT & W_CG) m[slot++] = cls_cg(v.x, v.y, v.z);
}

// '-' bytes.  Isolated ones are counted in D (pl_events); the edges of longer runs become coverage deltas in W: -1 where a
// run starts, +1 right after it ends, clamped to the window start (pl_edges; `ed` = edge positions of the unit, E = its '-'
// positions, w_addr = shared BYTE address of the unit's position 0 in W).
__device__ __forceinline__ void pl_edges(uint32_t ed, uint32_t E, uint32_t w_addr, uint32_t sinc, uint32_t w_floor) {
  while (ed) {
    const int i = 31 - __clz(ed);
    const uint32_t bit = 1u << i;
    ed ^= bit;
    red_shared(max(w_addr + 4u * (uint32_t)i, w_floor), (E & bit) ? 0u - sinc : sinc);
  }
}
// Streaming form (the next unit is not at hand): a '-' at position 31 is treated as a run edge, it may continue.
__device__ __forceinline__ uint32_t pl_dash(uint32_t E, uint32_t Eprev, uint32_t w_addr, uint32_t sinc, uint32_t w_floor) {
  const uint32_t prev = __funnelshift_l(Eprev, E, 1);                    // "the position before me is a '-'"
  const uint32_t iso = E & ~prev & ~(E >> 1) & 0x7FFFFFFFu;
  pl_edges((E ^ prev) & ~(iso | (iso << 1)), E, w_addr, sinc, w_floor);
  return iso;
}
__device__ __forceinline__ uint32_t pl_dashes(const uint4& v) { return v.x & v.y & ~v.z & v.w; }

// m[k] = calls of requested context k, iso = isolated '-' positions, methnow = methylated calls to be counted right now
// (0 in the fused kernels).  d_addr / t_addr = shared BYTE addresses of the unit's position 0 in D and in T[0][strand].
#ifndef EA_PL_STEPS
#define EA_PL_STEPS 1
#endif
template <int NC>
__device__ __forceinline__ void pl_events(const uint32_t (&m)[3], uint32_t iso, uint32_t methnow, uint32_t d_addr, uint32_t t_addr,
                                          uint32_t ctx_stride, uint32_t sinc) {
  // two loops without address / value selects: calls, then isolated '-' bytes.  A lane rarely holds more than one or two
  // events of a kind per unit: the first EA_PL_STEPS are straight-line predicated code, the loop takes the rest.
  uint32_t ev = m[0] | m[1] | m[2];
#pragma unroll
  for (int k = 0; k < EA_PL_STEPS; k++) {
    int i;
    const uint32_t bit = top_bit(ev, i);
    ev ^= bit;
    uint32_t a = t_addr;
    if (NC > 1 && (m[1] & bit)) a = t_addr + ctx_stride;
    if (NC > 2 && (m[2] & bit)) a = t_addr + 2u * ctx_stride;
    red_shared_if(bit, a + 4u * (uint32_t)i, (methnow & bit) ? 0x10001u : 1u);
  }
  while (ev) {
    const int i = 31 - __clz(ev);
    const uint32_t bit = 1u << i;
    ev ^= bit;
    uint32_t a = t_addr;
    if (NC > 1 && (m[1] & bit)) a = t_addr + ctx_stride;
    if (NC > 2 && (m[2] & bit)) a = t_addr + 2u * ctx_stride;
    red_shared(a + 4u * (uint32_t)i, (methnow & bit) ? 0x10001u : 1u);
  }
#pragma unroll
  for (int k = 0; k < EA_PL_STEPS; k++) {
    int i;
    const uint32_t bit = top_bit(iso, i);
    iso ^= bit;
    red_shared_if(bit, d_addr + 4u * (uint32_t)i, sinc);
  }
  while (iso) {
    const int i = 31 - __clz(iso);
    iso ^= 1u << i;
    red_shared(d_addr + 4u * (uint32_t)i, sinc);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Finalize of one window, packed-compare form of window_finalize_rt (win_common.cuh): same state, same three barriers, same
// ordered compaction, leaves the state zeroed.  The row test `total_ctx > floor(cov/2)` (src/rcpp_cx_report.cpp:58-85) is
// `2*total > cov` and is evaluated for BOTH strands of a position at once on 16-bit fields: PRMT packs the two strands'
// totals, VIADDMNMX.U16x2 (max(2*total, coverage)) against the packed coverage differs from the coverage exactly in the fields where the test
// holds (totals and coverages are at most 32767: the dispatcher's precondition).  No divergent code before the row loop.
// flags: bit k = (position q0+k, '+'), bit 16+k = (position q0+k, '-');  PER <= 16.
// ---------------------------------------------------------------------------------------------------------------
template <int WANT, int PER, int NT, bool FUSED>
__device__ __forceinline__ void window_finalize_pk(uint32_t* __restrict__ W, uint32_t* __restrict__ D, uint32_t* __restrict__ T,
                                                   int S /* multiple of 4, <= PER * NT */, int w0, int win, const ReportOut& out,
                                                   uint32_t* s_scan /* 2 * NT/32 words */, long long* s_base, int* s_npass) {
  constexpr int NC = n_ctx<WANT>();
  constexpr int NW = NT / 32;
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  static_assert(PER % 4 == 0 && PER <= 16, "PER");
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
  {  // + the coverage of the warps in front: a shuffle scan over the NW per-warp totals instead of NW loads per thread
    const uint32_t v = lane < NW ? s_scan[lane] : 0u;
    uint32_t vi = v;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
      const uint32_t o = __shfl_up_sync(FULL, vi, d);
      if (lane >= d) vi += o;
    }
    ex += __shfl_sync(FULL, vi - v, warp);
  }

  uint32_t flags = 0;
#pragma unroll
  for (int c = 0; c < NC; c++)
#pragma unroll
    for (int v4 = 0; v4 < PER / 4; v4++) {
      uint4 tp = zero4, tm = zero4;
      if (q0 + v4 * 4 < S) {
        uint4* pp = reinterpret_cast<uint4*>(&T[(c * 2 + 0) * SG + kGuardLo + q0 + v4 * 4]);
        uint4* pm = reinterpret_cast<uint4*>(&T[(c * 2 + 1) * SG + kGuardLo + q0 + v4 * 4]);
        tp = *pp; tm = *pm;
      }
      const uint32_t a[4] = {tp.x, tp.y, tp.z, tp.w}, b[4] = {tm.x, tm.y, tm.z, tm.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int k = v4 * 4 + j;
        const uint32_t tot = __byte_perm(a[j], b[j], 0x5410);         // total '+' | total '-' << 16
        const uint32_t cv = cov[k] + ex;
        const uint32_t df = __viaddmax_u16x2(tot, tot, cv) ^ cv;      // max(2 * total, coverage): non-zero field <=> 2 * total > coverage
        flags += __vimin3_u16x2(df, 0x00010001u, 0x00010001u) << k;                // fields never collide: at most one context wins a position
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
  int my_off = rinc - mine, nrows_total;
  {
    const int v = lane < NW ? (int)s_scan[NW + lane] : 0;
    int vi = v;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
      const int o = __shfl_up_sync(FULL, vi, d);
      if (lane >= d) vi += o;
    }
    my_off += __shfl_sync(FULL, vi - v, warp);
    nrows_total = __shfl_sync(FULL, vi, NW - 1);
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
    long long di = wbase + my_off;
    uint32_t pos_bits = (flags | (flags >> 16)) & 0xFFFFu;
    while (pos_bits) {  // ascending position, '+' before '-'
      const int k = __ffs(pos_bits) - 1;
      pos_bits &= pos_bits - 1u;
#pragma unroll
      for (int s = 0; s < 2; s++) {
        if (!((flags >> (16 * s + k)) & 1u)) continue;
        int code = (WANT & W_CHH) ? 2 : ((WANT & W_CHG) ? 6 : 7);
        uint32_t val = T[s * SG + kGuardLo + q0 + k];
        if (NC > 1) {  // which of the requested contexts holds the majority
          uint32_t cv = 0;
#pragma unroll
          for (int kk = 0; kk < PER; kk++) cv = (kk == k) ? cov[kk] : cv;
          cv += ex;
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
        out.scratch[di++] = make_int4(w0 + q0 + k, (s + 1) | (code << 8), (int)(val >> 16), (int)((val & 0xFFFFu) - (val >> 16)));
      }
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

// ---------------------------------------------------------------------------------------------------------------
// Window kernel on planar tiles: one CTA per genomic window, a warp per tile, ONE LANE PER ROW.
//   MAXU > 0 : rows of at most MAXU units.  All units of a row are loaded up front and stay in registers (missing units
//              as zeros, so the code below is branch-free per unit): first the class counts of the whole row and its
//              pass flag (src/rcpp_threshold_reads.cpp:39-70), then ONE sweep of events in which a passing row's
//              methylated calls ride on the same shared-memory reduction as the call itself (total | methylated << 16).
//   MAXU == 0: any row length up to 1024 positions, single streaming sweep; the pass flag is known only at the row's end,
//              so passing rows read their in-window units a second time (L2 hits) to add their methylated calls
//              (SURVEY.md 8a invariant 4: only the methylated count depends on `pass`).
// The pass/fail chain runs for 32 rows per instruction through the exact integer tables of win_common.cuh (shared memory).
// Preconditions (dispatcher): rows <= 1024 positions, at most 32767 rows overlap a window, S a multiple of 32, <= PER*256.
// ---------------------------------------------------------------------------------------------------------------
#ifndef EA_PL_MINB
#define EA_PL_MINB 5
#endif

template <int TCTX, int WANT, bool FUSED, int MAXU, int PER>
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
  // the 32-bit shared address of the window state, taken ONCE through an opaque asm: left to itself the compiler
  // rematerialises the generic->shared conversion (S2R + LEA + ...) at every use instead of spending a register (3 % of
  // the kernel's instructions in round 2's first profile)
  uint32_t w_base = smem_u32(W);
  asm volatile("mov.u32 %0, %0;" : "+r"(w_base));
  const uint32_t d_base = w_base + 4u * (uint32_t)SG, t_base = w_base + 8u * (uint32_t)SG;
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
      if (overlaps) {
        red_shared(w_floor + 4u * (uint32_t)max(rel0, 0), sinc);
        const int e = rel0 + len;
        if (e < S) red_shared(w_floor + 4u * (uint32_t)e, 0u - sinc);
      }

#ifndef EA_PL_NO_PREFETCH
      if (t + kRowWarps < t_hi && lane < 2) {  // the warp's next tile: its unit rows (one bulk L2 prefetch) and its row metadata
        const long long nt0 = pl.toff[t + kRowWarps], nt1 = pl.toff[t + kRowWarps + 1];
        if (lane == 0 && nt1 > nt0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pl.units + nt0 * 32), "r"((uint32_t)(nt1 - nt0) * 512u) : "memory");
        if (lane == 1) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.start + (t + kRowWarps) * 32));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pl.ls + (t + kRowWarps) * 32));
        }
      }
#endif
      if (KEEP) {
        uint4 v[KU];
#pragma unroll
        for (int u = 0; u < KU; u++) {
          v[u] = make_uint4(0, 0, 0, 0);
          if (u < nu) v[u] = ldg_plunit(src + u * 32);
        }
        bool pass = valid && (pass_in == nullptr || pass_in[r] != 0);
        if (FUSED) {
          uint32_t n_in = 0, n_inm = 0, n_out = 0, n_outm = 0;
#pragma unroll
          for (int u = 0; u < KU; u++) {
            const uint32_t ci = cls_in<TCTX>(v[u].x, v[u].y, v[u].z), co = cls_out<TCTX>(v[u].x, v[u].y, v[u].z);
            n_in += __popc(ci); n_inm += __popc(ci & ~v[u].w);
            if (TCTX != TC_CX) { n_out += __popc(co); n_outm += __popc(co & ~v[u].w); }
          }
          // the reference's chain (src/rcpp_threshold_reads.cpp:39-70) without its short circuits: both table entries are
          // always inside the staged range, so all four conditions can be evaluated for 32 rows without a branch
          const uint32_t t_m = s_mmin[n_in], t_o = s_omax[n_out];
          pass = (n_inm != 0u) & (n_in >= thr.min_n) & (n_inm >= t_m) & ((n_outm == 0u) | (n_outm <= t_o));
          if (overlaps && rel0 >= 0) {  // the window that holds the row's start owns its flag
            pass_out[r] = pass ? 1 : 0;
            npass_w += pass;
          }
        }
        const uint32_t passm = pass ? 0xFFFFFFFFu : 0u;
        // '-' masks with exact neighbours across unit borders: E of the previous / this / the next unit are all in registers.
        // ALLIN: every row of the tile lies completely inside the window (the common case: only the tiles at a window's
        // edges hold rows that cross it) - no per-unit window test, no entry rule.
        auto sweep = [&](auto allin_t) {
          constexpr bool ALLIN = decltype(allin_t)::value;
          uint32_t Em1 = 0u, isom1 = 0u, E = pl_dashes(v[0]);
#pragma unroll
          for (int u = 0; u < KU; u++) {
            const uint32_t Enext = u + 1 < KU ? pl_dashes(v[u + 1 < KU ? u + 1 : u]) : 0u;
            const uint32_t prev = __funnelshift_l(Em1, E, 1), next = __funnelshift_r(E, Enext, 1);
            const uint32_t iso = E & ~prev & ~next;
            const uint32_t ed = (E ^ prev) & ~(iso | __funnelshift_l(isom1, iso, 1));
            const int rel = rel0 + 32 * u;
            if (ALLIN || (unsigned)(rel + 31) < (unsigned)(S + 31)) {  // -32 < rel < S (units of lanes without a row are zero: no events)
              const uint32_t off4 = 128u * (uint32_t)u;
              if (!ALLIN && u > 0 && rel <= 0 && ((Em1 & ~isom1) >> 31)) red_shared(w_floor, 0u - sinc);  // a '-' run is open where the row enters the window
              uint32_t m[3];
              pl_want<WANT>(v[u], m);
              pl_edges(ed, E, wq + off4, sinc, w_floor);
              pl_events<NC>(m, iso, (m[0] | m[1] | m[2]) & ~v[u].w & passm, dq + off4, tq + off4, ctx_stride, sinc);
            }
            Em1 = E; isom1 = iso; E = Enext;
          }
          return (Em1 & ~isom1) >> 31;   // the row's last position is the end of a '-' run
        };
#ifndef EA_PL_NO_ALLIN
        const bool allin = __all_sync(0xFFFFFFFFu, !overlaps || (rel0 >= 0 && rel0 + 32 * KU < S));
#else
        const bool allin = false;
#endif
        const uint32_t tail_run = allin ? sweep(std::true_type{}) : sweep(std::false_type{});
        if (overlaps && nu == KU && (len & 31) == 0) {  // a '-' run that reaches the end of a row filling all KU units
          if (tail_run && rel0 + len < S) red_shared(max(wq + 4u * (uint32_t)len, w_floor), sinc);
        }
      } else {
        // Streaming form, two sweeps over the row's units: (A, fused only) class counts of the whole row with four loads in
        // flight -> pass flag; (B) the events of the in-window units with the flag known (the second sweep's loads hit L2).
        bool pass = valid && (pass_in == nullptr || pass_in[r] != 0);
        if (FUSED) {
          uint32_t n_in = 0, n_inm = 0, n_out = 0, n_outm = 0;
          for (int u = 0; u < nu; u += 4) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
              v[k] = make_uint4(0, 0, 0, 0);
              if (u + k < nu) v[k] = ldg_plunit(src + (u + k) * 32);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const uint32_t ci = cls_in<TCTX>(v[k].x, v[k].y, v[k].z), co = cls_out<TCTX>(v[k].x, v[k].y, v[k].z);
              n_in += __popc(ci); n_inm += __popc(ci & ~v[k].w);
              if (TCTX != TC_CX) { n_out += __popc(co); n_outm += __popc(co & ~v[k].w); }
            }
          }
          __syncwarp();
          pass = overlaps && eval_pass_staged(n_inm, n_in, n_outm, n_out, thr, s_mmin, s_omax, (uint32_t)TBL);
          if (overlaps && rel0 >= 0) {  // the window that holds the row's start owns its flag
            pass_out[r] = pass ? 1 : 0;
            npass_w += pass;
          }
        }
        const uint32_t passm = pass ? 0xFFFFFFFFu : 0u;
        const int u_lo = rel0 <= -32 ? (-rel0) >> 5 : 0;           // first unit with rel > -32
        const int u_hi = min(nu, (S - rel0 + 31) >> 5);            // first unit with rel >= S
        const int u_first = u_lo > 0 ? u_lo - 1 : 0;               // one unit earlier: its last position decides whether a '-' run is open
        uint32_t Eprev = 0;
        uint4 cur = make_uint4(0, 0, 0, 0), nxt = make_uint4(0, 0, 0, 0);
        if (u_first < u_hi) cur = ldg_plunit(src + u_first * 32);
        for (int u = u_first; u < u_hi; u++) {
          if (u + 1 < u_hi) nxt = ldg_plunit(src + (u + 1) * 32);
          const uint4 v = cur;
          const uint32_t E = pl_dashes(v);
          if (u >= u_lo) {
            const uint32_t off4 = 128u * (uint32_t)u;
            if (u == u_lo && (Eprev >> 31)) red_shared(w_floor, 0u - sinc);  // a '-' run is open where the row enters the window
            uint32_t m[3];
            pl_want<WANT>(v, m);
            const uint32_t iso = pl_dash(E, Eprev, wq + off4, sinc, w_floor);
            pl_events<NC>(m, iso, (m[0] | m[1] | m[2]) & ~v.w & passm, dq + off4, tq + off4, ctx_stride, sinc);
          }
          Eprev = E;
          cur = nxt;
        }
        __syncwarp();
        if (overlaps && u_hi == nu) {   // the row ends inside the window: a '-' run that reaches a row end on a unit border
          const int e = rel0 + len;
          if ((len & 31) == 0 && (Eprev >> 31) && e < S) red_shared(max(wq + 4u * (uint32_t)len, w_floor), sinc);
        }
      }
    }
    if (FUSED) {
      for (int d = 16; d >= 1; d >>= 1) npass_w += __shfl_down_sync(0xFFFFFFFFu, npass_w, d);
      if (lane == 0 && npass_w) atomicAdd(&s_npass, npass_w);
    }
    __syncthreads();
    window_finalize_pk<WANT, PER, kRowThreads, FUSED>(W, D, T, S, w0, win, out, s_scan, &s_base, &s_npass);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Long-read window kernel on planar ROWS (tiles of one row): a WARP per row, rows dealt round robin, the lanes read 32
// consecutive units of the row per step (512 contiguous bytes, the next step's units requested before the current ones
// are processed) and push their unit's events.  Pass flags are an input (thresholding a long read needs the counts of
// the whole read: k_threshold_plr below; the reference advises threshold.reads=FALSE for long reads anyway,
// R/generateCytosineReport.R:97-98).  The '-' state crosses units through a shuffle (and a carried word across steps).
// Preconditions (dispatcher): at most 32767 rows overlap a window, S a multiple of 32 and <= PER * 256.
// ---------------------------------------------------------------------------------------------------------------
#ifndef EA_PLR_MINB
#define EA_PLR_MINB 5
#endif
template <int WANT, int PER>
__global__ void __launch_bounds__(kRowThreads, EA_PLR_MINB)
k_window_plr(Planar pl, WinTable wt, const uint8_t* __restrict__ pass_in, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int S = wt.S;
  const int SG = kGuardLo + S + kGuardHi;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_scan[2 * kRowWarps];
  __shared__ long long s_base;
  __shared__ int s_npass;  // unused (no fused thresholding); the finalize takes the pointer

  uint32_t* const W = reinterpret_cast<uint32_t*>(smem_raw);  // [SG]
  uint32_t* const D = W + SG;                                 // [SG]
  uint32_t* const T = D + SG;                                 // [NC][2][SG]
  const int kZero4 = (NC + 1) * 2 * SG / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t w_base = smem_u32(W), d_base = smem_u32(D), t_base = smem_u32(T);
  const uint32_t w_floor = w_base + 4u * kGuardLo;
  const uint32_t ctx_stride = 8u * (uint32_t)SG;

  for (int i = threadIdx.x; i < kZero4; i += kRowThreads) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) s_npass = 0;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    __syncthreads();
    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    for (int r = lo + warp; r < hi; r += kRowWarps) {
      const long long u0 = pl.toff[r];                          // warp-uniform loads
      const int start = pl.start[r];
      const uint32_t lsv = pl.ls[r];
      const int len = (int)(lsv & 0x7FFFFFFFu);
      const uint32_t strand0 = lsv >> 31;
      const int rel0 = start - w0;
      if (len <= 0 || rel0 + len <= 0) continue;                // rel0 < S by construction of hi
      const int nu = (len + 31) >> 5;
      const int u_lo = rel0 <= -32 ? (-rel0) >> 5 : 0;          // first unit with rel > -32
      const int u_hi = min(nu, (S - rel0 + 31) >> 5);           // first unit with rel >= S
      const int u_first = u_lo > 0 ? u_lo - 1 : 0;              // one unit earlier: is a '-' run open where the row enters the window?
      const uint32_t sinc = strand0 ? 0x10000u : 1u;
      if (lane == 0) {  // the row's coverage deltas
        red_shared(w_floor + 4u * (uint32_t)max(rel0, 0), sinc);
        const int e = rel0 + len;
        if (e < S) red_shared(w_floor + 4u * (uint32_t)e, 0u - sinc);
      }
      const uint32_t passm = (pass_in == nullptr || pass_in[r] != 0) ? 0xFFFFFFFFu : 0u;
      const uint4* src = pl.units + u0;
      const uint32_t t_row = t_base + 4u * strand0 * (uint32_t)SG;
      uint32_t Ecarry = 0u;
      int u = u_first + lane;
      uint4 v = make_uint4(0, 0, 0, 0), nv = make_uint4(0, 0, 0, 0);
      if (u < u_hi) v = ldg_plunit(src + u);
      for (int base = u_first; base < u_hi; base += 32) {        // warp-uniform trip count
        const int un = u + 32;
        if (un < u_hi) nv = ldg_plunit(src + un);
        const uint32_t E = u < u_hi ? pl_dashes(v) : 0u;
        uint32_t Eprev = __shfl_up_sync(FULL, E, 1);
        if (lane == 0) Eprev = Ecarry;
        Ecarry = __shfl_sync(FULL, E, 31);
        if (u >= u_lo && u < u_hi) {
          const uint32_t rel4 = 4u * (uint32_t)(rel0 + 32 * u + kGuardLo);   // >= 4: rel0 + 32u > -32
          if (u == u_lo && (Eprev >> 31)) red_shared(w_floor, 0u - sinc);    // a '-' run is open where the row enters the window
          uint32_t m[3];
          pl_want<WANT>(v, m);
          const uint32_t iso = pl_dash(E, Eprev, w_base + rel4, sinc, w_floor);
          pl_events<NC>(m, iso, (m[0] | m[1] | m[2]) & ~v.w & passm, d_base + rel4, t_row + rel4, ctx_stride, sinc);
          if (u == nu - 1 && (len & 31) == 0 && (E >> 31) && rel0 + len < S)  // a '-' run that reaches the row end on a unit border
            red_shared(max(w_base + 4u * (uint32_t)(rel0 + len + kGuardLo), w_floor), sinc);
        }
        __syncwarp();
        v = nv; u = un;
      }
    }
    __syncthreads();
    window_finalize_pk<WANT, PER, kRowThreads, false>(W, D, T, S, w0, win, out, s_scan, &s_base, &s_npass);
  }
}

// Stand-alone per-read kernel on planar ROWS (long reads): a warp per row, four units in flight per lane, class counts
// reduced with shuffles; the pass/fail chain through the staged tables where the counts fit them, else with the
// reference's IEEE double divisions (win_common.cuh:eval_pass_staged).
template <int TCTX>
__global__ void __launch_bounds__(256)
k_threshold_plr(Planar pl, ThrParams thr, ThrTables tbl, uint8_t* __restrict__ pass_out, double* __restrict__ beta_out,
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
  for (int64_t r = warp; r < pl.n; r += n_warps) {
    const long long u0 = pl.toff[r];
    const int nu = (int)(((pl.ls[r] & 0x7FFFFFFFu) + 31u) >> 5);
    const uint4* src = pl.units + u0;
    uint32_t n_in = 0, n_inm = 0, n_out = 0, n_outm = 0;
    for (int u = lane; u < nu; u += 128) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        v[k] = make_uint4(0, 0, 0, 0);
        if (u + 32 * k < nu) v[k] = ldg_plunit(src + u + 32 * k);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t ci = cls_in<TCTX>(v[k].x, v[k].y, v[k].z), co = cls_out<TCTX>(v[k].x, v[k].y, v[k].z);
        n_in += __popc(ci); n_inm += __popc(ci & ~v[k].w);
        if (TCTX != TC_CX) { n_out += __popc(co); n_outm += __popc(co & ~v[k].w); }
      }
    }
    __syncwarp();
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      n_in += __shfl_xor_sync(0xFFFFFFFFu, n_in, d); n_inm += __shfl_xor_sync(0xFFFFFFFFu, n_inm, d);
      n_out += __shfl_xor_sync(0xFFFFFFFFu, n_out, d); n_outm += __shfl_xor_sync(0xFFFFFFFFu, n_outm, d);
    }
    if (lane == 0) {
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
