// Window kernel with VERTICAL (bit-sliced) counters: the aggregation of rcpp_cx_report (src/rcpp_cx_report.cpp:117-128 and
// 58-85) without a single atomic operation and without per-position state in shared memory.
//
// The event kernels (planar.cuh: k_window_pl, win_common.cuh) turn every call into a shared-memory reduction.  Their cost
// grows with the number of calls (CX reports: ~10 % of all positions) and same-address reductions serialise on deep data
// (amplicons / capture islands: 1000 reads over one position).  Here the work is turned around:
//
//   * a window of S = 32 * NB positions is cut into NB aligned BLOCKS of 32 positions; a lane owns one (block, strand)
//     pair (LPB = 256 / NB lanes per block: 2 strands x GP row subgroups);
//   * the window's rows arrive in BATCHES of whole tiles (planar.cuh layout) staged in shared memory.  Phase A, one lane
//     per row: class counts of the whole row -> the reference's pass/fail chain (src/rcpp_threshold_reads.cpp:39-70) ->
//     a 16-byte record {window-relative start, length | pass << 31, unit base, units} appended to the record list of the
//     row's strand (ordered: warp ballots + one block-wide exclusive scan);
//   * phase B, one lane per (block, strand, subgroup): binary search for the records that can reach the block, then for
//     every such row the 32-position slice of its four bit planes aligned to the block (two 128-bit shared loads + four
//     funnel shifts).  `not '-'`, `call of context c` and `methylated call of a passing row` are one or two LOP3 each over
//     the 32 positions, and each of these 32-bit words is ADDED INTO A BIT-SLICED COUNTER held in registers with
//     carry-save adders (a full adder over 32 positions = two LOP3): about 4 logic instructions per row, block and counter,
//     whatever the depth and however many calls there are;
//   * finalize: the GP partial counters of a (block, strand) are added with shuffles, the row test
//     2 * total > coverage is a bit-sliced comparator (two LOP3 per plane for 32 positions), and only the few positions
//     that are report rows are un-sliced.  Coverage is counted directly (a position is covered by a row iff its code is
//     not 0 and not 11), so there are no coverage deltas, no prefix sums and nothing to zero.
//
// K = planes per counter = bits of the largest number of rows that overlap one window (dispatcher).
#pragma once
#include "planar.cuh"

namespace ea {

constexpr int kVsThreads = 256;
constexpr int kVsWarps = kVsThreads / 32;
constexpr int kVsBatchTiles = 8;                       // <= 256 rows per batch: one row per thread in phase A
constexpr int kVsCapUnitRows = 40;                     // staged unit rows per batch (x 512 bytes)
constexpr int kVsRecCap = 32 * kVsBatchTiles;

#ifndef EA_VS_MINB
#define EA_VS_MINB 3
#endif

// full adder over 32 positions: (a, b, c) -> sum in a, carry returned
__device__ __forceinline__ uint32_t vs_fa(uint32_t& a, uint32_t b, uint32_t c) {
  const uint32_t carry = (a & b) | (c & (a | b));   // majority: one LOP3
  a = a ^ b ^ c;                                    // one LOP3
  return carry;
}

// K-plane bit-sliced counter of 32 positions
template <int K>
struct VsCounter {
  uint32_t p[K];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < K; i++) p[i] = 0u;
  }
  // add four one-bit-per-position words (carry-save tree: two full adders on plane 0, one on plane 1, ripple from plane 2)
  __device__ __forceinline__ void add4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t t0 = vs_fa(p[0], a, b);
    const uint32_t t1 = vs_fa(p[0], c, d);
    uint32_t f = K > 1 ? vs_fa(p[1], t0, t1) : 0u;
#pragma unroll
    for (int i = 2; i < K; i++) {
      const uint32_t carry = p[i] & f;
      p[i] ^= f;
      f = carry;
    }
  }
  // this += other (both K planes; the sum fits K planes by the dispatcher's choice of K)
  __device__ __forceinline__ void add(const VsCounter<K>& o) {
    uint32_t carry = 0u;
#pragma unroll
    for (int i = 0; i < K; i++) carry = vs_fa(p[i], o.p[i], carry);
  }
  __device__ __forceinline__ uint32_t value_at(int pos) const {
    uint32_t v = 0u;
#pragma unroll
    for (int i = 0; i < K; i++) v |= ((p[i] >> pos) & 1u) << i;
    return v;
  }
};

// positions where 2 * tot > cov (bit-sliced comparator, most significant plane first)
template <int K>
__device__ __forceinline__ uint32_t vs_majority(const VsCounter<K>& tot, const VsCounter<K>& cov) {
  uint32_t gt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
  for (int i = K; i >= 0; i--) {
    const uint32_t a = i >= 1 ? tot.p[i >= 1 ? i - 1 : 0] : 0u;   // plane i of 2 * tot
    const uint32_t b = i < K ? cov.p[i < K ? i : 0] : 0u;
    gt |= eq & a & ~b;
    eq &= ~(a ^ b);
  }
  return gt;
}

struct VsRec { int32_t rel0; uint32_t lp; uint32_t ubase; uint32_t nu; };   // lp = length | pass << 31

template <int TCTX, int WANT, bool FUSED, int K, int LOGNB>
__global__ void __launch_bounds__(kVsThreads, EA_VS_MINB)
k_window_vs(Planar pl, WinTable wt, ThrParams thr, ThrTables tbl, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
            ReportOut out, int max_len) {
  constexpr int NC = n_ctx<WANT>();
  constexpr int NB = 1 << LOGNB;              // blocks per window
  constexpr int LPB = kVsThreads / NB;        // lanes per block
  constexpr int GP = LPB / 2;                 // row subgroups per (block, strand)
  static_assert(GP >= 1 && LPB <= 32, "window size");
  constexpr uint32_t FULL = 0xFFFFFFFFu;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4* const buf = reinterpret_cast<uint4*>(smem_raw);                                   // [kVsCapUnitRows * 32] staged units
  VsRec* const rec = reinterpret_cast<VsRec*>(smem_raw + (size_t)kVsCapUnitRows * 512);   // [2][kVsRecCap]
  __shared__ int s_wcnt[kVsWarps][2];
  __shared__ int s_cnt[NB + 1];         // rows per block -> exclusive prefix
  __shared__ long long s_base;
  __shared__ int s_npass;
  __shared__ uint16_t s_mmin[FUSED ? kThrTable : 1], s_omax[FUSED ? kThrTable : 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int blk = tid / LPB;                  // block of the window this lane works for
  const int strand = (tid / GP) & 1;
  const int sub = tid % GP;
  const int pb = 32 * blk;

  if (FUSED) stage_thr_tables(tbl, s_mmin, s_omax, kThrTable - 1);
  if (tid == 0) s_npass = 0;

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    VsCounter<K> c_cov, c_tot[NC], c_meth[NC];
    c_cov.clear();
#pragma unroll
    for (int c = 0; c < NC; c++) { c_tot[c].clear(); c_meth[c].clear(); }
    int npass_w = 0;

    const int t_end = (hi + 31) >> 5;
    int tb = lo >> 5;
    while (tb < t_end) {
      // ---- batch = tiles [tb, te): at most kVsBatchTiles tiles and kVsCapUnitRows unit rows ----
      // (lane k looks at tile tb + k: one coalesced load of the tile table and a ballot instead of a serial scan)
      const long long u_tb = pl.toff[tb];
      long long u_k = u_tb;
      if (lane <= kVsBatchTiles && tb + lane <= t_end) u_k = pl.toff[tb + lane];
      const uint32_t fits = __ballot_sync(FULL, lane >= 1 && lane <= kVsBatchTiles && tb + lane <= t_end && u_k - u_tb <= kVsCapUnitRows);
      const int te = tb + max(1, __popc(fits));       // tiles are monotone: the fitting ones are a prefix (a tile never exceeds the capacity)
      const int n_units = (int)(__shfl_sync(FULL, u_k, te - tb) - u_tb) * 32;
      __syncthreads();  // the previous batch (or window) is done with the buffers
      {
        const uint4* src = pl.units + u_tb * 32;
        for (int i = tid; i < n_units; i += kVsThreads) buf[i] = ldg_plunit(src + i);
      }
      // ---- phase A: one lane per row ----
      const int tl = tid >> 5;                       // tile of the batch
      const int r = (tb + tl) * 32 + lane;
      const bool valid = tb + tl < te && r >= lo && r < hi;
      int start = 0;
      uint32_t lsv = 0;
      if (valid) { start = pl.start[r]; lsv = pl.ls[r]; }
      const int len = (int)(lsv & 0x7FFFFFFFu);
      const int rstrand = (int)(lsv >> 31);
      const int rel0 = start - w0;
      const bool overlaps = valid && len > 0 && rel0 + len > 0;  // rel0 < S by construction of hi
      const int nu = (len + 31) >> 5;
      uint32_t ubase = 0;
      if (tb + tl < te) ubase = (uint32_t)(pl.toff[tb + tl] - u_tb) * 32u + (uint32_t)lane;
      bool pass = valid && (pass_in == nullptr || pass_in[r] != 0);
      __syncthreads();  // staged units visible
      if (FUSED) {
        uint32_t n_in = 0, n_inm = 0, n_out = 0, n_outm = 0;
        if (overlaps) {
          for (int u = 0; u < nu; u++) {
            const uint4 v = buf[ubase + 32u * (uint32_t)u];
            const uint32_t ci = cls_in<TCTX>(v.x, v.y, v.z), co = cls_out<TCTX>(v.x, v.y, v.z);
            n_in += __popc(ci); n_inm += __popc(ci & ~v.w);
            if (TCTX != TC_CX) { n_out += __popc(co); n_outm += __popc(co & ~v.w); }
          }
        }
        pass = overlaps && eval_pass_staged(n_inm, n_in, n_outm, n_out, thr, s_mmin, s_omax, kThrTable - 1);
        if (overlaps && rel0 >= 0) {  // the window that holds the row's start owns its flag
          pass_out[r] = pass ? 1 : 0;
          npass_w += pass;
        }
      }
      // ordered record lists per strand
      const uint32_t m0 = __ballot_sync(FULL, overlaps && rstrand == 0), m1 = __ballot_sync(FULL, overlaps && rstrand == 1);
      if (lane == 0) { s_wcnt[warp][0] = __popc(m0); s_wcnt[warp][1] = __popc(m1); }
      __syncthreads();
      int base = 0, n_rec[2] = {0, 0};
#pragma unroll
      for (int w = 0; w < kVsWarps; w++) {
        const int c0 = s_wcnt[w][0], c1 = s_wcnt[w][1];
        n_rec[0] += c0; n_rec[1] += c1;
        if (w < warp) base += rstrand ? c1 : c0;
      }
      if (overlaps) {
        const uint32_t lt = (1u << lane) - 1u;
        const int rank = base + __popc((rstrand ? m1 : m0) & lt);
        VsRec rc;
        rc.rel0 = rel0; rc.lp = (uint32_t)len | (pass ? 0x80000000u : 0u); rc.ubase = ubase; rc.nu = (uint32_t)nu;
        *reinterpret_cast<uint4*>(&rec[rstrand * kVsRecCap + rank]) = *reinterpret_cast<const uint4*>(&rc);
      }
      __syncthreads();

      // ---- phase B: one lane per (block, strand, subgroup) ----
      {
        const VsRec* lst = rec + strand * kVsRecCap;
        const int n = n_rec[strand];
        // first record that starts behind the block / first record that can reach it (records are sorted by rel0)
        int a = 0, b = n;
        while (a < b) { const int m = (a + b) >> 1; if (lst[m].rel0 < pb + 32) a = m + 1; else b = m; }
        const int i_hi = a;
        a = 0; b = i_hi;
        while (a < b) { const int m = (a + b) >> 1; if (lst[m].rel0 + max_len <= pb) a = m + 1; else b = m; }
        for (int i = a + sub; i < i_hi; i += 4 * GP) {
          uint32_t x_cov[4], x_tot[NC][4], x_meth[NC][4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            x_cov[q] = 0u;
#pragma unroll
            for (int c = 0; c < NC; c++) { x_tot[c][q] = 0u; x_meth[c][q] = 0u; }
            const int idx = i + q * GP;
            if (idx < i_hi) {
              const uint4 rv = *reinterpret_cast<const uint4*>(&lst[idx]);
              const int o = pb - (int)rv.x;                      // position within the row of the block's first position (> -32)
              const int L = (int)(rv.y & 0x7FFFFFFFu);
              if (o < L) {
                const int j = o >> 5;
                const uint32_t sh = (uint32_t)o & 31u;
                uint4 lo4 = make_uint4(0, 0, 0, 0), hi4 = make_uint4(0, 0, 0, 0);
                if (j >= 0) lo4 = buf[rv.z + 32u * (uint32_t)j];
                if (j + 1 < (int)rv.w) hi4 = buf[rv.z + 32u * (uint32_t)(j + 1)];
                uint4 v;
                v.x = __funnelshift_r(lo4.x, hi4.x, sh); v.y = __funnelshift_r(lo4.y, hi4.y, sh);
                v.z = __funnelshift_r(lo4.z, hi4.z, sh); v.w = __funnelshift_r(lo4.w, hi4.w, sh);
                x_cov[q] = v.z | (v.y & ~(v.x & v.w));           // code neither 0 (outside the row) nor 11 ('-')
                uint32_t m[3];
                pl_want<WANT>(v, m);
                const uint32_t pm = (rv.y >> 31) ? ~v.w : 0u;     // methylated calls count for passing rows only
#pragma unroll
                for (int c = 0; c < NC; c++) { x_tot[c][q] = m[c]; x_meth[c][q] = m[c] & pm; }
              }
            }
          }
          c_cov.add4(x_cov[0], x_cov[1], x_cov[2], x_cov[3]);
#pragma unroll
          for (int c = 0; c < NC; c++) {
            c_tot[c].add4(x_tot[c][0], x_tot[c][1], x_tot[c][2], x_tot[c][3]);
            c_meth[c].add4(x_meth[c][0], x_meth[c][1], x_meth[c][2], x_meth[c][3]);
          }
        }
      }
      tb = te;
    }

    // ---- finalize: add the subgroups' counters, row test, ordered compaction ----
#pragma unroll
    for (int d = 1; d < GP; d <<= 1) {
      VsCounter<K> o;
#pragma unroll
      for (int i = 0; i < K; i++) o.p[i] = __shfl_xor_sync(FULL, c_cov.p[i], d);
      c_cov.add(o);
#pragma unroll
      for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int i = 0; i < K; i++) o.p[i] = __shfl_xor_sync(FULL, c_tot[c].p[i], d);
        c_tot[c].add(o);
#pragma unroll
        for (int i = 0; i < K; i++) o.p[i] = __shfl_xor_sync(FULL, c_meth[c].p[i], d);
        c_meth[c].add(o);
      }
    }
    uint32_t win_mask[NC];
    uint32_t rows_mask = 0u;
#pragma unroll
    for (int c = 0; c < NC; c++) { win_mask[c] = vs_majority<K>(c_tot[c], c_cov); rows_mask |= win_mask[c]; }
    const uint32_t other = __shfl_xor_sync(FULL, rows_mask, GP);   // the same block's other strand
    if (FUSED) {
      for (int d = 16; d >= 1; d >>= 1) npass_w += __shfl_down_sync(FULL, npass_w, d);
      if (lane == 0 && npass_w) atomicAdd(&s_npass, npass_w);
    }
    if (sub == 0 && strand == 0) s_cnt[blk] = __popc(rows_mask) + __popc(other);
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the NB block counts (NB <= 64: two per lane)
      constexpr int PERL = (NB + 31) / 32;
      int v[PERL], sum = 0;
#pragma unroll
      for (int k = 0; k < PERL; k++) { v[k] = lane * PERL + k < NB ? s_cnt[lane * PERL + k] : 0; sum += v[k]; }
      int inc = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += o;
      }
      int ex = inc - sum;
      const int total = __shfl_sync(FULL, inc, 31);
#pragma unroll
      for (int k = 0; k < PERL; k++) { if (lane * PERL + k < NB) s_cnt[lane * PERL + k] = ex; ex += v[k]; }
      if (lane == 0) {
        s_cnt[NB] = total;
        const unsigned long long bb = total ? atomicAdd(out.cursor, (unsigned long long)total) : 0ull;
        s_base = (long long)bb;
        out.w_base[win] = (long long)bb;
        out.w_cnt[win] = total;
        if (FUSED) {
          if (s_npass) atomicAdd(out.n_pass, (unsigned long long)s_npass);
          s_npass = 0;
        }
      }
    }
    __syncthreads();
    const long long wbase = s_base;
    const int nrows_total = s_cnt[NB];
    if (rows_mask && wbase + nrows_total <= out.cap) {
      int4* dst = out.scratch + wbase + s_cnt[blk];
      uint32_t mm = rows_mask;
      for (int nth = 0; mm; nth++) {  // rows of a block in (position, strand) order: '+' before '-'; the GP lanes that hold the sums take turns
        const int p = __ffs(mm) - 1;
        mm &= mm - 1u;
        if ((nth & (GP - 1)) != sub) continue;
        const uint32_t below = (1u << p) - 1u;
        const int rank = __popc(rows_mask & below) + __popc(other & below) + (strand ? (int)((other >> p) & 1u) : 0);
        int code = (WANT & W_CHH) ? 2 : ((WANT & W_CHG) ? 6 : 7);
        uint32_t tot = 0u, meth = 0u;
        int slot = 0;
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
          if (!((WANT >> cc) & 1)) continue;
          if ((win_mask[slot] >> p) & 1u) {
            code = cc == 0 ? 2 : (cc == 1 ? 6 : 7);
            tot = c_tot[slot].value_at(p);
            meth = c_meth[slot].value_at(p);
          }
          slot++;
        }
        dst[rank] = make_int4(w0 + pb + p, (strand + 1) | (code << 8), (int)meth, (int)(tot - meth));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Long reads (planar rows, tiles of one row): a WARP per window of 1024 positions, lane b owns block b of the window.
// Every row that crosses the window is warp-uniform work: its metadata are uniform loads, lane b reads the unit that
// holds its block's first position (the 32 lanes read 512 contiguous bytes), takes the next unit from its neighbour with
// a shuffle, aligns the four planes with funnel shifts and adds `not '-'`, `call` and `methylated call of a passing row`
// into the bit-sliced counters of the row's strand (a warp-uniform branch).  No shared memory, no barriers, no atomics
// but the window's claim on the scratch rows; coverage is counted directly, so nothing is proportional to the number of
// POSITIONS except one comparator pass per block - on 10-30 kb reads at a few-fold depth the per-position finalize of the
// event kernels is 40 % of their instructions.  Pass flags are an input (k_threshold_plr).
// K planes per counter: at most 2^K - 1 rows may overlap a window (dispatcher).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kVsrS = 1024;
#ifndef EA_VSR_MINB
#define EA_VSR_MINB 3
#endif

template <int K>
__device__ __forceinline__ void vs_add1(VsCounter<K>& c, uint32_t x) {
#pragma unroll
  for (int i = 0; i < K; i++) {
    const uint32_t carry = c.p[i] & x;
    c.p[i] ^= x;
    x = carry;
  }
}

template <int WANT, int K>
struct VsrStrand {
  VsCounter<K> cov, tot[n_ctx<WANT>()], meth[n_ctx<WANT>()];
  __device__ __forceinline__ void clear() {
    cov.clear();
#pragma unroll
    for (int c = 0; c < n_ctx<WANT>(); c++) { tot[c].clear(); meth[c].clear(); }
  }
  __device__ __forceinline__ void add(const uint4& v, uint32_t passm) {
    vs_add1<K>(cov, v.z | (v.y & ~(v.x & v.w)));   // code neither 0 (outside the row) nor 11 ('-')
    uint32_t m[3];
    pl_want<WANT>(v, m);
#pragma unroll
    for (int c = 0; c < n_ctx<WANT>(); c++) { vs_add1<K>(tot[c], m[c]); vs_add1<K>(meth[c], m[c] & ~v.w & passm); }
  }
};

template <int WANT, int K>
__global__ void __launch_bounds__(kVsThreads, EA_VSR_MINB)
k_window_vsr(Planar pl, WinTable wt, const uint8_t* __restrict__ pass_in, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  constexpr uint32_t FULL = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int pb = 32 * lane;
  const int gw = (int)((blockIdx.x * (unsigned)kVsThreads + threadIdx.x) >> 5), n_gw = (int)((gridDim.x * (unsigned)kVsThreads) >> 5);
  for (int win = gw; win < wt.n; win += n_gw) {
    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];
    VsrStrand<WANT, K> sp, sm;   // '+' and '-' strand
    sp.clear(); sm.clear();
    // Software pipeline over the window's rows: the metadata of row r+2 and the units of row r+1 are requested before row r
    // is added up (a window holds few rows, each costing a dependent metadata -> unit round trip; with three CTAs of
    // registers per SM there are not enough warps to hide it otherwise).
    struct Meta { long long u0; int start; uint32_t lsv; uint32_t passm; };
    auto load_meta = [&](int r) {
      Meta m;
      m.u0 = 0; m.start = 0; m.lsv = 0u; m.passm = 0u;
      if (r < hi) {
        m.u0 = pl.toff[r]; m.start = pl.start[r]; m.lsv = pl.ls[r];
        m.passm = (pass_in == nullptr || pass_in[r] != 0) ? 0xFFFFFFFFu : 0u;
      }
      return m;
    };
    struct Units { uint4 lo4, hi31; uint32_t sh; bool live; uint32_t strand, passm; };
    auto issue = [&](const Meta& m) {   // request the units of a row: lane b the one that holds its block's first position
      Units u;
      u.lo4 = make_uint4(0, 0, 0, 0); u.hi31 = make_uint4(0, 0, 0, 0);
      const int len = (int)(m.lsv & 0x7FFFFFFFu);
      const int rel0 = m.start - w0;
      u.live = len > 0 && rel0 + len > 0;                     // warp-uniform; rel0 < S by construction of hi
      u.strand = m.lsv >> 31; u.passm = m.passm;
      const int nu = (len + 31) >> 5;
      const int o = pb - rel0;                                // position within the row of this lane's first position
      const int j = o >> 5;                                   // (floor) unit that holds it; lane b+1 looks at unit j+1
      u.sh = (uint32_t)o & 31u;                               // the same in every lane
      if (u.live) {
        if (j >= 0 && j < nu) u.lo4 = ldg_plunit(pl.units + m.u0 + j);
        if (lane == 31 && u.sh != 0u && j + 1 >= 0 && j + 1 < nu) u.hi31 = ldg_plunit(pl.units + m.u0 + j + 1);
      }
      return u;
    };
    Meta m1 = load_meta(lo), m2 = load_meta(lo + 1);
    Units cur = issue(m1);
    for (int r = lo; r < hi; r++) {
      const Meta m3 = load_meta(r + 2);
      const Units nxt = issue(m2);
      if (cur.live) {
        uint4 hi4;
        hi4.x = __shfl_down_sync(FULL, cur.lo4.x, 1); hi4.y = __shfl_down_sync(FULL, cur.lo4.y, 1);
        hi4.z = __shfl_down_sync(FULL, cur.lo4.z, 1); hi4.w = __shfl_down_sync(FULL, cur.lo4.w, 1);
        if (lane == 31) hi4 = cur.hi31;
        uint4 v;
        v.x = __funnelshift_r(cur.lo4.x, hi4.x, cur.sh); v.y = __funnelshift_r(cur.lo4.y, hi4.y, cur.sh);
        v.z = __funnelshift_r(cur.lo4.z, hi4.z, cur.sh); v.w = __funnelshift_r(cur.lo4.w, hi4.w, cur.sh);
        if (cur.strand) sm.add(v, cur.passm); else sp.add(v, cur.passm);   // warp-uniform branch
      }
      cur = nxt; m2 = m3;
    }
    // ---- rows of the window: bit-sliced majority test, ordered compaction within the warp ----
    uint32_t win_p[NC], win_m[NC], rows_p = 0u, rows_m = 0u;
#pragma unroll
    for (int c = 0; c < NC; c++) {
      win_p[c] = vs_majority<K>(sp.tot[c], sp.cov); rows_p |= win_p[c];
      win_m[c] = vs_majority<K>(sm.tot[c], sm.cov); rows_m |= win_m[c];
    }
    const int mine = __popc(rows_p) + __popc(rows_m);
    int inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += o;
    }
    const int total = __shfl_sync(FULL, inc, 31);
    unsigned long long bb = 0ull;
    if (lane == 0) {
      bb = total ? atomicAdd(out.cursor, (unsigned long long)total) : 0ull;
      out.w_base[win] = (long long)bb;
      out.w_cnt[win] = total;
    }
    bb = __shfl_sync(FULL, bb, 0);
    if (mine && (long long)bb + total <= out.cap) {
      int4* dst = out.scratch + bb + (inc - mine);
      uint32_t both = rows_p | rows_m;
      while (both) {  // ascending position, '+' before '-'
        const int p = __ffs(both) - 1;
        both &= both - 1u;
#pragma unroll
        for (int s = 0; s < 2; s++) {
          if (!(((s ? rows_m : rows_p) >> p) & 1u)) continue;
          int code = (WANT & W_CHH) ? 2 : ((WANT & W_CHG) ? 6 : 7);
          uint32_t tot = 0u, meth = 0u;
          int slot = 0;
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            if (!((WANT >> cc) & 1)) continue;
            if (((s ? win_m[slot] : win_p[slot]) >> p) & 1u) {
              code = cc == 0 ? 2 : (cc == 1 ? 6 : 7);
              tot = s ? sm.tot[slot].value_at(p) : sp.tot[slot].value_at(p);
              meth = s ? sm.meth[slot].value_at(p) : sp.meth[slot].value_at(p);
            }
            slot++;
          }
          *dst++ = make_int4(w0 + pb + p, (s + 1) | (code << 8), (int)meth, (int)(tot - meth));
        }
      }
    }
  }
}

}  // namespace ea
