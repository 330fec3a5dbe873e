// Short-read window kernel: one THREAD per read over TMA-staged tiles.
//
// Why a second formulation next to k_window (kernels.cuh): with ~150-byte reads the lane-group
// kernel spends most of its issue slots on per-row overheads that are paid per warp instruction
// (row metadata, cross-lane reductions, pass evaluation) and on a divergent event loop.  Here
//   * each warp bulk-copies a tile of up to 32 consecutive padded rows (one contiguous, 16-byte
//     aligned byte range of the CSR) into shared memory with ONE cp.async.bulk (TMA) completed on a
//     warp-private mbarrier, then every lane walks its own row with LDS.128: no shuffles, row
//     metadata loads are coalesced over 32 rows, the IEEE-double pass chain runs for 32 rows at once;
//   * only cytosine calls of the requested contexts become shared-memory atomics ({total, methylated
//     of passing reads} packed in one 32-bit word per (context, strand, position));
//   * coverage is NOT accumulated: by SURVEY.md 8a invariant 3 a row exists iff total_ctx >
//     floor(cov/2), so cov is only needed where total_ctx > 0.  Those few candidates (~2% of the
//     positions for a CpG report) are compacted in position order and every thread PULLS the
//     coverage of one candidate from the rows that cover it (rows are sorted, so they are a short
//     contiguous range), counting bytes whose code is not 11 ('-' and the 0x0B padding).
// Preconditions (checked by the host dispatcher): every padded row fits a tile, at most 65535
// candidate rows per window, moderate depth (the pull walks depth x max_len/avg_len rows).
#pragma once
#include "kernels.cuh"

namespace ea {

constexpr int kTprThreads = 256;
constexpr int kTprWarps = kTprThreads / 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(phase)
      : "memory");
}
// 1-D TMA: global -> shared bulk copy, bytes and both addresses multiples of 16
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NWARPS>
__device__ __forceinline__ int block_exscan_n(int v, int* total, int* s_warp /* NWARPS+1 ints */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int x = lane < NWARPS ? s_warp[lane] : 0;
    int xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, xi, d);
      if (lane >= d) xi += o;
    }
    if (lane < NWARPS) s_warp[lane] = xi - x;
    if (lane == NWARPS - 1) s_warp[NWARPS] = xi;
  }
  __syncthreads();
  *total = s_warp[NWARPS];
  return s_warp[warp] + inc - v;
}

// push the requested-context calls of one chunk; rel = window-relative position of the chunk's byte 0
template <int WANT>
__device__ __forceinline__ void push_calls(const Planes& p, uint32_t* __restrict__ tm, int S, int strand0, int rel, bool pass) {
  uint32_t cand = cls_want<WANT>(p.A0, p.B0, p.C0) | cls_want<WANT>(p.A1, p.B1, p.C1);
  while (cand) {
    const int i = __ffs(cand) - 1;
    cand &= cand - 1u;
    const int q = rel + ((i & 1) << 3) + (i & 4) + (i >> 3);
    if ((unsigned)q < (unsigned)S) {
      const uint32_t code = (((i & 1) ? p.P1 : p.P0) >> (i & 28)) & 15u;
      const int slot = n_ctx<WANT>() == 1 ? 0 : slot_of<WANT>(code & 7u);
      atomicAdd(&tm[(slot * 2 + strand0) * S + q], (pass && code < 8u) ? 0x10001u : 1u);
    }
  }
}

template <int TCTX, int WANT, bool FUSED>
__global__ void __launch_bounds__(kTprThreads, 2)
k_window_tpr(Csr csr, WinTable wt, ThrParams thr, const uint8_t* __restrict__ pass_in, uint8_t* __restrict__ pass_out,
             int tile_bytes, int max_span, ReportOut out) {
  constexpr int NC = n_ctx<WANT>();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[kTprWarps];
  __shared__ int s_scan[kTprWarps + 1];
  __shared__ long long s_base;
  __shared__ int s_npass;

  constexpr int S = NC == 1 ? 4096 : 2048;  // == wt.S (host asserts)
  uint32_t* tm = reinterpret_cast<uint32_t*>(smem_raw);                     // [NC][2][S]
  unsigned char* tiles = smem_raw + sizeof(uint32_t) * NC * 2 * S;           // [kTprWarps][tile_bytes]; candidate list later
  uint32_t* list = reinterpret_cast<uint32_t*>(tiles);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* my_tile = tiles + (size_t)warp * tile_bytes;
  uint64_t* my_bar = &s_bar[warp];
  uint32_t phase = 0;

  if (lane == 0) mbar_init(my_bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  for (int win = blockIdx.x; win < wt.n; win += gridDim.x) {
    // generic-proxy writes of the previous window (candidate list) must be ordered before this window's TMA writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < NC * 2 * S / 4; i += kTprThreads) reinterpret_cast<uint4*>(tm)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) s_npass = 0;
    __syncthreads();

    const int w0 = wt.pos0[win];
    const int lo = wt.lo[win], hi = wt.hi[win];

    // ---------------- rows: 32 per warp step, staged through shared memory by TMA ----------------
    for (int base = lo + warp * 32; base < hi; base += kTprWarps * 32) {
      const int r = base + lane;
      const bool valid = r < hi;
      const int rc = valid ? r : hi;  // off[hi] exists (off has n+1 entries)
      const long long o0 = csr.off[rc];
      long long o1 = __shfl_down_sync(0xFFFFFFFFu, o0, 1);
      if (lane == 31) o1 = valid ? csr.off[r + 1] : o0;
      int start = 0, strand0 = 0;
      if (valid) { start = csr.start[r]; strand0 = csr.strand[r] - 1; }
      const int nvalid = min(32, hi - base);
      int first = 0;
      while (first < nvalid) {
        // rows [first, first+k) of this step whose bytes fit one tile
        const long long of = __shfl_sync(0xFFFFFFFFu, o0, first);
        const uint32_t fit = __ballot_sync(0xFFFFFFFFu, valid && lane >= first && (o1 - of) <= (long long)tile_bytes) >> first;
        const int k = (~fit == 0u) ? 32 : max(1, __ffs(~fit) - 1);  // leading run of fitting rows (>=1 by the dispatcher's precondition)
        const long long oe = __shfl_sync(0xFFFFFFFFu, o1, first + k - 1);
        const uint32_t bytes = (uint32_t)(oe - of);
        if (bytes) {
          if (lane == 0) {
            mbar_expect_tx(my_bar, bytes);
            tma_load_1d(my_tile, csr.data + of, bytes, my_bar);
          }
          mbar_wait(my_bar, phase);
          phase ^= 1u;
        }
        if (valid && lane >= first && lane < first + k) {
          const int nch = (int)((o1 - o0) >> 4);
          const unsigned char* row = my_tile + (int)(o0 - of);
          const int rel0 = start - w0;
          const bool overlaps = rel0 + nch * 16 > 0;
          if (overlaps) {
            bool pass = true;
            if (FUSED) {
              uint32_t c0 = 0, c1 = 0;
              for (int c = 0; c < nch; c++) {
                Planes p;
                make_planes(*reinterpret_cast<const uint4*>(row + c * 16), p);
                count_chunk<TCTX>(p, c0, c1);
              }
              pass = eval_pass(c0 >> 16, c0 & 0xFFFFu, c1 >> 16, c1 & 0xFFFFu, thr);
              if (rel0 >= 0) {  // the window that holds the row's start owns its flag
                pass_out[r] = pass ? 1 : 0;
                if (pass) atomicAdd(&s_npass, 1);
              }
            } else {
              pass = pass_in ? (pass_in[r] != 0) : true;
            }
            const int c_first = rel0 >= 0 ? 0 : ((-rel0) >> 4);
            const int c_end = rel0 + nch * 16 > S ? ((S - rel0 + 15) >> 4) : nch;
            for (int c = c_first; c < c_end; c++) {
              Planes p;
              make_planes(*reinterpret_cast<const uint4*>(row + c * 16), p);
              push_calls<WANT>(p, tm, S, strand0, rel0 + c * 16, pass);
            }
          }
        }
        __syncwarp();
        first += k;
      }
    }
    __syncthreads();

    // ---------------- candidates: (pos,strand) with any requested-context call, in position order ----------------
    constexpr int PER = S / kTprThreads;  // positions per thread
    const int q0 = threadIdx.x * PER;
    uint32_t nz = 0;  // bit 2k+s
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
      for (int c = 0; c < NC; c++)
#pragma unroll
        for (int v4 = 0; v4 < PER / 4; v4++) {
          const uint4 v = *reinterpret_cast<const uint4*>(&tm[(c * 2 + s) * S + q0 + v4 * 4]);
          if (v.x) nz |= 1u << (2 * (v4 * 4 + 0) + s);
          if (v.y) nz |= 1u << (2 * (v4 * 4 + 1) + s);
          if (v.z) nz |= 1u << (2 * (v4 * 4 + 2) + s);
          if (v.w) nz |= 1u << (2 * (v4 * 4 + 3) + s);
        }
    int n_cand;
    int my = block_exscan_n<kTprWarps>(__popc(nz), &n_cand, s_scan);
    // (the tiles are dead now: the list aliases them; capacity S*2 entries <= tile area, checked by the host)
    for (uint32_t m = nz; m; m &= m - 1u) {
      const int b = __ffs(m) - 1;
      list[my++] = (uint32_t)(q0 + (b >> 1)) | ((uint32_t)(b & 1) << 16);
    }
    __syncthreads();

    // ---------------- pull the coverage of every candidate from the rows covering it ----------------
    for (int idx = threadIdx.x; idx < n_cand; idx += kTprThreads) {
      const uint32_t e = list[idx];
      const int q = e & 0xFFFF, s = (e >> 16) & 1;
      const int p = w0 + q;
      int a = lo, b = hi;  // first row with start > p
      while (a < b) {
        const int m = (a + b) >> 1;
        if (__ldg(&csr.start[m]) <= p) a = m + 1; else b = m;
      }
      uint32_t cov = 0;
      for (int r = a - 1; r >= lo; r--) {
        const int st = __ldg(&csr.start[r]);
        if (st <= p - max_span) break;
        if (__ldg(&csr.strand[r]) - 1 != s) continue;
        const long long ro = __ldg(&csr.off[r]);
        const int d = p - st;
        if (d < (int)(__ldg(&csr.off[r + 1]) - ro)) cov += ((__ldg(&csr.data[ro + d]) & 15u) != 11u);
      }
      const uint32_t half = cov >> 1;
      uint32_t res = e;
      int slot = 0;
#pragma unroll
      for (int cc = 0; cc < 3; cc++) {
        if (!((WANT >> cc) & 1)) continue;
        if ((tm[(slot * 2 + s) * S + q] & 0xFFFFu) > half) res |= (uint32_t)(slot + 1) << 20 | (uint32_t)(cc == 0 ? 2 : (cc == 1 ? 6 : 7)) << 24;
        slot++;
      }
      list[idx] = res;
    }
    __syncthreads();

    // ---------------- ordered emission ----------------
    const int chunk = (n_cand + kTprThreads - 1) / kTprThreads;
    const int i0 = min(n_cand, (int)threadIdx.x * chunk), i1 = min(n_cand, i0 + chunk);
    int mine = 0;
    for (int i = i0; i < i1; i++) mine += (list[i] >> 20) != 0;
    int nrows_total;
    int my_off = block_exscan_n<kTprWarps>(mine, &nrows_total, s_scan);
    if (threadIdx.x == 0) {
      const unsigned long long bb = nrows_total ? atomicAdd(out.cursor, (unsigned long long)nrows_total) : 0ull;
      s_base = (long long)bb;
      out.w_base[win] = (long long)bb;
      out.w_cnt[win] = nrows_total;
      if (FUSED && s_npass) atomicAdd(out.n_pass, (unsigned long long)s_npass);
    }
    __syncthreads();
    const long long wbase = s_base;
    if (mine && wbase + nrows_total <= out.cap) {
      long long dst = wbase + my_off;
      for (int i = i0; i < i1; i++) {
        const uint32_t e = list[i];
        if (!(e >> 20)) continue;
        const int q = e & 0xFFFF, s = (e >> 16) & 1, slot = (int)((e >> 20) & 15u) - 1, code = (int)(e >> 24);
        const uint32_t v = tm[(slot * 2 + s) * S + q];
        out.scratch[dst++] = make_int4(w0 + q, (s + 1) | (code << 8), (int)(v >> 16), (int)((v & 0xFFFFu) - (v >> 16)));
      }
    }
  }
}

}  // namespace ea
