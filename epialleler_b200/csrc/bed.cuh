// BED-assisted read -> target matching and per-target read counts:
//   k_match_target  replaces rcpp_match_amplicon / rcpp_match_capture (/root/reference/src/rcpp_match_target.cpp:16-81),
//                   which scan every BED row for every read (O(N*B)) and keep the FIRST row in file order that matches;
//   k_bed_counts    replaces the data.table grouping of .getBedReport (/root/reference/R/internal.R:529-561): reads per
//                   (BED row, pass, strand), with one extra bucket for unmatched reads.
// The BED rows are sorted once on the host by (chromosome, start) and by (chromosome, end), each copy keeping its file
// order index; a read then only visits the rows that can match (binary search + a short scan bounded by a running
// maximum of the ends) and keeps the smallest file index among the rows that do - the reference's "first match".
#pragma once
#include "kernels.cuh"

namespace ea {

constexpr int32_t kNaInt = INT32_MIN;  // R's NA_integer_

struct BedIndex {
  // sorted by (chr, start): key = chr << 32 | (start + 2^31); pmax = running maximum of `end` within the chromosome
  const int64_t* s_key; const int32_t* s_start; const int32_t* s_end; const int32_t* s_pmax; const int32_t* s_idx;
  // sorted by (chr, end)
  const int64_t* e_key; const int32_t* e_idx;
  int32_t n;  // rows with a chromosome known to the BAM
};

__host__ __device__ __forceinline__ int64_t bed_key(int32_t chr, int64_t pos) {
  return ((int64_t)chr << 32) + (pos + 2147483648ll);
}

// first index in [0, n) with key[i] >= k
__device__ __forceinline__ int lower_bound_key(const int64_t* __restrict__ key, int n, int64_t k) {
  int a = 0, b = n;
  while (a < b) {
    const int m = (a + b) >> 1;
    if (key[m] < k) a = m + 1; else b = m;
  }
  return a;
}

// mode 0: amplicon, param = tolerance; mode 1: capture, param = min overlap.  One thread per read.
__global__ void __launch_bounds__(256)
k_match_target(Csr csr, BedIndex bed, int mode, int param, int32_t* __restrict__ match) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= csr.n) return;
  const int32_t rc = csr.rname[r];
  const int64_t rs = csr.start[r];
  const int64_t re = rs + csr.len[r] - 1;  // read_end = start + size - 1 (rcpp_match_target.cpp:33,68)
  int32_t best = INT32_MAX;
  if (mode == 0) {
    if (param >= 0) {
      // |read_start - start_i| <= tol
      for (int i = lower_bound_key(bed.s_key, bed.n, bed_key(rc, rs - param)), e = lower_bound_key(bed.s_key, bed.n, bed_key(rc, rs + param + 1)); i < e; i++)
        best = min(best, bed.s_idx[i]);
      // |read_end - end_i| <= tol
      for (int i = lower_bound_key(bed.e_key, bed.n, bed_key(rc, re - param)), e = lower_bound_key(bed.e_key, bed.n, bed_key(rc, re + param + 1)); i < e; i++)
        best = min(best, bed.e_idx[i]);
    }
  } else {
    // min(re, ce) - max(rs, cs) + 1 >= mo.  Necessary: cs <= re - mo + 1 and ce >= rs + mo - 1.
    const int first = lower_bound_key(bed.s_key, bed.n, bed_key(rc, -2147483648ll));
    const int last = lower_bound_key(bed.s_key, bed.n, bed_key(rc, re - (int64_t)param + 2));  // first row with cs > re - mo + 1
    const int64_t need_end = rs + (int64_t)param - 1;
    for (int i = last - 1; i >= first && (int64_t)bed.s_pmax[i] >= need_end; i--) {
      const int64_t cs = bed.s_start[i], ce = bed.s_end[i];
      const int64_t overlap = (re < ce ? re : ce) - (rs > cs ? rs : cs) + 1;
      if (overlap >= (int64_t)param) best = min(best, bed.s_idx[i]);
    }
  }
  match[r] = best == INT32_MAX ? kNaInt : best + 1;
}

// counts[(row or nbed for NA) * 4 + pass * 2 + (strand - 1)] += 1.  Small tables are privatised in shared memory.
__global__ void __launch_bounds__(256)
k_bed_counts(Csr csr, const int32_t* __restrict__ match, const uint8_t* __restrict__ pass /* null = all TRUE */, int nbed,
             unsigned long long* __restrict__ counts, int smem_entries) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned int* s_cnt = reinterpret_cast<unsigned int*>(smem_raw);
  const int n_entries = (nbed + 1) * 4;
  const bool priv = smem_entries >= n_entries;
  if (priv) {
    for (int i = threadIdx.x; i < n_entries; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
  }
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < csr.n; r += (int64_t)gridDim.x * blockDim.x) {
    const int32_t m = match[r];
    const int row = m == kNaInt ? nbed : m - 1;
    const int p = pass ? (pass[r] != 0) : 1;
    const int e = row * 4 + p * 2 + (csr.strand[r] - 1);
    // rows are sorted, so the 32 consecutive reads of a warp fall into a handful of buckets (most of them into the same
    // target, or into the "unmatched" bucket): one atomic per distinct bucket and warp instead of one per read
    const unsigned peers = __match_any_sync(__activemask(), e);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) {
      const unsigned c = (unsigned)__popc(peers);
      if (priv) atomicAdd(&s_cnt[e], c); else atomicAdd(&counts[e], (unsigned long long)c);
    }
  }
  if (priv) {
    __syncthreads();
    for (int i = threadIdx.x; i < n_entries; i += blockDim.x)
      if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
  }
}

}  // namespace ea
