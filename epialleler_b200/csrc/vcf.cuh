// Base frequencies at SNV positions: the device side of rcpp_get_base_freqs
// (/root/reference/src/rcpp_get_base_freqs.cpp:15-61).  The reference walks reads and VCF entries with one shared cursor
// and assumes both are sorted; here every read binary-searches the sorted VCF keys (chromosome << 32 | position) for the
// entries inside [start, start + len - 1] and adds one to cell (entry, base + 5*(strand-1) + 10*pass), base = HTSlib's
// seq_nt16_int of the HIGH nibble of the packed SEQ|XM byte (A,C,G,T = 0..3, anything else 4).
//
// Entries whose chromosome is NA never match, but they are not inert in the reference: its cursor treats them as "before
// the read" and moves onto them as soon as a read's scan reaches them (rcpp_get_base_freqs.cpp:44-47), which hides every
// earlier entry from all LATER reads.  A scan reaches an NA entry exactly when the read's end is not before the last
// valid entry in front of it, so entry i is visible to rows 0..limit(i) only, limit(i) = the first row whose running
// maximum of (rname, end) reaches the key of the last valid entry before the first NA entry that follows i
// (k_vcf_na_limits: one binary search per NA entry on the handle's prefix maximum of row ends).
#pragma once
#include "kernels.cuh"

namespace ea {

// keys: sorted valid entries; idx: their rows in the caller's table; counts: [20][n_vcf] (column-major like R's matrix)
// maxend[r] = rname << 32 | running maximum of (start + len) over rows 0..r (0 for empty rows): monotone.
// na_key[j] = key of the last valid entry in front of NA entry j; limit_out[j] = first row whose inclusive end reaches it.
__global__ void k_vcf_na_limits(const uint64_t* __restrict__ maxend, int64_t n_rows, const uint64_t* __restrict__ na_key, int32_t n_na,
                                int32_t* __restrict__ limit_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_na) return;
  const uint64_t K = na_key[j];
  int64_t a = 0, z = n_rows;
  while (a < z) {
    const int64_t m = (a + z) >> 1;
    if (maxend[m] <= K) a = m + 1; else z = m;  // inclusive end = maxend - 1 >= K  <=>  maxend > K
  }
  limit_out[j] = (int32_t)min(a, (int64_t)0x7FFFFFFF);
}

__global__ void k_base_freqs(Csr csr, const uint8_t* __restrict__ pass, const uint64_t* __restrict__ keys, const int32_t* __restrict__ idx,
                             const int32_t* __restrict__ limit /* per valid entry: last row that sees it */,
                             int32_t n_valid, int64_t n_vcf, unsigned int* __restrict__ counts) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= csr.n) return;
  const int32_t len = csr.len[r];
  if (len <= 0) return;
  const uint64_t lo_key = ((uint64_t)(uint32_t)csr.rname[r] << 32) | (uint32_t)csr.start[r];
  const uint64_t hi_key = lo_key + (uint32_t)(len - 1);
  int a = 0, z = n_valid;
  while (a < z) {
    const int m = (a + z) >> 1;
    if (keys[m] < lo_key) a = m + 1; else z = m;
  }
  if (a >= n_valid || keys[a] > hi_key) return;
  const uint8_t* row = csr.data + csr.off[r];
  const int col0 = (csr.strand[r] == 2 ? 5 : 0) + ((pass == nullptr || pass[r]) ? 10 : 0);
  for (int k = a; k < n_valid; k++) {
    const uint64_t key = keys[k];
    if (key > hi_key) break;
    if (r > limit[k]) continue;
    const uint32_t code = row[key - lo_key] >> 4;
    const int base = (0x4444444344424104ull >> (4 * code)) & 15;  // seq_nt16_int: 1->0 (A), 2->1 (C), 4->2 (G), 8->3 (T), else 4
    atomicAdd(&counts[(int64_t)(col0 + base) * n_vcf + idx[k]], 1u);
  }
}

__global__ void k_counts_to_double(const unsigned int* __restrict__ counts, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)counts[i];
}

}  // namespace ea
