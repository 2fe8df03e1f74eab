// radix_sort.cuh -- stable LSD radix sort of (u32 key, u32 value) pairs for the cell-index sorts of the hot path (elevation fuse
// runs; 21-24 key bits), written for two things the library sort could not give:
//   * the element count lives in DEVICE memory (no host synchronisation between producing the keys and sorting them), and
//   * the first pass can DROP a sentinel key: invalid points (filtered by the sensor model, outside the map, or -- on a
//     tile-sharded grid -- owned by another GPU) never enter the sort, so every rank sorts only its own share.
// 8-bit digits; per pass: per-block digit histogram -> exclusive scan of the [digit][block] table (two small kernels) -> stable
// scatter (rank inside the block from warp match-any groups, in element order).  Blocks own RS_TILE consecutive elements.
#pragma once
#include "common.cuh"

namespace mrslam {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;                        // elements per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // 2048 elements per block
constexpr int RS_BINS = 256;
constexpr int RS_SCAN_CHUNK = 1024;                // entries of the [digit][block] table scanned by one block

struct RadixScratch {
  DevBuf<unsigned> hist;     // [RS_BINS][nblk] digit counts, then exclusive offsets inside their scan chunk
  DevBuf<unsigned> chunk;    // per scan chunk: total, then exclusive prefix
  DevBuf<int> count;         // [2] element counts (ping-pong between passes)
};

__device__ __forceinline__ int rs_count(const int* d_n, int n_max) { return d_n ? min(*d_n, n_max) : n_max; }

// per-block digit histogram; hist[d * nblk + b]
static __global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const unsigned* __restrict__ keys, const int* __restrict__ d_n, int n_max, int nblk,
                                                        int shift, unsigned drop_key, int use_drop, unsigned* __restrict__ hist) {
  __shared__ unsigned s_h[RS_BINS];
  const int n = rs_count(d_n, n_max);
  const int b = blockIdx.x;
  s_h[threadIdx.x] = 0;
  __syncthreads();
  const int base = b * RS_TILE;
  if (base < n) {
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
      // warp-aggregated: neighbouring points land in the same few digits (cells of one scan line), which would serialise
      // per-element shared-memory atomics; one atomic per group of equal digits in the warp instead
      const int e = base + i * RS_THREADS + threadIdx.x;
      bool live = e < n;
      unsigned k = live ? keys[e] : 0u;
      if (use_drop && k == drop_key) live = false;
      const unsigned d = (k >> shift) & 0xffu;
      const unsigned grp = __match_any_sync(FULL, live ? d : (256u + (threadIdx.x & 31u)));
      if (live && (grp & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&s_h[d], (unsigned)__popc(grp));
    }
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblk + b] = s_h[threadIdx.x];
}

// exclusive scan inside chunks of RS_SCAN_CHUNK consecutive table entries (digit-major order = output order); chunk totals out
static __global__ void __launch_bounds__(RS_SCAN_CHUNK) k_rs_scan_chunks(unsigned* __restrict__ hist, int total, unsigned* __restrict__ chunk) {
  __shared__ unsigned s_w[RS_SCAN_CHUNK / 32];
  const int e = blockIdx.x * RS_SCAN_CHUNK + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned v = e < total ? hist[e] : 0u;
  unsigned x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(FULL, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_w[w] = x;
  __syncthreads();
  if (w == 0) {
    unsigned t = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(FULL, t, o);
      if (lane >= o) t += y;
    }
    s_w[lane] = t;  // inclusive over warps
  }
  __syncthreads();
  const unsigned warp_base = w ? s_w[w - 1] : 0u;
  if (e < total) hist[e] = warp_base + x - v;
  if (threadIdx.x == RS_SCAN_CHUNK - 1) chunk[blockIdx.x] = warp_base + x;
}

// exclusive scan of the chunk totals by one block (any number of chunks: serial over tiles of 1024); grand total -> *d_n_out
static __global__ void __launch_bounds__(1024) k_rs_scan_totals(unsigned* __restrict__ chunk, int nchunks, int* __restrict__ d_n_out) {
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int c0 = 0; c0 < nchunks; c0 += 1024) {
    const int e = c0 + threadIdx.x;
    const unsigned v = e < nchunks ? chunk[e] : 0u;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(FULL, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_w[w] = x;
    __syncthreads();
    if (w == 0) {
      unsigned t = s_w[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(FULL, t, o);
        if (lane >= o) t += y;
      }
      s_w[lane] = t;
    }
    __syncthreads();
    const unsigned carry = s_carry;
    const unsigned incl = carry + (w ? s_w[w - 1] : 0u) + x;
    if (e < nchunks) chunk[e] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && d_n_out) *d_n_out = (int)s_carry;
}

// stable scatter: destination = offset of (digit, block) + rank of the element among the block's elements of that digit that
// precede it.  Warp w owns elements [w*256, w*256+256) of the tile, visited 32 at a time in order.
static __global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const unsigned* __restrict__ keys, const unsigned* __restrict__ vals,
                                                           const int* __restrict__ d_n, int n_max, int nblk, int shift, unsigned drop_key,
                                                           int use_drop, const unsigned* __restrict__ hist, const unsigned* __restrict__ chunk,
                                                           unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out) {
  __shared__ unsigned s_cnt[RS_THREADS / 32][RS_BINS];  // per warp: running count per digit, then exclusive base over warps
  __shared__ unsigned s_off[RS_BINS];                   // global offset of (digit, this block)
  const int n = rs_count(d_n, n_max);
  const int b = blockIdx.x, base = b * RS_TILE;
  if (base >= n) return;  // block-uniform
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int i = 0; i < RS_THREADS / 32; i++) s_cnt[i][threadIdx.x] = 0;
  {
    const size_t t = (size_t)threadIdx.x * nblk + b;
    s_off[threadIdx.x] = hist[t] + chunk[t / RS_SCAN_CHUNK];
  }
  __syncthreads();
  unsigned k[RS_ITEMS], v[RS_ITEMS], rank[RS_ITEMS];
  bool live[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const int e = base + w * (32 * RS_ITEMS) + i * 32 + lane;
    live[i] = e < n;
    k[i] = live[i] ? keys[e] : 0u;
    v[i] = live[i] ? vals[e] : 0u;
    if (use_drop && k[i] == drop_key) live[i] = false;
    const unsigned d = (k[i] >> shift) & 0xffu;
    // lanes with the same digit (dead lanes get a group of their own: digit 256 + lane can never match a live one)
    const unsigned grp = __match_any_sync(FULL, live[i] ? d : (256u + (unsigned)lane));
    const unsigned prior = live[i] ? s_cnt[w][d] : 0u;
    rank[i] = prior + __popc(grp & lt);
    __syncwarp();
    if (live[i] && (grp & lt) == 0) s_cnt[w][d] = prior + __popc(grp);  // the group's first lane
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive prefix of every digit over the warps
    unsigned run = 0;
#pragma unroll
    for (int i = 0; i < RS_THREADS / 32; i++) {
      const unsigned c = s_cnt[i][threadIdx.x];
      s_cnt[i][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++)
    if (live[i]) {
      const unsigned d = (k[i] >> shift) & 0xffu;
      const unsigned dst = s_off[d] + s_cnt[w][d] + rank[i];
      keys_out[dst] = k[i];
      vals_out[dst] = v[i];
    }
}

// scratch for up to n_max elements (allocations are ordered on the library stream, see DevBuf)
inline int radix_reserve(RadixScratch& S, int n_max) {
  const int nblk = std::max(1, (n_max + RS_TILE - 1) / RS_TILE);
  const int total = RS_BINS * nblk;
  MR_CHECK(S.hist.reserve((size_t)total));
  MR_CHECK(S.chunk.reserve((size_t)((total + RS_SCAN_CHUNK - 1) / RS_SCAN_CHUNK)));
  MR_CHECK(S.count.reserve(2));
  return MRSLAM_OK;
}

// Sorts by key bits [begin_bit, end_bit).  n_max = capacity (grid size); d_n_in = element count on the device (null: n_max).
// Elements whose key equals drop_key are removed in the first pass when use_drop != 0.  Buffers a = input, b = scratch; the
// sorted pairs end up in *keys_sorted / *vals_sorted (one of the two buffers) and their count in *d_n_sorted (device).
inline int radix_sort_pairs(RadixScratch& S, unsigned* keys_a, unsigned* vals_a, unsigned* keys_b, unsigned* vals_b, int n_max,
                            const int* d_n_in, int begin_bit, int end_bit, unsigned drop_key, int use_drop, cudaStream_t st,
                            unsigned** keys_sorted, unsigned** vals_sorted, const int** d_n_sorted) {
  const int nblk = std::max(1, (n_max + RS_TILE - 1) / RS_TILE);
  const int total = RS_BINS * nblk;
  const int nchunks = (total + RS_SCAN_CHUNK - 1) / RS_SCAN_CHUNK;
  MR_CHECK(radix_reserve(S, n_max));
  const int* d_n = d_n_in;
  unsigned *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
  int pass = 0;
  for (int shift = begin_bit; shift < end_bit; shift += 8, pass++) {
    const int drop = (pass == 0) ? use_drop : 0;
    k_rs_hist<<<nblk, RS_THREADS, 0, st>>>(ki, d_n, n_max, nblk, shift, drop_key, drop, S.hist.p);
    k_rs_scan_chunks<<<nchunks, RS_SCAN_CHUNK, 0, st>>>(S.hist.p, total, S.chunk.p);
    int* d_out = S.count.p + (pass & 1);
    k_rs_scan_totals<<<1, 1024, 0, st>>>(S.chunk.p, nchunks, d_out);
    k_rs_scatter<<<nblk, RS_THREADS, 0, st>>>(ki, vi, d_n, n_max, nblk, shift, drop_key, drop, S.hist.p, S.chunk.p, ko, vo);
    count_launch(4);
    d_n = d_out;
    std::swap(ki, ko);
    std::swap(vi, vo);
  }
  MR_CUDA(cudaGetLastError());
  *keys_sorted = ki;
  *vals_sorted = vi;
  *d_n_sorted = d_n;
  return MRSLAM_OK;
}

}  // namespace mrslam
