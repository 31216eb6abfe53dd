// sort.cuh — the ENTRY ARRAY of a batch: every query of every scan of the batch, ordered by the Z-order (Morton)
// code of the target-grid cell it falls in (scans interleaved).
//
// Why: the neighbour search of icp_sweep_coop (sweep.cuh) is warp-cooperative — the 32 (or 8) queries of a group
// share ONE staged set of candidate target points.  That only pays when the queries of a group are close
// together, and how close they are is set by the QUERY density.  One HDL-64 sweep has ~8 returns per m^2 of
// surface against ~25 map points per m^2; the 32 sweeps of a batch together have ~250 per m^2.  Sorting the
// batch's queries by cell ACROSS scans turns a group of 32 consecutive entries into "the queries of ~one cell",
// whose candidate set is the ~25 map points around that cell (reference: the per-scan loop around
// icp.align(), src/icpslam/icp_odometer.cpp:198 — the scans are independent, so any order is legal).
//
// The sort is a stable LSD radix sort (11-bit digits) written for determinism: equal keys keep their input
// order (scan, then original index), so the entry array — and with it every group and every sum — is a pure
// function of the inputs.  Per pass: sort_hist (per-chunk digit histogram) -> exclusive scan of the
// [digit][chunk] table (grid.cuh scan_*) -> sort_scatter (stable ranks: per-warp histograms + match_any).
#pragma once
#include "common.cuh"
#include "grid.cuh"

namespace b2 {

constexpr int kSortBits = 11;
constexpr int kSortBins = 1 << kSortBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortPerWarp = 512;                        // consecutive elements owned by one warp
constexpr int kSortChunk = kSortWarps * kSortPerWarp;    // 4096 elements per CTA

// Z-order (Morton) code of a cell: bit l of x, y, z interleaved from the low bits up; axes that have run out of bits
// are skipped, so the code has bits.x + bits.y + bits.z bits.  Consecutive codes are close in space at every scale,
// which is what keeps the union box of a warp's compacted work list small.
__device__ __forceinline__ unsigned int morton_key(int cx, int cy, int cz, int bx, int by, int bz) {
  unsigned int key = 0;
  int pos = 0;
  const int top = max(bx, max(by, bz));
  for (int l = 0; l < top; ++l) {
    if (l < bx) key |= (unsigned int)((cx >> l) & 1) << pos++;
    if (l < by) key |= (unsigned int)((cy >> l) & 1) << pos++;
    if (l < bz) key |= (unsigned int)((cz >> l) & 1) << pos++;
  }
  return key;
}

struct KeyParams {
  int bx, by, bz;  // bits per axis after `drop`
  int drop;        // low cell-coordinate bits left out of the key (only when the full code would not fit 32 bits)
  int cell_bits;   // bx + by + bz: the segment sits above them
};

// key of every query: (segment << cell_bits) | Morton code of the cell of guess * p in the scan's target grid;
// value = global id (ent_off + index)
__global__ void __launch_bounds__(256) entry_keys(const ScanTask* __restrict__ tasks, KeyParams kp,
                                                  unsigned int* __restrict__ keys, unsigned int* __restrict__ vals) {
  const ScanTask& t = tasks[blockIdx.y];
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= t.n) return;
  const float4 p = __ldg(t.src + i);
  const float4 q = xform_f(t.state->Tinc, p.x, p.y, p.z);  // Tinc holds the initial guess before the first sweep
  const GridView& g = t.grid;
  const int cx = cell_coord(q.x, g.ox, g.inv_cell, g.nx) >> kp.drop;
  const int cy = cell_coord(q.y, g.oy, g.inv_cell, g.ny) >> kp.drop;
  const int cz = cell_coord(q.z, g.oz, g.inv_cell, g.nz) >> kp.drop;
  keys[t.ent_off + i] = ((unsigned int)t.seg << kp.cell_bits) | morton_key(cx, cy, cz, kp.bx, kp.by, kp.bz);
  vals[t.ent_off + i] = (unsigned int)(t.ent_off + i);
}

__global__ void __launch_bounds__(kSortThreads) sort_hist(const unsigned int* __restrict__ keys, int n, int shift,
                                                          int nchunk, int* __restrict__ hist) {
  __shared__ int h[kSortBins];
  for (int b = threadIdx.x; b < kSortBins; b += kSortThreads) h[b] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSortChunk;
  for (int k = threadIdx.x; k < kSortChunk; k += kSortThreads) {
    const int i = base + k;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & (kSortBins - 1)], 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kSortBins; b += kSortThreads) hist[(size_t)b * nchunk + blockIdx.x] = h[b];
}

// `scanned` = exclusive prefix sums of hist ([digit][chunk] order): the first output slot of (digit, chunk).
__global__ void __launch_bounds__(kSortThreads) sort_scatter(const unsigned int* __restrict__ keys,
                                                             const unsigned int* __restrict__ vals, int n, int shift,
                                                             int nchunk, const int* __restrict__ scanned,
                                                             unsigned int* __restrict__ keys_out,
                                                             unsigned int* __restrict__ vals_out) {
  extern __shared__ int wh[];  // [kSortWarps][kSortBins]: per-warp digit counts, then per-warp running output slots
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b = threadIdx.x; b < kSortWarps * kSortBins; b += kSortThreads) wh[b] = 0;
  __syncthreads();
  const int wbase = blockIdx.x * kSortChunk + warp * kSortPerWarp;
  int* mine = wh + warp * kSortBins;
  for (int r = 0; r < kSortPerWarp; r += 32) {
    const int i = wbase + r + lane;
    if (i < n) atomicAdd(&mine[(keys[i] >> shift) & (kSortBins - 1)], 1);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kSortBins; b += kSortThreads) {
    int run = scanned[(size_t)b * nchunk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const int c = wh[w * kSortBins + b];
      wh[w * kSortBins + b] = run;
      run += c;
    }
  }
  __syncthreads();
  // every warp walks its elements in order; inside a round, lanes with the same digit take consecutive slots in
  // lane order (match_any), so equal digits keep their input order: the pass is stable
  for (int r = 0; r < kSortPerWarp; r += 32) {
    const int i = wbase + r + lane;
    const bool valid = i < n;
    const unsigned int key = valid ? keys[i] : 0u;
    const unsigned int digit = valid ? ((key >> shift) & (kSortBins - 1)) : (unsigned int)(kSortBins + lane);
    const unsigned int peers = __match_any_sync(0xFFFFFFFFu, digit);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    int slot = 0;
    if (valid && lane == leader) {
      slot = mine[digit];
      mine[digit] = slot + __popc(peers);
    }
    slot = __shfl_sync(0xFFFFFFFFu, slot, leader);
    __syncwarp();
    if (valid) {
      keys_out[slot + rank] = key;
      vals_out[slot + rank] = vals[i];
    }
  }
}

// Entry e <- the query with global id vals[e] (or e itself when the batch is not sorted): its source point, its
// scan, its original index, and the inverse map pos[scan][index] = e that the per-scan kernels read through.
__global__ void __launch_bounds__(256) entry_fill(const ScanTask* __restrict__ tasks, int nscan,
                                                  const unsigned int* __restrict__ vals, int E,
                                                  float4* __restrict__ ent_src, unsigned char* __restrict__ ent_sid,
                                                  int* __restrict__ ent_orig) {
  __shared__ int s_off[kMaxScans + 1];
  for (int k = threadIdx.x; k <= nscan; k += blockDim.x) s_off[k] = k < nscan ? tasks[k].ent_off : E;
  __syncthreads();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int gid = vals ? (int)vals[e] : e;
  int lo = 0, hi = nscan;  // last scan whose offset is <= gid
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s_off[mid] <= gid) lo = mid; else hi = mid;
  }
  const ScanTask& t = tasks[lo];
  const int i = gid - s_off[lo];
  const float4 p = __ldg(t.src + i);
  ent_src[e] = make_float4(p.x, p.y, p.z, 1.0f);
  ent_sid[e] = (unsigned char)lo;
  ent_orig[e] = i;
  t.pos[i] = e;
}

}  // namespace b2
