// common.cuh — device-side types and float helpers shared by every kernel of libb2icp.
//
// Float semantics: PCL/FLANN/Eigen on baseline x86-64 evaluate the point transform and the squared
// distance with one IEEE rounding per operation and no FMA.
// All arithmetic whose result decides a correspondence index or a gate goes through the *_rn
// intrinsics below so that nvcc cannot contract a*b+c into an FMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

// CTA shape of the fused sweep (and of the search kernels that share its scratch layout).  Measured on B200, streamed
// leg of bench.py (scans/s): 256 threads x 3 CTAs/SM (80 registers, the round-1 shape) 21 674; 128 x 6 (80 registers)
// 21 972; 128 x 5 (96 registers, no spills, 20 warps/SM) 22 130; 192 x 3 (96 registers, 18 warps/SM) 20 984.  At 80
// registers the search loop re-materialises addresses and re-reads pointers on every trip; 96 is where that stops.
#ifndef B2_SWEEP_THREADS
#define B2_SWEEP_THREADS 128
#endif
#ifndef B2_SWEEP_MIN_CTAS
#define B2_SWEEP_MIN_CTAS 5
#endif
#ifndef B2_PHASE_A_UNROLL
#define B2_PHASE_A_UNROLL 1
#endif
constexpr int kPhaseAUnroll = B2_PHASE_A_UNROLL;
constexpr int kSweepThreads = B2_SWEEP_THREADS;    // CTA size of the fused sweep kernels
constexpr int kSweepMinCtas = B2_SWEEP_MIN_CTAS;   // resident CTAs per SM the sweep is compiled for
constexpr int kNumSums = 17;         // n, sum(src)[3], sum(dst)[3], sum(dst*src^T)[9], sum(d2)
constexpr unsigned long long kInfKey = 0x7F8000007FFFFFFFull;  // d2 = +inf, idx = INT_MAX

// Uniform neighbour grid over the target cloud (built by grid.cuh).  Cells are ordered x-fastest,
// so the cells (x0..x1, y, z) of one row are one contiguous range of `pts`.
struct GridView {
  const float4* pts;      // target points sorted by cell; .w carries the ORIGINAL index (int bits)
  const int* cell_start;  // [ncell + 1] exclusive prefix sums of per-cell counts
  float ox, oy, oz;       // min corner
  float cell, inv_cell;   // cell edge and its reciprocal
  float slack;            // subtracted from every face distance: absorbs binning/face rounding
  int nx, ny, nz;
  int n;                  // number of points
};

// Device-resident state of one ICP run (one scan).  Written by the last CTA of every sweep.
struct IcpState {
  float Tinc[16];    // transform the NEXT sweep applies to the running cloud (row-major)
  float final_T[16]; // final_transformation_ (Matrix4f)
  double mse;        // mean gated d2 of the last sweep
  double prev_mse;   // DefaultConvergenceCriteria::correspondences_prev_mse_
  double fitness_sum;
  unsigned long long fitness_cnt;
  int iter;          // nr_iterations_
  int done;          // loop has ended (converged or failed): later sweeps return immediately
  int converged;     // converged_
  int status;        // b2icp_status of the loop
  int n_corr;        // gated correspondences of the last sweep
  unsigned int ticket;   // last-CTA election counter
  unsigned int unresolved;  // P2P loop: queries that needed a real search, summed over the iterations
  int pad;           // set to 1 by a sweep thread that met a non-finite coordinate
};

struct IcpConfig {
  double max2;             // max_correspondence_distance^2 (gate, compared in double like PCL)
  double rot_thresh;       // 1 - transformation_epsilon
  double trans_thresh;     // transformation_epsilon (compared with the SQUARED translation)
  double mse_abs;          // 1e-12
  double mse_rel;          // euclidean_fitness_epsilon
  float bound2;            // float upper bound of max2 used for pruning
  int max_iterations;
  int min_corr;            // 3 (ICP)
  int max_rings;           // ring budget before a query is handed to the fallback / cell span of a box search
  float margin_frac;       // nncache.cuh: search-radius margin as a fraction of the grid cell
};

// One scan: source cloud, running (in-place transformed) cloud, outputs, state.
struct ScanTask {
  GridView grid;
  const float4* src;   // source as uploaded (w ignored)
  float4* cur;         // input_transformed: rewritten in place by every sweep; .w = lower bound on the distance from
                       // the point to every target point other than its cached neighbours (nncache.cuh)
  int* corr_idx;       // [n] target index of the last sweep, -1 = gated out
  float* corr_d2;      // [n] float d2 of the last sweep
  float4* c0;          // [n] nearest target point found for cur[i]: xyz + original index (int bits, -1 = none)
  float4* c1;          // [n] runner-up, same layout (nncache.cuh)
  float4* c2;          // [n] third nearest (used when kCacheK == 3)
  double* partials;    // [ceil(n / 32)][kNumSums] per-warp sums of one sweep
  IcpState* state;
  int n;
  int pad;             // 1: c0 / c1 / cur.w hold the certificate of cur (left by the point-to-point loop)
};

// Per-query state is streamed once per iteration and is larger than what the L2 can keep next to the
// target cloud and its cell table: evict-first accesses keep it from pushing those out.
template <class T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }
template <class T>
__device__ __forceinline__ void st_stream(T* p, const T& v) { __stcs(p, v); }

__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }

// FLANN L2_Simple: result = ((0 + dx*dx) + dy*dy) + dz*dz, each op rounded to float.
__device__ __forceinline__ float sqdist3(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = fsub(qx, px), dy = fsub(qy, py), dz = fsub(qz, pz);
  return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

// pcl::transformPointCloud / Eigen 4x4 * 4x1 in float: ((m0*x + m1*y) + m2*z) + m3.
__device__ __forceinline__ float4 xform_f(const float* __restrict__ T, float x, float y, float z) {
  float4 o;
  o.x = fadd(fadd(fadd(fmul(T[0], x), fmul(T[1], y)), fmul(T[2], z)), T[3]);
  o.y = fadd(fadd(fadd(fmul(T[4], x), fmul(T[5], y)), fmul(T[6], z)), T[7]);
  o.z = fadd(fadd(fadd(fmul(T[8], x), fmul(T[9], y)), fmul(T[10], z)), T[11]);
  o.w = 1.0f;
  return o;
}

// (d2, idx) packed so that an unsigned 64-bit min is the lexicographic min: non-negative floats
// order like their bit patterns, ties fall to the smaller index (canonical rule).
__device__ __forceinline__ unsigned long long pack_key(float d2, int idx) {
  return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)idx;
}
__device__ __forceinline__ float key_d2(unsigned long long k) { return __uint_as_float((unsigned int)(k >> 32)); }
__device__ __forceinline__ int key_idx(unsigned long long k) { return (int)(unsigned int)(k & 0xFFFFFFFFull); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, k, o);
    k = other < k ? other : k;
  }
  return k;
}

}  // namespace b2
