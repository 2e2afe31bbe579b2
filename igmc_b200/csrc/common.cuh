// Shared device helpers for the IGMC hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IGMC_WARP 32
#define IGMC_FULL 0xffffffffu

// ---- error codes written to the device-side error word -------------------------------------
enum {
  IGMC_OK = 0,
  IGMC_ERR_NODE_CAP = 1,    // a subgraph side exceeded the node-list capacity
  IGMC_ERR_EDGE_CAP = 2,    // batch edge capacity exceeded
  IGMC_ERR_NODE_TOTAL = 3,  // batch node capacity exceeded
  IGMC_ERR_SMEM_NODES = 4,  // subgraph larger than the model kernel's shared-memory plan
  IGMC_ERR_BAD_BATCH = 5,   // edges not grouped by graph / node ids outside their graph
};

__device__ __forceinline__ void igmc_set_err(int* err, int code) {
  if (err) atomicCAS(err, 0, code);
}

// ---- counter-based hashing (bit-exact twin: oracle/extract_np.py::splitmix64/hash_keys) ----
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// per (seed, pair, side, hop) stream state
__host__ __device__ __forceinline__ uint64_t sample_state(uint64_t seed, int64_t pair_id, int side, int hop) {
  uint64_t tag = (uint64_t)pair_id * 16ull + (uint64_t)hop * 2ull + (uint64_t)side;
  return splitmix64(seed ^ splitmix64(tag));
}
__host__ __device__ __forceinline__ uint32_t sample_key(uint64_t state, int node) {
  return (uint32_t)(splitmix64(state + (uint64_t)(uint32_t)node) >> 32);
}

// edge-dropout draw for directed edge `eid` of the current step: keep iff hash >= thresh
// (thresh = p * 2^32).  Twin: igmc_b200/models.py::edge_keep_reference.
__host__ __device__ __forceinline__ bool edge_keep(uint64_t seed, uint32_t eid, uint32_t thresh) {
  return (uint32_t)(splitmix64(seed + (uint64_t)eid) >> 32) >= thresh;
}

// ---- warp / block primitives -----------------------------------------------------------------
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(IGMC_FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(IGMC_FULL, v, o);
  return v;
}
__device__ __forceinline__ int warp_incl_scan_i(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(IGMC_FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one int per thread.  `ws` is >= 33 ints of shared scratch.
// Returns the exclusive prefix; *total receives the block sum.  Ends with a barrier so `ws`
// can be reused immediately.
__device__ __forceinline__ int block_excl_scan_i(int v, int* ws, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int incl = warp_incl_scan_i(v, lane);
  if (lane == 31) ws[w] = incl;
  __syncthreads();
  if (w == 0) {
    int s = lane < nw ? ws[lane] : 0;
    int inc = warp_incl_scan_i(s, lane);
    ws[lane] = inc - s;
    if (lane == 31) ws[32] = inc;
  }
  __syncthreads();
  int excl = ws[w] + incl - v;
  *total = ws[32];
  __syncthreads();
  return excl;
}

__device__ __forceinline__ int block_sum_i(int v, int* ws) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum_i(v);
  if (lane == 0) ws[w] = v;
  __syncthreads();
  int s = 0;
  if (w == 0) {
    s = lane < nw ? ws[lane] : 0;
    s = warp_sum_i(s);
    if (lane == 0) ws[32] = s;
  }
  __syncthreads();
  s = ws[32];
  __syncthreads();
  return s;
}

#define IGMC_CUDA_CHECK_LAUNCH()                      \
  do {                                                \
    cudaError_t e__ = cudaGetLastError();             \
    if (e__ != cudaSuccess) return (int)e__ + 1000;   \
  } while (0)
