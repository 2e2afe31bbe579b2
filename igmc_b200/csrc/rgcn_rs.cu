// v2 fused relational message passing: relation-space aggregate + dense transform, one thread-block
// CLUSTER per enclosing subgraph (1/2/4 CTAs split the destination nodes), up to 32 warps per CTA.
//
// Same math and C-ABI as csrc/rgcn.cu (which stays as the generic path for many relations); this is the
// fast path for num_relations <= RS_MAX_R.  Per layer (reference: tanh(RGCNConv) models.py:200-202,
// PyG 1.4.2 semantics SURVEY.md A.1), with W_r = sum_b att[r,b] basis[b] formed once per CTA in shared memory:
//   AGG[v,r,:] = sum_{(u->v) of type r, kept} h[u,:]      warp owns 4 destination nodes; 8-lane groups x float4,
//                                                          edge lists staged in shared memory, software-pipelined
//   h'[v]      = tanh( 1/deg(v) * sum_r AGG[v,r,:] W_r + h[v] root + bias )    lane = (node, 4 output channels)
// Node features of the WHOLE subgraph stay in shared memory; CTAs of a cluster exchange their rows through
// L2 (ld/st.cg) + barrier.cluster once per layer.  Backward = same two passes on the out-lists with W_r^T,
// a K=n weight-gradient tile GEMM, and the (att,basis) chain rule applied to the per-CTA dW_r.
// No float atomics; every reduction has a fixed order (bitwise run-to-run deterministic).
#include <cooperative_groups.h>

#include "common.cuh"
#include "../../include/igmc_b200.h"

namespace cg = cooperative_groups;

namespace rs {

constexpr int HID = IGMC_HIDDEN;
constexpr int L1O = IGMC_LIN1_OUT;
constexpr int RS_MAX_R = 12;
constexpr int GN = 4;          // destination nodes per warp group (one 8-lane group each)
constexpr int TW = 32;         // node tile of the weight-gradient GEMM
constexpr int WP = 33;         // padded row stride of the transposition scratch
constexpr uint32_t DROPPED = 0xFFFFFFFFu;

// activation tiles [n][32]: column XOR-swizzled by the row in 4-float groups (a row stays one 128 B line,
// a fixed column group over 8 consecutive rows hits 8 different bank groups)
__device__ __forceinline__ int hix(int v, int c) { return (v << 5) + (c ^ ((v & 7) << 2)); }
__host__ __device__ __forceinline__ int a4(int x) { return (x + 3) & ~3; }

struct Keep {
  bool active;
  const uint8_t* mask;
  uint64_t seed;
  uint32_t thresh;
  __device__ __forceinline__ bool keep(int e) const {
    return mask ? (mask[e] != 0) : edge_keep(seed, (uint32_t)e, thresh);
  }
};
__device__ __forceinline__ Keep make_keep(const igmc_dropout_t& D, int training) {
  Keep K;
  K.mask = D.edge_keep;
  K.seed = D.seed_dev ? *D.seed_dev : D.seed;
  K.active = training && (D.adj_dropout > 0.0f || D.edge_keep != nullptr);
  double t = (double)D.adj_dropout * 4294967296.0;
  K.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  return K;
}

// entry of list position p, DROPPED if the (possibly mirrored) edge is dropped this step
__device__ __forceinline__ uint32_t load_entry(const uint32_t* __restrict__ adj, const int32_t* __restrict__ eid,
                                               int p, const Keep& K, bool mirror, int eb, int m_half) {
  uint32_t ent = __ldg(adj + p);
  if (K.active) {
    int e = __ldg(eid + p);
    if (mirror) { const int el = e - eb; e = eb + (el < m_half ? el + m_half : el - m_half); }
    if (!K.keep(e)) ent = DROPPED;
  }
  return ent;
}

// Edge lists of the own nodes [lo,hi): staged once per kernel in shared memory, dropped edges compacted out,
// with block-local offsets lptr[0..n_own].  Falls back to the global arrays when they do not fit.
struct Lists {
  const uint32_t* lst;   // staged entries (nullptr: not staged)
  const int* lptr;       // [n_own+1] offsets into lst
  const uint32_t* adj;   // global fallback
  const int32_t* eid;
  const int32_t* ptr;
  bool mirror;
  int eb, m_half;
};

// kept[] (optional, may be null) receives the kept entry count per own node.  Block-wide; ends with a barrier.
__device__ __forceinline__ Lists stage_lists(const uint32_t* adj, const int32_t* eid, const int32_t* ptr, int nb, int lo,
                                             int hi, const Keep& K, bool mirror, int eb, int m_half, uint32_t* lbuf,
                                             int lcap, int* lptr, int* ws, float* invdeg_out, float* invdeg_glob) {
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
  const int n_own = hi - lo;
  Lists Ls;
  Ls.adj = adj; Ls.eid = eid; Ls.ptr = ptr; Ls.mirror = mirror; Ls.eb = eb; Ls.m_half = m_half;
  Ls.lptr = lptr;
  // pass 1: kept count per node -> lptr[i+1] (count), degree
  for (int i = warp; i < n_own; i += nwarps) {
    const int p0 = ptr[nb + lo + i], p1 = ptr[nb + lo + i + 1];
    int kept = p1 - p0;
    if (K.active) {
      kept = 0;
      for (int p = p0 + lane; p < p1; p += 32) kept += load_entry(adj, eid, p, K, mirror, eb, m_half) != DROPPED;
      kept = warp_sum_i(kept);
    }
    if (lane == 0) {
      lptr[i + 1] = kept;
      if (invdeg_out) {
        const float id = 1.f / (float)max(kept, 1);
        invdeg_out[lo + i] = id;
        invdeg_glob[nb + lo + i] = id;
      }
    }
  }
  if (tid == 0) lptr[0] = 0;
  __syncthreads();
  // exclusive scan of the counts (in place: lptr[i+1] becomes the end offset of node i)
  int running = 0;
  for (int base = 0; base < n_own; base += NT) {
    const int i = base + tid;
    const int c = i < n_own ? lptr[i + 1] : 0;
    int tot;
    const int ex = block_excl_scan_i(c, ws, &tot);
    if (i < n_own) lptr[i + 1] = running + ex + c;
    running += tot;
  }
  __syncthreads();
  const int total = n_own > 0 ? lptr[n_own] : 0;
  Ls.lst = nullptr;
  if (total <= lcap) {
    // pass 2: compacted copy (warp per node, order preserved)
    for (int i = warp; i < n_own; i += nwarps) {
      const int p0 = ptr[nb + lo + i], p1 = ptr[nb + lo + i + 1];
      int o = lptr[i];
      for (int c = p0; c < p1; c += 32) {
        const int p = c + lane;
        uint32_t ent = DROPPED;
        if (p < p1) ent = load_entry(adj, eid, p, K, mirror, eb, m_half);
        const unsigned bal = __ballot_sync(IGMC_FULL, ent != DROPPED);
        if (ent != DROPPED) lbuf[o + __popc(bal & ((1u << lane) - 1u))] = ent;
        o += __popc(bal);
      }
    }
    Ls.lst = lbuf;
  }
  __syncthreads();
  return Ls;
}

// Relation-space aggregate of the warp's nodes [base, base+cnt) (cnt <= GN, base relative to the own range)
// into its staging rows stg[s][r*inp + k] (row stride SS): 8-lane group s walks node base+s, float4 per lane.
__device__ __forceinline__ void zero_stage(float* __restrict__ stg, int SS, int lane) {
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = lane * 4; i < GN * SS; i += 128) *reinterpret_cast<float4*>(stg + i) = z4;
  __syncwarp();
}

__device__ __forceinline__ void gather_staged(const uint32_t* __restrict__ lst, const int* __restrict__ lptr, int lbase,
                                              int cnt, int lane, const float* __restrict__ feat,
                                              float* __restrict__ stg, int SS, int inp) {
  zero_stage(stg, SS, lane);
  const int q = lane & 7, s = lane >> 3;
  const int fo = 4 * q;
  int p = 0, p1 = 0;
  if (s < cnt && fo < inp) { p = lptr[lbase + s]; p1 = lptr[lbase + s + 1]; }
  float* row = stg + s * SS + fo;
  for (; p < p1; ++p) {
    const uint32_t ent = lst[p];
    const int src = (int)(ent & 0xffffu);
    const float4 a = *reinterpret_cast<const float4*>(feat + (src << 5) + (fo ^ ((src & 7) << 2)));
    float4* d = reinterpret_cast<float4*>(row + (int)(ent >> 16) * inp);
    float4 t = *d;
    t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
    *d = t;
  }
  __syncwarp();
}

// fallback: lists read from global memory, dropout draws evaluated per edge
__device__ __forceinline__ void gather_global(const Lists& Ls, const Keep& K, int nb, int gbase, int cnt, int lane,
                                              const float* __restrict__ feat, float* __restrict__ stg, int SS, int inp) {
  zero_stage(stg, SS, lane);
  const int q = lane & 7, s = lane >> 3;
  const int fo = 4 * q;
  int p = 0, p1 = 0;
  if (s < cnt && fo < inp) { p = Ls.ptr[nb + gbase + s]; p1 = Ls.ptr[nb + gbase + s + 1]; }
  float* row = stg + s * SS + fo;
  for (; p < p1; ++p) {
    const uint32_t ent = load_entry(Ls.adj, Ls.eid, p, K, Ls.mirror, Ls.eb, Ls.m_half);
    if (ent == DROPPED) continue;
    const int src = (int)(ent & 0xffffu);
    const float4 a = *reinterpret_cast<const float4*>(feat + (src << 5) + (fo ^ ((src & 7) << 2)));
    float4* d = reinterpret_cast<float4*>(row + (int)((ent >> 16) & 0xffu) * inp);
    float4 t = *d;
    t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
    *d = t;
  }
  __syncwarp();
}

// acc[j] (4 output channels c4..c4+3 of one node) += sum_kk a_row[kk] * W[kk][c4+j]; K multiple of 4,
// W row-major [K][32]
__device__ __forceinline__ void gemm_rows(const float* __restrict__ a_row, int K, const float* __restrict__ W, int c4,
                                          float (&acc)[4]) {
  const float* w = W + c4;
#pragma unroll 2
  for (int kk = 0; kk < K; kk += 4) {
    const float4 a = *reinterpret_cast<const float4*>(a_row + kk);
    const float4 w0 = *reinterpret_cast<const float4*>(w + (kk + 0) * HID);
    const float4 w1 = *reinterpret_cast<const float4*>(w + (kk + 1) * HID);
    const float4 w2 = *reinterpret_cast<const float4*>(w + (kk + 2) * HID);
    const float4 w3 = *reinterpret_cast<const float4*>(w + (kk + 3) * HID);
    acc[0] = fmaf(a.x, w0.x, acc[0]); acc[1] = fmaf(a.x, w0.y, acc[1]); acc[2] = fmaf(a.x, w0.z, acc[2]); acc[3] = fmaf(a.x, w0.w, acc[3]);
    acc[0] = fmaf(a.y, w1.x, acc[0]); acc[1] = fmaf(a.y, w1.y, acc[1]); acc[2] = fmaf(a.y, w1.z, acc[2]); acc[3] = fmaf(a.y, w1.w, acc[3]);
    acc[0] = fmaf(a.z, w2.x, acc[0]); acc[1] = fmaf(a.z, w2.y, acc[1]); acc[2] = fmaf(a.z, w2.z, acc[2]); acc[3] = fmaf(a.z, w2.w, acc[3]);
    acc[0] = fmaf(a.w, w3.x, acc[0]); acc[1] = fmaf(a.w, w3.y, acc[1]); acc[2] = fmaf(a.w, w3.z, acc[2]); acc[3] = fmaf(a.w, w3.w, acc[3]);
  }
}
// same with the A operand taken from row v of a swizzled activation tile
__device__ __forceinline__ void gemm_hrow(const float* __restrict__ Hbuf, int v, int K, const float* __restrict__ W,
                                          int c4, float (&acc)[4]) {
  const float* w = W + c4;
  const float* hrow = Hbuf + (v << 5);
  const int sw = (v & 7) << 2;
#pragma unroll 2
  for (int kk = 0; kk < K; kk += 4) {
    const float4 a = *reinterpret_cast<const float4*>(hrow + (kk ^ sw));
    const float4 w0 = *reinterpret_cast<const float4*>(w + (kk + 0) * HID);
    const float4 w1 = *reinterpret_cast<const float4*>(w + (kk + 1) * HID);
    const float4 w2 = *reinterpret_cast<const float4*>(w + (kk + 2) * HID);
    const float4 w3 = *reinterpret_cast<const float4*>(w + (kk + 3) * HID);
    acc[0] = fmaf(a.x, w0.x, acc[0]); acc[1] = fmaf(a.x, w0.y, acc[1]); acc[2] = fmaf(a.x, w0.z, acc[2]); acc[3] = fmaf(a.x, w0.w, acc[3]);
    acc[0] = fmaf(a.y, w1.x, acc[0]); acc[1] = fmaf(a.y, w1.y, acc[1]); acc[2] = fmaf(a.y, w1.z, acc[2]); acc[3] = fmaf(a.y, w1.w, acc[3]);
    acc[0] = fmaf(a.z, w2.x, acc[0]); acc[1] = fmaf(a.z, w2.y, acc[1]); acc[2] = fmaf(a.z, w2.z, acc[2]); acc[3] = fmaf(a.z, w2.w, acc[3]);
    acc[0] = fmaf(a.w, w3.x, acc[0]); acc[1] = fmaf(a.w, w3.y, acc[1]); acc[2] = fmaf(a.w, w3.z, acc[2]); acc[3] = fmaf(a.w, w3.w, acc[3]);
  }
}

#define IGMC_STAMP(i_) do { if (S.prof && threadIdx.x == 0) S.prof[(size_t)blockIdx.x * 32 + (i_)] = clock64(); } while (0)

struct Split { int lo, hi; };
__device__ __forceinline__ Split own_range(int n, int rank, int CL) {
  const int per = ((n + CL - 1) / CL + GN - 1) / GN * GN;   // multiple of the warp group
  Split s;
  s.lo = min(n, rank * per);
  s.hi = min(n, s.lo + per);
  return s;
}

// ------------------------------------------------------------------------------------------------
// per-step weight preparation:  wprep[l][0] = [W_r rows (r*inp+k) ; root rows] (forward, row-major [.][32]),
//                               wprep[l][1] = [W_r^T rows (r*32+j) ; root^T rows] (backward, layers >= 1)
// W_r = sum_b att[r,b] basis[b].  One small launch per step instead of R*NB global loads per element per CTA.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t wprep_slab(int R) { return (size_t)(R + 1) * HID * HID; }

__global__ void __launch_bounds__(256)
k_prep_weights(igmc_model_t M, const float* __restrict__ params, float* __restrict__ wprep) {
  const int l = blockIdx.x >> 1, dir = blockIdx.x & 1;
  const int R = M.num_relations, NB = M.num_bases;
  const int in = l == 0 ? M.in_dim0 : HID, inp = a4(in);
  const float* bs = params + M.off_basis[l];
  const float* at = params + M.off_att[l];
  const float* rt = params + M.off_root[l];
  float* out = wprep + ((size_t)l * 2 + dir) * wprep_slab(R);
  __shared__ float att_s[256];
  for (int i = threadIdx.x; i < R * NB && i < 256; i += 256) att_s[i] = at[i];
  __syncthreads();
  if (dir == 0) {
    const int K1 = R * inp;
    for (int idx = threadIdx.x; idx < (K1 + inp) * HID; idx += 256) {
      const int j = idx & 31, row = idx >> 5;
      float w = 0.f;
      if (row < K1) {
        const int r = row / inp, k = row - r * inp;
        if (k < in)
          for (int b = 0; b < NB; ++b) w = fmaf(att_s[r * NB + b], bs[(b * in + k) * HID + j], w);
      } else {
        const int k = row - K1;
        if (k < in) w = rt[k * HID + j];
      }
      out[idx] = w;
    }
  } else if (l > 0) {
    for (int idx = threadIdx.x; idx < (R + 1) * HID * HID; idx += 256) {   // out[(r*32+j)][k] = W_r[k][j]
      const int k = idx & 31, j = (idx >> 5) & 31, r = idx >> 10;
      float w = 0.f;
      if (r < R) {
        for (int b = 0; b < NB; ++b) w = fmaf(att_s[r * NB + b], bs[(b * HID + k) * HID + j], w);
      } else {
        w = rt[k * HID + j];
      }
      out[idx] = w;
    }
  }
}

__device__ __forceinline__ void copy_f4(float* __restrict__ dst, const float* __restrict__ src, int nfloats) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < (nfloats >> 2); i += blockDim.x) d4[i] = __ldg(s4 + i);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
k_forward_rs(igmc_model_t M, const float* __restrict__ params, const uint8_t* __restrict__ node_label,
             const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ edge_ptr, igmc_adj_t A, int n_cap,
             int lcap, igmc_dropout_t D, int training, igmc_saved_t S, const float* __restrict__ y, float loss_scale,
             float* __restrict__ dpred, float* __restrict__ sqerr, int* err) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int g = blockIdx.x / CL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5, NT = blockDim.x;
  const int L = M.num_layers, R = M.num_relations, CW = HID * L, F = 2 * CW;
  const int in0 = M.in_dim0, in0p = a4(in0);
  const int SSmax = R * HID + 4;
  const int own_cap = ((n_cap + CL - 1) / CL + GN - 1) / GN * GN;
  float* H = smem;                                   // [n_cap][32]
  float* Hn = H + (size_t)n_cap * HID;               // [n_cap][32]
  float* W = Hn + (size_t)n_cap * HID;               // [(R+1)*32][32] row-major
  float* stg_all = W + (size_t)(R + 1) * HID * HID;  // [nwarps][GN][SSmax]
  float* bias_s = stg_all + (size_t)nwarps * GN * SSmax;
  float* invdeg = bias_s + HID;                      // [n_cap]
  float* feat_s = invdeg + a4(n_cap);
  float* hid_s = feat_s + a4(F);
  int* lptr = reinterpret_cast<int*>(hid_s + L1O);              // [own_cap+1]
  uint32_t* lbuf = reinterpret_cast<uint32_t*>(lptr + a4(own_cap + 1));   // [lcap]
  __shared__ int s_t[2];
  __shared__ int ws[34];

  const int nb = node_ptr[g], n = node_ptr[g + 1] - nb;
  const int eb = edge_ptr[g], m_half = (edge_ptr[g + 1] - eb) >> 1;
  if (n > n_cap) {   // uniform over the cluster
    if (tid == 0) igmc_set_err(err, IGMC_ERR_SMEM_NODES);
    return;
  }
  const Keep K = make_keep(D, training);
  const Split own = own_range(n, rank, CL);
  IGMC_STAMP(0);

  if (tid == 0) { s_t[0] = 0x7fffffff; s_t[1] = 0x7fffffff; }
  __syncthreads();
  for (int idx = tid; idx < n * HID; idx += NT) {
    const int v = idx >> 5, c = idx & 31;
    const int lab = node_label[nb + v];
    H[hix(v, c)] = (c == lab && c < in0) ? 1.f : 0.f;
    if (c == 0 && lab == 0) atomicMin(&s_t[0], v);
    if (c == 0 && lab == 1) atomicMin(&s_t[1], v);
  }
  // in-lists of the own nodes -> shared memory; kept in-degree (dropout_adj is applied once, models.py:193)
  const Lists Ls = stage_lists(A.in_adj, A.in_eid, A.in_ptr, nb, own.lo, own.hi, K, false, eb, m_half, lbuf, lcap,
                               lptr, ws, invdeg, S.inv_deg);
  const int tu = s_t[0], ti = s_t[1];
  if (tu >= n || ti >= n) {
    if (tid == 0) igmc_set_err(err, IGMC_ERR_BAD_BATCH);
    return;
  }

  float* stg = stg_all + (size_t)warp * GN * SSmax;
  IGMC_STAMP(1);
  for (int l = 0; l < L; ++l) {
    const int inp = l == 0 ? in0p : HID;
    const int K1 = R * inp, SS = K1 + 4;
    copy_f4(W, S.wprep + (size_t)l * 2 * wprep_slab(R), (K1 + inp) * HID);
    if (tid < HID) bias_s[tid] = params[M.off_bias[l] + tid];
    __syncthreads();
    IGMC_STAMP(2 + 4 * l);
    for (int lb = warp * GN; lb < own.hi - own.lo; lb += nwarps * GN) {
      const int base = own.lo + lb;
      const int cnt = min(GN, own.hi - base);
      if (Ls.lst) gather_staged(Ls.lst, Ls.lptr, lb, cnt, lane, H, stg, SS, inp);
      else gather_global(Ls, K, nb, base, cnt, lane, H, stg, SS, inp);
      const int s = lane >> 3, c4 = (lane & 7) * 4;
      const int v = base + min(s, cnt - 1);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      gemm_rows(stg + s * SS, K1, W, c4, acc);
      const float id = invdeg[v];
      acc[0] *= id; acc[1] *= id; acc[2] *= id; acc[3] *= id;
      gemm_hrow(H, v, inp, W + K1 * HID, c4, acc);
      if (s < cnt) {
        const float4 o = make_float4(tanhf(acc[0] + bias_s[c4]), tanhf(acc[1] + bias_s[c4 + 1]),
                                     tanhf(acc[2] + bias_s[c4 + 2]), tanhf(acc[3] + bias_s[c4 + 3]));
        *reinterpret_cast<float4*>(Hn + hix(v, c4)) = o;
        __stcg(reinterpret_cast<float4*>(S.states + (size_t)(nb + v) * CW + l * HID + c4), o);
      }
      if (S.zsave) {   // 1/deg-scaled aggregate, reused by the weight-gradient GEMM
        for (int s2 = 0; s2 < cnt; ++s2) {
          const float id2 = invdeg[base + s2];
          float4* zs = reinterpret_cast<float4*>(S.zsave + ((size_t)l * S.node_cap + nb + base + s2) * (size_t)(R * HID));
          const float4* src = reinterpret_cast<const float4*>(stg + s2 * SS);
          for (int k4 = lane; k4 < (K1 >> 2); k4 += 32) {
            float4 t = src[k4];
            t.x *= id2; t.y *= id2; t.z *= id2; t.w *= id2;
            zs[k4] = t;
          }
        }
      }
      __syncwarp();
    }
    IGMC_STAMP(3 + 4 * l);
    __syncthreads();
    IGMC_STAMP(4 + 4 * l);
    // make the own rows visible to the other CTAs of the cluster, then fetch theirs
    if (CL > 1) {
      __threadfence();
      cluster.sync();
      for (int idx = tid; idx < n * 8; idx += NT) {
        const int v = idx >> 3, c4 = (idx & 7) * 4;
        if (v >= own.lo && v < own.hi) continue;
        const float4 t = __ldcg(reinterpret_cast<const float4*>(S.states + (size_t)(nb + v) * CW + l * HID + c4));
        *reinterpret_cast<float4*>(Hn + hix(v, c4)) = t;
      }
    }
    __syncthreads();
    IGMC_STAMP(5 + 4 * l);
    float* t = H; H = Hn; Hn = t;
  }

  if (rank != 0) return;
  // ---- readout (models.py:205-215), one CTA of the cluster ----
  for (int c = tid; c < F; c += NT) {
    const int node = c < CW ? tu : ti;
    const float v = __ldcg(S.states + (size_t)(nb + node) * CW + (c < CW ? c : c - CW));
    feat_s[c] = v;
    S.feat[(size_t)g * F + c] = v;
  }
  if (tid == 0) { S.target[2 * g] = nb + tu; S.target[2 * g + 1] = nb + ti; }
  __syncthreads();
  const float* W1 = params + M.off_lin1_w;
  for (int o = warp; o < L1O; o += nwarps) {
    float s = 0.f;
    for (int i = lane; i < F; i += 32) s = fmaf(W1[(size_t)o * F + i], feat_s[i], s);
    s = warp_sum_f(s);
    if (lane == 0) {
      float h = fmaxf(s + params[M.off_lin1_b + o], 0.f);
      float scale = 1.f;
      if (training && (D.hidden_dropout > 0.f || D.hidden_keep)) {
        bool keep;
        if (D.hidden_keep) keep = D.hidden_keep[(size_t)g * L1O + o] != 0;
        else {
          double t = (double)D.hidden_dropout * 4294967296.0;
          keep = edge_keep(K.seed ^ 0x5bd1e995a5a5a5a5ull, (uint32_t)(g * L1O + o),
                           t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t);
        }
        const float p = D.hidden_keep ? 0.5f : D.hidden_dropout;
        scale = keep ? 1.f / (1.f - p) : 0.f;
      }
      hid_s[o] = h * scale;
      S.hid[(size_t)g * L1O + o] = h * scale;
      S.hid_gscale[(size_t)g * L1O + o] = h > 0.f ? scale : 0.f;
    }
  }
  __syncthreads();
  if (warp == 0) {
    float s = 0.f;
    for (int o = lane; o < L1O; o += 32) s = fmaf(params[M.off_lin2_w + o], hid_s[o], s);
    s = warp_sum_f(s);
    if (lane == 0) {
      const float out = (s + params[M.off_lin2_b]) * M.multiply_by;
      S.pred[g] = out;
      if (y) {
        const float diff = out - y[g];
        if (sqerr) sqerr[g] = diff * diff;
        if (dpred) dpred[g] = 2.f * diff * loss_scale * M.multiply_by;
      }
    }
  }
  IGMC_STAMP(2 + 4 * L);
}

// K = n_own weight-gradient tile GEMM.  The block is cut into 256-thread slices that take every nsl-th node of
// a tile; inside a slice a thread owns rows {kg + 32 i, i < NRW} x channels c0..c0+3 of
//   dW[kk][j] = sum_v A[v][kk] dpre[v][j],   A[v] = [saved 1/deg-scaled AGG | h_{l-1}[v]].
// Slices are summed into dW (shared) in slice order; d bias likewise into dB (shared).
template <int NRW>
__device__ __forceinline__ void wgrad(const igmc_saved_t& S, const uint8_t* __restrict__ node_label, int l, int nb,
                                      int lo, int n_own, int K1, int inp, int in0, int CW, int R, int TS, int KR,
                                      float* __restrict__ tile, float* __restrict__ dW, float* __restrict__ dB,
                                      const float* __restrict__ DP) {
  const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
  const int nsl = NT >> 8, slice = tid >> 8, t = tid & 255;
  const int c0 = (t & 7) * 4, kg = t >> 3;
  float acc[NRW][4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NRW; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
  for (int t0 = 0; t0 < n_own; t0 += TW) {
    const int rows = min(TW, n_own - t0);
    for (int r_ = warp; r_ < rows; r_ += nwarps) {
      const int v = lo + t0 + r_;
      const float4* src = reinterpret_cast<const float4*>(S.zsave + ((size_t)l * S.node_cap + nb + v) * (size_t)(R * HID));
      float4* dst = reinterpret_cast<float4*>(tile + r_ * TS);
      for (int k4 = lane; k4 < (K1 >> 2); k4 += 32) dst[k4] = src[k4];
      if (lane < inp) {
        float hv;
        if (l > 0) hv = __ldcg(S.states + (size_t)(nb + v) * CW + (l - 1) * HID + lane);
        else hv = (lane == (int)node_label[nb + v] && lane < in0) ? 1.f : 0.f;
        tile[r_ * TS + K1 + lane] = hv;
      }
    }
    __syncthreads();
    for (int r_ = slice; r_ < rows; r_ += nsl) {
      const float4 d = *reinterpret_cast<const float4*>(DP + hix(t0 + r_, c0));
      const float* arow = tile + r_ * TS + kg;
#pragma unroll
      for (int i = 0; i < NRW; ++i) {
        if (i < NRW - 1 || kg + 32 * i < KR) {
          const float a = arow[32 * i];
          acc[i][0] = fmaf(a, d.x, acc[i][0]); acc[i][1] = fmaf(a, d.y, acc[i][1]);
          acc[i][2] = fmaf(a, d.z, acc[i][2]); acc[i][3] = fmaf(a, d.w, acc[i][3]);
        }
      }
      if (kg == 0) { accb[0] += d.x; accb[1] += d.y; accb[2] += d.z; accb[3] += d.w; }
    }
    __syncthreads();
  }
  for (int sl = 0; sl < nsl; ++sl) {
    if (slice == sl) {
#pragma unroll
      for (int i = 0; i < NRW; ++i) {
        const int kk = kg + 32 * i;
        if (kk < KR) {
          float4* o = reinterpret_cast<float4*>(dW + kk * HID + c0);
          float4 v4 = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          if (sl > 0) { const float4 p4 = *o; v4.x += p4.x; v4.y += p4.y; v4.z += p4.z; v4.w += p4.w; }
          *o = v4;
        }
      }
      if (kg == 0) {
        float4* o = reinterpret_cast<float4*>(dB + c0);
        float4 v4 = make_float4(accb[0], accb[1], accb[2], accb[3]);
        if (sl > 0) { const float4 p4 = *o; v4.x += p4.x; v4.y += p4.y; v4.z += p4.z; v4.w += p4.w; }
        *o = v4;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
k_backward_rs(igmc_model_t M, const float* __restrict__ params, const uint8_t* __restrict__ node_label,
              const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ edge_ptr, igmc_adj_t A, int n_cap,
              int lcap, igmc_dropout_t D, igmc_saved_t S, const float* __restrict__ dpred, float* __restrict__ gpart,
              float* __restrict__ dhid_out, float* __restrict__ dstate, int* err) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int g = blockIdx.x / CL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5, NT = blockDim.x;
  const int L = M.num_layers, R = M.num_relations, NB = M.num_bases, CW = HID * L, F = 2 * CW;
  const int in0 = M.in_dim0, in0p = a4(in0);
  const int SSmax = R * HID + 4;
  const int own_cap = ((n_cap + CL - 1) / CL + GN - 1) / GN * GN;
  float* DPS = smem;                                    // [n_cap][32]   dpre / deg   (all nodes)
  float* DP = DPS + (size_t)n_cap * HID;                // [own_cap][32] dpre         (own nodes)
  float* Wt = DP + (size_t)own_cap * HID;               // [(R+1)*32][32] transposed weights, row-major
  float* stg_all = Wt + (size_t)(R + 1) * HID * HID;    // [nwarps][GN][SSmax] | tile + dW + dB
  size_t stage_fl = (size_t)nwarps * GN * SSmax;        // must mirror bwd_base_fl()
  {
    const size_t need = (size_t)TW * (SSmax + HID) + ((size_t)(R + 1) * HID + 1) * HID;
    if (need > stage_fl) stage_fl = need;
  }
  float* att_s = stg_all + stage_fl;
  float* invdeg = att_s + a4(R * NB);
  float* dfeat = invdeg + a4(n_cap);
  float* dhid_s = dfeat + a4(F);
  int* lptr = reinterpret_cast<int*>(dhid_s + L1O);             // [own_cap+1]
  uint32_t* lbuf = reinterpret_cast<uint32_t*>(lptr + a4(own_cap + 1));   // [lcap]
  __shared__ int ws[34];

  const int nb = node_ptr[g], n = node_ptr[g + 1] - nb;
  const int eb = edge_ptr[g], m_half = (edge_ptr[g + 1] - eb) >> 1;
  if (n > n_cap) {
    if (tid == 0) igmc_set_err(err, IGMC_ERR_SMEM_NODES);
    return;
  }
  const Keep K = make_keep(D, 1);
  const bool sym = A.symmetric != 0;
  const Split own = own_range(n, rank, CL);
  const int n_own = own.hi - own.lo;
  const int tu = S.target[2 * g] - nb, ti = S.target[2 * g + 1] - nb;
  float* gp = gpart + ((size_t)g * CL + rank) * M.conv_param_count;
  IGMC_STAMP(0);
  // out-lists of the own nodes (symmetric batches: the in-lists with mirrored edge ids)
  const Lists Ls = stage_lists(sym ? A.in_adj : A.out_adj, sym ? A.in_eid : A.out_eid, sym ? A.in_ptr : A.out_ptr, nb,
                               own.lo, own.hi, K, sym, eb, m_half, lbuf, lcap, lptr, ws, nullptr, nullptr);

  // ---- readout backward (every CTA needs d feat to seed its target rows) ----
  const float dp = dpred[g];
  for (int o = tid; o < L1O; o += NT) {
    const float d = dp * params[M.off_lin2_w + o] * S.hid_gscale[(size_t)g * L1O + o];
    dhid_s[o] = d;
    if (rank == 0) dhid_out[(size_t)g * L1O + o] = d;
  }
  for (int v = tid; v < n; v += NT) invdeg[v] = S.inv_deg[nb + v];
  __syncthreads();
  {
    const float* W1 = params + M.off_lin1_w;
    for (int i = tid; i < F; i += NT) {
      float s = 0.f;
#pragma unroll 8
      for (int o = 0; o < L1O; ++o) s = fmaf(W1[(size_t)o * F + i], dhid_s[o], s);
      dfeat[i] = s;
    }
  }
  __syncthreads();

  float* stg = stg_all + (size_t)warp * GN * SSmax;
  IGMC_STAMP(1);
  for (int l = L - 1; l >= 0; --l) {
    const int in = l == 0 ? in0 : HID, inp = l == 0 ? in0p : HID;
    const int K1 = R * inp;
    const int sb = 2 + 5 * (L - 1 - l);
    // (0) d h_l of all nodes (top layer: readout rows only; below: exchanged through dstate),
    //     d pre = d h (1 - h^2);  DPS = d pre / deg (gather source), DP = d pre of the own rows
    for (int idx = tid; idx < n * 8; idx += NT) {
      const int v = idx >> 3, c4 = (idx & 7) * 4;
      float4 dh;
      if (l == L - 1) {
        dh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v == tu) { dh.x += dfeat[l * HID + c4]; dh.y += dfeat[l * HID + c4 + 1]; dh.z += dfeat[l * HID + c4 + 2]; dh.w += dfeat[l * HID + c4 + 3]; }
        if (v == ti) { dh.x += dfeat[CW + l * HID + c4]; dh.y += dfeat[CW + l * HID + c4 + 1]; dh.z += dfeat[CW + l * HID + c4 + 2]; dh.w += dfeat[CW + l * HID + c4 + 3]; }
      } else {
        dh = __ldcg(reinterpret_cast<const float4*>(dstate + ((size_t)l * S.node_cap + nb + v) * HID + c4));
      }
      const float4 h = __ldcg(reinterpret_cast<const float4*>(S.states + (size_t)(nb + v) * CW + l * HID + c4));
      const float4 dpre = make_float4(dh.x * (1.f - h.x * h.x), dh.y * (1.f - h.y * h.y), dh.z * (1.f - h.z * h.z),
                                      dh.w * (1.f - h.w * h.w));
      const float id = invdeg[v];
      *reinterpret_cast<float4*>(DPS + hix(v, c4)) = make_float4(dpre.x * id, dpre.y * id, dpre.z * id, dpre.w * id);
      if (v >= own.lo && v < own.hi) *reinterpret_cast<float4*>(DP + hix(v - own.lo, c4)) = dpre;
    }
    for (int idx = tid; idx < R * NB; idx += NT) att_s[idx] = params[M.off_att[l] + idx];
    if (l > 0) copy_f4(Wt, S.wprep + ((size_t)l * 2 + 1) * wprep_slab(R), (R + 1) * HID * HID);
    __syncthreads();
    IGMC_STAMP(sb);

    // (1) data gradient of the own nodes:  d h_{l-1}[u] = sum_r Q[u,r,:] W_r^T + dpre[u] root^T,
    //     Q[u,r,:] = sum_{(u->d) of type r, kept} dpre[d,:]/deg(d)
    if (l > 0) {
      const int SS = K1 + 4;
      for (int lb = warp * GN; lb < n_own; lb += nwarps * GN) {
        const int base = own.lo + lb;
        const int cnt = min(GN, own.hi - base);
        if (Ls.lst) gather_staged(Ls.lst, Ls.lptr, lb, cnt, lane, DPS, stg, SS, HID);
        else gather_global(Ls, K, nb, base, cnt, lane, DPS, stg, SS, HID);
        const int s = lane >> 3, c4 = (lane & 7) * 4;
        const int u = base + min(s, cnt - 1);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        gemm_rows(stg + s * SS, K1, Wt, c4, acc);
        gemm_hrow(DP, u - own.lo, HID, Wt + K1 * HID, c4, acc);
        if (s < cnt) {
          if (u == tu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += dfeat[(l - 1) * HID + c4 + j];
          }
          if (u == ti) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] += dfeat[CW + (l - 1) * HID + c4 + j];
          }
          __stcg(reinterpret_cast<float4*>(dstate + ((size_t)(l - 1) * S.node_cap + nb + u) * HID + c4),
                 make_float4(acc[0], acc[1], acc[2], acc[3]));
        }
        __syncwarp();
      }
      __syncthreads();
    }
    IGMC_STAMP(sb + 1);

    // (2) weight gradients over the own nodes:  dW[kk][j] = sum_v A[v][kk] dpre[v][j]
    //     A[v] = [ AGG'[v,r,k] (saved, 1/deg-scaled) | h_{l-1}[v,k] ],   rows kk < KR = (R+1)*inp
    {
      const int KR = K1 + inp, TS = KR + 4;
      float* tile = stg_all;                       // [TW][TS]
      float* dW = stg_all + TW * (SSmax + HID);    // [KR][32]
      float* dB = dW + (size_t)(R + 1) * HID * HID;  // [32]
      const int nrw = (KR + 31) >> 5;              // rows per thread (uniform): R+1 for the 32-wide layers
#define IGMC_WGRAD(N_) wgrad<N_>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, dB, DP)
      switch (nrw) {
        case 1: IGMC_WGRAD(1); break;
        case 2: IGMC_WGRAD(2); break;
        case 3: IGMC_WGRAD(3); break;
        case 4: IGMC_WGRAD(4); break;
        case 5: IGMC_WGRAD(5); break;
        case 6: IGMC_WGRAD(6); break;
        case 7: IGMC_WGRAD(7); break;
        case 8: IGMC_WGRAD(8); break;
        case 9: IGMC_WGRAD(9); break;
        case 10: IGMC_WGRAD(10); break;
        case 11: IGMC_WGRAD(11); break;
        case 12: IGMC_WGRAD(12); break;
        default: IGMC_WGRAD(13); break;
      }
#undef IGMC_WGRAD
      IGMC_STAMP(sb + 2);
      const float* bs = params + M.off_basis[l];
      // d basis[b][k][j] = sum_r att[r,b] dW_r[k][j]
      for (int row = warp; row < NB * in; row += nwarps) {
        const int b = row / in, k = row - b * in;
        float s = 0.f;
        for (int r = 0; r < R; ++r) s = fmaf(att_s[r * NB + b], dW[(r * inp + k) * HID + lane], s);
        gp[M.off_basis[l] + row * HID + lane] = s;
      }
      // d root[k][j], d bias[j]
      for (int idx = tid; idx < in * HID; idx += NT) gp[M.off_root[l] + idx] = dW[K1 * HID + idx];
      if (tid < HID) gp[M.off_bias[l] + tid] = dB[tid];
      // d att[r,b] = < dW_r , basis[b] >   (warp per (r,b), fixed-order tree)
      for (int rb = warp; rb < R * NB; rb += nwarps) {
        const int r = rb / NB, b = rb - r * NB;
        float s = 0.f;
        for (int k = 0; k < in; ++k) s = fmaf(dW[(r * inp + k) * HID + lane], bs[(b * in + k) * HID + lane], s);
        s = warp_sum_f(s);
        if (lane == 0) gp[M.off_att[l] + rb] = s;
      }
    }
    __syncthreads();
    IGMC_STAMP(sb + 3);
    // (3) d h_{l-1} rows are in dstate: publish to the cluster before the next layer reads them
    if (CL > 1) {
      __threadfence();
      cluster.sync();
    } else {
      __syncthreads();
    }
    IGMC_STAMP(sb + 4);
  }
}

size_t fwd_base_fl(int n_cap, int R, int L, int nwarps, int CL) {
  const size_t SSmax = (size_t)R * HID + 4, F = 2 * HID * L;
  const size_t own_cap = (size_t)(((n_cap + CL - 1) / CL + GN - 1) / GN * GN);
  return 2 * (size_t)n_cap * HID + (size_t)(R + 1) * HID * HID + (size_t)nwarps * GN * SSmax + HID + a4(n_cap) +
         a4((int)F) + L1O + a4((int)own_cap + 1);
}
size_t bwd_base_fl(int n_cap, int R, int NB, int L, int nwarps, int CL) {
  const size_t SSmax = (size_t)R * HID + 4, F = 2 * HID * L;
  const size_t own_cap = (size_t)(((n_cap + CL - 1) / CL + GN - 1) / GN * GN);
  size_t stage = (size_t)nwarps * GN * SSmax;
  const size_t need = (size_t)TW * (SSmax + HID) + ((size_t)(R + 1) * HID + 1) * HID;   // tile + dW + dB
  if (need > stage) stage = need;
  return (size_t)n_cap * HID + own_cap * HID + (size_t)(R + 1) * HID * HID + stage + a4(R * NB) + a4(n_cap) +
         a4((int)F) + L1O + a4((int)own_cap + 1);
}

}  // namespace rs

// ---- host-side dispatch helpers used by rgcn.cu's extern "C" entry points ----------------------------------
int rs_supported(const igmc_model_t* M) { return M->num_relations <= rs::RS_MAX_R; }

// threads per CTA, dynamic shared memory and the edge-list staging capacity for a plan
int rs_plan(const igmc_model_t* M, int n_cap, int cluster, int backward, int* threads, size_t* smem, int* lcap) {
  const size_t limit = 227 * 1024;
  const int KR = (M->num_relations + 1) * rs::HID;
  for (int nt = 1024; nt >= 256; nt >>= 1) {
    const size_t base = 4 * (backward ? rs::bwd_base_fl(n_cap, M->num_relations, M->num_bases, M->num_layers, nt >> 5, cluster)
                                      : rs::fwd_base_fl(n_cap, M->num_relations, M->num_layers, nt >> 5, cluster));
    // the weight-gradient slices are 256 threads wide and keep (R+1) x 4 accumulators per thread:
    // 64 registers (1024 threads) are enough up to R = 7
    if (nt < 256) continue;
    if (backward && nt > 512 && KR > 8 * rs::HID) continue;
    if (base + 4096 > limit) continue;
    size_t lc = (limit - base) / 4;
    if (lc > 16384) lc = 16384;
    *threads = nt;
    *lcap = (int)lc;
    *smem = base + lc * 4;
    return 0;
  }
  return -3;
}

template <class Kern, class... Args>
static int launch_cluster(Kern kern, int grid, int threads, size_t smem, int cluster, cudaStream_t st, Args... args) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) return (int)e + 1000;
  return 0;
}

int rs_forward(const igmc_model_t* M, const float* params, const uint8_t* node_label, const int32_t* node_ptr,
               const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap, const igmc_dropout_t* D, int training,
               const igmc_saved_t* S, const float* y, float loss_scale, float* dpred, float* sqerr, int cluster,
               int* err, cudaStream_t st) {
  int threads, lcap;
  size_t smem;
  int rc = rs_plan(M, n_cap, cluster, 0, &threads, &smem, &lcap);
  if (rc) return rc;
  return launch_cluster(rs::k_forward_rs, B * cluster, threads, smem, cluster, st, *M, params, node_label, node_ptr,
                        edge_ptr, *A, n_cap, lcap, *D, training, *S, y, loss_scale, dpred, sqerr, err);
}

int rs_backward(const igmc_model_t* M, const float* params, const uint8_t* node_label, const int32_t* node_ptr,
                const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap, const igmc_dropout_t* D,
                const igmc_saved_t* S, const float* dpred, float* gpart, float* dhid, float* dstate, int cluster,
                int* err, cudaStream_t st) {
  int threads, lcap;
  size_t smem;
  int rc = rs_plan(M, n_cap, cluster, 1, &threads, &smem, &lcap);
  if (rc) return rc;
  return launch_cluster(rs::k_backward_rs, B * cluster, threads, smem, cluster, st, *M, params, node_label, node_ptr,
                        edge_ptr, *A, n_cap, lcap, *D, *S, dpred, gpart, dhid, dstate, err);
}

int rs_prep_weights(const igmc_model_t* M, const float* params, float* wprep, cudaStream_t st) {
  rs::k_prep_weights<<<M->num_layers * 2, 256, 0, st>>>(*M, params, wprep);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e + 1000;
}
