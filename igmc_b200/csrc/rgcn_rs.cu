// v2 fused relational message passing: relation-space aggregate + dense transform, one thread-block
// CLUSTER per enclosing subgraph (1/2/4 CTAs split the destination nodes), 16 warps per CTA.
//
// Same math and C-ABI as csrc/rgcn.cu (which stays as the generic path for many relations); this is the
// fast path for num_relations <= RS_MAX_R.  Per layer (reference: tanh(RGCNConv) models.py:200-202,
// PyG 1.4.2 semantics SURVEY.md A.1), with W_r = sum_b att[r,b] basis[b] formed once per CTA in shared memory:
//   AGG[v,r,:] = sum_{(u->v) of type r, kept} h[u,:]            warp owns 8 destination nodes; 8-lane groups x float4
//   h'[v]      = tanh( 1/deg(v) * sum_r AGG[v,r,:] W_r + h[v] root + bias )    lane = (node, 8 output channels)
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
constexpr int GN = 8;          // destination nodes per warp group
constexpr int TW = 32;         // node tile of the weight-gradient GEMM
constexpr uint32_t DROPPED = 0xFFFFFFFFu;

__device__ __forceinline__ int hix(int v, int c) { return (v << 5) + (c ^ ((v & 7) << 2)); }
// weights [row][32] with the 4-float column groups XOR-swizzled by the row (conflict-free float4 row reads
// AND cheap transposed writes)
__device__ __forceinline__ int wix(int row, int c) { return (row << 5) + ((((c >> 2) ^ (row & 7)) << 2) | (c & 3)); }
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

// entry of node list position p, DROPPED if the (possibly mirrored) edge is dropped this step
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

// Relation-space aggregate of up to GN nodes [base, base+cnt) into the warp's staging rows
// stg[s][r*inp + k] (row stride SS).  8-lane groups walk 4 node lists concurrently, float4 per lane.
// SRC_SCALED: feature rows already carry their 1/deg factor (backward).  Returns nothing; lists are
// (type, neighbour)-sorted but no order is assumed.
__device__ __forceinline__ void gather_group(const uint32_t* __restrict__ adj, const int32_t* __restrict__ eid,
                                             const int32_t* __restrict__ ptr, int nb, int base, int cnt,
                                             const Keep& K, bool mirror, int eb, int m_half, int lane,
                                             const float* __restrict__ feat, float* __restrict__ stg, int SS,
                                             int inp, int R) {
  const int q = lane & 7, gq = lane >> 3;
  // zero the staging rows
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = lane * 4; i < GN * SS; i += 128) *reinterpret_cast<float4*>(stg + i) = z4;
  __syncwarp();
  const bool lane_on = (4 * q) < inp;
#pragma unroll
  for (int round = 0; round < GN / 4; ++round) {
    const int s = round * 4 + gq;
    int p = 0, p1 = 0;
    if (s < cnt) { p = ptr[nb + base + s]; p1 = ptr[nb + base + s + 1]; }
    float* row = stg + s * SS + 4 * q;
    for (; p < p1; ++p) {
      const uint32_t ent = load_entry(adj, eid, p, K, mirror, eb, m_half);
      if (ent == DROPPED || !lane_on) continue;
      const int src = (int)(ent & 0xffffu), ty = (int)((ent >> 16) & 0xffu);
      const float4 a = *reinterpret_cast<const float4*>(feat + hix(src, 4 * q));
      float4* d = reinterpret_cast<float4*>(row + ty * inp);
      float4 t = *d;
      t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
      *d = t;
    }
  }
  __syncwarp();
}

// acc[j] (8 output channels c8..c8+7 of node slot s) += sum_kk stg[s][kk] * W[kk][c8+j]
__device__ __forceinline__ void gemm_rows(const float* __restrict__ a_row, int K, const float* __restrict__ W,
                                          int row0, int c8, float (&acc)[8]) {
#pragma unroll 4
  for (int kk = 0; kk < K; ++kk) {
    const float a = a_row[kk];
    const float4 w0 = *reinterpret_cast<const float4*>(W + wix(row0 + kk, c8));
    const float4 w1 = *reinterpret_cast<const float4*>(W + wix(row0 + kk, c8 + 4));
    acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]); acc[2] = fmaf(a, w0.z, acc[2]);
    acc[3] = fmaf(a, w0.w, acc[3]); acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
    acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
  }
}
// same with the A operand read from a swizzled activation row
__device__ __forceinline__ void gemm_hrow(const float* __restrict__ Hbuf, int v, int K, const float* __restrict__ W,
                                          int row0, int c8, float (&acc)[8]) {
#pragma unroll 4
  for (int kk = 0; kk < K; ++kk) {
    const float a = Hbuf[hix(v, kk)];
    const float4 w0 = *reinterpret_cast<const float4*>(W + wix(row0 + kk, c8));
    const float4 w1 = *reinterpret_cast<const float4*>(W + wix(row0 + kk, c8 + 4));
    acc[0] = fmaf(a, w0.x, acc[0]); acc[1] = fmaf(a, w0.y, acc[1]); acc[2] = fmaf(a, w0.z, acc[2]);
    acc[3] = fmaf(a, w0.w, acc[3]); acc[4] = fmaf(a, w1.x, acc[4]); acc[5] = fmaf(a, w1.y, acc[5]);
    acc[6] = fmaf(a, w1.z, acc[6]); acc[7] = fmaf(a, w1.w, acc[7]);
  }
}

struct Split { int lo, hi; };
__device__ __forceinline__ Split own_range(int n, int rank, int CL) {
  const int per = ((n + CL - 1) / CL + GN - 1) / GN * GN;   // multiple of the warp group
  Split s;
  s.lo = min(n, rank * per);
  s.hi = min(n, s.lo + per);
  return s;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
k_forward_rs(igmc_model_t M, const float* __restrict__ params, const uint8_t* __restrict__ node_label,
             const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ edge_ptr, igmc_adj_t A, int n_cap,
             igmc_dropout_t D, int training, igmc_saved_t S, const float* __restrict__ y, float loss_scale,
             float* __restrict__ dpred, float* __restrict__ sqerr, int* err) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int g = blockIdx.x / CL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5, NT = blockDim.x;
  const int L = M.num_layers, R = M.num_relations, NB = M.num_bases, CW = HID * L, F = 2 * CW;
  const int in0 = M.in_dim0, in0p = a4(in0);
  const int SSmax = R * HID + 4;
  float* H = smem;                                   // [n_cap][32]
  float* Hn = H + (size_t)n_cap * HID;               // [n_cap][32]
  float* W = Hn + (size_t)n_cap * HID;               // [(R+1)*32][32]
  float* stg_all = W + (size_t)(R + 1) * HID * HID;  // [nwarps][GN][SSmax]
  float* att_s = stg_all + (size_t)nwarps * GN * SSmax;
  float* bias_s = att_s + a4(R * NB);
  float* invdeg = bias_s + HID;                      // [n_cap]
  float* feat_s = invdeg + a4(n_cap);
  float* hid_s = feat_s + a4(F);
  __shared__ int s_t[2];

  const int nb = node_ptr[g], n = node_ptr[g + 1] - nb;
  const int eb = edge_ptr[g], m_half = (edge_ptr[g + 1] - eb) >> 1;
  if (n > n_cap) {   // uniform over the cluster
    if (tid == 0) igmc_set_err(err, IGMC_ERR_SMEM_NODES);
    return;
  }
  const Keep K = make_keep(D, training);
  const Split own = own_range(n, rank, CL);

  if (tid == 0) { s_t[0] = 0x7fffffff; s_t[1] = 0x7fffffff; }
  __syncthreads();
  for (int idx = tid; idx < n * HID; idx += NT) {
    const int v = idx >> 5, c = idx & 31;
    const int lab = node_label[nb + v];
    H[hix(v, c)] = (c == lab && c < in0) ? 1.f : 0.f;
    if (c == 0 && lab == 0) atomicMin(&s_t[0], v);
    if (c == 0 && lab == 1) atomicMin(&s_t[1], v);
  }
  // kept in-degree of the own nodes (dropout_adj is applied once, models.py:193)
  for (int v = own.lo + warp; v < own.hi; v += nwarps) {
    const int p0 = A.in_ptr[nb + v], p1 = A.in_ptr[nb + v + 1];
    int kept = 0;
    if (K.active) {
      for (int p = p0 + lane; p < p1; p += 32)
        kept += load_entry(A.in_adj, A.in_eid, p, K, false, eb, m_half) != DROPPED;
      kept = warp_sum_i(kept);
    } else {
      kept = p1 - p0;
    }
    if (lane == 0) {
      const float id = 1.f / (float)max(kept, 1);
      invdeg[v] = id;
      S.inv_deg[nb + v] = id;
    }
  }
  __syncthreads();
  const int tu = s_t[0], ti = s_t[1];
  if (tu >= n || ti >= n) {
    if (tid == 0) igmc_set_err(err, IGMC_ERR_BAD_BATCH);
    return;
  }

  float* stg = stg_all + (size_t)warp * GN * SSmax;
  for (int l = 0; l < L; ++l) {
    const int in = l == 0 ? in0 : HID, inp = l == 0 ? in0p : HID;
    const int K1 = R * inp, SS = K1 + 4;
    // W_r = sum_b att[r,b] basis[b]  (rows r*inp+k), then root rows; zero rows for the k padding
    {
      const float* bs = params + M.off_basis[l];
      const float* at = params + M.off_att[l];
      const float* rt = params + M.off_root[l];
      for (int idx = tid; idx < K1 * HID; idx += NT) {
        const int j = idx & 31, row = idx >> 5, r = row / inp, k = row - r * inp;
        float w = 0.f;
        if (k < in)
          for (int b = 0; b < NB; ++b) w = fmaf(at[r * NB + b], bs[(b * in + k) * HID + j], w);
        W[wix(row, j)] = w;
      }
      for (int idx = tid; idx < inp * HID; idx += NT) {
        const int j = idx & 31, k = idx >> 5;
        W[wix(K1 + k, j)] = k < in ? rt[k * HID + j] : 0.f;
      }
      if (tid < HID) bias_s[tid] = params[M.off_bias[l] + tid];
    }
    __syncthreads();
    for (int base = own.lo + warp * GN; base < own.hi; base += nwarps * GN) {
      const int cnt = min(GN, own.hi - base);
      gather_group(A.in_adj, A.in_eid, A.in_ptr, nb, base, cnt, K, false, eb, m_half, lane, H, stg, SS, inp, R);
      const int s = lane >> 2, c8 = (lane & 3) * 8;
      const int v = base + min(s, cnt - 1);
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      gemm_rows(stg + s * SS, K1, W, 0, c8, acc);
      const float id = invdeg[v];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= id;
      gemm_hrow(H, v, inp, W, K1, c8, acc);
      if (s < cnt) {
        float4 o0 = make_float4(tanhf(acc[0] + bias_s[c8]), tanhf(acc[1] + bias_s[c8 + 1]),
                                tanhf(acc[2] + bias_s[c8 + 2]), tanhf(acc[3] + bias_s[c8 + 3]));
        float4 o1 = make_float4(tanhf(acc[4] + bias_s[c8 + 4]), tanhf(acc[5] + bias_s[c8 + 5]),
                                tanhf(acc[6] + bias_s[c8 + 6]), tanhf(acc[7] + bias_s[c8 + 7]));
        *reinterpret_cast<float4*>(Hn + hix(v, c8)) = o0;
        *reinterpret_cast<float4*>(Hn + hix(v, c8 + 4)) = o1;
        float* gs = S.states + (size_t)(nb + v) * CW + l * HID + c8;
        __stcg(reinterpret_cast<float4*>(gs), o0);
        __stcg(reinterpret_cast<float4*>(gs + 4), o1);
      }
      if (S.zsave) {   // 1/deg-scaled aggregate, reused by the weight-gradient GEMM
        for (int s2 = 0; s2 < cnt; ++s2) {
          const float id2 = invdeg[base + s2];
          float* zs = S.zsave + ((size_t)l * S.node_cap + nb + base + s2) * (size_t)(R * HID);
          for (int kk = lane; kk < K1; kk += 32) zs[kk] = stg[s2 * SS + kk] * id2;
        }
      }
      __syncwarp();
    }
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
    float* t = H; H = Hn; Hn = t;
  }

  if (rank != 0) return;
  // ---- readout (models.py:205-215), one CTA of the cluster ----
  if (CL == 1) __threadfence_block();
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
}

// K = n_own weight-gradient tile GEMM: thread owns rows {kg + i*KG, i < NR} x channels c0..c0+3 of
// dW[kk][j] = sum_v A[v][kk] dpre[v][j];  A[v] = [saved 1/deg-scaled AGG | h_{l-1}[v]].  Result -> dW (shared),
// d bias -> gbias (global).
template <int NR>
__device__ __forceinline__ void wgrad(const igmc_saved_t& S, const uint8_t* __restrict__ node_label, int l, int nb,
                                      int lo, int n_own, int K1, int inp, int in0, int CW, int R, int TS, int KR,
                                      float* __restrict__ tile, float* __restrict__ dW, const float* __restrict__ DP,
                                      float* __restrict__ gbias, int c0, int kg, int KG, int tid, int NT) {
  float acc[NR][4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NR; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
  for (int t0 = 0; t0 < n_own; t0 += TW) {
    const int rows = min(TW, n_own - t0);
    for (int idx = tid; idx < rows * K1; idx += NT) {
      const int r_ = idx / K1, kk = idx - r_ * K1;
      tile[r_ * TS + kk] = S.zsave[((size_t)l * S.node_cap + nb + lo + t0 + r_) * (size_t)(R * HID) + kk];
    }
    for (int idx = tid; idx < rows * inp; idx += NT) {
      const int r_ = idx / inp, k = idx - r_ * inp;
      const int v = lo + t0 + r_;
      float hv;
      if (l > 0) hv = __ldcg(S.states + (size_t)(nb + v) * CW + (l - 1) * HID + k);
      else hv = (k == (int)node_label[nb + v] && k < in0) ? 1.f : 0.f;
      tile[r_ * TS + K1 + k] = hv;
    }
    __syncthreads();
    for (int r_ = 0; r_ < rows; ++r_) {
      const float4 d = *reinterpret_cast<const float4*>(DP + hix(t0 + r_, c0));
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int kk = kg + i * KG;
        if (i < NR - 1 || kk < KR) {
          const float a = tile[r_ * TS + kk];
          acc[i][0] = fmaf(a, d.x, acc[i][0]); acc[i][1] = fmaf(a, d.y, acc[i][1]);
          acc[i][2] = fmaf(a, d.z, acc[i][2]); acc[i][3] = fmaf(a, d.w, acc[i][3]);
        }
      }
      if (kg == 0) { accb[0] += d.x; accb[1] += d.y; accb[2] += d.z; accb[3] += d.w; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int kk = kg + i * KG;
    if (kk < KR)
      *reinterpret_cast<float4*>(dW + kk * HID + c0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  if (kg == 0) *reinterpret_cast<float4*>(gbias + c0) = make_float4(accb[0], accb[1], accb[2], accb[3]);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
k_backward_rs(igmc_model_t M, const float* __restrict__ params, const uint8_t* __restrict__ node_label,
              const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ edge_ptr, igmc_adj_t A, int n_cap,
              igmc_dropout_t D, igmc_saved_t S, const float* __restrict__ dpred, float* __restrict__ gpart,
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
  float* Wt = DP + (size_t)own_cap * HID;               // [(R+1)*32][32] transposed weights
  float* stg_all = Wt + (size_t)(R + 1) * HID * HID;    // [nwarps][GN][SSmax]  | weight-grad tile | dW
  size_t stage_fl = (size_t)nwarps * GN * SSmax;        // must mirror bwd_smem()
  {
    const size_t need = (size_t)TW * (SSmax + HID) + ((size_t)(R + 1) * HID + 1) * HID;
    if (need > stage_fl) stage_fl = need;
  }
  float* att_s = stg_all + stage_fl;
  float* invdeg = att_s + a4(R * NB);
  float* dfeat = invdeg + a4(n_cap);
  float* dhid_s = dfeat + a4(F);

  const int nb = node_ptr[g], n = node_ptr[g + 1] - nb;
  const int eb = edge_ptr[g], m_half = (edge_ptr[g + 1] - eb) >> 1;
  if (n > n_cap) {
    if (tid == 0) igmc_set_err(err, IGMC_ERR_SMEM_NODES);
    return;
  }
  const Keep K = make_keep(D, 1);
  const bool sym = A.symmetric != 0;
  const int32_t* optr = sym ? A.in_ptr : A.out_ptr;
  const uint32_t* oadj = sym ? A.in_adj : A.out_adj;
  const int32_t* oeid = sym ? A.in_eid : A.out_eid;
  const Split own = own_range(n, rank, CL);
  const int n_own = own.hi - own.lo;
  const int tu = S.target[2 * g] - nb, ti = S.target[2 * g + 1] - nb;
  float* gp = gpart + ((size_t)g * CL + rank) * M.conv_param_count;

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
      for (int o = 0; o < L1O; ++o) s = fmaf(W1[(size_t)o * F + i], dhid_s[o], s);
      dfeat[i] = s;
    }
  }
  __syncthreads();

  float* stg = stg_all + (size_t)warp * GN * SSmax;
  for (int l = L - 1; l >= 0; --l) {
    const int in = l == 0 ? in0 : HID, inp = l == 0 ? in0p : HID;
    const int K1 = R * inp, SS = K1 + 4;
    // (0) d h_l of all nodes (top layer: readout rows only; below: exchanged through dstate),
    //     d pre = d h (1 - h^2);  DPS = d pre / deg (gather source), DP = d pre of the own rows
    for (int idx = tid; idx < n * HID; idx += NT) {
      const int v = idx >> 5, c = idx & 31;
      float dh;
      if (l == L - 1) {
        dh = 0.f;
        if (v == tu) dh += dfeat[l * HID + c];
        if (v == ti) dh += dfeat[CW + l * HID + c];
      } else {
        dh = __ldcg(dstate + ((size_t)l * S.node_cap + nb + v) * HID + c);
      }
      const float h = __ldcg(S.states + (size_t)(nb + v) * CW + l * HID + c);
      const float dpre = dh * (1.f - h * h);
      DPS[hix(v, c)] = dpre * invdeg[v];
      if (v >= own.lo && v < own.hi) DP[hix(v - own.lo, c)] = dpre;
    }
    for (int idx = tid; idx < R * NB; idx += NT) att_s[idx] = params[M.off_att[l] + idx];
    __syncthreads();

    // (1) data gradient of the own nodes:  d h_{l-1}[u] = sum_r Q[u,r,:] W_r^T + dpre[u] root^T,
    //     Q[u,r,:] = sum_{(u->d) of type r, kept} dpre[d,:]/deg(d)
    if (l > 0) {
      const float* bs = params + M.off_basis[l];
      const float* rt = params + M.off_root[l];
      for (int idx = tid; idx < R * HID * HID; idx += NT) {   // Wt[(r*32+j)][k] = W_r[k][j]
        const int j = idx & 31, k = (idx >> 5) & 31, r = idx >> 10;
        float w = 0.f;
        for (int b = 0; b < NB; ++b) w = fmaf(att_s[r * NB + b], bs[(b * HID + k) * HID + j], w);
        Wt[wix(r * HID + j, k)] = w;
      }
      for (int idx = tid; idx < HID * HID; idx += NT) {
        const int j = idx & 31, k = idx >> 5;
        Wt[wix(R * HID + j, k)] = rt[idx];
      }
      __syncthreads();
      for (int base = own.lo + warp * GN; base < own.hi; base += nwarps * GN) {
        const int cnt = min(GN, own.hi - base);
        gather_group(oadj, oeid, optr, nb, base, cnt, K, sym, eb, m_half, lane, DPS, stg, SS, HID, R);
        const int s = lane >> 2, c8 = (lane & 3) * 8;
        const int u = base + min(s, cnt - 1);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        gemm_rows(stg + s * SS, K1, Wt, 0, c8, acc);
        gemm_hrow(DP, u - own.lo, HID, Wt, K1, c8, acc);
        if (s < cnt) {
          if (u == tu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += dfeat[(l - 1) * HID + c8 + j];
          }
          if (u == ti) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += dfeat[CW + (l - 1) * HID + c8 + j];
          }
          float* gd = dstate + ((size_t)(l - 1) * S.node_cap + nb + u) * HID + c8;
          __stcg(reinterpret_cast<float4*>(gd), make_float4(acc[0], acc[1], acc[2], acc[3]));
          __stcg(reinterpret_cast<float4*>(gd + 4), make_float4(acc[4], acc[5], acc[6], acc[7]));
        }
        __syncwarp();
      }
      __syncthreads();
    }

    // (2) weight gradients over the own nodes:  dW[kk][j] = sum_v A[v][kk] dpre[v][j]
    //     A[v] = [ AGG'[v,r,k] (saved, 1/deg-scaled) | h_{l-1}[v,k] ],   rows kk < KR = (R+1)*inp
    {
      const int KR = K1 + inp, TS = KR + 4;
      float* tile = stg_all;                       // [TW][TS]
      float* dW = stg_all + TW * (SSmax + HID);    // [KR][32] (+ bias row), written after the loop
      const int c0 = (tid & 7) * 4, kg = tid >> 3, KG = NT >> 3;
      const int nr = (KR + KG - 1) / KG;           // rows per thread (uniform): 3 for R=5 at 512 threads
      switch (nr) {
        case 1: wgrad<1>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 2: wgrad<2>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 3: wgrad<3>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 4: wgrad<4>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 5: wgrad<5>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 6: wgrad<6>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 7: wgrad<7>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 8: wgrad<8>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 9: wgrad<9>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 10: wgrad<10>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        case 11: wgrad<11>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
        default: wgrad<12>(S, node_label, l, nb, own.lo, n_own, K1, inp, in0, CW, R, TS, KR, tile, dW, DP, gp + M.off_bias[l], c0, kg, KG, tid, NT); break;
      }
      __syncthreads();
      const float* bs = params + M.off_basis[l];
      // d basis[b][k][j] = sum_r att[r,b] dW_r[k][j]
      for (int idx = tid; idx < NB * in * HID; idx += NT) {
        const int j = idx & 31, k = (idx >> 5) % in, b = (idx >> 5) / in;
        float s = 0.f;
        for (int r = 0; r < R; ++r) s = fmaf(att_s[r * NB + b], dW[(r * inp + k) * HID + j], s);
        gp[M.off_basis[l] + idx] = s;
      }
      // d root[k][j]
      for (int idx = tid; idx < in * HID; idx += NT) gp[M.off_root[l] + idx] = dW[(K1 + (idx >> 5)) * HID + (idx & 31)];
      // d att[r,b] = < dW_r , basis[b] >   (warp per (r,b), fixed-order tree)
      for (int rb = warp; rb < R * NB; rb += nwarps) {
        const int r = rb / NB, b = rb - r * NB;
        float s = 0.f;
        for (int idx = lane; idx < in * HID; idx += 32)
          s = fmaf(dW[(r * inp + (idx >> 5)) * HID + (idx & 31)], bs[b * in * HID + idx], s);
        s = warp_sum_f(s);
        if (lane == 0) gp[M.off_att[l] + rb] = s;
      }
    }
    // (3) d h_{l-1} rows are in dstate: publish to the cluster before the next layer reads them
    if (CL > 1) {
      __threadfence();
      cluster.sync();
    } else {
      __threadfence_block();
      __syncthreads();
    }
  }
}

size_t fwd_smem(int n_cap, int R, int NB, int L, int nwarps) {
  const size_t SSmax = (size_t)R * HID + 4, F = 2 * HID * L;
  size_t fl = 2 * (size_t)n_cap * HID + (size_t)(R + 1) * HID * HID + (size_t)nwarps * GN * SSmax + a4(R * NB) + HID +
              a4(n_cap) + a4((int)F) + L1O;
  return fl * sizeof(float);
}
size_t bwd_smem(int n_cap, int R, int NB, int L, int nwarps, int CL) {
  const size_t SSmax = (size_t)R * HID + 4, F = 2 * HID * L;
  const size_t own_cap = (size_t)(((n_cap + CL - 1) / CL + GN - 1) / GN * GN);
  size_t stage = (size_t)nwarps * GN * SSmax;
  const size_t need = (size_t)TW * (SSmax + HID) + ((size_t)(R + 1) * HID + 1) * HID;   // tile + dW
  if (need > stage) stage = need;
  size_t fl = (size_t)n_cap * HID + own_cap * HID + (size_t)(R + 1) * HID * HID + stage + a4(R * NB) + a4(n_cap) +
              a4((int)F) + L1O;
  return fl * sizeof(float);
}

}  // namespace rs

// ---- host-side dispatch helpers used by rgcn.cu's extern "C" entry points ----------------------------------
int rs_supported(const igmc_model_t* M) { return M->num_relations <= rs::RS_MAX_R; }

int rs_plan(const igmc_model_t* M, int n_cap, int cluster, int backward, int* threads, size_t* smem) {
  const size_t limit = 227 * 1024;
  for (int nt = 512; nt >= 128; nt >>= 1) {
    const size_t b = backward ? rs::bwd_smem(n_cap, M->num_relations, M->num_bases, M->num_layers, nt >> 5, cluster)
                              : rs::fwd_smem(n_cap, M->num_relations, M->num_bases, M->num_layers, nt >> 5);
    // the weight-gradient mapping needs ceil(KR / (threads/8)) <= 12
    const int KR = (M->num_relations + 1) * rs::HID;
    if (backward && (KR + (nt >> 3) - 1) / (nt >> 3) > 12) continue;
    if (b <= limit) { *threads = nt; *smem = b; return 0; }
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
  int threads;
  size_t smem;
  int rc = rs_plan(M, n_cap, cluster, 0, &threads, &smem);
  if (rc) return rc;
  return launch_cluster(rs::k_forward_rs, B * cluster, threads, smem, cluster, st, *M, params, node_label, node_ptr,
                        edge_ptr, *A, n_cap, *D, training, *S, y, loss_scale, dpred, sqerr, err);
}

int rs_backward(const igmc_model_t* M, const float* params, const uint8_t* node_label, const int32_t* node_ptr,
                const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap, const igmc_dropout_t* D,
                const igmc_saved_t* S, const float* dpred, float* gpart, float* dhid, float* dstate, int cluster,
                int* err, cudaStream_t st) {
  int threads;
  size_t smem;
  int rc = rs_plan(M, n_cap, cluster, 1, &threads, &smem);
  if (rc) return rc;
  return launch_cluster(rs::k_backward_rs, B * cluster, threads, smem, cluster, st, *M, params, node_label, node_ptr,
                        edge_ptr, *A, n_cap, *D, *S, dpred, gpart, dhid, dstate, err);
}
