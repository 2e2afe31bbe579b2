// Fused relational message passing for IGMC: one CTA per enclosing subgraph, node features resident
// in shared memory for all layers.
//
// Replaces (reference call sites): dropout_adj (models.py:193-198), 4x tanh(RGCNConv) (models.py:200-202;
// PyG 1.4.2 RGCNConv = index_select of a per-edge [in,out] weight + bmm + scatter_mean),
// concat + target-row readout + lin1/relu/dropout/lin2 (models.py:203-215) and their autograd.
//
// Formulation (never materialises the per-edge weight):  with W_r = sum_b att[r,b] basis[b],
//   Z[v,b,:]  = 1/deg(v) * sum_{(u->v, type r) kept} att[r,b] * h[u,:]          (segment reduce, run per type)
//   pre[v,:]  = [Z[v] | h[v]] . [basis ; root] + bias ,   h'[v] = tanh(pre[v])     (dense K=(NB+1)*in GEMM tile)
// deg(v) = number of kept incoming edges of ANY type (aggr='mean').  Backward is the same pair of
// passes on the transposed lists with [basis ; root]^T, plus K=n weight-gradient GEMM tiles.
// All reductions have a fixed order: results are run-to-run deterministic.
#include "common.cuh"
#include "../../include/igmc_b200.h"

namespace {

constexpr int MP_THREADS = 256;
constexpr int MP_WARPS = MP_THREADS / 32;
constexpr int TM = 64;   // node tile of the dense phases
constexpr int WS = 36;   // row stride of transposed weights in backward
constexpr int HID = IGMC_HIDDEN;
constexpr int L1O = IGMC_LIN1_OUT;
constexpr uint32_t DROPPED = 0xFFFFFFFFu;

// Activation tiles [n][32] live in shared memory with an XOR swizzle of the column by the row:
// a row stays one contiguous 128 B line (lane = channel: conflict-free, float4 groups intact) and
// a fixed channel over 8 consecutive rows hits 8 different banks (column reads of the dense phases).
__device__ __forceinline__ int hix(int v, int c) { return (v << 5) + (c ^ ((v & 7) << 2)); }

struct EdgeKeep {
  bool active;
  const uint8_t* mask;
  uint64_t seed;
  uint32_t thresh;
  __device__ __forceinline__ bool keep(int e) const {
    return mask ? (mask[e] != 0) : edge_keep(seed, (uint32_t)e, thresh);
  }
};

__device__ __forceinline__ EdgeKeep make_keep(const igmc_dropout_t& D, int training) {
  EdgeKeep K;
  K.mask = D.edge_keep;
  K.seed = D.seed_dev ? *D.seed_dev : D.seed;
  K.active = training && (D.adj_dropout > 0.0f || D.edge_keep != nullptr);
  double t = (double)D.adj_dropout * 4294967296.0;
  K.thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  return K;
}

// Walks list [p0,p1) of one node with the whole warp; body(nbr, type) is called warp-uniformly for
// every kept entry, in list order.  `mirror`: the entry describes edge nbr->x but the caller wants
// the reverse edge x->nbr of a symmetric batch, whose id is the mirror image inside the graph.
template <class Body>
__device__ __forceinline__ int list_foreach(const uint32_t* __restrict__ adj, const int32_t* __restrict__ eid,
                                            int p0, int p1, const EdgeKeep& K, bool mirror, int eb, int m_half,
                                            int lane, Body&& body) {
  int kept = 0;
  for (int c = p0; c < p1; c += 32) {
    const int p = c + lane;
    uint32_t ent = DROPPED;
    if (p < p1) {
      ent = adj[p];
      if (K.active) {
        int e = eid[p];
        if (mirror) { const int el = e - eb; e = eb + (el < m_half ? el + m_half : el - m_half); }
        if (!K.keep(e)) ent = DROPPED;
      }
    }
    const int cnt = min(32, p1 - c);
    for (int q = 0; q < cnt; ++q) {
      const uint32_t en = __shfl_sync(IGMC_FULL, ent, q);
      if (en == DROPPED) continue;
      ++kept;
      body((int)(en & 0xffffu), (int)((en >> 16) & 0xffu));
    }
  }
  return kept;
}

// Type-run segment reduce: acc[b] = sum over kept entries of att[type][b] * feat[nbr][lane] * scale[nbr]
template <int NB, bool SCALE>
__device__ __forceinline__ int gather_runs(const uint32_t* __restrict__ adj, const int32_t* __restrict__ eid,
                                           int p0, int p1, const EdgeKeep& K, bool mirror, int eb, int m_half,
                                           int lane, const float* __restrict__ feat, const float* __restrict__ scale,
                                           const float* __restrict__ att_s, float (&acc)[NB]) {
  float run = 0.f;
  int cur = -1;
#pragma unroll
  for (int b = 0; b < NB; ++b) acc[b] = 0.f;
  const int kept = list_foreach(adj, eid, p0, p1, K, mirror, eb, m_half, lane, [&](int nbr, int ty) {
    if (ty != cur) {
      if (cur >= 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] = fmaf(att_s[cur * NB + b], run, acc[b]);
      }
      run = 0.f;
      cur = ty;
    }
    float v = feat[hix(nbr, lane)];
    if (SCALE) v *= scale[nbr];
    run += v;
  });
  if (cur >= 0) {
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = fmaf(att_s[cur * NB + b], run, acc[b]);
  }
  return kept;
}

__device__ __forceinline__ size_t align4(size_t x) { return (x + 3) & ~(size_t)3; }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(MP_THREADS)
k_forward(igmc_model_t M, const float* __restrict__ params, const uint8_t* __restrict__ node_label,
          const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ edge_ptr, igmc_adj_t A, int n_cap,
          igmc_dropout_t D, int training, igmc_saved_t S, const float* __restrict__ y, float loss_scale,
          float* __restrict__ dpred, float* __restrict__ sqerr, int* err) {
  extern __shared__ __align__(16) float smem[];
  constexpr int ZS = NB * HID + 4;
  const int L = M.num_layers, R = M.num_relations, CW = HID * L, F = 2 * CW;
  float* HA = smem;
  float* HB = HA + (size_t)n_cap * HID;
  float* Zs = HB + (size_t)n_cap * HID;
  float* Wc = Zs + TM * ZS;
  float* att_s = Wc + (NB + 1) * HID * HID;
  float* bias_s = att_s + align4((size_t)R * NB);
  float* invdeg = bias_s + HID;
  float* feat_s = invdeg + align4(n_cap);
  float* hid_s = feat_s + align4(F);
  __shared__ int s_t[2];

  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = node_ptr[g], n = node_ptr[g + 1] - nb;
  const int eb = edge_ptr[g], m_half = (edge_ptr[g + 1] - eb) >> 1;
  if (n > n_cap) {
    if (tid == 0) igmc_set_err(err, IGMC_ERR_SMEM_NODES);
    return;
  }
  const EdgeKeep K = make_keep(D, training);
  const int in0 = M.in_dim0;

  // h_0 = one-hot(node label) (util_functions.py:286) padded to 32 columns; target rows
  if (tid == 0) { s_t[0] = 0x7fffffff; s_t[1] = 0x7fffffff; }
  __syncthreads();
  for (int idx = tid; idx < n * HID; idx += MP_THREADS) {
    const int v = idx >> 5, c = idx & 31;
    const int lab = node_label[nb + v];
    HA[hix(v, c)] = (c == lab && c < in0) ? 1.f : 0.f;
    if (c == 0 && lab == 0) atomicMin(&s_t[0], v);   // users = x[:,0]==1 (models.py:205)
    if (c == 0 && lab == 1) atomicMin(&s_t[1], v);   // items = x[:,1]==1 (models.py:206)
  }
  // kept in-degree (same for every layer: dropout_adj is applied once, models.py:193)
  for (int v = warp; v < n; v += MP_WARPS) {
    const int p0 = A.in_ptr[nb + v], p1 = A.in_ptr[nb + v + 1];
    int kept = p1 - p0;
    if (K.active) kept = list_foreach(A.in_adj, A.in_eid, p0, p1, K, false, eb, m_half, lane, [](int, int) {});
    if (lane == 0) {
      const float id = 1.f / (float)max(kept, 1);
      invdeg[v] = id;
      S.inv_deg[nb + v] = id;
    }
  }
  __syncthreads();
  const int tu = s_t[0], ti = s_t[1];
  if (M.readout == 0 && (tu >= n || ti >= n)) {
    if (tid == 0) igmc_set_err(err, IGMC_ERR_BAD_BATCH);
    return;
  }

  for (int l = 0; l < L; ++l) {
    const int in = l == 0 ? in0 : HID;
    const int K1 = NB * in;
    // [basis ; root] rows, att, bias -> shared
    {
      const float* bs = params + M.off_basis[l];
      const float* rt = params + M.off_root[l];
      for (int idx = tid; idx < K1 * HID; idx += MP_THREADS) Wc[idx] = bs[idx];
      for (int idx = tid; idx < in * HID; idx += MP_THREADS) Wc[K1 * HID + idx] = rt[idx];
      for (int idx = tid; idx < R * NB; idx += MP_THREADS) att_s[idx] = params[M.off_att[l] + idx];
      if (tid < HID) bias_s[tid] = params[M.off_bias[l] + tid];
    }
    __syncthreads();
    for (int t0 = 0; t0 < n; t0 += TM) {
      const int tend = min(t0 + TM, n);
      // ---- phase G: basis-space aggregate of the tile's nodes (warp per destination node) ----
      for (int v = t0 + warp; v < tend; v += MP_WARPS) {
        float z[NB];
        gather_runs<NB, false>(A.in_adj, A.in_eid, A.in_ptr[nb + v], A.in_ptr[nb + v + 1], K, false, eb, m_half,
                               lane, HA, nullptr, att_s, z);
        const float id = invdeg[v];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          z[b] *= id;
          if (lane < in) Zs[(v - t0) * ZS + b * in + lane] = z[b];
          if (S.zsave)
            S.zsave[((size_t)l * S.node_cap + nb + v) * (NB * HID) + b * HID + lane] = z[b];
        }
      }
      __syncthreads();
      // ---- phase T: [Z | h] . [basis ; root] + bias, tanh (thread = 2 nodes x 4 channels) ----
      {
        const int r0 = (tid >> 3) * 2, c0 = (tid & 7) * 4;
        float a0[4], a1[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { a0[c] = bias_s[c0 + c]; a1[c] = a0[c]; }
        const float* z0 = Zs + r0 * ZS;
        const float* z1 = z0 + ZS;
        for (int kk = 0; kk < K1; ++kk) {
          const float x0 = z0[kk], x1 = z1[kk];
          const float4 w = *reinterpret_cast<const float4*>(Wc + kk * HID + c0);
          a0[0] = fmaf(x0, w.x, a0[0]); a0[1] = fmaf(x0, w.y, a0[1]); a0[2] = fmaf(x0, w.z, a0[2]); a0[3] = fmaf(x0, w.w, a0[3]);
          a1[0] = fmaf(x1, w.x, a1[0]); a1[1] = fmaf(x1, w.y, a1[1]); a1[2] = fmaf(x1, w.z, a1[2]); a1[3] = fmaf(x1, w.w, a1[3]);
        }
        const int v0 = t0 + r0, v1 = v0 + 1;
        const int q0 = min(v0, n - 1), q1 = min(v1, n - 1);
        for (int k = 0; k < in; ++k) {
          const float x0 = HA[hix(q0, k)], x1 = HA[hix(q1, k)];
          const float4 w = *reinterpret_cast<const float4*>(Wc + (K1 + k) * HID + c0);
          a0[0] = fmaf(x0, w.x, a0[0]); a0[1] = fmaf(x0, w.y, a0[1]); a0[2] = fmaf(x0, w.z, a0[2]); a0[3] = fmaf(x0, w.w, a0[3]);
          a1[0] = fmaf(x1, w.x, a1[0]); a1[1] = fmaf(x1, w.y, a1[1]); a1[2] = fmaf(x1, w.z, a1[2]); a1[3] = fmaf(x1, w.w, a1[3]);
        }
        if (v0 < tend) {
          float4 o = make_float4(tanhf(a0[0]), tanhf(a0[1]), tanhf(a0[2]), tanhf(a0[3]));
          *reinterpret_cast<float4*>(HB + hix(v0, c0)) = o;
          *reinterpret_cast<float4*>(S.states + (size_t)(nb + v0) * CW + l * HID + c0) = o;
        }
        if (v1 < tend) {
          float4 o = make_float4(tanhf(a1[0]), tanhf(a1[1]), tanhf(a1[2]), tanhf(a1[3]));
          *reinterpret_cast<float4*>(HB + hix(v1, c0)) = o;
          *reinterpret_cast<float4*>(S.states + (size_t)(nb + v1) * CW + l * HID + c0) = o;
        }
      }
      __syncthreads();
    }
    float* t = HA; HA = HB; HB = t;
  }

  if (M.readout != 0) return;   // concat_states only: an external readout (csrc/sortpool.cu) takes over
  // ---- readout: concat rows of the target user and item (models.py:205-207) ----
  for (int c = tid; c < F; c += MP_THREADS) {
    const int node = c < CW ? tu : ti;
    const float v = S.states[(size_t)(nb + node) * CW + (c < CW ? c : c - CW)];
    feat_s[c] = v;
    S.feat[(size_t)g * F + c] = v;
  }
  if (tid == 0) { S.target[2 * g] = nb + tu; S.target[2 * g + 1] = nb + ti; }
  __syncthreads();
  const float* W1 = params + M.off_lin1_w;
  for (int o = warp; o < L1O; o += MP_WARPS) {
    float s = 0.f;
    for (int i = lane; i < F; i += 32) s = fmaf(W1[(size_t)o * F + i], feat_s[i], s);
    s = warp_sum_f(s);
    if (lane == 0) {
      float h = fmaxf(s + params[M.off_lin1_b + o], 0.f);                 // relu(lin1) models.py:211
      float scale = 1.f;
      if (training && (D.hidden_dropout > 0.f || D.hidden_keep)) {        // F.dropout models.py:212
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
      const float out = (s + params[M.off_lin2_b]) * M.multiply_by;       // models.py:213-215
      S.pred[g] = out;
      if (y) {
        const float diff = out - y[g];
        if (sqerr) sqerr[g] = diff * diff;
        if (dpred) dpred[g] = 2.f * diff * loss_scale * M.multiply_by;   // d mse / d lin2-output
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(MP_THREADS)
k_backward(igmc_model_t M, const float* __restrict__ params, const uint8_t* __restrict__ node_label,
           const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ edge_ptr, igmc_adj_t A, int n_cap,
           igmc_dropout_t D, igmc_saved_t S, const float* __restrict__ dpred, float* __restrict__ gpart,
           float* __restrict__ dhid_out, int* err) {
  extern __shared__ __align__(16) float smem[];
  constexpr int ZS = NB * HID + 4;
  const int L = M.num_layers, R = M.num_relations, CW = HID * L, F = 2 * CW;
  float* GA = smem;                               // d h_l  -> d pre_l
  float* GB = GA + (size_t)n_cap * HID;           // d h_{l-1}
  float* HP = GB + (size_t)n_cap * HID;           // h_{l-1}
  float* Zs = HP + (size_t)n_cap * HID;           // tile scratch [TM][ZS]
  float* Wt = Zs + TM * ZS;                       // transposed weights, (NB+1)*32 rows x WS  |  W2 [32][ZS]
  float* att_s = Wt + (NB + 1) * HID * WS;
  float* invdeg = att_s + align4((size_t)R * NB);
  float* dfeat = invdeg + align4(n_cap);
  float* dhid_s = dfeat + align4(F);
  float* datt_w = dhid_s + L1O;                   // [MP_WARPS][R*NB]

  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = node_ptr[g], n = node_ptr[g + 1] - nb;
  const int eb = edge_ptr[g], m_half = (edge_ptr[g + 1] - eb) >> 1;
  if (n > n_cap) {
    if (tid == 0) igmc_set_err(err, IGMC_ERR_SMEM_NODES);
    return;
  }
  const EdgeKeep K = make_keep(D, 1);
  const bool sym = A.symmetric != 0;
  const int32_t* optr = sym ? A.in_ptr : A.out_ptr;
  const uint32_t* oadj = sym ? A.in_adj : A.out_adj;
  const int32_t* oeid = sym ? A.in_eid : A.out_eid;
  const int in0 = M.in_dim0;
  const bool ext = M.readout != 0;   // d concat_states comes from an external readout (S.dstate)
  const int tu = ext ? -1 : S.target[2 * g] - nb, ti = ext ? -1 : S.target[2 * g + 1] - nb;
  float* gp = gpart + (size_t)g * M.conv_param_count;

  // ---- readout backward: d hid, d feat (models.py:211-213) ----
  if (!ext) {
    const float dp = dpred[g];
    for (int o = tid; o < L1O; o += MP_THREADS) {
      const float d = dp * params[M.off_lin2_w + o] * S.hid_gscale[(size_t)g * L1O + o];
      dhid_s[o] = d;
      dhid_out[(size_t)g * L1O + o] = d;
    }
  }
  for (int v = tid; v < n; v += MP_THREADS) invdeg[v] = S.inv_deg[nb + v];
  __syncthreads();
  if (!ext) {
    const float* W1 = params + M.off_lin1_w;
    for (int i = tid; i < F; i += MP_THREADS) {
      float s = 0.f;
      for (int o = 0; o < L1O; ++o) s = fmaf(W1[(size_t)o * F + i], dhid_s[o], s);
      dfeat[i] = s;
    }
  }
  __syncthreads();
  // d h_L : only the two target rows receive gradient from the IGMC readout
  for (int idx = tid; idx < n * HID; idx += MP_THREADS) {
    const int v = idx >> 5, c = idx & 31;
    float gval = 0.f;
    if (ext) gval = S.dstate[(size_t)(nb + v) * CW + (L - 1) * HID + c];
    if (v == tu) gval += dfeat[(L - 1) * HID + c];
    if (v == ti) gval += dfeat[CW + (L - 1) * HID + c];
    GA[hix(v, c)] = gval;
  }
  __syncthreads();

  for (int l = L - 1; l >= 0; --l) {
    const int in = l == 0 ? in0 : HID;
    // (0) d pre = d h * (1 - h^2) ; h_{l-1} ; d h_{l-1} seeded with the readout rows
    for (int idx = tid; idx < n * HID; idx += MP_THREADS) {
      const int v = idx >> 5, c = idx & 31;
      const float h = S.states[(size_t)(nb + v) * CW + l * HID + c];
      GA[hix(v, c)] *= (1.f - h * h);
      float hp, gval = 0.f;
      if (l > 0) {
        hp = S.states[(size_t)(nb + v) * CW + (l - 1) * HID + c];
        if (ext) gval = S.dstate[(size_t)(nb + v) * CW + (l - 1) * HID + c];
        if (v == tu) gval += dfeat[(l - 1) * HID + c];
        if (v == ti) gval += dfeat[CW + (l - 1) * HID + c];
      } else {
        hp = (c == (int)node_label[nb + v] && c < in0) ? 1.f : 0.f;
      }
      HP[hix(v, c)] = hp;
      GB[hix(v, c)] = gval;
    }
    for (int idx = tid; idx < R * NB; idx += MP_THREADS) att_s[idx] = params[M.off_att[l] + idx];
    for (int idx = tid; idx < MP_WARPS * R * NB; idx += MP_THREADS) datt_w[idx] = 0.f;
    __syncthreads();

    // (1) data gradient: d h_{l-1} += [G | dpre] . [basis ; root]^T , G = out-list aggregate of dpre/deg
    if (l > 0) {
      const float* bs = params + M.off_basis[l];
      const float* rt = params + M.off_root[l];
      for (int idx = tid; idx < NB * HID * HID; idx += MP_THREADS) {   // basis[b][k][j] -> Wt[(b*32+j)][k]
        const int j = idx & 31, k = (idx >> 5) & 31, b = idx >> 10;
        Wt[(b * HID + j) * WS + k] = bs[idx];
      }
      for (int idx = tid; idx < HID * HID; idx += MP_THREADS) {        // root[k][j] -> Wt[(NB*32+j)][k]
        const int j = idx & 31, k = idx >> 5;
        Wt[(NB * HID + j) * WS + k] = rt[idx];
      }
      __syncthreads();
      for (int t0 = 0; t0 < n; t0 += TM) {
        const int tend = min(t0 + TM, n);
        for (int u = t0 + warp; u < tend; u += MP_WARPS) {
          float gz[NB];
          gather_runs<NB, true>(oadj, oeid, optr[nb + u], optr[nb + u + 1], K, sym, eb, m_half, lane, GA, invdeg,
                                att_s, gz);
#pragma unroll
          for (int b = 0; b < NB; ++b) Zs[(u - t0) * ZS + b * HID + lane] = gz[b];
        }
        __syncthreads();
        {
          const int r0 = (tid >> 3) * 2, c0 = (tid & 7) * 4;
          float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
          const float* z0 = Zs + r0 * ZS;
          const float* z1 = z0 + ZS;
          for (int kk = 0; kk < NB * HID; ++kk) {
            const float x0 = z0[kk], x1 = z1[kk];
            const float4 w = *reinterpret_cast<const float4*>(Wt + kk * WS + c0);
            a0[0] = fmaf(x0, w.x, a0[0]); a0[1] = fmaf(x0, w.y, a0[1]); a0[2] = fmaf(x0, w.z, a0[2]); a0[3] = fmaf(x0, w.w, a0[3]);
            a1[0] = fmaf(x1, w.x, a1[0]); a1[1] = fmaf(x1, w.y, a1[1]); a1[2] = fmaf(x1, w.z, a1[2]); a1[3] = fmaf(x1, w.w, a1[3]);
          }
          const int v0 = t0 + r0, v1 = v0 + 1;
          const int q0 = min(v0, n - 1), q1 = min(v1, n - 1);
          for (int j = 0; j < HID; ++j) {
            const float x0 = GA[hix(q0, j)], x1 = GA[hix(q1, j)];
            const float4 w = *reinterpret_cast<const float4*>(Wt + (NB * HID + j) * WS + c0);
            a0[0] = fmaf(x0, w.x, a0[0]); a0[1] = fmaf(x0, w.y, a0[1]); a0[2] = fmaf(x0, w.z, a0[2]); a0[3] = fmaf(x0, w.w, a0[3]);
            a1[0] = fmaf(x1, w.x, a1[0]); a1[1] = fmaf(x1, w.y, a1[1]); a1[2] = fmaf(x1, w.z, a1[2]); a1[3] = fmaf(x1, w.w, a1[3]);
          }
          if (v0 < tend) {
            float4* o = reinterpret_cast<float4*>(GB + hix(v0, c0));
            float4 t = *o; t.x += a0[0]; t.y += a0[1]; t.z += a0[2]; t.w += a0[3]; *o = t;
          }
          if (v1 < tend) {
            float4* o = reinterpret_cast<float4*>(GB + hix(v1, c0));
            float4 t = *o; t.x += a1[0]; t.y += a1[1]; t.z += a1[2]; t.w += a1[3]; *o = t;
          }
        }
        __syncthreads();
      }
    }

    // (2) weight gradients: d[basis ; root] = [Z | h_{l-1}]^T . dpre  (K = n), d bias = sum_v dpre
    {
      const int kg = tid >> 3, c0 = (tid & 7) * 4;   // thread: rows {kg + 32 i}, channels c0..c0+3
      float acc[NB + 1][4];
      float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i <= NB; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
      for (int t0 = 0; t0 < n; t0 += TM) {
        const int tend = min(t0 + TM, n), rows = tend - t0;
        for (int idx = tid; idx < rows * (NB * HID); idx += MP_THREADS) {
          const int r = idx / (NB * HID), c = idx - r * (NB * HID);
          Zs[r * ZS + c] = S.zsave[((size_t)l * S.node_cap + nb + t0 + r) * (NB * HID) + c];
        }
        __syncthreads();
        if (kg < in) {
          for (int r = 0; r < rows; ++r) {
            const float4 d = *reinterpret_cast<const float4*>(GA + hix(t0 + r, c0));
#pragma unroll
            for (int i = 0; i < NB; ++i) {
              const float zv = Zs[r * ZS + i * HID + kg];
              acc[i][0] = fmaf(zv, d.x, acc[i][0]); acc[i][1] = fmaf(zv, d.y, acc[i][1]);
              acc[i][2] = fmaf(zv, d.z, acc[i][2]); acc[i][3] = fmaf(zv, d.w, acc[i][3]);
            }
            const float hv = HP[hix(t0 + r, kg)];
            acc[NB][0] = fmaf(hv, d.x, acc[NB][0]); acc[NB][1] = fmaf(hv, d.y, acc[NB][1]);
            acc[NB][2] = fmaf(hv, d.z, acc[NB][2]); acc[NB][3] = fmaf(hv, d.w, acc[NB][3]);
            if (kg == 0) { accb[0] += d.x; accb[1] += d.y; accb[2] += d.z; accb[3] += d.w; }
          }
        }
        __syncthreads();
      }
      if (kg < in) {
#pragma unroll
        for (int i = 0; i < NB; ++i)
          *reinterpret_cast<float4*>(gp + M.off_basis[l] + (i * in + kg) * HID + c0) =
              make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(gp + M.off_root[l] + kg * HID + c0) =
            make_float4(acc[NB][0], acc[NB][1], acc[NB][2], acc[NB][3]);
        if (kg == 0)
          *reinterpret_cast<float4*>(gp + M.off_bias[l] + c0) = make_float4(accb[0], accb[1], accb[2], accb[3]);
      }
    }

    // (3) d att[r,b] = sum_v sum_{in-runs of type r} < sum_run h_{l-1}[src,:] , dZ'[v,b,:] >,
    //     dZ'[v,b,k] = 1/deg(v) * sum_j dpre[v,j] basis[b][k][j]
    {
      const float* bs = params + M.off_basis[l];
      for (int idx = tid; idx < NB * in * HID; idx += MP_THREADS) {   // basis[b][k][j] -> W2[j][b*32+k]
        const int j = idx & 31, k = (idx >> 5) % in, b = (idx >> 5) / in;
        Wt[j * ZS + b * HID + k] = bs[idx];
      }
      __syncthreads();
      for (int t0 = 0; t0 < n; t0 += TM) {
        const int tend = min(t0 + TM, n);
        {
          // thread: node r = tid/4, columns [cg*8NB, (cg+1)*8NB) ; only k < in are meaningful
          const int r = tid >> 2, cg = tid & 3;
          constexpr int CPT = 8 * NB;
          float a[CPT];
#pragma unroll
          for (int c = 0; c < CPT; ++c) a[c] = 0.f;
          const int v = t0 + r;
          if (v < tend) {
            for (int j = 0; j < HID; ++j) {
              const float x = GA[hix(v, j)];
              const float* w = Wt + j * ZS + cg * CPT;
#pragma unroll
              for (int c = 0; c < CPT; c += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(w + c);
                a[c] = fmaf(x, w4.x, a[c]); a[c + 1] = fmaf(x, w4.y, a[c + 1]);
                a[c + 2] = fmaf(x, w4.z, a[c + 2]); a[c + 3] = fmaf(x, w4.w, a[c + 3]);
              }
            }
            const float id = invdeg[v];
#pragma unroll
            for (int c = 0; c < CPT; ++c) Zs[r * ZS + cg * CPT + c] = a[c] * id;
          }
        }
        __syncthreads();
        for (int v = t0 + warp; v < tend; v += MP_WARPS) {
          float dz[NB];
#pragma unroll
          for (int b = 0; b < NB; ++b) dz[b] = (lane < in) ? Zs[(v - t0) * ZS + b * HID + lane] : 0.f;
          float run = 0.f;
          int cur = -1;
          float* dw = datt_w + warp * (R * NB);
          auto flush = [&]() {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const float p = warp_sum_f(run * dz[b]);
              if (lane == 0) dw[cur * NB + b] += p;
            }
          };
          list_foreach(A.in_adj, A.in_eid, A.in_ptr[nb + v], A.in_ptr[nb + v + 1], K, false, eb, m_half, lane,
                       [&](int nbr, int ty) {
                         if (ty != cur) {
                           if (cur >= 0) flush();
                           run = 0.f;
                           cur = ty;
                         }
                         run += HP[hix(nbr, lane)];
                       });
          if (cur >= 0) flush();
        }
        __syncthreads();
      }
      for (int idx = tid; idx < R * NB; idx += MP_THREADS) {
        float s = 0.f;
        for (int w = 0; w < MP_WARPS; ++w) s += datt_w[w * (R * NB) + idx];
        gp[M.off_att[l] + idx] = s;
      }
    }
    __syncthreads();
    float* t = GA; GA = GB; GB = t;
  }
}

size_t fwd_smem_bytes(int n_cap, int R, int NB, int L) {
  const size_t ZS = NB * HID + 4, F = 2 * HID * L;
  size_t fl = 2 * (size_t)n_cap * HID + TM * ZS + (size_t)(NB + 1) * HID * HID + ((R * NB + 3) & ~3) + HID +
              ((n_cap + 3) & ~3) + ((F + 3) & ~3) + L1O;
  return fl * sizeof(float);
}
size_t bwd_smem_bytes(int n_cap, int R, int NB, int L) {
  const size_t ZS = NB * HID + 4, F = 2 * HID * L;
  size_t wt = (size_t)(NB + 1) * HID * WS;
  if (HID * ZS > wt) wt = HID * ZS;
  size_t fl = 3 * (size_t)n_cap * HID + TM * ZS + wt + ((R * NB + 3) & ~3) + ((n_cap + 3) & ~3) +
              ((F + 3) & ~3) + L1O + (size_t)MP_WARPS * R * NB;
  return fl * sizeof(float);
}

int check_model(const igmc_model_t* M) {
  if (M->num_layers < 1 || M->num_layers > IGMC_MAX_LAYERS) return -10;
  if (M->num_bases != 2 && M->num_bases != 4) return -11;
  if (M->num_relations < 1 || M->num_relations > 256) return -12;
  if (M->in_dim0 < 1 || M->in_dim0 > HID) return -13;
  if (M->readout != 0 && M->readout != 1) return -19;
  return 0;
}

}  // namespace

// relation-space / cluster kernels (csrc/rgcn_rs.cu)
int rs_supported(const igmc_model_t* M);
int rs_plan(const igmc_model_t* M, int n_cap, int cluster, int backward, int* threads, size_t* smem, int* lcap, int* chunk);
int rs_forward(const igmc_model_t* M, const float* params, const uint8_t* node_label, const int32_t* node_ptr,
               const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap, const igmc_dropout_t* D, int training,
               const igmc_saved_t* S, const float* y, float loss_scale, float* dpred, float* sqerr, int cluster,
               const igmc_stage_t* stage, int* err, cudaStream_t st);
int rs_backward(const igmc_model_t* M, const float* params, const uint8_t* node_label, const int32_t* node_ptr,
                const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap, const igmc_dropout_t* D,
                const igmc_saved_t* S, const float* dpred, float* gpart, float* dhid, int cluster,
                const igmc_stage_t* stage, int* err, cudaStream_t st);
int rs_stage_plan(const igmc_model_t* M, int n_cap, int cluster, int backward, igmc_stage_t* img);
int rs_stage_lists(const igmc_model_t* M, const int32_t* node_ptr, const int32_t* edge_ptr, const igmc_adj_t* A, int B,
                   int n_cap, const igmc_dropout_t* D, int training, const igmc_stage_t* fwd, const igmc_stage_t* bwd,
                   cudaStream_t st);

int rs_prep_weights(const igmc_model_t* M, const float* params, float* wprep, cudaStream_t st);
int rs_train(const igmc_model_t* M, const float* params, const uint8_t* node_label, const int32_t* node_ptr,
             const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap, const igmc_dropout_t* D,
             const igmc_saved_t* S, const float* y, float loss_scale, float* dpred, float* sqerr, float* gpart,
             float* dhid, int cluster, const igmc_stage_t* stage_f, const igmc_stage_t* stage_b, int* err,
             cudaStream_t st);
int rs_gate_wait(int* gate, int target, int timeout_us, cudaStream_t st);

extern "C" int igmc_gate_wait(int32_t* gate, int target, int timeout_us, void* stream) {
  if (!gate || target < 0 || timeout_us < 0) return -21;
  return rs_gate_wait(gate, target, timeout_us, (cudaStream_t)stream);
}

extern "C" int igmc_prep_weights(const igmc_model_t* M, const float* params, float* wprep, void* stream) {
  int rc = check_model(M);
  if (rc) return rc;
  if (!rs_supported(M)) return -16;
  return rs_prep_weights(M, params, wprep, (cudaStream_t)stream);
}

extern "C" int igmc_model_plan(const igmc_model_t* M, int n_cap, int cluster, int backward) {
  int rc = check_model(M);
  if (rc) return rc;
  if (cluster == 0) {
    const size_t b = backward ? bwd_smem_bytes(n_cap, M->num_relations, M->num_bases, M->num_layers)
                              : fwd_smem_bytes(n_cap, M->num_relations, M->num_bases, M->num_layers);
    return b > 227 * 1024 ? -3 : (int)b;
  }
  if (cluster < 1 || cluster > 4) return -15;
  if (!rs_supported(M)) return -16;
  int threads, lcap, chunk;
  size_t smem;
  rc = rs_plan(M, n_cap, cluster, backward, &threads, &smem, &lcap, &chunk);
  return rc ? rc : (int)smem;
}

extern "C" int igmc_stage_plan(const igmc_model_t* M, int n_cap, int cluster, int backward, igmc_stage_t* img) {
  int rc = check_model(M);
  if (rc) return rc;
  if (cluster < 1 || cluster > 4) return -15;
  if (!rs_supported(M)) return -16;
  return rs_stage_plan(M, n_cap, cluster, backward, img);
}

extern "C" int igmc_stage_lists(const igmc_model_t* M, const int32_t* node_ptr, const int32_t* edge_ptr,
                                const igmc_adj_t* A, int B, int n_cap, const igmc_dropout_t* D, int training,
                                const igmc_stage_t* fwd, const igmc_stage_t* bwd, int* err, void* stream) {
  (void)err;
  if (B <= 0) return 0;
  int rc = check_model(M);
  if (rc) return rc;
  if (!rs_supported(M)) return -16;
  return rs_stage_lists(M, node_ptr, edge_ptr, A, B, n_cap, D, training, fwd, bwd, (cudaStream_t)stream);
}

extern "C" int igmc_raw_grad_count(const igmc_model_t* M) {
  int rc = check_model(M);
  if (rc) return rc;
  return igmc_raw_count(M->num_relations, M->in_dim0, M->num_layers);
}

extern "C" int igmc_forward(const igmc_model_t* M, const float* params, const uint8_t* node_label,
                            const int32_t* node_ptr, const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap,
                            const igmc_dropout_t* D, int training, const igmc_saved_t* S, const float* y,
                            float loss_scale, float* dpred, float* sqerr, int cluster, const igmc_stage_t* stage,
                            int* err, void* stream) {
  if (B <= 0) return 0;
  int rc = check_model(M);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (cluster > 0) {
    if (!rs_supported(M)) return -16;
    if (!S->wprep) return -17;
    return rs_forward(M, params, node_label, node_ptr, edge_ptr, A, B, n_cap, D, training, S, y, loss_scale, dpred,
                      sqerr, cluster, stage, err, st);
  }
  const size_t smem = fwd_smem_bytes(n_cap, M->num_relations, M->num_bases, M->num_layers);
  if (smem > 227 * 1024) return -3;
  if (M->num_bases == 4) {
    cudaFuncSetAttribute(k_forward<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_forward<4><<<B, MP_THREADS, smem, st>>>(*M, params, node_label, node_ptr, edge_ptr, *A, n_cap, *D, training, *S,
                                              y, loss_scale, dpred, sqerr, err);
  } else {
    cudaFuncSetAttribute(k_forward<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_forward<2><<<B, MP_THREADS, smem, st>>>(*M, params, node_label, node_ptr, edge_ptr, *A, n_cap, *D, training, *S,
                                              y, loss_scale, dpred, sqerr, err);
  }
  IGMC_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int igmc_forward_backward(const igmc_model_t* M, const float* params, const uint8_t* node_label,
                                     const int32_t* node_ptr, const int32_t* edge_ptr, const igmc_adj_t* A, int B,
                                     int n_cap, const igmc_dropout_t* D, const igmc_saved_t* S, const float* y,
                                     float loss_scale, float* dpred, float* sqerr, float* gpart, float* dhid,
                                     int cluster, const igmc_stage_t* stage_fwd, const igmc_stage_t* stage_bwd,
                                     int* err, void* stream) {
  if (B <= 0) return 0;
  int rc = check_model(M);
  if (rc) return rc;
  if (cluster <= 0 || M->readout != 0 || !rs_supported(M)) return -16;   // cluster plans, IGMC readout
  if (!S->wprep || !S->zsave || !y || !dpred) return -17;
  return rs_train(M, params, node_label, node_ptr, edge_ptr, A, B, n_cap, D, S, y, loss_scale, dpred, sqerr, gpart, dhid,
                  cluster, stage_fwd, stage_bwd, err, (cudaStream_t)stream);
}

extern "C" int igmc_backward(const igmc_model_t* M, const float* params, const uint8_t* node_label,
                             const int32_t* node_ptr, const int32_t* edge_ptr, const igmc_adj_t* A, int B, int n_cap,
                             const igmc_dropout_t* D, const igmc_saved_t* S, const float* dpred, float* gpart,
                             float* dhid, int cluster, const igmc_stage_t* stage, int* err, void* stream) {
  if (B <= 0) return 0;
  int rc = check_model(M);
  if (rc) return rc;
  if (!S->zsave) return -14;
  if (M->readout != 0 && !S->dstate) return -17;
  cudaStream_t st = (cudaStream_t)stream;
  if (cluster > 0) {
    if (!rs_supported(M)) return -16;
    if (!S->wprep) return -17;
    return rs_backward(M, params, node_label, node_ptr, edge_ptr, A, B, n_cap, D, S, dpred, gpart, dhid, cluster, stage,
                       err, st);
  }
  const size_t smem = bwd_smem_bytes(n_cap, M->num_relations, M->num_bases, M->num_layers);
  if (smem > 227 * 1024) return -3;
  if (M->num_bases == 4) {
    cudaFuncSetAttribute(k_backward<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_backward<4><<<B, MP_THREADS, smem, st>>>(*M, params, node_label, node_ptr, edge_ptr, *A, n_cap, *D, *S, dpred,
                                               gpart, dhid, err);
  } else {
    cudaFuncSetAttribute(k_backward<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_backward<2><<<B, MP_THREADS, smem, st>>>(*M, params, node_label, node_ptr, edge_ptr, *A, n_cap, *D, *S, dpred,
                                               gpart, dhid, err);
  }
  IGMC_CUDA_CHECK_LAUNCH();
  return 0;
}
