// Batched enclosing-subgraph extraction (h = 1..3) over the device-resident rating CSR/CSC.
//
// Replaces, for a whole mini-batch in two launches, the reference's per-pair Python path
//   MyDynamicDataset.get            util_functions.py:138-145
//   subgraph_extraction_labeling    util_functions.py:208-247   (BFS, sampling, labels, induced sub-matrix)
//   construct_pyg_graph / one_hot   util_functions.py:280-297, 307-311
//   PyG Batch.from_data_list        (third party)
// and emits the already-collated batch in canonical form: per side the target
// first, then the fringe in ascending global id; undirected edges sorted by (u_local, v_local).
//
// Layout: one CTA per (user,item) pair.
//   k_extract_select_count : fringe discovery from CSC column j / CSR row i (coalesced int32
//       loads), uniform k-subset sampling by 4-pass radix-select over counter-hash keys
//       (no storage, any degree), item-id -> local-id table in shared memory, per-user-row
//       match count (warp per row).
//   k_extract_fill         : cross-graph prefix (each CTA sums its predecessors' counts),
//       second row scan with warp-ballot ordered compaction, writes x / labels / batch / y /
//       edge_index (int64, PyG order [u|v ; v|u]) / edge_type / node_ptr / edge_ptr.
// HBM-bound integer work: no tensor cores; int32 index + uint8 rating per nonzero.
#include "common.cuh"
#include "../../include/igmc_b200.h"

namespace {

constexpr int EX_THREADS = 1024;
constexpr uint16_t NONE16 = 0xFFFF;

// Select the node list of one side and hop from the sorted candidate list `nbr`: with `emit_target` (hop 1)
// out[0] = target and out[1..] = the (sampled) fringe ascending, the target itself excluded from the candidates;
// without (hops >= 2: `out` points behind the nodes found so far) out[0..] = the (sampled) fringe ascending.
__device__ void select_side(const int32_t* __restrict__ nbr, int len, int target, bool emit_target, int mnph,
                            double ratio, uint64_t state, int32_t* __restrict__ outp, int cap, int* n_out,
                            int* hist, int* ws, int* sh, int* err) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int base_ = emit_target ? 1 : 0;
  int32_t* out = outp + base_ - 1;          // the code below writes the fringe at out[1 + ...]
  if (tid == 0) sh[0] = -1;
  __syncthreads();
  if (emit_target)
    for (int p = tid; p < len; p += nt)
      if (nbr[p] == target) sh[0] = p;  // sorted, duplicate-free list: at most one hit
  __syncthreads();
  const int pos = sh[0];
  const int d = len - (pos >= 0 ? 1 : 0);
  int k = d;
  if (ratio < 1.0) k = (int)(ratio * (double)d);   // int(sample_ratio*len(fringe)), ref :223-224
  if (mnph >= 0 && mnph < k) k = mnph;             // strict '<', ref :226,:228
  if (k + base_ > cap) {
    if (tid == 0) { igmc_set_err(err, IGMC_ERR_NODE_CAP); if (emit_target) outp[0] = target; *n_out = base_; }
    __syncthreads();
    return;
  }
  if (tid == 0) { if (emit_target) outp[0] = target; *n_out = base_ + k; }
  if (k == d) {  // no sampling: ordered copy minus the target
    for (int p = tid; p < len; p += nt) {
      if (p == pos) continue;
      out[1 + p - ((pos >= 0 && p > pos) ? 1 : 0)] = nbr[p];
    }
    __syncthreads();
    return;
  }
  if (k == 0) { __syncthreads(); return; }

  // ---- radix-select the k-th smallest 32-bit key (ties resolved by ascending node id) ----
  uint32_t prefix = 0, mask = 0;
  int remaining = k;
  for (int pass = 3; pass >= 0; --pass) {
    for (int b = tid; b < 256; b += nt) hist[b] = 0;
    __syncthreads();
    for (int p = tid; p < len; p += nt) {
      if (p == pos) continue;
      uint32_t key = sample_key(state, nbr[p]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1);
    }
    __syncthreads();
    if (tid < 32) {
      int s = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += hist[8 * tid + i];
      int incl = warp_incl_scan_i(s, tid);
      int excl = incl - s;
      if (excl < remaining && remaining <= incl) {
        int c = excl, b = 8 * tid + 7;
        for (int i = 0; i < 8; ++i) {
          int hc = hist[8 * tid + i];
          if (c + hc >= remaining) { b = 8 * tid + i; break; }
          c += hc;
        }
        sh[1] = b;
        sh[2] = remaining - c;
      }
    }
    __syncthreads();
    prefix |= (uint32_t)sh[1] << (8 * pass);
    mask |= 255u << (8 * pass);
    remaining = sh[2];
    __syncthreads();
  }
  const uint32_t tau = prefix;  // keys < tau are all taken; `remaining` ties (key == tau) are taken
  int run_strict = 0, run_tie = 0;
  for (int base = 0; base < len; base += nt) {
    const int p = base + tid;
    int strict = 0, tie = 0, node = 0;
    if (p < len && p != pos) {
      node = nbr[p];
      uint32_t key = sample_key(state, node);
      strict = key < tau;
      tie = key == tau;
    }
    int tot;
    int ex = block_excl_scan_i(strict | (tie << 16), ws, &tot);
    const int sb = run_strict + (ex & 0xFFFF), tb = run_tie + (ex >> 16);
    if (strict || (tie && tb < remaining)) out[1 + sb + min(tb, remaining)] = node;
    run_strict += tot & 0xFFFF;
    run_tie += tot >> 16;
  }
  __syncthreads();
}

// ---- hops >= 2 (util_functions.py:216-235): bitmaps over one side's ids in shared memory -------------------------
// cand |= union of the CSR rows (CSC columns) of `nodes`  (warp per node, coalesced index reads)
__device__ void mark_neighbours(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                const int32_t* __restrict__ nodes, int cnt, uint32_t* cand) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int a = warp; a < cnt; a += nwarps) {
    const int v = nodes[a];
    for (int p = ptr[v] + lane; p < ptr[v + 1]; p += 32) {
      const int t = idx[p];
      atomicOr(&cand[t >> 5], 1u << (t & 31));
    }
  }
}
// fringe = cand & ~visited; visited |= fringe; list = ids of the fringe ascending.  Returns its length (block-wide).
__device__ int fringe_to_list(uint32_t* cand, uint32_t* vis, int words, int32_t* list, int* ws) {
  const int tid = threadIdx.x, nt = blockDim.x;
  int running = 0;
  for (int base = 0; base < words; base += nt) {
    const int w = base + tid;
    uint32_t f = 0;
    if (w < words) {
      f = cand[w] & ~vis[w];
      vis[w] |= f;
    }
    int tot;
    const int ex = block_excl_scan_i(__popc(f), ws, &tot);
    int o = running + ex;
    while (f) {
      const int b = __ffs(f) - 1;
      list[o++] = (w << 5) + b;
      f &= f - 1;
    }
    running += tot;
  }
  __syncthreads();
  return running;
}

__device__ __forceinline__ void resolve_pair(const igmc_pairs_t& P, int g, int* i, int* j, int* lab, int64_t* pid) {
  const int64_t src = P.idx ? P.idx[g] : (int64_t)g;
  *i = P.links_u[src];
  *j = P.links_v[src];
  *lab = P.links_label ? P.links_label[src] : 0;
  *pid = P.pair_id ? P.pair_id[g] : src;
}

__global__ void __launch_bounds__(EX_THREADS)
k_extract_select_count(igmc_csr_t G, igmc_pairs_t P, int B, int mnph, double ratio, uint64_t seed_val,
                       const uint64_t* __restrict__ seed_dev, int cap,
                       const int32_t* __restrict__ inj_nodes_u, const int32_t* __restrict__ inj_nodes_v,
                       const int32_t* __restrict__ inj_n_u, const int32_t* __restrict__ inj_n_v,
                       int32_t* __restrict__ nodes_u, int32_t* __restrict__ nodes_v,
                       int32_t* __restrict__ n_u, int32_t* __restrict__ n_v,
                       int32_t* __restrict__ row_cnt, int32_t* __restrict__ m_cnt, int32_t* __restrict__ col_cnt,
                       int h, int32_t* __restrict__ hop_off, int* err) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* colcnt = reinterpret_cast<int*>(smem_raw);                 // [cap]
  uint16_t* tab = reinterpret_cast<uint16_t*>(colcnt + cap);      // [num_items]
  // h > 1 only: visited / candidate bitmaps and the ascending fringe list, behind the item table
  const int wu = (G.num_users + 31) >> 5, wv = (G.num_items + 31) >> 5;
  uint32_t* vis_u = reinterpret_cast<uint32_t*>(smem_raw + (((size_t)cap * 4 + (size_t)G.num_items * 2 + 15) & ~(size_t)15));
  uint32_t* vis_v = vis_u + wu;
  uint32_t* cand = vis_v + wv;
  int32_t* flist = reinterpret_cast<int32_t*>(cand + max(wu, wv));
  __shared__ int hist[256];
  __shared__ int ws[34];
  __shared__ int sh[4];
  __shared__ int s_nu, s_nv, s_m;
  const int g = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const uint64_t seed = seed_dev ? *seed_dev : seed_val;
  int i, j, lab;
  int64_t pid;
  resolve_pair(P, g, &i, &j, &lab, &pid);
  int32_t* gu = nodes_u + (size_t)g * cap;
  int32_t* gv = nodes_v + (size_t)g * cap;
  if (tid == 0) s_m = 0;
  if (inj_nodes_u) {  // test hook: node lists supplied (the reference's own draw)
    const int nu = inj_n_u[g], nv = inj_n_v[g];
    if (nu > cap || nv > cap) {
      if (tid == 0) { igmc_set_err(err, IGMC_ERR_NODE_CAP); s_nu = 1; s_nv = 1; gu[0] = i; gv[0] = j; }
    } else {
      for (int t = tid; t < nu; t += nt) gu[t] = inj_nodes_u[(size_t)g * cap + t];
      for (int t = tid; t < nv; t += nt) gv[t] = inj_nodes_v[(size_t)g * cap + t];
      if (tid == 0) { s_nu = nu; s_nv = nv; }
    }
    __syncthreads();
  } else {
    // user fringe = users who rated item j (CSC column j); item fringe = items rated by user i (CSR row i)
    select_side(G.row_idx + G.col_ptr[j], G.col_ptr[j + 1] - G.col_ptr[j], i, true, mnph, ratio,
                sample_state(seed, pid, 0, 1), gu, cap, &s_nu, hist, ws, sh, err);
    select_side(G.col_idx + G.row_ptr[i], G.row_ptr[i + 1] - G.row_ptr[i], j, true, mnph, ratio,
                sample_state(seed, pid, 1, 1), gv, cap, &s_nv, hist, ws, sh, err);
  }
  __syncthreads();
  if (hop_off && tid == 0) {   // nodes within distance d per side: [users 0..3 | items 0..3]
    int32_t* ho = hop_off + (size_t)g * 2 * (IGMC_MAX_HOP + 1);
    ho[0] = 1; ho[IGMC_MAX_HOP + 1] = 1;
    for (int d = 1; d <= IGMC_MAX_HOP; ++d) { ho[d] = s_nu; ho[IGMC_MAX_HOP + 1 + d] = s_nv; }
  }
  if (h > 1 && !inj_nodes_u) {
    // visited = the targets and the WHOLE hop-1 fringes (the reference marks a fringe visited before it samples it,
    // util_functions.py:220-221)
    for (int w = tid; w < wu + wv; w += nt) vis_u[w] = 0;   // vis_u and vis_v are contiguous
    __syncthreads();
    if (tid == 0) { atomicOr(&vis_u[i >> 5], 1u << (i & 31)); atomicOr(&vis_v[j >> 5], 1u << (j & 31)); }
    for (int p = G.col_ptr[j] + tid; p < G.col_ptr[j + 1]; p += nt) { const int t = G.row_idx[p]; atomicOr(&vis_u[t >> 5], 1u << (t & 31)); }
    for (int p = G.row_ptr[i] + tid; p < G.row_ptr[i + 1]; p += nt) { const int t = G.col_idx[p]; atomicOr(&vis_v[t >> 5], 1u << (t & 31)); }
    __syncthreads();
    int lo_u = 1, hi_u = s_nu, lo_v = 1, hi_v = s_nv;
    for (int d = 2; d <= h; ++d) {
      // both new fringes come from the PREVIOUS fringes (tuple assignment, util_functions.py:217)
      for (int w = tid; w < wv; w += nt) cand[w] = 0;
      __syncthreads();
      mark_neighbours(G.row_ptr, G.col_idx, gu + lo_u, hi_u - lo_u, cand);
      __syncthreads();
      int len = fringe_to_list(cand, vis_v, wv, flist, ws);
      select_side(flist, len, -1, false, mnph, ratio, sample_state(seed, pid, 1, d), gv + hi_v, cap - hi_v, &s_nv,
                  hist, ws, sh, err);
      __syncthreads();
      const int add_v = s_nv;
      for (int w = tid; w < wu; w += nt) cand[w] = 0;
      __syncthreads();
      mark_neighbours(G.col_ptr, G.row_idx, gv + lo_v, hi_v - lo_v, cand);
      __syncthreads();
      len = fringe_to_list(cand, vis_u, wu, flist, ws);
      select_side(flist, len, -1, false, mnph, ratio, sample_state(seed, pid, 0, d), gu + hi_u, cap - hi_u, &s_nu,
                  hist, ws, sh, err);
      __syncthreads();
      const int add_u = s_nu;
      if (add_u == 0 && add_v == 0) break;        // util_functions.py:230-231
      lo_u = hi_u; hi_u += add_u; lo_v = hi_v; hi_v += add_v;
      if (hop_off && tid == 0) {
        int32_t* ho = hop_off + (size_t)g * 2 * (IGMC_MAX_HOP + 1);
        for (int dd = d; dd <= IGMC_MAX_HOP; ++dd) { ho[dd] = hi_u; ho[IGMC_MAX_HOP + 1 + dd] = hi_v; }
      }
      __syncthreads();
    }
    if (tid == 0) { s_nu = hi_u; s_nv = hi_v; }
    __syncthreads();
  }
  const int nu = s_nu, nv = s_nv;
  // item-id -> local-id table
  for (int t = tid; t < G.num_items; t += nt) tab[t] = NONE16;
  for (int b = tid; b < nv; b += nt) colcnt[b] = 0;
  __syncthreads();
  for (int b = tid; b < nv; b += nt) tab[gv[b]] = (uint16_t)b;
  __syncthreads();
  // per-row match count (warp per user row); bit31 = row contains the target item j
  for (int a = warp; a < nu; a += nwarps) {
    const int u = gu[a];
    const int s = G.row_ptr[u], e = G.row_ptr[u + 1];
    int cnt = 0, hasj = 0;
    for (int p = s + lane; p < e; p += 32) {
      const uint16_t b = tab[G.col_idx[p]];
      if (b != NONE16) {
        if (b == 0) { hasj = 1; if (a != 0) { ++cnt; atomicAdd(&colcnt[0], 1); } }   // (0,0) = target edge: dropped (ref :238)
        else { ++cnt; atomicAdd(&colcnt[b], 1); }
      }
    }
    cnt = warp_sum_i(cnt);
    hasj = __any_sync(IGMC_FULL, hasj);
    if (lane == 0) {
      row_cnt[(size_t)g * cap + a] = cnt | (hasj ? (int)0x80000000 : 0);
      atomicAdd(&s_m, cnt);
    }
  }
  __syncthreads();
  for (int b = tid; b < nv; b += nt) col_cnt[(size_t)g * cap + b] = colcnt[b];
  if (tid == 0) { n_u[g] = nu; n_v[g] = nv; m_cnt[g] = s_m; }
}

__global__ void __launch_bounds__(EX_THREADS)
k_extract_fill(igmc_csr_t G, igmc_pairs_t P, int B, int cap,
               const int32_t* __restrict__ nodes_u, const int32_t* __restrict__ nodes_v,
               const int32_t* __restrict__ n_u, const int32_t* __restrict__ n_v,
               const int32_t* __restrict__ row_cnt, const int32_t* __restrict__ m_cnt,
               const int32_t* __restrict__ col_cnt, int h, const int32_t* __restrict__ hop_off,
               const float* __restrict__ class_values, igmc_batch_out_t O, int* err) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* rowoff = reinterpret_cast<int*>(smem_raw);                 // [cap]
  int* colfill = rowoff + cap;                                    // [cap] running fill pointer of the item lists
  uint16_t* tab = reinterpret_cast<uint16_t*>(colfill + cap);     // [num_items]
  __shared__ int ws[34];
  const int g = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  int nsum = 0, msum = 0;
  for (int q = tid; q < g; q += nt) { nsum += n_u[q] + n_v[q]; msum += m_cnt[q]; }
  const int Nbase = block_sum_i(nsum, ws);
  const int Mbase = block_sum_i(msum, ws);
  const int nu = n_u[g], nv = n_v[g], m = m_cnt[g], n = nu + nv;
  const int32_t* gu = nodes_u + (size_t)g * cap;
  const int32_t* gv = nodes_v + (size_t)g * cap;
  const bool overflow = (Nbase + n > O.node_cap) || (2 * (Mbase + m) > O.edge_cap);
  const bool want_adj = O.adj_in_ptr != nullptr;
  if (g == B - 1 && tid == 0) {
    O.counts[0] = Nbase + n;
    O.counts[1] = 2 * (Mbase + m);
    O.node_ptr[B] = Nbase + n;
    O.edge_ptr[B] = 2 * (Mbase + m);
    if (want_adj && Nbase + n <= O.node_cap) O.adj_in_ptr[Nbase + n] = 2 * (Mbase + m);
  }
  if (tid == 0) { O.node_ptr[g] = Nbase; O.edge_ptr[g] = 2 * Mbase; }
  if (overflow) {
    if (tid == 0) igmc_set_err(err, (Nbase + n > O.node_cap) ? IGMC_ERR_NODE_TOTAL : IGMC_ERR_EDGE_CAP);
    return;
  }
  int i, j, lab;
  int64_t pid;
  resolve_pair(P, g, &i, &j, &lab, &pid);
  if (tid == 0) {
    O.y[g] = class_values[lab];
    O.graph_nu[g] = nu;
  }
  // node labels: user at distance d -> 2d, item -> 2d+1 (ref :245); h=1: 0 / 1 targets, 2 / 3 the others
  int hu[IGMC_MAX_HOP + 1], hv[IGMC_MAX_HOP + 1];   // nodes within distance d, per side
#pragma unroll
  for (int d = 0; d <= IGMC_MAX_HOP; ++d) {
    hu[d] = (h > 1 && hop_off) ? hop_off[(size_t)g * 2 * (IGMC_MAX_HOP + 1) + d] : (d == 0 ? 1 : nu);
    hv[d] = (h > 1 && hop_off) ? hop_off[(size_t)g * 2 * (IGMC_MAX_HOP + 1) + IGMC_MAX_HOP + 1 + d] : (d == 0 ? 1 : nv);
  }
  for (int t = tid; t < n; t += nt) {
    int label;
    if (t < nu) {
      int d = 0;
#pragma unroll
      for (int q = 0; q < IGMC_MAX_HOP; ++q) d += (t >= hu[q]) ? 1 : 0;
      label = 2 * d;
    } else {
      int d = 0;
#pragma unroll
      for (int q = 0; q < IGMC_MAX_HOP; ++q) d += (t - nu >= hv[q]) ? 1 : 0;
      label = 2 * d + 1;
    }
    const size_t row = (size_t)Nbase + t;
    O.node_label[row] = (uint8_t)label;
    O.batch[row] = g;
    O.node_gid[row] = t < nu ? gu[t] : gv[t - nu];
    if (O.x) {
      for (int f = 0; f < O.feat_dim; ++f) O.x[row * O.feat_dim + f] = (f == label) ? 1.0f : 0.0f;
    }
  }
  // exclusive scan of the per-row match counts
  int running = 0;
  for (int base = 0; base < nu; base += nt) {
    const int a = base + tid;
    const int c = a < nu ? (row_cnt[(size_t)g * cap + a] & 0x7fffffff) : 0;
    int tot;
    const int ex = block_excl_scan_i(c, ws, &tot);
    if (a < nu) {
      rowoff[a] = running + ex;
      if (want_adj) O.adj_in_ptr[Nbase + a] = 2 * Mbase + running + ex;       // user a: in-edges from its items
    }
    running += tot;
  }
  if (want_adj) {   // item b: in-edges from users, stored after the m user-side entries of this graph
    running = 0;
    for (int base = 0; base < nv; base += nt) {
      const int b = base + tid;
      const int c = b < nv ? col_cnt[(size_t)g * cap + b] : 0;
      int tot;
      const int ex = block_excl_scan_i(c, ws, &tot);
      if (b < nv) {
        colfill[b] = running + ex;
        O.adj_in_ptr[Nbase + nu + b] = 2 * Mbase + m + running + ex;
      }
      running += tot;
    }
  }
  for (int t = tid; t < G.num_items; t += nt) tab[t] = NONE16;
  __syncthreads();
  for (int b = tid; b < nv; b += nt) tab[gv[b]] = (uint16_t)b;
  __syncthreads();
  int64_t* ei0 = O.edge_index;
  int64_t* ei1 = O.edge_index + O.edge_cap;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int a = warp; a < nu; a += nwarps) {
    const int u = gu[a];
    const int s = G.row_ptr[u], e = G.row_ptr[u + 1];
    const int jfirst = (a != 0 && (row_cnt[(size_t)g * cap + a] < 0)) ? 1 : 0;  // j sorts first (v_local 0)
    const int base = rowoff[a];
    int seen = 0;
    // h > 1: local item ids ascend with (hop, global id), the row is in global-id order -> a match's rank is the
    // number of matches of earlier hops plus its position among the matches of its own hop
    int hop_base[IGMC_MAX_HOP + 1] = {0, 0, 0, 0}, hop_seen[IGMC_MAX_HOP + 1] = {0, 0, 0, 0};
    if (h > 1) {
      int c[IGMC_MAX_HOP + 1] = {0, 0, 0, 0};
      for (int p0 = s; p0 < e; p0 += 32) {
        const int p = p0 + lane;
        uint16_t b = NONE16;
        if (p < e) b = tab[G.col_idx[p]];
        const bool m_ = (b != NONE16) && b != 0;
#pragma unroll
        for (int d = 1; d <= IGMC_MAX_HOP; ++d)
          c[d] += __popc(__ballot_sync(IGMC_FULL, m_ && (int)b >= hv[d - 1] && (int)b < hv[d]));
      }
      hop_base[1] = 0;
#pragma unroll
      for (int d = 2; d <= IGMC_MAX_HOP; ++d) hop_base[d] = hop_base[d - 1] + c[d - 1];
    }
    for (int p0 = s; p0 < e; p0 += 32) {
      const int p = p0 + lane;
      uint16_t b = NONE16;
      if (p < e) b = tab[G.col_idx[p]];
      const bool match = (b != NONE16) && !(a == 0 && b == 0);
      const bool isj = match && b == 0;
      const unsigned bal = __ballot_sync(IGMC_FULL, match && !isj);
      int hrank = 0;
      if (h > 1) {
#pragma unroll
        for (int d = 1; d <= IGMC_MAX_HOP; ++d) {
          const bool in_d = match && !isj && (int)b >= hv[d - 1] && (int)b < hv[d];
          const unsigned bd = __ballot_sync(IGMC_FULL, in_d);
          if (in_d) hrank = hop_base[d] + hop_seen[d] + __popc(bd & lt_mask);
          hop_seen[d] += __popc(bd);
        }
      }
      if (match) {
        const int rank = isj ? 0 : (h > 1 ? jfirst + hrank : jfirst + seen + __popc(bal & lt_mask));
        const int64_t r = G.rating[p];
        const size_t e1 = (size_t)2 * Mbase + base + rank, e2 = e1 + m;
        const int64_t un = Nbase + a, vn = Nbase + nu + b;
        ei0[e1] = un; ei1[e1] = vn; O.edge_type[e1] = r;
        ei0[e2] = vn; ei1[e2] = un; O.edge_type[e2] = r;
        if (want_adj) {
          // key = type | neighbour | edge id : lists are rank-sorted by (type, neighbour) below.
          // user a's list (slots [rowoff[a], +cnt)): in-edge item->user is the mirrored copy e2
          O.adj_tmp[e1] = ((uint64_t)r << 56) | ((uint64_t)(nu + b) << 32) | (uint64_t)(uint32_t)e2;
          // item b's list: in-edge user->item is e1; unordered placement
          const int slot = atomicAdd(&colfill[b], 1);
          O.adj_tmp[(size_t)2 * Mbase + m + slot] = ((uint64_t)r << 56) | ((uint64_t)a << 32) | (uint64_t)(uint32_t)e1;
        }
      }
      seen += __popc(bal);
    }
  }
  if (want_adj) {   // deterministic lists sorted by (type, neighbour): rank-sort (lists are short)
    __syncthreads();
    for (int v = warp; v < n; v += nwarps) {
      int k;
      size_t beg;
      if (v < nu) {
        k = row_cnt[(size_t)g * cap + v] & 0x7fffffff;
        beg = (size_t)2 * Mbase + rowoff[v];
      } else {
        k = col_cnt[(size_t)g * cap + (v - nu)];
        beg = (size_t)2 * Mbase + m + (colfill[v - nu] - k);
      }
      for (int i = lane; i < k; i += 32) {
        const uint64_t key = O.adj_tmp[beg + i];
        int rank = 0;
        for (int q = 0; q < k; ++q) rank += (O.adj_tmp[beg + q] < key) ? 1 : 0;
        O.adj_in[beg + rank] = (uint32_t)((key >> 32) & 0xffffu) | ((uint32_t)(key >> 56) << 16);
        O.adj_eid[beg + rank] = (int32_t)(uint32_t)(key & 0xffffffffu);
      }
    }
  }
}


// Batch assembly from the static store (one CTA per output graph).
__global__ void __launch_bounds__(256)
k_assemble(igmc_store_t S, const int64_t* __restrict__ idx, int B, igmc_batch_out_t O, int* err) {
  __shared__ int ws[34];
  const int g = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  int nsum = 0, esum = 0;
  for (int q = tid; q < g; q += nt) {
    const int64_t s = idx[q];
    nsum += S.node_off[s + 1] - S.node_off[s];
    esum += S.edge_off[s + 1] - S.edge_off[s];
  }
  const int Nbase = block_sum_i(nsum, ws);
  const int Ebase = block_sum_i(esum, ws);
  const int64_t s = idx[g];
  const int n0 = S.node_off[s], n = S.node_off[s + 1] - n0;
  const int e0 = S.edge_off[s], e = S.edge_off[s + 1] - e0;
  if (g == B - 1 && tid == 0) {
    O.counts[0] = Nbase + n;
    O.counts[1] = Ebase + e;
    O.node_ptr[B] = Nbase + n;
    O.edge_ptr[B] = Ebase + e;
    if (O.adj_in_ptr && Nbase + n <= O.node_cap) O.adj_in_ptr[Nbase + n] = Ebase + e;
  }
  if (tid == 0) { O.node_ptr[g] = Nbase; O.edge_ptr[g] = Ebase; }
  if (Nbase + n > O.node_cap || Ebase + e > O.edge_cap) {
    if (tid == 0) igmc_set_err(err, (Nbase + n > O.node_cap) ? IGMC_ERR_NODE_TOTAL : IGMC_ERR_EDGE_CAP);
    return;
  }
  if (tid == 0) { O.y[g] = S.y[s]; O.graph_nu[g] = S.graph_nu[s]; }
  const int32_t* aptr = S.adj_ptr + n0 + s;   // per graph n+1 entries -> offset n0 + s
  for (int t = tid; t < n; t += nt) {
    const size_t row = (size_t)Nbase + t;
    const int label = S.node_label[n0 + t];
    O.node_label[row] = (uint8_t)label;
    O.batch[row] = g;
    O.node_gid[row] = S.node_gid[n0 + t];
    if (O.x)
      for (int f = 0; f < O.feat_dim; ++f) O.x[row * O.feat_dim + f] = (f == label) ? 1.0f : 0.0f;
    if (O.adj_in_ptr) O.adj_in_ptr[row] = Ebase + aptr[t];
  }
  int64_t* ei0 = O.edge_index;
  int64_t* ei1 = O.edge_index + O.edge_cap;
  for (int t = tid; t < e; t += nt) {
    const size_t k = (size_t)Ebase + t;
    ei0[k] = (int64_t)Nbase + S.edge_src[e0 + t];
    ei1[k] = (int64_t)Nbase + S.edge_dst[e0 + t];
    O.edge_type[k] = S.edge_type[e0 + t];
    if (O.adj_in_ptr) {
      O.adj_in[k] = S.adj_in[e0 + t];
      O.adj_eid[k] = Ebase + S.adj_eid[e0 + t];
    }
  }
}

}  // namespace

extern "C" int igmc_extract_batch(const igmc_csr_t* G, const igmc_pairs_t* P, int B, int h, int max_nodes_per_hop,
                                  double sample_ratio, uint64_t seed, const uint64_t* seed_dev, int cap,
                                  const int32_t* inj_nodes_u, const int32_t* inj_nodes_v,
                                  const int32_t* inj_n_u, const int32_t* inj_n_v,
                                  const igmc_extract_ws_t* W, const float* class_values,
                                  const igmc_batch_out_t* O, int* err, void* stream) {
  if (B <= 0) return 0;
  if (cap < 1 || cap > 65534) return -2;
  if (h < 1 || h > IGMC_MAX_HOP) return -4;
  if (h > 1 && (inj_nodes_u || !W->hop_off)) return -4;   // injected node lists carry no hop boundaries
  cudaStream_t st = (cudaStream_t)stream;
  size_t smemA = (size_t)cap * sizeof(int) + (size_t)G->num_items * sizeof(uint16_t);
  if (h > 1) {   // visited / candidate bitmaps + the fringe list of the larger side
    const size_t wu = (G->num_users + 31) / 32, wv = (G->num_items + 31) / 32;
    smemA = ((smemA + 15) & ~(size_t)15) + (wu + wv + (wu > wv ? wu : wv)) * 4 +
            (size_t)(G->num_users > G->num_items ? G->num_users : G->num_items) * 4;
    if (smemA > 220 * 1024) return -3;
  }
  const size_t smemB = 2 * (size_t)cap * sizeof(int) + (size_t)G->num_items * sizeof(uint16_t);
  if (smemB > 220 * 1024) return -3;  // item table does not fit in shared memory
  cudaFuncSetAttribute(k_extract_select_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemA);
  cudaFuncSetAttribute(k_extract_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemB);
  k_extract_select_count<<<B, EX_THREADS, smemA, st>>>(*G, *P, B, max_nodes_per_hop, sample_ratio, seed, seed_dev, cap,
                                                       inj_nodes_u, inj_nodes_v, inj_n_u, inj_n_v,
                                                       W->nodes_u, W->nodes_v, W->n_u, W->n_v, W->row_cnt, W->m_cnt,
                                                       W->col_cnt, h, W->hop_off, err);
  IGMC_CUDA_CHECK_LAUNCH();
  k_extract_fill<<<B, EX_THREADS, smemB, st>>>(*G, *P, B, cap, W->nodes_u, W->nodes_v, W->n_u, W->n_v,
                                               W->row_cnt, W->m_cnt, W->col_cnt, h, W->hop_off, class_values, *O, err);
  IGMC_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int igmc_assemble_batch(const igmc_store_t* S, const int64_t* idx, int B, const igmc_batch_out_t* O, int* err,
                                   void* stream) {
  if (B <= 0) return 0;
  k_assemble<<<B, 256, 0, (cudaStream_t)stream>>>(*S, idx, B, *O, err);
  IGMC_CUDA_CHECK_LAUNCH();
  return 0;
}
