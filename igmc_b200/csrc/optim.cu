// Gradient assembly (per-CTA partial rows -> flat bucket), adjacent-rating regulariser and Adam.
//
// Replaces autograd's accumulation + `loss += ARR * sum((w[1:]-w[:-1])**2)` (train_eval.py:167-174)
// + `torch.optim.Adam.step` (train_eval.py:54,177) with TWO launches on the flat parameter bucket, which is
// also what the one NCCL all-reduce per step operates on.
#include "common.cuh"
#include "../../include/igmc_b200.h"

namespace {

constexpr int HID = IGMC_HIDDEN;
constexpr int L1O = IGMC_LIN1_OUT;

// dW_r[kj] of the adjacent-rating regulariser for one (layer, kj):  reg = sum_{r<R-1} ||W_{r+1}-W_r||^2,
// W_r = sum_b att[r,b] basis[b];  d reg / d W_r = 2 (W_r - W_{r-1}) [r>0] - 2 (W_{r+1} - W_r) [r<R-1].
// `pair` (optional) receives (W_{r+1}[kj]-W_r[kj])^2, the contribution of the pair (r, r+1) to reg.
__device__ __forceinline__ float arr_dw(const float* __restrict__ att, const float* __restrict__ bs, int NB, int KJ,
                                        int R, int r, int kj, float* pair) {
  float wm = 0.f, w0 = 0.f, wp = 0.f;
  for (int b = 0; b < NB; ++b) {
    const float bv = bs[b * KJ + kj];
    w0 = fmaf(att[r * NB + b], bv, w0);
    if (r > 0) wm = fmaf(att[(r - 1) * NB + b], bv, wm);
    if (r < R - 1) wp = fmaf(att[(r + 1) * NB + b], bv, wp);
  }
  float dw = 0.f;
  if (r > 0) dw += 2.f * (w0 - wm);
  if (r < R - 1) dw -= 2.f * (wp - w0);
  if (pair) *pair = r < R - 1 ? (wp - w0) * (wp - w0) : 0.f;
  return dw;
}

__device__ __forceinline__ float block_sum_f256(float v, float* red) {   // 256-thread block, fixed order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum_f(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// Blocks [0, PB): one thread per parameter.
//   conv params = fixed-order sum of the per-CTA partial rows (+ the ARR term for basis entries, thread-local);
//   lin1/lin2 from the saved readout factors; att entries are left to the ARR blocks when arr != 0.
// Blocks [PB, PB + L*R) (arr != 0): block (l, r) owns att[l][r][:]:  sum of the partial rows + arr * <dW_r, basis[b]>,
//   and the regulariser value of the pair (r, r+1).  The last of them to finish writes the loss
//   (sum_g sqerr * loss_scale + arr * sum of the pair values, in fixed order).
__global__ void __launch_bounds__(256)
k_grad_reduce(igmc_model_t M, const float* __restrict__ params, int B, int rows, int PB,
              const float* __restrict__ gpart, const float* __restrict__ dhid, const float* __restrict__ feat,
              const float* __restrict__ hid, const float* __restrict__ dpred, const float* __restrict__ sqerr,
              float loss_scale, float arr, float grad_scale, float* __restrict__ grad, float* __restrict__ loss_out,
              float* __restrict__ reg_ws) {
  const int PC = M.conv_param_count, F = 2 * HID * M.num_layers;
  const int R = M.num_relations, NB = M.num_bases;
  if ((int)blockIdx.x < PB) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < M.param_count) {
      float s = 0.f;
      bool skip = false;
      if (p < PC) {
        for (int g = 0; g < rows; ++g) s += gpart[(size_t)g * PC + p];
        if (arr != 0.f) {
          for (int l = 0; l < M.num_layers; ++l) {
            const int in = l == 0 ? M.in_dim0 : HID, KJ = in * HID;
            const int q = p - M.off_basis[l];
            if (q >= 0 && q < NB * KJ) {   // basis[b][kj]:  += arr * sum_r att[r,b] dW_r[kj]
              const int b = q / KJ, kj = q - b * KJ;
              const float* att = params + M.off_att[l];
              float t = 0.f;
              for (int r = 0; r < R; ++r)
                t = fmaf(att[r * NB + b], arr_dw(att, params + M.off_basis[l], NB, KJ, R, r, kj, nullptr), t);
              s = fmaf(arr, t, s);
            }
            const int qa = p - M.off_att[l];
            if (qa >= 0 && qa < R * NB) skip = true;   // written by the ARR block (l, r)
          }
        }
      } else if (M.readout != 0) {
        skip = true;   // readout parameters belong to the external readout's gradient kernel
      } else if (p >= M.off_lin1_w && p < M.off_lin1_w + L1O * F) {
        const int q = p - M.off_lin1_w, o = q / F, i = q - o * F;      // lin1.weight[o][i]
        for (int g = 0; g < B; ++g) s = fmaf(dhid[(size_t)g * L1O + o], feat[(size_t)g * F + i], s);
      } else if (p >= M.off_lin1_b && p < M.off_lin1_b + L1O) {
        const int o = p - M.off_lin1_b;
        for (int g = 0; g < B; ++g) s += dhid[(size_t)g * L1O + o];
      } else if (p >= M.off_lin2_w && p < M.off_lin2_w + L1O) {
        const int o = p - M.off_lin2_w;
        for (int g = 0; g < B; ++g) s = fmaf(dpred[g], hid[(size_t)g * L1O + o], s);
      } else if (p == M.off_lin2_b) {
        for (int g = 0; g < B; ++g) s += dpred[g];
      }
      if (!skip) grad[p] = s * grad_scale;
    }
    if (arr == 0.f && blockIdx.x == 0 && threadIdx.x == 0 && loss_out) {
      float s = 0.f;
      if (sqerr)
        for (int g = 0; g < B; ++g) s += sqerr[g];
      loss_out[0] = s * loss_scale;
    }
    return;
  }
  // ---- ARR block (l, r) ----
  const int lr = blockIdx.x - PB, l = lr / R, r = lr - l * R;
  const int in = l == 0 ? M.in_dim0 : HID, KJ = in * HID;
  const int tid = threadIdx.x;
  extern __shared__ float sm_arr[];                 // att [R*NB] | basis [NB*KJ] | dW_r [KJ]
  float* att = sm_arr;
  float* bs = att + ((R * NB + 3) & ~3);
  float* dw = bs + NB * KJ;
  __shared__ float red[8];
  __shared__ int s_last;
  for (int i = tid; i < R * NB; i += 256) att[i] = params[M.off_att[l] + i];
  for (int i = tid; i < NB * KJ; i += 256) bs[i] = params[M.off_basis[l] + i];
  __syncthreads();
  float pair_sum = 0.f;
  for (int kj = tid; kj < KJ; kj += 256) {
    float pr;
    dw[kj] = arr_dw(att, bs, NB, KJ, R, r, kj, &pr);
    pair_sum += pr;
  }
  __syncthreads();
  for (int b = 0; b < NB; ++b) {
    float s = 0.f;
    for (int kj = tid; kj < KJ; kj += 256) s = fmaf(dw[kj], bs[b * KJ + kj], s);
    const float dot = block_sum_f256(s, red);
    // column sum of the partial rows for att[l][r][b]
    float c = 0.f;
    for (int g = tid; g < rows; g += 256) c += gpart[(size_t)g * PC + M.off_att[l] + r * NB + b];
    const float col = block_sum_f256(c, red);
    if (tid == 0) grad[M.off_att[l] + r * NB + b] = (col + arr * dot) * grad_scale;
  }
  const float reg = block_sum_f256(pair_sum, red);
  if (tid == 0) {
    reg_ws[lr] = reg;
    __threadfence();
    int* ticket = reinterpret_cast<int*>(reg_ws + IGMC_MAX_LAYERS * 256);
    s_last = (atomicAdd(ticket, 1) == (int)(gridDim.x - PB) - 1);
  }
  __syncthreads();
  if (s_last && tid == 0) {
    __threadfence();
    float s = 0.f;
    for (int q = 0; q < (int)(gridDim.x - PB); ++q) s += __ldcg(reg_ws + q);
    float mse = 0.f;
    if (sqerr)
      for (int g = 0; g < B; ++g) mse += sqerr[g];
    if (loss_out) loss_out[0] = mse * loss_scale + arr * s;
    *reinterpret_cast<int*>(reg_ws + IGMC_MAX_LAYERS * 256) = 0;   // re-arm for the next step
  }
}

// torch.optim.Adam step; the last block to finish increments the device-side step counter.
__global__ void k_adam(float* __restrict__ params, const float* __restrict__ grad, float* __restrict__ m,
                       float* __restrict__ v, int64_t* __restrict__ step_count, int* __restrict__ ticket, int n,
                       float lr_val, const float* __restrict__ lr_dev, float b1, float b2, float eps, float wd,
                       float grad_mul, const float* __restrict__ loss_in, float* __restrict__ loss_acc,
                       float loss_weight) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && loss_acc) loss_acc[0] += loss_in[0] * loss_weight;   // epoch loss bookkeeping (train_eval.py:176)
  const int64_t step_i = step_count[0] + 1;
  if (i < n) {
    const float lr = lr_dev ? *lr_dev : lr_val;
    const double step = (double)step_i;
    const float bc1 = (float)(1.0 - pow((double)b1, step));
    const float bc2 = (float)(1.0 - pow((double)b2, step));
    float g = grad[i] * grad_mul;
    const float p = params[i];
    if (wd != 0.f) g = fmaf(wd, p, g);
    const float mi = b1 * m[i] + (1.f - b1) * g;
    const float vi = b2 * v[i] + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    params[i] = p - (lr / bc1) * (mi / denom);
  }
  __syncthreads();   // every thread of the block has read step_count
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
      step_count[0] = step_i;
      *ticket = 0;
    }
  }
}

}  // namespace

extern "C" int igmc_grad_reduce(const igmc_model_t* M, const float* params, int B, int gpart_rows, const float* gpart,
                                const float* dhid, const float* feat, const float* hid, const float* dpred,
                                const float* sqerr, float loss_scale, float arr, float grad_scale, float* grad,
                                float* loss_out, float* reg_ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int PB = (M->param_count + 255) / 256;
  int blocks = PB;
  size_t smem = 0;
  if (arr != 0.f) {
    if (!reg_ws) return -18;
    blocks += M->num_layers * M->num_relations;
    smem = (size_t)(((M->num_relations * M->num_bases + 3) & ~3) + (M->num_bases + 1) * HID * HID) * sizeof(float);
  }
  k_grad_reduce<<<blocks, 256, smem, st>>>(*M, params, B, gpart_rows, PB, gpart, dhid, feat, hid, dpred, sqerr,
                                          loss_scale, arr, grad_scale, grad, loss_out, reg_ws);
  IGMC_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int igmc_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                              int64_t* step_count, int n, float lr, const float* lr_dev, float beta1, float beta2,
                              float eps, float weight_decay, float grad_mul, const float* loss_in, float* loss_acc,
                              float loss_weight, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  // step_count points at [int64 step | int32 ticket]: the word after the counter is the kernel's completion ticket
  int* ticket = reinterpret_cast<int*>(step_count + 1);
  k_adam<<<(n + 255) / 256, 256, 0, st>>>(params, grad, exp_avg, exp_avg_sq, step_count, ticket, n, lr, lr_dev, beta1,
                                          beta2, eps, weight_decay, grad_mul, loss_in, loss_acc, loss_weight);
  IGMC_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int igmc_build_info(void) {
#ifdef IGMC_SM_ARCH
  return IGMC_SM_ARCH;
#else
  return 0;
#endif
}
