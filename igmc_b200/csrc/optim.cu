// Gradient assembly (per-graph partials -> flat bucket), adjacent-rating regulariser and Adam.
//
// Replaces autograd's accumulation + `loss += ARR * sum((w[1:]-w[:-1])**2)` (train_eval.py:167-174)
// + `torch.optim.Adam.step` (train_eval.py:54,177) with three small launches on the flat parameter
// bucket, which is also what the one NCCL all-reduce per step operates on.
#include "common.cuh"
#include "../../include/igmc_b200.h"

namespace {

constexpr int HID = IGMC_HIDDEN;
constexpr int L1O = IGMC_LIN1_OUT;

// dW_r[kj] of the adjacent-rating regulariser for one (layer, kj):  reg = sum_{r<R-1} ||W_{r+1}-W_r||^2,
// W_r = sum_b att[r,b] basis[b];  d reg / d W_r = 2 (W_r - W_{r-1}) [r>0] - 2 (W_{r+1} - W_r) [r<R-1].
__device__ __forceinline__ float arr_dw(const float* __restrict__ att, const float* __restrict__ bs, int NB, int KJ,
                                        int R, int r, int kj, float* reg_pair) {
  float wm = 0.f, w0 = 0.f, wp = 0.f;
  for (int b = 0; b < NB; ++b) {
    const float bv = bs[b * KJ + kj];
    w0 = fmaf(att[r * NB + b], bv, w0);
    if (r > 0) wm = fmaf(att[(r - 1) * NB + b], bv, wm);
    if (r < R - 1) wp = fmaf(att[(r + 1) * NB + b], bv, wp);
  }
  float dw = 0.f;
  if (r > 0) dw += 2.f * (w0 - wm);
  if (r < R - 1) { dw -= 2.f * (wp - w0); if (reg_pair) *reg_pair = (wp - w0) * (wp - w0); }
  else if (reg_pair) *reg_pair = 0.f;
  return dw;
}

// grad[p] for every parameter: conv params = fixed-order sum of the per-CTA partial rows (+ the ARR term
// for basis entries, thread-local); lin1/lin2 from the saved readout factors.  One thread per parameter.
__global__ void k_grad_reduce(igmc_model_t M, const float* __restrict__ params, int B, int rows,
                              const float* __restrict__ gpart, const float* __restrict__ dhid,
                              const float* __restrict__ feat, const float* __restrict__ hid,
                              const float* __restrict__ dpred, const float* __restrict__ sqerr, float loss_scale,
                              float arr, float grad_scale, float* __restrict__ grad, float* __restrict__ loss_out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int PC = M.conv_param_count, F = 2 * HID * M.num_layers;
  if (p < M.param_count) {
    float s = 0.f;
    if (p < PC) {
      for (int g = 0; g < rows; ++g) s += gpart[(size_t)g * PC + p];
      if (arr != 0.f) {
        for (int l = 0; l < M.num_layers; ++l) {
          const int in = l == 0 ? M.in_dim0 : HID, KJ = in * HID;
          const int q = p - M.off_basis[l];
          if (q >= 0 && q < M.num_bases * KJ) {   // basis[b][kj]:  += arr * sum_r att[r,b] dW_r[kj]
            const int b = q / KJ, kj = q - b * KJ;
            const float* att = params + M.off_att[l];
            float t = 0.f;
            for (int r = 0; r < M.num_relations; ++r)
              t = fmaf(att[r * M.num_bases + b],
                       arr_dw(att, params + M.off_basis[l], M.num_bases, KJ, M.num_relations, r, kj, nullptr), t);
            s = fmaf(arr, t, s);
          }
        }
      }
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
    grad[p] = s * grad_scale;
  }
  if (p == 0 && loss_out) {
    float s = 0.f;
    if (sqerr)
      for (int g = 0; g < B; ++g) s += sqerr[g];
    loss_out[0] = s * loss_scale;
  }
}

// ARR, att part + regulariser value: one CTA per layer, one warp per (r,b) (looped).
//   d att[r,b] += arr * < dW_r , basis[b] > ;   reg_l = sum_{r<R-1} ||W_{r+1}-W_r||^2
// The last CTA to finish adds arr * sum_l reg_l to loss_out in layer order (deterministic).
__global__ void __launch_bounds__(256)
k_arr_att(igmc_model_t M, const float* __restrict__ params, float arr, float grad_scale, float* __restrict__ grad,
          float* __restrict__ loss_out, float* __restrict__ reg_ws) {
  const int l = blockIdx.x, R = M.num_relations, NB = M.num_bases;
  const int in = l == 0 ? M.in_dim0 : HID, KJ = in * HID;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ float sm_arr[];                 // att [R*NB] | basis [NB*KJ]  (latency-bound otherwise)
  float* att = sm_arr;
  float* bs = sm_arr + ((R * NB + 3) & ~3);
  for (int i = tid; i < R * NB; i += 256) att[i] = params[M.off_att[l] + i];
  for (int i = tid; i < NB * KJ; i += 256) bs[i] = params[M.off_basis[l] + i];
  __shared__ float red[8];
  __shared__ int s_last;
  __syncthreads();
  for (int rb = warp; rb < R * NB; rb += 8) {
    const int r = rb / NB, b = rb - r * NB;
    float s = 0.f;
    for (int kj = lane; kj < KJ; kj += 32) s = fmaf(arr_dw(att, bs, NB, KJ, R, r, kj, nullptr), bs[b * KJ + kj], s);
    s = warp_sum_f(s);
    if (lane == 0) grad[M.off_att[l] + rb] += arr * grad_scale * s;
  }
  float reg = 0.f;
  for (int idx = tid; idx < (R - 1) * KJ; idx += 256) {
    const int r = idx / KJ, kj = idx - r * KJ;
    float pr;
    arr_dw(att, bs, NB, KJ, R, r, kj, &pr);
    reg += pr;
  }
  reg = warp_sum_f(reg);
  if (lane == 0) red[warp] = reg;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    reg_ws[l] = s;
    __threadfence();
    int* ticket = reinterpret_cast<int*>(reg_ws + IGMC_MAX_LAYERS);
    s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && tid == 0) {
    __threadfence();
    float s = 0.f;
    for (int q = 0; q < (int)gridDim.x; ++q) s += __ldcg(reg_ws + q);
    if (loss_out) loss_out[0] += arr * s;
    *reinterpret_cast<int*>(reg_ws + IGMC_MAX_LAYERS) = 0;   // re-arm for the next step
  }
}

__global__ void k_adam(float* __restrict__ params, const float* __restrict__ grad, float* __restrict__ m,
                       float* __restrict__ v, const int64_t* __restrict__ step_count, int n, float lr_val,
                       const float* __restrict__ lr_dev, float b1, float b2, float eps, float wd, float grad_mul) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr = lr_dev ? *lr_dev : lr_val;
  const double step = (double)(step_count[0] + 1);
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

__global__ void k_inc_step(int64_t* step_count) { step_count[0] += 1; }

}  // namespace

extern "C" int igmc_grad_reduce(const igmc_model_t* M, const float* params, int B, int gpart_rows, const float* gpart,
                                const float* dhid, const float* feat, const float* hid, const float* dpred,
                                const float* sqerr, float loss_scale, float arr, float grad_scale, float* grad,
                                float* loss_out, float* reg_ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (M->param_count + 255) / 256;
  k_grad_reduce<<<blocks, 256, 0, st>>>(*M, params, B, gpart_rows, gpart, dhid, feat, hid, dpred, sqerr, loss_scale,
                                        arr, grad_scale, grad, loss_out);
  IGMC_CUDA_CHECK_LAUNCH();
  if (arr != 0.f) {
    if (!reg_ws) return -18;
    const size_t smem = (size_t)(((M->num_relations * M->num_bases + 3) & ~3) + M->num_bases * HID * HID) * sizeof(float);
    k_arr_att<<<M->num_layers, 256, smem, st>>>(*M, params, arr, grad_scale, grad, loss_out, reg_ws);
    IGMC_CUDA_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int igmc_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                              int64_t* step_count, int n, float lr, const float* lr_dev, float beta1, float beta2,
                              float eps, float weight_decay, float grad_mul, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  k_adam<<<(n + 255) / 256, 256, 0, st>>>(params, grad, exp_avg, exp_avg_sq, step_count, n, lr, lr_dev, beta1, beta2, eps,
                                          weight_decay, grad_mul);
  IGMC_CUDA_CHECK_LAUNCH();
  k_inc_step<<<1, 1, 0, st>>>(step_count);
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
