// Fused optimizer update and episodic reset over the per-utterance trainable vectors.
//   optimizer_step  torch/optim/adam.py:484-545 (_single_tensor_adam) as called by REF/main.py:206, incl. the
//                   reference's duplicate-parameter semantics: a tensor listed k times by collect_params
//                   (REF/main.py:88-94) receives k sequential sub-steps per step() with the same gradient and a
//                   shared state -- fused here in registers, so HBM traffic is 28 B/param regardless of k.
//   params_reset    REF/main.py:147-155 (load_model_and_optimizer): restore the trainables from the pristine copy,
//                   clear both Adam moments (the snapshot is taken before any step, so state is empty).
// Element-wise; HBM-bound in train_feature mode (4.6 M params/utterance), launch-bound in LayerNorm-only mode.
#include "kernels.cuh"
#include <math.h>

namespace {

constexpr int MAXK = 12;    // largest multiplicity: 7 = an encoder Linear under train_all, 10 = a conv weight under train_all + train_feature (REF/main.py:81-100: once per enclosing module and flag)
struct AdamConsts {
  float step_size[MAXK + 1][MAXK];   // [k][j]: lr / (1 - beta1^s), s = k*step_index + j + 1
  float bc2_sqrt[MAXK + 1][MAXK];    // sqrt(1 - beta2^s)
  float inv_bc2_sqrt[MAXK + 1][MAXK];
};

__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void adam_kernel(AdamArgs a, AdamConsts c) {
  const long long total = a.n * a.n_utts;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i % a.n;
    const int k = a.mult ? a.mult[e] : 1;
    if (k == 0) continue;
    float p = a.P[i];
    const float g = a.G[i];
    if (a.kind == 1) {
      for (int j = 0; j < k; ++j) p -= a.lr * (g + a.weight_decay * p);          // torch/optim/sgd.py: grad.add(param, alpha=wd)
    } else {
      float m = a.Mom[i], v = a.Var[i];
      for (int j = 0; j < k; ++j) {
        float gj = g;
        if (a.weight_decay != 0.f) {
          if (a.kind == 2) gj = g + a.weight_decay * p;    // Adam: grad = grad.add(param, alpha=weight_decay)
          else p *= 1.0f - a.lr * a.weight_decay;          // AdamW: param.mul_(1 - lr * weight_decay)
        }
        m = m + (gj - m) * (1.0f - a.beta1);                // exp_avg.lerp_(grad, 1 - beta1)
        v = v * a.beta2 + (1.0f - a.beta2) * gj * gj;       // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
        const float denom = sqrtf(v) / c.bc2_sqrt[k][j] + a.eps;
        p = p - c.step_size[k][j] * (m / denom);            // param.addcdiv_(exp_avg, denom, value=-step_size)
      }
      a.Mom[i] = m;
      a.Var[i] = v;
    }
    a.P[i] = p;
    if (a.shadow) a.shadow[i] = __float2bfloat16(p);
  }
}

// float4 version: one utterance per blockIdx.y, 4 consecutive parameters per thread (all segment offsets are multiples
// of 4).  28 B/param of state traffic + 2 B for the bf16 operand copy, independent of the multiplicity k.
__global__ void __launch_bounds__(256)
adam_vec4_kernel(AdamArgs a, AdamConsts c) {
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e >= a.n) return;
  const long long i = (long long)blockIdx.y * a.n + e;
  float4 p = *reinterpret_cast<const float4*>(a.P + i);
  const float4 g = __ldcs(reinterpret_cast<const float4*>(a.G + i));       // streamed once: evict-first
  uchar4 k4 = make_uchar4(1, 1, 1, 1);
  if (a.mult) k4 = __ldg(reinterpret_cast<const uchar4*>(a.mult + e));
  float pv[4] = {p.x, p.y, p.z, p.w};
  const float gv[4] = {g.x, g.y, g.z, g.w};
  const int kv[4] = {k4.x, k4.y, k4.z, k4.w};
  if (a.kind == 1) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
      for (int j = 0; j < kv[t]; ++j) pv[t] -= a.lr * (gv[t] + a.weight_decay * pv[t]);
  } else {
    float4 m4 = __ldcs(reinterpret_cast<const float4*>(a.Mom + i)), v4 = __ldcs(reinterpret_cast<const float4*>(a.Var + i));
    float mv[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    const float omb1 = 1.0f - a.beta1, omb2 = 1.0f - a.beta2;
    const float decay = 1.0f - a.lr * a.weight_decay;
    if (kv[0] == kv[1] && kv[1] == kv[2] && kv[2] == kv[3] && a.weight_decay == 0.f) {
      // the usual case (segment sizes are multiples of 4): the four elements advance through the k sub-steps
      // together.  With k = 4 on every conv weight (REF/main.py:88-94) the update is instruction-bound unless the
      // square root and the two divisions are single MUFU operations.
      const int k = kv[0];
      float gg[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) gg[t] = omb2 * gv[t] * gv[t];
      for (int j = 0; j < k; ++j) {
        const float ss = c.step_size[k][j], ib = c.inv_bc2_sqrt[k][j];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          mv[t] = fmaf(gv[t] - mv[t], omb1, mv[t]);
          vv[t] = fmaf(vv[t], a.beta2, gg[t]);
          const float denom = fmaf(sqrt_approx(vv[t]), ib, a.eps);
          pv[t] = fmaf(-ss, __fdividef(mv[t], denom), pv[t]);
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int k = kv[t];
        for (int j = 0; j < k; ++j) {
          float gj = gv[t];
          if (a.weight_decay != 0.f) {
            if (a.kind == 2) gj = fmaf(a.weight_decay, pv[t], gj);   // Adam: L2 term joins the gradient
            else pv[t] *= decay;                                     // AdamW: decoupled decay
          }
          mv[t] = fmaf(gj - mv[t], omb1, mv[t]);
          vv[t] = fmaf(vv[t], a.beta2, omb2 * gj * gj);
          const float denom = fmaf(sqrt_approx(vv[t]), c.inv_bc2_sqrt[k][j], a.eps);
          pv[t] = fmaf(-c.step_size[k][j], __fdividef(mv[t], denom), pv[t]);
        }
      }
    }
    if (kv[0] | kv[1] | kv[2] | kv[3]) {
      __stcs(reinterpret_cast<float4*>(a.Mom + i), make_float4(mv[0], mv[1], mv[2], mv[3]));
      __stcs(reinterpret_cast<float4*>(a.Var + i), make_float4(vv[0], vv[1], vv[2], vv[3]));
    }
  }
  if (kv[0] | kv[1] | kv[2] | kv[3]) *reinterpret_cast<float4*>(a.P + i) = make_float4(pv[0], pv[1], pv[2], pv[3]);
  const uint2 packed = make_uint2(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]));
  if (a.shadow) *reinterpret_cast<uint2*>(a.shadow + i) = packed;
  for (int s = 0; s < a.n_seg; ++s) {
    const long long o = e - a.seg[s].off;
    if (o >= 0 && o < a.seg[s].size) *reinterpret_cast<uint2*>(a.seg[s].dst + (long long)blockIdx.y * a.seg[s].size + o) = packed;
  }
}

__global__ void reset_kernel(float* __restrict__ P, const float* __restrict__ P0, float* __restrict__ Mom,
                             float* __restrict__ Var, bf16* __restrict__ shadow, long long n, int n_utts) {
  const long long total = n * n_utts;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float p = P0[i % n];
    P[i] = p;
    Mom[i] = 0.f;
    Var[i] = 0.f;
    if (shadow) shadow[i] = __float2bfloat16(p);
  }
}

int grid_for(long long total) {
  long long b = (total + 255) / 256;
  long long cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

int optimizer_step(const AdamArgs& a, cudaStream_t stream) {
  SUTA_CHECK_ARG(a.P && a.G && a.n > 0 && a.n_utts > 0);
  SUTA_CHECK_ARG(a.kind == 1 || (a.Mom && a.Var));
  AdamConsts c;
  for (int k = 1; k <= MAXK; ++k)
    for (int j = 0; j < k; ++j) {
      double s = (double)k * a.step_index + j + 1;
      double bc1 = 1.0 - pow((double)a.beta1, s), bc2 = 1.0 - pow((double)a.beta2, s);
      c.step_size[k][j] = (float)((double)a.lr / bc1);
      c.bc2_sqrt[k][j] = (float)sqrt(bc2);
      c.inv_bc2_sqrt[k][j] = (float)(1.0 / sqrt(bc2));
    }
  bool vec = a.n % 4 == 0 && a.n_seg <= 8;
  for (int s = 0; s < a.n_seg; ++s) vec = vec && a.seg[s].off % 4 == 0 && a.seg[s].size % 4 == 0 && a.seg[s].dst;
  if (vec) {
    dim3 grid((unsigned)((a.n / 4 + 255) / 256), (unsigned)a.n_utts);
    adam_vec4_kernel<<<grid, 256, 0, stream>>>(a, c);
  } else {
    SUTA_CHECK_ARG(a.n_seg == 0);
    adam_kernel<<<grid_for(a.n * a.n_utts), 256, 0, stream>>>(a, c);
  }
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int params_reset(float* P, const float* P0, float* Mom, float* Var, bf16* shadow, long long n, int n_utts,
                 cudaStream_t stream) {
  SUTA_CHECK_ARG(P && P0 && Mom && Var && n > 0 && n_utts > 0);
  reset_kernel<<<grid_for(n * n_utts), 256, 0, stream>>>(P, P0, Mom, Var, shadow, n, n_utts);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
