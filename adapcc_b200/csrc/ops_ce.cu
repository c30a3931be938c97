// Fused softmax cross-entropy (forward + backward, in place) for the chunked LM head of the
// GPT-2 workload (adapcc_b200/models/gpt2.py).
//
// One CTA per row of bf16 logits [rows, stride] (stride = vocab padded to a multiple of 8 so every
// row is 16-byte aligned): pass 1 computes the online max / sum-exp with 128-bit loads, pass 2
// re-reads the row (L2 resident: a row is ~100 KB) and overwrites it with
// d loss / d logits = (softmax - onehot) * valid, also in bf16. Per-row losses go to a float
// array. Columns >= vocab are padding: they read as -inf and get zero gradient. Compared with the
// eager chain (float cast, logsumexp, softmax, scatter, mask, cast back: ~17 bytes/logit of HBM
// traffic) this moves 4 bytes/logit.
#include <cuda_bf16.h>

#include "common.h"
#include "device_prims.cuh"

namespace adapcc {

__device__ __forceinline__ void online_combine(float& m, float& s, float m2, float s2) {
  const float mx = fmaxf(m, m2);
  s = s * __expf(m - mx) + s2 * __expf(m2 - mx);
  m = mx;
}

__global__ void __launch_bounds__(512)
fused_ce_kernel(__nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels,
                float* __restrict__ row_loss, int vocab, int stride, const float* __restrict__ grad_scale) {
  const float gs = grad_scale ? *grad_scale : 1.f;       // d(total loss)/d(row loss), folded into the gradient
  const int row = blockIdx.x;
  __nv_bfloat16* p = logits + (long long)row * stride;
  const long long label = labels[row];
  const bool valid = label >= 0 && label < vocab;
  const int nvec = stride / 8;

  float m = -INFINITY, s = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    uint4 q = reinterpret_cast<const uint4*>(p)[v];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    float x[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[2 * i] = __uint_as_float(w[i] << 16);
      x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    float lm = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (v * 8 + i >= vocab) x[i] = -INFINITY;
      lm = fmaxf(lm, x[i]);
    }
    if (lm > -INFINITY) {
      float ls = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) ls += __expf(x[i] - lm);
      online_combine(m, s, lm, ls);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    if (m2 > -INFINITY || m > -INFINITY) online_combine(m, s, m2, s2);
  }
  __shared__ float sm[16], ss[16];
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
  __syncthreads();
  m = sm[0]; s = ss[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (sm[w] > -INFINITY || m > -INFINITY) online_combine(m, s, sm[w], ss[w]);
  const float inv = 1.f / s;
  if (threadIdx.x == 0) {
    const float xl = valid ? __bfloat162float(p[label]) : 0.f;
    row_loss[row] = valid ? (__logf(s) + m - xl) : 0.f;
  }
  __syncthreads();   // the label logit above must be read before the row is overwritten
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    uint4 q = reinterpret_cast<const uint4*>(p)[v];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = __uint_as_float(w[i] << 16), b = __uint_as_float(w[i] & 0xffff0000u);
      const int c = v * 8 + 2 * i;
      float ga = (valid && c < vocab) ? (__expf(a - m) * inv - (c == label ? 1.f : 0.f)) * gs : 0.f;
      float gb = (valid && c + 1 < vocab) ? (__expf(b - m) * inv - (c + 1 == label ? 1.f : 0.f)) * gs : 0.f;
      __nv_bfloat162 r = __floats2bfloat162_rn(ga, gb);
      o[i] = *reinterpret_cast<uint32_t*>(&r);
    }
    reinterpret_cast<uint4*>(p)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// v2: the row is staged in shared memory (a 50 304-wide bf16 row is 100 KB; two CTAs fit per SM), so
// global memory is read once and written once and every logit costs ONE ex2 instead of two: pass 1
// loads + row max, pass 2 e = exp(x - max) (kept in smem as bf16 — the output precision) + sum,
// pass 3 gradient = e / sum - onehot. ncu on v1 showed it SFU/ALU-bound (83 % SM, 32 % DRAM).
__global__ void __launch_bounds__(512)
fused_ce_smem_kernel(__nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels,
                     float* __restrict__ row_loss, int vocab, int stride, const float* __restrict__ grad_scale) {
  const float gs = grad_scale ? *grad_scale : 1.f;
  extern __shared__ uint4 srow[];
  const int row = blockIdx.x;
  __nv_bfloat16* p = logits + (long long)row * stride;
  const long long label = labels[row];
  const bool valid = label >= 0 && label < vocab;
  const int nvec = stride / 8;
  __shared__ float red[16];
  __shared__ float bcast[2];

  if (!valid) {
    // ignored row (label -100): loss 0, gradient exactly 0 — nothing to read, no exponentials; the row is only
    // overwritten with zeros for the backward GEMMs. PersonaChat-shaped batches ignore 7 of 8 rows.
    // (`valid` is uniform over the CTA, so the whole block leaves before the first barrier.)
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) st16(reinterpret_cast<uint4*>(p) + v, z);
    if (threadIdx.x == 0) row_loss[row] = 0.f;
    return;
  }

  float m = -INFINITY;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    uint4 q = ld16(reinterpret_cast<const uint4*>(p) + v);
    srow[v] = q;
    float x[8];
    unpack<__nv_bfloat16>(q, x);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (v * 8 + i < vocab) m = fmaxf(m, x[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mm = fmaxf(mm, red[w]);
    bcast[0] = mm;
  }
  __syncthreads();
  m = bcast[0];
  const float xl = valid ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(srow)[label]) : 0.f;
  __syncthreads();                       // label logit read before the row turns into exponentials

  float s = 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float x[8];
    unpack<__nv_bfloat16>(srow[v], x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i] = (v * 8 + i < vocab) ? __expf(x[i] - m) : 0.f;
      s += x[i];
    }
    srow[v] = pack<__nv_bfloat16>(x);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float ss = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) ss += red[w];
    bcast[1] = ss;
    row_loss[row] = valid ? (__logf(ss) + m - xl) : 0.f;
  }
  __syncthreads();
  const float inv = valid ? gs / bcast[1] : 0.f;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    float e[8];
    unpack<__nv_bfloat16>(srow[v], e);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = v * 8 + i;
      e[i] = e[i] * inv - ((valid && c == label) ? gs : 0.f);
    }
    st16(reinterpret_cast<uint4*>(p) + v, pack<__nv_bfloat16>(e));
  }
}

}  // namespace adapcc

// grad_scale (optional device scalar): the gradient written into `logits` is multiplied by it (e.g. 1 / #scored
// rows of a mean loss), so the GEMMs that consume it produce final gradients; row losses stay unscaled.
extern "C" int adapcc_fused_ce_scaled(void* logits, const long long* labels, float* row_loss, int rows, int vocab,
                                      int stride, const float* grad_scale, void* stream) {
  using namespace adapcc;
  if (rows <= 0) return 0;
  if (stride % 8 != 0 || (reinterpret_cast<uintptr_t>(logits) & 15)) {
    set_error("fused_ce: stride must be a multiple of 8 and logits 16-byte aligned");
    return -1;
  }
  if (vocab > stride) { set_error("fused_ce: vocab > stride"); return -1; }
  const size_t smem = (size_t)stride * 2;
  static int smem_ok = -1;
  if (smem_ok < 0) {
    smem_ok = cudaFuncSetAttribute(fused_ce_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) ==
              cudaSuccess;
    (void)cudaGetLastError();
  }
  const char* force = getenv("ADAPCC_CE_V1");
  if (smem_ok && smem <= 200 * 1024 && !(force && atoi(force)))
    fused_ce_smem_kernel<<<rows, 512, smem, (cudaStream_t)stream>>>((__nv_bfloat16*)logits, labels, row_loss, vocab,
                                                                    stride, grad_scale);
  else
    fused_ce_kernel<<<rows, 512, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)logits, labels, row_loss, vocab, stride,
                                                            grad_scale);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int adapcc_fused_ce(void* logits, const long long* labels, float* row_loss, int rows, int vocab,
                               int stride, void* stream) {
  return adapcc_fused_ce_scaled(logits, labels, row_loss, rows, vocab, stride, nullptr, stream);
}
