// Sharded optimizer step fused with its collective (ZeRO-1 over NVSwitch) — the engine's default at N > 1 (parity-tested).
//
// Data-parallel training ends every step with: all-reduce(grads) -> every rank runs the SAME AdamW over ALL
// parameters (the reference: DDP + optimizer.step per rank, /root/reference/train_ddp.py:37-54). On one NVSwitch
// box that is N-fold redundant HBM traffic (30 bytes per parameter per rank) and an all-reduce that moves twice
// what is needed. Here instead:
//   backward : bucket hooks REDUCE-SCATTER (rank r gets the averaged slice r of the bucket in place — the
//              existing direct kernels with root = self; NVLS multimem.ld_reduce when bound)
//   step end : sum of squares of my slices -> 4-byte all-reduce (global grad norm; doubles as the "every rank
//              has finished reading the old parameters" barrier)
//   THIS FILE: AdamW over my slices only (fp32 master/m/v for 1/N of the parameters), and the updated bf16
//              parameters leave the kernel through multimem.st on the multicast alias of the parameter
//              buffer — the switch replicates them into every rank's copy (the all-gather), no separate pass.
//              Fallback without multicast: one 128-bit store per peer.
//   then     : one device barrier before the next forward.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.h"
#include "device_prims.cuh"

namespace adapcc {

struct ZeroPeers {
  char* p[kMaxRanks];       // p[i]: base of rank i's parameter buffer in my address space (fallback path)
  int n;
};

struct ZeroAdam {
  float lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale;
};

// Elements [0, n) of a shard; n % 8 == 0. `param_off_bytes` = byte offset of the shard inside the parameter buffer
// (identical on every rank: the buffer is symmetric). master/m/v/grad point at the shard's first element.
__global__ void __launch_bounds__(512)
zero_adamw_bcast_kernel(char* __restrict__ param_mc, ZeroPeers peers, long long param_off_bytes,
                        const __nv_bfloat16* __restrict__ grad, float* __restrict__ master, float* __restrict__ m,
                        float* __restrict__ v, long long n, ZeroAdam a, const float* __restrict__ sumsq,
                        const int* __restrict__ step_ptr) {
  const float t = (float)(*step_ptr);
  const float bias1 = 1.f - powf(a.beta1, t), bias2 = 1.f - powf(a.beta2, t);
  float coef = a.grad_scale;
  if (a.max_norm > 0.f && sumsq != nullptr) {
    const float norm = sqrtf(*sumsq) * a.grad_scale;
    coef *= fminf(1.f, a.max_norm / (norm + 1e-6f));
  }
  const long long nv = n / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const long long e0 = i * 8;
    float g[8], w[8], mm[8], vv[8];
    unpack<__nv_bfloat16>(ld16(grad + e0), g);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(w + 4 * h) = *reinterpret_cast<const float4*>(master + e0 + 4 * h);
      *reinterpret_cast<float4*>(mm + 4 * h) = *reinterpret_cast<const float4*>(m + e0 + 4 * h);
      *reinterpret_cast<float4*>(vv + 4 * h) = *reinterpret_cast<const float4*>(v + e0 + 4 * h);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float gk = g[k] * coef;
      mm[k] = a.beta1 * mm[k] + (1.f - a.beta1) * gk;
      vv[k] = a.beta2 * vv[k] + (1.f - a.beta2) * gk * gk;
      w[k] = w[k] * (1.f - a.lr * a.weight_decay) - a.lr * (mm[k] / bias1) / (sqrtf(vv[k] / bias2) + a.eps);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(master + e0 + 4 * h) = *reinterpret_cast<const float4*>(w + 4 * h);
      *reinterpret_cast<float4*>(m + e0 + 4 * h) = *reinterpret_cast<const float4*>(mm + 4 * h);
      *reinterpret_cast<float4*>(v + e0 + 4 * h) = *reinterpret_cast<const float4*>(vv + 4 * h);
    }
    const uint4 packed = pack<__nv_bfloat16>(w);
    const long long off = param_off_bytes + e0 * 2;
    if (param_mc != nullptr) {
      mc_st16(param_mc + off, packed);                         // the switch writes every rank's copy
    } else {
      for (int r = 0; r < peers.n; ++r) st16(peers.p[r] + off, packed);
    }
  }
}

}  // namespace adapcc

using namespace adapcc;

extern "C" {

// param_mc: multicast alias of the (symmetric) bf16 parameter buffer, or NULL -> peer_params[0..n_peers) are
// written one by one (must include this rank's own buffer). Everything 16-byte aligned, n % 8 == 0.
int adapcc_zero_adamw_bcast(void* param_mc, void* const* peer_params, int n_peers, long long param_off_bytes,
                            const void* grad, float* master, float* m, float* v, long long n, float lr, float beta1,
                            float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                            const float* sumsq, const int* step_ptr, void* stream) {
  if (n <= 0) return 0;
  if (n % 8 != 0 || (param_off_bytes & 15)) { set_error("zero_adamw: shard must be a multiple of 8 elements at a 16-byte offset"); return -1; }
  if (((uintptr_t)grad | (uintptr_t)master | (uintptr_t)m | (uintptr_t)v | (uintptr_t)param_mc) & 15) { set_error("zero_adamw: buffers must be 16-byte aligned"); return -1; }
  if (step_ptr == nullptr) { set_error("zero_adamw: needs the device step counter"); return -1; }
  if (param_mc == nullptr && (n_peers <= 0 || n_peers > kMaxRanks || peer_params == nullptr)) { set_error("zero_adamw: no multicast alias and no peer list"); return -1; }
  ZeroPeers peers{};
  peers.n = param_mc ? 0 : n_peers;
  for (int r = 0; r < peers.n; ++r) {
    if ((uintptr_t)peer_params[r] & 15) { set_error("zero_adamw: peer buffer %d misaligned", r); return -1; }
    peers.p[r] = (char*)peer_params[r];
  }
  ZeroAdam a{lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale};
  int blocks = (int)std::min<long long>(592, (n / 8 + 511) / 512);
  if (blocks < 1) blocks = 1;
  zero_adamw_bcast_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>((char*)param_mc, peers, param_off_bytes,
                                                                    (const __nv_bfloat16*)grad, master, m, v, n, a,
                                                                    sumsq, step_ptr);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

}  // extern "C"
