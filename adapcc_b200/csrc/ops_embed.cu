// Fused embedding sum (forward + backward) for the transformer workloads.
//
// GPT-2's input layer is x[t] = wte[ids[t]] + wpe[pos[t]] + wte[token_type[t]]
// (/root/reference/models/gpt2/train_gpt2_ddp.py:157-159 -> transformers' GPT2Model): three gathers and two adds
// forward, and — the expensive part — three sort-based `embedding_dense_backward` passes (16 radix-sort launches,
// segment scans, 0.35 ms per step on B200) plus two 77 MB accumulate kernels backward. Here:
//
//   forward : one kernel, one warp per token row, the K looked-up rows summed in fp32 registers, bf16 out.
//   backward: three small launches, no sort, no host-visible sizes (CUDA-graph capturable):
//     1. claim   : every lookup (k, t) does atomicMin(owner[table_k, idx], lookup id) — the smallest lookup id that
//                  hits a table row becomes the row's OWNER for this step;
//     2. scatter : every lookup adds its dY row into the owner's fp32 scratch row with red.global.add.v4.f32
//                  (duplicates — repeated tokens, the two token-type rows hit by thousands of tokens — accumulate in
//                  fp32, like ATen's segment reduction, not in bf16); rows hit many times within a 64-lookup tile are
//                  pre-summed in registers first;
//     3. commit  : each owner adds its scratch row into the table's bf16 gradient row (which may already hold another
//                  contribution, e.g. the tied LM head's dW), zeroes the scratch row and releases the owner slot, so
//                  both work buffers are back to their initial state for the next step.
//   Lookups that share a table (ids and token types both index wte) share owner slots, so a row hit through both is
//   still committed exactly once.
#include <cuda_bf16.h>

#include <climits>

#include "common.h"
#include "device_prims.cuh"

namespace adapcc {

constexpr int kEmbedMaxLookups = 4;

struct EmbedFwdArgs {
  const __nv_bfloat16* table[kEmbedMaxLookups];
  const long long* idx[kEmbedMaxLookups];
  int K;
};

struct EmbedBwdArgs {
  const long long* idx[kEmbedMaxLookups];
  __nv_bfloat16* grad[kEmbedMaxLookups];
  int owner_base[kEmbedMaxLookups];      // first owner slot of lookup k's table
  int K;
};

// y[t, :] = sum_k table_k[idx_k[t], :]; D % 8 == 0; one warp per row.
__global__ void __launch_bounds__(256)
embed_sum_fwd_kernel(const __grid_constant__ EmbedFwdArgs a, int n, int D, __nv_bfloat16* __restrict__ out) {
  const int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int nvec = D >> 3;
  const __nv_bfloat16* rows[kEmbedMaxLookups];
#pragma unroll
  for (int k = 0; k < kEmbedMaxLookups; ++k)
    rows[k] = k < a.K ? a.table[k] + a.idx[k][warp] * (long long)D : nullptr;
  for (int c = lane; c < nvec; c += 32) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int k = 0; k < kEmbedMaxLookups; ++k) {
      if (k < a.K) {
        float f[8];
        unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(rows[k] + c * 8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
    }
    st16(out + (long long)warp * D + c * 8, pack<__nv_bfloat16>(acc));
  }
}

__global__ void __launch_bounds__(256)
embed_bwd_claim_kernel(const __grid_constant__ EmbedBwdArgs a, int n, int* __restrict__ owner) {
  const int i = (int)(blockIdx.x * (long long)blockDim.x + threadIdx.x);
  if (i >= a.K * n) return;
  const int k = i / n, t = i - k * n;
  atomicMin(owner + a.owner_base[k] + (int)a.idx[k][t], i);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// scratch[owner, :] += dY[t, :] for a TILE of 64 consecutive lookups per CTA. Rows hit by many lookups of the tile (the
// two token-type rows are hit by every token) are first summed in registers and leave as ONE reduction per tile and
// thread group instead of one per lookup: up to 4 "hot" owners per tile (>= 4 hits), found with 64 x 64 compares in
// shared memory. Everything else (token ids, positions: mostly unique inside a tile) goes straight to
// red.global.add.v4.f32. First version (one warp per lookup, every lookup its own reductions): 136 us of the GPT-2 step,
// 4096-deep same-address contention on the token-type rows.
constexpr int kEmbTile = 64, kEmbHot = 4, kEmbGroup = 128;       // 2 thread groups x 128 threads; group g: lookups j = g mod 2

__global__ void __launch_bounds__(2 * kEmbGroup)
embed_bwd_scatter_kernel(const __grid_constant__ EmbedBwdArgs a, int n, int D, const __nv_bfloat16* __restrict__ dy,
                         const int* __restrict__ owner, float* __restrict__ scratch) {
  __shared__ int s_own[kEmbTile];
  __shared__ int s_row[kEmbTile];
  __shared__ int s_hot_key[kEmbHot];
  __shared__ int s_hot_of[kEmbTile];
  __shared__ int s_nhot;
  const int total = a.K * n;
  const int i0 = blockIdx.x * kEmbTile;
  const int j = threadIdx.x;
  if (j == 0) s_nhot = 0;
  if (j < kEmbTile) {
    const int i = i0 + j;
    int own = -1, row = 0;
    if (i < total) {
      const int k = i / n;
      row = i - k * n;
      own = owner[a.owner_base[k] + (int)a.idx[k][row]];
    }
    s_own[j] = own;
    s_row[j] = row;
  }
  __syncthreads();
  if (j < kEmbTile && s_own[j] >= 0) {
    const int own = s_own[j];
    int cnt = 0;
    bool first = true;
    for (int i2 = 0; i2 < kEmbTile; ++i2) {
      if (s_own[i2] == own) {
        ++cnt;
        if (i2 < j) first = false;
      }
    }
    if (first && cnt >= 4) {
      const int slot = atomicAdd(&s_nhot, 1);
      if (slot < kEmbHot) s_hot_key[slot] = own;
    }
  }
  __syncthreads();
  const int nhot = min(s_nhot, kEmbHot);
  if (j < kEmbTile) {
    int h = -1;
    for (int q = 0; q < nhot; ++q)
      if (s_hot_key[q] == s_own[j]) h = q;
    s_hot_of[j] = h;
  }
  __syncthreads();
  const int g = threadIdx.x / kEmbGroup, tg = threadIdx.x - g * kEmbGroup;
  const int nvec = D >> 3;
  for (int c = tg; c < nvec; c += kEmbGroup) {
    float acc[kEmbHot][8];
#pragma unroll
    for (int q = 0; q < kEmbHot; ++q)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[q][e] = 0.f;
    for (int jj = g; jj < kEmbTile; jj += 2) {
      const int own = s_own[jj];
      if (own < 0) continue;
      float f[8];
      unpack<__nv_bfloat16>(ld16(dy + (long long)s_row[jj] * D + c * 8), f);
      const int h = s_hot_of[jj];
      if (h < 0) {
        float* dst = scratch + (long long)own * D + c * 8;
        red_add_v4(dst, f[0], f[1], f[2], f[3]);
        red_add_v4(dst + 4, f[4], f[5], f[6], f[7]);
      } else {
#pragma unroll
        for (int q = 0; q < kEmbHot; ++q)
          if (q == h) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[q][e] += f[e];
          }
      }
    }
#pragma unroll
    for (int q = 0; q < kEmbHot; ++q)
      if (q < nhot) {
        float* dst = scratch + (long long)s_hot_key[q] * D + c * 8;
        red_add_v4(dst, acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        red_add_v4(dst + 4, acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
      }
  }
}

// one warp per lookup; only owners act: grad[idx, :] += scratch[i, :]; scratch[i, :] = 0; owner slot released
__global__ void __launch_bounds__(256)
embed_bwd_commit_kernel(const __grid_constant__ EmbedBwdArgs a, int n, int D, int* __restrict__ owner,
                        float* __restrict__ scratch) {
  const int i = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= a.K * n) return;
  const int k = i / n, t = i - k * n;
  const long long row = a.idx[k][t];
  int* slot = owner + a.owner_base[k] + (int)row;
  if (*slot != i) return;                       // warp-uniform: every lane reads the same word
  float* s = scratch + (long long)i * D;
  __nv_bfloat16* g = a.grad[k] + row * (long long)D;
  const int nvec = D >> 3;
  for (int c = lane; c < nvec; c += 32) {
    float f[8];
    unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(g + c * 8), f);
    const float4 s0 = *reinterpret_cast<const float4*>(s + c * 8), s1 = *reinterpret_cast<const float4*>(s + c * 8 + 4);
    f[0] += s0.x; f[1] += s0.y; f[2] += s0.z; f[3] += s0.w;
    f[4] += s1.x; f[5] += s1.y; f[6] += s1.z; f[7] += s1.w;
    *reinterpret_cast<uint4*>(g + c * 8) = pack<__nv_bfloat16>(f);
    *reinterpret_cast<float4*>(s + c * 8) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(s + c * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  if (lane == 0) *slot = INT_MAX;
}

__global__ void embed_fill_int_kernel(int* p, long long n, int v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace adapcc

using namespace adapcc;

extern "C" {

// tables[k]: bf16 [rows_k, D]; idx[k]: int64 [n]; out: bf16 [n, D]
int adapcc_embed_sum_fwd(void* const* tables, void* const* idx, int K, int n, int D, void* out, void* stream) {
  if (n <= 0) return 0;
  if (K < 1 || K > kEmbedMaxLookups || D % 8 != 0) { set_error("embed_sum_fwd: 1 <= K <= %d lookups, D %% 8 == 0", kEmbedMaxLookups); return -1; }
  EmbedFwdArgs a{};
  a.K = K;
  for (int k = 0; k < K; ++k) {
    if (((uintptr_t)tables[k] | (uintptr_t)out) & 15) { set_error("embed_sum_fwd: tables / output must be 16-byte aligned"); return -1; }
    a.table[k] = (const __nv_bfloat16*)tables[k];
    a.idx[k] = (const long long*)idx[k];
  }
  const int wpb = 8;
  embed_sum_fwd_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(a, n, D, (__nv_bfloat16*)out);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

// owner: int32 [sum of distinct table rows], all INT_MAX between calls; scratch: fp32 [K * n, D], all zero between
// calls (adapcc_embed_bwd_reset establishes both). grads[k]: bf16 [rows_k, D], accumulated into.
int adapcc_embed_sum_bwd(const void* dy, void* const* idx, void* const* grads, const int* owner_base, int K, int n,
                         int D, int* owner, float* scratch, void* stream) {
  if (n <= 0) return 0;
  if (K < 1 || K > kEmbedMaxLookups || D % 8 != 0) { set_error("embed_sum_bwd: 1 <= K <= %d lookups, D %% 8 == 0", kEmbedMaxLookups); return -1; }
  if ((long long)K * n >= INT_MAX) { set_error("embed_sum_bwd: too many lookups"); return -1; }
  EmbedBwdArgs a{};
  a.K = K;
  for (int k = 0; k < K; ++k) {
    if (((uintptr_t)grads[k] | (uintptr_t)dy | (uintptr_t)scratch) & 15) { set_error("embed_sum_bwd: buffers must be 16-byte aligned"); return -1; }
    a.idx[k] = (const long long*)idx[k];
    a.grad[k] = (__nv_bfloat16*)grads[k];
    a.owner_base[k] = owner_base[k];
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int total = K * n;
  embed_bwd_claim_kernel<<<(total + 255) / 256, 256, 0, s>>>(a, n, owner);
  const int wpb = 8;
  embed_bwd_scatter_kernel<<<(total + kEmbTile - 1) / kEmbTile, 2 * kEmbGroup, 0, s>>>(a, n, D, (const __nv_bfloat16*)dy, owner,
                                                                                   scratch);
  embed_bwd_commit_kernel<<<(total + wpb - 1) / wpb, wpb * 32, 0, s>>>(a, n, D, owner, scratch);
  CUDA_TRY(cudaGetLastError());
  count_launch(3);
  return 0;
}

int adapcc_embed_bwd_reset(int* owner, long long n_owner, float* scratch, long long scratch_elems, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (n_owner > 0) embed_fill_int_kernel<<<(int)std::min<long long>(1024, (n_owner + 255) / 256), 256, 0, s>>>(owner, n_owner, INT_MAX);
  if (scratch_elems > 0) CUDA_TRY(cudaMemsetAsync(scratch, 0, (size_t)scratch_elems * sizeof(float), s));
  CUDA_TRY(cudaGetLastError());
  if (n_owner > 0) count_launch();
  return 0;
}

}  // extern "C"
