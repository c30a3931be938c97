// tcgen05 / TMEM / TMA GEMM with a fused epilogue. Numerically validated on B200 (tests/test_gpu_tcgen05.py: five
// shapes incl. ragged M, both tile widths, K up to 3072, autograd); SASS: UTCHMMA, UTCBAR, UTMALDG.2D, LDTM.x32.
// Variants 0-2 of this file are timed in profiles/gemm_tcgen05.md; the tuned kernel on the model path is variant 3 (gemm_tcgen05_pp.cu).
//
// out[M, N] = act(A[M, K] · W[N, K]^T + bias[N])          (bf16 in, fp32 accumulate, bf16 out)
// optionally also pre[M, N] = A · W^T + bias               (what the activation's backward needs)
//
// This is the transformer MLP up-projection with its bias and GELU folded into the GEMM epilogue: the
// accumulator tile never leaves the SM between the tensor-core pass and the activation — it sits in
// TMEM, is read back with tcgen05.ld, gets bias + GELU in registers and goes to HBM once (stock path:
// cuBLASLt GEMM+bias writes [M, N], a separate GELU kernel re-reads and re-writes it;
// profiles/gpt2_step_kernels_torch_profiler.md: 24 such elementwise launches, ~0.5 ms of a 9.6 ms step).
// The reference has no GEMM of its own (its workloads call torch/HF modules,
// /root/reference/models/gpt2/train_gpt2_ddp.py:157-159); this is a B200-side addition.
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0, one lane : TMA producer  — cp.async.bulk.tensor.2d of the A (128 x 64) and W (BN x 64) K-slabs,
//                                       128B-swizzled, into a 4-stage smem ring; completion on `full[s]`
//   warp 1           : TMEM allocator (BN fp32 columns) and, one lane, the MMA issuer —
//                      4 x tcgen05.mma.cta_group::1.kind::f16 (M128 x N{BN} x K16) per stage, accumulator in
//                      TMEM; tcgen05.commit releases the stage (`empty[s]`) and finally signals `acc_full`
//   warps 2..5       : epilogue — each warp owns the TMEM lane quarter (warp % 4), reads 32 columns at a
//                      time (tcgen05.ld.32x32b.x32), bias + GELU(tanh) in fp32, packs bf16, 64-byte stores
// All mbarrier waits are bounded by wall clock (trap after 2 s instead of hanging the GPU).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <mutex>

#include "common.h"

namespace adapcc {
namespace tc {

constexpr int kBM = 128;          // tile rows   (UMMA M, cta_group::1)
constexpr int kBK = 64;           // K slab      (64 bf16 = 128 B = one swizzle-128B row)
constexpr int kUmmaK = 16;        // K per tcgen05.mma for 16-bit inputs
constexpr int kStages = 4;
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded by WALL CLOCK (%globaltimer), not by poll count: how long one try_wait may suspend is
// implementation-defined. A protocol bug traps after 2 s instead of hanging the box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
#pragma unroll 1
  for (;;) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (t - t0 > 2000000000ull) __trap();
  }
}

// ---- TMA -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c_inner, int c_outer,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}

// ---- tcgen05 -------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {       // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t taddr) {           // whole warp, the allocating one
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem desc] · B[smem desc]; `accumulate` = 0 overwrites the accumulator.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane quarter base + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- descriptors (bit layouts: cute/arch/mma_sm100_desc.hpp of the vendored CUTLASS 4.5) ----------
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B, 8-row groups
// 1024 B apart (SBO); LBO is unused for swizzled K-major layouts; version 1 = sm_100.
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // [0,14)  start address / 16
  d |= (uint64_t)1 << 16;                              // [16,30) leading byte offset / 16 (canonical value 1)
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;        // [32,46) stride byte offset / 16
  d |= (uint64_t)1 << 46;                              // [46,48) descriptor version
  d |= (uint64_t)2 << 61;                              // [61,64) SWIZZLE_128B
  return d;
}
// Instruction descriptor: D fp32, A/B bf16, both K-major, dense, M x N.
__host__ __device__ constexpr uint32_t instr_desc_bf16(int m, int n) {
  return (1u << 4)                 // [4,6)   D format  = F32
         | (1u << 7)               // [7,10)  A format  = BF16
         | (1u << 10)              // [10,13) B format  = BF16
         | ((uint32_t)(n >> 3) << 17)   // [17,23) N / 8
         | ((uint32_t)(m >> 4) << 24);  // [24,29) M / 16
}

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) = x * sigmoid(2 u)
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return __fdividef(x, 1.f + __expf(-2.f * u));
}

// d/dx of the tanh-form GELU above: s + x s (1 - s) 2k (1 + 3c x^2), s = sigmoid(2u)
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k = 0.7978845608028654f, c = 0.044715f;
  const float u = k * (x + c * x * x * x);
  const float sg = __fdividef(1.f, 1.f + __expf(-2.f * u));
  return sg + x * sg * (1.f - sg) * 2.f * k * (1.f + 3.f * c * x * x);
}

template <int BN>
struct Smem {
  static constexpr uint32_t kABytes = kBM * kBK * 2;            // 16 KB
  static constexpr uint32_t kBBytes = BN * kBK * 2;             // 16 / 32 KB
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kBarOffset = kStages * kStageBytes;  // full[S], empty[S], acc_full, tmem ptr
  static constexpr uint32_t kTotal = kBarOffset + (2 * kStages + 1) * 8 + 16;
  static constexpr uint32_t kDynamic = kTotal + 1024;            // slack for the manual 1024-B alignment
};

// ACT: 0 = identity, 1 = GELU(tanh)                      -> out = act(acc + bias), optional pre = acc + bias
//      2 = dGELU   : out = acc * gelu'(aux)               (backward of the MLP: dY.W2 with the activation's derivative
//                                                          applied in the epilogue; aux = saved pre-activation)
//      3 = residual: out = acc + bias + aux               (projection + residual add)
// For ACT >= 2 the `pre` argument is an INPUT (aux, same shape as out) and nothing is written to it.
template <int BN, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bias_act_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                             const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                             __nv_bfloat16* __restrict__ pre, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // swizzle-128B tiles need 1024-B alignment
  using S = Smem<BN>;
  const uint32_t bar0 = base + S::kBarOffset;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  const uint32_t acc_full = bar0 + 8u * (2 * kStages);
  const uint32_t tmem_slot = acc_full + 8u;
  auto smem_a = [&](int s) { return base + (uint32_t)s * S::kStageBytes; };
  auto smem_b = [&](int s) { return base + (uint32_t)s * S::kStageBytes + S::kABytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x, m_blk = blockIdx.y;
  const int num_kb = K / kBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        mbar_wait(empty(s), ph ^ 1u);                        // first pass over the ring falls through
        mbar_expect_tx(full(s), S::kStageBytes);
        tma_load_2d(smem_a(s), &map_a, kb * kBK, m_blk * kBM, full(s));
        tma_load_2d(smem_b(s), &map_w, kb * kBK, n_blk * BN, full(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = instr_desc_bf16(kBM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (uint32_t)(kb / kStages) & 1u;
        mbar_wait(full(s), ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          // advancing K inside the 128-B swizzle atom = advancing the start address (32 B per UMMA_K)
          const uint64_t da = smem_desc_k_sw128(smem_a(s) + (uint32_t)k * kUmmaK * 2);
          const uint64_t db = smem_desc_k_sw128(smem_b(s) + (uint32_t)k * kUmmaK * 2);
          umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty(s));                               // smem slot reusable once these MMAs retire
      }
      umma_commit(acc_full);                                 // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 -> TMEM lane quarters 2, 3, 0, 1 =====
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int row = m_blk * kBM + q * 32 + lane;
    const size_t row_off = (size_t)row * (size_t)N + (size_t)n_blk * BN;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t acc[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), acc);
      float bf[32];
      if (bias != nullptr) {                                   // 32 bias values = 4 x 16-byte loads (64-B aligned)
        const uint4* bp = reinterpret_cast<const uint4*>(bias + (size_t)n_blk * BN + c * 32);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint4 t = __ldg(bp + v);
          const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
            bf[v * 8 + 2 * j] = f.x;
            bf[v * 8 + 2 * j + 1] = f.y;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) bf[i] = 0.f;
      }
      uint32_t packed_out[16], packed_pre[16];
      if (ACT >= 2) {
        // aux tile: this thread's 32 bf16 of the same row / columns (64 contiguous bytes)
        uint32_t ax[16];
        if (row < M) {
          const uint4* ap = reinterpret_cast<const uint4*>(pre + row_off + c * 32);
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 t = __ldg(ap + v);
            ax[4 * v] = t.x; ax[4 * v + 1] = t.y; ax[4 * v + 2] = t.z; ax[4 * v + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) ax[i] = 0u;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ax[i]));
          float o0, o1;
          if (ACT == 2) {
            o0 = __uint_as_float(acc[2 * i]) * gelu_tanh_grad(a2.x);
            o1 = __uint_as_float(acc[2 * i + 1]) * gelu_tanh_grad(a2.y);
          } else {
            // bf16-round the projection first, then add: what `x + linear(h)` computes on bf16 tensors
            const float2 pr = __bfloat1622float2(__floats2bfloat162_rn(__uint_as_float(acc[2 * i]) + bf[2 * i],
                                                                       __uint_as_float(acc[2 * i + 1]) + bf[2 * i + 1]));
            o0 = pr.x + a2.x;
            o1 = pr.y + a2.y;
          }
          __nv_bfloat162 po = __floats2bfloat162_rn(o0, o1);
          packed_out[i] = *reinterpret_cast<uint32_t*>(&po);
        }
      } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float u0 = __uint_as_float(acc[2 * i]) + bf[2 * i];
        const float u1 = __uint_as_float(acc[2 * i + 1]) + bf[2 * i + 1];
        __nv_bfloat162 pu = __floats2bfloat162_rn(u0, u1);
        packed_pre[i] = *reinterpret_cast<uint32_t*>(&pu);
        if (ACT == 1) {
          // the activation sees the bf16-rounded pre-activation, exactly like gelu(linear(x)) on bf16 tensors
          const float2 r = __bfloat1622float2(pu);
          __nv_bfloat162 po = __floats2bfloat162_rn(gelu_tanh(r.x), gelu_tanh(r.y));
          packed_out[i] = *reinterpret_cast<uint32_t*>(&po);
        } else {
          packed_out[i] = packed_pre[i];
        }
      }
      }
      if (row < M) {
        uint4* o = reinterpret_cast<uint4*>(out + row_off + c * 32);
#pragma unroll
        for (int v = 0; v < 4; ++v)
          o[v] = make_uint4(packed_out[4 * v], packed_out[4 * v + 1], packed_out[4 * v + 2], packed_out[4 * v + 3]);
        if (ACT < 2 && pre != nullptr) {
          uint4* p = reinterpret_cast<uint4*>(pre + row_off + c * 32);
#pragma unroll
          for (int v = 0; v < 4; ++v)
            p[v] = make_uint4(packed_pre[4 * v], packed_pre[4 * v + 1], packed_pre[4 * v + 2], packed_pre[4 * v + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_free<BN>(tmem_base);
}

// =====================================================================================================
// Variant 1 (persistent; validated on B200 in tests/test_gpu_tcgen05.py, 64-87 us at 8192x3072x768; selected explicitly):
//   grid = min(tiles, SMs); every CTA walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...
//   TMEM holds TWO accumulators (2 x BN columns): the MMA lane fills accumulator (i & 1) of its i-th tile
//   while the epilogue warps drain the other one -> bias/GELU/stores overlap the next tile's MMAs, and the
//   smem ring keeps streaming across tile boundaries (the producer never waits for an epilogue).
//   Extra barriers: acc_full[2] (MMA -> epilogue, tcgen05.commit) and acc_empty[2] (epilogue -> MMA, one
//   arrive per epilogue warp).
// =====================================================================================================
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// bias + activation + bf16 stores of one 32-column chunk of one accumulator row (held by one thread)
template <int ACT>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&acc)[32], const __nv_bfloat16* __restrict__ bias32,
                                               __nv_bfloat16* __restrict__ out32, __nv_bfloat16* __restrict__ pre32,
                                               bool row_valid) {
  float bf[32];
  if (bias32 != nullptr) {
    const uint4* bp = reinterpret_cast<const uint4*>(bias32);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint4 t = __ldg(bp + v);
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        bf[v * 8 + 2 * j] = f.x;
        bf[v * 8 + 2 * j + 1] = f.y;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) bf[i] = 0.f;
  }
  uint32_t packed_out[16], packed_pre[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float u0 = __uint_as_float(acc[2 * i]) + bf[2 * i];
    const float u1 = __uint_as_float(acc[2 * i + 1]) + bf[2 * i + 1];
    __nv_bfloat162 pu = __floats2bfloat162_rn(u0, u1);
    packed_pre[i] = *reinterpret_cast<uint32_t*>(&pu);
    if (ACT == 1) {
      const float2 r = __bfloat1622float2(pu);
      __nv_bfloat162 po = __floats2bfloat162_rn(gelu_tanh(r.x), gelu_tanh(r.y));
      packed_out[i] = *reinterpret_cast<uint32_t*>(&po);
    } else {
      packed_out[i] = packed_pre[i];
    }
  }
  if (row_valid) {
    uint4* o = reinterpret_cast<uint4*>(out32);
#pragma unroll
    for (int v = 0; v < 4; ++v)
      o[v] = make_uint4(packed_out[4 * v], packed_out[4 * v + 1], packed_out[4 * v + 2], packed_out[4 * v + 3]);
    if (pre32 != nullptr) {
      uint4* p = reinterpret_cast<uint4*>(pre32);
#pragma unroll
      for (int v = 0; v < 4; ++v)
        p[v] = make_uint4(packed_pre[4 * v], packed_pre[4 * v + 1], packed_pre[4 * v + 2], packed_pre[4 * v + 3]);
    }
  }
}

template <int BN>
struct SmemP {
  static constexpr uint32_t kABytes = kBM * kBK * 2;
  static constexpr uint32_t kBBytes = BN * kBK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kBarOffset = kStages * kStageBytes;   // full[S], empty[S], acc_full[2], acc_empty[2], tmem ptr
  static constexpr uint32_t kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16;
  static constexpr uint32_t kDynamic = kTotal + 1024;
};

template <int BN, int ACT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bias_act_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap map_a,
                                        const __grid_constant__ CUtensorMap map_w,
                                        const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                        __nv_bfloat16* __restrict__ pre, int M, int N, int K, int tiles_n,
                                        int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  using S = SmemP<BN>;
  constexpr int kTmemCols = 2 * BN;                              // 256 or 512: a power of two
  const uint32_t bar0 = base + S::kBarOffset;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto acc_full = [&](int a) { return bar0 + 8u * (2 * kStages + a); };
  auto acc_empty = [&](int a) { return bar0 + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * kStages + 4);
  auto smem_a = [&](int s) { return base + (uint32_t)s * S::kStageBytes; };
  auto smem_b = [&](int s) { return base + (uint32_t)s * S::kStageBytes + S::kABytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = K / kBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 4);                                // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: one continuous K-slab stream over all of this CTA's tiles =====
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = (int)(it % kStages);
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(empty(s), ph ^ 1u);
          mbar_expect_tx(full(s), S::kStageBytes);
          tma_load_2d(smem_a(s), &map_a, kb * kBK, m_blk * kBM, full(s));
          tma_load_2d(smem_b(s), &map_w, kb * kBK, n_blk * BN, full(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = instr_desc_bf16(kBM, BN);
      uint32_t it = 0, local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const uint32_t as = local & 1u, aph = (local >> 1) & 1u;
        mbar_wait(acc_empty(as), aph ^ 1u);                      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * (uint32_t)BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = (int)(it % kStages);
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(full(s), ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t da = smem_desc_k_sw128(smem_a(s) + (uint32_t)k * kUmmaK * 2);
            const uint64_t db = smem_desc_k_sw128(smem_b(s) + (uint32_t)k * kUmmaK * 2);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty(s));
        }
        umma_commit(acc_full(as));
      }
    }
  } else {
    // ===== epilogue warps 2..5 =====
    const int q = warp & 3;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      const uint32_t as = local & 1u, aph = (local >> 1) & 1u;
      mbar_wait(acc_full(as), aph);
      tc_fence_after();
      const int row = m_blk * kBM + q * 32 + lane;
      const size_t row_off = (size_t)row * (size_t)N + (size_t)n_blk * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)BN + (uint32_t)(c * 32), acc);
        epilogue_chunk<ACT>(acc, bias ? bias + (size_t)n_blk * BN + c * 32 : nullptr, out + row_off + c * 32,
                            pre ? pre + row_off + c * 32 : nullptr, row < M);
      }
      // all of this warp's TMEM reads of accumulator `as` are complete (wait::ld inside tmem_ld_32x32)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(as));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_free<kTmemCols>(tmem_base);
}

// =====================================================================================================
// Variant 2 (CTA pair, cta_group::2; validated on B200 in tests/test_gpu_tcgen05.py, 90-113 us at 8192x3072x768):
//   a cluster of two CTAs (two SMs of one TPC) owns a 256 x 256 output tile. CTA r stages ITS 128 rows of A and ITS
//   128 rows of W per K-slab (32 KB per stage instead of 48 KB), the leader (cluster rank 0) issues
//   tcgen05.mma.cta_group::2 with M = 256: the tensor cores of both SMs read both shared memories, CTA r's TMEM
//   receives rows 128 r .. 128 r + 127 of the accumulator. Per output element the pair pulls half as many operand
//   bytes through L2 as two independent 128 x 256 CTAs (arithmetic intensity 128 instead of 85 FLOP/B) — the
//   reason the library GEMMs it competes with are 2-CTA kernels.
//   Barriers: full[s] lives in the leader and counts the TMA bytes of BOTH CTAs (the peer's loads signal the
//   leader's barrier: mbarrier address with the CTA bit cleared); empty[s] and acc_full exist in both CTAs and are
//   signalled together by tcgen05.commit ... multicast::cluster with mask 0b11.
// =====================================================================================================
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;        // clears the CTA-rank bit of a shared::cluster address -> CTA 0

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem) {   // one whole warp in EACH CTA, same dst offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// this CTA's slab; completion bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, int c_inner, int c_outer,
                                                 uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar & kPeerBitMask), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  const unsigned short mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

struct SmemPair {
  static constexpr uint32_t kABytes = kBM * kBK * 2;             // this CTA's 128 rows of A
  static constexpr uint32_t kBBytes = 128 * kBK * 2;             // this CTA's 128 rows of W
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;     // 32 KB
  static constexpr int kStagesPair = 6;                          // 192 KB
  static constexpr uint32_t kBarOffset = kStagesPair * kStageBytes;
  static constexpr uint32_t kTotal = kBarOffset + (2 * kStagesPair + 1) * 8 + 16;
  static constexpr uint32_t kDynamic = kTotal + 1024;
};

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bias_act_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                                  const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                  __nv_bfloat16* __restrict__ pre, int M, int N, int K, int tiles_n) {
  constexpr int BN = 256;                                         // accumulator columns per CTA (full tile width)
  constexpr int kS = SmemPair::kStagesPair;
  extern __shared__ uint8_t smem_raw[];
  // both CTAs must end up with IDENTICAL offsets (descriptors and barrier addresses are shared): the dynamic smem
  // window starts at the same shared-space address in every CTA of a kernel, so the same rounding applies
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  using S = SmemPair;
  const uint32_t bar0 = base + S::kBarOffset;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kS + s); };
  const uint32_t acc_full = bar0 + 8u * (2 * kS);
  const uint32_t tmem_slot = acc_full + 8u;
  auto smem_a = [&](int s) { return base + (uint32_t)s * S::kStageBytes; };
  auto smem_b = [&](int s) { return base + (uint32_t)s * S::kStageBytes + S::kABytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();                         // 0 = leader
  const int pair = blockIdx.x >> 1;
  const int m_blk = pair / tiles_n, n_blk = pair % tiles_n;       // 256 x 256 tile of the pair
  const int row0 = m_blk * 256 + (int)cta * 128;                  // this CTA's rows of A / of the output
  const int wrow0 = n_blk * 256 + (int)cta * 128;                 // this CTA's rows of W (columns of the output tile)
  const int num_kb = K / kBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < kS; ++s) {
      mbar_init(full(s), 1);                                      // only the leader's copy is ever used
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_2cta<BN>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();                                             // the peer's barriers exist before anyone signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs; bytes are counted on the leader's full[s]) =====
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kS;
        const uint32_t ph = (uint32_t)(kb / kS) & 1u;
        mbar_wait(empty(s), ph ^ 1u);                             // my own copy: commit arrives on both CTAs
        if (cta == 0) mbar_expect_tx(full(s), 2 * S::kStageBytes);
        tma_load_2d_2cta(smem_a(s), &map_a, kb * kBK, row0, full(s));
        tma_load_2d_2cta(smem_b(s), &map_w, kb * kBK, wrow0, full(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta == 0) {
      // ===== MMA issuer: leader only =====
      constexpr uint32_t idesc = instr_desc_bf16(256, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kS;
        const uint32_t ph = (uint32_t)(kb / kS) & 1u;
        mbar_wait(full(s), ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          const uint64_t da = smem_desc_k_sw128(smem_a(s) + (uint32_t)k * kUmmaK * 2);
          const uint64_t db = smem_desc_k_sw128(smem_b(s) + (uint32_t)k * kUmmaK * 2);
          umma_bf16_2cta(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit_2cta(empty(s));
      }
      umma_commit_2cta(acc_full);
    }
  } else {
    // ===== epilogue (both CTAs, each drains its own 128 accumulator rows) =====
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int row = row0 + q * 32 + lane;
    const size_t row_off = (size_t)row * (size_t)N + (size_t)n_blk * BN;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t acc[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), acc);
      epilogue_chunk<ACT>(acc, bias ? bias + (size_t)n_blk * BN + c * 32 : nullptr, out + row_off + c * 32,
                          pre ? pre + row_off + c * 32 : nullptr, row < M);
    }
  }
  tc_fence_before();
  cluster_sync_all();                                             // both CTAs are done with TMEM and with each other's smem
  if (warp == 1) tmem_free_2cta<BN>(tmem_base);
}

// ---- host ----------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  });
  return fn;
}

// row-major [rows, cols] bf16 matrix, box = (64 columns, box_rows rows), 128-byte swizzle
static int make_map(CUtensorMap* map, const void* ptr, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { set_error("gemm_tcgen05: cuTensorMapEncodeTiled unavailable"); return -1; }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_tcgen05: cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
  return 0;
}

template <int BN, int ACT>
static int launch(const CUtensorMap& ma, const CUtensorMap& mw, const void* bias, void* out, void* pre, int M, int N,
                  int K, cudaStream_t s) {
  auto kern = gemm_bias_act_tcgen05_kernel<BN, ACT>;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<BN>::kDynamic));
    configured = true;
  }
  dim3 grid(N / BN, (M + kBM - 1) / kBM);
  kern<<<grid, kThreads, Smem<BN>::kDynamic, s>>>(ma, mw, (const __nv_bfloat16*)bias, (__nv_bfloat16*)out,
                                                   (__nv_bfloat16*)pre, M, N, K);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

template <int BN, int ACT>
static int launch_persistent(const CUtensorMap& ma, const CUtensorMap& mw, const void* bias, void* out, void* pre,
                             int M, int N, int K, cudaStream_t s) {
  auto kern = gemm_bias_act_tcgen05_persistent_kernel<BN, ACT>;
  static bool configured = false;
  static int sms = 0;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemP<BN>::kDynamic));
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  const int tiles_n = N / BN, tiles_m = (M + kBM - 1) / kBM;
  const int num_tiles = tiles_n * tiles_m;
  const int grid = num_tiles < sms ? num_tiles : sms;
  kern<<<grid, kThreads, SmemP<BN>::kDynamic, s>>>(ma, mw, (const __nv_bfloat16*)bias, (__nv_bfloat16*)out,
                                                   (__nv_bfloat16*)pre, M, N, K, tiles_n, num_tiles);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

template <int ACT>
static int launch_pair(const CUtensorMap& ma, const CUtensorMap& mw, const void* bias, void* out, void* pre, int M,
                       int N, int K, cudaStream_t s) {
  auto kern = gemm_bias_act_tcgen05_pair_kernel<ACT>;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemPair::kDynamic));
    configured = true;
  }
  const int tiles_n = N / 256, tiles_m = (M + 255) / 256;
  kern<<<2 * tiles_n * tiles_m, kThreads, SmemPair::kDynamic, s>>>(ma, mw, (const __nv_bfloat16*)bias,
                                                                   (__nv_bfloat16*)out, (__nv_bfloat16*)pre, M, N, K,
                                                                   tiles_n);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tc
}  // namespace adapcc

using namespace adapcc;

extern "C" {

// out = act(a[M,K] @ w[N,K]^T + bias); pre (optional) = the pre-activation. bf16 row-major, 16-byte aligned.
// act 2: out = (a @ w^T) * gelu'(aux); act 3: out = a @ w^T + bias + aux — aux is passed in `pre` (an input then).
// Constraints of this first version: K % 64 == 0, N % 128 == 0 (256-wide tiles when N % 256 == 0).
// variant 0: one tile per CTA (validated on B200). variant 1: persistent CTAs, double-buffered TMEM accumulator;
// variant 2: CTA pairs (cta_group::2, 256 x 256 tile per pair) — all three validated on B200 (tests/test_gpu_tcgen05.py).
int adapcc_gemm_bias_act_v(const void* a, const void* w, const void* bias, void* out, void* pre, int M, int N, int K,
                           int act, int variant, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (K % tc::kBK != 0 || N % 128 != 0) { set_error("gemm_tcgen05: need K %% 64 == 0 and N %% 128 == 0 (got K=%d N=%d)", K, N); return -1; }
  if (act < 0 || act > 3) { set_error("gemm_tcgen05: act must be 0 (none), 1 (gelu_tanh), 2 (dgelu * aux) or 3 (+ aux)"); return -1; }
  if (act >= 2 && pre == nullptr) { set_error("gemm_tcgen05: act %d needs the aux tensor (passed as `pre`)", act); return -1; }
  if (act >= 2 && variant != 0) { set_error("gemm_tcgen05: act %d is only built for variant 0", act); return -1; }
  if (variant < 0 || variant > 2) { set_error("gemm_tcgen05: variant must be 0, 1 or 2"); return -1; }
  if (variant == 2 && (N % 256 != 0 || act > 1)) { set_error("gemm_tcgen05: variant 2 (CTA pairs) needs N %% 256 == 0 and act 0/1"); return -1; }
  if (((uintptr_t)a | (uintptr_t)w | (uintptr_t)out | (uintptr_t)pre | (uintptr_t)bias) & 15) { set_error("gemm_tcgen05: operands must be 16-byte aligned"); return -1; }
  const int bn = (N % 256 == 0) ? 256 : 128;
  CUtensorMap ma, mw;
  if (tc::make_map(&ma, a, M, K, tc::kBM)) return -1;
  if (tc::make_map(&mw, w, N, K, bn)) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  if (variant == 2) {
    CUtensorMap mw2;                                   // each CTA of a pair loads 128 rows of W per K-slab
    if (tc::make_map(&mw2, w, N, K, 128)) return -1;
    return act ? tc::launch_pair<1>(ma, mw2, bias, out, pre, M, N, K, s) : tc::launch_pair<0>(ma, mw2, bias, out, pre, M, N, K, s);
  }
  if (variant == 1) {
    if (bn == 256) return act ? tc::launch_persistent<256, 1>(ma, mw, bias, out, pre, M, N, K, s) : tc::launch_persistent<256, 0>(ma, mw, bias, out, pre, M, N, K, s);
    return act ? tc::launch_persistent<128, 1>(ma, mw, bias, out, pre, M, N, K, s) : tc::launch_persistent<128, 0>(ma, mw, bias, out, pre, M, N, K, s);
  }
#define TC_LAUNCH(BN_)                                                                  \
  switch (act) {                                                                        \
    case 0: return tc::launch<BN_, 0>(ma, mw, bias, out, pre, M, N, K, s);              \
    case 1: return tc::launch<BN_, 1>(ma, mw, bias, out, pre, M, N, K, s);              \
    case 2: return tc::launch<BN_, 2>(ma, mw, bias, out, pre, M, N, K, s);              \
    default: return tc::launch<BN_, 3>(ma, mw, bias, out, pre, M, N, K, s);             \
  }
  if (bn == 256) { TC_LAUNCH(256) }
  TC_LAUNCH(128)
#undef TC_LAUNCH
}

int adapcc_gemm_bias_act(const void* a, const void* w, const void* bias, void* out, void* pre, int M, int N, int K,
                         int act, void* stream) {
  return adapcc_gemm_bias_act_v(a, w, bias, out, pre, M, N, K, act, 0, stream);
}

}  // extern "C"
