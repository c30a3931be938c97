// tcgen05 GEMM, variant 3: PERSISTENT CTA PAIRS with a double-buffered TMEM accumulator and a coalescing epilogue.
//
//   out[M, N] = act(A[M, K] . W[N, K]^T + bias[N])      bf16 in / out, fp32 accumulation in TMEM
//   ACT 0: identity        1: GELU(tanh), optionally also pre = A.W^T + bias (for the backward)
//       2: dGELU: out = (A.W^T) * gelu'(aux)  — the MLP backward dH = (dY.W2) (.) gelu'(pre) — and, optionally,
//          colsum[N] += column sums of `out` (= the bias gradient of the up-projection, accumulated in fp32)
//
// What the measurements of variants 0-2 said (gpurun call 1 of round 2, 8192 x 3072 x 768, B200):
//   * one 128 x 256 tile per CTA streams 453 MB of operands out of L2 per call — 64 us at the ~7 TB/s the L2 delivers,
//     exactly what the persistent variant 1 measured: operand REUSE, not issue rate, is the limit -> CTA pairs
//     (tcgen05.mma.cta_group::2, 256 x 256 per pair: 302 MB);
//   * writing `pre` cost +23 us: every thread stored 64-byte row fragments 6 KB apart (32 half-filled lines per
//     warp instruction) -> rows are transposed through padded shared memory and leave as 128-byte segments;
//   * four epilogue warps (one per scheduler) could not hide the TMEM-load / SFU / store latencies -> eight, and GELU
//     costs one MUFU (tanh.approx) instead of two (ex2 + rcp).
//
// (The epilogue stages 32 columns at a time through ONE padded tile per warp, which leaves room for six 32 KB operand
// stages: with four, the MMA issuer waited on `full[s]` half the time — ncu: tensor pipe 51 %, epilogue idle on acc_full.)
//
// Roles in each CTA of a pair (320 threads): warp 0 lane 0 = TMA producer (its 128 rows of A and of W per 64-wide
// K slab, completion counted on the LEADER's `full` barrier); warp 1 = TMEM allocation and, in the leader, lane 0 = the
// MMA issuer (4 x tcgen05.mma.cta_group::2 M256 x N256 x K16 per slab; tcgen05.commit multicast releases the slab in
// both CTAs and finally signals `acc_full`); warps 2..9 = epilogue (TMEM lane quarter = warp % 4, column half =
// (warp - 2) / 4) draining accumulator (tile & 1) while the tensor cores fill the other one. `acc_empty` lives in the
// leader and counts the 16 epilogue warps of BOTH CTAs (the peer arrives remotely).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <mutex>

#include "common.h"

namespace adapcc {
namespace tc3 {

constexpr int kBK = 64;            // K slab: 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kStages = 6;          // 6 x 32 KB per CTA: the refill round trip (MMA done -> commit -> producer wake -> TMA ->
                                   // complete_tx -> issuer wake, ~3300 cycles measured) must fit in (stages - 1) x 512 MMA cycles
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kTileM = 256, kTileN = 256;       // per pair
constexpr uint32_t kABytes = 128 * kBK * 2, kBBytes = 128 * kBK * 2, kStageBytes = kABytes + kBBytes;   // per CTA
constexpr uint32_t kRowPitch = 80;              // staging row: 32 bf16 (64 B) + 16 B pad: row writes conflict-free
constexpr uint32_t kStageTile = 32 * kRowPitch; // one warp's 32 x 32 chunk; ONE tile per warp, reused for out / pre / aux
constexpr uint32_t kEpiBytes = kEpiWarps * kStageTile;
constexpr uint32_t kBiasBytes = kEpiWarps * 128 * 4;            // each epilogue warp's 128 bias values of the tile, as fp32
constexpr uint32_t kBarOffset = kStages * kStageBytes + kEpiBytes + kBiasBytes;
// full[S] empty[S] acc_full[2] acc_empty[2] + tmem slot
constexpr uint32_t kSmemTotal = kBarOffset + (2 * kStages + 4) * 8 + 16;
constexpr uint32_t kSmemDynamic = kSmemTotal + 1024;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in CTA 0

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      // default (acquire.cta) semantics on purpose: a cluster-scope acquire makes ptxas emit CCTL.IVALL (an L1
      // invalidate) after EVERY successful wait — 17 % of all stall samples in the first ncu capture of this kernel, most
      // of them in the MMA issuer's per-slab wait (tensor pipe 50 % busy). Nothing read after these waits needs it: the
      // operands arrive through the async proxy (complete_tx), and acc_empty only orders TMEM reuse.
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {      // wall-clock bounded: trap, never hang
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
// arrive on the barrier at `local_bar`'s offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
  // relaxed: the only thing ordered by this arrive is "my tcgen05.ld of the accumulator have completed", which
  // tcgen05.wait::ld + fence::before_thread_sync already guarantee; a release would also drain the tile's global stores
  // (ncu: 12 % of the epilogue's issue slots stalled on membar)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, int c_inner, int c_outer,
                                                 uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar & kPeerBitMask), "r"(c_inner), "r"(c_outer) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  const unsigned short mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr) {     // K-major, 128-byte swizzle (sm_100)
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t instr_desc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 0.5 x (1 + tanh(k (x + c x^3)))
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);      // k x + k c x^3
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_fast(u), hx);
}
// d/dx: 0.5 (1 + t) + 0.5 x (1 - t^2) k (1 + 3 c x^2)
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float x2 = x * x;
  const float t = tanh_fast(x * fmaf(0.0356774081f, x2, 0.7978845608f));
  const float du = fmaf(0.1070322243f, x2, 0.7978845608f);             // k (1 + 3 c x^2)
  return fmaf(0.5f * x * (1.f - t * t), du, 0.5f + 0.5f * t);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

// One warp moves its staged 32 x 32 bf16 chunk to / from global memory, 8 rows (8 x 64 B) per instruction.
__device__ __forceinline__ void chunk_to_global(uint32_t stage, __nv_bfloat16* __restrict__ g, size_t ld, int row0,
                                                int rows_valid, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2), ch = lane & 3;
    const uint4 v = ld_shared_v4(stage + (uint32_t)r * kRowPitch + (uint32_t)ch * 16);
    if (r < rows_valid) *reinterpret_cast<uint4*>(g + (size_t)(row0 + r) * ld + ch * 8) = v;
  }
}
__device__ __forceinline__ void chunk_from_global(uint32_t stage, const __nv_bfloat16* __restrict__ g, size_t ld,
                                                  int row0, int rows_valid, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2), ch = lane & 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid) v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)(row0 + r) * ld + ch * 8));
    st_shared_v4(stage + (uint32_t)r * kRowPitch + (uint32_t)ch * 16, v);
  }
}

template <int ACT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_pair_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                            const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                            __nv_bfloat16* __restrict__ pre, float* __restrict__ colsum, int M, int N, int K,
                            int tiles_n, int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // identical in both CTAs of the pair
  const uint32_t bar0 = base + kBarOffset;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto acc_full = [&](int a) { return bar0 + 8u * (2 * kStages + a); };
  auto acc_empty = [&](int a) { return bar0 + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * kStages + 4);
  auto smem_a = [&](int s) { return base + (uint32_t)s * kStageBytes; };
  auto smem_b = [&](int s) { return base + (uint32_t)s * kStageBytes + kABytes; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
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
      mbar_init(acc_empty(a), 2 * kEpiWarps);                       // the epilogue warps of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: one continuous slab stream over all tiles of this pair =====
      uint32_t it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        const int row0 = m_blk * kTileM + (int)cta * 128, wrow0 = n_blk * kTileN + (int)cta * 128;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = (int)(it % kStages);
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(empty(s), ph ^ 1u);
          if (cta == 0) mbar_expect_tx(full(s), 2 * kStageBytes);
          tma_load_2d_2cta(smem_a(s), &map_a, kb * kBK, row0, full(s));
          tma_load_2d_2cta(smem_b(s), &map_w, kb * kBK, wrow0, full(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta == 0) {
      // ===== MMA issuer (leader) =====
      constexpr uint32_t idesc = instr_desc_bf16(256, kTileN);
      uint32_t it = 0, local = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
        const uint32_t as = local & 1u, use = local >> 1;
        mbar_wait(acc_empty(as), (use & 1u) ^ 1u);                  // both CTAs drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * (uint32_t)kTileN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = (int)(it % kStages);
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(full(s), ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t da = smem_desc_k_sw128(smem_a(s) + (uint32_t)k * kUmmaK * 2);
            const uint64_t db = smem_desc_k_sw128(smem_b(s) + (uint32_t)k * kUmmaK * 2);
            umma_bf16_2cta(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta(empty(s));
        }
        umma_commit_2cta(acc_full(as));
      }
    }
  } else {
    // ===== epilogue: 8 warps, TMEM lane quarter q, column half h =====
    const int e = warp - 2, q = warp & 3, half = e >> 2;
    const uint32_t st = base + kStages * kStageBytes + (uint32_t)e * kStageTile;     // this warp's 32 x 32 staging tile
    const uint32_t my_row = st + (uint32_t)lane * kRowPitch;
    const uint32_t st_bias = base + kStages * kStageBytes + kEpiBytes + (uint32_t)e * 512;
    uint32_t local = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++local) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      const uint32_t as = local & 1u, use = local >> 1;
      const int row0 = m_blk * kTileM + (int)cta * 128 + q * 32;    // this warp's 32 rows; lane = row
      const int rows_valid = min(32, M - row0);
      if (ACT != 2) {
        // this warp's 128 bias values of the tile -> fp32 in shared memory once (every lane needs all of them for its
        // row: broadcast LDS.128 below instead of 16-byte global loads + bf16 unpacking per 32 columns)
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias != nullptr) {
          const uint2 t = __ldg(reinterpret_cast<const uint2*>(bias + n_blk * kTileN + half * 128 + lane * 4));
          const float2 lo = unpack_bf16x2(t.x), hi = unpack_bf16x2(t.y);
          bv = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(st_bias + (uint32_t)lane * 16), "f"(bv.x), "f"(bv.y),
                     "f"(bv.z), "f"(bv.w) : "memory");
        __syncwarp();
      }
      mbar_wait(acc_full(as), use & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {                              // 32-column groups of this warp's 128-column half
        const int col0 = n_blk * kTileN + half * 128 + cc * 32;
        uint32_t ax[16];
        if (ACT == 2) {
          // aux (saved pre-activation): coalesced global -> staging, then every lane takes its own row
          chunk_from_global(st, pre + col0, (size_t)N, row0, rows_valid, lane);
          __syncwarp();
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 t = ld_shared_v4(my_row + (uint32_t)v * 16);
            ax[4 * v] = t.x; ax[4 * v + 1] = t.y; ax[4 * v + 2] = t.z; ax[4 * v + 3] = t.w;
          }
          __syncwarp();                                             // the staging tile is about to hold `out`
        }
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)kTileN + (uint32_t)(half * 128 + cc * 32), acc);
        uint32_t po[16], pp[16];
        if (ACT == 2) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 a2 = unpack_bf16x2(ax[i]);
            po[i] = pack_bf16x2(__uint_as_float(acc[2 * i]) * gelu_tanh_grad(a2.x),
                                __uint_as_float(acc[2 * i + 1]) * gelu_tanh_grad(a2.y));
          }
        } else {
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            float4 b4;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b4.x), "=f"(b4.y), "=f"(b4.z), "=f"(b4.w)
                         : "r"(st_bias + (uint32_t)(cc * 32 + v * 4) * 4) : "memory");
            const float r0 = __uint_as_float(acc[4 * v]) + b4.x, r1 = __uint_as_float(acc[4 * v + 1]) + b4.y;
            const float r2 = __uint_as_float(acc[4 * v + 2]) + b4.z, r3 = __uint_as_float(acc[4 * v + 3]) + b4.w;
            pp[2 * v] = pack_bf16x2(r0, r1);
            pp[2 * v + 1] = pack_bf16x2(r2, r3);
            if (ACT == 1) {         // GELU on the fp32 pre-activation (torch applies it to the bf16-rounded one: the two
                                    // differ by less than the output's own bf16 rounding)
              po[2 * v] = pack_bf16x2(gelu_tanh(r0), gelu_tanh(r1));
              po[2 * v + 1] = pack_bf16x2(gelu_tanh(r2), gelu_tanh(r3));
            } else {
              po[2 * v] = pp[2 * v];
              po[2 * v + 1] = pp[2 * v + 1];
            }
          }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v)
          st_shared_v4(my_row + (uint32_t)v * 16, make_uint4(po[4 * v], po[4 * v + 1], po[4 * v + 2], po[4 * v + 3]));
        __syncwarp();
        chunk_to_global(st, out + col0, (size_t)N, row0, rows_valid, lane);
        if (ACT == 2 && colsum != nullptr) {
          // bias gradient: lane l sums column l of the staged bf16 chunk over the valid rows
          float s0 = 0.f;
          for (int r = 0; r < rows_valid; ++r) {
            unsigned short w;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(w) : "r"(st + (uint32_t)r * kRowPitch + (uint32_t)lane * 2));
            s0 += __uint_as_float((uint32_t)w << 16);
          }
          atomicAdd(colsum + col0 + lane, s0);
        }
        __syncwarp();                                               // everybody has read the tile: reuse it
        if (ACT <= 1 && pre != nullptr) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            st_shared_v4(my_row + (uint32_t)v * 16, make_uint4(pp[4 * v], pp[4 * v + 1], pp[4 * v + 2], pp[4 * v + 3]));
          __syncwarp();
          chunk_to_global(st, pre + col0, (size_t)N, row0, rows_valid, lane);
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty(as), 0);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

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
static int make_map(CUtensorMap* map, const void* ptr, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { set_error("gemm_tcgen05_pp: cuTensorMapEncodeTiled unavailable"); return -1; }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_tcgen05_pp: cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
  return 0;
}

template <int ACT>
static int launch(const CUtensorMap& ma, const CUtensorMap& mw, const void* bias, void* out, void* pre, float* colsum,
                  int M, int N, int K, cudaStream_t s) {
  auto kern = gemm_pair_persistent_kernel<ACT>;
  static bool configured = false;
  static int sms = 0;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemDynamic));
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  const int tiles_n = N / kTileN, tiles_m = (M + kTileM - 1) / kTileM;
  const int num_tiles = tiles_n * tiles_m;
  const int pairs = std::min(num_tiles, sms / 2);
  kern<<<2 * pairs, kThreads, kSmemDynamic, s>>>(ma, mw, (const __nv_bfloat16*)bias, (__nv_bfloat16*)out,
                                                 (__nv_bfloat16*)pre, colsum, M, N, K, tiles_n, num_tiles);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace tc3
}  // namespace adapcc

using namespace adapcc;

// act 0 / 1: out = act(a @ w^T + bias), pre (optional output) = a @ w^T + bias.
// act 2: out = (a @ w^T) * gelu'(pre)  (pre = aux INPUT), colsum (optional, fp32 [N], accumulated into) += column sums.
// bf16 row-major operands, 16-byte aligned; K % 64 == 0, N % 256 == 0.
extern "C" int adapcc_gemm_pp(const void* a, const void* w, const void* bias, void* out, void* pre, float* colsum,
                              int M, int N, int K, int act, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (K % tc3::kBK != 0 || N % tc3::kTileN != 0) { set_error("gemm_pp: need K %% 64 == 0 and N %% 256 == 0 (got K=%d N=%d)", K, N); return -1; }
  if (act < 0 || act > 2) { set_error("gemm_pp: act must be 0 (none), 1 (gelu) or 2 (dgelu)"); return -1; }
  if (act == 2 && pre == nullptr) { set_error("gemm_pp: act 2 needs the saved pre-activation"); return -1; }
  if (((uintptr_t)a | (uintptr_t)w | (uintptr_t)out | (uintptr_t)pre | (uintptr_t)bias) & 15) { set_error("gemm_pp: operands must be 16-byte aligned"); return -1; }
  CUtensorMap ma, mw;
  if (tc3::make_map(&ma, a, M, K, 128)) return -1;
  if (tc3::make_map(&mw, w, N, K, 128)) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  switch (act) {
    case 0: return tc3::launch<0>(ma, mw, bias, out, pre, colsum, M, N, K, s);
    case 1: return tc3::launch<1>(ma, mw, bias, out, pre, colsum, M, N, K, s);
    default: return tc3::launch<2>(ma, mw, bias, out, pre, colsum, M, N, K, s);
  }
}
