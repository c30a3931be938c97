// Direct (switch-topology) collectives on symmetric memory: one-shot, two-shot and NVLS
// all-reduce / reduce, plus direct broadcast. These are the "uniform NVSwitch" schedules
// the synthesizer can pick besides the reference-style trees. Every kernel fuses what the
// reference runs as separate engine ops with host syncs in between
// (/root/reference/csrc/allreduce.cu:568-654, /root/reference/csrc/run.cu:103-127):
//   stage-in  : user tensor -> symmetric window, fused dtype cast (fp32 -> bf16 wire)
//   transfer  : in-kernel peer ld/st (or multimem) over NVLink, fused reduction in fp32
//   stage-out : symmetric window -> user tensor, fused 1/N scale + cast back
// and a "zero-copy" mode when the tensor already lives in the symmetric heap.
#pragma once
#include "device_prims.cuh"

namespace adapcc {

constexpr int kThreads = 512;
constexpr int kUnroll = 4;

enum DirectFlags : int {
  F_ZERO_COPY = 1,   // tensor is inside the symmetric window (data[] already point at it)
  F_ROOT_ONLY = 2,   // reduce-to-root: only active index `root` receives the result
};

// Work partition: packs [0, npacks) are cut into `nslices` contiguous slices of `pps`
// packs; inside a slice, pack j belongs to block (j / kThreads) % gridDim.x in EVERY
// phase, so block b only ever consumes data that block b of a peer produced and a
// per-block barrier is sufficient.
struct Partition {
  long long npacks, pps;
  int nslices;
  long long pack0 = 0;     // first pack of the piece this partition covers (0 = whole message)
  __device__ __forceinline__ long long slice_begin(int s) const { return pack0 + (long long)s * pps; }
  __device__ __forceinline__ long long slice_count(int s) const {
    long long b = (long long)s * pps;
    long long c = npacks - b;
    return c < 0 ? 0 : (c > pps ? pps : c);
  }
};

// Every helper below takes a sub-grid view (bid of nb CTAs); the default is the whole grid. The
// pipelined staged kernel (kernels_pipelined.cuh) runs stagers and link CTAs as two sub-grids.
struct SubGrid {
  int bid, nb;
  __device__ __forceinline__ SubGrid() : bid(blockIdx.x), nb(gridDim.x) {}
  __device__ __forceinline__ SubGrid(int b, int n) : bid(b), nb(n) {}
};

template <typename U, typename W>
__device__ __forceinline__ void stage_in(const Partition& P, const U* __restrict__ in, long long n,
                                         bool vec_ok, char* __restrict__ local, SubGrid g = SubGrid()) {
  constexpr int kEpp = WireTraits<W>::kEpp;
  const long long stride = (long long)g.nb * kThreads;
  for (int s = 0; s < P.nslices; ++s) {
    const long long base = P.slice_begin(s), cnt = P.slice_count(s);
    for (long long j0 = (long long)g.bid * kThreads + threadIdx.x; j0 < cnt; j0 += stride * kUnroll) {
      float f[kUnroll][kEpp];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = j0 + u * stride;
        if (j < cnt) load_user<U, kEpp>(in, (base + j) * kEpp, n, vec_ok, f[u]);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = j0 + u * stride;
        if (j < cnt) st16(local + (base + j) * 16, pack<W>(f[u]));
      }
    }
  }
}

template <typename U, typename W>
__device__ __forceinline__ void stage_out(const Partition& P, U* __restrict__ out, long long n,
                                          bool vec_ok, const char* __restrict__ local, float scale,
                                          SubGrid g = SubGrid()) {
  constexpr int kEpp = WireTraits<W>::kEpp;
  const long long stride = (long long)g.nb * kThreads;
  for (int s = 0; s < P.nslices; ++s) {
    const long long base = P.slice_begin(s), cnt = P.slice_count(s);
    for (long long j0 = (long long)g.bid * kThreads + threadIdx.x; j0 < cnt; j0 += stride * kUnroll) {
      uint4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = j0 + u * stride;
        if (j < cnt) v[u] = ld16(local + (base + j) * 16);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = j0 + u * stride;
        if (j < cnt) {
          float f[kEpp];
          unpack<W>(v[u], f);
#pragma unroll
          for (int i = 0; i < kEpp; ++i) f[i] *= scale;
          store_user<U, kEpp>(out, (base + j) * kEpp, n, vec_ok, f);
        }
      }
    }
  }
}

// Zero-copy tensors need not be a whole number of 16-byte packs (a slice of a heap tensor, a DDP bucket with an
// odd element count). The last pack is then PARTIAL: it is loaded as 16 bytes (the bytes past the tensor are inside
// the symmetric heap, so the read is safe and its lanes are simply dropped) but must be stored element by element,
// or every peer's memory right after the tensor would be overwritten with reduced neighbour data.
struct TailPack {
  long long pack;    // global index of the partial last pack, -1 when the message is a whole number of packs
  int elems;         // valid wire elements in it
};
template <typename W>
__device__ __forceinline__ TailPack tail_of(long long n, bool zero_copy) {
  constexpr int kEpp = WireTraits<W>::kEpp;
  TailPack t;
  t.elems = (int)(n % kEpp);
  t.pack = (zero_copy && t.elems) ? n / kEpp : -1;
  return t;
}
template <typename W>
__device__ __forceinline__ void st_tail(char* dst, const float* f, int elems) {
  W* d = reinterpret_cast<W*>(dst);
  for (int i = 0; i < elems; ++i) d[i] = from_float<W>(f[i]);
}

// ----------------------------------------------------------------------------------
// phase-1 bodies, specialised on NR = upper bound of participants (2/4/8/16) so that
// NR x UN = 16 independent 128-bit peer loads are in flight per thread: NVLink round trips
// are ~2 us, bandwidth comes from bytes in flight, not from thread count.
// ----------------------------------------------------------------------------------
template <typename U, typename W, int OP, int NR>
__device__ __forceinline__ void one_shot_phase1(const DevComm& c, long long npacks, U* __restrict__ out,
                                                long long n, bool out_vec, float scale) {
  constexpr int kEpp = WireTraits<W>::kEpp;
  constexpr int UN = 16 / NR;
  const int na = c.n_active;
  const long long stride = (long long)gridDim.x * kThreads;
  const char* peers[NR];
#pragma unroll
  for (int a = 0; a < NR; ++a) peers[a] = a < na ? c.data[c.active_ranks[a]] : nullptr;
  for (long long j0 = (long long)blockIdx.x * kThreads + threadIdx.x; j0 < npacks; j0 += stride * UN) {
    uint4 v[UN][NR];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long j = j0 + u * stride;
      if (j < npacks) {
#pragma unroll
        for (int a = 0; a < NR; ++a)
          if (a < na) v[u][a] = ld16(peers[a] + j * 16);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long j = j0 + u * stride;
      if (j < npacks) {
        float acc[kEpp];
#pragma unroll
        for (int i = 0; i < kEpp; ++i) acc[i] = red_identity<OP>();
#pragma unroll
        for (int a = 0; a < NR; ++a)   // fixed rank order: bit-identical on every replica
          if (a < na) {
            float f[kEpp];
            unpack<W>(v[u][a], f);
#pragma unroll
            for (int i = 0; i < kEpp; ++i) acc[i] = red_apply<OP>(acc[i], f[i]);
          }
#pragma unroll
        for (int i = 0; i < kEpp; ++i) acc[i] *= scale;
        store_user<U, kEpp>(out, j * kEpp, n, out_vec, acc);
      }
    }
  }
}

template <typename W, int OP, int NR>
__device__ __forceinline__ void two_shot_phase1(const DevComm& c, long long base, long long cnt,
                                                bool zero_copy, float scale, bool root_only, int root,
                                                SubGrid g = SubGrid(), TailPack tail = TailPack{-1, 0}) {
  constexpr int kEpp = WireTraits<W>::kEpp;
  constexpr int UN = 16 / NR;
  const int na = c.n_active, me = c.my_index;
  const long long stride = (long long)g.nb * kThreads;
  char* peers[NR];
#pragma unroll
  for (int a = 0; a < NR; ++a) {
    int idx = me + a;                       // stagger: rank r starts with its own window
    if (idx >= na) idx -= na;
    peers[a] = a < na ? c.data[c.active_ranks[idx]] : nullptr;
  }
  char* const root_ptr = root_only ? c.data[c.active_ranks[root]] : nullptr;
  for (long long j0 = (long long)g.bid * kThreads + threadIdx.x; j0 < cnt; j0 += stride * UN) {
    uint4 v[UN][NR];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long j = j0 + u * stride;
      if (j < cnt) {
#pragma unroll
        for (int a = 0; a < NR; ++a)
          if (a < na) v[u][a] = ld16(peers[a] + (base + j) * 16);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long j = j0 + u * stride;
      if (j < cnt) {
        const long long off = (base + j) * 16;
        float acc[kEpp];
#pragma unroll
        for (int i = 0; i < kEpp; ++i) acc[i] = red_identity<OP>();
#pragma unroll
        for (int a = 0; a < NR; ++a)
          if (a < na) {
            float f[kEpp];
            unpack<W>(v[u][a], f);
#pragma unroll
            for (int i = 0; i < kEpp; ++i) acc[i] = red_apply<OP>(acc[i], f[i]);
          }
        if (zero_copy) {
#pragma unroll
          for (int i = 0; i < kEpp; ++i) acc[i] *= scale;
        }
        const uint4 r = pack<W>(acc);
        if (base + j == tail.pack) {           // partial last pack of a zero-copy tensor: bounded stores
          if (root_only) {
            st_tail<W>(root_ptr + off, acc, tail.elems);
          } else {
            for (int a = 0; a < NR; ++a)
              if (a < na) st_tail<W>(peers[a] + off, acc, tail.elems);
          }
        } else if (root_only) {
          st16(root_ptr + off, r);
        } else {
#pragma unroll
          for (int a = 0; a < NR; ++a)
            if (a < na) st16(peers[a] + off, r);
        }
      }
    }
  }
}

template <typename W, int OP>
__device__ __forceinline__ void nvls_phase1(const DevComm& c, long long base, long long cnt, bool zero_copy,
                                            float scale, bool root_only, int root, SubGrid g = SubGrid(),
                                            TailPack tail = TailPack{-1, 0}) {
  constexpr int kEpp = WireTraits<W>::kEpp;
  constexpr int UN = 8;
  const long long stride = (long long)g.nb * kThreads;
  char* const root_ptr = root_only ? c.data[c.active_ranks[root]] : nullptr;
  for (long long j0 = (long long)g.bid * kThreads + threadIdx.x; j0 < cnt; j0 += stride * UN) {
    uint4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long j = j0 + u * stride;
      if (j < cnt) v[u] = mc_ld_reduce<W, OP>(c.mc_data + (base + j) * 16);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long j = j0 + u * stride;
      if (j < cnt) {
        if (zero_copy && scale != 1.f) {
          float f[kEpp];
          unpack<W>(v[u], f);
#pragma unroll
          for (int i = 0; i < kEpp; ++i) f[i] *= scale;
          v[u] = pack<W>(f);
        }
        if (base + j == tail.pack) {           // partial last pack: unicast, element by element
          float f[kEpp];
          unpack<W>(v[u], f);
          if (root_only) {
            st_tail<W>(root_ptr + (base + j) * 16, f, tail.elems);
          } else {
            for (int a = 0; a < c.n_active; ++a) st_tail<W>(c.data[c.active_ranks[a]] + (base + j) * 16, f, tail.elems);
          }
        } else if (root_only) st16(root_ptr + (base + j) * 16, v[u]);
        else mc_st16(c.mc_data + (base + j) * 16, v[u]);
      }
    }
  }
}

// ----------------------------------------------------------------------------------
// all-reduce / reduce, direct algorithms
// ----------------------------------------------------------------------------------
template <typename U, typename W, int OP, int ALGO, int NR>
__global__ void __launch_bounds__(kThreads, 1)
allreduce_direct_kernel(const __grid_constant__ DevComm c, const U* __restrict__ in, U* __restrict__ out,
                        long long n, float scale, int flags, int root) {
  constexpr int kEpp = WireTraits<W>::kEpp;
  BarrierState epoch = barrier_begin(c);
  const int na = c.n_active, me = c.my_index;
  const bool zero_copy = flags & F_ZERO_COPY;
  const bool root_only = flags & F_ROOT_ONLY;
  const bool in_vec = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  const bool out_vec = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  char* const local = c.data[c.rank];

  Partition P;
  P.npacks = (n + kEpp - 1) / kEpp;
  P.nslices = (ALGO == ONE_SHOT) ? 1 : na;
  P.pps = (P.npacks + P.nslices - 1) / P.nslices;

  // ---- phase 0: publish my contribution -------------------------------------------
  if (!zero_copy) stage_in<U, W>(P, in, n, in_vec, local);
  block_barrier(c, epoch);

  // ---- phase 1: reduce (+ redistribute) -------------------------------------------
  if (ALGO == ONE_SHOT) {
    // every rank pulls every peer's window and reduces locally; the result goes straight
    // to the user tensor (no stage-out pass).
    if (!root_only || me == root) one_shot_phase1<U, W, OP, NR>(c, P.npacks, out, n, out_vec, scale);
    // peers may still be reading my window: nobody leaves before everyone is done
    block_barrier(c, epoch);
  } else {
    const long long base = P.slice_begin(me), cnt = P.slice_count(me);
    const TailPack tail = tail_of<W>(n, zero_copy);
    if (ALGO == TWO_SHOT) two_shot_phase1<W, OP, NR>(c, base, cnt, zero_copy, scale, root_only, root, SubGrid(), tail);
    else nvls_phase1<W, OP>(c, base, cnt, zero_copy, scale, root_only, root, SubGrid(), tail);
    block_barrier(c, epoch);
    // ---- phase 2: hand the result back to the caller ------------------------------
    if (!zero_copy && (!root_only || me == root))
      stage_out<U, W>(P, out, n, out_vec, local, scale);
  }
  finish_op(c, epoch);
}

// ----------------------------------------------------------------------------------
// broadcast, direct: root pushes (multimem.st when available, else one store per peer)
// ----------------------------------------------------------------------------------
template <typename U, typename W>
__global__ void __launch_bounds__(kThreads, 1)
broadcast_direct_kernel(const __grid_constant__ DevComm c, U* __restrict__ buf, long long n, int root,
                        int use_mc, int flags) {
  constexpr int kEpp = WireTraits<W>::kEpp;
  BarrierState epoch = barrier_begin(c);
  const int na = c.n_active, me = c.my_index;
  const bool zero_copy = flags & F_ZERO_COPY;
  const bool vec = (reinterpret_cast<uintptr_t>(buf) & 15) == 0;
  Partition P;
  P.npacks = (n + kEpp - 1) / kEpp;
  P.nslices = 1;
  P.pps = P.npacks;
  const long long stride = (long long)gridDim.x * kThreads;
  const TailPack tail = tail_of<W>(n, zero_copy);

  // everyone must have entered the op before the root overwrites their window
  block_barrier(c, epoch);
  if (me == root) {
    for (long long j0 = (long long)blockIdx.x * kThreads + threadIdx.x; j0 < P.npacks; j0 += stride * kUnroll) {
      uint4 v[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = j0 + u * stride;
        if (j < P.npacks) {
          if (zero_copy) {
            v[u] = ld16(c.data[c.rank] + j * 16);
          } else {
            float f[kEpp];
            load_user<U, kEpp>(buf, j * kEpp, n, vec, f);
            v[u] = pack<W>(f);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long j = j0 + u * stride;
        if (j < P.npacks) {
          if (j == tail.pack) {                // partial last pack of a zero-copy tensor
            float f[kEpp];
            unpack<W>(v[u], f);
            for (int a = 0; a < na; ++a)
              if (a != me) st_tail<W>(c.data[c.active_ranks[a]] + j * 16, f, tail.elems);
          } else if (use_mc) {
            mc_st16(c.mc_data + j * 16, v[u]);
          } else {
#pragma unroll
            for (int a = 0; a < kMaxRanks; ++a)
              if (a < na && a != me) st16(c.data[c.active_ranks[a]] + j * 16, v[u]);
          }
        }
      }
    }
  }
  block_barrier(c, epoch);
  if (!zero_copy && me != root) stage_out<U, W>(P, buf, n, vec, c.data[c.rank], 1.f);
  finish_op(c, epoch);
}


// ----------------------------------------------------------------------------------
// all-to-all (equal splits): rank r pushes block p of its input straight into slot r of peer
// p's window (128-bit NVLink stores), one barrier, local copy-out. The reference declares
// ALLTOALL=4 but never implements it (adapcc.py:59-61 calls a missing method).
// ----------------------------------------------------------------------------------
template <typename U>
__global__ void __launch_bounds__(kThreads, 1)
alltoall_kernel(const __grid_constant__ DevComm c, const U* __restrict__ in, U* __restrict__ out,
                long long per_peer) {
  constexpr int kEpp = WireTraits<U>::kEpp;
  BarrierState epoch = barrier_begin(c);
  const int na = c.n_active, me = c.my_index;
  const bool in_vec = (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (per_peer % kEpp) == 0;
  const bool out_vec = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (per_peer % kEpp) == 0;
  const long long ppacks = (per_peer + kEpp - 1) / kEpp;       // packs per block (padded)
  const long long stride = (long long)gridDim.x * kThreads;
  block_barrier(c, epoch);                                     // every window is free again
  for (int a = 0; a < na; ++a) {
    int idx = me + a; if (idx >= na) idx -= na;                // stagger destinations
    char* dst = c.data[c.active_ranks[idx]] + (long long)me * ppacks * 16;
    const U* src = in + (long long)idx * per_peer;
    for (long long j = (long long)blockIdx.x * kThreads + threadIdx.x; j < ppacks; j += stride) {
      float f[kEpp];
      load_user<U, kEpp>(src, j * kEpp, per_peer, in_vec, f);
      st16(dst + j * 16, pack<U>(f));
    }
  }
  block_barrier(c, epoch);
  const char* local = c.data[c.rank];
  for (int a = 0; a < na; ++a) {
    for (long long j = (long long)blockIdx.x * kThreads + threadIdx.x; j < ppacks; j += stride) {
      float f[kEpp];
      unpack<U>(ld16(local + ((long long)a * ppacks + j) * 16), f);
      store_user<U, kEpp>(out + (long long)a * per_peer, j * kEpp, per_peer, out_vec, f);
    }
  }
  finish_op(c, epoch);
}

}  // namespace adapcc
