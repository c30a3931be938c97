// Strategy-driven tree collectives: one kernel executes the synthesised set of parallel
// reduction / broadcast trees ("transmissions"), chunk-pipelined, entirely on the device.
//
// Behavioural parity with the reference's allreduce/reduce/boardcast contexts
// (/root/reference/csrc/allreduce.cu:201-666, reduce.cu:153-367, boardcast.cu:152-318):
//   * the tensor is cut into one contiguous slice per tree, each slice into chunks,
//   * reduce runs leaf -> root, broadcast root -> leaf on the same tree, and a chunk starts
//     its way down as soon as the root has reduced it (reduce/broadcast pipelining),
//   * relay control: ranks that are not active contribute nothing but still forward.
// What is different (B200-first): no pthread per tree, no copy engine, no IPC events, no
// host sync per chunk. A parent PULLS its children's partial sums with 128-bit peer loads
// over NVLink straight out of the child's symmetric window (so there is no per-child
// staging slot), reduces in fp32 registers and publishes the chunk with one
// st.release.sys flag. Tails are handled (reference defect: dropped by integer division).
#pragma once
#include "device_prims.cuh"
#include "kernels_direct.cuh"

namespace adapcc {

struct TreeRole {
  int parent;                    // effective parent (world rank) or -1 when root
  int n_children;                // children pulled from in the reduce phase
  int children[kMaxChildren];
  int flags;
};

struct TreePlan {
  int n_trees;
  int do_reduce, do_bcast;
  long long chunk_packs;
  long long slice_begin[kMaxTrees + 1];   // in packs; slice t = [begin[t], begin[t+1])
  TreeRole role[kMaxTrees];
};

template <typename U, typename W, int OP>
__global__ void __launch_bounds__(kThreads, 1)
tree_collective_kernel(const __grid_constant__ DevComm c, const __grid_constant__ TreePlan plan,
                       const U* __restrict__ in, U* __restrict__ out, long long n, float scale) {
  constexpr int kEpp = WireTraits<W>::kEpp;
  uint32_t epoch = c.bar_epoch[blockIdx.x];
  const unsigned long long q = *c.seq;
  const bool in_vec = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  const bool out_vec = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  char* const local = c.data[c.rank];
  unsigned long long* const my_rflag = c.flag[c.rank] + blockIdx.x;
  unsigned long long* const my_bflag = c.flag[c.rank] + kMaxBlocks + blockIdx.x;

  long long max_k = 0;
  for (int t = 0; t < plan.n_trees; ++t) {
    const long long len = plan.slice_begin[t + 1] - plan.slice_begin[t];
    const long long nk = (len + plan.chunk_packs - 1) / plan.chunk_packs;
    max_k = nk > max_k ? nk : max_k;
  }
  const long long n_items = max_k * plan.n_trees;

  unsigned long long m = 0;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++m) {
    const int t = (int)(item % plan.n_trees);
    const long long k = item / plan.n_trees;
    const long long p0 = plan.slice_begin[t] + k * plan.chunk_packs;
    long long pc = plan.slice_begin[t + 1] - p0;
    if (pc <= 0) continue;
    if (pc > plan.chunk_packs) pc = plan.chunk_packs;
    const unsigned long long token = (q << 24) | (m + 1);
    const TreeRole& role = plan.role[t];
    const bool is_root = role.parent < 0;
    const int nc = role.n_children;

    // ------------------------------ reduce phase ---------------------------------
    if (plan.do_reduce && (role.flags & TR_IN_REDUCE)) {
      if (nc > 0) {
        if ((int)threadIdx.x < nc) wait_flag64(c, c.flag[role.children[threadIdx.x]] + blockIdx.x, token);
        __syncthreads();
      }
      const bool has_local = role.flags & TR_HAS_LOCAL;
      const bool to_user = is_root && (role.flags & TR_WANT_RESULT);
      const bool to_window = !is_root || (plan.do_bcast && (role.flags & TR_PUBLISH));
      for (long long j = threadIdx.x; j < pc; j += kThreads) {
        const long long pk = p0 + j;
        uint4 v[kMaxChildren];
#pragma unroll
        for (int a = 0; a < kMaxChildren; ++a)
          if (a < nc) v[a] = ld16(c.data[role.children[a]] + pk * 16);
        float acc[kEpp];
        if (has_local) {
          load_user<U, kEpp>(in, pk * kEpp, n, in_vec, acc);
        } else {
#pragma unroll
          for (int i = 0; i < kEpp; ++i) acc[i] = red_identity<OP>();
        }
#pragma unroll
        for (int a = 0; a < kMaxChildren; ++a)
          if (a < nc) {
            float f[kEpp];
            unpack<W>(v[a], f);
#pragma unroll
            for (int i = 0; i < kEpp; ++i) acc[i] = red_apply<OP>(acc[i], f[i]);
          }
        if (is_root) {
#pragma unroll
          for (int i = 0; i < kEpp; ++i) acc[i] *= scale;
        }
        if (to_window) st16(local + pk * 16, pack<W>(acc));
        if (to_user) store_user<U, kEpp>(out, pk * kEpp, n, out_vec, acc);
      }
      __syncthreads();
      if (threadIdx.x == 0) st_release_sys64(is_root ? my_bflag : my_rflag, token);
    } else if (!plan.do_reduce && plan.do_bcast && is_root) {
      // pure broadcast: the root publishes its tensor chunk by chunk
      for (long long j = threadIdx.x; j < pc; j += kThreads) {
        float f[kEpp];
        load_user<U, kEpp>(in, (p0 + j) * kEpp, n, in_vec, f);
        st16(local + (p0 + j) * 16, pack<W>(f));
      }
      __syncthreads();
      if (threadIdx.x == 0) st_release_sys64(my_bflag, token);
    }

    // ----------------------------- broadcast phase -------------------------------
    if (plan.do_bcast && !is_root && (role.flags & TR_IN_BCAST)) {
      if (threadIdx.x == 0) wait_flag64(c, c.flag[role.parent] + kMaxBlocks + blockIdx.x, token);
      __syncthreads();
      const bool publish = role.flags & TR_PUBLISH;
      const bool want = role.flags & TR_WANT_RESULT;
      const char* src = c.data[role.parent];
      for (long long j0 = threadIdx.x; j0 < pc; j0 += kThreads * kUnroll) {
        uint4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const long long j = j0 + (long long)u * kThreads;
          if (j < pc) v[u] = ld16(src + (p0 + j) * 16);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const long long j = j0 + (long long)u * kThreads;
          if (j < pc) {
            if (publish) st16(local + (p0 + j) * 16, v[u]);
            if (want) {
              float f[kEpp];
              unpack<W>(v[u], f);
              store_user<U, kEpp>(out, (p0 + j) * kEpp, n, out_vec, f);
            }
          }
        }
      }
      if (publish) {
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys64(my_bflag, token);
      }
    }
  }
  // windows are reused by the next op: nobody leaves while a peer may still be pulling
  block_barrier(c, epoch);
  finish_op(c, epoch);
}

// Keeps a non-participating rank's op sequence number in step with the others.
__global__ void skip_op_kernel(const __grid_constant__ DevComm c) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *c.seq = *c.seq + 1;
}

}  // namespace adapcc
