// check_p2p — stand-alone P2P sanity / bandwidth binary (no Python, no MPI).
//
// Role of the reference's CUDA-aware-MPI ping-pong (/root/reference/units-test/check-p2p/
// check_mpi_p2p.cu: rank 0 sends a device buffer to rank 1 with MPI_Send and back). On an NVSwitch
// box the question is different: can every GPU's SMs load from / store to every other GPU's memory,
// are peer atomics native, and what do SM-issued peer loads and stores sustain — those are the
// operations the collective kernels are made of. One process drives all visible GPUs:
//
//   for every ordered pair (i, j):  kernel on GPU i  STORES a pattern into GPU j's buffer,
//                                   kernel on GPU j  verifies it locally,
//                                   kernel on GPU i  LOADS GPU j's buffer and verifies it,
//                                   timed 128-bit peer-read and peer-write sweeps (CUDA events)
//
// Output: capability matrix, read/write GB/s matrices, "P2P OK" / "P2P FAILED" (exit code 1).
//   ./check_p2p [--mb 256] [--iters 5] [--quick]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));        \
      exit(2);                                                                                    \
    }                                                                                             \
  } while (0)

__device__ __forceinline__ unsigned pattern(size_t i, unsigned salt) {
  return (unsigned)(i * 2654435761u) ^ salt;
}

__global__ void fill_kernel(unsigned* dst, size_t n, unsigned salt) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = pattern(i, salt);
}

__global__ void verify_kernel(const unsigned* src, size_t n, unsigned salt, unsigned long long* bad) {
  unsigned long long local = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    local += src[i] != pattern(i, salt);
  if (local) atomicAdd(bad, local);
}

// 128-bit streaming read of a (peer) buffer; the xor keeps the loads alive
__global__ void read_kernel(const uint4* __restrict__ src, size_t nvec, unsigned* sink) {
  unsigned acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

__global__ void write_kernel(uint4* __restrict__ dst, size_t nvec, unsigned salt) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = make_uint4(salt, (unsigned)i, salt ^ 0xffffffffu, 0u);
}

__global__ void atomic_kernel(unsigned* peer_counter, int adds) {
  for (int k = 0; k < adds; ++k) atomicAdd_system(peer_counter, 1u);
}

int main(int argc, char** argv) {
  size_t mb = 256;
  int iters = 5;
  bool quick = false;
  for (int a = 1; a < argc; ++a) {
    if (!strcmp(argv[a], "--mb") && a + 1 < argc) mb = strtoull(argv[++a], nullptr, 10);
    else if (!strcmp(argv[a], "--iters") && a + 1 < argc) iters = atoi(argv[++a]);
    else if (!strcmp(argv[a], "--quick")) quick = true;
    else { fprintf(stderr, "usage: %s [--mb N] [--iters K] [--quick]\n", argv[0]); return 2; }
  }
  if (quick) { mb = 16; iters = 2; }
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n == 0) { printf("no CUDA device\n"); return 2; }
  const size_t bytes = mb << 20, words = bytes / 4, nvec = bytes / 16;
  printf("check_p2p: %d GPU(s), %zu MiB per buffer, %d timed iterations\n", n, mb, iters);

  std::vector<unsigned*> buf(n);
  std::vector<unsigned long long*> bad(n);
  std::vector<unsigned*> sink(n);
  std::vector<int> sms(n);
  for (int i = 0; i < n; ++i) {
    CK(cudaSetDevice(i));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, i));
    sms[i] = p.multiProcessorCount;
    printf("  GPU %d: %s, sm_%d%d, %d SMs, %.0f GiB\n", i, p.name, p.major, p.minor, p.multiProcessorCount,
           p.totalGlobalMem / 1073741824.0);
    CK(cudaMalloc(&buf[i], bytes));
    CK(cudaMalloc(&bad[i], sizeof(unsigned long long)));
    CK(cudaMalloc(&sink[i], sizeof(unsigned)));
    CK(cudaMemset(bad[i], 0, sizeof(unsigned long long)));
  }
  // capability matrix + enable
  std::vector<int> can(n * n, 0), native_atomic(n * n, 0);
  for (int i = 0; i < n; ++i) {
    CK(cudaSetDevice(i));
    for (int j = 0; j < n; ++j) {
      if (i == j) { can[i * n + j] = 1; native_atomic[i * n + j] = 1; continue; }
      CK(cudaDeviceCanAccessPeer(&can[i * n + j], i, j));
      if (can[i * n + j]) {
        cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
        (void)cudaGetLastError();
        CK(cudaDeviceGetP2PAttribute(&native_atomic[i * n + j], cudaDevP2PAttrNativeAtomicSupported, i, j));
      }
    }
  }
  printf("peer access (row = accessing GPU; A = access + native atomics, a = access only, . = none)\n");
  for (int i = 0; i < n; ++i) {
    printf("  %d: ", i);
    for (int j = 0; j < n; ++j) printf("%c ", !can[i * n + j] ? '.' : native_atomic[i * n + j] ? 'A' : 'a');
    printf("\n");
  }

  bool ok = true;
  std::vector<double> rd(n * n, 0.0), wr(n * n, 0.0);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      if (!can[i * n + j]) { if (i != j) ok = false; continue; }
      const unsigned salt = 0x9e3779b9u * (unsigned)(i * n + j + 1);
      const int grid = sms[i] * 4;
      // (1) i stores into j, j verifies locally
      CK(cudaSetDevice(i));
      fill_kernel<<<grid, 512>>>(buf[j], words, salt);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      CK(cudaSetDevice(j));
      CK(cudaMemset(bad[j], 0, sizeof(unsigned long long)));
      verify_kernel<<<sms[j] * 4, 512>>>(buf[j], words, salt, bad[j]);
      CK(cudaGetLastError());
      unsigned long long nb = 0;
      CK(cudaMemcpy(&nb, bad[j], sizeof(nb), cudaMemcpyDeviceToHost));
      if (nb) { printf("  FAIL: %llu words stored by GPU %d into GPU %d read back wrong\n", nb, i, j); ok = false; }
      // (2) i loads j's buffer and verifies
      CK(cudaSetDevice(i));
      CK(cudaMemset(bad[i], 0, sizeof(unsigned long long)));
      verify_kernel<<<grid, 512>>>(buf[j], words, salt, bad[i]);
      CK(cudaGetLastError());
      CK(cudaMemcpy(&nb, bad[i], sizeof(nb), cudaMemcpyDeviceToHost));
      if (nb) { printf("  FAIL: %llu words of GPU %d's buffer loaded wrong by GPU %d\n", nb, j, i); ok = false; }
      // (3) system-scope atomics land
      if (native_atomic[i * n + j]) {
        CK(cudaSetDevice(j));
        CK(cudaMemset(buf[j], 0, 4));
        CK(cudaDeviceSynchronize());
        CK(cudaSetDevice(i));
        atomic_kernel<<<8, 64>>>(buf[j], 4);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        unsigned got = 0;
        CK(cudaMemcpy(&got, buf[j], 4, cudaMemcpyDeviceToHost));
        if (got != 8u * 64u * 4u) { printf("  FAIL: peer atomics %d -> %d counted %u of %u\n", i, j, got, 8u * 64u * 4u); ok = false; }
      }
      // (4) timed sweeps (events on the issuing GPU's stream)
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      float ms = 0.f;
      read_kernel<<<grid, 512>>>((const uint4*)buf[j], nvec, sink[i]);          // warm-up
      CK(cudaEventRecord(e0));
      for (int k = 0; k < iters; ++k) read_kernel<<<grid, 512>>>((const uint4*)buf[j], nvec, sink[i]);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      rd[i * n + j] = (double)bytes * iters / (ms * 1e-3) / 1e9;
      write_kernel<<<grid, 512>>>((uint4*)buf[j], nvec, salt);                   // warm-up
      CK(cudaEventRecord(e0));
      for (int k = 0; k < iters; ++k) write_kernel<<<grid, 512>>>((uint4*)buf[j], nvec, salt + k);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      wr[i * n + j] = (double)bytes * iters / (ms * 1e-3) / 1e9;
      CK(cudaGetLastError());
      CK(cudaEventDestroy(e0));
      CK(cudaEventDestroy(e1));
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    printf("%s GB/s (row = issuing GPU, column = memory owner; diagonal = local HBM)\n",
           pass == 0 ? "SM-issued 128-bit READ" : "SM-issued 128-bit WRITE");
    for (int i = 0; i < n; ++i) {
      printf("  %d: ", i);
      for (int j = 0; j < n; ++j) printf("%8.1f", (pass == 0 ? rd : wr)[i * n + j]);
      printf("\n");
    }
  }
  for (int i = 0; i < n; ++i) {
    CK(cudaSetDevice(i));
    CK(cudaFree(buf[i]));
    CK(cudaFree(bad[i]));
    CK(cudaFree(sink[i]));
  }
  printf("%s\n", ok ? "P2P OK" : "P2P FAILED");
  return ok ? 0 : 1;
}
