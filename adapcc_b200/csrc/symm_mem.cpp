#include "symm_mem.h"

#include <unistd.h>

#include "driver_api.h"

namespace adapcc {

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

bool SymmContext::all_agree(bool mine) {
  if (world_ == 1) return mine;
  char m = mine ? 1 : 0;
  std::vector<char> all(world_);
  if (boot_.allgather(&m, 1, all.data())) return false;
  for (char c : all)
    if (!c) return false;
  return true;
}

int SymmContext::init(const std::string& name, int rank, int world, int device) {
  rank_ = rank;
  world_ = world;
  device_ = device;
  if (world > kMaxRanks) { set_error("world %d > kMaxRanks %d", world, kMaxRanks); return -1; }
  CUDA_TRY(cudaSetDevice(device));
  CUDA_TRY(cudaFree(0));
  if (boot_.init(name, rank, world)) return -1;

  const DriverApi& d = driver();
  bool vmm = d.ok, mc = d.has_multicast;
  if (vmm) {
    CUdevice dev;
    int v = 0;
    if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) vmm = false;
    if (vmm && (d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED,
                                     dev) != CUDA_SUCCESS || !v))
      vmm = false;
    v = 0;
    if (!vmm || d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS || !v)
      mc = false;
  } else {
    mc = false;
  }
  const char* e = getenv("ADAPCC_SYMM_BACKEND");
  if (e && std::string(e) == "ipc") vmm = mc = false;
  const char* m = getenv("ADAPCC_DISABLE_MULTICAST");
  if (m && atoi(m)) mc = false;
  // P2P reachability is established (or fails loudly) when the peers' allocations are mapped below:
  // cuMemSetAccess / cudaIpcOpenMemHandle reject devices outside the NVLink / PCIe P2P domain.
  vmm_ok_ = all_agree(vmm);
  mc_ok_ = vmm_ok_ && all_agree(mc);
  ADAPCC_LOG(1, "rank %d/%d dev %d: vmm=%d multicast=%d", rank, world, device, (int)vmm_ok_,
             (int)mc_ok_);
  return 0;
}

void SymmContext::destroy() { boot_.close_all(); }

int SymmContext::alloc(size_t bytes, bool want_mc, SymmBuffer* out) {
  *out = SymmBuffer();
  if (bytes == 0) { set_error("symm alloc of 0 bytes"); return -1; }
  if (vmm_ok_) {
    // a multicast object needs >= 2 devices (cuMulticastCreate rejects numDevices = 1)
    int rc = alloc_vmm(bytes, want_mc && mc_ok_ && world_ > 1, out);
    if (rc == 0) return 0;
    // alloc_vmm only fails collectively (all ranks agree) before any mapping is kept.
    ADAPCC_LOG(1, "rank %d: VMM symmetric alloc failed (%s); falling back to cudaIpc", rank_,
               get_error());
    vmm_ok_ = false;
    mc_ok_ = false;
  }
  return alloc_ipc(bytes, out);
}

int SymmContext::alloc_vmm(size_t bytes, bool want_mc, SymmBuffer* out) {
  const DriverApi& d = driver();
  CUdevice dev;
  CU_TRY(d.DeviceGet(&dev, device_));

  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device_;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;

  size_t gran = 0;
  bool ok = d.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) ==
            CUDA_SUCCESS;
  if (!ok || gran == 0) gran = 2u << 20;
  size_t mc_gran = 0;
  if (want_mc) {
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)world_;
    mp.size = round_up(bytes, gran);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (d.MulticastGetGranularity(&mc_gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS ||
        mc_gran == 0)
      want_mc = false;
    else if (mc_gran > gran)
      gran = mc_gran;
  }
  want_mc = all_agree(want_mc);
  size_t size = round_up(bytes, gran);

  CUmemGenericAllocationHandle mine = 0;
  int my_fd = -1;
  bool local_ok = d.MemCreate(&mine, size, &prop, 0) == CUDA_SUCCESS;
  if (local_ok)
    local_ok = d.MemExportToShareableHandle(&my_fd, mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) ==
               CUDA_SUCCESS;
  if (!all_agree(local_ok)) {
    if (my_fd >= 0) ::close(my_fd);
    if (mine) d.MemRelease(mine);
    set_error("cuMemCreate/export(%zu bytes, posix fd) failed on some rank", size);
    return -1;
  }

  std::vector<int> fds(world_, -1);
  if (world_ > 1) {
    if (boot_.exchange_fds(my_fd, fds)) return -1;
  } else {
    fds[0] = my_fd;
  }

  bool map_ok = true;
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (int p = 0; p < world_ && map_ok; ++p) {
    CUmemGenericAllocationHandle h = mine;
    if (p != rank_) {
      CUresult r = d.MemImportFromShareableHandle(&h, (void*)(uintptr_t)fds[p],
                                                  CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      if (r != CUDA_SUCCESS) {
        set_error("cuMemImportFromShareableHandle(rank %d): %s", p, cu_error_string(r));
        map_ok = false;
        break;
      }
    }
    out->handles[p] = h;
    CUdeviceptr va = 0;
    CUresult r = d.MemAddressReserve(&va, size, gran, 0, 0);
    if (r == CUDA_SUCCESS) r = d.MemMap(va, size, 0, h, 0);
    if (r == CUDA_SUCCESS) r = d.MemSetAccess(va, size, &acc, 1);
    if (r != CUDA_SUCCESS) {
      set_error("map peer %d buffer: %s", p, cu_error_string(r));
      map_ok = false;
      break;
    }
    out->peers[p] = (void*)va;
  }
  for (int p = 0; p < world_; ++p)
    if (fds[p] >= 0) ::close(fds[p]);
  if (!all_agree(map_ok)) {
    // leave whatever was mapped to process teardown; report collectively
    if (map_ok) set_error("peer mapping failed on another rank");
    return -1;
  }
  out->size = size;
  out->backend = SYMM_VMM;
  CUDA_TRY(cudaMemset(out->peers[rank_], 0, size));
  CUDA_TRY(cudaDeviceSynchronize());

  if (want_mc) {
    if (setup_multicast(size, out) != 0) {
      ADAPCC_LOG(1, "rank %d: multicast setup failed (%s); NVLS disabled", rank_, get_error());
      out->mc = nullptr;
      mc_ok_ = false;
    }
  }
  if (world_ > 1 && boot_.barrier()) return -1;
  return 0;
}

int SymmContext::setup_multicast(size_t size, SymmBuffer* out) {
  const DriverApi& d = driver();
  CUdevice dev;
  CU_TRY(d.DeviceGet(&dev, device_));
  CUmulticastObjectProp mp{};
  mp.numDevices = (unsigned)world_;
  mp.size = size;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  mp.flags = 0;

  CUmemGenericAllocationHandle mc = 0;
  int fd = -1;
  bool ok = true;
  if (rank_ == 0) {
    CUresult r = d.MulticastCreate(&mc, &mp);
    if (r == CUDA_SUCCESS)
      r = d.MemExportToShareableHandle(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) {
      set_error("cuMulticastCreate/export: %s", cu_error_string(r));
      ok = false;
    }
  }
  if (!all_agree(ok)) { if (ok) set_error("multicast create failed on rank 0"); return -1; }
  if (world_ > 1) {
    int got = -1;
    if (boot_.bcast_fd(0, fd, &got)) return -1;
    if (rank_ != 0) {
      CUresult r = d.MemImportFromShareableHandle(&mc, (void*)(uintptr_t)got,
                                                  CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      if (r != CUDA_SUCCESS) { set_error("import multicast handle: %s", cu_error_string(r)); ok = false; }
      ::close(got);
    }
  }
  if (fd >= 0) ::close(fd);
  if (ok) {
    CUresult r = d.MulticastAddDevice(mc, dev);
    if (r != CUDA_SUCCESS) { set_error("cuMulticastAddDevice: %s", cu_error_string(r)); ok = false; }
  }
  // every device must be added before any memory is bound
  if (!all_agree(ok)) { if (ok) set_error("multicast add-device failed on another rank"); return -1; }
  {
    CUresult r = d.MulticastBindMem(mc, 0, out->handles[rank_], 0, size, 0);
    if (r != CUDA_SUCCESS) { set_error("cuMulticastBindMem: %s", cu_error_string(r)); ok = false; }
  }
  if (!all_agree(ok)) { if (ok) set_error("multicast bind failed on another rank"); return -1; }
  CUdeviceptr va = 0;
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  size_t gran = 2u << 20;
  CUresult r = d.MemAddressReserve(&va, size, gran, 0, 0);
  if (r == CUDA_SUCCESS) r = d.MemMap(va, size, 0, mc, 0);
  if (r == CUDA_SUCCESS) r = d.MemSetAccess(va, size, &acc, 1);
  if (r != CUDA_SUCCESS) { set_error("map multicast va: %s", cu_error_string(r)); ok = false; }
  if (!all_agree(ok)) { if (ok) set_error("multicast map failed on another rank"); return -1; }
  out->mc = (void*)va;
  out->mc_handle = mc;
  out->mc_bound = true;
  return 0;
}

int SymmContext::alloc_ipc(size_t bytes, SymmBuffer* out) {
  size_t size = round_up(bytes, 2u << 20);
  void* mine = nullptr;
  bool ok = cudaMalloc(&mine, size) == cudaSuccess;
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  if (ok && world_ > 1) ok = cudaIpcGetMemHandle(&h, mine) == cudaSuccess;
  if (!all_agree(ok)) {
    set_error("cudaMalloc/cudaIpcGetMemHandle(%zu) failed: %s", size,
              cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  CUDA_TRY(cudaMemset(mine, 0, size));
  CUDA_TRY(cudaDeviceSynchronize());
  std::vector<cudaIpcMemHandle_t> all(world_);
  if (world_ > 1) {
    if (boot_.allgather(&h, sizeof(h), all.data())) return -1;
  }
  bool map_ok = true;
  for (int p = 0; p < world_; ++p) {
    if (p == rank_) { out->peers[p] = mine; continue; }
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[p], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
      map_ok = false;
      break;
    }
    out->peers[p] = ptr;
  }
  if (!all_agree(map_ok)) { if (map_ok) set_error("cudaIpc mapping failed on another rank"); return -1; }
  out->size = size;
  out->backend = SYMM_CUDA_IPC;
  out->mc = nullptr;
  if (world_ > 1 && boot_.barrier()) return -1;
  return 0;
}

int SymmContext::free(SymmBuffer* buf) {
  if (!buf || buf->size == 0) return 0;
  cudaDeviceSynchronize();
  if (world_ > 1) boot_.barrier();
  if (buf->backend == SYMM_VMM) {
    const DriverApi& d = driver();
    CUdevice dev;
    d.DeviceGet(&dev, device_);
    if (buf->mc) {
      d.MemUnmap((CUdeviceptr)buf->mc, buf->size);
      d.MemAddressFree((CUdeviceptr)buf->mc, buf->size);
    }
    if (buf->mc_bound) d.MulticastUnbind(buf->mc_handle, dev, 0, buf->size);
    if (buf->mc_handle) d.MemRelease(buf->mc_handle);
    for (int p = 0; p < world_; ++p) {
      if (!buf->peers[p]) continue;
      d.MemUnmap((CUdeviceptr)buf->peers[p], buf->size);
      d.MemAddressFree((CUdeviceptr)buf->peers[p], buf->size);
      if (buf->handles[p]) d.MemRelease(buf->handles[p]);
    }
  } else if (buf->backend == SYMM_CUDA_IPC) {
    for (int p = 0; p < world_; ++p) {
      if (!buf->peers[p]) continue;
      if (p == rank_) cudaFree(buf->peers[p]);
      else cudaIpcCloseMemHandle(buf->peers[p]);
    }
  }
  *buf = SymmBuffer();
  if (world_ > 1) boot_.barrier();
  return 0;
}

}  // namespace adapcc
