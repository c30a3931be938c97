// Single-node rendezvous over Unix-domain sockets (abstract namespace).
//
// Replaces the reference's MPI host-hash allgather (/root/reference/csrc/init.cu:30-51)
// and the two N x N TCP "barrier meshes" (/root/reference/csrc/trans.cu:102-230). It is used
// only at setup/teardown (handle exchange, host barriers); the data path never touches
// the host. SCM_RIGHTS fd passing is what lets CUDA VMM allocations (and the multicast
// object) be shared between the per-GPU processes without MPI or CUDA IPC events.
#pragma once
#include <string>
#include <vector>

namespace adapcc {

class Bootstrap {
 public:
  Bootstrap() = default;
  ~Bootstrap();
  // Establish the full mesh. `name` must be identical on every rank and unique per
  // context. Returns 0 on success.
  int init(const std::string& name, int rank, int world, int timeout_ms = 60000);
  void close_all();

  int rank() const { return rank_; }
  int world() const { return world_; }

  // all[i*n .. (i+1)*n) receives rank i's blob.
  int allgather(const void* mine, size_t n, void* all);
  // Every rank contributes one fd; fds_out[i] is a local duplicate of rank i's fd
  // (fds_out[rank] == my_fd).
  int exchange_fds(int my_fd, std::vector<int>& fds_out);
  // Root passes an fd to everyone (returns it in *fd_out; root gets its own back).
  int bcast_fd(int root, int my_fd, int* fd_out);
  int barrier();

 private:
  int send_bytes(int peer, const void* p, size_t n);
  int recv_bytes(int peer, void* p, size_t n);
  int send_fd(int peer, int fd);
  int recv_fd(int peer, int* fd);

  int rank_ = -1, world_ = 0;
  int listen_fd_ = -1;
  std::vector<int> socks_;  // socks_[peer]
};

}  // namespace adapcc
