#include "bootstrap.h"

#include <errno.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include <cstring>

#include "common.h"

namespace adapcc {

static int64_t now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (int64_t)ts.tv_sec * 1000 + ts.tv_nsec / 1000000;
}

static socklen_t make_addr(const std::string& name, int rank, sockaddr_un* addr) {
  memset(addr, 0, sizeof(*addr));
  addr->sun_family = AF_UNIX;
  // abstract namespace: leading NUL, no filesystem entry to clean up
  std::string s = "adapcc-" + name + "-" + std::to_string(rank);
  if (s.size() > sizeof(addr->sun_path) - 2) s.resize(sizeof(addr->sun_path) - 2);
  memcpy(addr->sun_path + 1, s.data(), s.size());
  return (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + s.size());
}

Bootstrap::~Bootstrap() { close_all(); }

void Bootstrap::close_all() {
  for (int& s : socks_) {
    if (s >= 0) ::close(s);
    s = -1;
  }
  if (listen_fd_ >= 0) ::close(listen_fd_);
  listen_fd_ = -1;
}

int Bootstrap::init(const std::string& name, int rank, int world, int timeout_ms) {
  rank_ = rank;
  world_ = world;
  socks_.assign(world, -1);
  if (world == 1) return 0;

  listen_fd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (listen_fd_ < 0) { set_error("bootstrap: socket(): %s", strerror(errno)); return -1; }
  sockaddr_un addr;
  socklen_t alen = make_addr(name, rank, &addr);
  if (::bind(listen_fd_, (sockaddr*)&addr, alen) != 0) {
    set_error("bootstrap: bind(%s,%d): %s", name.c_str(), rank, strerror(errno));
    return -1;
  }
  if (::listen(listen_fd_, world + 4) != 0) {
    set_error("bootstrap: listen(): %s", strerror(errno));
    return -1;
  }

  // Higher rank connects to lower rank. connect() is retried until the peer's
  // listening socket exists.
  int64_t deadline = now_ms() + timeout_ms;
  for (int peer = 0; peer < rank; ++peer) {
    sockaddr_un pa;
    socklen_t pl = make_addr(name, peer, &pa);
    int s = -1;
    while (true) {
      s = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
      if (s < 0) { set_error("bootstrap: socket(): %s", strerror(errno)); return -1; }
      if (::connect(s, (sockaddr*)&pa, pl) == 0) break;
      ::close(s);
      s = -1;
      if (now_ms() > deadline) {
        set_error("bootstrap: rank %d timed out connecting to rank %d (%s)", rank, peer,
                  strerror(errno));
        return -1;
      }
      usleep(2000);
    }
    int32_t me = rank;
    if (::send(s, &me, sizeof(me), MSG_NOSIGNAL) != (ssize_t)sizeof(me)) {
      set_error("bootstrap: hello send failed: %s", strerror(errno));
      return -1;
    }
    socks_[peer] = s;
  }
  for (int n = rank + 1; n < world; ++n) {
    pollfd pfd{listen_fd_, POLLIN, 0};
    int64_t left = deadline - now_ms();
    if (left < 0) left = 0;
    int pr = ::poll(&pfd, 1, (int)left);
    if (pr <= 0) { set_error("bootstrap: rank %d timed out in accept", rank); return -1; }
    int s = ::accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
    if (s < 0) { set_error("bootstrap: accept(): %s", strerror(errno)); return -1; }
    int32_t who = -1;
    if (::recv(s, &who, sizeof(who), MSG_WAITALL) != (ssize_t)sizeof(who) || who <= rank ||
        who >= world || socks_[who] != -1) {
      set_error("bootstrap: bad hello (%d)", who);
      return -1;
    }
    socks_[who] = s;
  }
  return barrier();
}

int Bootstrap::send_bytes(int peer, const void* p, size_t n) {
  const char* c = (const char*)p;
  while (n) {
    ssize_t k = ::send(socks_[peer], c, n, MSG_NOSIGNAL);
    if (k <= 0) {
      if (k < 0 && errno == EINTR) continue;
      set_error("bootstrap: send to %d failed: %s", peer, strerror(errno));
      return -1;
    }
    c += k;
    n -= (size_t)k;
  }
  return 0;
}

int Bootstrap::recv_bytes(int peer, void* p, size_t n) {
  char* c = (char*)p;
  while (n) {
    ssize_t k = ::recv(socks_[peer], c, n, 0);
    if (k <= 0) {
      if (k < 0 && errno == EINTR) continue;
      set_error("bootstrap: recv from %d failed: %s", peer, k == 0 ? "closed" : strerror(errno));
      return -1;
    }
    c += k;
    n -= (size_t)k;
  }
  return 0;
}

int Bootstrap::send_fd(int peer, int fd) {
  char byte = 'F';
  iovec iov{&byte, 1};
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  cm->cmsg_level = SOL_SOCKET;
  cm->cmsg_type = SCM_RIGHTS;
  cm->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(cm), &fd, sizeof(int));
  while (true) {
    ssize_t k = ::sendmsg(socks_[peer], &msg, MSG_NOSIGNAL);
    if (k == 1) return 0;
    if (k < 0 && errno == EINTR) continue;
    set_error("bootstrap: sendmsg(fd) to %d failed: %s", peer, strerror(errno));
    return -1;
  }
}

int Bootstrap::recv_fd(int peer, int* fd) {
  char byte = 0;
  iovec iov{&byte, 1};
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof(ctrl));
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  while (true) {
    ssize_t k = ::recvmsg(socks_[peer], &msg, MSG_CMSG_CLOEXEC);
    if (k == 1) break;
    if (k < 0 && errno == EINTR) continue;
    set_error("bootstrap: recvmsg(fd) from %d failed: %s", peer,
              k == 0 ? "closed" : strerror(errno));
    return -1;
  }
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  if (!cm || cm->cmsg_level != SOL_SOCKET || cm->cmsg_type != SCM_RIGHTS) {
    set_error("bootstrap: no fd in message from %d", peer);
    return -1;
  }
  memcpy(fd, CMSG_DATA(cm), sizeof(int));
  return 0;
}

int Bootstrap::allgather(const void* mine, size_t n, void* all) {
  char* out = (char*)all;
  memcpy(out + (size_t)rank_ * n, mine, n);
  // Blobs are small (handles, a few hundred bytes): they fit the socket buffers, so
  // send-all-then-receive-all cannot deadlock.
  for (int p = 0; p < world_; ++p)
    if (p != rank_ && send_bytes(p, mine, n)) return -1;
  for (int p = 0; p < world_; ++p)
    if (p != rank_ && recv_bytes(p, out + (size_t)p * n, n)) return -1;
  return 0;
}

int Bootstrap::exchange_fds(int my_fd, std::vector<int>& fds_out) {
  fds_out.assign(world_, -1);
  fds_out[rank_] = my_fd;
  for (int p = 0; p < world_; ++p)
    if (p != rank_ && send_fd(p, my_fd)) return -1;
  for (int p = 0; p < world_; ++p)
    if (p != rank_ && recv_fd(p, &fds_out[p])) return -1;
  return 0;
}

int Bootstrap::bcast_fd(int root, int my_fd, int* fd_out) {
  if (rank_ == root) {
    for (int p = 0; p < world_; ++p)
      if (p != rank_ && send_fd(p, my_fd)) return -1;
    *fd_out = my_fd;
    return 0;
  }
  return recv_fd(root, fd_out);
}

int Bootstrap::barrier() {
  char token = 'B';
  std::vector<char> all((size_t)world_ > 0 ? world_ : 1);
  return allgather(&token, 1, all.data());
}

}  // namespace adapcc
