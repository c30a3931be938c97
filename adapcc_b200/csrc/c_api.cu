// C ABI of libadapcc.so.
//
// Two layers:
//  (1) adapcc_* : the stream/dtype-aware API the Python package binds with ctypes.
//  (2) initThreads / exitThreads / allreduce / reduce / boardcast / updateActive : the six
//      symbols of the reference's communicator.so (/root/reference/csrc/run.cu:19-174), same
//      signatures, fp32, blocking — implemented on top of (1), no MPI, no threads.
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include "bootstrap.h"
#include "comm_context.h"
#include "driver_api.h"

using namespace adapcc;

extern "C" {

const char* adapcc_last_error() { return get_error(); }

int adapcc_version() { return 100; }
long long adapcc_launch_count() { return launch_count(); }

void* adapcc_ctx_create(const char* name, int rank, int world, int device, unsigned long long staging_bytes,
                        unsigned long long heap_bytes) {
  CommContext* c = new CommContext();
  if (c->init(name ? name : "default", rank, world, device, (size_t)staging_bytes, (size_t)heap_bytes)) {
    delete c;
    return nullptr;
  }
  return c;
}

int adapcc_ctx_destroy(void* h) {
  if (!h) return 0;
  CommContext* c = static_cast<CommContext*>(h);
  c->destroy();
  delete c;
  return 0;
}

// info[0]=symm backend (0 VMM, 1 cudaIpc) [1]=multicast on staging [2]=multicast on heap
// [3]=rank [4]=world [5]=number of strategy trees
int adapcc_ctx_info(void* h, int* info) {
  CommContext* c = static_cast<CommContext*>(h);
  info[0] = c->symm_backend();
  info[1] = c->has_multicast();
  info[2] = c->heap_multicast();
  info[3] = c->rank();
  info[4] = c->world();
  info[5] = (int)c->strategy().trees.size();
  return 0;
}

void* adapcc_ctx_heap_ptr(void* h) { return static_cast<CommContext*>(h)->heap_ptr(); }
unsigned long long adapcc_ctx_heap_bytes(void* h) { return static_cast<CommContext*>(h)->heap_bytes(); }
unsigned long long adapcc_ctx_staging_bytes(void* h) { return static_cast<CommContext*>(h)->staging_bytes(); }
void* adapcc_ctx_peer_heap_ptr(void* h, int r) { return static_cast<CommContext*>(h)->peer_heap_ptr(r); }
// multicast alias of the symmetric heap (NULL when no multicast object is bound)
void* adapcc_ctx_heap_mc_ptr(void* h) { return static_cast<CommContext*>(h)->heap_mc_ptr(); }
void* adapcc_ctx_peer_staging_ptr(void* h, int r) { return static_cast<CommContext*>(h)->peer_staging_ptr(r); }
int adapcc_ctx_last_algo(void* h) { return static_cast<CommContext*>(h)->last_algo; }

// keys: 0 max_blocks, 1 one_shot_max_bytes, 2 nvls_min_bytes, 3 relay_mode, 4 timeout_ms,
// 5 tree_blocks, 6 tree_chunk_max_bytes
int adapcc_ctx_set_tunable(void* h, int key, long long value) {
  CommContext* c = static_cast<CommContext*>(h);
  switch (key) {
    case 0: c->tun.max_blocks = (int)std::max<long long>(1, std::min<long long>(value, kMaxBlocks)); break;
    case 1: c->tun.one_shot_max_bytes = value; break;
    case 2: c->tun.nvls_min_bytes = value; break;
    case 3: c->tun.relay_mode = (int)value; break;
    case 4: c->tun.timeout_ms = value; break;
    case 5: c->tun.tree_blocks = (int)std::max<long long>(1, std::min<long long>(value, kMaxBlocks)); break;
    case 6: c->tun.tree_chunk_max_bytes = value; break;
    case 7: c->tun.nvls_min_ranks = (int)value; break;
    case 8: c->tun.force_kernel = (int)value; break;
    case 9: c->tun.pipe_min_bytes = value; break;
    case 10: c->tun.pipe_stagers = (int)value; break;
    case 11: c->tun.pipe_links = (int)value; break;
    case 12: c->tun.pipe_piece_bytes = value; break;
    case 13: c->tun.pipe_nvls = (int)value; break;
    case 14: c->tun.ll_max_bytes = value; break;
    default: set_error("unknown tunable %d", key); return -1;
  }
  return 0;
}

int adapcc_ctx_load_strategy(void* h, const char* path) {
  return static_cast<CommContext*>(h)->load_strategy_file(path);
}
int adapcc_ctx_load_strategy_text(void* h, const char* xml) {
  return static_cast<CommContext*>(h)->load_strategy_text(xml);
}

static std::vector<int> sorted_active(const int* active, int n) {
  std::vector<int> v(active, active + n);
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
  return v;
}

int adapcc_allreduce(void* h, const void* in, void* out, long long count, int dtype, int wire, int op,
                     int algo, const int* active, int n_active, void* stream) {
  return static_cast<CommContext*>(h)->allreduce(in, out, count, dtype, wire, op, algo,
                                                 sorted_active(active, n_active), (cudaStream_t)stream);
}
int adapcc_reduce(void* h, const void* in, void* out, long long count, int dtype, int wire, int op, int algo,
                  int root, const int* active, int n_active, void* stream) {
  return static_cast<CommContext*>(h)->reduce(in, out, count, dtype, wire, op, algo, root,
                                              sorted_active(active, n_active), (cudaStream_t)stream);
}
int adapcc_broadcast(void* h, void* buf, long long count, int dtype, int root, const int* active, int n_active,
                     void* stream) {
  return static_cast<CommContext*>(h)->broadcast(buf, count, dtype, root, sorted_active(active, n_active),
                                                 (cudaStream_t)stream);
}
int adapcc_alltoall(void* h, const void* in, void* out, long long per_peer, int dtype, const int* active, int n_active,
                    void* stream) {
  return static_cast<CommContext*>(h)->alltoall(in, out, per_peer, dtype, sorted_active(active, n_active),
                                                (cudaStream_t)stream);
}
int adapcc_tree_collective(void* h, int prim, const void* in, void* out, long long count, int dtype, int wire,
                           int op, long long chunk_bytes, const int* active, int n_active, void* stream) {
  return static_cast<CommContext*>(h)->tree_collective(prim, in, out, count, dtype, wire, op, chunk_bytes,
                                                       sorted_active(active, n_active), (cudaStream_t)stream);
}
int adapcc_tree_relay_persistent(void* h, int n_buckets, const long long* counts, const long long* chunk_bytes,
                                 int wire, int op, const int* active, int n_active, void* stream) {
  return static_cast<CommContext*>(h)->tree_relay_persistent(n_buckets, counts, chunk_bytes, wire, op,
                                                             sorted_active(active, n_active), (cudaStream_t)stream);
}
// opt-in low-latency path (context created with ADAPCC_LL=1): all ranks, <= 32 KB
int adapcc_allreduce_ll(void* h, const void* in, void* out, long long count, int dtype, int op, void* stream) {
  return static_cast<CommContext*>(h)->allreduce_ll(in, out, count, dtype, op, (cudaStream_t)stream);
}
int adapcc_ctx_has_ll(void* h) { return static_cast<CommContext*>(h)->has_ll() ? 1 : 0; }
int adapcc_skip_op(void* h, void* stream) { return static_cast<CommContext*>(h)->skip_op((cudaStream_t)stream); }
int adapcc_ctx_check(void* h, void* stream) { return static_cast<CommContext*>(h)->check((cudaStream_t)stream); }
int adapcc_ctx_host_barrier(void* h) {
  CommContext* c = static_cast<CommContext*>(h);
  return c->world() > 1 ? c->symm().boot().barrier() : 0;
}

// ---- symmetric-heap allocator for torch.cuda.MemPool (CUDAPluggableAllocator ABI) ---------------
// DDP gradient buckets allocated inside `use_mem_pool` land in the symmetric heap, which makes
// the comm hook's all-reduce zero-copy. Bump allocation: every rank performs the same sequence of
// allocations, so offsets match across ranks; blocks are returned when the context is destroyed.
namespace {
std::mutex g_pool_mu;
CommContext* g_pool_ctx = nullptr;
size_t g_pool_off = 0;
size_t g_pool_limit = 0;
}  // namespace

int adapcc_pool_bind(void* h, unsigned long long start_offset) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  g_pool_ctx = static_cast<CommContext*>(h);
  g_pool_off = (size_t)start_offset;
  g_pool_limit = g_pool_ctx ? g_pool_ctx->heap_bytes() : 0;
  return 0;
}
unsigned long long adapcc_pool_offset() {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  return g_pool_off;
}
void* adapcc_pool_alloc(ssize_t size, int device, void* stream) {
  (void)device; (void)stream;
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (!g_pool_ctx || !g_pool_ctx->heap_ptr()) { set_error("pool_alloc: no symmetric heap bound"); return nullptr; }
  size_t off = (g_pool_off + 511) & ~(size_t)511;
  if (off + (size_t)size > g_pool_limit) {
    set_error("pool_alloc: symmetric heap exhausted (%zu + %zd > %zu); raise heap_mb", off, size, g_pool_limit);
    return nullptr;
  }
  g_pool_off = off + (size_t)size;
  return (char*)g_pool_ctx->heap_ptr() + off;
}
void adapcc_pool_free(void* ptr, ssize_t size, int device, void* stream) {
  (void)ptr; (void)size; (void)device; (void)stream;   // bump allocator: reclaimed with the context
}

// ---- bootstrap self-test (no GPU needed): every rank of a `world`-process group calls this with the
// same name. Exercises the full mesh: allgather, fd exchange + fd broadcast over SCM_RIGHTS (each fd
// is a pipe whose content names its owner, read back through the RECEIVED descriptor), barrier.
// Returns 0, or a negative step number on failure (details in adapcc_last_error()).
int adapcc_bootstrap_selftest(const char* name, int rank, int world, int timeout_ms) {
  Bootstrap bs;
  if (bs.init(name, rank, world, timeout_ms)) return -1;
  std::vector<long long> all(world, -1);
  const long long mine = 1000 + 7LL * rank;
  if (bs.allgather(&mine, sizeof(mine), all.data())) return -2;
  for (int r = 0; r < world; ++r)
    if (all[r] != 1000 + 7LL * r) { set_error("allgather: slot %d holds %lld", r, all[r]); return -2; }
  auto make_pipe = [](int tag, int* rd) -> int {      // a readable fd whose content is `tag`
    int p[2];
    if (pipe(p)) return -1;
    const int n = (int)write(p[1], &tag, sizeof(tag));
    close(p[1]);
    *rd = p[0];
    return n == (int)sizeof(tag) ? 0 : -1;
  };
  auto read_tag = [](int fd) -> int {
    int tag = -1;
    if (read(fd, &tag, sizeof(tag)) != (ssize_t)sizeof(tag)) return -1;
    return tag;
  };
  // fd broadcast from the last rank
  const int root = world - 1;
  int my_fd = -1, got = -1;
  if (rank == root && make_pipe(4242, &my_fd)) { set_error("pipe failed"); return -3; }
  if (bs.bcast_fd(root, my_fd, &got)) return -3;
  if (bs.barrier()) return -3;
  // one pipe, world readers: only check that the descriptor is valid on every rank; the root reads it
  if (got < 0 || fcntl(got, F_GETFD) < 0) { set_error("bcast_fd: invalid descriptor"); return -3; }
  if (rank == root && read_tag(got) != 4242) { set_error("bcast_fd: wrong content"); return -3; }
  if (got != my_fd) close(got);
  if (my_fd >= 0) close(my_fd);
  // all-to-all fd exchange: rank r's pipe holds (r + 1) tags of value 500 + r, one per reader
  int p[2];
  if (pipe(p)) { set_error("pipe failed"); return -4; }
  for (int i = 0; i < world; ++i) {
    const int tag = 500 + rank;
    if (write(p[1], &tag, sizeof(tag)) != (ssize_t)sizeof(tag)) { set_error("pipe write failed"); return -4; }
  }
  close(p[1]);
  std::vector<int> fds;
  if (bs.exchange_fds(p[0], fds)) return -4;
  if ((int)fds.size() != world) { set_error("exchange_fds: %d descriptors", (int)fds.size()); return -4; }
  int rc = 0;
  for (int r = 0; r < world; ++r) {
    const int tag = read_tag(fds[r]);                 // every rank consumes exactly one tag per pipe
    if (tag != 500 + r) { set_error("exchange_fds: fd of rank %d carried %d", r, tag); rc = -4; }
  }
  if (bs.barrier()) return -5;
  for (int r = 0; r < world; ++r)
    if (fds[r] >= 0 && fds[r] != p[0]) close(fds[r]);
  close(p[0]);
  bs.close_all();
  return rc;
}

// ---- strategy / relay-control queries (no GPU needed; used by tests and the control plane)
// out[0..3] = hasRecv, hasLocal, hasKernel, hasSend ; out[4] = number of active recvs,
// out[5..] = those child ranks. Returns the number of trees, or -1.
int adapcc_relay_control(const char* xml_text, int world, int tree, int rank, const int* active, int n_active,
                         int* out, int out_cap) {
  Strategy s;
  if (!s.load(xml_text, world)) return -1;
  if (tree < 0 || tree >= (int)s.trees.size()) { set_error("tree index out of range"); return -1; }
  int maxr = world > 0 ? world : 0;
  for (int i = 0; i < n_active; ++i) maxr = std::max(maxr, active[i] + 1);
  for (int x : s.trees[tree].nodes) maxr = std::max(maxr, x + 1);
  std::vector<bool> act(maxr, false);
  for (int i = 0; i < n_active; ++i) act[active[i]] = true;
  RelayControl rc = relay_control(s.trees[tree], rank, act);
  if (out_cap < 5 + (int)rc.active_recvs.size()) { set_error("output too small"); return -1; }
  out[0] = rc.has_recv; out[1] = rc.has_local; out[2] = rc.has_kernel; out[3] = rc.has_send;
  out[4] = (int)rc.active_recvs.size();
  for (size_t i = 0; i < rc.active_recvs.size(); ++i) out[5 + i] = rc.active_recvs[i];
  return (int)s.trees.size();
}

// out[0]=parent out[1]=flags out[2]=n_children out[3..]=children. Returns #trees or -1.
int adapcc_tree_role(const char* xml_text, int world, int tree, int rank, const int* active, int n_active,
                     int prim, int relay_mode, int* out, int out_cap) {
  Strategy s;
  if (!s.load(xml_text, world)) return -1;
  if (tree < 0 || tree >= (int)s.trees.size()) { set_error("tree index out of range"); return -1; }
  int maxr = world > 0 ? world : 0;
  for (int i = 0; i < n_active; ++i) maxr = std::max(maxr, active[i] + 1);
  for (int x : s.trees[tree].nodes) maxr = std::max(maxr, x + 1);
  std::vector<bool> act(maxr, false);
  for (int i = 0; i < n_active; ++i) act[active[i]] = true;
  HostTreeRole r = tree_role(s.trees[tree], rank, act, prim, relay_mode);
  if (out_cap < 3 + (int)r.children.size()) { set_error("output too small"); return -1; }
  out[0] = r.parent; out[1] = r.flags; out[2] = (int)r.children.size();
  for (size_t i = 0; i < r.children.size(); ++i) out[3 + i] = r.children[i];
  return (int)s.trees.size();
}

// ---------------------------------------------------------------------------------------
// Reference-compatible ABI
// ---------------------------------------------------------------------------------------
namespace {
std::mutex g_mu;
std::map<int, CommContext*> g_ctx;       // one context per primitive, like the reference
std::vector<int> g_extra_active;         // updateActive() additions for the next op
int g_init_count = 0;

int env_int(const char* a, const char* b, int dflt) {
  const char* v = getenv(a);
  if (!v && b) v = getenv(b);
  return v ? atoi(v) : dflt;
}
}  // namespace

void initThreads(int prim, char* filename, int sockPort) {
  std::lock_guard<std::mutex> lk(g_mu);
  const int rank = env_int("RANK", "OMPI_COMM_WORLD_RANK", 0);
  const int world = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1);
  const int local = env_int("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", 0);
  if (prim != ALLREDUCE && prim != REDUCE && prim != BOARDCAST) {
    fprintf(stderr, "[Rank %d]initThreads: primitive %d has no transmission context in the C ABI; "
                    "use the Python workflow (detect/profile)\n", rank, prim);
    return;
  }
  if (g_ctx.count(prim)) return;
  const char* sb = getenv("ADAPCC_STAGING_MB");
  size_t staging = (size_t)(sb ? atoll(sb) : 256) << 20;
  std::string name = "abi-" + std::to_string(sockPort) + "-" + std::to_string(prim) + "-" +
                     std::to_string(g_init_count++);
  CommContext* c = new CommContext();
  if (c->init(name, rank, world, local, staging, 0) || c->load_strategy_file(filename ? filename : "")) {
    fprintf(stderr, "[Rank %d]initThreads failed: %s\n", rank, get_error());
    delete c;
    return;
  }
  g_ctx[prim] = c;
  printf("[Rank %d]transmission context ready (prim %d, %zu trees)\n", rank, prim, c->strategy().trees.size());
}

void exitThreads(int prim) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_ctx.find(prim);
  if (it == g_ctx.end()) return;
  it->second->destroy();
  delete it->second;
  g_ctx.erase(it);
}

static void abi_collective(int prim, void* tensor, int size, int chunkBytes, int* activeGPU, int numGPU) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_ctx.find(prim);
  if (it == g_ctx.end()) { fprintf(stderr, "[adapcc] no context for primitive %d\n", prim); return; }
  std::vector<int> act(activeGPU, activeGPU + numGPU);
  for (int r : g_extra_active) act.push_back(r);
  g_extra_active.clear();
  std::sort(act.begin(), act.end());
  act.erase(std::unique(act.begin(), act.end()), act.end());
  CommContext* c = it->second;
  int rc = c->tree_collective(prim, tensor, tensor, size, F32, F32, SUM, chunkBytes, act, 0);
  if (!rc) rc = c->check(0);
  if (rc) fprintf(stderr, "[adapcc] primitive %d failed: %s\n", prim, get_error());
}

void allreduce(void* tensor, int size, int chunkBytes, int* activeGPU, int numGPU) {
  abi_collective(ALLREDUCE, tensor, size, chunkBytes, activeGPU, numGPU);
}
void reduce(void* tensor, int size, int chunkBytes, int* activeGPU, int numGPU) {
  abi_collective(REDUCE, tensor, size, chunkBytes, activeGPU, numGPU);
}
void boardcast(void* tensor, int size, int chunkBytes, int* activeGPU, int numGPU) {
  abi_collective(BOARDCAST, tensor, size, chunkBytes, activeGPU, numGPU);
}
void updateActive(int myRank) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_extra_active.push_back(myRank);
}

}  // extern "C"
