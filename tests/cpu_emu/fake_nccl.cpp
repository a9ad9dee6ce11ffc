// TEST INFRASTRUCTURE (see cuda_emu.hpp): an in-process stand-in for the eight NCCL entry points lbm_comm.cu binds, so
// that the slab ring — ghost rows, moment / normal halos, pressure packets, the immersed body's row exchange — can be
// run on host cores with one THREAD per rank (tests/cpu_emu/ring_threads.py).  The emulated CUDA runtime executes
// every call immediately, so a send is a copy into a mailbox keyed (communicator, source, destination) and a receive
// takes the oldest message of its key, blocking until it is there; inside ncclGroupStart/End the receives are held back
// until the group ends (a group posts its operations in any order without deadlock, as NCCL guarantees).  A rank that
// waits longer than FAKE_NCCL_TIMEOUT_S (default 60) for a message aborts the process with the pending key: a ring
// that would hang on the GPU box fails the CPU test instead.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

extern "C"
{
typedef struct FakeComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef void* cudaStream_t;
}

namespace
{
std::mutex g_mu;
std::condition_variable g_cv;
std::map<std::tuple<std::string, int, int>, std::deque<std::vector<char>>> g_box;  // (communicator id, src, dst) -> messages
unsigned long long g_next_id = 1;

struct Pending
{
  void* dst;
  size_t bytes;
  int peer;
};
thread_local int tl_group_depth = 0;
thread_local std::vector<std::pair<struct FakeComm*, Pending>> tl_recvs;

size_t type_bytes(int t) { return t == 8 ? 8 : (t == 7 ? 4 : (t == 0 || t == 1 ? 1 : (t == 2 || t == 3 ? 4 : 8))); }
}  // namespace

struct FakeComm
{
  std::string id;
  int rank, n;
};

static int do_recv(FakeComm* c, const Pending& p)
{
  static const int timeout_s = std::getenv("FAKE_NCCL_TIMEOUT_S") ? std::atoi(std::getenv("FAKE_NCCL_TIMEOUT_S")) : 60;
  std::unique_lock<std::mutex> lk(g_mu);
  auto key = std::make_tuple(c->id, p.peer, c->rank);
  const bool ok = g_cv.wait_for(lk, std::chrono::seconds(timeout_s), [&] { return !g_box[key].empty(); });
  if (!ok)
  {
    std::fprintf(stderr, "fake_nccl: rank %d of %d waited %d s for a message from rank %d that was never sent (a hang on real NCCL)\n",
                 c->rank, c->n, timeout_s, p.peer);
    std::abort();
  }
  std::vector<char>& m = g_box[key].front();
  if (m.size() != p.bytes)
  {
    std::fprintf(stderr, "fake_nccl: rank %d expects %zu bytes from rank %d, the message holds %zu\n", c->rank, p.bytes, p.peer, m.size());
    std::abort();
  }
  std::memcpy(p.dst, m.data(), p.bytes);
  g_box[key].pop_front();
  return 0;
}

extern "C"
{
ncclResult_t ncclGetUniqueId(ncclUniqueId* u)
{
  std::lock_guard<std::mutex> lk(g_mu);
  std::memset(u->internal, 0, sizeof(u->internal));
  std::snprintf(u->internal, sizeof(u->internal), "fake-nccl-%llu", g_next_id++);
  return 0;
}
ncclResult_t ncclCommInitRank(ncclComm_t* c, int n, ncclUniqueId u, int rank)
{
  *c = new FakeComm{std::string(u.internal), rank, n};
  return 0;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) { delete c; return 0; }
ncclResult_t ncclSend(const void* buf, size_t count, int type, int peer, ncclComm_t c, cudaStream_t)
{
  const size_t bytes = count * type_bytes(type);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_box[std::make_tuple(c->id, c->rank, peer)].emplace_back((const char*)buf, (const char*)buf + bytes);
  }
  g_cv.notify_all();
  return 0;
}
ncclResult_t ncclRecv(void* buf, size_t count, int type, int peer, ncclComm_t c, cudaStream_t)
{
  Pending p{buf, count * type_bytes(type), peer};
  if (tl_group_depth > 0) { tl_recvs.emplace_back(c, p); return 0; }
  return do_recv(c, p);
}
ncclResult_t ncclGroupStart() { tl_group_depth++; return 0; }
ncclResult_t ncclGroupEnd()
{
  if (--tl_group_depth > 0) return 0;
  std::vector<std::pair<FakeComm*, Pending>> todo;
  todo.swap(tl_recvs);
  for (auto& r : todo) do_recv(r.first, r.second);
  return 0;
}
const char* ncclGetErrorString(ncclResult_t) { return "fake nccl"; }
}
