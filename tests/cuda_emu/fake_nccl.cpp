// TEST INFRASTRUCTURE: an in-process stand-in for libnccl with the eleven entry points comm.cuh binds, so
// that the library's multi-device code (group contexts, per-rank communicators: scene broadcast, stripe
// gather) can run in the host-emulated build.  "Devices" are just ordinals, memory is host memory.
// Point-to-point: a send copies its payload into the group's mailbox at once (eager), a receive waits for
// the matching (src, dst) entry -- that serves both one thread driving every rank inside ncclGroupStart/End
// (sends are executed before receives at ncclGroupEnd) and one thread per rank.
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

struct emu_stream;
typedef emu_stream *cudaStream_t;
#include "nccl.h"

namespace {
struct Msg { int src, dst; std::vector<unsigned char> data; };
struct Group {
  int n = 0, joined = 0;
  std::mutex m;
  std::condition_variable cv;
  std::deque<Msg> mail;
  // broadcast slot
  unsigned long long bc_gen = 0;
  int bc_left = 0;
  std::vector<unsigned char> bc_data;
  std::vector<unsigned long long> bc_seen;      // per rank: broadcasts consumed so far
};
}  // namespace
struct ncclComm { Group *g; int rank; };

namespace {
size_t dt_size(ncclDataType_t t) {
  switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclFloat16: return 2;
    case ncclInt64: case ncclUint64: return 8;
    default: return 4;
  }
}
struct Op { int kind; ncclComm *c; const void *src; void *dst; size_t bytes; int peer; };   // 0 send, 1 recv, 2 broadcast
thread_local int depth = 0;
thread_local std::vector<Op> pending;
std::mutex g_ids_m;
std::map<std::string, Group *> g_ids;
unsigned long long g_next_id = 1;

void do_send(const Op &o) {
  Group *g = o.c->g;
  Msg m{o.c->rank, o.peer, std::vector<unsigned char>((const unsigned char *)o.src, (const unsigned char *)o.src + o.bytes)};
  std::lock_guard<std::mutex> lk(g->m);
  g->mail.push_back(std::move(m));
  g->cv.notify_all();
}
ncclResult_t do_recv(const Op &o) {
  Group *g = o.c->g;
  std::unique_lock<std::mutex> lk(g->m);
  for (;;) {
    for (auto it = g->mail.begin(); it != g->mail.end(); ++it)
      if (it->src == o.peer && it->dst == o.c->rank) {
        if (it->data.size() != o.bytes) return ncclInvalidArgument;
        std::memcpy(o.dst, it->data.data(), o.bytes);
        g->mail.erase(it);
        return ncclSuccess;
      }
    g->cv.wait(lk);
  }
}
ncclResult_t do_bcast(const Op &o) {
  Group *g = o.c->g;
  std::unique_lock<std::mutex> lk(g->m);
  const int r = o.c->rank;
  if (r == o.peer) {                       // root: wait until the previous broadcast was consumed, then publish
    g->cv.wait(lk, [&] { return g->bc_left == 0; });
    g->bc_data.assign((const unsigned char *)o.src, (const unsigned char *)o.src + o.bytes);
    g->bc_gen += 1;
    g->bc_seen[r] = g->bc_gen;
    g->bc_left = g->n - 1;
    if (o.dst != o.src) std::memcpy(o.dst, o.src, o.bytes);
    g->cv.notify_all();
    return ncclSuccess;
  }
  g->cv.wait(lk, [&] { return g->bc_gen > g->bc_seen[r]; });
  if (g->bc_data.size() != o.bytes) return ncclInvalidArgument;
  std::memcpy(o.dst, g->bc_data.data(), o.bytes);
  g->bc_seen[r] = g->bc_gen;
  g->bc_left -= 1;
  g->cv.notify_all();
  return ncclSuccess;
}
ncclResult_t run(std::vector<Op> &ops) {
  // one thread may hold several ranks' operations: publishers first, then consumers
  ncclResult_t rc = ncclSuccess;
  for (auto &o : ops) if (o.kind == 0) do_send(o);
  for (auto &o : ops) if (o.kind == 2 && o.c->rank == o.peer) { auto r = do_bcast(o); if (r) rc = r; }
  for (auto &o : ops) if (o.kind == 2 && o.c->rank != o.peer) { auto r = do_bcast(o); if (r) rc = r; }
  for (auto &o : ops) if (o.kind == 1) { auto r = do_recv(o); if (r) rc = r; }
  ops.clear();
  return rc;
}
ncclResult_t post(Op o) {
  pending.push_back(o);
  return depth > 0 ? ncclSuccess : run(pending);
}
}  // namespace

extern "C" {
ncclResult_t ncclGetVersion(int *v) { *v = 0; return ncclSuccess; }
const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL error"; }
ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  std::lock_guard<std::mutex> lk(g_ids_m);
  std::memset(id, 0, sizeof *id);
  const unsigned long long k = g_next_id++;
  std::memcpy(id->internal, &k, sizeof k);
  return ncclSuccess;
}
ncclResult_t ncclCommInitAll(ncclComm_t *comms, int ndev, const int *) {
  Group *g = new Group();
  g->n = g->joined = ndev;
  g->bc_seen.assign((size_t)ndev, 0);
  for (int r = 0; r < ndev; ++r) comms[r] = new ncclComm{g, r};
  return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
  Group *g;
  {
    std::lock_guard<std::mutex> lk(g_ids_m);
    Group *&slot = g_ids[std::string(id.internal, sizeof id.internal)];
    if (!slot) { slot = new Group(); slot->n = nranks; slot->bc_seen.assign((size_t)nranks, 0); }
    g = slot;
  }
  std::unique_lock<std::mutex> lk(g->m);
  g->joined += 1;
  g->cv.notify_all();
  g->cv.wait(lk, [&] { return g->joined >= g->n; });      // like the real call: returns when every rank has joined
  *comm = new ncclComm{g, rank};
  return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) { delete c; return ncclSuccess; }     // (the group object is leaked: tests)
ncclResult_t ncclGroupStart(void) { depth += 1; return ncclSuccess; }
ncclResult_t ncclGroupEnd(void) { return --depth == 0 ? run(pending) : ncclSuccess; }
ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t c, cudaStream_t) {
  return post(Op{0, c, buf, nullptr, count * dt_size(dt), peer});
}
ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t dt, int peer, ncclComm_t c, cudaStream_t) {
  return post(Op{1, c, nullptr, buf, count * dt_size(dt), peer});
}
ncclResult_t ncclBroadcast(const void *send, void *recv, size_t count, ncclDataType_t dt, int root, ncclComm_t c, cudaStream_t) {
  return post(Op{2, c, send, recv, count * dt_size(dt), root});
}
}
