// fake_nccl.cpp -- TEST INFRASTRUCTURE.  The handful of NCCL entry points gtsb_dist.cu binds with
// dlsym, for ranks that are THREADS of one process driving the emulated library (tests/test_sim.py):
// a communicator is a rendezvous group; a collective copies every rank's contribution aside, waits
// for all ranks, computes each rank's result from the copies, and waits once more before the copies
// are reused.  Streams are immediate in the emulation, so a collective is complete when it returns.
#include <nccl.h>
#include <string.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

struct Group {
  int world = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t gen = 0;
  std::vector<std::vector<char>> stage;
  uint64_t splits = 0;
  std::string id;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const uint64_t g = gen;
    if (++arrived == world) {
      arrived = 0;
      gen++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};

struct Comm {
  Group *g;
  int rank;
};

std::mutex reg_m;
std::map<std::string, Group *> registry;
uint64_t next_id = 1;

Group *group_of(const std::string &id, int world) {
  std::lock_guard<std::mutex> lk(reg_m);
  Group *&g = registry[id];
  if (g == nullptr) {
    g = new Group();
    g->world = world;
    g->stage.resize(world);
    g->id = id;
  }
  return g;
}

size_t size_of(ncclDataType_t t) {
  switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    case ncclInt64: case ncclUint64: case ncclFloat64: return 8;
    default: return 0;
  }
}

template <typename T>
void reduce(T *out, const std::vector<std::vector<char>> &stage, size_t count, ncclRedOp_t op) {
  for (size_t i = 0; i < count; i++) {
    T acc = reinterpret_cast<const T *>(stage[0].data())[i];
    for (size_t r = 1; r < stage.size(); r++) {
      const T v = reinterpret_cast<const T *>(stage[r].data())[i];
      if (op == ncclSum) acc = (T) (acc + v);
      else if (op == ncclMin) acc = v < acc ? v : acc;
      else if (op == ncclMax) acc = v > acc ? v : acc;
      else if (op == ncclProd) acc = (T) (acc * v);
    }
    out[i] = acc;
  }
}

}  // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  std::lock_guard<std::mutex> lk(reg_m);
  memset(id, 0, sizeof(*id));
  const uint64_t v = next_id++;
  memcpy(id->internal, "cusim", 5);
  memcpy(id->internal + 8, &v, 8);
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return ncclInvalidArgument;
  Comm *c = new Comm{group_of(std::string(id.internal, sizeof(id.internal)), nranks), rank};
  *comm = reinterpret_cast<ncclComm_t>(c);
  c->g->barrier();                         // as the real call: returns once every rank has joined
  return ncclSuccess;
}

// one colour only (every rank of the parent joins the child): what gtsb_dist_init asks for
ncclResult_t ncclCommSplit(ncclComm_t comm, int color, int key, ncclComm_t *newcomm, ncclConfig_t *) {
  Comm *c = reinterpret_cast<Comm *>(comm);
  (void) color;
  (void) key;
  uint64_t n;
  {
    std::lock_guard<std::mutex> lk(c->g->m);
    n = c->g->splits;
  }
  c->g->barrier();                         // every rank has read the same split number
  if (c->rank == 0) {
    std::lock_guard<std::mutex> lk(c->g->m);
    c->g->splits++;
  }
  Comm *d = new Comm{group_of(c->g->id + "/split" + std::to_string(n), c->g->world), c->rank};
  *newcomm = reinterpret_cast<ncclComm_t>(d);
  c->g->barrier();
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  delete reinterpret_cast<Comm *>(comm);   // groups stay registered (a few hundred bytes per test)
  return ncclSuccess;
}

const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "fake NCCL error"; }

ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t comm, cudaStream_t) {
  Comm *c = reinterpret_cast<Comm *>(comm);
  const size_t bytes = count * size_of(t);
  if (size_of(t) == 0) return ncclInvalidArgument;
  c->g->stage[c->rank].assign(static_cast<const char *>(send), static_cast<const char *>(send) + bytes);
  c->g->barrier();
  for (int r = 0; r < c->g->world; r++)
    if (bytes) memcpy(static_cast<char *>(recv) + (size_t) r * bytes, c->g->stage[r].data(), bytes);
  c->g->barrier();
  return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t comm,
                           cudaStream_t) {
  Comm *c = reinterpret_cast<Comm *>(comm);
  const size_t bytes = count * size_of(t);
  if (size_of(t) == 0) return ncclInvalidArgument;
  c->g->stage[c->rank].assign(static_cast<const char *>(send), static_cast<const char *>(send) + bytes);
  c->g->barrier();
  switch (t) {
    case ncclUint8: reduce(static_cast<uint8_t *>(recv), c->g->stage, count, op); break;
    case ncclInt8: reduce(static_cast<int8_t *>(recv), c->g->stage, count, op); break;
    case ncclUint32: reduce(static_cast<uint32_t *>(recv), c->g->stage, count, op); break;
    case ncclInt32: reduce(static_cast<int32_t *>(recv), c->g->stage, count, op); break;
    case ncclUint64: reduce(static_cast<uint64_t *>(recv), c->g->stage, count, op); break;
    case ncclInt64: reduce(static_cast<int64_t *>(recv), c->g->stage, count, op); break;
    case ncclFloat32: reduce(static_cast<float *>(recv), c->g->stage, count, op); break;
    case ncclFloat64: reduce(static_cast<double *>(recv), c->g->stage, count, op); break;
    default: return ncclInvalidArgument;
  }
  c->g->barrier();
  return ncclSuccess;
}

ncclResult_t ncclBroadcast(const void *send, void *recv, size_t count, ncclDataType_t t, int root, ncclComm_t comm,
                           cudaStream_t) {
  Comm *c = reinterpret_cast<Comm *>(comm);
  const size_t bytes = count * size_of(t);
  if (c->rank == root) c->g->stage[root].assign(static_cast<const char *>(send), static_cast<const char *>(send) + bytes);
  c->g->barrier();
  if (bytes) memcpy(recv, c->g->stage[root].data(), bytes);
  c->g->barrier();
  return ncclSuccess;
}

// bound by gtsb_dist.cu but not called on any current path
ncclResult_t ncclSend(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
ncclResult_t ncclRecv(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
ncclResult_t ncclGroupStart() { return ncclSuccess; }
ncclResult_t ncclGroupEnd() { return ncclSuccess; }

}
