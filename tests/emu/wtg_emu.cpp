// TEST INFRASTRUCTURE — host-compiled instantiation of the device state-transition bodies
// (wittgenstein_b200/csrc/wtg_logic.cuh) with a 1-lane coop, driven by the same Engine
// orchestration.  Purpose: debug the exact-order logic of the tick pipeline against the oracle
// on a machine without a GPU.  It is NOT part of the product: the package never loads it, and it
// exports wtgemu_* symbols only.  Scans and the multisplit are plain sequential loops here; the
// CUDA kernels that implement them are validated on the GPU by tests/ -m gpu.
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../wittgenstein_b200/csrc/wtg_engine.hpp"

namespace wtg {

class HostBackend : public Backend {
 public:
  long long launches = 0;
  void* alloc(size_t bytes) override { return std::calloc(1, bytes); }
  void release(void* p) override {
    if (shm.count(p))
      releaseShared(p);
    else
      std::free(p);
  }
  void upload(void* dst, const void* src, size_t bytes) override { std::memcpy(dst, src, bytes); }
  void download(void* dst, const void* src, size_t bytes) override { std::memcpy(dst, src, bytes); }
  void sync() override {}

  // Exchange regions of node-sharded runs live in POSIX shared memory, so that the multi-process form of the sharded engine
  // (one shard per process, handles exchanged through torch.distributed: DistributedGSFSignature / DistributedCasperIMD) can
  // be driven over gloo on a machine without a GPU — the host counterpart of cudaIpcGetMemHandle / cudaIpcOpenMemHandle.
  struct Shm {
    std::string name;
    size_t bytes;
    bool owner;
  };
  std::map<void*, Shm> shm;
  void* allocShared(size_t bytes) override {
    static std::atomic<int> counter{0};
    char name[64];
    std::snprintf(name, sizeof(name), "/wtgemu_%d_%d", (int)getpid(), counter.fetch_add(1));
    int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) throw std::runtime_error("shm_open failed");
    if (ftruncate(fd, (off_t)bytes) != 0) {
      close(fd);
      shm_unlink(name);
      throw std::runtime_error("ftruncate failed");
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);  // zero-filled
    close(fd);
    if (p == MAP_FAILED) {
      shm_unlink(name);
      throw std::runtime_error("mmap failed");
    }
    shm[p] = Shm{name, bytes, true};
    return p;
  }
  void exportShared(void* p, unsigned char* handle) override {  // [0,64) name, [64,72) pid, [72,80) address, [80,88) size
    std::memset(handle, 0, 128);
    auto it = shm.find(p);
    if (it == shm.end()) throw std::runtime_error("not a shared region");
    std::memcpy(handle, it->second.name.c_str(), it->second.name.size() + 1);
    long long pid = (long long)getpid();
    std::memcpy(handle + 64, &pid, sizeof(pid));
    std::memcpy(handle + 72, &p, sizeof(p));
    unsigned long long sz = it->second.bytes;
    std::memcpy(handle + 80, &sz, sizeof(sz));
  }
  void* importShared(const unsigned char* handle) override {
    long long pid;
    std::memcpy(&pid, handle + 64, sizeof(pid));
    if (pid == (long long)getpid()) {  // a shard of this process: same address space
      void* p;
      std::memcpy(&p, handle + 72, sizeof(p));
      return p;
    }
    unsigned long long sz;
    std::memcpy(&sz, handle + 80, sizeof(sz));
    int fd = shm_open(reinterpret_cast<const char*>(handle), O_RDWR, 0600);
    if (fd < 0) throw std::runtime_error("shm_open of a peer's region failed");
    void* p = mmap(nullptr, (size_t)sz, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) throw std::runtime_error("mmap of a peer's region failed");
    shm[p] = Shm{std::string(reinterpret_cast<const char*>(handle)), (size_t)sz, false};
    return p;
  }
  void releaseShared(void* p) {
    auto it = shm.find(p);
    if (it == shm.end()) return;
    munmap(p, it->second.bytes);
    if (it->second.owner) shm_unlink(it->second.name.c_str());
    shm.erase(it);
  }
  ~HostBackend() override {
    while (!shm.empty()) releaseShared(shm.begin()->first);
  }

  void pairScan(const Dev& d, int which) {
    int M = scanCount(d, which);
    Pair run{0, 0};
    for (int j = 0; j < M; ++j) {
      Pair v = scanLoad(d, which, j);
      scanStore(d, which, j, run);
      run.a += v.a;
      run.b += v.b;
    }
    scanTotals(d, which, run);
  }
  // a shard that hit an error still publishes its header (which carries the error) and both signals of the pass, so the
  // other shards stop at once instead of waiting for their time-out
  bool bail(const Dev& d, int phasesDone) {
    if (!d.ctl->error) return false;
    if (d.G > 1) {
      if (phasesDone < 1) {
        xPublishHeader(d);
        xSignal(d, 0);
      }
      if (phasesDone < 2) {
        if (d.allCap > 0) xPublishAllCount(d);
        xSignal(d, 1);
      }
    }
    return true;
  }
  void tick(const Dev& d, int mode) override {
    CoopSerial c;
    std::vector<uint32_t> keep((size_t)std::max(1, d.qcap));
    if (mode == 3) {
      // host-injected sends: control block and descriptors are already in place (Engine::inject)
    } else if (d.ffwd && mode == 1)
      tickBeginFfwd(d, c);
    else if (d.farCap > 0 && !d.ffwd)
      tickBeginFar(d, c, mode);
    else
      tickBegin(d, mode);
    if (bail(d, 0)) return;
    if (d.proto == PROTO_GSF && mode != 3) {
      for (int n = d.n0; n < d.n0 + d.nLoc; ++n)
        if (gsfCondMark(d, n)) gsfCondScanQueue(d, c, n);
      int per = d.workCap / ARENA_STRIPES, tot = stripedTotal(d.ctl->workCnt, per);
      for (int t = 0; t < tot; ++t) gsfScoreItem(d, c, d.workList[stripedIndex(d.ctl->workCnt, per, t)]);
      for (int n = d.n0; n < d.n0 + d.nLoc; ++n)
        if (d.condDue[n]) gsfCondSelect(d, c, n, keep.data());
    }
    if (d.proto == PROTO_HANDEL && mode != 3) {
      HScratch sc;
      for (int n = 0; n < d.N; ++n)
        if (hCondMark(d, n)) hCondScanQueue(d, c, n);
      int per = d.workCap / ARENA_STRIPES, tot = stripedTotal(d.ctl->workCnt, per);
      for (int t = 0; t < tot; ++t) hScoreItem(d, c, d.workList[stripedIndex(d.ctl->workCnt, per, t)]);
      for (int n = 0; n < d.N; ++n) hCondSelect(d, c, n, &sc);
      const char* force = std::getenv("WTG_EMU_FORCE_SERIAL_PICK");
      bool serial = force && force[0] == '1';
      pairScan(d, 2);
      if (!serial)
        for (int n = 0; n < d.N; ++n)
          if (hCondPick(d, n, (u64)d.hDrawBase[n], false) > d.condDraws[n]) serial = true;
      if (!serial) {
        for (int n = 0; n < d.N; ++n) hCondPick(d, n, (u64)d.hDrawBase[n], true);
      } else {
        u64 idx = 0;
        for (int n = 0; n < d.N; ++n) idx += (u64)hCondPick(d, n, idx, true);
      }
    }
    if (mode != 2 && mode != 3) {
      int nEv = d.ctl->nEv;
      const bool coopDispatch = d.allCap > 0;
      for (int i = 0; i < nEv; ++i) coopDispatch ? dispatchCountCoop(d, c, i) : dispatchCount(d, i);
      pairScan(d, 0);
      if (bail(d, 0)) return;
      for (int i = 0; i < nEv; ++i) coopDispatch ? dispatchScatterCoop(d, c, i) : dispatchScatter(d, i);
      for (int n = d.n0; n < d.n0 + d.nLoc; ++n) nodeProcess(d, c, n, 0);
      if (d.proto == PROTO_CASPER && d.cRandomTies && d.ctl->tieCnt > 0 && !d.ctl->error) casperResolveTies(d, c);
      if (d.proto == PROTO_CASPER && d.G == 1 && !d.ctl->error) casperRenumber(d, c);
    }
    pairScan(d, 1);
    if (d.G > 1) {  // node-sharded: exchange 1 (items -> global creation / draw offsets)
      for (int i = 0; i <= d.ctl->nItems; ++i) xPublishItem(d, i);
      xPublishHeader(d);
      xSignal(d, 0);
      for (int q = 0; q < d.G; ++q) xWaitOne(d, 0, q);
      if (!d.ctl->error) {
        for (int i = 0; i < d.ctl->nItems; ++i) xOffsets(d, i);
        xTotals(d);
      }
      if (bail(d, 1)) return;
    }
    if (d.ctl->error) return;
    {
    for (int n = d.n0; n < d.n0 + d.nLoc; ++n) emitCond(d, n);
    if (d.shufCap > 0) {
      int per = d.descCap / ARENA_STRIPES, tot = stripedTotal(d.ctl->descCnt, per);
      for (int t = 0; t < tot; ++t) shuffleCheck(d, stripedIndex(d.ctl->descCnt, per, t));
      shuffleSerial(d);
    }
    {
      int per = d.descCap / ARENA_STRIPES, tot = stripedTotal(d.ctl->descCnt, per);
      for (int t = 0; t < tot; ++t) emitDesc(d, stripedIndex(d.ctl->descCnt, per, t));
    }
    if (d.allCap > 0 && d.G > 1) {  // node-sharded sendAll: publish the descriptors; built after the envelope exchange
      int cnt = std::min(d.ctl->allCnt, d.xAllCap);
      for (int j = 0; j < cnt; ++j) xPublishAll(d, j);
      xPublishAllCount(d);
    } else if (d.allCap > 0) {
      std::vector<int> tmp((size_t)d.N), hist((size_t)ALL_HIST);
      int cnt = std::min(d.ctl->allCnt, d.allCap);
      for (int j = 0; j < cnt; ++j) emitAll(d, c, d.allList[j], tmp.data(), hist.data());
    }
    }
    if (d.G > 1) {  // exchange 2: every shard has stored its envelopes into the destination shards' arrays
      xSignal(d, 1);
      for (int q = 0; q < d.G; ++q) xWaitOne(d, 1, q);
      if (d.ctl->error) return;
      for (int g = 0; g < d.ctl->totalSlots; ++g)
        if (xNeedsIngest(d, g)) xIngest(d, c, g);
      if (d.allCap > 0) {
        std::vector<int> tmp((size_t)d.N), hist((size_t)ALL_HIST);
        int cnt = xAllTotal(d);
        for (int k = 0; k < cnt; ++k) xBuildAll(d, c, k, tmp.data(), hist.data());
      }
    }
    if (d.ctl->error) return;
    // multisplit: stable append into the ring in creation order
    int G = d.ctl->totalSlots;
    for (int g = 0; g < G; ++g) {
      int t = d.newTarget[g];
      if (t < 0) continue;
      if (d.G > 1) d.newTarget[g] = -1;  // the array is indexed by the global creation index: clean for the next pass
      int slot = t & (d.ring - 1);
      int pos = d.bucketCount[slot];
      if (pos >= d.bcap) {
        setError(d, ERR_BUCKET_OVERFLOW, t);
        continue;
      }
      d.buckets[(size_t)slot * d.bcap + pos] = d.newEv[g];
      if (d.G > 1) d.bucketKey[(size_t)slot * d.bcap + pos] = orderKey((unsigned)d.ctl->xseq, (unsigned)g);
      d.bucketCount[slot] = pos + 1;
    }
    {
      int per = d.freeCap / ARENA_STRIPES, tot = stripedTotal(d.ctl->freeCnt, per);
      for (int t = 0; t < tot; ++t) freeApply(d, stripedIndex(d.ctl->freeCnt, per, t));
    }
    tickEnd(d, mode);
    launches += 1;
  }
  void gsfInitNodes(const Dev& d) override {
    for (int n = d.n0; n < d.n0 + d.nLoc; ++n) gsfInitNodeBody(d, n);
  }
  void rngCandidates(const Dev& d, unsigned long long s0, unsigned long long count, int maxBound,
                     std::vector<unsigned long long>& out) override {
    int cap = (int)(count / 256 + 65536);
    std::vector<u64> buf((size_t)cap);
    int cnt = 0;
    rngCandidateChunk(d, s0, 0, count, maxBound, buf.data(), &cnt, cap);
    if (cnt > cap) throw std::runtime_error("rng candidate list overflow");
    out.assign(buf.begin(), buf.begin() + cnt);
  }
  void gsfShufflePeers(const Dev& d, unsigned long long s0, const int* liveRank, const unsigned long long* rejOrd, int nRej) override {
    for (int l = d.L - 1; l >= 1; --l)
      for (int n = d.n0; n < d.n0 + d.nLoc; ++n) {
        if (d.peerBits == 16)
          gsfShuffleLevel<uint16_t>(d, n, l, s0, liveRank, rejOrd, nRej);
        else
          gsfShuffleLevel<uint32_t>(d, n, l, s0, liveRank, rejOrd, nRej);
      }
  }
};

Backend* makeBackend(int) { return new HostBackend(); }
long long backendLaunches(Backend* b) { return static_cast<HostBackend*>(b)->launches; }

}  // namespace wtg

#define WTG_API(name) wtgemu_##name
#include "../../wittgenstein_b200/csrc/wtg_capi.inl"

// debugging aid of the host build only: the entries of the bucket of millisecond t (kind, to, from, meta per entry)
extern "C" int wtgemu_debug_bucket(void* h, int t, unsigned* out, int cap) {
  wtg::Engine& e = static_cast<NetHandle*>(h)->eng;
  int slot = t & (e.d.ring - 1);
  int n = e.d.bucketCount[slot];
  for (int i = 0; i < n && i < cap; ++i) {
    const wtg::Ev& ev = e.d.buckets[(size_t)slot * e.d.bcap + i];
    out[4 * i] = ev.kind;
    out[4 * i + 1] = ev.to;
    out[4 * i + 2] = ev.from;
    out[4 * i + 3] = ev.meta;
  }
  return n;
}
extern "C" int wtgemu_debug_recs(void* h, unsigned from, int* out, int cap) {  // records sent by `from`: idx, n, cur, then (dest, arrival) pairs
  wtg::Engine& e = static_cast<NetHandle*>(h)->eng;
  int k = 0;
  for (int r = 0; r < e.d.recCap && k + 40 < cap; ++r) {
    const wtg::MultiRec& rc = e.d.rec[r];
    if (rc.n == 0 || rc.from != from) continue;
    out[k++] = r;
    out[k++] = (int)rc.n;
    out[k++] = (int)rc.cur;
    for (unsigned i = 0; i < rc.n; ++i) {
      out[k++] = (int)e.d.recDest[rc.off + i];
      out[k++] = e.d.recArrival[rc.off + i];
    }
    out[k++] = -1;
  }
  return k;
}
