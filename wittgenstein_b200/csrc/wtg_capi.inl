// wittgenstein_b200 — C ABI implementation (see include/wtg.h for the contract and the reference
// interface each entry point replaces).  Included by the CUDA backend translation unit with
// WTG_API(name) = wtg_##name; the tests/emu debugging build includes it with another prefix.
#include <cstring>
#include <string>

namespace wtg {
Backend* makeBackend(int device);
long long backendLaunches(Backend* b);
}  // namespace wtg

namespace {
thread_local std::string g_lastError;
struct NetHandle {
  wtg::Engine eng;
  explicit NetHandle(int device = -1) : eng(wtg::makeBackend(device)) {}
};
template <class F>
int guard(F f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_lastError = e.what();
    return -1;
  } catch (...) {
    g_lastError = "unknown error";
    return -1;
  }
}
}  // namespace

extern "C" {

const char* WTG_API(last_error)(void) { return g_lastError.c_str(); }

void* WTG_API(create)(void) {
  try {
    return new NetHandle();
  } catch (const std::exception& e) {
    g_lastError = e.what();
    return nullptr;
  }
}
// a network on a given CUDA device (several networks, one per GPU, may be driven from one process: one caller thread each)
void* WTG_API(create_on)(int device) {
  try {
    return new NetHandle(device);
  } catch (const std::exception& e) {
    g_lastError = e.what();
    return nullptr;
  }
}
void WTG_API(destroy)(void* h) { delete static_cast<NetHandle*>(h); }

#define ENG (static_cast<NetHandle*>(h)->eng)

// ---- node-sharded simulation: shard `rank` of `world` (a power of two <= 8) of ONE network.  Every shard is configured and
//      initialised with identical calls (same seed, builder, latency, protocol parameters), from its own thread or process;
//      after init() the shards exchange the 128-byte handles of their exchange regions and link; then every shard calls
//      run_ms with the same arguments.  Node-indexed read-backs of a shard cover its own ids [n0, n0 + nLoc). ----
void* WTG_API(shard_create)(int rank, int world, int device) {
  try {
    NetHandle* nh = new NetHandle(device);
    try {
      nh->eng.setShard(rank, world);
    } catch (...) {
      delete nh;
      throw;
    }
    return nh;
  } catch (const std::exception& e) {
    g_lastError = e.what();
    return nullptr;
  }
}
int WTG_API(shard_export)(void* h, unsigned char* handle128) {
  return guard([&] {
    ENG.exportExchange(handle128);
    return 0;
  });
}
// handles: world x 128 bytes in rank order (shards of the same process are mapped directly, the others through CUDA IPC)
int WTG_API(shard_link)(void* h, const unsigned char* handles) {
  return guard([&] {
    ENG.linkExchange(handles);
    return 0;
  });
}
int WTG_API(shard_range)(void* h, int* n0, int* nLoc) {
  return guard([&] {
    ENG.requireInited();
    *n0 = ENG.d.n0;
    *nLoc = ENG.d.nLoc;
    return 0;
  });
}
int WTG_API(device)(void* h) { return ENG.be->deviceId(); }

int WTG_API(set_seed)(void* h, long long seed) {
  return guard([&] {
    ENG.setSeed(seed);
    return 0;
  });
}
int WTG_API(set_network_latency)(void* h, const char* name) {
  return guard([&] {
    ENG.requireNotInited();
    ENG.hm.setLatencyByName(name);
    return 0;
  });
}
int WTG_API(set_network_latency_measured)(void* h, const int* proportions, const int* values, int n) {
  return guard([&] {
    ENG.requireNotInited();
    ENG.hm.setLatencyMeasured(proportions, values, n);
    return 0;
  });
}
int WTG_API(set_node_builder)(void* h, const char* name) {
  return guard([&] {
    ENG.requireNotInited();
    ENG.hm.setBuilderByName(name);
    return 0;
  });
}
int WTG_API(set_msg_discard_time)(void* h, int ms) {
  return guard([&] {
    ENG.requireNotInited();
    ENG.msgDiscardTime = ms;
    return 0;
  });
}
// capacities: key in {"bcap","qcap","pool_slots_per_node","desc_cap","rec_cap","ring"}
int WTG_API(set_tunable)(void* h, const char* key, long long v) {
  return guard([&] {
    ENG.requireNotInited();
    std::string k = key;
    if (k == "bcap") ENG.tun.bcap = v;
    else if (k == "qcap") ENG.tun.qcap = v;
    else if (k == "pool_slots_per_node") ENG.tun.poolSlotsPerNode = v;
    else if (k == "desc_cap") ENG.tun.descCap = v;
    else if (k == "rec_cap") ENG.tun.recCap = v;
    else if (k == "ring") ENG.tun.ring = v;
    else if (k == "casper_votes") ENG.tun.casperVotes = v;
    else if (k == "force_shuffle_serial") ENG.forceShufSerial = v != 0;
    else if (k == "casper_blocks") ENG.tun.casperBlocks = v;
    else if (k == "stage_words") ENG.tun.stageWords = v;
    else if (k == "far_cap") ENG.tun.farCap = v;
    else if (k == "peer_bits_32") ENG.tun.peerBits32 = v;
    else throw std::invalid_argument("unknown tunable " + k);
    return 0;
  });
}
int WTG_API(pingpong_init)(void* h, int nodeCt) {
  return guard([&] {
    ENG.pingpongInit(nodeCt);
    return 0;
  });
}
int WTG_API(gsf_init)(void* h, const int* params7) {
  return guard([&] {
    wtg::GsfParams p{params7[0], params7[1], params7[2], params7[3], params7[4], params7[5], params7[6]};
    ENG.gsfInit(p);
    return 0;
  });
}
// params6 = { nodeCount, threshold, pairingTime, signatureSize, replyTimeout, candidateCount }
int WTG_API(sanfermin_construct)(void* h, const int* params6) {
  return guard([&] {
    wtg::SfParams p{params6[0], params6[1], params6[2], params6[3], params6[4], params6[5]};
    ENG.sanferminConstruct(p);
    return 0;
  });
}
int WTG_API(sanfermin_init)(void* h) {
  return guard([&] {
    ENG.sanferminInit();
    return 0;
  });
}
// params6 = { nodeCount, threshold, pairingTime, signatureSize, timeout, candidateCount }
int WTG_API(cappos_init)(void* h, const int* p) {
  return guard([&] {
    wtg::CapposParams cp{p[0], p[1], p[2], p[3], p[4], p[5]};
    ENG.capposInit(cp);
    return 0;
  });
}
// per node: currentPrefixLength, totalNumberOfSigs(-1), done, thresholdDone, isSwapping, mask of cached levels ; thresholdAt
int WTG_API(cappos_node_scalars)(void* h, int* cpl, int* sigs, int* done, int* thrDone, int* swapping, int* cacheMask, long long* thresholdAt) {
  return guard([&] {
    ENG.requireInited();
    if (ENG.d.proto != wtg::PROTO_CAPPOS) throw std::logic_error("not a SanFerminCappos network");
    size_t n = (size_t)ENG.d.N;
    std::vector<int> fl(n), cache(n * 32);
    std::vector<uint32_t> mask(n);
    ENG.fetch(cpl, ENG.d.sfCpl, n);
    ENG.fetch(fl.data(), ENG.d.sfFlags, n);
    ENG.fetch(mask.data(), ENG.d.sfCacheMask, n);
    ENG.fetch(cache.data(), ENG.d.sfCache, n * 32);
    ENG.fetch(thresholdAt, ENG.d.sfThresholdAt, n);
    for (size_t i = 0; i < n; ++i) {
      int s = 1;
      for (int l = 0; l < 32; ++l)
        if ((mask[i] >> l) & 1u) s += cache[i * 32 + (size_t)l];
      sigs[i] = s;
      swapping[i] = fl[i] & 1;
      done[i] = (fl[i] >> 1) & 1;
      thrDone[i] = (fl[i] >> 2) & 1;
      cacheMask[i] = (int)mask[i];
    }
    return 0;
  });
}
// Collections.shuffle(list, rnd) for a java.util.Random whose 48-bit state is `state` (host-side run of the code the emit
// kernel uses for shuffled multi-sends, nextInt's rejection loop included); returns the number of values drawn
int WTG_API(java_shuffle)(unsigned long long state, int n, int* inout) {
  return guard([&] {
    if (n < 0) throw std::invalid_argument("n");
    unsigned long long ja[48], jc[48];
    wtg::lcgJumpTables((uint64_t*)ja, (uint64_t*)jc);
    std::vector<uint32_t> v((size_t)n);
    for (int i = 0; i < n; ++i) v[(size_t)i] = (uint32_t)inout[i];
    int used = wtg::javaShuffleAt((const wtg::u64*)ja, (const wtg::u64*)jc, state & ((1ULL << 48) - 1), 0, v.data(), n);
    for (int i = 0; i < n; ++i) inout[i] = (int)v[(size_t)i];
    return used;
  });
}
// params6 = { cycleLength, randomOnTies, blockProducersCount, attestersPerRound, blockConstructionTime, attestationConstructionTime }
int WTG_API(casper_construct)(void* h, const int* p) {
  return guard([&] {
    wtg::CasperParams cp{p[0], p[1], p[2], p[3], p[4], p[5]};
    ENG.casperConstruct(cp);
    return 0;
  });
}
int WTG_API(casper_init)(void* h, int byzDelay) {
  return guard([&] {
    ENG.casperInit(byzDelay);
    return 0;
  });
}
// kind: 3 ByzBlockProducer, 4 ByzBlockProducerSF, 5 ByzBlockProducerNS, 6 ByzBlockProducerWF (CasperIMD.java:511-707)
int WTG_API(casper_init_byz)(void* h, int kind, int byzDelay) {
  return guard([&] {
    ENG.casperInit(byzDelay, kind);
    return 0;
  });
}
static void requireCasper(wtg::Engine& e) {
  e.requireInited();
  if (e.d.proto != wtg::PROTO_CASPER) throw std::logic_error("not a CasperIMD network");
}
static inline unsigned long long casperAttMix(int attester, int height, int head) {
  return (unsigned long long)(unsigned)attester * 0x9E3779B97F4A7C15ULL + (unsigned long long)(unsigned)height * 0xC2B2AE3D27D4EB4FULL +
         (unsigned long long)(unsigned)head * 0x165667B19E3779F9ULL;
}
int WTG_API(casper_block_count)(void* h) {
  return guard([&] {
    requireCasper(ENG);
    wtg::CasperG g;
    ENG.fetch(&g, ENG.d.cg, 1);
    return g.nBlocks;
  });
}
// per block (index == id, genesis first): height, parent id (-1), producer node id (-1), proposalTime, attestations included
int WTG_API(casper_blocks)(void* h, int* height, int* parent, int* producer, int* proposalTime, int* included) {
  return guard([&] {
    requireCasper(ENG);
    wtg::CasperG g;
    ENG.fetch(&g, ENG.d.cg, 1);
    size_t nb = (size_t)std::min(g.nBlocks, ENG.d.cMaxBlocks);
    ENG.fetch(height, ENG.d.cbHeight, nb);
    ENG.fetch(parent, ENG.d.cbParent, nb);
    ENG.fetch(producer, ENG.d.cbProducer, nb);
    ENG.fetch(proposalTime, ENG.d.cbTime, nb);
    std::vector<unsigned long long> inc(nb * (size_t)ENG.d.cAttWords);
    ENG.fetch(inc.data(), ENG.d.cbIncluded, inc.size());
    for (size_t b = 0; b < nb; ++b) {
      int c = 0;
      for (int w = 0; w < ENG.d.cAttWords; ++w) c += __builtin_popcountll(inc[b * (size_t)ENG.d.cAttWords + (size_t)w]);
      included[b] = c;
    }
    return (int)nb;
  });
}
// the attestations a block includes, as (attester node id, attestation height) pairs in attestation-index order
int WTG_API(casper_block_attestations)(void* h, int block, int* attester, int* height, int cap) {
  return guard([&] {
    requireCasper(ENG);
    if (block < 0 || block >= ENG.d.cMaxBlocks) throw std::invalid_argument("block");
    std::vector<unsigned long long> inc((size_t)ENG.d.cAttWords);
    ENG.fetch(inc.data(), ENG.d.cbIncluded + (size_t)block * ENG.d.cAttWords, inc.size());
    std::vector<int> ah((size_t)ENG.d.cMaxAtts);
    ENG.fetch(ah.data(), ENG.d.attHeight, ah.size());
    int k = 0;
    for (int a = 0; a < ENG.d.cMaxAtts; ++a)
      if ((inc[(size_t)a >> 6] >> (a & 63)) & 1ULL) {
        if (k < cap) {
          attester[k] = ENG.d.cFirstAtt + a % ENG.d.cAttCount;
          height[k] = ah[(size_t)a];
        }
        ++k;
      }
    return k;
  });
}
// per node: head block id, attestations received, distinct heads among them (attestationsByHead.size()), blocks received
// (genesis included), |blocksToReevaluate|, and an order-free hash of the received attestations (attester, height, head)
int WTG_API(casper_node_state)(void* h, int* head, int* attsReceived, int* headsWithAtts, int* blocksReceived, int* toReevaluate,
                               unsigned long long* attHash) {
  return guard([&] {
    requireCasper(ENG);
    const wtg::Dev& d = ENG.d;
    size_t n = (size_t)d.nLoc, o = (size_t)d.n0;  // a shard reports its own nodes
    ENG.fetch(head, d.cHead + o, n);
    std::vector<int> ah((size_t)d.cMaxAtts), ahd((size_t)d.cMaxAtts);
    ENG.fetch(ah.data(), d.attHeight, ah.size());
    ENG.fetch(ahd.data(), d.attHead, ahd.size());
    std::vector<unsigned long long> br(n * (size_t)d.cBlkWords), tr(n * (size_t)d.cBlkWords);
    ENG.fetch(br.data(), d.cBlkRecv + o * (size_t)d.cBlkWords, br.size());
    ENG.fetch(tr.data(), d.cToReeval + o * (size_t)d.cBlkWords, tr.size());
    std::vector<unsigned long long> row((size_t)d.cAttWords);
    std::vector<unsigned char> seen((size_t)d.cMaxBlocks);
    for (size_t i = 0; i < n; ++i) {
      int b = 0, t = 0;
      for (int w = 0; w < d.cBlkWords; ++w) {
        b += __builtin_popcountll(br[i * (size_t)d.cBlkWords + (size_t)w]);
        t += __builtin_popcountll(tr[i * (size_t)d.cBlkWords + (size_t)w]);
      }
      blocksReceived[i] = b;
      toReevaluate[i] = t;
      ENG.fetch(row.data(), d.cAttRecv + (o + i) * (size_t)d.cAttWords, row.size());
      std::fill(seen.begin(), seen.end(), 0);
      int cnt = 0, heads = 0;
      unsigned long long hs = 0;
      for (int w = 0; w < d.cAttWords; ++w) {
        unsigned long long bits = row[(size_t)w];
        while (bits) {
          int a = w * 64 + __builtin_ctzll(bits);
          bits &= bits - 1;
          ++cnt;
          int hb = ahd[(size_t)a];
          if (!seen[(size_t)hb]) {
            seen[(size_t)hb] = 1;
            ++heads;
          }
          hs += casperAttMix(d.cFirstAtt + a % d.cAttCount, ah[(size_t)a], hb);
        }
      }
      attsReceived[i] = cnt;
      headsWithAtts[i] = heads;
      attHash[i] = hs;
    }
    return 0;
  });
}
int WTG_API(casper_heads)(void* h, int* head) {
  return guard([&] {
    requireCasper(ENG);
    ENG.fetch(head, ENG.d.cHead + ENG.d.n0, (size_t)ENG.d.nLoc);
    return 0;
  });
}
// out9 = { toSend, h, late, onTime, delay, onDirectFather, onOlderAncestor, incNotTheBestFather, skipped } of node 1
int WTG_API(casper_byz)(void* h, int* out5) {
  return guard([&] {
    requireCasper(ENG);
    wtg::CasperG g;
    ENG.fetch(&g, ENG.d.cg, 1);
    out5[0] = g.byzToSend;
    out5[1] = g.byzH;
    out5[2] = g.byzLate;
    out5[3] = g.byzOnTime;
    out5[4] = ENG.d.cByzDelay;
    out5[5] = g.byzDirect;
    out5[6] = g.byzOlder;
    out5[7] = g.byzNotBest;
    out5[8] = g.byzSkipped;
    return 0;
  });
}
// params11 = { nodeCount, threshold, pairingTime, levelWaitTime, extraCycle, disseminationPeriodMs, fastPath, nodesDown,
//              desynchronizedStart, byzantineSuicide, hiddenByzantine }
int WTG_API(handel_init)(void* h, const int* p) {
  return guard([&] {
    wtg::HandelParams hp{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10]};
    ENG.handelInit(hp);
    return 0;
  });
}
// network.send(msg, from, to) / network.send(msg, from, dests) (Network.java:353-366) issued by the caller between two windows
int WTG_API(send)(void* h, int type, unsigned long long payload, int from, const int* to, int n) {
  return guard([&] {
    if (n <= 0) return 0;  // send(m, from, emptyList) is a no-op (:354-356)
    wtg::Engine::HostSend hs{from, (uint32_t)type, payload, std::vector<int>(to, to + n)};
    ENG.inject({hs});
    return 0;
  });
}
// network.send(msg, sendTime, from, to) / send(msg, sendTime, from, dests, delaysBetweenMessage) (Network.java:369-382, 420-447)
int WTG_API(send_at)(void* h, int type, unsigned long long payload, int from, const int* to, int n, int sendTime, int delayBetween) {
  return guard([&] {
    if (n <= 0) return 0;
    if (sendTime <= 0) throw std::invalid_argument("sendTime <= time");
    wtg::Engine::HostSend hs{from, (uint32_t)type, payload, std::vector<int>(to, to + n), sendTime, delayBetween};
    ENG.inject({hs});
    return 0;
  });
}
// network.sendAll(msg, from) (Network.java:345-347)
int WTG_API(send_all)(void* h, int type, unsigned long long payload, int from) {
  return guard([&] {
    wtg::Engine::HostSend hs{from, (uint32_t)type, payload, {}};
    ENG.inject({hs});
    return 0;
  });
}
int WTG_API(run_ms)(void* h, int ms) {
  return guard([&] { return ENG.runMs(ms); });
}
int WTG_API(time)(void* h) { return ENG.time; }
int WTG_API(node_count)(void* h) { return ENG.d.N; }
int WTG_API(msgs_size)(void* h) {
  return guard([&] { return ENG.msgsSize(); });
}
int WTG_API(msgs_size_at)(void* h, int t) {
  return guard([&] { return ENG.msgsSizeAt(t); });
}
int WTG_API(peek_messages)(void* h, int* from, int* to, int* sentAt, int* arrivingAt, int* kind, int* msgType, int cap) {
  return guard([&] {
    std::vector<wtg::Engine::PeekRow> rows;
    long long total = ENG.peekMessages(rows, cap < 0 ? 0 : cap);
    for (size_t i = 0; i < rows.size(); ++i) {
      if (from) from[i] = rows[i].from;
      if (to) to[i] = rows[i].to;
      if (sentAt) sentAt[i] = rows[i].sentAt;
      if (arrivingAt) arrivingAt[i] = rows[i].arrivingAt;
      if (kind) kind[i] = rows[i].kind;
      if (msgType) msgType[i] = (int)rows[i].meta;
    }
    return (int)std::min<long long>(total, 0x7fffffffLL);
  });
}
int WTG_API(stop_node)(void* h, int id) {
  return guard([&] {
    ENG.setDown(id, true);
    return 0;
  });
}
int WTG_API(start_node)(void* h, int id) {
  return guard([&] {
    ENG.setDown(id, false);
    return 0;
  });
}
int WTG_API(partition)(void* h, float part) {
  return guard([&] {
    ENG.partition(part);
    return 0;
  });
}
int WTG_API(end_partition)(void* h) {
  return guard([&] {
    ENG.endPartition();
    return 0;
  });
}
unsigned long long WTG_API(rng_state)(void* h) {
  unsigned long long r = 0;
  guard([&] {
    r = ENG.inited ? ENG.readCtl().rng : ENG.hm.rd.seed;
    return 0;
  });
  return r;
}

// ---- read-back -------------------------------------------------------------------------------
// out5N: msgReceived[N], msgSent[N], bytesSent[N], bytesReceived[N], doneAt[N]
int WTG_API(node_counters)(void* h, long long* out5N) {
  return guard([&] {
    ENG.requireInited();
    size_t n = (size_t)ENG.d.nLoc, o = (size_t)ENG.d.n0;  // a shard reports its own nodes: out5N holds 5 x nLoc values
    ENG.fetch(out5N + 0 * n, ENG.d.msgReceived + o, n);
    ENG.fetch(out5N + 1 * n, ENG.d.msgSent + o, n);
    ENG.fetch(out5N + 2 * n, ENG.d.bytesSent + o, n);
    ENG.fetch(out5N + 3 * n, ENG.d.bytesReceived + o, n);
    ENG.fetch(out5N + 4 * n, ENG.d.doneAt + o, n);
    return 0;
  });
}
int WTG_API(node_attrs)(void* h, int* x, int* y, int* extra, int* city, double* speed, unsigned char* down) {
  return guard([&] {
    ENG.requireInited();
    for (int i = 0; i < ENG.d.N; ++i) {
      const wtg::HostNode& n = ENG.hm.nodes[(size_t)i];
      if (x) x[i] = n.x;
      if (y) y[i] = n.y;
      if (extra) extra[i] = n.extra;
      if (city) city[i] = ENG.hm.builder == wtg::HostModel::B_AWS ? n.city : ENG.hm.builder == wtg::HostModel::B_CITIES ? 100 + n.city : -1;  // AWS: region; CITIES: 100 + index of the latency table
      if (speed) speed[i] = n.speed;
      if (down) down[i] = n.down ? 1 : 0;
    }
    return 0;
  });
}
int WTG_API(pingpong_pongs)(void* h, int* out) {
  return guard([&] {
    ENG.requireInited();
    if (ENG.d.proto != wtg::PROTO_PINGPONG) throw std::logic_error("not a PingPong network");
    ENG.fetch(out, ENG.d.pong, (size_t)ENG.d.N);
    return 0;
  });
}
// aggValue, currentPrefixLength, done, thresholdDone, sentRequests, receivedRequests, isSwapping (int[N] each), thresholdAt (int64[N])
int WTG_API(sanfermin_node_scalars)(void* h, int* agg, int* cpl, int* done, int* thrDone, int* sentReq, int* recvReq, int* swapping,
                                    long long* thresholdAt) {
  return guard([&] {
    ENG.requireInited();
    if (ENG.d.proto != wtg::PROTO_SANFERMIN) throw std::logic_error("not a SanFerminSignature network");
    size_t n = (size_t)ENG.d.N;
    std::vector<int> fl(n);
    ENG.fetch(agg, ENG.d.sfAgg, n);
    ENG.fetch(cpl, ENG.d.sfCpl, n);
    ENG.fetch(fl.data(), ENG.d.sfFlags, n);
    ENG.fetch(sentReq, ENG.d.sfSentReq, n);
    ENG.fetch(recvReq, ENG.d.sfRecvReq, n);
    ENG.fetch(thresholdAt, ENG.d.sfThresholdAt, n);
    for (size_t i = 0; i < n; ++i) {
      swapping[i] = fl[i] & 1;
      done[i] = (fl[i] >> 1) & 1;
      thrDone[i] = (fl[i] >> 2) & 1;
    }
    return 0;
  });
}
static void requireHandel(wtg::Engine& e) {
  e.requireInited();
  if (e.d.proto != wtg::PROTO_HANDEL) throw std::logic_error("not a Handel network");
}
// out9N: startAt, nodePairingTime, sigsChecked, sigQueueSize, msgFiltered, currWindowSize, addedCycle, totalSigSize(), sum of toVerifyAgg sizes
int WTG_API(handel_node_scalars)(void* h, int* out9N) {
  return guard([&] {
    requireHandel(ENG);
    size_t n = (size_t)ENG.d.N;
    ENG.fetch(out9N + 0 * n, ENG.d.hStartAt, n);
    ENG.fetch(out9N + 1 * n, ENG.d.pairing, n);
    ENG.fetch(out9N + 2 * n, ENG.d.hSigsChecked, n);
    ENG.fetch(out9N + 3 * n, ENG.d.hSigQueueSize, n);
    ENG.fetch(out9N + 4 * n, ENG.d.hMsgFiltered, n);
    ENG.fetch(out9N + 5 * n, ENG.d.hWindow, n);
    ENG.fetch(out9N + 6 * n, ENG.d.hAddedCycle, n);
    // totalSigSize() = |totalOutgoing(last)| + |totalIncoming(last)| (:349-352) = sum of |totalIncoming| over all levels
    ENG.fetch(out9N + 7 * n, ENG.d.hTotal, n);
    ENG.fetch(out9N + 8 * n, ENG.d.qLen, n);
    return 0;
  });
}
// which: 0 totalIncoming, 1 lastAggVerified, 2 verifiedIndSignatures, 3 toVerifyInd, 4 finishedPeers (unions over levels), 5 blacklist
int WTG_API(handel_rows)(void* h, int which, unsigned long long* outNW) {
  return guard([&] {
    requireHandel(ENG);
    const unsigned long long* src = which == 0 ? ENG.d.hTotInc : which == 1 ? ENG.d.hLastAgg : which == 2 ? ENG.d.hVerInd : which == 3 ? ENG.d.hToVerInd : which == 4 ? ENG.d.hFinPeers : ENG.d.hBlack;
    ENG.fetch(outNW, src, (size_t)ENG.d.N * ENG.d.W64);
    return 0;
  });
}
// N*L arrays: posInLevel, outgoingFinished, suicideBizAfter
int WTG_API(handel_level_scalars)(void* h, int* pos, int* outFin, int* biz) {
  return guard([&] {
    requireHandel(ENG);
    size_t n = (size_t)ENG.d.N * ENG.d.L;
    ENG.fetch(pos, ENG.d.hPos, n);
    ENG.fetch(outFin, ENG.d.hOutFin, n);
    ENG.fetch(biz, ENG.d.hBiz, n);
    return 0;
  });
}
int WTG_API(handel_peers)(void* h, int node, int level, int* out, int cap) {
  return guard([&] {
    requireHandel(ENG);
    if (node < 0 || node >= ENG.d.N || level < 0 || level >= ENG.d.L) throw std::invalid_argument("node/level");
    if (level == 0 || ENG.hm.nodes[(size_t)node].down) return 0;
    int size = 1 << (level - 1);
    int cnt = size < cap ? size : cap;
    std::vector<unsigned> tmp((size_t)cnt);
    ENG.fetch(tmp.data(), (const unsigned*)ENG.d.peers + (size_t)node * (size_t)(ENG.d.N - 1) + (size_t)(size - 1), (size_t)cnt);
    for (int i = 0; i < cnt; ++i) out[i] = (int)tmp[(size_t)i];
    return size;
  });
}
int WTG_API(handel_ranks)(void* h, int node, int* outN) {
  return guard([&] {
    requireHandel(ENG);
    if (node < 0 || node >= ENG.d.N) throw std::invalid_argument("node");
    ENG.fetch(outN, ENG.d.hRanks + (size_t)node * ENG.d.N, (size_t)ENG.d.N);
    return 0;
  });
}
static void requireGsf(wtg::Engine& e) {
  e.requireInited();
  if (e.d.proto != wtg::PROTO_GSF) throw std::logic_error("not a GSFSignature network");
}
int WTG_API(gsf_levels)(void* h) { return ENG.d.L; }
int WTG_API(handel_levels)(void* h) { return ENG.d.L; }
int WTG_API(gsf_verified)(void* h, unsigned long long* outNW) {
  return guard([&] {
    requireGsf(ENG);
    ENG.fetch(outNW, ENG.d.verified + (size_t)ENG.d.n0 * ENG.d.W64, (size_t)ENG.d.nLoc * ENG.d.W64);
    return 0;
  });
}
// which: 0 verified, 1 individualSignatures (seen), 2 indivVerifiedSig
int WTG_API(gsf_rows)(void* h, int which, unsigned long long* outNW) {
  return guard([&] {
    requireGsf(ENG);
    const unsigned long long* src = which == 0 ? ENG.d.verified : which == 1 ? ENG.d.indivSeen : ENG.d.indivVer;
    ENG.fetch(outNW, src + (size_t)ENG.d.n0 * ENG.d.W64, (size_t)ENG.d.nLoc * ENG.d.W64);
    return 0;
  });
}
int WTG_API(gsf_node_scalars)(void* h, int* pairing, int* sigChecked, int* sigQueueSize, int* toVerifySize, int* card) {
  return guard([&] {
    requireGsf(ENG);
    size_t n = (size_t)ENG.d.nLoc, o = (size_t)ENG.d.n0;
    if (pairing) ENG.fetch(pairing, ENG.d.pairing + o, n);
    if (sigChecked) ENG.fetch(sigChecked, ENG.d.sigChecked + o, n);
    if (sigQueueSize) ENG.fetch(sigQueueSize, ENG.d.sigQueueSize + o, n);
    if (toVerifySize) ENG.fetch(toVerifySize, ENG.d.qLen + o, n);
    if (card) ENG.fetch(card, ENG.d.totalCard + o, n);
    return 0;
  });
}
// arrays of N*L, row-major by node
int WTG_API(gsf_level_scalars)(void* h, int* pos, int* remaining, int* card) {
  return guard([&] {
    requireGsf(ENG);
    size_t n = (size_t)ENG.d.nLoc * ENG.d.L, o = (size_t)ENG.d.n0 * ENG.d.L;
    if (pos) ENG.fetch(pos, ENG.d.pos + o, n);
    if (remaining) ENG.fetch(remaining, ENG.d.remaining + o, n);
    if (card) ENG.fetch(card, ENG.d.cntVer + o, n);
    return 0;
  });
}
int WTG_API(gsf_peers)(void* h, int node, int level, int* out, int cap) {
  return guard([&] {
    requireGsf(ENG);
    if (node < 0 || node >= ENG.d.N || level < 0 || level >= ENG.d.L) throw std::invalid_argument("node/level");
    if (node < ENG.d.n0 || node >= ENG.d.n0 + ENG.d.nLoc) throw std::invalid_argument("node belongs to another shard");
    if (level == 0 || ENG.hm.nodes[(size_t)node].down) return 0;
    int size = 1 << (level - 1);
    size_t off = (size_t)node * (size_t)(ENG.d.N - 1) + (size_t)(size - 1);
    int cnt = size < cap ? size : cap;
    if (ENG.d.peerBits == 16) {
      std::vector<unsigned short> tmp((size_t)cnt);
      ENG.fetch(tmp.data(), (const unsigned short*)ENG.d.peers + off, (size_t)cnt);
      int sib = wtg::levelBlock(node ^ (1 << (level - 1)), level).base;
      for (int i = 0; i < cnt; ++i) out[i] = sib + tmp[(size_t)i];
    } else {
      std::vector<unsigned> tmp((size_t)cnt);
      ENG.fetch(tmp.data(), (const unsigned*)ENG.d.peers + off, (size_t)cnt);
      for (int i = 0; i < cnt; ++i) out[i] = (int)tmp[(size_t)i];
    }
    return size;
  });
}
// stats (int64 x 26; the last two: updateWords, reEvaluatedEntries): deliveries, tasks, condRuns, draws, evalEntries, evalWords, updates, cycles, sends,
// multiSends, sendWords, events, maxQueue, maxBucket, maxInbox, recTop, recDestTop, kernelLaunches,
// minPoolFree (over levels), initDraws, ring, bcap, qcap, peerBits
int WTG_API(stats)(void* h, long long* out24) {  // out24 must hold 26 values
  return guard([&] {
    ENG.requireInited();
    wtg::Ctl c = ENG.readCtl();
    long long minFree = -1;
    for (int l = wtg::INLINE_MAX_LEVEL + 1; l < ENG.d.L; ++l)
      if (minFree < 0 || c.poolMinFree[l] < minFree) minFree = c.poolMinFree[l];
    std::vector<unsigned long long> st = ENG.readStats();
    long long v[24] = {(long long)st[wtg::ST_DELIVERIES], (long long)st[wtg::ST_TASKS], (long long)st[wtg::ST_CONDRUNS], (long long)c.statDraws,
                       (long long)st[wtg::ST_EVALENTRIES], (long long)st[wtg::ST_EVALWORDS], (long long)st[wtg::ST_UPDATES],
                       (long long)st[wtg::ST_CYCLES], (long long)st[wtg::ST_SENDS], (long long)st[wtg::ST_MULTISENDS],
                       (long long)st[wtg::ST_SENDWORDS], (long long)c.statEvents, (long long)st[wtg::ST_MAXQUEUE], c.maxBucket,
                       (long long)st[wtg::ST_MAXINBOX], c.recTop, c.recDestTop, wtg::backendLaunches(ENG.be.get()),
                       minFree, (long long)ENG.initDraws, ENG.d.ring, ENG.d.bcap, ENG.d.qcap, ENG.d.peerBits};
    std::memcpy(out24, v, sizeof(v));
    out24[24] = (long long)st[wtg::ST_UPDATEWORDS];
    out24[25] = (long long)st[wtg::ST_EVALPOOL];
    return 0;
  });
}

// device-side stopwatch around whatever is enqueued between the two calls (CUDA events on the engine stream)
int WTG_API(timer_start)(void* h) {
  return guard([&] {
    ENG.be->timerStart();
    return 0;
  });
}
double WTG_API(timer_stop_ms)(void* h) {
  double r = -1.0;
  guard([&] {
    r = ENG.be->timerStopMs();
    return 0;
  });
  return r;
}
// per-kernel event timing (adds a cudaEvent pair around every launch; disables graph replay while on)
int WTG_API(profile_enable)(void* h, int on) {
  return guard([&] {
    ENG.be->profileEnable(on != 0);
    return 0;
  });
}
// returns the number of kernels; ms[i], launches[i] accumulated since enable; names[i] static strings
int WTG_API(profile_read)(void* h, double* ms, long long* launches, const char** names, int cap) {
  return guard([&] { return ENG.be->profileRead(ms, launches, names, cap); });
}

#undef ENG
}  // extern "C"
