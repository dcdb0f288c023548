// ORACLE — TEST INFRASTRUCTURE ONLY.  C interface (ctypes) over the CPU restatement.
// Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
#include <chrono>
#include <cstring>
#include <string>

#include "casper.hpp"
#include "protocols.hpp"

using namespace wo;

static thread_local std::string g_err;
#define WO_TRY try {
#define WO_CATCH(ret)                 \
  }                                   \
  catch (const std::exception& e) {   \
    g_err = e.what();                 \
    return ret;                       \
  }

extern "C" {

const char* wo_last_error() { return g_err.c_str(); }

// ---- JDK primitives ---------------------------------------------------------------------
void wo_random_next_ints(int64_t seed, int n, int32_t* out) {
  JavaRandom r(seed);
  for (int i = 0; i < n; ++i) out[i] = r.nextInt();
}
void wo_random_next_bounded(int64_t seed, int n, int32_t bound, int32_t* out) {
  JavaRandom r(seed);
  for (int i = 0; i < n; ++i) out[i] = r.nextInt(bound);
}
double wo_random_next_double(int64_t seed, int skip) {
  JavaRandom r(seed);
  for (int i = 0; i < skip; ++i) r.nextInt();
  return r.nextDouble();
}
void wo_shuffle(int64_t seed, int n, int32_t* inout) {
  JavaRandom r(seed);
  javaShuffle(inout, n, r);
}
uint64_t wo_lcg_advance(uint64_t seed48, uint64_t n) { return lcgAdvance(seed48, n); }
int wo_pseudo_random(int nodeId, int seed) { return Network::getPseudoRandom(nodeId, seed); }
int32_t wo_string_hash(const char* s) { return javaStringHash(s); }

// AWS builder city order (HashMap iteration order) as a ';'-joined string
int wo_aws_city_order(char* buf, int cap) {
  NodeBuilder nb = makeAwsCityBuilder();
  std::string s;
  for (auto& c : nb.citiesInfo) s += c.name + ";";
  if (static_cast<int>(s.size()) + 1 > cap) return -1;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return static_cast<int>(nb.citiesInfo.size());
}
void wo_aws_cumulative(float* out) {
  NodeBuilder nb = makeAwsCityBuilder();
  for (size_t i = 0; i < nb.citiesInfo.size(); ++i) out[i] = nb.citiesInfo[i].cumulativeProbability;
}

// latency table: latency(kind, param, from attrs, to attrs, delta) without building a network
int wo_latency(const char* latencyName, int fx, int fy, int fextra, const char* fcity, int tx, int ty, int textra,
               const char* tcity, int delta) {
  WO_TRY
  NetworkLatency nl = networkLatencyByName(latencyName);
  Node f, t;
  f.nodeId = 0;
  t.nodeId = 1;
  f.x = fx;
  f.y = fy;
  f.extraLatency = fextra;
  f.cityName = fcity;
  t.x = tx;
  t.y = ty;
  t.extraLatency = textra;
  t.cityName = tcity;
  return nl.getLatency(f, t, delta);
  WO_CATCH(-1000000)
}
double wo_gpd_inverse(double shape, double location, double scale, double y) {
  return GeneralizedParetoDistribution(shape, location, scale).inverseF(y);
}

// ---- PingPong ---------------------------------------------------------------------------
void* wo_pp_create(int nodeCt, const char* nodeBuilderName, const char* networkLatencyName) {
  WO_TRY
  PingPong::Params p;
  p.nodeCt = nodeCt;
  p.nodeBuilderName = nodeBuilderName ? nodeBuilderName : "";
  p.latencyNull = networkLatencyName == nullptr;
  p.networkLatencyName = networkLatencyName ? networkLatencyName : "";
  return new PingPong(p);
  WO_CATCH(nullptr)
}
void wo_pp_destroy(void* h) { delete static_cast<PingPong*>(h); }
void wo_pp_set_seed(void* h, int64_t s) { static_cast<PingPong*>(h)->network.rd.setSeed(s); }
int wo_pp_init(void* h) {
  WO_TRY
  static_cast<PingPong*>(h)->init();
  return 0;
  WO_CATCH(-1)
}
int wo_pp_run_ms(void* h, int ms) {
  WO_TRY
  return static_cast<PingPong*>(h)->network.runMs(ms) ? 1 : 0;
  WO_CATCH(-1)
}
int wo_pp_time(void* h) { return static_cast<PingPong*>(h)->network.time; }
uint64_t wo_pp_rng_state(void* h) { return static_cast<PingPong*>(h)->network.rd.seed; }
int wo_pp_msgs_size(void* h) { return static_cast<PingPong*>(h)->network.msgs.size(); }
void wo_pp_pongs(void* h, int32_t* out) {
  auto* p = static_cast<PingPong*>(h);
  for (size_t i = 0; i < p->nodes.size(); ++i) out[i] = p->nodes[i]->pong;
}
// network.msgs.peekMessages() (Network.java:279-286): rows sorted by (arrivingAt, from, to); returns the total
static int peekRows(const Network& net, int32_t* from, int32_t* to, int32_t* sentAt, int32_t* arrivingAt, int32_t* isTask, int cap) {
  std::vector<EnvelopeInfo> v = net.msgs.peekMessages();
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) {
    from[i] = v[i].from;
    to[i] = v[i].to;
    sentAt[i] = v[i].sentAt;
    arrivingAt[i] = v[i].arrivingAt;
    isTask[i] = v[i].isTask ? 1 : 0;
  }
  return (int)v.size();
}
int wo_pp_peek_messages(void* h, int32_t* from, int32_t* to, int32_t* sentAt, int32_t* arrivingAt, int32_t* isTask, int cap) {
  return peekRows(static_cast<PingPong*>(h)->network, from, to, sentAt, arrivingAt, isTask, cap);
}
static void nodeCounters(const std::vector<Node*>& nodes, int64_t* out5N) {
  size_t n = nodes.size();
  for (size_t i = 0; i < n; ++i) {
    out5N[0 * n + i] = nodes[i]->msgReceived;
    out5N[1 * n + i] = nodes[i]->msgSent;
    out5N[2 * n + i] = nodes[i]->bytesSent;
    out5N[3 * n + i] = nodes[i]->bytesReceived;
    out5N[4 * n + i] = nodes[i]->doneAt;
  }
}
static void nodeAttrs(const std::vector<Node*>& nodes, int32_t* x, int32_t* y, int32_t* extra, int32_t* city, double* speed,
                      uint8_t* down) {
  for (size_t i = 0; i < nodes.size(); ++i) {
    if (x) x[i] = nodes[i]->x;
    if (y) y[i] = nodes[i]->y;
    if (extra) extra[i] = nodes[i]->extraLatency;
    if (city) {  // AWS builder: region index; CITIES builder: 100 + index of the latency table (cf. wtg_node_attrs)
      int reg = nodes[i]->cityIdx < 0 ? -1 : awsRegionOf(nodes[i]->cityName);
      bool aws = false;
      if (reg >= 0)
        for (auto& c : awsCitiesPutOrder()) aws |= c.region == reg && c.mercX == nodes[i]->x && c.mercY == nodes[i]->y;
      city[i] = nodes[i]->cityIdx < 0 ? -1 : aws ? reg : 100 + latencyCityIndex(nodes[i]->cityName);
    }
    if (speed) speed[i] = nodes[i]->speedRatio;
    if (down) down[i] = nodes[i]->down ? 1 : 0;
  }
}
void wo_pp_node_counters(void* h, int64_t* out5N) { nodeCounters(static_cast<PingPong*>(h)->network.allNodes, out5N); }
void wo_pp_node_attrs(void* h, int32_t* x, int32_t* y, int32_t* extra, int32_t* city, double* speed, uint8_t* down) {
  nodeAttrs(static_cast<PingPong*>(h)->network.allNodes, x, y, extra, city, speed, down);
}

// ---- GSFSignature -----------------------------------------------------------------------
void* wo_gsf_create(int nodeCount, int threshold, int pairingTime, int timeoutPerLevelMs, int periodDurationMs,
                    int acceleratedCallsCount, int nodesDown, const char* nodeBuilderName, const char* networkLatencyName) {
  WO_TRY
  GSFSignature::Params p = GSFSignature::makeParams(nodeCount, threshold, pairingTime, timeoutPerLevelMs, periodDurationMs,
                                                    acceleratedCallsCount, nodesDown, nodeBuilderName ? nodeBuilderName : "",
                                                    networkLatencyName ? networkLatencyName : "");
  p.latencyNull = networkLatencyName == nullptr;
  return new GSFSignature(p);
  WO_CATCH(nullptr)
}
void wo_gsf_destroy(void* h) { delete static_cast<GSFSignature*>(h); }
void wo_gsf_set_seed(void* h, int64_t s) { static_cast<GSFSignature*>(h)->network.rd.setSeed(s); }
int wo_gsf_init(void* h) {
  WO_TRY
  static_cast<GSFSignature*>(h)->init();
  return 0;
  WO_CATCH(-1)
}
int wo_gsf_init_fast(void* h, int threads) {
  WO_TRY
  static_cast<GSFSignature*>(h)->initFast(threads);
  return 0;
  WO_CATCH(-1)
}
int wo_gsf_run_ms(void* h, int ms) {
  WO_TRY
  return static_cast<GSFSignature*>(h)->network.runMs(ms) ? 1 : 0;
  WO_CATCH(-1)
}
// run `steps` x runMs(ms); returns wall seconds of the runMs calls only
double wo_gsf_run_timed(void* h, int ms, int steps) {
  auto* p = static_cast<GSFSignature*>(h);
  auto t0 = std::chrono::steady_clock::now();
  try {
    for (int i = 0; i < steps; ++i) p->network.runMs(ms);
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
int wo_gsf_time(void* h) { return static_cast<GSFSignature*>(h)->network.time; }
int wo_gsf_msgs_size(void* h) { return static_cast<GSFSignature*>(h)->network.msgs.size(); }
int64_t wo_gsf_msgs_live(void* h) { return static_cast<GSFSignature*>(h)->network.msgs.live; }
int wo_gsf_peek_messages(void* h, int32_t* from, int32_t* to, int32_t* sentAt, int32_t* arrivingAt, int32_t* isTask, int cap) {
  return peekRows(static_cast<GSFSignature*>(h)->network, from, to, sentAt, arrivingAt, isTask, cap);
}
int wo_gsf_msgs_size_at(void* h, int t) {
  WO_TRY
  return static_cast<GSFSignature*>(h)->network.msgs.sizeAt(t);
  WO_CATCH(-1)
}
int wo_gsf_continue_if(void* h) { return static_cast<GSFSignature*>(h)->continueIf() ? 1 : 0; }
int wo_gsf_levels(void* h, int node) { return static_cast<int>(static_cast<GSFSignature*>(h)->node(node).levels.size()); }
void wo_gsf_node_counters(void* h, int64_t* out5N) { nodeCounters(static_cast<GSFSignature*>(h)->network.allNodes, out5N); }
void wo_gsf_node_attrs(void* h, int32_t* x, int32_t* y, int32_t* extra, int32_t* city, double* speed, uint8_t* down) {
  nodeAttrs(static_cast<GSFSignature*>(h)->network.allNodes, x, y, extra, city, speed, down);
}
// per-node protocol scalars: pairing, sigChecked, sigQueueSize, toVerify.size(), cardinality(verifiedSignatures)
void wo_gsf_node_scalars(void* h, int32_t* pairing, int32_t* sigChecked, int32_t* sigQueueSize, int32_t* toVerifySize,
                         int32_t* card) {
  auto* p = static_cast<GSFSignature*>(h);
  for (size_t i = 0; i < p->nodes.size(); ++i) {
    auto& n = *p->nodes[i];
    if (pairing) pairing[i] = n.nodePairingTime;
    if (sigChecked) sigChecked[i] = n.sigChecked;
    if (sigQueueSize) sigQueueSize[i] = n.sigQueueSize;
    if (toVerifySize) toVerifySize[i] = static_cast<int>(n.toVerify.size());
    if (card) card[i] = n.verifiedSignatures.cardinality();
  }
}
// verifiedSignatures of every node as N rows of `words` uint64 (little-endian bit order: bit i of the set = bit i%64 of word i/64)
void wo_gsf_verified(void* h, uint64_t* out, int words) {
  auto* p = static_cast<GSFSignature*>(h);
  for (size_t i = 0; i < p->nodes.size(); ++i)
    for (int w = 0; w < words; ++w) out[i * static_cast<size_t>(words) + static_cast<size_t>(w)] = p->nodes[i]->verifiedSignatures.rawWord(w);
}
// which: 0 = verifiedSignatures of the level, 1 = individualSignatures, 2 = indivVerifiedSig, 3 = waitedSigs ; OR over levels
void wo_gsf_level_rows(void* h, int which, uint64_t* out, int words) {
  auto* p = static_cast<GSFSignature*>(h);
  for (size_t i = 0; i < p->nodes.size(); ++i) {
    uint64_t* row = out + i * static_cast<size_t>(words);
    for (int w = 0; w < words; ++w) row[w] = 0;
    for (auto& l : p->nodes[i]->levels) {
      const JBitSet& b = which == 0 ? l.verifiedSignatures : which == 1 ? l.individualSignatures : which == 2 ? l.indivVerifiedSig : l.waitedSigs;
      for (int w = 0; w < words; ++w) row[w] |= b.rawWord(w);
    }
  }
}
// per (node, level): posInLevel, remainingCalls, cardinality(level.verifiedSignatures); arrays of N*L (row-major by node)
void wo_gsf_level_scalars(void* h, int L, int32_t* pos, int32_t* remaining, int32_t* card) {
  auto* p = static_cast<GSFSignature*>(h);
  for (size_t i = 0; i < p->nodes.size(); ++i)
    for (int l = 0; l < L; ++l) {
      size_t k = i * static_cast<size_t>(L) + static_cast<size_t>(l);
      auto& n = *p->nodes[i];
      bool has = l < static_cast<int>(n.levels.size());
      if (pos) pos[k] = has ? n.levels[static_cast<size_t>(l)].posInLevel : 0;
      if (remaining) remaining[k] = has ? n.levels[static_cast<size_t>(l)].remainingCalls : 0;
      if (card) card[k] = has ? n.levels[static_cast<size_t>(l)].verifiedSignatures.cardinality() : 0;
    }
}
int wo_gsf_peers(void* h, int node, int level, int32_t* out, int cap) {
  auto* p = static_cast<GSFSignature*>(h);
  auto& n = p->node(node);
  if (level >= static_cast<int>(n.levels.size())) return 0;
  auto& pe = n.levels[static_cast<size_t>(level)].peers;
  int c = std::min<int>(cap, static_cast<int>(pe.size()));
  for (int i = 0; i < c; ++i) out[i] = static_cast<int32_t>(pe[static_cast<size_t>(i)]);
  return static_cast<int>(pe.size());
}
uint64_t wo_gsf_rng_state(void* h) { return static_cast<GSFSignature*>(h)->network.rd.seed; }
// stats: [deliveries, tasks, condRuns, draws, evalEntries, evalBytes, updates, cycles, sends, multiSends, sendBytes, maxQueue]
void wo_gsf_stats(void* h, int64_t* out12) {
  auto* p = static_cast<GSFSignature*>(h);
  out12[0] = p->network.statDeliveries;
  out12[1] = p->network.statTasks;
  out12[2] = p->network.statCondRuns;
  out12[3] = p->network.statDraws;
  out12[4] = p->statEvalEntries;
  out12[5] = p->statEvalBytes;
  out12[6] = p->statUpdates;
  out12[7] = p->statCycles;
  out12[8] = p->statSends;
  out12[9] = p->statMultiSends;
  out12[10] = p->statSendBytes;
  out12[11] = p->statMaxQueue;
}
// test hook used by the restated PT/GSFSignatureTest.testGetLastFinishedLevel
int wo_gsf_test_last_finished(void* h) {
  auto* p = static_cast<GSFSignature*>(h);
  auto& n0 = p->node(0);
  int r0 = n0.getLastFinishedLevel().cardinality();
  n0.levels[1].verifiedSignatures.or_(n0.levels[1].waitedSigs);
  int r1 = n0.getLastFinishedLevel().cardinality();
  n0.levels[2].verifiedSignatures.set(2);
  int r2 = n0.getLastFinishedLevel().cardinality();
  n0.levels[2].verifiedSignatures.set(3);
  int r3 = n0.getLastFinishedLevel().cardinality();
  return r0 * 1000 + r1 * 100 + r2 * 10 + r3;
}


// ---- SanFerminSignature -------------------------------------------------------------------
void* wo_sf_create(int nodeCount, int threshold, int pairingTime, int signatureSize, int replyTimeout, int candidateCount,
                   const char* nodeBuilderName, const char* networkLatencyName) {
  WO_TRY
  SanFerminSignature::Params p;
  p.nodeCount = nodeCount;
  p.threshold = threshold;
  p.pairingTime = pairingTime;
  p.signatureSize = signatureSize;
  p.replyTimeout = replyTimeout;
  p.candidateCount = candidateCount;
  p.nodeBuilderName = nodeBuilderName ? nodeBuilderName : "";
  p.latencyNull = networkLatencyName == nullptr;
  p.networkLatencyName = networkLatencyName ? networkLatencyName : "";
  return new SanFerminSignature(p);
  WO_CATCH(nullptr)
}
void wo_sf_destroy(void* h) { delete static_cast<SanFerminSignature*>(h); }
void wo_sf_set_seed(void* h, int64_t s) { static_cast<SanFerminSignature*>(h)->network.rd.setSeed(s); }
int wo_sf_init(void* h) {
  WO_TRY
  static_cast<SanFerminSignature*>(h)->init();
  return 0;
  WO_CATCH(-1)
}
int wo_sf_run_ms(void* h, int ms) {
  WO_TRY
  return static_cast<SanFerminSignature*>(h)->network.runMs(ms) ? 1 : 0;
  WO_CATCH(-1)
}
int wo_sf_time(void* h) { return static_cast<SanFerminSignature*>(h)->network.time; }
int64_t wo_sf_msgs_live(void* h) { return static_cast<SanFerminSignature*>(h)->network.msgs.live; }
uint64_t wo_sf_rng_state(void* h) { return static_cast<SanFerminSignature*>(h)->network.rd.seed; }
void wo_sf_node_counters(void* h, int64_t* out5N) { nodeCounters(static_cast<SanFerminSignature*>(h)->network.allNodes, out5N); }
void wo_sf_node_attrs(void* h, int32_t* x, int32_t* y, int32_t* extra, int32_t* city, double* speed, uint8_t* down) {
  nodeAttrs(static_cast<SanFerminSignature*>(h)->network.allNodes, x, y, extra, city, speed, down);
}
// per node: aggValue, currentPrefixLength, done, thresholdDone, sentRequests, receivedRequests, isSwapping ; thresholdAt (int64)
void wo_sf_node_scalars(void* h, int32_t* agg, int32_t* cpl, int32_t* done, int32_t* thrDone, int32_t* sentReq, int32_t* recvReq,
                        int32_t* swapping, int64_t* thresholdAt) {
  auto* p = static_cast<SanFerminSignature*>(h);
  for (size_t i = 0; i < p->nodes.size(); ++i) {
    auto& n = *p->nodes[i];
    agg[i] = n.aggValue;
    cpl[i] = n.currentPrefixLength;
    done[i] = n.done ? 1 : 0;
    thrDone[i] = n.thresholdDone ? 1 : 0;
    sentReq[i] = n.sentRequests;
    recvReq[i] = n.receivedRequests;
    swapping[i] = n.isSwapping ? 1 : 0;
    thresholdAt[i] = n.thresholdAt;
  }
}
// SanFerminHelper KATs (PT/SanFerminTest.java): candidate / own set of a node, pickNextNodes
void wo_sf_helper_sets(int nodeId, int setSize, int level, int32_t* out4) {
  SanFerminHelper h(nodeId, setSize, nullptr);
  auto c = h.getCandidateSet(level);
  auto o = h.getOwnSet(level);
  out4[0] = c.first;
  out4[1] = c.second;
  out4[2] = o.first;
  out4[3] = o.second;
}
int wo_sf_helper_pick(int nodeId, int setSize, int level, int howMany, int calls, int32_t* out, int cap) {
  JavaRandom rd(0);
  SanFerminHelper h(nodeId, setSize, &rd);
  std::vector<int> r;
  for (int i = 0; i < calls; ++i) r = h.pickNextNodes(level, howMany);
  int n = std::min<int>(cap, static_cast<int>(r.size()));
  for (int i = 0; i < n; ++i) out[i] = r[static_cast<size_t>(i)];
  return static_cast<int>(r.size());
}


// ---- Handel -------------------------------------------------------------------------------
// params10 = nodeCount, threshold, pairingTime, levelWaitTime, extraCycle, disseminationPeriodMs, fastPath, nodesDown,
//            desynchronizedStart, byzantineSuicide
void* wo_handel_create(const int* params10,  /* 11 values: the 11th is hiddenByzantine */ const char* nodeBuilderName, const char* networkLatencyName) {
  WO_TRY
  Handel::Params p;
  p.nodeCount = params10[0];
  p.threshold = params10[1];
  p.pairingTime = params10[2];
  p.levelWaitTime = params10[3];
  p.extraCycle = params10[4];
  p.disseminationPeriodMs = params10[5];
  p.fastPath = params10[6];
  p.nodesDown = params10[7];
  p.desynchronizedStart = params10[8];
  p.byzantineSuicide = params10[9] != 0;
  p.hiddenByzantine = params10[10] != 0;
  p.nodeBuilderName = nodeBuilderName ? nodeBuilderName : "";
  p.latencyNull = networkLatencyName == nullptr;
  p.networkLatencyName = networkLatencyName ? networkLatencyName : "";
  return new Handel(p);
  WO_CATCH(nullptr)
}
void wo_handel_destroy(void* h) { delete static_cast<Handel*>(h); }
void wo_handel_set_seed(void* h, int64_t s) { static_cast<Handel*>(h)->network.rd.setSeed(s); }
int wo_handel_init(void* h) {
  WO_TRY
  static_cast<Handel*>(h)->init();
  return 0;
  WO_CATCH(-1)
}
int wo_handel_run_ms(void* h, int ms) {
  WO_TRY
  return static_cast<Handel*>(h)->network.runMs(ms) ? 1 : 0;
  WO_CATCH(-1)
}
double wo_handel_run_timed(void* h, int ms, int steps) {
  auto* p = static_cast<Handel*>(h);
  auto t0 = std::chrono::steady_clock::now();
  try {
    for (int i = 0; i < steps; ++i) p->network.runMs(ms);
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
int wo_handel_time(void* h) { return static_cast<Handel*>(h)->network.time; }
int64_t wo_handel_msgs_live(void* h) { return static_cast<Handel*>(h)->network.msgs.live; }
uint64_t wo_handel_rng_state(void* h) { return static_cast<Handel*>(h)->network.rd.seed; }
int wo_handel_continue_if(void* h) { return static_cast<Handel*>(h)->continueIf() ? 1 : 0; }
int wo_handel_levels(void* h) { return static_cast<int>(static_cast<Handel*>(h)->node(0).levels.size()); }
void wo_handel_node_counters(void* h, int64_t* out5N) { nodeCounters(static_cast<Handel*>(h)->network.allNodes, out5N); }
void wo_handel_node_attrs(void* h, int32_t* x, int32_t* y, int32_t* extra, int32_t* city, double* speed, uint8_t* down) {
  nodeAttrs(static_cast<Handel*>(h)->network.allNodes, x, y, extra, city, speed, down);
}
// per node: startAt, pairing, sigsChecked, sigQueueSize, msgFiltered, currWindowSize, addedCycle, totalSigSize, sum of toVerifyAgg sizes
void wo_handel_node_scalars(void* h, int32_t* out9N) {
  auto* p = static_cast<Handel*>(h);
  size_t n = p->nodes.size();
  for (size_t i = 0; i < n; ++i) {
    auto& nd = *p->nodes[i];
    int q = 0;
    for (auto& l : nd.levels) q += static_cast<int>(l.toVerifyAgg.size());
    int v[9] = {nd.startAt, nd.nodePairingTime, nd.sigsChecked, nd.sigQueueSize, nd.msgFiltered, nd.currWindowSize, nd.addedCycle, nd.totalSigSize(), q};
    for (int k = 0; k < 9; ++k) out9N[static_cast<size_t>(k) * n + i] = v[k];
  }
}
// which: 0 totalIncoming, 1 lastAggVerified, 2 verifiedIndSignatures, 3 toVerifyInd, 4 finishedPeers (OR over levels), 5 blacklist
void wo_handel_rows(void* h, int which, uint64_t* out, int words) {
  auto* p = static_cast<Handel*>(h);
  for (size_t i = 0; i < p->nodes.size(); ++i) {
    uint64_t* row = out + i * static_cast<size_t>(words);
    for (int w = 0; w < words; ++w) row[w] = 0;
    if (which == 5) {
      for (int w = 0; w < words; ++w) row[w] = p->nodes[i]->blacklist.rawWord(w);
      continue;
    }
    for (auto& l : p->nodes[i]->levels) {
      const JBitSet& b = which == 0 ? l.totalIncoming : which == 1 ? l.lastAggVerified : which == 2 ? l.verifiedIndSignatures : which == 3 ? l.toVerifyInd : l.finishedPeers;
      for (int w = 0; w < words; ++w) row[w] |= b.rawWord(w);
    }
  }
}
// per (node, level): posInLevel, outgoingFinished, suicideBizAfter, toVerifyAgg.size()
void wo_handel_level_scalars(void* h, int L, int32_t* pos, int32_t* fin, int32_t* biz, int32_t* qsz) {
  auto* p = static_cast<Handel*>(h);
  for (size_t i = 0; i < p->nodes.size(); ++i)
    for (int l = 0; l < L; ++l) {
      size_t k = i * static_cast<size_t>(L) + static_cast<size_t>(l);
      auto& lv = p->nodes[i]->levels[static_cast<size_t>(l)];
      pos[k] = lv.posInLevel;
      fin[k] = lv.outgoingFinished ? 1 : 0;
      biz[k] = lv.suicideBizAfter;
      qsz[k] = static_cast<int>(lv.toVerifyAgg.size());
    }
}
int wo_handel_peers(void* h, int node, int level, int32_t* out, int cap) {
  auto& pe = static_cast<Handel*>(h)->node(node).levels[static_cast<size_t>(level)].peers;
  int c = std::min<int>(cap, static_cast<int>(pe.size()));
  for (int i = 0; i < c; ++i) out[i] = pe[static_cast<size_t>(i)];
  return static_cast<int>(pe.size());
}
void wo_handel_ranks(void* h, int node, int32_t* out) {
  auto& r = static_cast<Handel*>(h)->node(node).receptionRanks;
  for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
}


// ---- CasperIMD ------------------------------------------------------------------------------
// params6 = { cycleLength, randomOnTies, blockProducersCount, attestersPerRound, blockConstructionTime, attestationConstructionTime }
void* wo_casper_create(const int* p, const char* nodeBuilderName, const char* networkLatencyName) {
  WO_TRY
  return new CasperIMD(CasperIMD::makeParams(p[0], p[1] != 0, p[2], p[3], p[4], p[5], nodeBuilderName ? nodeBuilderName : "",
                                             networkLatencyName ? networkLatencyName : "", networkLatencyName == nullptr));
  WO_CATCH(nullptr)
}
void wo_casper_destroy(void* h) { delete static_cast<CasperIMD*>(h); }
void wo_casper_set_seed(void* h, int64_t s) { static_cast<CasperIMD*>(h)->network.rd.setSeed(s); }
int wo_casper_init(void* h, int byzDelay) {
  WO_TRY
  auto* ci = static_cast<CasperIMD*>(h);
  ci->init(ci->newByzWF(byzDelay));
  return 0;
  WO_CATCH(-1)
}
int wo_casper_init_byz(void* h, int kind, int byzDelay) {  // 3 plain, 4 SF, 5 NS, 6 WF
  WO_TRY
  auto* ci = static_cast<CasperIMD*>(h);
  CasperIMD::ByzBlockProducer* b = kind == 3 ? ci->newByz(byzDelay) : kind == 4 ? static_cast<CasperIMD::ByzBlockProducer*>(ci->newByzSF(byzDelay))
                                   : kind == 5 ? static_cast<CasperIMD::ByzBlockProducer*>(ci->newByzNS(byzDelay))
                                               : static_cast<CasperIMD::ByzBlockProducer*>(ci->newByzWF(byzDelay));
  ci->init(b);
  return 0;
  WO_CATCH(-1)
}
int wo_casper_run_ms(void* h, int ms) {
  WO_TRY
  return static_cast<CasperIMD*>(h)->network.runMs(ms) ? 1 : 0;
  WO_CATCH(-1)
}
double wo_casper_run_timed(void* h, int ms, int step) {  // wall seconds of `ms` simulated ms in runMs(step) slices
  WO_TRY
  auto* ci = static_cast<CasperIMD*>(h);
  auto t0 = std::chrono::steady_clock::now();
  for (int done = 0; done < ms; done += step) ci->network.runMs(std::min(step, ms - done));
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  WO_CATCH(-1.0)
}
int wo_casper_time(void* h) { return static_cast<CasperIMD*>(h)->network.time; }
int wo_casper_node_count(void* h) { return static_cast<int>(static_cast<CasperIMD*>(h)->network.allNodes.size()); }
int64_t wo_casper_msgs_live(void* h) { return static_cast<CasperIMD*>(h)->network.msgs.live; }
int wo_casper_peek_messages(void* h, int32_t* from, int32_t* to, int32_t* sentAt, int32_t* arrivingAt, int32_t* isTask, int cap) {
  return peekRows(static_cast<CasperIMD*>(h)->network, from, to, sentAt, arrivingAt, isTask, cap);
}
int wo_casper_msgs_size_at(void* h, int t) { return static_cast<CasperIMD*>(h)->network.msgs.sizeAt(t); }
uint64_t wo_casper_rng_state(void* h) { return static_cast<CasperIMD*>(h)->network.rd.seed; }
int64_t wo_casper_deliveries(void* h) { return static_cast<CasperIMD*>(h)->network.statDeliveries; }
void wo_casper_node_counters(void* h, int64_t* out5N) { nodeCounters(static_cast<CasperIMD*>(h)->network.allNodes, out5N); }
void wo_casper_node_attrs(void* h, int32_t* x, int32_t* y, int32_t* extra, int32_t* city, double* speed, uint8_t* down) {
  nodeAttrs(static_cast<CasperIMD*>(h)->network.allNodes, x, y, extra, city, speed, down);
}
void wo_casper_stop_node(void* h, int id) { static_cast<CasperIMD*>(h)->network.getNodeById(id).stop(); }
void wo_casper_start_node(void* h, int id) { static_cast<CasperIMD*>(h)->network.getNodeById(id).start(); }
int wo_casper_partition(void* h, float part) {
  WO_TRY
  static_cast<CasperIMD*>(h)->network.partition(part);
  return 0;
  WO_CATCH(-1)
}
int wo_casper_block_count(void* h) { return 1 + static_cast<int>(static_cast<CasperIMD*>(h)->blocks.size()); }
// blocks in id order, genesis first; ids are reported relative to the instance (genesis 0, first block 1)
void wo_casper_blocks(void* h, int32_t* height, int32_t* parent, int32_t* producer, int32_t* proposalTime, int32_t* included) {
  auto* ci = static_cast<CasperIMD*>(h);
  height[0] = 0;
  parent[0] = -1;
  producer[0] = -1;
  proposalTime[0] = 0;
  included[0] = 0;
  for (size_t i = 0; i < ci->blocks.size(); ++i) {
    auto& b = *ci->blocks[i];
    size_t k = static_cast<size_t>(b.id);
    height[k] = b.height;
    parent[k] = static_cast<int32_t>(b.parent->id);
    producer[k] = b.producer->nodeId;
    proposalTime[k] = b.proposalTime;
    int c = 0;
    for (auto& kv : b.attestationsByHeight) c += static_cast<int>(kv.second.size());
    included[k] = c;
  }
}
int wo_casper_block_attestations(void* h, int block, int32_t* attester, int32_t* height, int cap) {
  auto* ci = static_cast<CasperIMD*>(h);
  if (block <= 0) return 0;
  auto& b = *ci->blocks.at(static_cast<size_t>(block - 1));
  int k = 0;
  for (auto& kv : b.attestationsByHeight)
    for (auto* a : kv.second) {
      if (k < cap) {
        attester[k] = a->attester->nodeId;
        height[k] = a->height;
      }
      ++k;
    }
  return k;
}
static inline uint64_t casperAttMix(int attester, int height, int head) {
  return static_cast<uint64_t>(static_cast<uint32_t>(attester)) * 0x9E3779B97F4A7C15ULL +
         static_cast<uint64_t>(static_cast<uint32_t>(height)) * 0xC2B2AE3D27D4EB4FULL +
         static_cast<uint64_t>(static_cast<uint32_t>(head)) * 0x165667B19E3779F9ULL;
}
void wo_casper_node_state(void* h, int32_t* head, int32_t* attsReceived, int32_t* headsWithAtts, int32_t* blocksReceived,
                          int32_t* toReevaluate, uint64_t* attHash) {
  auto* ci = static_cast<CasperIMD*>(h);
  for (size_t i = 0; i < ci->network.allNodes.size(); ++i) {
    auto* n = static_cast<CasperIMD::CasperNode*>(ci->network.allNodes[i]);
    head[i] = static_cast<int32_t>(n->head->id);
    int cnt = 0;
    uint64_t hs = 0;
    for (auto& kv : n->attestationsByHead)
      for (auto* a : kv.second) {
        ++cnt;
        hs += casperAttMix(a->attester->nodeId, a->height, static_cast<int>(a->head->id));
      }
    attsReceived[i] = cnt;
    headsWithAtts[i] = static_cast<int32_t>(n->attestationsByHead.size());
    blocksReceived[i] = static_cast<int32_t>(n->blocksReceivedByBlockId.size());
    toReevaluate[i] = static_cast<int32_t>(n->blocksToReevaluate.size());
    attHash[i] = hs;
  }
}
void wo_casper_byz(void* h, int32_t* out5) {
  auto* ci = static_cast<CasperIMD*>(h);
  auto* b = static_cast<CasperIMD::ByzBlockProducer*>(ci->bps.at(0));
  auto* wf = dynamic_cast<CasperIMD::ByzBlockProducerWF*>(b);
  auto* ns = dynamic_cast<CasperIMD::ByzBlockProducerNS*>(b);
  out5[0] = b->toSend;
  out5[1] = b->h;
  out5[2] = wf ? wf->late : 0;
  out5[3] = wf ? wf->onTime : 0;
  out5[4] = b->delay;
  out5[5] = b->onDirectFather;
  out5[6] = b->onOlderAncestor;
  out5[7] = b->incNotTheBestFather;
  out5[8] = ns ? ns->skipped : 0;
}

// ---- network controls of any protocol (Node.stop/start, Network.partition/endPartition/setMsgDiscardTime) ----------
// op: 0 stop(arg), 1 start(arg), 2 partition(arg / 10000.f), 3 endPartition, 4 setMsgDiscardTime(arg)
static int netCtl(Network& net, int op, int arg) {
  WO_TRY
  switch (op) {
    case 0: net.getNodeById(arg).stop(); break;
    case 1: net.getNodeById(arg).start(); break;
    case 2: net.partition(static_cast<float>(arg) / 10000.f); break;
    case 3: net.endPartition(); break;
    case 4: net.setMsgDiscardTime(arg); break;
    default: throw IllegalArgument("op");
  }
  return 0;
  WO_CATCH(-1)
}
// network.setNetworkLatency(distribProp, distribVal) — MeasuredNetworkLatency (Network.java:665-667), after construction, before init()
static int setMeasured(Network& net, const int* props, const int* vals, int n) {
  WO_TRY
  net.setNetworkLatency(NetworkLatency::measured(std::vector<int>(props, props + n), std::vector<int>(vals, vals + n)));
  return 0;
  WO_CATCH(-1)
}
int wo_pp_set_latency_measured(void* h, const int* p, const int* v, int n) { return setMeasured(static_cast<PingPong*>(h)->network, p, v, n); }
int wo_gsf_set_latency_measured(void* h, const int* p, const int* v, int n) { return setMeasured(static_cast<GSFSignature*>(h)->network, p, v, n); }
int wo_pp_net_ctl(void* h, int op, int arg) { return netCtl(static_cast<PingPong*>(h)->network, op, arg); }
int wo_gsf_net_ctl(void* h, int op, int arg) { return netCtl(static_cast<GSFSignature*>(h)->network, op, arg); }
int wo_sf_net_ctl(void* h, int op, int arg) { return netCtl(static_cast<SanFerminSignature*>(h)->network, op, arg); }
int wo_handel_net_ctl(void* h, int op, int arg) { return netCtl(static_cast<Handel*>(h)->network, op, arg); }

// ---- SanFerminCappos ---------------------------------------------------------------------------
// params6 = { nodeCount, threshold, pairingTime, signatureSize, timeout, candidateCount }  (the constructor's order, :86-94)
void* wo_cappos_create(const int* p6, const char* nodeBuilderName, const char* networkLatencyName) {
  WO_TRY
  SanFerminCappos::Params p;
  p.nodeCount = p6[0];
  p.threshold = p6[1];
  p.pairingTime = p6[2];
  p.signatureSize = p6[3];
  p.timeout = p6[4];
  p.candidateCount = p6[5];
  p.nodeBuilderName = nodeBuilderName ? nodeBuilderName : "";
  p.latencyNull = networkLatencyName == nullptr;
  p.networkLatencyName = networkLatencyName ? networkLatencyName : "";
  return new SanFerminCappos(p);
  WO_CATCH(nullptr)
}
void wo_cappos_destroy(void* h) { delete static_cast<SanFerminCappos*>(h); }
void wo_cappos_set_seed(void* h, int64_t s) { static_cast<SanFerminCappos*>(h)->network.rd.setSeed(s); }
int wo_cappos_init(void* h) {
  WO_TRY
  static_cast<SanFerminCappos*>(h)->init();
  return 0;
  WO_CATCH(-1)
}
int wo_cappos_run_ms(void* h, int ms) {
  WO_TRY
  return static_cast<SanFerminCappos*>(h)->network.runMs(ms) ? 1 : 0;
  WO_CATCH(-1)
}
int wo_cappos_time(void* h) { return static_cast<SanFerminCappos*>(h)->network.time; }
int64_t wo_cappos_msgs_live(void* h) { return static_cast<SanFerminCappos*>(h)->network.msgs.live; }
uint64_t wo_cappos_rng_state(void* h) { return static_cast<SanFerminCappos*>(h)->network.rd.seed; }
void wo_cappos_node_counters(void* h, int64_t* out5N) { nodeCounters(static_cast<SanFerminCappos*>(h)->network.allNodes, out5N); }
void wo_cappos_node_attrs(void* h, int32_t* x, int32_t* y, int32_t* extra, int32_t* city, double* speed, uint8_t* down) {
  nodeAttrs(static_cast<SanFerminCappos*>(h)->network.allNodes, x, y, extra, city, speed, down);
}
int wo_cappos_net_ctl(void* h, int op, int arg) { return netCtl(static_cast<SanFerminCappos*>(h)->network, op, arg); }
// per node: currentPrefixLength, totalNumberOfSigs(-1), done, thresholdDone, isSwapping, cached levels (bit mask) ; thresholdAt
void wo_cappos_node_scalars(void* h, int32_t* cpl, int32_t* sigs, int32_t* done, int32_t* thrDone, int32_t* swapping, int32_t* cacheMask,
                            int64_t* thresholdAt) {
  auto* p = static_cast<SanFerminCappos*>(h);
  for (size_t i = 0; i < p->nodes.size(); ++i) {
    auto& n = *p->nodes[i];
    cpl[i] = n.currentPrefixLength;
    sigs[i] = n.totalNumberOfSigs(-1);
    done[i] = n.done ? 1 : 0;
    thrDone[i] = n.thresholdDone ? 1 : 0;
    swapping[i] = n.isSwapping ? 1 : 0;
    int m = 0;
    for (auto& kv : n.signatureCache) m |= 1 << kv.first;
    cacheMask[i] = m;
    thresholdAt[i] = n.thresholdAt;
  }
}

// PingPong: network.send(new Ping()/Pong(), from, dests) issued by the caller (type 1 = Ping, 2 = Pong)
int wo_pp_send(void* h, int type, int from, const int32_t* to, int n) {
  WO_TRY
  auto* p = static_cast<PingPong*>(h);
  std::vector<Node*> dests;
  for (int i = 0; i < n; ++i) dests.push_back(&p->network.getNodeById(to[i]));
  MessagePtr m;
  if (type == 1)
    m = std::make_shared<PingPong::Ping>();
  else
    m = std::make_shared<PingPong::Pong>();
  p->network.send(m, p->network.getNodeById(from), dests);
  return 0;
  WO_CATCH(-1)
}
int wo_pp_send_at(void* h, int type, int from, const int32_t* to, int n, int sendTime, int delay) {
  WO_TRY
  auto* p = static_cast<PingPong*>(h);
  MessagePtr m;
  if (type == 1)
    m = std::make_shared<PingPong::Ping>();
  else
    m = std::make_shared<PingPong::Pong>();
  if (n == 1) {
    p->network.send(m, sendTime, p->network.getNodeById(from), p->network.getNodeById(to[0]));
  } else {
    std::vector<Node*> dests;
    for (int i = 0; i < n; ++i) dests.push_back(&p->network.getNodeById(to[i]));
    p->network.send(m, sendTime, p->network.getNodeById(from), dests, delay);
  }
  return 0;
  WO_CATCH(-1)
}
// CasperIMD: like the others, but endPartition is BlockChainNetwork.endPartition (every node re-sends its head)
int wo_casper_net_ctl(void* h, int op, int arg) {
  if (op == 3) {
    WO_TRY
    static_cast<CasperIMD*>(h)->endPartition();
    return 0;
    WO_CATCH(-1)
  }
  return netCtl(static_cast<CasperIMD*>(h)->network, op, arg);
}
}  // extern "C"
