// wittgenstein_b200 — C++ host-side mirror of the reference classes for the accelerated path, over the C ABI of wtg.h.
//
// The reference is Java and no JVM exists in the build image, so the host side above the C ABI is written in C++ with the
// reference's names, argument order and error behaviour (unchecked exceptions -> wtg_b200::WtgError):
//   core/Network.java            -> Network          (rd.setSeed, runMs, run, time, msgs.size(), partition, ...)
//   protocols/PingPong.java      -> PingPong / PingPongParameters
//   protocols/GSFSignature.java  -> GSFSignature / GSFSignatureParameters
//   protocols/SanFerminSignature.java -> SanFerminSignature / SanFerminSignatureParameters
//   protocols/SanFerminCappos.java -> SanFerminCappos / SanFerminCapposParameters
//   protocols/Handel.java        -> Handel / HandelParameters
//   protocols/CasperIMD.java     -> CasperIMD / CasperParemeters (the reference's spelling)
// Header-only; link with wittgenstein_b200/libwtg_b200.so.  tests/cpp/mirror_parity.cpp drives it against the CPU oracle.
#pragma once
#include <cstdint>
#include <functional>
#include <stdexcept>
#include <thread>
#include <string>
#include <vector>

#include "wtg.h"

namespace wtg_b200 {

struct WtgError : std::runtime_error {  // IllegalArgumentException / IllegalStateException of the reference
  using std::runtime_error::runtime_error;
};
inline int check(int rc) {
  if (rc < 0) throw WtgError(wtg_last_error());
  return rc;
}
inline const char* cstr(const std::string& s) { return s.empty() ? nullptr : s.c_str(); }  // "" stands for Java null

struct NodeCounters {  // Node.java:72-79
  std::vector<long long> msgReceived, msgSent, bytesSent, bytesReceived, doneAt;
};

// core/Network.java
class Network {
 public:
  struct Rd {  // network.rd
    Network* net;
    void setSeed(long long seed) { check(wtg_set_seed(net->h, seed)); }  // RunMultipleTimes.java:47
  } rd{this};

  Network() : h(wtg_create()) {
    if (!h) throw WtgError(wtg_last_error());
  }
  ~Network() { wtg_destroy(h); }
  Network(const Network&) = delete;
  Network& operator=(const Network&) = delete;

  void setNetworkLatency(const std::string& registryName) { check(wtg_set_network_latency(h, cstr(registryName))); }
  void setNetworkLatency(const std::vector<int>& distribProp, const std::vector<int>& distribVal) {  // Network.java:665-667
    check(wtg_set_network_latency_measured(h, distribProp.data(), distribVal.data(), (int)distribProp.size()));
  }
  void setNodeBuilder(const std::string& registryName) { check(wtg_set_node_builder(h, cstr(registryName))); }
  void setMsgDiscardTime(int ms) { check(wtg_set_msg_discard_time(h, ms)); }
  void setTunable(const std::string& key, long long v) { check(wtg_set_tunable(h, key.c_str(), v)); }

  bool runMs(int ms) { return check(wtg_run_ms(h, ms)) == 1; }  // Network.java:318-338
  bool run(int seconds) { return runMs(seconds * 1000); }       // :306-308
  int time() const { return wtg_time(h); }
  int nodeCount() const { return wtg_node_count(h); }
  struct Msgs {  // network.msgs
    const Network* net;
    int size() const { return check(wtg_msgs_size(net->h)); }
    int sizeAt(int t) const { return check(wtg_msgs_size_at(net->h, t)); }
  } msgs{this};

  // network.send(msg, from, to) / send(msg, from, dests) / sendAll(msg, from) issued by the caller (Network.java:341-366);
  // msgType / payload name a message of the running protocol (see wtg.h)
  void send(int msgType, int from, int to, unsigned long long payload = 0) { check(wtg_send(h, msgType, payload, from, &to, 1)); }
  void send(int msgType, int from, const std::vector<int>& dests, unsigned long long payload = 0) {
    check(wtg_send(h, msgType, payload, from, dests.data(), (int)dests.size()));
  }
  void sendAll(int msgType, int from, unsigned long long payload = 0) { check(wtg_send_all(h, msgType, payload, from)); }
  // send(msg, sendTime, from, dests, delaysBetweenMessage) (Network.java:420-447)
  void send(int msgType, int sendTime, int from, const std::vector<int>& dests, int delayBetweenMessages, unsigned long long payload = 0) {
    check(wtg_send_at(h, msgType, payload, from, dests.data(), (int)dests.size(), sendTime, delayBetweenMessages));
  }
  void stopNode(int id) { check(wtg_stop_node(h, id)); }    // node.stop()
  void startNode(int id) { check(wtg_start_node(h, id)); }  // node.start()
  void partition(float part) { check(wtg_partition(h, part)); }
  void endPartition() { check(wtg_end_partition(h)); }
  unsigned long long rngState() const { return wtg_rng_state(h); }

  NodeCounters counters() const {
    size_t n = (size_t)nodeCount();
    std::vector<long long> raw(5 * n);
    check(wtg_node_counters(h, raw.data()));
    NodeCounters c;
    c.msgReceived.assign(raw.begin(), raw.begin() + n);
    c.msgSent.assign(raw.begin() + n, raw.begin() + 2 * n);
    c.bytesSent.assign(raw.begin() + 2 * n, raw.begin() + 3 * n);
    c.bytesReceived.assign(raw.begin() + 3 * n, raw.begin() + 4 * n);
    c.doneAt.assign(raw.begin() + 4 * n, raw.end());
    return c;
  }
  std::vector<unsigned char> down() const {  // Node.isDown()
    std::vector<unsigned char> d((size_t)nodeCount());
    check(wtg_node_attrs(h, nullptr, nullptr, nullptr, nullptr, nullptr, d.data()));
    return d;
  }
  wtg_net* handle() const { return h; }

 private:
  wtg_net* h;
};

// protocols/PingPong.java:34-50
struct PingPongParameters {
  int nodeCt = 1000;
  std::string nodeBuilderName, networkLatencyName;
};
class PingPong {
 public:
  explicit PingPong(const PingPongParameters& p) : params(p) {
    net.setNodeBuilder(p.nodeBuilderName);
    net.setNetworkLatency(p.networkLatencyName);
  }
  Network& network() { return net; }
  void init() { check(wtg_pingpong_init(net.handle(), params.nodeCt)); }  // :82-87
  std::vector<int> pong() {
    std::vector<int> v((size_t)params.nodeCt);
    check(wtg_pingpong_pongs(net.handle(), v.data()));
    return v;
  }
  const PingPongParameters params;

 private:
  Network net;
};

// protocols/GSFSignature.java:27-107
struct GSFSignatureParameters {
  int nodeCount, threshold, pairingTime, timeoutPerLevelMs, periodDurationMs, acceleratedCallsCount, nodesDown;
  std::string nodeBuilderName, networkLatencyName;
};
class GSFSignature {
 public:
  explicit GSFSignature(const GSFSignatureParameters& p) : params(p) {
    net.setNodeBuilder(p.nodeBuilderName);
    net.setNetworkLatency(p.networkLatencyName);
  }
  Network& network() { return net; }
  void init() {  // :611-635
    int a[7] = {params.nodeCount, params.threshold, params.pairingTime, params.timeoutPerLevelMs, params.periodDurationMs,
                params.acceleratedCallsCount, params.nodesDown};
    check(wtg_gsf_init(net.handle(), a));
  }
  int words() const { return params.nodeCount < 64 ? 1 : params.nodeCount / 64; }
  std::vector<unsigned long long> verifiedSignatures() {  // GSFNode.verifiedSignatures, N rows of words()
    std::vector<unsigned long long> v((size_t)params.nodeCount * (size_t)words());
    check(wtg_gsf_verified(net.handle(), v.data()));
    return v;
  }
  struct Scalars {
    std::vector<int> nodePairingTime, sigChecked, sigQueueSize, toVerifySize, cardinality;
  };
  Scalars scalars() {
    size_t n = (size_t)params.nodeCount;
    Scalars s{std::vector<int>(n), std::vector<int>(n), std::vector<int>(n), std::vector<int>(n), std::vector<int>(n)};
    check(wtg_gsf_node_scalars(net.handle(), s.nodePairingTime.data(), s.sigChecked.data(), s.sigQueueSize.data(), s.toVerifySize.data(),
                               s.cardinality.data()));
    return s;
  }
  bool continueIf() {  // newConfIf :670-682
    Scalars s = scalars();
    std::vector<unsigned char> d = net.down();
    for (int i = 0; i < params.nodeCount; ++i)
      if (!d[(size_t)i] && s.cardinality[(size_t)i] < params.threshold) return true;
    return false;
  }
  const GSFSignatureParameters params;

 private:
  Network net;
};

// protocols/SanFerminSignature.java:41-110
struct SanFerminSignatureParameters {
  int nodeCount, threshold, pairingTime, signatureSize, replyTimeout, candidateCount;
  std::string nodeBuilderName, networkLatencyName;
};
class SanFerminSignature {
 public:
  explicit SanFerminSignature(const SanFerminSignatureParameters& p) : params(p) {  // the constructor builds the nodes (:112-129)
    net.setNodeBuilder(p.nodeBuilderName);
    net.setNetworkLatency(p.networkLatencyName);
    int a[6] = {p.nodeCount, p.threshold, p.pairingTime, p.signatureSize, p.replyTimeout, p.candidateCount};
    check(wtg_sanfermin_construct(net.handle(), a));
  }
  Network& network() { return net; }
  void init() { check(wtg_sanfermin_init(net.handle())); }
  const SanFerminSignatureParameters params;

 private:
  Network net;
};

// protocols/SanFerminCappos.java:43-104
struct SanFerminCapposParameters {
  int nodeCount, threshold, pairingTime, signatureSize, timeout, candidateCount;
  std::string nodeBuilderName, networkLatencyName;
};
class SanFerminCappos {
 public:
  explicit SanFerminCappos(const SanFerminCapposParameters& p) : params(p) {
    net.setNodeBuilder(p.nodeBuilderName);
    net.setNetworkLatency(p.networkLatencyName);
  }
  Network& network() { return net; }
  void init() {  // :120-134
    int a[6] = {params.nodeCount, params.threshold, params.pairingTime, params.signatureSize, params.timeout, params.candidateCount};
    check(wtg_cappos_init(net.handle(), a));
  }
  const SanFerminCapposParameters params;

 private:
  Network net;
};

// protocols/Handel.java:22-142
struct HandelParameters {
  int nodeCount, threshold, pairingTime, levelWaitTime, extraCycle, disseminationPeriodMs, fastPath, nodesDown;
  std::string nodeBuilderName, networkLatencyName;
  int desynchronizedStart = 0;
  bool byzantineSuicide = false, hiddenByzantine = false;
};
class Handel {
 public:
  explicit Handel(const HandelParameters& p) : params(p) {
    net.setNodeBuilder(p.nodeBuilderName);
    net.setNetworkLatency(p.networkLatencyName);
  }
  Network& network() { return net; }
  void init() {  // :957-1014
    int a[11] = {params.nodeCount, params.threshold, params.pairingTime, params.levelWaitTime, params.extraCycle,
                 params.disseminationPeriodMs, params.fastPath, params.nodesDown, params.desynchronizedStart,
                 params.byzantineSuicide ? 1 : 0, params.hiddenByzantine ? 1 : 0};
    check(wtg_handel_init(net.handle(), a));
  }
  std::vector<unsigned long long> totalIncoming() {  // union over levels, N rows of N/64 words
    std::vector<unsigned long long> v((size_t)params.nodeCount * (size_t)(params.nodeCount < 64 ? 1 : params.nodeCount / 64));
    check(wtg_handel_rows(net.handle(), 0, v.data()));
    return v;
  }
  const HandelParameters params;

 private:
  Network net;
};

// protocols/CasperIMD.java:18-71
struct CasperParemeters {
  int cycleLength = 4;
  bool randomOnTies = true;
  int blockProducersCount = 2, attestersPerRound = 20, blockConstructionTime = 1000, attestationConstructionTime = 1;
  std::string nodeBuilderName, networkLatencyName;
};
class CasperIMD {
 public:
  enum ByzKind { ByzBlockProducer = 3, ByzBlockProducerSF = 4, ByzBlockProducerNS = 5, ByzBlockProducerWF = 6 };
  explicit CasperIMD(const CasperParemeters& p) : params(p) {  // adds the observer (:81-88)
    net.setNodeBuilder(p.nodeBuilderName);
    net.setNetworkLatency(p.networkLatencyName);
    int a[6] = {p.cycleLength, p.randomOnTies ? 1 : 0, p.blockProducersCount, p.attestersPerRound, p.blockConstructionTime,
                p.attestationConstructionTime};
    check(wtg_casper_construct(net.handle(), a));
  }
  Network& network() { return net; }
  int nodeCount() const { return 1 + params.blockProducersCount + params.attestersPerRound * params.cycleLength; }
  void init() { init(ByzBlockProducerWF, 0); }  // :472-476
  void init(ByzKind kind, int delay) { check(wtg_casper_init_byz(net.handle(), (int)kind, delay)); }
  std::vector<int> heads() {  // BlockChainNode.head (block id) of every node; node 0 is network.observer
    std::vector<int> v((size_t)nodeCount());
    check(wtg_casper_heads(net.handle(), v.data()));
    return v;
  }
  struct Blocks {
    std::vector<int> height, parent, producer, proposalTime, included;
  };
  Blocks blocks() {
    size_t nb = (size_t)check(wtg_casper_block_count(net.handle()));
    Blocks b{std::vector<int>(nb), std::vector<int>(nb), std::vector<int>(nb), std::vector<int>(nb), std::vector<int>(nb)};
    check(wtg_casper_blocks(net.handle(), b.height.data(), b.parent.data(), b.producer.data(), b.proposalTime.data(), b.included.data()));
    return b;
  }
  const CasperParemeters params;

 private:
  Network net;
};

// core/utils/StatsHelper.java:83-134
struct SimpleStats {
  long long min = 0, max = 0, avg = 0;
};
inline SimpleStats getStatsOn(const std::vector<long long>& values) {  // Java long arithmetic: avg = total / count, truncating
  if (values.empty()) throw WtgError("no live node");
  SimpleStats s;
  s.min = s.max = values[0];
  long long tot = 0;
  for (long long v : values) {
    tot += v;
    if (v < s.min) s.min = v;
    if (v > s.max) s.max = v;
  }
  s.avg = tot / (long long)values.size();
  return s;
}
inline SimpleStats avg(const std::vector<SimpleStats>& stats) {  // StatsHelper.avg :32-52
  if (stats.empty()) throw WtgError("no stats");
  if (stats.size() == 1) return stats[0];
  SimpleStats s;
  for (const SimpleStats& x : stats) {
    s.min += x.min;
    s.max += x.max;
    s.avg += x.avg;
  }
  s.min /= (long long)stats.size();
  s.max /= (long long)stats.size();
  s.avg /= (long long)stats.size();
  return s;
}

// core/RunMultipleTimes.java:41-85 for a protocol mirror P (constructible from its parameter struct, with network(), init()):
// runCount seeded runs of `runMs(10) while (maxTime == 0 || time < maxTime) && (!didSomething || contIf(p))`, the doneAt and
// msgReceived stats of the live nodes averaged over the runs.  Up to `concurrency` runs are in flight at once, each on its own
// engine instance and CUDA stream.
template <class P, class Params>
struct RunMultipleTimes {
  Params params;
  int runCount, maxTime;
  std::vector<int> endTimes;
  struct Result {
    SimpleStats doneAt, msgReceived;
  };
  Result run(const std::function<bool(P&)>& contIf, int concurrency = 8) {
    std::vector<Result> per((size_t)runCount);
    endTimes.assign((size_t)runCount, 0);
    std::vector<std::string> errors((size_t)runCount);
    auto one = [&](int i) {
      try {
        P c(params);
        c.network().rd.setSeed(i);  // :47
        c.init();
        bool did;
        do {
          did = c.network().runMs(10);
        } while ((maxTime == 0 || c.network().time() < maxTime) && (!did || (contIf && contIf(c))));
        NodeCounters cnt = c.network().counters();
        std::vector<unsigned char> down = c.network().down();
        std::vector<long long> d, m;
        for (size_t n = 0; n < down.size(); ++n)
          if (!down[n]) {
            d.push_back(cnt.doneAt[n]);
            m.push_back(cnt.msgReceived[n]);
          }
        per[(size_t)i] = Result{getStatsOn(d), getStatsOn(m)};
        endTimes[(size_t)i] = c.network().time();
      } catch (const std::exception& e) {
        errors[(size_t)i] = e.what();
      }
    };
    for (int i0 = 0; i0 < runCount; i0 += concurrency) {
      std::vector<std::thread> th;
      for (int i = i0; i < runCount && i < i0 + concurrency; ++i) th.emplace_back(one, i);
      for (auto& t : th) t.join();
    }
    for (int i = 0; i < runCount; ++i)
      if (!errors[(size_t)i].empty()) throw WtgError("Failed execution for random seed of " + std::to_string(i) + ": " + errors[(size_t)i]);
    std::vector<SimpleStats> d, m;
    for (const Result& r : per) {
      d.push_back(r.doneAt);
      m.push_back(r.msgReceived);
    }
    return Result{avg(d), avg(m)};
  }
};

}  // namespace wtg_b200
