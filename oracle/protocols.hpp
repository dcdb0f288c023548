// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference protocols on the hot path:
//   protocols/PingPong.java       -> PingPong
//   protocols/GSFSignature.java   -> GSFSignature (GSFNode, SFLevel, SendSigs)
//   protocols/SanFerminSignature.java + SanFerminHelper.java -> SanFerminSignature, SanFerminHelper
//   protocols/Handel.java         -> Handel (HNode, HLevel, SendSigs, SigToVerify, HiddenByzantine)
// Line references are to those files.  PARITY STATUS: structure / schedule / liveness are
// pinned by the reference's own tests (PT/GSFSignatureTest.java, PT/PingPongTest.java,
// restated in tests/test_oracle_protocols.py); the protocol END STATE (bitmaps, doneAt) is
// "parity unpinned" — the reference's tests hold no golden end state and no JVM is available.
#pragma once
#include <atomic>
#include <map>
#include <memory>
#include <unordered_map>
#include <unordered_set>
#include <string>
#include <thread>
#include <vector>

#include "core.hpp"
#include "jbitset.hpp"

namespace wo {

// ----------------------------------------------------------------------------------------
// PingPong  (protocols/PingPong.java)
// ----------------------------------------------------------------------------------------
struct PingPong {
  struct Params {
    int nodeCt = 1000;
    std::string nodeBuilderName;  // empty == null
    std::string networkLatencyName;
    bool latencyNull = true;
  };
  struct PingPongNode : Node {
    int pong = 0;
    PingPong* p;
    PingPongNode(PingPong* pp) : Node(pp->network.rd, pp->nb), p(pp) {}
  };
  struct Pong : Message {
    void action(Network&, Node&, Node& to) override { static_cast<PingPongNode&>(to).pong++; }  // :77-79
  };
  struct Ping : Message {
    void action(Network& network, Node& from, Node& to) override {  // :73-75
      network.send(std::make_shared<Pong>(), to, from);
    }
  };

  Params params;
  Network network;
  NodeBuilder nb;
  std::vector<std::unique_ptr<PingPongNode>> nodes;

  explicit PingPong(const Params& p) : params(p) {  // :52-57
    nb = nodeBuilderByName(p.nodeBuilderName);
    network.setNetworkLatency(networkLatencyByName(p.networkLatencyName, p.latencyNull));
  }
  void init() {  // :82-87
    for (int i = 0; i < params.nodeCt; i++) {
      nodes.push_back(std::make_unique<PingPongNode>(this));
      network.addNode(nodes.back().get());
    }
    network.sendAll(std::make_shared<Ping>(), network.getNodeById(0));
  }
};

// ----------------------------------------------------------------------------------------
// GSFSignature  (protocols/GSFSignature.java)
// ----------------------------------------------------------------------------------------
inline int roundPow2(int n) {  // core/utils/MoreMath.java:13-19
  int res = 1;
  while ((res << 1) > 0 && (res << 1) <= n) res <<= 1;  // Integer.highestOneBit
  if (res != n) res <<= 1;
  return res;
}

struct GSFSignature {
  struct Params {  // :27-107
    int nodeCount = 32768 / 32;
    int threshold = static_cast<int>((32768 / 32) * 0.99);
    int pairingTime = 3;
    int timeoutPerLevelMs = 50;
    int periodDurationMs = 10;
    int acceleratedCallsCount = 10;
    int nodesDown = 0;
    std::string nodeBuilderName;
    std::string networkLatencyName;
    bool latencyNull = false;
  };
  static Params makeParams(int nodeCount, int threshold, int pairingTime, int timeoutPerLevelMs, int periodDurationMs,
                           int acceleratedCallsCount, int nodesDown, const std::string& nb, const std::string& nl) {
    if (nodesDown >= nodeCount || nodesDown < 0 || threshold > nodeCount || (nodesDown + threshold > nodeCount))
      throw IllegalArgument("nodeCount/threshold");  // :69-74
    Params p;
    p.nodeCount = nodeCount;
    p.threshold = threshold;
    p.pairingTime = pairingTime;
    p.timeoutPerLevelMs = timeoutPerLevelMs;
    p.periodDurationMs = periodDurationMs;
    p.acceleratedCallsCount = acceleratedCallsCount;
    p.nodesDown = nodesDown;
    p.nodeBuilderName = nb;
    p.networkLatencyName = nl;
    return p;
  }

  struct GSFNode;
  struct SFLevel;

  struct SendSigs : Message, std::enable_shared_from_this<SendSigs> {  // :137-164
    JBitSet sigs;  // mutable and shared between receivers of a multi-dest send
    GSFNode* from;
    int level;
    int size_;
    SendSigs(GSFNode* f, const JBitSet& s, const SFLevel& l);
    int size() const override { return size_; }
    void action(Network&, Node& from_, Node& to) override;
  };
  using SendSigsPtr = std::shared_ptr<SendSigs>;

  struct SFLevel {  // :235-356
    GSFNode* node;
    int level;
    std::vector<uint32_t> peers;  // node ids
    JBitSet waitedSigs, verifiedSignatures, individualSignatures, indivVerifiedSig;
    int waitedCard = 0;  // == waitedSigs.cardinality() (immutable after construction)
    int posInLevel = 0;
    int remainingCalls = 0;

    int expectedSigs() const { return waitedCard; }  // :286-288
    bool hasStarted(const JBitSet& toSend) const;    // :291-311
    void doCycle(const JBitSet& toSend);             // :313-323
    std::vector<GSFNode*> getRemainingPeers(int peersCt);  // :325-349
  };

  struct GSFNode : Node {  // :166-604
    GSFSignature* p;
    std::vector<SendSigsPtr> toVerify;
    std::vector<SFLevel> levels;
    JBitSet verifiedSignatures;
    int nodePairingTime;
    bool done = false;
    int sigChecked = 0;
    int sigQueueSize = 0;

    explicit GSFNode(GSFSignature* pp)
        : Node(pp->network.rd, pp->nb), p(pp), nodePairingTime(static_cast<int>(std::max(1.0, pp->params.pairingTime * speedRatio))) {
      verifiedSignatures.set(nodeId);  // :176-179
    }
    JBitSet allSigsAtLevel(int round) const;  // :359-372
    void initLevel(JavaRandom& rd);           // :181-191 (rd == network.rd in the reference)
    JBitSet getLastFinishedLevel() const;     // :193-210
    void doCycle();                           // :212-224
    static bool include(const JBitSet& large, const JBitSet& small) {  // :374-378
      JBitSet a = large;
      a.and_(small);
      return a.equals(small);
    }
    void updateVerifiedSignatures(GSFNode* from, int level, JBitSet& sigs);  // :384-460
    int evaluateSig(const SFLevel& l, const JBitSet& sig) const;             // :482-534
    void onNewSig(GSFNode* from, const SendSigsPtr& ssigs);                  // :537-555
    void checkSigs();                                                       // :557-583
  };

  Params params;
  Network network;
  NodeBuilder nb;
  std::vector<std::unique_ptr<GSFNode>> nodes;
  // statistics for sizing / roofline accounting (not in the reference)
  int64_t statEvalEntries = 0, statEvalBytes = 0, statUpdates = 0, statCycles = 0, statSends = 0, statMultiSends = 0;
  int64_t statSendBytes = 0, statMaxQueue = 0;

  explicit GSFSignature(const Params& p) : params(p) {  // :109-114
    nb = nodeBuilderByName(p.nodeBuilderName);
    network.setNetworkLatency(networkLatencyByName(p.networkLatencyName, p.latencyNull));
  }
  GSFNode& node(int i) { return *nodes[static_cast<size_t>(i)]; }

  void init() {  // :611-635
    for (int i = 0; i < params.nodeCount; i++) {
      nodes.push_back(std::make_unique<GSFNode>(this));
      network.addNode(nodes.back().get());
    }
    for (int setDown = 0; setDown < params.nodesDown;) {
      int down = network.rd.nextInt(params.nodeCount);
      Node& n = *network.allNodes[static_cast<size_t>(down)];
      if (!n.isDown() && down != 1) {
        n.stop();
        setDown++;
      }
    }
    for (auto& up : nodes) {
      GSFNode* n = up.get();
      if (!n->isDown()) {
        n->initLevel(network.rd);
        network.registerPeriodicTask([n] { n->doCycle(); }, 1, params.periodDurationMs, *n);
        network.registerConditionalTask([n] { n->checkSigs(); }, 1, n->nodePairingTime, *n,
                                        [n] { return !n->toVerify.empty(); }, [n] { return !n->done; });
      }
    }
  }
  // Same result as init(), built for large N (test infrastructure convenience, not a reference code path):
  // a sequential pre-pass steps network.rd through every shuffle draw (no swaps) to record the RNG state at
  // the start of each node, then the per-node initLevel() calls run on `threads` threads from those states.
  // tests/test_oracle_protocols.py checks initFast() == init() state for state.
  void initFast(int threads);

  // newConfIf :670-682 — true while some live node is below the threshold
  bool continueIf() const {
    for (auto& n : nodes)
      if (!n->isDown() && n->verifiedSignatures.cardinality() < params.threshold) return true;
    return false;
  }
};

inline GSFSignature::SendSigs::SendSigs(GSFNode* f, const JBitSet& s, const SFLevel& l)
    : sigs(s), from(f), level(l.level), size_(1 + l.expectedSigs() / 8 + 96) {}  // :145-153 (clone; size)

inline void GSFSignature::SendSigs::action(Network&, Node& from_, Node& to) {  // :161-163
  GSFNode& t = static_cast<GSFNode&>(to);
  t.onNewSig(static_cast<GSFNode*>(&from_), shared_from_this());
}

inline JBitSet GSFSignature::GSFNode::allSigsAtLevel(int round) const {  // :359-372
  if (round < 1) throw IllegalArgument("round");
  JBitSet res;
  int cMask = (1 << round) - 1;
  int start = (cMask | nodeId) ^ cMask;
  int end = nodeId | cMask;
  end = std::min(end, p->params.nodeCount - 1);
  res.setRange(start, end + 1);
  res.set(nodeId, false);
  return res;
}

inline void GSFSignature::GSFNode::initLevel(JavaRandom& rd) {  // :181-191, SFLevel ctors :260-280, randomSubset :462-476
  int roundedPow2NodeCount = roundPow2(p->params.nodeCount);
  JBitSet allPreviousNodes;
  levels.emplace_back();
  {
    SFLevel& l0 = levels.back();  // level 0: only our own signature, no peer (:260-267)
    l0.node = this;
    l0.level = 0;
    l0.waitedSigs.set(nodeId);
    l0.waitedCard = 1;
    l0.verifiedSignatures.set(nodeId);
    l0.remainingCalls = 0;
  }
  for (int l = 1; (1LL << l) <= roundedPow2NodeCount; l++) {
    allPreviousNodes.or_(levels.back().waitedSigs);
    SFLevel nl;
    nl.node = this;
    nl.level = levels.back().level + 1;
    nl.waitedSigs = allSigsAtLevel(nl.level);
    nl.waitedSigs.andNot(allPreviousNodes);
    nl.waitedCard = nl.waitedSigs.cardinality();
    // randomSubset(waitedSigs, MAX_VALUE): ids in increasing order, then Collections.shuffle(res, network.rd)
    nl.peers.reserve(static_cast<size_t>(nl.waitedCard));
    for (int cur = nl.waitedSigs.nextSetBit(0); cur >= 0; cur = nl.waitedSigs.nextSetBit(cur + 1))
      nl.peers.push_back(static_cast<uint32_t>(cur));
    javaShuffle(nl.peers, rd);
    nl.remainingCalls = static_cast<int>(nl.peers.size());
    levels.push_back(std::move(nl));
  }
}

inline JBitSet GSFSignature::GSFNode::getLastFinishedLevel() const {  // :193-210
  JBitSet res;
  const SFLevel* sfl = &levels[0];
  bool done_ = false;
  while (!done_) {
    if (sfl->waitedSigs.equals(sfl->verifiedSignatures)) {
      res.or_(sfl->waitedSigs);
      if (sfl->level < static_cast<int>(levels.size()) - 1)
        sfl = &levels[static_cast<size_t>(sfl->level + 1)];
      else
        done_ = true;
    } else {
      done_ = true;
    }
  }
  return res;
}

inline void GSFSignature::GSFNode::doCycle() {  // :212-224
  ++p->statCycles;
  JBitSet toSend = getLastFinishedLevel();
  for (SFLevel& sfl : levels) {
    sfl.doCycle(toSend);
    toSend.or_(sfl.verifiedSignatures);
  }
}

inline bool GSFSignature::SFLevel::hasStarted(const JBitSet& toSend) const {  // :291-311
  if (node->p->network.time >= level * node->p->params.timeoutPerLevelMs) return true;
  if (toSend.cardinality() >= expectedSigs()) return true;
  return false;
}

inline void GSFSignature::SFLevel::doCycle(const JBitSet& toSend) {  // :313-323
  if (remainingCalls == 0 || !hasStarted(toSend)) return;
  std::vector<GSFNode*> dest = getRemainingPeers(1);
  if (!dest.empty()) {
    auto ss = std::make_shared<SendSigs>(node, toSend, *this);
    ++node->p->statSends;
    node->p->statSendBytes += static_cast<int64_t>(toSend.storedBytes());
    node->p->network.send(ss, *node, *dest[0]);
  }
}

inline std::vector<GSFSignature::GSFNode*> GSFSignature::SFLevel::getRemainingPeers(int peersCt) {  // :325-349
  std::vector<GSFNode*> res;
  while (peersCt > 0 && remainingCalls > 0) {
    remainingCalls--;
    GSFNode* pn = &node->p->node(static_cast<int>(peers[static_cast<size_t>(posInLevel++)]));
    if (posInLevel >= static_cast<int>(peers.size())) posInLevel = 0;
    // `count == null || true` (:338): the "skip finished peers" branch is dead code
    res.push_back(pn);
    peersCt--;
  }
  return res;
}

inline void GSFSignature::GSFNode::updateVerifiedSignatures(GSFNode* from, int level, JBitSet& sigsRef) {  // :384-460
  ++p->statUpdates;
  SFLevel* sfl = &levels[static_cast<size_t>(level)];
  JBitSet* sigs = &sigsRef;  // aliases the (possibly shared) message payload, like the reference
  JBitSet local;

  if (sigs->cardinality() == 1) sfl->indivVerifiedSig.set(from->nodeId);
  sigs->or_(sfl->indivVerifiedSig);

  bool resetRemaining = false;
  if (sigs->cardinality() > sfl->expectedSigs()) {
    for (size_t i = 1; i < levels.size() && include(*sigs, levels[i].waitedSigs); i++) {
      SFLevel& l = levels[i];
      if (!l.verifiedSignatures.equals(l.waitedSigs)) {
        l.verifiedSignatures.or_(l.waitedSigs);
        verifiedSignatures.or_(l.waitedSigs);
        resetRemaining = true;
      }
      if (resetRemaining) l.remainingCalls = static_cast<int>(l.peers.size());
    }
    local = sfl->waitedSigs;  // sigs = (BitSet) sfl.waitedSigs.clone();
    sigs = &local;
  }

  if (sfl->verifiedSignatures.cardinality() > 0 && !sigs->intersects(sfl->verifiedSignatures)) {
    sigs->or_(sfl->verifiedSignatures);
  }

  if (sigs->cardinality() > sfl->verifiedSignatures.cardinality() || resetRemaining) {
    for (size_t i = static_cast<size_t>(sfl->level); i < levels.size(); i++)
      levels[i].remainingCalls = static_cast<int>(levels[i].peers.size());

    sfl->verifiedSignatures.andNot(sfl->waitedSigs);
    sfl->verifiedSignatures.or_(*sigs);

    verifiedSignatures.andNot(sfl->waitedSigs);
    verifiedSignatures.or_(*sigs);

    if (p->params.acceleratedCallsCount > 0) {
      JBitSet bestToSend = getLastFinishedLevel();
      while (include(bestToSend, sfl->waitedSigs) && sfl->level < static_cast<int>(levels.size()) - 1) {
        sfl = &levels[static_cast<size_t>(sfl->level + 1)];
        auto sendSigs = std::make_shared<SendSigs>(this, bestToSend, *sfl);
        std::vector<GSFNode*> peers = sfl->getRemainingPeers(p->params.acceleratedCallsCount);
        if (!peers.empty()) {
          ++p->statMultiSends;
          std::vector<Node*> dests(peers.begin(), peers.end());
          p->network.send(sendSigs, *this, dests);
        }
      }
    }
    if (doneAt == 0 && verifiedSignatures.cardinality() >= p->params.threshold) doneAt = p->network.time;
  }
}

inline int GSFSignature::GSFNode::evaluateSig(const SFLevel& l, const JBitSet& sig) const {  // :482-534
  int newTotal = 0, addedSigs = 0;
  if (l.verifiedSignatures.cardinality() >= l.expectedSigs()) return 0;
  JBitSet withIndiv = l.indivVerifiedSig;
  withIndiv.or_(sig);
  if (l.verifiedSignatures.cardinality() == 0) {
    newTotal = sig.cardinality();
    addedSigs = newTotal;
  } else {
    if (sig.intersects(l.verifiedSignatures)) {
      newTotal = withIndiv.cardinality();
      addedSigs = newTotal - l.verifiedSignatures.cardinality();
    } else {
      withIndiv.or_(l.verifiedSignatures);
      newTotal = withIndiv.cardinality();
      addedSigs = newTotal - l.verifiedSignatures.cardinality();
    }
  }
  if (addedSigs <= 0) {
    if (sig.cardinality() == 1 && !sig.intersects(l.indivVerifiedSig)) return 1;
    return 0;
  }
  if (newTotal == l.expectedSigs()) return 1000000 - l.level * 10;
  return 100000 - l.level * 100 + addedSigs;
}

inline void GSFSignature::GSFNode::onNewSig(GSFNode* from, const SendSigsPtr& ssigs) {  // :537-555
  SFLevel& l = levels[static_cast<size_t>(ssigs->level)];
  // l.received.put(from, 1) feeds only the dead branch of getRemainingPeers (:337-345)
  toVerify.push_back(ssigs);
  if (!l.individualSignatures.get(from->nodeId)) {
    JBitSet indiv;
    indiv.set(from->nodeId);
    toVerify.push_back(std::make_shared<SendSigs>(from, indiv, l));
    l.individualSignatures.set(from->nodeId);
  }
  sigQueueSize = static_cast<int>(toVerify.size());
  if (sigQueueSize > p->statMaxQueue) p->statMaxQueue = sigQueueSize;
}

inline void GSFSignature::GSFNode::checkSigs() {  // :557-583
  SendSigsPtr best;
  int score = 0;
  size_t w = 0;
  for (size_t i = 0; i < toVerify.size(); ++i) {
    SendSigsPtr& cur = toVerify[i];
    const SFLevel& l = levels[static_cast<size_t>(cur->level)];
    ++p->statEvalEntries;
    p->statEvalBytes += static_cast<int64_t>(cur->sigs.storedBytes());
    int ns = evaluateSig(l, cur->sigs);
    bool remove = false;
    if (ns > score) {
      score = ns;
      best = cur;
    } else if (ns == 0) {
      remove = true;  // it.remove()
    }
    if (!remove) {
      if (w != i) toVerify[w] = std::move(toVerify[i]);
      ++w;
    }
  }
  toVerify.resize(w);

  if (best) {
    // toVerify.remove(best): first element equal (identity) to best
    for (size_t i = 0; i < toVerify.size(); ++i)
      if (toVerify[i] == best) {
        toVerify.erase(toVerify.begin() + static_cast<long>(i));
        break;
      }
    sigChecked++;
    sigQueueSize = static_cast<int>(toVerify.size());
    SendSigsPtr tBest = best;
    GSFNode* self = this;
    p->network.registerTask([self, tBest] { self->updateVerifiedSignatures(tBest->from, tBest->level, tBest->sigs); },
                            p->network.time + nodePairingTime, *this);
  }
}

inline void GSFSignature::initFast(int threads) {
  for (int i = 0; i < params.nodeCount; i++) {
    nodes.push_back(std::make_unique<GSFNode>(this));
    network.addNode(nodes.back().get());
  }
  for (int setDown = 0; setDown < params.nodesDown;) {
    int down = network.rd.nextInt(params.nodeCount);
    Node& n = *network.allNodes[static_cast<size_t>(down)];
    if (!n.isDown() && down != 1) {
      n.stop();
      setDown++;
    }
  }
  const int N = params.nodeCount;
  const int rounded = roundPow2(N);
  std::vector<uint64_t> startSeed(static_cast<size_t>(N), 0);
  JavaRandom r = network.rd;
  for (int id = 0; id < N; ++id) {
    if (nodes[static_cast<size_t>(id)]->isDown()) continue;
    startSeed[static_cast<size_t>(id)] = r.seed;
    for (int l = 1; (1LL << l) <= rounded; l++) {
      // |waitedSigs| of level l: the sibling half of the 2^l block, clipped to [0, N)
      int mask = (1 << l) - 1, half = 1 << (l - 1);
      int start = (id | mask) ^ mask;
      int sibLo = (id & half) ? start : start + half;
      int sibHi = std::min(sibLo + half, N);
      int size = std::max(0, sibHi - sibLo);
      // the draws of Collections.shuffle without the swaps.  nextInt(i) can only loop when
      // u >= 2^31 - i (java.util.Random.nextInt rejection test), so the exact test runs on those draws only.
      for (int i = size; i > 1; i--) {
        int32_t u = r.next(31);
        if ((i & (i - 1)) != 0 && static_cast<uint32_t>(u) >= 0x80000000u - static_cast<uint32_t>(i)) {
          for (;;) {
            int32_t rem = u % i;
            if (static_cast<int32_t>(static_cast<uint32_t>(u) - static_cast<uint32_t>(rem) + static_cast<uint32_t>(i - 1)) >= 0) break;
            u = r.next(31);
          }
        }
      }
    }
  }
  std::atomic<int> next{0};
  auto work = [&]() {
    for (;;) {
      int id = next.fetch_add(1);
      if (id >= N) break;
      GSFNode* n = nodes[static_cast<size_t>(id)].get();
      if (n->isDown()) continue;
      JavaRandom rr(0);
      rr.seed = startSeed[static_cast<size_t>(id)];
      n->initLevel(rr);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < std::max(1, threads); ++t) pool.emplace_back(work);
  for (auto& t : pool) t.join();
  network.rd.draws += r.draws - network.rd.draws;
  network.rd.seed = r.seed;
  for (auto& up : nodes) {
    GSFNode* n = up.get();
    if (!n->isDown()) {
      network.registerPeriodicTask([n] { n->doCycle(); }, 1, params.periodDurationMs, *n);
      network.registerConditionalTask([n] { n->checkSigs(); }, 1, n->nodePairingTime, *n,
                                      [n] { return !n->toVerify.empty(); }, [n] { return !n->done; });
    }
  }
}

// ----------------------------------------------------------------------------------------
// SanFerminHelper  (protocols/SanFerminHelper.java) — node ids == indices in allNodes
// ----------------------------------------------------------------------------------------
inline int javaLog2(int n) {  // MoreMath.log2
  if (n <= 0) throw IllegalArgument("n");
  int r = 0;
  while ((1 << (r + 1)) > 0 && (1 << (r + 1)) <= n) ++r;
  return r;
}
struct SanFerminHelper {
  int nodeId = 0, setSize = 0;
  std::string binaryId;
  std::map<int, JBitSet> usedNodes;  // :25  level -> picked indices
  JavaRandom* rd = nullptr;

  static std::string toBinaryID(int nodeId, int setSize) {  // :169-172 + leftPadWithZeroes :159-167
    int log2 = javaLog2(setSize);
    std::string bin;
    for (unsigned v = static_cast<unsigned>(nodeId); v != 0; v >>= 1) bin.insert(bin.begin(), static_cast<char>('0' + (v & 1)));
    if (bin.empty()) bin = "0";
    if (static_cast<int>(bin.size()) > log2) throw IllegalState("StringIndexOutOfBounds in leftPadWithZeroes");
    return std::string(static_cast<size_t>(log2) - bin.size(), '0') + bin;
  }
  SanFerminHelper() = default;
  SanFerminHelper(int id, int n, JavaRandom* r) : nodeId(id), setSize(n), binaryId(toBinaryID(id, n)), rd(r) {}

  // [min, max) of allNodes.subList(min, max)
  std::pair<int, int> getOwnSet(int level) const {  // :46-64
    int mn = 0, mx = setSize;
    for (int currLevel = 0; currLevel <= level && mn <= mx; currLevel++) {
      int m = (mx + mn) / 2;
      char ch = binaryId.at(static_cast<size_t>(currLevel));
      if (ch == '0')
        mx = m;
      else if (ch == '1')
        mn = m;
      if (mx == mn) break;
      if (mx - 1 == 0 || mn == setSize) break;
    }
    return {mn, mx};
  }
  std::pair<int, int> getCandidateSet(int level) const {  // :70-96
    int mn = 0, mx = setSize;
    for (int currLevel = 0; currLevel <= level && mn <= mx; currLevel++) {
      int m = (mx + mn) / 2;
      char ch = binaryId.at(static_cast<size_t>(currLevel));
      if (ch == '0') {
        if (currLevel == level)
          mn = m;
        else
          mx = m;
      } else if (ch == '1') {
        if (currLevel == level)
          mx = m;
        else
          mn = m;
      }
      if (mx == mn) break;
      if (mx - 1 == 0 || mn == setSize) break;
    }
    return {mn, mx};
  }
  bool isCandidate(int node, int level) const {  // :99-101
    auto c = getCandidateSet(level);
    return node >= c.first && node < c.second;
  }
  std::vector<int> pickNextNodes(int level, int howMany) {  // :123-157
    auto cs = getCandidateSet(level);
    std::vector<int> candidateSet;
    for (int i = cs.first; i < cs.second; ++i) candidateSet.push_back(i);
    auto own = getOwnSet(level);
    int idx = (nodeId >= own.first && nodeId < own.second) ? nodeId - own.first : -1;
    if (idx == -1 || (own.second - own.first) < idx) throw IllegalState("pickNextNodes");
    std::vector<int> newList;
    JBitSet& set = usedNodes[level];  // getOrDefault + put: the same object is kept per level
    if (!set.get(idx)) {
      newList.push_back(candidateSet.at(static_cast<size_t>(idx)));
      candidateSet.erase(candidateSet.begin() + idx);
      set.set(idx);
    }
    int taken = 0;
    for (int i = 0; i < static_cast<int>(candidateSet.size()) && taken < howMany; ++i) {
      if (!set.get(i)) {
        set.set(i);
        newList.push_back(candidateSet[static_cast<size_t>(i)]);
        ++taken;
      }
    }
    javaShuffle(newList, *rd);
    return newList;
  }
};

// ----------------------------------------------------------------------------------------
// SanFerminSignature  (protocols/SanFerminSignature.java)
// ----------------------------------------------------------------------------------------
struct SanFerminSignature {
  struct Params {  // :41-110
    int nodeCount = 32768 / 32, powerOfTwo = 10, threshold = 32768 / 32, pairingTime = 2, signatureSize = 48, replyTimeout = 300;
    int candidateCount = 1;
    bool shuffledLists = false;
    std::string nodeBuilderName, networkLatencyName;
    bool latencyNull = true;
  };
  enum Status { OK, NO };
  struct SanFerminNode;
  struct SwapReply : Message {  // :518-541
    SanFerminSignature* p;
    Status status;
    int level, aggValue;
    SwapReply(SanFerminSignature* pp, Status s, int l, int a) : p(pp), status(s), level(l), aggValue(a) {}
    void action(Network&, Node& from, Node& to) override;
    int size() const override { return 4 + p->params.signatureSize; }
  };
  struct SwapRequest : Message {  // :543-564
    SanFerminSignature* p;
    int level, aggValue;
    SwapRequest(SanFerminSignature* pp, int l, int a) : p(pp), level(l), aggValue(a) {}
    void action(Network&, Node& from, Node& to) override;
    int size() const override { return 4 + p->params.signatureSize; }
  };
  struct SanFerminNode : Node {  // :148-511
    SanFerminSignature* p;
    std::string binaryId;
    int currentPrefixLength;
    SanFerminHelper candidateTree;
    std::unordered_map<int, int> signatureCache, futurSigs;
    std::unordered_set<int> pendingNodes;
    bool isSwapping = false;
    int aggValue = 1;
    int64_t thresholdAt = 0;
    bool thresholdDone = false, done = false;
    int sentRequests = 0, receivedRequests = 0;

    explicit SanFerminNode(SanFerminSignature* pp)
        : Node(pp->network.rd, pp->nb), p(pp), binaryId(SanFerminHelper::toBinaryID(nodeId, pp->params.nodeCount)),
          currentPrefixLength(pp->params.powerOfTwo) {}

    void onSwapRequest(SanFerminNode& node, const SwapRequest& request) {  // :229-268
      receivedRequests++;
      if (done || request.level != currentPrefixLength) {
        auto it = signatureCache.find(request.level);
        if (it != signatureCache.end()) {
          sendSwapReply(node, OK, request.level, it->second);
        } else {
          sendSwapReply(node, NO, currentPrefixLength, 0);
          bool isCandidate = candidateTree.isCandidate(node.nodeId, request.level);
          if (isCandidate) signatureCache[request.level] = request.aggValue;
        }
        return;
      }
      if (isSwapping) {
        sendSwapReply(node, OK, request.level, aggValue);
        return;
      }
      bool isCandidate = candidateTree.isCandidate(node.nodeId, currentPrefixLength);
      bool goodLevel = request.level == currentPrefixLength;
      if (isCandidate && goodLevel) transition(request.aggValue);
    }
    void onSwapReply(SanFerminNode& from, const SwapReply& reply) {  // :270-323
      if (reply.level != currentPrefixLength || done) return;
      if (isSwapping) return;
      switch (reply.status) {
        case OK:
          if (!pendingNodes.count(from.nodeId)) {
            bool isCandidate = candidateTree.isCandidate(from.nodeId, currentPrefixLength);
            bool goodLevel = reply.level == currentPrefixLength;
            if (isCandidate && goodLevel) transition(reply.aggValue);
            return;
          }
          transition(reply.aggValue);
          break;
        case NO:
          if (pendingNodes.count(from.nodeId)) {
            std::vector<int> nodes = candidateTree.pickNextNodes(currentPrefixLength, p->params.candidateCount);
            sendToNodes(nodes);
          }
          break;
      }
    }
    void sendToNodes(const std::vector<int>& candidates) {  // :329-373
      if (candidates.empty()) return;
      for (int c : candidates) pendingNodes.insert(c);
      sentRequests += static_cast<int>(candidates.size());
      auto r = std::make_shared<SwapRequest>(p, currentPrefixLength, aggValue);
      std::vector<Node*> dests;
      for (int c : candidates) dests.push_back(p->nodes[static_cast<size_t>(c)].get());
      p->network.send(r, *this, dests);
      int currLevel = currentPrefixLength;
      SanFerminNode* self = this;
      p->network.registerTask(
          [self, currLevel] {
            if (!self->done && self->currentPrefixLength == currLevel) {
              std::vector<int> newList = self->candidateTree.pickNextNodes(self->currentPrefixLength, self->p->params.candidateCount);
              self->sendToNodes(newList);
            }
          },
          p->network.time + p->params.replyTimeout, *this);
    }
    void goNextLevel() {  // :383-423
      if (done) return;
      bool enoughSigs = aggValue >= p->params.threshold;
      bool noMoreSwap = currentPrefixLength == 0;
      if (enoughSigs && !thresholdDone) {
        thresholdDone = true;
        thresholdAt = p->network.time + p->params.pairingTime * 2;
      }
      if (noMoreSwap && !done) {
        doneAt = p->network.time + p->params.pairingTime * 2;
        done = true;
        return;
      }
      currentPrefixLength--;
      signatureCache[currentPrefixLength] = aggValue;
      isSwapping = false;
      pendingNodes.clear();
      auto fs = futurSigs.find(currentPrefixLength);
      if (fs != futurSigs.end()) {
        aggValue += fs->second;
        goNextLevel();
        return;
      }
      std::vector<int> newList = candidateTree.pickNextNodes(currentPrefixLength, p->params.candidateCount);
      sendToNodes(newList);
    }
    void sendSwapReply(SanFerminNode& n, Status s, int level, int value) {  // :425-432
      auto r = std::make_shared<SwapReply>(p, s, level, value);
      std::vector<Node*> d{&n};
      p->network.send(r, *this, d);
    }
    void transition(int toAggregate) {  // :438-455
      isSwapping = true;
      SanFerminNode* self = this;
      p->network.registerTask(
          [self, toAggregate] {
            self->aggValue += toAggregate;
            self->goNextLevel();
          },
          p->network.time + p->params.pairingTime, *this);
    }
  };

  Params params;
  Network network;
  NodeBuilder nb;
  std::vector<std::unique_ptr<SanFerminNode>> nodes;

  // nodes are built in the constructor (:112-129): a later rd.setSeed() does not affect them
  explicit SanFerminSignature(const Params& pr) : params(pr) {
    params.powerOfTwo = javaLog2(params.nodeCount);
    nb = nodeBuilderByName(params.nodeBuilderName);
    network.setNetworkLatency(networkLatencyByName(params.networkLatencyName, params.latencyNull));
    for (int i = 0; i < params.nodeCount; i++) {
      nodes.push_back(std::make_unique<SanFerminNode>(this));
      network.addNode(nodes.back().get());
    }
    for (auto& n : nodes) n->candidateTree = SanFerminHelper(n->nodeId, params.nodeCount, &network.rd);
  }
  void init() {  // :136-138
    for (auto& up : nodes) {
      SanFerminNode* n = up.get();
      network.registerTask([n] { n->goNextLevel(); }, 1, *n);
    }
  }
};
inline void SanFerminSignature::SwapReply::action(Network&, Node& from, Node& to) {
  static_cast<SanFerminNode&>(to).onSwapReply(static_cast<SanFerminNode&>(from), *this);
}
inline void SanFerminSignature::SwapRequest::action(Network&, Node& from, Node& to) {
  static_cast<SanFerminNode&>(to).onSwapRequest(static_cast<SanFerminNode&>(from), *this);
}

// ----------------------------------------------------------------------------------------
// SanFerminCappos  (protocols/SanFerminCappos.java)
// ----------------------------------------------------------------------------------------
struct SanFerminCappos {
  struct Params {  // :43-104
    int nodeCount = 32768 / 16, pairingTime = 2, signatureSize = 48, candidateCount = 50, threshold = 32768 / 32, timeout = 150;
    std::string nodeBuilderName, networkLatencyName;
    bool latencyNull = true;
  };
  struct SanFerminNode;
  struct Swap : Message {  // :439-463
    SanFerminCappos* p;
    bool wantReply;
    int level, aggValue;
    Swap(SanFerminCappos* pp, int l, int a, bool r) : p(pp), wantReply(r), level(l), aggValue(a) {}
    void action(Network&, Node& from, Node& to) override;
    int size() const override { return 4 + p->params.signatureSize; }
  };
  struct SanFerminNode : Node {  // :144-437
    SanFerminCappos* p;
    SanFerminHelper helper;
    int currentPrefixLength;
    std::map<int, std::vector<int>> signatureCache;  // HashMap<Integer, List<Integer>>: only containsKey / max / filtered sums are observed
    bool isSwapping = false;
    int aggValue = 1;
    int64_t thresholdAt = 0;
    bool thresholdDone = false, done = false;

    explicit SanFerminNode(SanFerminCappos* pp)
        : Node(pp->network.rd, pp->nb), p(pp), currentPrefixLength(javaLog2(pp->params.nodeCount)) {}

    void onSwap(SanFerminNode& from, const Swap& swap) {  // :201-241
      bool wantReply = swap.wantReply;
      if (done || swap.level != currentPrefixLength) {
        bool isValueCached = signatureCache.count(swap.level) != 0;
        if (wantReply && isValueCached) {
          sendSwap({&from}, swap.level, getBestCachedSig(swap.level), false);
        } else {
          if (helper.isCandidate(from.nodeId, swap.level)) putCachedSig(swap.level, swap.aggValue);
        }
        return;
      }
      if (wantReply) sendSwap({&from}, swap.level, totalNumberOfSigs(swap.level), false);
      bool isCandidate = helper.isCandidate(from.nodeId, currentPrefixLength);
      if (isCandidate && !isSwapping) transition(swap.level, swap.aggValue);
    }
    void tryNextNodes(const std::vector<int>& candidates) {  // :248-296
      if (candidates.empty()) return;
      for (int c : candidates)
        if (!helper.isCandidate(c, currentPrefixLength)) throw IllegalState("tryNextNodes: not a candidate");
      std::vector<Node*> dests;
      for (int c : candidates) dests.push_back(p->nodes[static_cast<size_t>(c)].get());
      sendSwap(dests, currentPrefixLength, totalNumberOfSigs(currentPrefixLength + 1), true);
      int currLevel = currentPrefixLength;
      p->network.registerTask(
          [this, currLevel] {
            if (!done && currentPrefixLength == currLevel) tryNextNodes(helper.pickNextNodes(currentPrefixLength, p->params.candidateCount));
          },
          p->network.time + p->params.timeout, *this);
    }
    void goNextLevel() {  // :306-344
      if (done) return;
      bool enoughSigs = totalNumberOfSigs(currentPrefixLength) >= p->params.threshold;
      bool noMoreSwap = currentPrefixLength == 0;
      if (enoughSigs && !thresholdDone) {
        thresholdDone = true;
        thresholdAt = p->network.time + p->params.pairingTime * 2;
      }
      if (noMoreSwap && !done) {
        doneAt = p->network.time + p->params.pairingTime * 2;
        p->finishedNodes.push_back(this);
        done = true;
        return;
      }
      currentPrefixLength--;
      isSwapping = false;
      if (signatureCache.count(currentPrefixLength)) {
        goNextLevel();
        return;
      }
      tryNextNodes(helper.pickNextNodes(currentPrefixLength, p->params.candidateCount));
    }
    void sendSwap(const std::vector<Node*>& nodes, int level, int value, bool wantReply) {  // :346-349
      p->network.send(std::make_shared<Swap>(p, level, value, wantReply), *this, nodes);
    }
    int totalNumberOfSigs(int level) const {  // :351-358
      int sum = 0;
      for (auto& kv : signatureCache)
        if (kv.first >= level) sum += *std::max_element(kv.second.begin(), kv.second.end());
      return sum + 1;
    }
    void transition(int level, int toAggregate) {  // :364-374
      isSwapping = true;
      p->network.registerTask(
          [this, level, toAggregate] {
            putCachedSig(level, toAggregate);
            goNextLevel();
          },
          p->network.time + p->params.pairingTime, *this);
    }
    int getBestCachedSig(int level) const {  // :376-380
      const auto& v = signatureCache.at(level);
      return *std::max_element(v.begin(), v.end());
    }
    void putCachedSig(int level, int value) {  // :382-393
      signatureCache[level].push_back(value);
      bool enoughSigs = totalNumberOfSigs(currentPrefixLength) >= p->params.threshold;
      if (enoughSigs && !thresholdDone) {
        thresholdDone = true;
        thresholdAt = p->network.time + p->params.pairingTime * 2;
      }
    }
  };

  Params params;
  Network network;
  NodeBuilder nb;
  std::vector<std::unique_ptr<SanFerminNode>> nodes;
  std::vector<SanFerminNode*> finishedNodes;

  explicit SanFerminCappos(const Params& pr) : params(pr) {  // :106-112
    nb = nodeBuilderByName(pr.nodeBuilderName);
    network.setNetworkLatency(networkLatencyByName(pr.networkLatencyName, pr.latencyNull));
  }
  void init() {  // :120-134
    for (int i = 0; i < params.nodeCount; i++) {
      nodes.push_back(std::make_unique<SanFerminNode>(this));
      network.addNode(nodes.back().get());
    }
    for (auto& n : nodes) n->helper = SanFerminHelper(n->nodeId, params.nodeCount, &network.rd);
    for (auto& up : nodes) {
      SanFerminNode* n = up.get();
      network.registerTask([n] { n->goNextLevel(); }, 1, *n);
    }
  }
};
inline void SanFerminCappos::Swap::action(Network&, Node& from, Node& to) {
  static_cast<SanFerminNode&>(to).onSwap(static_cast<SanFerminNode&>(from), *this);
}

// ----------------------------------------------------------------------------------------
// Handel  (protocols/Handel.java), HiddenByzantine (:840-917) included.
// ----------------------------------------------------------------------------------------
struct Handel {
  struct Params {  // :22-142
    int nodeCount = 32, threshold = 31, pairingTime = 3, levelWaitTime = 50, extraCycle = 10, disseminationPeriodMs = 10;
    int fastPath = 10, nodesDown = 0;
    std::string nodeBuilderName, networkLatencyName;
    bool latencyNull = false;
    int desynchronizedStart = 0;
    bool byzantineSuicide = false, hiddenByzantine = false;
    int windowInitial = 16, windowMinimum = 1, windowMaximum = 128;  // WindowParameters() :157-159
  };
  static void validate(const Params& p) {  // :112-124
    if (p.nodesDown >= p.nodeCount || p.nodesDown < 0 || p.threshold > p.nodeCount || (p.nodesDown + p.threshold > p.nodeCount))
      throw IllegalArgument("nodeCount/threshold");
    if (__builtin_popcount(static_cast<unsigned>(p.nodeCount)) != 1) throw IllegalArgument("We support only power of two nodes in this simulation");
    if (p.byzantineSuicide && p.hiddenByzantine) throw IllegalArgument("Only one attack at a time");
  }
  // WindowParameters.newSize + ScoringExp(2, 4) :168-200
  int windowNewSize(int curr, bool correct) const {
    int updated = correct ? static_cast<int>(std::ceil(static_cast<double>(curr) * 2.0)) : static_cast<int>(std::floor(static_cast<double>(curr) / 4.0));
    if (updated > params.windowMaximum) return params.windowMaximum;
    if (updated < params.windowMinimum) return params.windowMinimum;
    return updated;
  }

  struct HNode;
  struct HLevel;
  struct SigToVerify {  // :919-938
    int from, level, rank;
    JBitSet sig;
    bool badSig;
  };
  using SigPtr = std::shared_ptr<SigToVerify>;
  struct SendSigs : Message {  // :239-276
    int level;
    JBitSet sigs;
    bool levelFinished;
    int size_;
    bool badSig = false;
    SendSigs(const JBitSet& s, const HLevel& l);
    int size() const override { return size_; }
    void action(Network&, Node& from, Node& to) override;
  };
  struct HLevel {  // :364-643
    HNode* node = nullptr;
    int level = 0, size = 0;
    std::vector<int> peers;  // node ids in emission order
    JBitSet waitedSigs, lastAggVerified, totalIncoming, verifiedIndSignatures, toVerifyInd, finishedPeers, totalOutgoing;
    std::vector<SigPtr> toVerifyAgg;
    bool outgoingFinished = false;
    int posInLevel = 0;
    int suicideBizAfter = -1;

    int expectedSigs() const { return size; }
    bool incomingComplete() const { return waitedSigs.equals(totalIncoming); }     // :520-522
    bool outgoingComplete() const { return totalOutgoing.cardinality() == size; }  // :524-526
    bool isOpen() const;                                                           // :454-468
    void doCycle();                                                                // :470-480
    std::vector<int> getRemainingPeers(int peersCt);                               // :482-504
    int sizeIfIncluded(const SigToVerify& sig) const {                             // :528-536
      JBitSet c = sig.sig;
      if (!c.intersects(totalIncoming)) c.or_(totalIncoming);
      c.or_(verifiedIndSignatures);
      return c.cardinality();
    }
    SigPtr createSuicideByzantineSig(int maxRank);  // :538-559
    SigPtr bestToVerify();                          // :566-630
  };
  struct HNode : Node {  // :278-838
    Handel* p;
    int startAt;
    std::vector<HLevel> levels;
    int nodePairingTime;
    std::vector<int> receptionRanks;
    JBitSet blacklist;
    int currWindowSize, addedCycle;
    bool done = false;
    int sigsChecked = 0, sigQueueSize = 0, msgFiltered = 0;
    // HiddenByzantine (:840-917): one per honest node when params.hiddenByzantine (:303)
    bool hasHidden = false, hbNoByzantinePeers = false;
    SigPtr hbLast;
    HNode* firstByzantine(HLevel& l);            // :844-858
    SigPtr attack(const SigPtr& currentBest);    // :861-916

    HNode(Handel* pp, int startAt_)
        : Node(pp->network.rd, pp->nb), p(pp), startAt(startAt_),
          nodePairingTime(static_cast<int>(std::max(1.0, pp->params.pairingTime * speedRatio))),
          receptionRanks(static_cast<size_t>(pp->params.nodeCount), 0), currWindowSize(pp->params.windowInitial),
          addedCycle(pp->params.extraCycle) {}
    JBitSet allSigsAtLevel(int round) const {  // :667-680
      JBitSet res;
      int cMask = (1 << round) - 1;
      int start = (cMask | nodeId) ^ cMask;
      int end = std::min(nodeId | cMask, p->params.nodeCount - 1);
      res.setRange(start, end + 1);
      res.set(nodeId, false);
      return res;
    }
    void initLevel() {  // :319-329, HLevel ctors :409-431
      int rounded = roundPow2(p->params.nodeCount);
      JBitSet allPreviousNodes;
      levels.reserve(40);
      levels.emplace_back();
      HLevel& l0 = levels.back();
      l0.node = this;
      l0.level = 0;
      l0.size = 1;
      l0.outgoingFinished = true;
      l0.lastAggVerified.set(nodeId);
      l0.verifiedIndSignatures.set(nodeId);
      l0.totalIncoming.set(nodeId);
      l0.suicideBizAfter = p->params.byzantineSuicide ? 0 : -1;
      for (int l = 1; (1LL << l) <= rounded; l++) {
        allPreviousNodes.or_(levels.back().waitedSigs);
        HLevel nl;
        nl.node = this;
        nl.level = levels.back().level + 1;
        nl.waitedSigs.or_(allSigsAtLevel(nl.level));
        nl.waitedSigs.andNot(allPreviousNodes);
        nl.totalOutgoing.set(nodeId);
        nl.size = nl.waitedSigs.cardinality();
        nl.suicideBizAfter = p->params.byzantineSuicide ? 0 : -1;
        levels.push_back(std::move(nl));
      }
    }
    void dissemination() {  // :331-343
      if (doneAt > 0) {
        if (addedCycle > 0)
          addedCycle--;
        else
          return;
      }
      for (HLevel& sfl : levels) sfl.doCycle();
    }
    bool hasSigToVerify() const { return sigQueueSize != 0; }
    int totalSigSize() const {  // :349-352
      const HLevel& last = levels.back();
      return last.totalOutgoing.cardinality() + last.totalIncoming.cardinality();
    }
    int score(const HLevel& l, const JBitSet& sig) const {  // :651-664
      if (l.lastAggVerified.cardinality() >= l.expectedSigs()) return 0;
      if (!l.lastAggVerified.intersects(sig)) return l.lastAggVerified.cardinality() + sig.cardinality();
      JBitSet withIndiv = l.verifiedIndSignatures;
      withIndiv.or_(sig);
      return std::max(0, withIndiv.cardinality() - l.lastAggVerified.cardinality());
    }
    static bool include(const JBitSet& big, const JBitSet& small) {  // BitSetUtils.include
      JBitSet b = small;
      b.or_(big);
      return b.equals(big);
    }
    void updateVerifiedSignatures(const SigPtr& vs);  // :686-750
    void onNewSig(HNode& from, const SendSigs& ssigs);  // :753-786
    void checkSigs();                                   // :792-837
  };

  Params params;
  Network network;
  NodeBuilder nb;
  std::vector<std::unique_ptr<HNode>> nodes;

  explicit Handel(const Params& pr) : params(pr) {  // :212-216
    validate(params);
    network.setNetworkLatency(networkLatencyByName(params.networkLatencyName, params.latencyNull));
  }
  HNode& node(int i) { return *nodes[static_cast<size_t>(i)]; }

  void setReceivingRanks() {  // :940-948
    std::vector<int> expected(static_cast<size_t>(params.nodeCount));
    for (int i = 0; i < params.nodeCount; ++i) expected[static_cast<size_t>(i)] = i;
    for (auto& n : nodes) {
      javaShuffle(expected, network.rd);
      for (size_t i = 0; i < expected.size(); i++) n->receptionRanks[static_cast<size_t>(expected[i])] = static_cast<int>(i);
    }
  }
  void init() {  // :957-1014
    nb = nodeBuilderByName(params.nodeBuilderName);
    std::vector<bool> badNodes = Network::chooseBadNodes(network.rd, params.nodeCount, params.nodesDown);
    for (int i = 0; i < params.nodeCount; i++) {
      int startAt = params.desynchronizedStart == 0 ? 0 : network.rd.nextInt(params.desynchronizedStart);
      nodes.push_back(std::make_unique<HNode>(this, startAt));
      if (badNodes[static_cast<size_t>(i)]) nodes.back()->stop();
      nodes.back()->hasHidden = params.hiddenByzantine && !badNodes[static_cast<size_t>(i)];  // :303, :968
      network.addNode(nodes.back().get());
    }
    for (auto& up : nodes) {
      HNode* n = up.get();
      n->initLevel();
      if (!n->isDown()) {
        network.registerPeriodicTask([n] { n->dissemination(); }, n->startAt + 1, params.disseminationPeriodMs, *n);
        network.registerConditionalTask([n] { n->checkSigs(); }, n->startAt + 1, n->nodePairingTime, *n,
                                        [n] { return n->hasSigToVerify(); }, [n] { return !n->done; });
      }
    }
    setReceivingRanks();
    // emission lists: contact first the peers that gave you a good reception rank (:991-1013).
    // The reference allocates a nodeCount-sized List[] per (sender, level); here one array is reused and only the
    // touched ranks are visited (in increasing rank order, like the array walk of buildEmissionList) — same result.
    std::vector<std::vector<int>> emissionList(static_cast<size_t>(params.nodeCount));
    std::vector<int> touched;
    for (auto& up : nodes) {
      HNode* sender = up.get();
      if (sender->isDown()) continue;
      for (HLevel& l : sender->levels) {
        touched.clear();
        for (int cur = l.waitedSigs.nextSetBit(0); cur >= 0; cur = l.waitedSigs.nextSetBit(cur + 1)) {
          int recRank = node(cur).receptionRanks[static_cast<size_t>(sender->nodeId)];
          if (emissionList[static_cast<size_t>(recRank)].empty()) touched.push_back(recRank);
          emissionList[static_cast<size_t>(recRank)].push_back(cur);
        }
        if (!l.peers.empty()) throw IllegalState("peers not empty");  // buildEmissionList :506-518
        std::sort(touched.begin(), touched.end());
        for (int rk : touched) {
          auto& ranks = emissionList[static_cast<size_t>(rk)];
          if (ranks.size() > 1) javaShuffle(ranks, network.rd);
          l.peers.insert(l.peers.end(), ranks.begin(), ranks.end());
          ranks.clear();
        }
      }
    }
  }
  bool continueIf() const {  // Handel.newContIf :1044-1053
    for (auto& n : nodes)
      if (!n->isDown() && (n->doneAt == 0 || n->addedCycle > 0)) return true;
    return false;
  }
};

inline Handel::SendSigs::SendSigs(const JBitSet& s, const HLevel& l)  // :253-265
    : level(l.level), sigs(s), levelFinished(l.incomingComplete()), size_(1 + l.expectedSigs() / 8 + 96 * 2) {
  if (sigs.isEmpty() || sigs.cardinality() > l.size) throw IllegalState("bad level");
}
inline void Handel::SendSigs::action(Network&, Node& from, Node& to) {
  static_cast<HNode&>(to).onNewSig(static_cast<HNode&>(from), *this);
}
inline bool Handel::HLevel::isOpen() const {
  if (outgoingFinished) return false;
  if (node->p->network.time >= (level - 1) * node->p->params.levelWaitTime) return true;
  if (outgoingComplete()) return true;
  return false;
}
inline void Handel::HLevel::doCycle() {
  if (!isOpen()) return;
  std::vector<int> dest = getRemainingPeers(1);
  if (!dest.empty()) {
    auto ss = std::make_shared<SendSigs>(totalOutgoing, *this);
    node->p->network.send(ss, *node, node->p->node(dest[0]));
  }
}
inline std::vector<int> Handel::HLevel::getRemainingPeers(int peersCt) {
  std::vector<int> res;
  int start = posInLevel;
  while (peersCt > 0 && !outgoingFinished) {
    int pid = peers.at(static_cast<size_t>(posInLevel++));
    if (posInLevel >= static_cast<int>(peers.size())) posInLevel = 0;
    if (!finishedPeers.get(pid) && !node->blacklist.get(pid)) {
      res.push_back(pid);
      peersCt--;
    } else {
      if (posInLevel == start) outgoingFinished = true;
    }
  }
  return res;
}
inline Handel::SigPtr Handel::HLevel::createSuicideByzantineSig(int maxRank) {
  bool reset = false;
  for (int i = suicideBizAfter; i < static_cast<int>(peers.size()); i++) {
    int pid = peers[static_cast<size_t>(i)];
    if (node->p->node(pid).isDown() && !node->blacklist.get(pid)) {
      if (!reset) {
        suicideBizAfter = i;
        reset = true;
      }
      if (node->receptionRanks[static_cast<size_t>(pid)] < maxRank)
        return std::make_shared<SigToVerify>(SigToVerify{pid, level, node->receptionRanks[static_cast<size_t>(pid)], waitedSigs, true});
    }
  }
  if (!reset) suicideBizAfter = -1;
  return nullptr;
}
inline Handel::SigPtr Handel::HLevel::bestToVerify() {
  if (toVerifyAgg.empty()) return nullptr;
  if (node->currWindowSize < 1) throw IllegalState("window");
  int windowIndex = toVerifyAgg[0]->rank;  // Collections.min(..., comparingInt(getRank)).rank
  for (auto& s : toVerifyAgg) windowIndex = std::min(windowIndex, s->rank);
  if (suicideBizAfter >= 0) {
    SigPtr bSig = createSuicideByzantineSig(windowIndex + node->currWindowSize);
    if (bSig) {
      toVerifyAgg.push_back(bSig);
      node->sigQueueSize++;
      return bSig;
    }
  }
  int curSignatureSize = totalIncoming.cardinality();
  SigPtr bestOutside, bestInside;
  int bestScoreInside = 0;
  int removed = 0;
  std::vector<SigPtr> curatedList;
  for (auto& stv : toVerifyAgg) {
    int s = sizeIfIncluded(*stv);
    if (!node->blacklist.get(stv->from) && s > curSignatureSize) {
      curatedList.push_back(stv);
      if (stv->rank <= windowIndex + node->currWindowSize) {
        int sc = node->score(*this, stv->sig);
        if (sc > bestScoreInside) {
          bestScoreInside = sc;
          bestInside = stv;
        }
      } else {
        if (!bestOutside || stv->rank < bestOutside->rank) bestOutside = stv;
      }
    } else {
      removed++;
    }
  }
  if (removed > 0) {  // replaceToVerifyAgg :632-642
    int oldSize = static_cast<int>(toVerifyAgg.size());
    toVerifyAgg = curatedList;
    node->sigQueueSize -= oldSize;
    node->sigQueueSize += static_cast<int>(toVerifyAgg.size());
    if (node->sigQueueSize < 0) throw IllegalState("sigQueueSize<0");
  }
  if (bestInside) return bestInside;
  if (bestOutside) return bestOutside;
  return nullptr;
}
inline void Handel::HNode::updateVerifiedSignatures(const SigPtr& vs) {
  if (vs->badSig) {
    blacklist.set(vs->from);
    if (!p->params.byzantineSuicide) throw IllegalState("We should not have invalid signatures in this scenario");
    return;
  }
  HLevel& vsl = levels[static_cast<size_t>(vs->level)];
  if (!include(vsl.waitedSigs, vs->sig)) throw IllegalState("bad signature received");
  vsl.toVerifyInd.set(vs->from, false);
  for (size_t i = 0; i < vsl.toVerifyAgg.size(); ++i)  // toVerifyAgg.remove(vs): identity
    if (vsl.toVerifyAgg[i] == vs) {
      vsl.toVerifyAgg.erase(vsl.toVerifyAgg.begin() + static_cast<long>(i));
      break;
    }
  vsl.verifiedIndSignatures.set(vs->from);
  bool improved = false;
  if (!vsl.totalIncoming.get(vs->from)) {
    vsl.totalIncoming.set(vs->from);
    improved = true;
  }
  JBitSet all = vs->sig;
  all.or_(vsl.verifiedIndSignatures);
  if (all.cardinality() > vsl.verifiedIndSignatures.cardinality()) {
    improved = true;
    if (vsl.lastAggVerified.intersects(vs->sig)) vsl.lastAggVerified = JBitSet();
    vsl.lastAggVerified.or_(vs->sig);
    vsl.totalIncoming = JBitSet();
    vsl.totalIncoming.or_(vsl.lastAggVerified);
    vsl.totalIncoming.or_(vsl.verifiedIndSignatures);
  }
  if (!improved) return;
  bool justCompleted = vsl.incomingComplete();
  JBitSet cur;
  for (HLevel& l : levels) {
    if (l.level > vsl.level) {
      l.totalOutgoing = JBitSet();
      l.totalOutgoing.or_(cur);
      if (justCompleted && p->params.fastPath > 0 && !l.outgoingFinished && l.outgoingComplete()) {
        std::vector<int> peers = l.getRemainingPeers(p->params.fastPath);
        auto sendSigs = std::make_shared<SendSigs>(l.totalOutgoing, l);
        std::vector<Node*> dests;
        for (int pid : peers) dests.push_back(&p->node(pid));
        p->network.send(sendSigs, *this, dests);
      }
    }
    cur.or_(l.totalIncoming);
  }
  if (doneAt == 0 && cur.cardinality() >= p->params.threshold) doneAt = p->network.time;
}
inline void Handel::HNode::onNewSig(HNode& from, const SendSigs& ssigs) {
  if (doneAt > 0) {
    msgFiltered++;
    return;
  }
  if (p->network.time < startAt || blacklist.get(from.nodeId)) return;
  HLevel& l = levels[static_cast<size_t>(ssigs.level)];
  if (!include(l.waitedSigs, ssigs.sigs)) throw IllegalState("bad signatures received");
  JBitSet cs = ssigs.sigs;
  cs.and_(l.waitedSigs);
  if (!cs.equals(ssigs.sigs) || ssigs.sigs.isEmpty()) throw IllegalState("bad message");
  if (ssigs.levelFinished) l.finishedPeers.set(from.nodeId);
  if (!l.verifiedIndSignatures.get(from.nodeId)) l.toVerifyInd.set(from.nodeId);
  sigQueueSize++;
  l.toVerifyAgg.push_back(std::make_shared<SigToVerify>(
      SigToVerify{from.nodeId, l.level, receptionRanks[static_cast<size_t>(from.nodeId)], cs, ssigs.badSig}));
}
inline void Handel::HNode::checkSigs() {
  std::vector<SigPtr> byLevels;
  for (HLevel& l : levels) {
    SigPtr ss = l.bestToVerify();
    if (!ss) continue;
    byLevels.push_back(ss);
  }
  if (byLevels.empty()) return;
  SigPtr best = byLevels[static_cast<size_t>(p->network.rd.nextInt(static_cast<int>(byLevels.size())))];  // :788-790
  if (hasHidden && best->level == static_cast<int>(levels.size()) - 1) best = attack(best);  // :813-817
  HLevel& l = levels[static_cast<size_t>(best->level)];
  int newSize = p->windowNewSize(currWindowSize, !best->badSig);
  currWindowSize = std::min(newSize, l.size);
  int& rk = receptionRanks[static_cast<size_t>(best->from)];
  rk = static_cast<int>(static_cast<uint32_t>(rk) + static_cast<uint32_t>(p->params.nodeCount));
  if (rk < 0) rk = std::numeric_limits<int>::max();
  sigsChecked++;
  HNode* self = this;
  p->network.registerTask([self, best] { self->updateVerifiedSignatures(best); }, p->network.time + nodePairingTime, *this);
}

inline Handel::HNode* Handel::HNode::firstByzantine(HLevel& l) {
  HNode* best = nullptr;
  int bestRank = std::numeric_limits<int>::max();
  for (int pid : l.peers) {
    HNode* pn = p->nodes[static_cast<size_t>(pid)].get();
    if (pn->isDown() && receptionRanks[static_cast<size_t>(pid)] < bestRank && !l.totalIncoming.get(pid)) {
      bestRank = receptionRanks[static_cast<size_t>(pid)];
      best = pn;
      if (bestRank == 0) return pn;
    }
  }
  return best;
}
inline Handel::SigPtr Handel::HNode::attack(const SigPtr& currentBest) {
  if (hbNoByzantinePeers) return currentBest;
  if (hbLast == currentBest) {  // a previous attack finally worked
    hbLast = nullptr;
    return currentBest;
  }
  HLevel& l = levels[static_cast<size_t>(currentBest->level)];
  if (hbLast) {
    for (auto& s : l.toVerifyAgg)
      if (s == hbLast) return currentBest;  // still in the list: the attack failed this time
    if (!l.totalIncoming.get(hbLast->from)) throw IllegalState("byz signature pruned!");
    hbLast = nullptr;
  }
  HNode* fb = firstByzantine(l);
  if (fb == nullptr) {
    hbNoByzantinePeers = true;
    return currentBest;
  }
  if (receptionRanks[static_cast<size_t>(fb->nodeId)] >= currentBest->rank) return currentBest;
  JBitSet cur;
  cur.set(fb->nodeId);
  SigPtr bad = std::make_shared<SigToVerify>(SigToVerify{fb->nodeId, l.level, receptionRanks[static_cast<size_t>(fb->nodeId)], cur, false});
  l.toVerifyAgg.push_back(bad);
  sigQueueSize++;
  SigPtr newBest = l.bestToVerify();
  if (newBest != bad) hbLast = bad;
  return newBest;
}

}  // namespace wo
