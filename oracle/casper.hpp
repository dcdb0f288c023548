// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of protocols/CasperIMD.java with core/Block.java, core/BlockChainNode.java and
// core/BlockChainNetwork.java.  Line references are to those files.
//
// PARITY STATUS: the fork-choice rule, block building, attestation bookkeeping and the schedule are pinned by the
// reference's own tests (PT/CasperIMDTest.java, PT/CasperByzantineTest.java, restated in oracle/test_casper_kat.cpp).
// The end state of a long run is "parity unpinned" (no golden state in the reference, no JVM here).
//
// Two places where the reference itself is not deterministic across JVM runs, and what this restatement does:
//   * `blocksToReevaluate` is a HashSet<CasperBlock> and CasperBlock has identity hashCode, so the order of the
//     best(head, b) folds in reevaluateHead (CasperIMD.java:348-353) is JVM-dependent.  Here: ascending block id.
//   * `blocksReceivedByHeight.get(h).iterator().next()` (:554, :623) picks an identity-hash-ordered element.  Here: lowest id.
// `Block.blockId` is a process-wide static in the reference (Block.java:10); here it is per protocol instance and starts
// at 1 — only the relative order of ids is ever observed (:255).
#pragma once
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "core.hpp"

namespace wo {

struct CasperIMD {
  struct Params {  // :18-71
    static constexpr int SLOT_DURATION = 8000;
    int cycleLength = 4;
    bool randomOnTies = true;
    int blockProducersCount = 2;
    int attestersPerRound = 20;
    int attestersCount = 80;
    int blockConstructionTime = 1000;
    int attestationConstructionTime = 1;
    std::string nodeBuilderName;  // empty == null
    std::string networkLatencyName;
    bool latencyNull = true;
  };
  static Params makeParams(int cycleLength, bool randomOnTies, int blockProducersCount, int attestersPerRound,
                           int blockConstructionTime, int attestationConstructionTime, const std::string& nb, const std::string& nl,
                           bool latencyNull) {
    Params p;
    p.cycleLength = cycleLength;
    p.randomOnTies = randomOnTies;
    p.blockProducersCount = blockProducersCount;
    p.attestersPerRound = attestersPerRound;
    p.attestersCount = attestersPerRound * cycleLength;
    p.blockConstructionTime = blockConstructionTime;
    p.attestationConstructionTime = attestationConstructionTime;
    p.nodeBuilderName = nb;
    p.networkLatencyName = nl;
    p.latencyNull = latencyNull;
    return p;
  }

  struct CasperNode;
  struct Attestation;
  struct AttLess {
    bool operator()(const Attestation* a, const Attestation* b) const;
  };
  using AttSet = std::set<Attestation*, AttLess>;

  // core/Block.java + CasperIMD.CasperBlock (:151-194)
  struct CasperBlock {
    int height = 0;
    int proposalTime = 0;
    int64_t lastTxId = 0;
    int64_t id = 0;
    CasperBlock* parent = nullptr;
    CasperNode* producer = nullptr;
    bool valid = true;
    std::map<int, AttSet> attestationsByHeight;

    CasperBlock() = default;  // genesis: Block(int h) with h = 0 (Block.java:24-32)
    CasperBlock(CasperIMD& ci, CasperNode* prod, int h, CasperBlock* father, std::map<int, AttSet> atts, int time)
        : attestationsByHeight(std::move(atts)) {  // Block.java:38-56
      if (h <= 0) throw IllegalArgument("Only the genesis block has a special height");
      if (father != nullptr && time < father->proposalTime) throw IllegalArgument("bad time");
      if (father != nullptr && father->height >= h) throw IllegalArgument("Bad parent");
      producer = prod;
      height = h;
      id = ci.blockId++;
      parent = father;
      valid = true;
      lastTxId = time;
      proposalTime = time;
    }
    int64_t txCount() const {  // Block.java:59-68
      if (id == 0) return 0;
      int64_t res = lastTxId - parent->lastTxId;
      if (res < 0) throw IllegalState("bad txCount");
      return res;
    }
    bool hasDirectLink(const CasperBlock* b) const {  // Block.java:87-100
      if (b == this) return true;
      if (b->height == height) return false;
      const CasperBlock* older = height > b->height ? this : b;
      const CasperBlock* young = height < b->height ? this : b;
      while (older->height > young->height) older = older->parent;
      return older == young;
    }
  };
  struct BlockLess {
    bool operator()(const CasperBlock* a, const CasperBlock* b) const { return a->id < b->id; }
  };
  using BlockSet = std::set<CasperBlock*, BlockLess>;

  // :105-149
  struct Attestation : Message {
    CasperNode* attester;
    int height;
    std::set<int64_t> hs;
    CasperBlock* head;
    int64_t seq;  // creation order (oracle only: a total order for the sets)
    Attestation(CasperIMD& ci, CasperNode* att, int h);
    void action(Network&, Node&, Node& to) override { static_cast<CasperNode&>(to).onAttestation(this); }
    bool attests(const CasperBlock* cb) const { return hs.count(cb->id) != 0; }
  };

  // BlockChainNetwork.SendBlock (BlockChainNetwork.java:22-40)
  struct SendBlock : Message {
    CasperBlock* toSend;
    explicit SendBlock(CasperBlock* b) : toSend(b) {}
    void action(Network&, Node&, Node& to) override { static_cast<CasperNode&>(to).onBlock(toSend); }
  };

  enum NodeKind { OBSERVER = 0, PRODUCER = 1, ATTESTER = 2, BYZ = 3, BYZ_SF = 4, BYZ_NS = 5, BYZ_WF = 6 };

  // BlockChainNode (BlockChainNode.java) + CasperNode (:196-363)
  struct CasperNode : Node {
    CasperIMD& ci;
    bool byzantine;
    NodeKind kind;
    CasperBlock* genesis;
    std::map<int64_t, CasperBlock*> blocksReceivedByBlockId;
    std::map<int64_t, BlockSet> blocksReceivedByFatherId;
    std::map<int, BlockSet> blocksReceivedByHeight;
    CasperBlock* head;
    std::map<int64_t, AttSet> attestationsByHead;
    BlockSet blocksToReevaluate;

    CasperNode(CasperIMD& c, bool byz, NodeKind k)
        : Node(c.network.rd, c.nb), ci(c), byzantine(byz), kind(k), genesis(&c.genesis), head(&c.genesis) {
      blocksReceivedByBlockId[genesis->id] = genesis;  // BlockChainNode.java:22-27
    }

    // :205-257
    CasperBlock* best(CasperBlock* o1, CasperBlock* o2) {
      if (o1 == o2) return o1;
      if (o1->height == o2->height) throw IllegalState("two blocks for the same height");
      if (o1->hasDirectLink(o2)) return o1->height < o2->height ? o2 : o1;
      CasperBlock* b1 = o1;
      CasperBlock* b2 = o2;
      while (b1->parent != b2->parent) {
        if (b1->parent->height > b2->parent->height)
          b1 = b1->parent;
        else
          b2 = b2->parent;
      }
      CasperBlock* h = b1->parent;
      int b1Votes = countAttestations(o1, h);
      int b2Votes = countAttestations(o2, h);
      if (b1Votes > b2Votes) return o1;
      if (b1Votes < b2Votes) return o2;
      if (ci.params.randomOnTies) {
        ++ci.network.statDraws;
        return ci.network.rd.nextBoolean() ? o1 : o2;
      }
      return b1->id >= b2->id ? o1 : o2;
    }

    // :262-288
    int countAttestations(CasperBlock* start, CasperBlock* h) {
      AttSet a1;
      for (CasperBlock* cur = start; cur != h; cur = cur->parent) {
        for (int i = cur->height - 1; i > h->height; i--) {
          auto it = cur->attestationsByHeight.find(i);
          if (it == cur->attestationsByHeight.end()) continue;
          for (Attestation* a : it->second)
            if (a->attests(h)) a1.insert(a);
        }
        auto it = attestationsByHead.find(cur->id);
        if (it != attestationsByHead.end())
          for (Attestation* a : it->second)
            if (a->attests(h)) a1.insert(a);
      }
      return static_cast<int>(a1.size());
    }

    // BlockChainNode.onBlock (BlockChainNode.java:33-49)
    bool baseOnBlock(CasperBlock* b) {
      if (!b->valid) return false;
      auto ins = blocksReceivedByBlockId.emplace(b->id, b);
      if (!ins.second) return false;  // already received
      blocksReceivedByFatherId[b->parent->id].insert(b);
      blocksReceivedByHeight[b->height].insert(b);
      head = best(head, b);
      return true;
    }
    // CasperNode.onBlock :298-314.  delta is never negative (time >= 0, height >= 1): the deferral branch is dead.
    virtual bool onBlock(CasperBlock* b) {
      const int delta = ci.network.time - genesis->proposalTime + b->height * Params::SLOT_DURATION;
      if (delta >= 0) {
        blocksToReevaluate.insert(head);
        blocksToReevaluate.insert(b);
        return baseOnBlock(b);
      }
      ci.network.registerTask([this, b] { onBlock(b); }, delta * -1, *this);
      return false;
    }
    // :316-337
    void onAttestation(Attestation* a) {
      attestationsByHead[a->head->id].insert(a);
      if (blocksReceivedByBlockId.count(a->head->id)) blocksToReevaluate.insert(a->head);
    }
    // :348-353 (iteration order: see the header)
    void reevaluateHead() {
      for (CasperBlock* b : blocksToReevaluate) head = best(head, b);
      blocksToReevaluate.clear();
    }
    virtual std::function<void()> periodicTask() { return nullptr; }

    // BlockChainNode.java:54-76
    int txsCreatedInChain(const CasperBlock* h) const {
      int txs = 0;
      for (const CasperBlock* cur = h; cur != nullptr; cur = cur->parent)
        if (cur->producer == this) txs += static_cast<int>(cur->txCount());
      return txs;
    }
    int blocksCreatedInChain(const CasperBlock* h) const {
      int blocks = 0;
      for (const CasperBlock* cur = h; cur != nullptr; cur = cur->parent)
        if (cur->producer == this) blocks++;
      return blocks;
    }
  };

  // :365-442
  struct BlockProducer : CasperNode {
    explicit BlockProducer(CasperIMD& c) : CasperNode(c, false, PRODUCER) {}
    BlockProducer(CasperIMD& c, bool byz, NodeKind k) : CasperNode(c, byz, k) {}
    std::function<void()> periodicTask() override {
      return [this] {
        reevaluateHead();
        createAndSendBlock(ci.network.time / Params::SLOT_DURATION);
      };
    }
    // :383-428
    CasperBlock* buildBlock(CasperBlock* base, int height) {
      std::map<int, AttSet> res;
      for (int i = height - 1; i >= 0 && i >= height - ci.params.cycleLength; i--) res[i];
      AttSet allFromBlocks;
      for (CasperBlock* cur = base; cur != genesis && cur->height >= height - ci.params.cycleLength; cur = cur->parent)
        for (auto& kv : cur->attestationsByHeight) allFromBlocks.insert(kv.second.begin(), kv.second.end());
      for (CasperBlock* cur = base; cur != nullptr && cur->height >= height - ci.params.cycleLength; cur = cur->parent) {
        auto it = attestationsByHead.find(cur->id);
        if (it == attestationsByHead.end()) continue;
        for (Attestation* a : it->second)
          if (a->height < height && !allFromBlocks.count(a)) res[a->height].insert(a);
      }
      ci.blocks.push_back(std::make_unique<CasperBlock>(ci, this, height, base, std::move(res), ci.network.time));
      return ci.blocks.back().get();
    }
    // :430-436
    void createAndSendBlock(int height) {
      head = buildBlock(head, height);
      ci.network.sendAll(std::make_shared<SendBlock>(head), ci.network.time + ci.params.blockConstructionTime, *this);
    }
  };

  // :444-470
  struct Attester : CasperNode {
    explicit Attester(CasperIMD& c) : CasperNode(c, false, ATTESTER) {}
    std::function<void()> periodicTask() override {
      return [this] { vote(ci.network.time / Params::SLOT_DURATION); };
    }
    void vote(int height) {
      reevaluateHead();
      auto v = std::make_shared<Attestation>(ci, this, height);
      ci.attestations.push_back(v);
      ci.network.sendAll(v, ci.network.time + ci.params.attestationConstructionTime, *this);
    }
  };

  // :511-580
  struct ByzBlockProducer : BlockProducer {
    int toSend = 1;
    int h = 0;
    int delay;
    int onDirectFather = 0, onOlderAncestor = 0, incNotTheBestFather = 0;
    ByzBlockProducer(CasperIMD& c, int d, NodeKind k = BYZ) : BlockProducer(c, true, k), delay(d) {}
    void reevaluateH(int time) {  // :529-542
      reevaluateHead();
      while (head->height >= toSend) head = head->parent;
      int slotTime = time - delay;
      h = slotTime / Params::SLOT_DURATION;
      if (h != toSend) throw IllegalState("h != toSend");
    }
    CasperBlock* firstAtHeight(int hh) {  // blocksReceivedByHeight.get(hh).iterator().next(); NPE when absent
      auto it = blocksReceivedByHeight.find(hh);
      if (it == blocksReceivedByHeight.end() || it->second.empty()) throw IllegalState("NullPointerException");
      return *it->second.begin();
    }
    std::function<void()> periodicTask() override {  // :544-564
      return [this] {
        reevaluateH(ci.network.time);
        if (head->height == h - 1) {
          onDirectFather++;
        } else {
          onOlderAncestor++;
          CasperBlock* possibleFather = firstAtHeight(h - 1);
          if (possibleFather != nullptr && possibleFather->parent->height != h - 1) incNotTheBestFather++;
        }
        createAndSendBlock(toSend);
        toSend += ci.params.blockProducersCount;
      };
    }
  };
  struct ByzBlockProducerSF : ByzBlockProducer {  // :583-604
    ByzBlockProducerSF(CasperIMD& c, int d) : ByzBlockProducer(c, d, BYZ_SF) {}
    std::function<void()> periodicTask() override {
      return [this] {
        reevaluateH(ci.network.time);
        if (head->id != 0 && head->height == h - 1) {
          head = head->parent;
          onDirectFather++;
        } else {
          onOlderAncestor++;
        }
        createAndSendBlock(toSend);
        toSend += ci.params.blockProducersCount;
      };
    }
  };
  struct ByzBlockProducerNS : ByzBlockProducer {  // :610-640
    int skipped = 0;
    ByzBlockProducerNS(CasperIMD& c, int d) : ByzBlockProducer(c, d, BYZ_NS) {}
    std::function<void()> periodicTask() override {
      return [this] {
        reevaluateH(ci.network.time);
        if (head->id != 0 && head->height == h - 1 && head->parent->height == h - 3) {
          CasperBlock* b = firstAtHeight(h - 2);
          if (b != nullptr) {
            head = b;
            skipped++;
          }
        }
        createAndSendBlock(toSend);
        toSend += ci.params.blockProducersCount;
      };
    }
  };
  struct ByzBlockProducerWF : ByzBlockProducer {  // :647-707
    int late = 0, onTime = 0;
    ByzBlockProducerWF(CasperIMD& c, int d) : ByzBlockProducer(c, d, BYZ_WF) {}
    std::function<void()> periodicTask() override {
      return [this] {
        if (head == genesis && toSend == 1) {
          reevaluateH(ci.network.time);
          createAndSendBlock(h);
          toSend += ci.params.blockProducersCount;
        }
      };
    }
    bool onBlock(CasperBlock* b) override {  // :667-701
      if (!CasperNode::onBlock(b)) return false;
      if (b->height == toSend - 1) {
        int perfectDate = Params::SLOT_DURATION * toSend + delay;
        const int th = toSend;
        auto r = [this, b, th] {
          head = buildBlock(b, th);
          ci.network.sendAll(std::make_shared<SendBlock>(head), ci.network.time + ci.params.blockConstructionTime, *this);
        };
        toSend += ci.params.blockProducersCount;
        if (ci.network.time >= perfectDate) {
          r();
          late++;
        } else {
          ci.network.registerTask(r, perfectDate, *this);
          onTime++;
        }
      }
      return true;
    }
  };

  Params params;
  Network network;  // BlockChainNetwork: `observer` + endPartition resend (below)
  NodeBuilder nb;
  CasperBlock genesis;
  int64_t blockId = 1;
  int64_t attSeq = 0;
  std::vector<std::unique_ptr<CasperBlock>> blocks;
  std::vector<std::shared_ptr<Attestation>> attestations;
  std::vector<std::unique_ptr<CasperNode>> nodes;  // every node ever built (tests build some outside the network)
  CasperNode* observer = nullptr;
  std::vector<Attester*> attesters;
  std::vector<BlockProducer*> bps;

  // :81-88
  explicit CasperIMD(const Params& p) : params(p) {
    nb = nodeBuilderByName(p.nodeBuilderName);
    network.setNetworkLatency(networkLatencyByName(p.networkLatencyName, p.latencyNull));
    observer = add(std::make_unique<CasperNode>(*this, false, OBSERVER));  // network.addObserver(new CasperNode(false, genesis){})
    network.addNode(observer);
  }
  template <class T>
  T* add(std::unique_ptr<T> n) {
    T* raw = n.get();
    nodes.push_back(std::move(n));
    return raw;
  }
  BlockProducer* newBlockProducer() { return add(std::make_unique<BlockProducer>(*this)); }
  Attester* newAttester() { return add(std::make_unique<Attester>(*this)); }
  ByzBlockProducerWF* newByzWF(int delay) { return add(std::make_unique<ByzBlockProducerWF>(*this, delay)); }
  ByzBlockProducer* newByz(int delay) { return add(std::make_unique<ByzBlockProducer>(*this, delay)); }
  ByzBlockProducerSF* newByzSF(int delay) { return add(std::make_unique<ByzBlockProducerSF>(*this, delay)); }
  ByzBlockProducerNS* newByzNS(int delay) { return add(std::make_unique<ByzBlockProducerNS>(*this, delay)); }

  void init() { init(newByzWF(0)); }  // :472-476
  void init(ByzBlockProducer* byzantineNode) {  // :478-508
    const int SD = Params::SLOT_DURATION;
    bps.push_back(byzantineNode);
    network.addNode(byzantineNode);
    network.registerPeriodicTask(byzantineNode->periodicTask(), SD + byzantineNode->delay, SD * params.blockProducersCount, *byzantineNode);
    for (int i = 1; i < params.blockProducersCount; i++) {
      BlockProducer* n = newBlockProducer();
      bps.push_back(n);
      network.addNode(n);
      network.registerPeriodicTask(n->periodicTask(), SD * (i + 1), SD * params.blockProducersCount, *n);
    }
    for (int i = 0; i < params.attestersCount; i++) {
      Attester* n = newAttester();
      attesters.push_back(n);
      network.addNode(n);
      network.registerPeriodicTask(n->periodicTask(), SD * (1 + i % params.cycleLength) + 4000, SD * params.cycleLength, *n);
    }
  }
  // BlockChainNetwork.endPartition (BlockChainNetwork.java:46-54)
  void endPartition() {
    network.endPartition();
    for (Node* n : network.allNodes)
      if (n) network.sendAll(std::make_shared<SendBlock>(static_cast<CasperNode*>(n)->head), *n);
  }
};

inline bool CasperIMD::AttLess::operator()(const Attestation* a, const Attestation* b) const { return a->seq < b->seq; }

inline CasperIMD::Attestation::Attestation(CasperIMD& ci, CasperNode* att, int h)
    : attester(att), height(h), head(att->head), seq(ci.attSeq++) {  // :113-127
  for (CasperBlock* cur = att->head->parent; cur != nullptr && cur->height >= att->head->height - ci.params.cycleLength; cur = cur->parent)
    hs.insert(cur->id);
}

}  // namespace wo
