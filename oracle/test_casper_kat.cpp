// ORACLE — TEST INFRASTRUCTURE ONLY.
// Known-answer tests pinning oracle/casper.hpp against the reference's own unit tests:
//   PT = protocols/src/test/java/net/consensys/wittgenstein/protocols
//   PT/CasperIMDTest.java (all 11 cases), PT/CasperByzantineTest.java (both cases).
// Built and run by tests/test_oracle_casper.py ("make -C oracle kat_casper").
#include <cstdio>
#include <cstdlib>

#include "casper.hpp"

using namespace wo;

static int g_fail = 0;
#define CHECK(cond)                                               \
  do {                                                            \
    if (!(cond)) {                                                \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
      ++g_fail;                                                   \
    }                                                             \
  } while (0)
#define CHECK_EQ(a, b)                                                                           \
  do {                                                                                           \
    long long _a = (long long)(a), _b = (long long)(b);                                          \
    if (_a != _b) {                                                                              \
      std::printf("FAIL %s:%d  %s == %s  (%lld vs %lld)\n", __FILE__, __LINE__, #a, #b, _a, _b); \
      ++g_fail;                                                                                  \
    }                                                                                            \
  } while (0)

using Block = CasperIMD::CasperBlock;
using Att = CasperIMD::Attestation;

// Fixture of PT/CasperIMDTest.java:11-22
struct Fix {
  CasperIMD ci{CasperIMD::makeParams(5, false, 5, 80, 1000, 1, "", "", true)};
  CasperIMD::BlockProducer* bp1 = ci.newBlockProducer();
  CasperIMD::BlockProducer* bp2 = ci.newBlockProducer();
  CasperIMD::Attester* at1 = ci.newAttester();
  CasperIMD::Attester* at2 = ci.newAttester();
  Fix() { ci.network.time = 100000; }
  std::shared_ptr<Att> att(CasperIMD::CasperNode* a, int h) {
    auto v = std::make_shared<Att>(ci, a, h);
    ci.attestations.push_back(v);
    return v;
  }
};

static void testInit() {  // :24-44
  Fix f;
  f.ci.network.time = 0;
  f.ci.init(f.ci.newByzWF(0));
  CHECK_EQ(5 * 80, f.ci.params.attestersCount);
  auto& m = f.ci.network.msgs;
  CHECK_EQ(0, m.sizeAt(1));
  for (int t : {8000, 16000, 24000, 32000, 40000}) CHECK_EQ(1, m.sizeAt(t));
  CHECK_EQ(0, m.sizeAt(48000));
  for (int t : {12000, 20000, 28000, 36000, 44000}) CHECK_EQ(80, m.sizeAt(t));
  CHECK_EQ(0, m.sizeAt(52000));
}

static void testMerge() {  // :46-88
  Fix f;
  Block* b = f.bp1->buildBlock(f.bp1->head, 1);
  f.bp1->onBlock(b);
  CHECK(b == f.bp1->head);
  auto a1 = f.att(f.at1, 1);
  CHECK_EQ(0, a1->hs.size());
  f.at1->onBlock(b);
  CHECK(b == f.at1->head);
  f.at2->onBlock(b);
  a1 = f.att(f.at1, 1);
  CHECK_EQ(1, a1->hs.size());
  CHECK(a1->attests(&f.ci.genesis));
  CHECK(!a1->attests(b));
  a1 = f.att(f.at1, 2);
  CHECK_EQ(1, a1->hs.size());
  CHECK(a1->attests(&f.ci.genesis));
  CHECK(!a1->attests(b));
  f.bp1->onAttestation(a1.get());
  CHECK(f.bp1->attestationsByHead.count(b->id));
  CHECK_EQ(1, f.bp1->attestationsByHead[b->id].size());
  CHECK(f.bp1->attestationsByHead[b->id].count(a1.get()));
  b = f.bp1->buildBlock(f.bp1->head, 2);
  CHECK(!b->attestationsByHeight.count(2));
  b = f.bp1->buildBlock(f.bp1->head, 3);
  CHECK(b->attestationsByHeight.count(2));
  CHECK_EQ(1, b->attestationsByHeight[2].size());
  a1 = f.att(f.at1, 2);
  f.bp1->onAttestation(a1.get());
  b = f.bp1->buildBlock(f.bp1->head, 3);
  CHECK(b->attestationsByHeight.count(2));
  CHECK_EQ(2, b->attestationsByHeight[2].size());
}

static void testCompareNoAttester() {  // :90-104
  Fix f;
  Block* b = f.bp1->buildBlock(f.bp1->head, 1);
  f.bp1->onBlock(b);
  f.bp2->onBlock(b);
  Block* b1 = f.bp1->buildBlock(f.bp1->head, 2);
  Block* b2 = f.bp2->buildBlock(f.bp2->head, 3);
  f.bp2->onBlock(b2);
  CHECK(b2 == f.bp2->head);
  f.bp2->onBlock(b1);
  CHECK(b1 != f.bp2->head);
}

static void testCountAttestationReceived() {  // :106-121
  Fix f;
  Block* b = f.bp1->buildBlock(f.bp1->head, 1);
  f.bp1->onBlock(b);
  f.at1->onBlock(b);
  CHECK_EQ(0, f.bp1->countAttestations(b, &f.ci.genesis));
  auto a1 = f.att(f.at1, 2);
  f.bp1->onAttestation(a1.get());
  CHECK(f.bp1->attestationsByHead.count(b->id));
  CHECK_EQ(1, f.bp1->countAttestations(b, &f.ci.genesis));
}

static void testCountAttestationInBlock() {  // :123-145
  Fix f;
  Block* b = f.bp1->buildBlock(f.bp1->head, 1);
  f.bp1->onBlock(b);
  f.at1->onBlock(b);
  CHECK_EQ(0, f.bp2->countAttestations(b, &f.ci.genesis));
  auto a1 = f.att(f.at1, 2);
  f.bp1->onAttestation(a1.get());
  CHECK(f.bp1->attestationsByHead.count(b->id));
  b = f.bp1->buildBlock(f.bp1->head, 3);
  CHECK(b->attestationsByHeight.count(2));
  CHECK_EQ(1, b->attestationsByHeight[2].size());
  f.bp2->onBlock(b);
  CHECK(b == f.bp2->head);
  CHECK_EQ(1, f.bp2->countAttestations(b, &f.ci.genesis));
}

static void testTooFarAwayAttestation() {  // :147-166
  Fix f;
  Block* b = f.bp1->buildBlock(f.bp1->head, 1);
  f.bp1->onBlock(b);
  f.at1->onBlock(b);
  auto a1 = f.att(f.at1, 2);
  f.bp1->onAttestation(a1.get());
  CHECK(f.bp1->attestationsByHead.count(b->id));
  b = f.bp1->buildBlock(f.bp1->head, a1->height + f.ci.params.cycleLength);
  CHECK(b->attestationsByHeight.count(2));
  b = f.bp1->buildBlock(f.bp1->head, a1->height + f.ci.params.cycleLength + 1);
  CHECK(!b->attestationsByHeight.count(2));
}

static void testOtherBranchAttestation() {  // :168-189
  Fix f;
  Block* b1 = f.bp1->buildBlock(f.bp1->head, 1);
  f.bp1->onBlock(b1);
  f.bp2->onBlock(b1);
  f.at1->onBlock(b1);
  Block* b2 = f.bp1->buildBlock(f.bp1->head, 2);
  f.bp1->onBlock(b2);
  f.at1->onBlock(b2);
  auto a1 = f.att(f.at1, 2);
  CHECK(a1->hs.count(b1->id));
  f.bp2->onAttestation(a1.get());
  Block* b3 = f.bp2->buildBlock(f.bp2->head, 3);
  CHECK(b3->attestationsByHeight[2].empty());
  f.bp2->onBlock(b2);
  b3 = f.bp2->buildBlock(f.bp2->head, 3);
  CHECK(!b3->attestationsByHeight[2].empty());
}

static void testCompareWithAttester() {  // :191-212
  Fix f;
  Block* b1 = f.bp1->buildBlock(f.bp1->head, 1);
  f.bp1->onBlock(b1);
  f.bp2->onBlock(b1);
  f.at1->onBlock(b1);
  Block* b2 = f.bp1->buildBlock(f.bp1->head, 2);
  f.bp1->onBlock(b2);
  f.at1->onBlock(b2);
  auto a1 = f.att(f.at1, 2);
  f.bp1->onAttestation(a1.get());
  Block* b3 = f.bp1->buildBlock(f.bp1->head, 3);
  CHECK_EQ(1, b3->attestationsByHeight[2].size());
  Block* b4 = f.bp2->buildBlock(f.bp2->head, 4);
  f.bp2->onBlock(b4);
  CHECK(b4 == f.bp2->head);
  f.bp2->onBlock(b3);
  CHECK(b3 == f.bp2->head);
}

static void testCompareWithAttesterAttestationOnAParent() {  // :214-232
  Fix f;
  Block* b = f.bp1->buildBlock(f.bp1->head, 1);
  f.bp1->onBlock(b);
  f.bp2->onBlock(b);
  f.at1->onBlock(b);
  auto a1 = f.att(f.at1, 2);
  f.bp1->onAttestation(a1.get());
  Block* b1 = f.bp1->buildBlock(f.bp1->head, 3);
  CHECK_EQ(1, b1->attestationsByHeight[2].size());
  Block* b2 = f.bp2->buildBlock(f.bp2->head, 4);
  f.bp2->onBlock(b2);
  CHECK(b2 == f.bp2->head);
  f.bp2->onBlock(b1);
  CHECK(b2 == f.bp2->head);
}

static void testRevaluation() {  // :234-257
  Fix f;
  Block* b1 = f.bp1->buildBlock(f.bp1->head, 1);
  f.bp1->onBlock(b1);
  f.bp2->onBlock(b1);
  Block* b2 = f.bp1->buildBlock(f.bp1->head, 2);
  Block* b3 = f.bp1->buildBlock(f.bp1->head, 3);
  f.bp2->onBlock(b2);
  f.bp2->onBlock(b3);
  CHECK(b3 == f.bp2->head);
  f.at1->onBlock(b2);
  auto a1 = f.att(f.at1, 2);
  CHECK(a1->hs.count(b1->id));
  f.bp2->onAttestation(a1.get());
  CHECK(f.bp2->attestationsByHead.count(b2->id));
  CHECK_EQ(1, f.bp2->countAttestations(b2, b1));
  f.bp2->reevaluateHead();
  CHECK(b2 == f.bp2->head);
}

static void testCopy() {  // :259-276
  auto prm = CasperIMD::makeParams(5, false, 5, 80, 1000, 1, "", "", true);
  CasperIMD p1(prm), p2(prm);
  p1.init();
  p2.init();
  while (p1.network.time < 20000) {
    p1.network.runMs(10);
    p2.network.runMs(10);
    for (Node* n : p1.network.allNodes) {
      auto* n1 = static_cast<CasperIMD::CasperNode*>(n);
      auto* n2 = static_cast<CasperIMD::CasperNode*>(&p2.network.getNodeById(n1->nodeId));
      CHECK_EQ(n1->doneAt, n2->doneAt);
      CHECK_EQ(n1->isDown(), n2->isDown());
      CHECK_EQ(n1->head->proposalTime, n2->head->proposalTime);
      CHECK_EQ(n1->attestationsByHead.size(), n2->attestationsByHead.size());
      CHECK_EQ(n1->msgReceived, n2->msgReceived);
    }
  }
  CHECK(p1.observer->head->height >= 1);
}

// PT/CasperByzantineTest.java:12-36
static void testByzantineWF() {
  CasperIMD ci(CasperIMD::makeParams(1, false, 2, 2, 1000, 1, "", "", true));
  ci.network.networkLatency = NetworkLatency::ofKind(NetworkLatency::NO_LATENCY);
  auto* byz = ci.newByzWF(0);
  ci.init(byz);
  ci.network.run(9);
  CHECK(&ci.genesis == ci.observer->head);
  ci.network.run(1);
  CHECK(&ci.genesis != ci.observer->head);
  CHECK_EQ(1, ci.observer->head->height);
  CHECK(byz == ci.observer->head->producer);
  ci.network.run(8);
  CHECK_EQ(2, ci.observer->head->height);
  CHECK(byz != ci.observer->head->producer);
  ci.network.run(8);
  CHECK_EQ(3, ci.observer->head->height);
  CHECK(byz == ci.observer->head->producer);
}

// PT/CasperByzantineTest.java:38-65
static void testByzantineWFWithDelay() {
  CasperIMD ci(CasperIMD::makeParams(1, false, 2, 2, 1000, 1, "", "", true));
  ci.network.networkLatency = NetworkLatency::ofKind(NetworkLatency::NO_LATENCY);
  auto* byz = ci.newByzWF(-2000);
  ci.init(byz);
  ci.network.run(5);
  CHECK_EQ(0, byz->head->height);
  ci.network.run(1);
  CHECK_EQ(1, byz->head->height);
  CHECK_EQ(0, ci.observer->head->height);
  ci.network.run(2);
  CHECK_EQ(1, ci.observer->head->height);
  ci.network.run(9);
  CHECK_EQ(1, ci.observer->head->height);
  ci.network.run(1);
  CHECK_EQ(2, byz->head->height);
  CHECK(byz != byz->head->producer);
  ci.network.run(3);
  CHECK_EQ(2, byz->head->height);
  ci.network.run(1);
  CHECK_EQ(3, byz->head->height);
}

int main() {
  testInit();
  testMerge();
  testCompareNoAttester();
  testCountAttestationReceived();
  testCountAttestationInBlock();
  testTooFarAwayAttestation();
  testOtherBranchAttestation();
  testCompareWithAttester();
  testCompareWithAttesterAttestationOnAParent();
  testRevaluation();
  testCopy();
  testByzantineWF();
  testByzantineWFWithDelay();
  if (g_fail) {
    std::printf("CASPER KAT FAILED: %d checks\n", g_fail);
    return 1;
  }
  std::printf("CASPER KAT OK\n");
  return 0;
}
