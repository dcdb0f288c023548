// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build or call it.
//
// CPU restatement (single-threaded, like the reference: core/Network.java:10) of the
// reference engine layer L1+L0:
//   core/Network.java            -> Network, MessageStorage, MsgsSlot, send*/runMs/receiveUntil
//   core/Envelope.java           -> Envelope, SingleDest/MultipleDest/MultipleDestWithDelay
//   core/messages/*.java         -> Message, Task, PeriodicTask, ConditionalTask
//   core/NetworkLatency.java     -> NetworkLatency + samplers
//   core/utils/GeneralizedParetoDistribution.java
//   core/Node.java, core/NodeBuilder.java, core/geoinfo/{Geo,GeoAWS,CityInfo}.java
//   core/RegistryNodeBuilders.java, core/RegistryNetworkLatencies.java
// Each function cites the reference lines it follows.  PARITY STATUS: the reference (Java)
// cannot run in this environment (no JVM); this restatement is pinned against the
// reference's own unit-test assertions (tests/test_oracle_engine.py) — engine, latency and
// RNG layers are pinned; protocol end-states are "parity unpinned" (see DESIGN.md).
//
// Deliberate deviation (performance only, results identical): the conditional-task
// snapshot (Network.java:543-565) uses an index list instead of ArrayList.remove(), which
// is O(n) per removal in the JDK.  This makes the CPU baseline *faster* than the reference.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "jrandom.hpp"

namespace wo {

struct IllegalArgument : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct IllegalState : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ----------------------------------------------------------------------------------------
// GeneralizedParetoDistribution.inverseF  (core/utils/GeneralizedParetoDistribution.java:26-46)
// ----------------------------------------------------------------------------------------
struct GeneralizedParetoDistribution {
  double shape, location, scale;
  GeneralizedParetoDistribution(double sh, double lo, double sc) : shape(sh), location(lo), scale(sc) {
    if (scale <= 0.0) throw IllegalArgument("scale");
  }
  double inverseF(double y) const {
    const double ONE = 0.999999, ZERO = 0.000001;
    if (y < 0.0 || y > 1.0) throw IllegalArgument("y");
    if (y < ZERO) return location;
    if (y > ONE && shape >= 0) return std::numeric_limits<double>::infinity();
    if (y > ONE && shape < 0) return location - scale / shape;
    if (std::fabs(shape) < ZERO) return location - scale * std::log1p(-y);
    return location + scale / shape * (-1 + std::pow(1 - y, -shape));
  }
};

// ----------------------------------------------------------------------------------------
// Node  (core/Node.java)
// ----------------------------------------------------------------------------------------
constexpr int MAX_X = 2000;
constexpr int MAX_Y = 1112;
inline int maxDist() {  // Node.java:17-18
  return static_cast<int>(std::sqrt((MAX_X / 2.0) * (MAX_X / 2.0) + (MAX_Y / 2.0) * (MAX_Y / 2.0)));
}

struct NodeBuilder;

struct Node {
  int nodeId = 0;
  int x = 1, y = 1;
  int extraLatency = 0;
  double speedRatio = 1.0;
  int cityIdx = -1;  // index into the builder's city table; -1 == Node.DEFAULT_CITY ("world")
  std::string cityName = "world";
  bool down = false;
  int64_t doneAt = 0;
  int64_t msgReceived = 0, msgSent = 0, bytesSent = 0, bytesReceived = 0;

  Node() = default;
  Node(JavaRandom& rd, NodeBuilder& nb);  // Node.java:246-271
  virtual ~Node() = default;
  bool isDown() const { return down; }
  void stop() { down = true; }
  void start() { down = false; }
  // Node.java:278-282, toroidal distance
  int dist(const Node& n) const {
    int dx = std::min(std::abs(x - n.x), MAX_X - std::abs(x - n.x));
    int dy = std::min(std::abs(y - n.y), MAX_Y - std::abs(y - n.y));
    return static_cast<int>(std::sqrt(static_cast<double>(dx * dx + dy * dy)));
  }
};

// ----------------------------------------------------------------------------------------
// java.lang.String.hashCode + java.util.HashMap iteration order.
// HashMap<String,..> with <= 12 entries keeps table size 16; iteration = bucket index
// ((h ^ h>>>16) & 15) ascending, insertion order inside a bucket.
// ----------------------------------------------------------------------------------------
inline int32_t javaStringHash(const std::string& s) {
  uint32_t h = 0;
  for (unsigned char c : s) h = 31u * h + c;
  return static_cast<int32_t>(h);
}
inline std::vector<int> javaHashMapOrder(const std::vector<std::string>& keysInPutOrder) {
  size_t cap = 16;
  while (keysInPutOrder.size() > cap * 3 / 4) cap <<= 1;  // default load factor .75 (resize keeps relative order)
  std::vector<int> idx(keysInPutOrder.size());
  for (size_t i = 0; i < idx.size(); ++i) idx[i] = static_cast<int>(i);
  auto bucket = [&](int i) {
    uint32_t h = static_cast<uint32_t>(javaStringHash(keysInPutOrder[static_cast<size_t>(i)]));
    return static_cast<uint32_t>((h ^ (h >> 16)) & (cap - 1));
  };
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return bucket(a) < bucket(b); });
  return idx;
}

// ----------------------------------------------------------------------------------------
// NodeBuilder  (core/NodeBuilder.java), geoinfo/GeoAWS.java, geoinfo/Geo.java
// ----------------------------------------------------------------------------------------
struct Aspects {
  bool uniformSpeed = false;  // Node.SpeedRatioAspect(new Node.UniformSpeed())  RegistryNodeBuilders.java:59-61
  bool extraLatency = false;  // Node.ExtraLatencyAspect(tor)                      RegistryNodeBuilders.java:62-64
  double tor = 0.0;
};

struct CityInfo {
  std::string name;
  int mercX, mercY;
  float cumulativeProbability;
};

struct NodeBuilder {
  enum Kind { BASE, RANDOM_POSITION, CITY } kind = BASE;
  int nodeIds = 0;
  Aspects aspects;
  // CITY: entries in HashMap iteration order (NodeBuilder.java:113-116, Geo.java:10-19)
  std::vector<CityInfo> citiesInfo;
  int citiesListSize = 0;  // this.cities.size()
  // test hooks (anonymous subclasses in the reference tests override getX / getY)
  std::function<int(int)> getXOverride, getYOverride;

  int allocateNodeId() { return nodeIds++; }

  // NodeBuilder.java:128-139  weighted random selection over the HashMap entry set
  int getRandomCityIdx(int32_t rdInt) const {
    int size = citiesListSize;
    // Math.abs(Integer.MIN_VALUE) == Integer.MIN_VALUE
    int32_t a = (rdInt == std::numeric_limits<int32_t>::min()) ? rdInt : std::abs(rdInt);
    int rand = a % size;
    float p = static_cast<float>(rand) / static_cast<float>(size);
    for (size_t i = 0; i < citiesInfo.size(); ++i)
      if (p <= citiesInfo[i].cumulativeProbability) return static_cast<int>(i);
    return -1;  // null in the reference -> NPE
  }
  int getX(int32_t rdInt) const {
    if (getXOverride) return getXOverride(rdInt);
    switch (kind) {
      case RANDOM_POSITION: {  // NodeBuilder.java:82-87: long r = rdInt >> 16; abs; % MAX_X + 1
        int64_t r = static_cast<int64_t>(rdInt >> 16);
        r = r < 0 ? -r : r;
        return static_cast<int>(r % MAX_X + 1);
      }
      case CITY: {
        int c = getRandomCityIdx(rdInt);
        if (c < 0) throw IllegalState("no city");
        return citiesInfo[static_cast<size_t>(c)].mercX;
      }
      default:
        return 1;
    }
  }
  int getY(int32_t rdInt) const {
    if (getYOverride) return getYOverride(rdInt);
    switch (kind) {
      case RANDOM_POSITION: {  // NodeBuilder.java:90-95: long r = rdInt << 16 (int shift, then widened)
        int32_t sh = static_cast<int32_t>(static_cast<uint32_t>(rdInt) << 16);
        int64_t r = static_cast<int64_t>(sh);
        r = r < 0 ? -r : r;
        return static_cast<int>(r % MAX_Y + 1);
      }
      case CITY: {
        int c = getRandomCityIdx(rdInt);
        if (c < 0) throw IllegalState("no city");
        return citiesInfo[static_cast<size_t>(c)].mercY;
      }
      default:
        return 1;
    }
  }
};

// GeoAWS.java:12-22 (put order) + AwsRegionNetworkLatency.regionPerCity (NetworkLatency.java:90-102)
struct AwsCity {
  const char* name;
  int mercX, mercY, region;
};
inline const std::vector<AwsCity>& awsCitiesPutOrder() {
  static const std::vector<AwsCity> v = {
      {"Oregon", 271, 261, 0},     {"Virginia", 513, 316, 1},  {"Mumbai", 1344, 426, 2},
      {"Seoul", 1641, 312, 3},     {"Singapore", 1507, 532, 4}, {"Sydney", 1773, 777, 5},
      {"Tokyo", 1708, 316, 6},     {"Canada central", 422, 256, 7},
      {"Frankfurt", 985, 226, 8},  {"Ireland", 891, 200, 9},   {"London", 937, 205, 10}};
  return v;
}
inline int awsRegionOf(const std::string& city) {
  for (auto& c : awsCitiesPutOrder())
    if (city == c.name) return c.region;
  return -1;
}

inline NodeBuilder makeAwsCityBuilder() {
  NodeBuilder nb;
  nb.kind = NodeBuilder::CITY;
  const auto& put = awsCitiesPutOrder();
  std::vector<std::string> keys;
  for (auto& c : put) keys.push_back(c.name);
  std::vector<int> order = javaHashMapOrder(keys);
  // Geo.cityInfoMap (Geo.java:10-19): float accumulation in HashMap iteration order
  float cumulativeProbability = 0.f;
  int totalPopulation = static_cast<int>(put.size());
  for (int i : order) {
    const AwsCity& c = put[static_cast<size_t>(i)];
    cumulativeProbability = cumulativeProbability + static_cast<float>(1) * 1.f / static_cast<float>(totalPopulation);
    nb.citiesInfo.push_back({c.name, c.mercX, c.mercY, cumulativeProbability});
  }
  nb.citiesListSize = static_cast<int>(put.size());
  return nb;
}

// ----------------------------------------------------------------------------------------
// geoinfo/GeoAllCities.java + NodeBuilder.NodeBuilderWithCity over CSVLatencyReader.cities()
// (RegistryNodeBuilders.java:50-52).  The tables come from the reference's resources through
// scripts/make_city_data.py (text parsing only); everything order- or rounding-sensitive is here.
// ----------------------------------------------------------------------------------------
namespace citydata {
#include "city_data.inc"
}
// GeoAllCities.convertToMercatorX / Y (:54-75); mapWidth = MAX_X, mapHeight = MAX_Y
inline int mercatorX(double longitude) {
  const double mapWidth = MAX_X;
  int posX = static_cast<int>((longitude + 180) * (mapWidth / 360));
  if (posX < mapWidth / 2)
    posX = posX - 45;
  else
    posX = posX - 70;
  return posX;
}
inline int mercatorY(float latitude) {
  const double mapHeight = MAX_Y;
  // (int) Math.round(double): floor(x + 0.5)
  double v = (mapHeight / 2) - (latitude * mapHeight / 180);  // float * double -> double
  int posY = static_cast<int>(std::floor(v + 0.5));
  if (posY < 0.2 * mapHeight) posY = posY - 35;
  return posY;
}
inline int latencyCityIndex(const std::string& name) {
  for (int i = 0; i < citydata::kLatCityCount; ++i)
    if (name == citydata::kLatCities[i]) return i;
  return -1;
}
inline NodeBuilder makeAllCitiesBuilder() {
  NodeBuilder nb;
  nb.kind = NodeBuilder::CITY;
  // readCityInfo (:29-52): HashMap<String,int[]> filled in file order; population += 200000; int total
  std::vector<std::string> keys;
  int totalPopulation = 0;
  for (int i = 0; i < citydata::kGeoCityCount; ++i) {
    keys.push_back(citydata::kGeoCities[i].name);
    totalPopulation += citydata::kGeoCities[i].population + 200000;
  }
  // Geo.cityInfoMap (Geo.java:10-19): cumulative probability in the HashMap's iteration order, float arithmetic
  std::vector<int> order = javaHashMapOrder(keys);
  std::vector<CityInfo> all;
  float cumulativeProbability = 0.f;
  for (int i : order) {
    const auto& c = citydata::kGeoCities[i];
    cumulativeProbability = cumulativeProbability + static_cast<float>(c.population + 200000) * 1.f / static_cast<float>(totalPopulation);
    all.push_back({c.name, mercatorX(static_cast<double>(c.longitude)), mercatorY(c.latitude), cumulativeProbability});
  }
  // citiesInfo map -> copy (citiesPosition()) -> filter by the latency reader's cities (case-insensitive) -> Collectors.toMap:
  // every step is a HashMap over (a subset of) the same keys filled in the previous one's iteration order
  std::vector<CityInfo> kept;
  std::vector<std::string> keptKeys;
  for (const CityInfo& ci : all) {
    bool in = false;
    for (int k = 0; k < citydata::kLatCityCount && !in; ++k) {
      std::string a = ci.name, b = citydata::kLatCities[k];
      for (auto& ch : a) ch = static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
      for (auto& ch : b) ch = static_cast<char>(std::toupper(static_cast<unsigned char>(ch)));
      in = a == b;
    }
    if (in) {
      kept.push_back(ci);
      keptKeys.push_back(ci.name);
    }
  }
  for (int i : javaHashMapOrder(keptKeys)) nb.citiesInfo.push_back(kept[static_cast<size_t>(i)]);
  nb.citiesListSize = citydata::kLatCityCount;  // cities.size(): the reader's key set, whether or not cities.csv knows the city
  return nb;
}

// RegistryNodeBuilders.name / getByName (RegistryNodeBuilders.java:21-25, 71-81)
inline NodeBuilder nodeBuilderByName(const std::string& nameIn) {
  std::string name = nameIn;
  bool blank = true;
  for (char c : name)
    if (c != ' ' && c != '\t') blank = false;
  if (blank) name = "RANDOM_SPEED=CONSTANT_TOR=0.00";
  // <SITE>_SPEED=<CONSTANT|GAUSSIAN>_TOR=<d.dd>
  size_t p1 = name.find("_SPEED=");
  size_t p2 = name.find("_TOR=");
  if (p1 == std::string::npos || p2 == std::string::npos || p2 < p1) throw IllegalArgument(name + " not in the registry");
  std::string site = name.substr(0, p1);
  std::string speed = name.substr(p1 + 7, p2 - (p1 + 7));
  std::string torS = name.substr(p2 + 5);
  static const char* tors[] = {"0.00", "0.01", "0.10", "0.20", "0.33", "0.50", "0.60", "0.80", "1.00"};
  static const double torv[] = {0.0, 0.01, 0.10, 0.20, .33, .5, .6, .8, 1.0};
  double tor = -1;
  for (int i = 0; i < 9; ++i)
    if (torS == tors[i]) tor = torv[i];
  if (tor < 0) throw IllegalArgument(name + " not in the registry");
  NodeBuilder nb;
  if (site == "AWS")
    nb = makeAwsCityBuilder();
  else if (site == "RANDOM")
    nb.kind = NodeBuilder::RANDOM_POSITION;
  else if (site == "CITIES")
    nb = makeAllCitiesBuilder();
  else
    throw IllegalArgument(name + " not in the registry");
  if (speed == "GAUSSIAN")
    nb.aspects.uniformSpeed = true;  // sic: the GAUSSIAN label installs UniformSpeed
  else if (speed != "CONSTANT")
    throw IllegalArgument(name + " not in the registry");
  if (tor > 0.001) {
    nb.aspects.extraLatency = true;
    nb.aspects.tor = tor;
  }
  return nb;
}

// Node.java:246-271
inline Node::Node(JavaRandom& rd, NodeBuilder& nb) {
  nodeId = nb.allocateNodeId();
  if (nodeId < 0) throw IllegalArgument("bad nodeId");
  int32_t rdNode = rd.nextInt();
  if (nb.kind == NodeBuilder::CITY && !nb.getXOverride) {
    cityIdx = nb.getRandomCityIdx(rdNode);
    if (cityIdx < 0) throw IllegalState("no city");
    cityName = nb.citiesInfo[static_cast<size_t>(cityIdx)].name;
  }
  x = nb.getX(rdNode);
  y = nb.getY(rdNode);
  if (x <= 0 || x > MAX_X) throw IllegalArgument("bad x");
  if (y <= 0 || y > MAX_Y) throw IllegalArgument("bad y");
  // Aspects, in this order: speed ratio, then extra latency (Node.java:264-265)
  if (nb.aspects.uniformSpeed) {  // Node.java:233-238
    speedRatio = rd.nextBoolean() ? (rd.nextInt(67) + 33) / 100.0 : (rd.nextInt(200) + 100) / 100.0;
  }
  if (nb.aspects.extraLatency) {  // Node.java:158-160
    extraLatency = rd.nextDouble() < nb.aspects.tor ? 500 : 0;
  }
  if (speedRatio <= 0) throw IllegalArgument("speedRatio");
}

// ----------------------------------------------------------------------------------------
// NetworkLatency  (core/NetworkLatency.java)
// ----------------------------------------------------------------------------------------
struct NetworkLatency {
  enum Kind {
    BY_DISTANCE_W_JITTER = 0,
    AWS_REGION = 1,
    FIXED = 2,
    UNIFORM = 3,
    NO_LATENCY = 4,
    MEASURED = 5,
    ETHSCAN = 6,
    IC3 = 7,
    BY_CITY = 8,
    BY_CITY_W_JITTER = 9
  } kind = IC3;
  int param = 0;             // FIXED / UNIFORM
  int longDistrib[100] = {};  // MEASURED / ETHSCAN
  GeneralizedParetoDistribution gpd{1.4, -0.3, 0.35};

  static void checkDelta(int delta) {
    if (delta < 0 || delta > 99) throw IllegalArgument("delta");
  }
  double getJitter(int delta) const { return gpd.inverseF(delta / 100.0); }  // :59-61
  static double getFixedLatency(int dist) {                                  // :53-65
    const double earthPerimeter = 24860;
    const double pointValue = (earthPerimeter / 2) / maxDist();
    return pointValue * dist * 0.022 + 4.862;
  }
  // MeasuredNetworkLatency.setLatency :284-303
  void setMeasured(const std::vector<int>& proportions, const std::vector<int>& values) {
    int li = 0, cur = 0, sum = 0;
    for (size_t i = 0; i < proportions.size(); i++) {
      if (proportions[i] == 0) {
        cur = values[i];
        continue;
      }
      sum += proportions[i];
      int step = (values[i] - cur) / proportions[i];
      for (int ii = 0; ii < proportions[i]; ii++) {
        cur += step;
        if (li >= 100) throw IllegalArgument("li");
        longDistrib[li++] = cur;
      }
    }
    if (sum != 100) throw IllegalArgument("sum");
    if (li != 100) throw IllegalArgument("li");
  }

  // NetworkLatencyByCity.getLatency(cityFrom, cityTo) :188-198 (the converter resolved the "other direction" fallback)
  static float cityLatency(const Node& from, const Node& to) {
    if (from.cityIdx < 0 || to.cityIdx < 0) throw IllegalState("Can't use NetworkLatencyByCity model with default city location");
    int a = latencyCityIndex(from.cityName), b = latencyCityIndex(to.cityName);
    if (a < 0) throw IllegalArgument("Can't find latencies for " + from.cityName);
    if (b < 0) throw IllegalArgument("Can't find latencies for " + to.cityName);
    return citydata::kCityLatency[a][b];
  }
  int getExtendedLatency(const Node& from, const Node& to, int delta) const {
    switch (kind) {
      case BY_CITY: {  // :168-185
        if (from.nodeId == to.nodeId) return 1;
        float h = 0.5f * cityLatency(from, to);
        int r = static_cast<int>(std::floor(static_cast<double>(h) + 0.5));  // Math.round(float): floor(x + 1/2), exact
        return std::max(1, r);
      }
      case BY_CITY_W_JITTER: {  // :204-232
        if (from.nodeId == to.nodeId) return 1;
        float lat = cityLatency(from, to);
        double raw = getJitter(delta);
        if (from.cityName == to.cityName)
          raw += 10;
        else
          raw += lat;  // float widened to double
        return std::max(1, static_cast<int>(std::floor(0.5 * raw + 0.5)));  // (int) Math.round(double)
      }
      case BY_DISTANCE_W_JITTER: {  // :67-72
        checkDelta(delta);
        double raw = getFixedLatency(from.dist(to)) + getJitter(delta);
        return static_cast<int>(raw / 2);
      }
      case AWS_REGION: {  // :130-151
        int reg1 = awsRegionOf(from.cityName), reg2 = awsRegionOf(to.cityName);
        if (reg1 < 0 || reg2 < 0) throw IllegalArgument("not in our aws cities list");
        if (reg1 == reg2) return 1;
        static const int latencies[10][11] = {
            {0, 81, 216, 126, 165, 138, 97, 64, 164, 131, 141}, {0, 0, 182, 181, 232, 195, 167, 13, 88, 80, 75},
            {0, 0, 0, 152, 62, 223, 123, 194, 111, 122, 113},   {0, 0, 0, 0, 97, 133, 35, 184, 259, 254, 264},
            {0, 0, 0, 0, 0, 169, 69, 218, 162, 174, 171},       {0, 0, 0, 0, 0, 0, 105, 210, 282, 269, 271},
            {0, 0, 0, 0, 0, 0, 0, 156, 235, 222, 234},          {0, 0, 0, 0, 0, 0, 0, 0, 101, 78, 87},
            {0, 0, 0, 0, 0, 0, 0, 0, 0, 24, 13},                {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 12}};
        int minReg = std::min(reg1, reg2), maxReg = std::max(reg1, reg2);
        return std::max(1, latencies[minReg][maxReg] / 2 + static_cast<int>(getJitter(delta)));
      }
      case FIXED:  // :242-244
        return param;
      case UNIFORM:  // :262-264
        return static_cast<int>((delta / 99.0) * param);
      case NO_LATENCY:  // :272-274
        return 1;
      case MEASURED:  // :310-313
        checkDelta(delta);
        return longDistrib[delta];
      case ETHSCAN: {  // :375-377 wraps MeasuredNetworkLatency.getLatency -> extra latency counted twice
        if (&from == &to) return 1;
        checkDelta(delta);
        int base = from.extraLatency + to.extraLatency + longDistrib[delta];
        return std::max(1, base);
      }
      case IC3: {  // :404-416
        double dist = from.dist(to);
        double surface = dist * dist * M_PI;
        double totalSurface = MAX_X * MAX_Y;
        int position = static_cast<int>((surface * 100) / totalSurface);
        if (position <= 10) return 92 / 2;
        if (position <= 33) return 125 / 2;
        if (position <= 50) return 152 / 2;
        if (position <= 67) return 200 / 2;
        if (position <= 90) return 276 / 2;
        return 350 / 2;
      }
    }
    throw IllegalState("latency kind");
  }
  // NetworkLatency.getLatency :27-34
  int getLatency(const Node& from, const Node& to, int delta) const {
    if (&from == &to) return 1;
    int base = from.extraLatency + to.extraLatency;
    base += getExtendedLatency(from, to, delta);
    return std::max(1, base);
  }

  static NetworkLatency fixed(int f) {
    NetworkLatency l;
    l.kind = FIXED;
    l.param = std::max(1, f);
    return l;
  }
  static NetworkLatency uniform(int f) {
    NetworkLatency l;
    l.kind = UNIFORM;
    l.param = std::max(1, f);
    return l;
  }
  static NetworkLatency measured(const std::vector<int>& p, const std::vector<int>& v) {
    NetworkLatency l;
    l.kind = MEASURED;
    l.setMeasured(p, v);
    return l;
  }
  static NetworkLatency ethScan() {  // :366-372
    NetworkLatency l;
    l.setMeasured({16, 18, 17, 12, 8, 5, 4, 3, 3, 1, 1, 2, 1, 1, 8},
                  {250, 500, 1000, 1250, 1500, 1750, 2000, 2250, 2500, 2750, 4500, 6000, 8500, 9750, 10000});
    l.kind = ETHSCAN;
    return l;
  }
  static NetworkLatency ofKind(Kind k) {
    NetworkLatency l;
    l.kind = k;
    return l;
  }
};

// RegistryNetworkLatencies.getByName (RegistryNetworkLatencies.java:28-58)
inline NetworkLatency networkLatencyByName(const std::string& nameIn, bool isNull = false) {
  std::string name = isNull ? "NetworkLatencyByDistanceWJitter" : nameIn;
  for (int f : {0, 100, 200, 500, 1000, 2000, 4000, 8000}) {
    if (name == "NetworkFixedLatency(" + std::to_string(f) + ")") return NetworkLatency::fixed(f);
    if (name == "NetworkUniformLatency(" + std::to_string(f) + ")") return NetworkLatency::uniform(f);
  }
  if (name == "NetworkLatencyByDistanceWJitter") return NetworkLatency::ofKind(NetworkLatency::BY_DISTANCE_W_JITTER);
  if (name == "AwsRegionNetworkLatency") return NetworkLatency::ofKind(NetworkLatency::AWS_REGION);
  if (name == "NetworkNoLatency") return NetworkLatency::ofKind(NetworkLatency::NO_LATENCY);
  if (name == "EthScanNetworkLatency") return NetworkLatency::ethScan();
  if (name == "NetworkLatencyByCity") return NetworkLatency::ofKind(NetworkLatency::BY_CITY);
  if (name == "NetworkLatencyByCityWJitter") return NetworkLatency::ofKind(NetworkLatency::BY_CITY_W_JITTER);
  if (name == "IC3NetworkLatency") return NetworkLatency::ofKind(NetworkLatency::IC3);
  throw IllegalArgument("latency '" + name + "' not in the registry");
}

// ----------------------------------------------------------------------------------------
// Messages  (core/messages/*.java)
// ----------------------------------------------------------------------------------------
struct Network;

struct Message {
  virtual ~Message() = default;
  virtual void action(Network& network, Node& from, Node& to) = 0;  // Message.java:21
  virtual int size() const { return 1; }                            // Message.java:27-29
  virtual bool isTask() const { return false; }
};
using MessagePtr = std::shared_ptr<Message>;

struct Task : Message {  // Task.java
  std::function<void()> r;
  explicit Task(std::function<void()> f) : r(std::move(f)) {}
  int size() const override { return 0; }
  bool isTask() const override { return true; }
  void action(Network&, Node&, Node&) override { r(); }
};

struct ConditionalTask {  // ConditionalTask.java
  std::function<bool()> startIf, repeatIf;
  std::function<void()> r;
  int duration;
  int minStartTime;
  Node* from;
  bool removed = false;  // removed from Network.conditionalTasks
};

// ----------------------------------------------------------------------------------------
// Envelope  (core/Envelope.java)
// ----------------------------------------------------------------------------------------
struct Envelope {
  int sendTime;
  Envelope* nextSameTime = nullptr;
  MessagePtr message;
  int fromNodeId;
  explicit Envelope(int st, MessagePtr m, int from) : sendTime(st), message(std::move(m)), fromNodeId(from) {}
  virtual ~Envelope() = default;
  virtual int getNextDestId() const = 0;
  virtual int nextArrivalTime(const Network& network) const = 0;
  virtual void markRead() = 0;
  virtual bool hasNextReader() const = 0;
  // Envelope.infos(network) (:39): (destination, arrival) of every destination not yet served
  virtual void infos(const Network& network, std::vector<std::pair<int, int>>& out) const = 0;
};

struct EnvelopeInfo {  // core/EnvelopeInfo.java:8-14 (msg: the isTask flag is all the checker needs)
  int from, to, sentAt, arrivingAt;
  bool isTask;
};

struct MessageArrival {  // Network.java:392-410
  Node* dest;
  int arrival;
};

struct SingleDestEnvelope : Envelope {  // Envelope.java:230-301
  int toNodeId, arrivalTime;
  SingleDestEnvelope(MessagePtr m, const Node& from, const Node& to, int sendTime, int arrival)
      : Envelope(sendTime, std::move(m), from.nodeId), toNodeId(to.nodeId), arrivalTime(arrival) {}
  int getNextDestId() const override { return toNodeId; }
  int nextArrivalTime(const Network&) const override { return arrivalTime; }
  void markRead() override {}
  bool hasNextReader() const override { return false; }
  void infos(const Network&, std::vector<std::pair<int, int>>& out) const override { out.emplace_back(toNodeId, arrivalTime); }  // :297-300
};

struct MultipleDestEnvelope : Envelope {  // Envelope.java:57-155
  int randomSeed;
  std::vector<int> destIds;
  int curPos = 0;
  MultipleDestEnvelope(MessagePtr m, const Node& from, const std::vector<MessageArrival>& dests, int sendTime, int seed)
      : Envelope(sendTime, std::move(m), from.nodeId), randomSeed(seed) {
    for (auto& d : dests) destIds.push_back(d.dest->nodeId);
  }
  int getNextDestId() const override { return destIds[static_cast<size_t>(curPos)]; }
  int nextArrivalTime(const Network& network) const override;  // recomputed from the seed (:107-118)
  int arrivalTime(const Network& network, int destId) const;    // :120-124
  void markRead() override { curPos++; }
  bool hasNextReader() const override { return curPos < static_cast<int>(destIds.size()); }
  void infos(const Network& network, std::vector<std::pair<int, int>>& out) const override {  // :145-154
    for (size_t i = static_cast<size_t>(curPos); i < destIds.size(); ++i) out.emplace_back(destIds[i], arrivalTime(network, destIds[i]));
  }
};

struct MultipleDestWithDelayEnvelope : Envelope {  // Envelope.java:157-228
  std::vector<int> destIds, arrivalTime;
  int curPos = 0;
  MultipleDestWithDelayEnvelope(MessagePtr m, const Node& from, const std::vector<MessageArrival>& dests, int sendTime)
      : Envelope(sendTime, std::move(m), from.nodeId) {
    for (auto& d : dests) {
      destIds.push_back(d.dest->nodeId);
      arrivalTime.push_back(d.arrival);
    }
  }
  int getNextDestId() const override { return destIds[static_cast<size_t>(curPos)]; }
  int nextArrivalTime(const Network&) const override { return arrivalTime[static_cast<size_t>(curPos)]; }
  void markRead() override { curPos++; }
  bool hasNextReader() const override { return curPos < static_cast<int>(destIds.size()); }
  void infos(const Network&, std::vector<std::pair<int, int>>& out) const override {  // :219-227
    for (size_t i = static_cast<size_t>(curPos); i < destIds.size(); ++i) out.emplace_back(destIds[i], arrivalTime[i]);
  }
};

// ----------------------------------------------------------------------------------------
// Network  (core/Network.java)
// ----------------------------------------------------------------------------------------
struct Network {
  static constexpr int duration = 60 * 1000;  // :14

  struct MsgsSlot {  // :116-199
    int startTime, endTime;
    std::vector<Envelope*> msgsByMs;
    explicit MsgsSlot(int st) : startTime(st - (st % duration)), endTime(st + duration), msgsByMs(duration, nullptr) {}
    int getPos(int aTime) const {
      if (aTime < startTime || aTime >= startTime + duration) throw IllegalArgument("aTime");
      return aTime % duration;
    }
  };

  struct MessageStorage {  // :201-299
    Network& net;
    std::vector<std::unique_ptr<MsgsSlot>> msgsBySlot;
    int64_t live = 0;  // == size(), maintained incrementally (size() itself walks the lists like the reference)
    explicit MessageStorage(Network& n) : net(n) {}
    ~MessageStorage() { clearAll(); }

    int size() const {  // :204-210, 164-176
      int sz = 0;
      for (auto& ms : msgsBySlot)
        for (Envelope* m : ms->msgsByMs)
          for (; m != nullptr; m = m->nextSameTime) sz++;
      return sz;
    }
    int sizeAt(int t) {  // :212-220
      int sz = 0;
      for (Envelope* cur = peek(t); cur != nullptr; cur = cur->nextSameTime) sz++;
      return sz;
    }
    void cleanup() {  // :222-229
      while (!msgsBySlot.empty() && net.time >= msgsBySlot[0]->endTime) {
        for (Envelope* m : msgsBySlot[0]->msgsByMs)
          while (m) {
            Envelope* n = m->nextSameTime;
            delete m;
            --live;
            m = n;
          }
        msgsBySlot.erase(msgsBySlot.begin());
      }
      if (msgsBySlot.empty()) msgsBySlot.push_back(std::make_unique<MsgsSlot>(net.time));
    }
    void ensureSize(int aTime) {  // :231-235
      while (msgsBySlot.back()->endTime <= aTime) {
        int e = msgsBySlot.back()->endTime;
        msgsBySlot.push_back(std::make_unique<MsgsSlot>(e));
      }
    }
    MsgsSlot& findSlot(int aTime) {  // :237-245
      cleanup();
      ensureSize(aTime);
      int pos = (aTime - msgsBySlot[0]->startTime) / duration;
      if (pos >= static_cast<int>(msgsBySlot.size()) || pos < 0) throw IllegalState("pos");
      return *msgsBySlot[static_cast<size_t>(pos)];
    }
    void addMsg(Envelope* m) {  // :247-255, 134-148
      int na = m->nextArrivalTime(net);
      if (na < net.time) {
        delete m;
        throw IllegalState("Arriving in the past");
      }
      MsgsSlot& slot = findSlot(na);
      int pos = slot.getPos(na);
      m->nextSameTime = slot.msgsByMs[static_cast<size_t>(pos)];  // push-front: LIFO inside one ms
      slot.msgsByMs[static_cast<size_t>(pos)] = m;
      ++live;
    }
    Envelope* peek(int t) {  // :257-259
      MsgsSlot& s = findSlot(t);
      return s.msgsByMs[static_cast<size_t>(s.getPos(t))];
    }
    Envelope* poll(int t) {  // :261-263, 155-162
      MsgsSlot& s = findSlot(t);
      size_t pos = static_cast<size_t>(s.getPos(t));
      Envelope* m = s.msgsByMs[pos];
      if (m != nullptr) {
        s.msgsByMs[pos] = m->nextSameTime;
        --live;
      }
      return m;
    }
    void clearAll() {
      for (auto& ms : msgsBySlot)
        for (Envelope* m : ms->msgsByMs)
          while (m) {
            Envelope* n = m->nextSameTime;
            delete m;
            m = n;
          }
      msgsBySlot.clear();
      live = 0;
    }
    void clear() {  // :265-268
      clearAll();
      cleanup();
    }
    std::vector<EnvelopeInfo> peekMessages() const;  // :279-286 (defined after Message)
    Envelope* peekFirst() {  // :271-277
      for (auto& ms : msgsBySlot)
        for (Envelope* m : ms->msgsByMs)
          if (m) return m;
      return nullptr;
    }
  };

  MessageStorage msgs{*this};
  std::list<ConditionalTask> conditionalTasks;  // :23 (LinkedList, registration order)
  std::vector<ConditionalTask*> ctIndex;        // same order; entries flagged `removed` are skipped
  std::vector<Node*> allNodes;                  // :29
  JavaRandom rd{0};                             // :32
  std::vector<int> partitionsInX;               // :34
  int msgDiscardTime = std::numeric_limits<int>::max();  // :40
  NetworkLatency networkLatency;                // :43 default IC3
  int time = 0;                                 // :49

  // statistics (not in the reference): executed events, for msgs/sec and roofline accounting
  int64_t statDeliveries = 0, statTasks = 0, statCondRuns = 0, statDraws = 0;

  virtual ~Network() = default;

  // chooseBadNodes :52-64
  static std::vector<bool> chooseBadNodes(JavaRandom& rd, int nodeCount, int nodesDown) {
    std::vector<bool> bad(static_cast<size_t>(nodeCount), false);
    for (int setDown = 0; setDown < nodesDown;) {
      int down = rd.nextInt(nodeCount);
      if (down != 1 && !bad[static_cast<size_t>(down)]) {
        bad[static_cast<size_t>(down)] = true;
        setDown++;
      }
    }
    return bad;
  }

  Node& getNodeById(int id) { return *allNodes.at(static_cast<size_t>(id)); }
  void addNode(Node* n) {  // :651-659
    while (static_cast<int>(allNodes.size()) <= n->nodeId) allNodes.push_back(nullptr);
    if (allNodes[static_cast<size_t>(n->nodeId)] != nullptr) throw IllegalState("There is already a node with this id");
    allNodes[static_cast<size_t>(n->nodeId)] = n;
  }

  bool runMs(int ms) {  // :318-338
    if (ms <= 0) throw IllegalArgument("Should be greater than 0");
    if (time == 0)
      for (Node* n : allNodes)
        if (n && !n->isDown()) n->start();
    int endAt = static_cast<int>(static_cast<uint32_t>(time) + static_cast<uint32_t>(ms));
    if (endAt <= 0) throw IllegalState("Maximum time reached!");
    bool didSomething = receiveUntil(endAt);
    time = endAt;
    return didSomething;
  }
  void run(int seconds) { runMs(seconds * 1000); }

  void checkIn(const Node& n, const char* what) const {
    if (n.nodeId >= static_cast<int>(allNodes.size()) || allNodes[static_cast<size_t>(n.nodeId)] != &n) throw IllegalArgument(what);
  }

  // sendAll :341-347
  void sendAll(const MessagePtr& m, int sendTime, Node& from) { send(m, sendTime, from, allNodes, 0); }
  void sendAll(const MessagePtr& m, Node& from) { send(m, time + 1, from, allNodes, 0); }
  // send(m, from, dests) :353-362
  void send(const MessagePtr& m, Node& from, const std::vector<Node*>& dests) {
    if (dests.empty()) return;
    if (dests.size() == 1)
      send(m, time + 1, from, *dests[0]);
    else
      send(m, time + 1, from, dests, 0);
  }
  void send(const MessagePtr& m, Node& from, Node& to) { send(m, time + 1, from, to); }  // :364-366
  // single destination :369-382
  void send(const MessagePtr& mc, int sendTime, Node& from, Node& to) {
    checkIn(from, "The from node is not in the network");
    checkIn(to, "The to node is not in the network");
    int seed = rd.nextInt();
    ++statDraws;
    MessageArrival ms;
    if (createMessageArrival(*mc, from, to, sendTime, seed, ms)) {
      msgs.addMsg(new SingleDestEnvelope(mc, from, to, sendTime, ms.arrival));
    }
  }
  // sendArriveAt :384-390
  void sendArriveAt(const MessagePtr& mc, int arriveAt, Node& from, Node& to) {
    if (arriveAt <= time) throw IllegalArgument("wrong arrival time");
    msgs.addMsg(new SingleDestEnvelope(mc, from, to, time, arriveAt));
  }
  // multi destination :420-447
  void send(const MessagePtr& m, int sendTime, Node& from, const std::vector<Node*>& dests, int delaysBetweenMessage) {
    checkIn(from, "The from node is not in the network");
    int randomSeed = rd.nextInt();
    ++statDraws;
    std::vector<MessageArrival> da = createMessageArrivals(*m, sendTime, from, dests, randomSeed, delaysBetweenMessage);
    if (!da.empty()) {
      Envelope* msg;
      if (da.size() == 1)
        msg = new SingleDestEnvelope(m, from, *da[0].dest, sendTime, da[0].arrival);
      else if (delaysBetweenMessage == 0)
        msg = new MultipleDestEnvelope(m, from, da, sendTime, randomSeed);
      else
        msg = new MultipleDestWithDelayEnvelope(m, from, da, sendTime);
      msgs.addMsg(msg);
    }
  }
  // createMessageArrivals :449-467   (Collections.sort is a stable merge sort)
  std::vector<MessageArrival> createMessageArrivals(Message& m, int sendTime, Node& from, const std::vector<Node*>& dests,
                                                    int randomSeed, int delaysBetweenMessage) {
    std::vector<MessageArrival> da;
    da.reserve(dests.size());
    for (Node* n : dests) {
      MessageArrival ma;
      bool ok = createMessageArrival(m, from, *n, sendTime, randomSeed, ma);
      sendTime += delaysBetweenMessage + (delaysBetweenMessage > 0 ? 1 : 0);
      if (ok) da.push_back(ma);
    }
    std::stable_sort(da.begin(), da.end(), [](const MessageArrival& a, const MessageArrival& b) { return a.arrival < b.arrival; });
    return da;
  }
  // createMessageArrival :469-487
  bool createMessageArrival(Message& m, Node& from, Node& to, int sendTime, int randomSeed, MessageArrival& out) {
    if (sendTime <= time) throw IllegalState("sendTime <= time");
    from.msgSent++;
    from.bytesSent += m.size();
    if (partitionId(from) == partitionId(to) && !from.isDown() && !to.isDown()) {
      int nt = networkLatency.getLatency(from, to, getPseudoRandom(to.nodeId, randomSeed));
      if (nt < msgDiscardTime) {
        out.dest = &to;
        out.arrival = sendTime + nt;
        return true;
      }
    }
    return false;
  }
  // getPseudoRandom / hash :493-503
  static int getPseudoRandom(int nodeId, int randomSeed) {
    int32_t x = hash(nodeId) ^ randomSeed;
    int32_t r = x % 100;  // truncating, like Java
    return r < 0 ? -r : r;
  }
  static int32_t hash(int32_t a0) {
    uint32_t a = static_cast<uint32_t>(a0);
    a ^= (a << 13);
    a ^= (a >> 17);
    a ^= (a << 5);
    return static_cast<int32_t>(a);
  }

  // registerTask / registerPeriodicTask / registerConditionalTask :505-531
  void registerTask(std::function<void()> task, int startAt, Node& from) {
    auto sw = std::make_shared<Task>(std::move(task));
    msgs.addMsg(new SingleDestEnvelope(sw, from, from, time, startAt));
  }
  struct PeriodicTask : Task {  // PeriodicTask.java
    int period;
    Node* sender;
    std::function<bool()> continuationCondition;
    std::weak_ptr<PeriodicTask> self;
    PeriodicTask(std::function<void()> f, Node* from, int p, std::function<bool()> c)
        : Task(std::move(f)), period(p), sender(from), continuationCondition(std::move(c)) {}
    void action(Network& network, Node&, Node&) override {  // :40-47
      r();
      if (continuationCondition()) network.sendArriveAt(self.lock(), network.time + period, *sender, *sender);
    }
  };
  void registerPeriodicTask(std::function<void()> task, int startAt, int period, Node& from,
                            std::function<bool()> c = [] { return true; }) {
    auto sw = std::make_shared<PeriodicTask>(std::move(task), &from, period, std::move(c));
    sw->self = sw;
    msgs.addMsg(new SingleDestEnvelope(sw, from, from, time, startAt));
  }
  void registerConditionalTask(std::function<void()> task, int startAt, int dur, Node& from, std::function<bool()> startIf,
                               std::function<bool()> repeatIf) {
    conditionalTasks.push_back(ConditionalTask{std::move(startIf), std::move(repeatIf), std::move(task), dur, startAt, &from});
    ctIndex.push_back(&conditionalTasks.back());
  }

  // nextMessage :533-570.  `cts` = snapshot of conditionalTasks taken lazily once per call.
  Envelope* nextMessage(int until) {
    bool haveSnapshot = false;
    std::vector<ConditionalTask*>& cts = snapshot_;
    while (time <= until) {
      Envelope* m = msgs.poll(time);
      if (m != nullptr) return m;
      time++;
      if (!haveSnapshot) {
        cts.clear();
        size_t w = 0;
        for (ConditionalTask* ct : ctIndex)
          if (!ct->removed) {
            cts.push_back(ct);
            ctIndex[w++] = ct;
          }
        ctIndex.resize(w);
        haveSnapshot = true;
      }
      size_t keep = 0;
      for (size_t i = 0; i < cts.size(); ++i) {
        ConditionalTask* ct = cts[i];
        if (ct->minStartTime > until || ct->from->isDown()) continue;  // it.remove()
        if (ct->minStartTime <= time) {                                // it.remove(), then maybe run
          if (ct->startIf()) {
            ct->r();
            ++statCondRuns;
            ct->minStartTime = time + ct->duration;
            if (!ct->repeatIf()) ct->removed = true;  // conditionalTasks.remove(ct)
          }
          continue;
        }
        cts[keep++] = ct;
      }
      cts.resize(keep);
    }
    return nullptr;
  }

  // receiveUntil :587-637
  bool receiveUntil(int until) {
    int previousTime = time;
    Envelope* next = nextMessage(until);
    if (next == nullptr) return false;
    while (next != nullptr) {
      Envelope* m = next;
      int na = m->nextArrivalTime(*this);
      if (na != previousTime && time > na) throw IllegalState("time > arrival");
      Node& from = *allNodes[static_cast<size_t>(m->fromNodeId)];
      Node& to = *allNodes[static_cast<size_t>(m->getNextDestId())];
      if (!to.isDown() && partitionId(from) == partitionId(to)) {
        if (!m->message->isTask()) {
          if (m->message->size() == 0) throw IllegalState("Message size should be greater than zero");
          to.msgReceived++;
          to.bytesReceived += m->message->size();
          ++statDeliveries;
        } else {
          ++statTasks;
        }
        MessagePtr keepAlive = m->message;
        keepAlive->action(*this, from, to);
      }
      m->markRead();
      if (m->hasNextReader())
        msgs.addMsg(m);
      else
        delete m;
      previousTime = time;
      next = nextMessage(until);
    }
    return true;
  }

  int partitionId(const Node& to) const {  // :639-649
    int pId = 0;
    for (int x : partitionsInX) {
      if (x > to.x) return pId;
      pId++;
    }
    return pId;
  }
  void setNetworkLatency(const NetworkLatency& nl) {  // :669-677
    if (msgs.size() != 0) throw IllegalState("You can't change the latency while the system as on going messages");
    networkLatency = nl;
  }
  void partition(float part) {  // :693-703
    if (part <= 0 || part >= 1) throw IllegalArgument("part needs to be a percentage between 0 & 100 excluded");
    int xPoint = static_cast<int>(static_cast<float>(MAX_X) * part);
    if (std::find(partitionsInX.begin(), partitionsInX.end(), xPoint) != partitionsInX.end())
      throw IllegalArgument("this partition exists already");
    partitionsInX.push_back(xPoint);
    std::sort(partitionsInX.begin(), partitionsInX.end());
  }
  void endPartition() { partitionsInX.clear(); }
  void setMsgDiscardTime(int l) { msgDiscardTime = l; }

 private:
  std::vector<ConditionalTask*> snapshot_;
};

inline int MultipleDestEnvelope::arrivalTime(const Network& network, int destId) const {  // Envelope.java:120-124
  int rd = Network::getPseudoRandom(destId, randomSeed);
  const Node& f = *network.allNodes[static_cast<size_t>(fromNodeId)];
  const Node& t = *network.allNodes[static_cast<size_t>(destId)];
  return sendTime + network.networkLatency.getLatency(f, t, rd);
}
// Network.MessageStorage.peekMessages :279-286 — all slots, MsgsSlot.infos (:178-187: per ms, the list from its head), sorted
// with EnvelopeInfo.compareTo (EnvelopeInfo.java:33-48: arrivingAt, then from; its remaining keys compare arrivingAt again or
// identity hashes, i.e. no defined order) — the checker sorts by (arrivingAt, from, to, sentAt)
inline std::vector<EnvelopeInfo> Network::MessageStorage::peekMessages() const {
  std::vector<EnvelopeInfo> res;
  std::vector<std::pair<int, int>> tmp;
  for (auto& ms : msgsBySlot)
    for (Envelope* m : ms->msgsByMs)
      for (; m != nullptr; m = m->nextSameTime) {
        tmp.clear();
        m->infos(net, tmp);
        for (auto& da : tmp) res.push_back(EnvelopeInfo{m->fromNodeId, da.first, m->sendTime, da.second, m->message->isTask()});
      }
  std::stable_sort(res.begin(), res.end(), [](const EnvelopeInfo& a, const EnvelopeInfo& b) {
    if (a.arrivingAt != b.arrivingAt) return a.arrivingAt < b.arrivingAt;
    if (a.from != b.from) return a.from < b.from;
    if (a.to != b.to) return a.to < b.to;
    return a.sentAt < b.sentAt;
  });
  return res;
}
inline int MultipleDestEnvelope::nextArrivalTime(const Network& network) const {  // Envelope.java:107-118
  int destId = getNextDestId();
  int rd = Network::getPseudoRandom(destId, randomSeed);
  const Node& f = *network.allNodes[static_cast<size_t>(fromNodeId)];
  const Node& t = *network.allNodes[static_cast<size_t>(destId)];
  return sendTime + network.networkLatency.getLatency(f, t, rd);
}

}  // namespace wo
