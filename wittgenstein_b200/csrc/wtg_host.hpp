// wittgenstein_b200 — host-side model: everything the reference does sequentially at
// construction time on the single network RNG (node attributes, dead-node selection, the
// integer latency tables).  Pure C++17, no CUDA.
//
// Mirrors (file:line under /root/reference/core/src/main/java/net/consensys/wittgenstein/core):
//   Node.java:246-271            node constructor draw order          -> HostModel::buildNodes
//   NodeBuilder.java:77-148      random position / weighted city pick  -> builderX/Y/City
//   geoinfo/Geo.java:10-19, GeoAWS.java:12-22  cumulative probabilities in HashMap order
//   RegistryNodeBuilders.java:21-81, RegistryNetworkLatencies.java:28-58  name registries
//   NetworkLatency.java:49-417   samplers, folded into integer tables (delta in [0,99])
//   utils/GeneralizedParetoDistribution.java:26-46
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "wtg_types.h"

namespace wtg {

// java.util.Random (legacy 48-bit LCG)
struct JRandom {
  static constexpr uint64_t MULT = 0x5DEECE66DULL, ADD = 0xBULL, MASK = (1ULL << 48) - 1;
  uint64_t seed = 0;
  explicit JRandom(int64_t s = 0) { setSeed(s); }
  void setSeed(int64_t s) { seed = ((uint64_t)s ^ MULT) & MASK; }
  int32_t next(int bits) {
    seed = (seed * MULT + ADD) & MASK;
    return (int32_t)(uint32_t)(seed >> (48 - bits));
  }
  int32_t nextInt() { return next(32); }
  int32_t nextInt(int32_t bound) {
    int32_t r = next(31), m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
    for (int32_t u = r;; u = next(31)) {
      r = u % bound;
      if ((int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m) >= 0) return r;
    }
  }
  bool nextBoolean() { return next(1) != 0; }
  double nextDouble() { return (double)(((int64_t)next(26) << 27) + (int64_t)next(27)) * 0x1.0p-53; }
};

inline void lcgJumpTables(uint64_t* a, uint64_t* c) {  // a[i], c[i]: 2^i steps
  uint64_t ba = JRandom::MULT, bc = JRandom::ADD;
  for (int i = 0; i < 48; ++i) {
    a[i] = ba;
    c[i] = bc;
    bc = (bc * ba + bc) & JRandom::MASK;
    ba = (ba * ba) & JRandom::MASK;
  }
}

struct HostNode {
  int x = 1, y = 1, extra = 0, city = 0;
  double speed = 1.0;
  bool down = false;
};

struct AwsCityDef {
  const char* name;
  int x, y, region;
};
// GeoAWS.java:12-22 in put order; region ids of NetworkLatency.java:90-102
static const AwsCityDef kAwsCities[11] = {{"Oregon", 271, 261, 0},   {"Virginia", 513, 316, 1},       {"Mumbai", 1344, 426, 2},
                                          {"Seoul", 1641, 312, 3},   {"Singapore", 1507, 532, 4},     {"Sydney", 1773, 777, 5},
                                          {"Tokyo", 1708, 316, 6},   {"Canada central", 422, 256, 7}, {"Frankfurt", 985, 226, 8},
                                          {"Ireland", 891, 200, 9},  {"London", 937, 205, 10}};

struct HostModel {
  enum BuilderKind { B_BASE, B_RANDOM, B_AWS } builder = B_RANDOM;
  bool uniformSpeed = false;
  bool hasTor = false;
  double tor = 0.0;
  // AWS pick table: HashMap iteration order + float cumulative probabilities
  int awsOrder[11];
  float awsCum[11];

  JRandom rd{0};
  std::vector<HostNode> nodes;

  // latency
  int latKind = LAT_DIST;
  int latParam = 0;
  std::vector<int16_t> latTab, latBase, latJit;
  int latMax = 0;  // upper bound of getLatency over all pairs / deltas (ring sizing)
  bool latencySet = false;

  HostModel() { setLatencyByName("IC3NetworkLatency", false); }  // Network.java:43 default

  // ---- registries ----
  void setBuilderByName(const char* nameOrNull) {
    std::string name = nameOrNull ? nameOrNull : "";
    bool blank = true;
    for (char ch : name)
      if (ch != ' ' && ch != '\t' && ch != '\n') blank = false;
    if (blank) name = "RANDOM_SPEED=CONSTANT_TOR=0.00";
    size_t p1 = name.find("_SPEED="), p2 = name.find("_TOR=");
    if (p1 == std::string::npos || p2 == std::string::npos || p2 < p1) throw std::invalid_argument(name + " not in the registry");
    std::string site = name.substr(0, p1), speed = name.substr(p1 + 7, p2 - p1 - 7), torS = name.substr(p2 + 5);
    static const char* tors[] = {"0.00", "0.01", "0.10", "0.20", "0.33", "0.50", "0.60", "0.80", "1.00"};
    static const double torv[] = {0.0, 0.01, 0.10, 0.20, .33, .5, .6, .8, 1.0};
    double t = -1;
    for (int i = 0; i < 9; ++i)
      if (torS == tors[i]) t = torv[i];
    if (t < 0) throw std::invalid_argument(name + " not in the registry");
    if (site == "AWS") {
      builder = B_AWS;
      buildAwsPick();
    } else if (site == "RANDOM") {
      builder = B_RANDOM;
    } else if (site == "CITIES") {
      throw std::invalid_argument("CITIES node builder needs the WonderNetwork CSV data (not supported yet)");
    } else {
      throw std::invalid_argument(name + " not in the registry");
    }
    if (speed == "GAUSSIAN")
      uniformSpeed = true;  // RegistryNodeBuilders.java:59-61 installs UniformSpeed under this label
    else if (speed == "CONSTANT")
      uniformSpeed = false;
    else
      throw std::invalid_argument(name + " not in the registry");
    hasTor = t > 0.001;
    tor = t;
  }

  static int32_t strHash(const char* s) {
    uint32_t h = 0;
    for (; *s; ++s) h = 31u * h + (unsigned char)*s;
    return (int32_t)h;
  }
  void buildAwsPick() {
    // HashMap<String,..>(16): iteration by bucket (h ^ h>>>16) & 15, insertion order inside a bucket
    int idx[11];
    for (int i = 0; i < 11; ++i) idx[i] = i;
    auto bucket = [](int i) {
      uint32_t h = (uint32_t)strHash(kAwsCities[i].name);
      return (int)((h ^ (h >> 16)) & 15u);
    };
    std::stable_sort(idx, idx + 11, [&](int a, int b) { return bucket(a) < bucket(b); });
    float cum = 0.f;
    for (int i = 0; i < 11; ++i) {
      cum = cum + (float)1 * 1.f / (float)11;  // Geo.java:14 float accumulation
      awsOrder[i] = idx[i];
      awsCum[i] = cum;
    }
  }
  int awsPick(int32_t rdInt) const {  // NodeBuilder.java:128-139
    int32_t a = rdInt == std::numeric_limits<int32_t>::min() ? rdInt : (rdInt < 0 ? -rdInt : rdInt);
    int rand = a % 11;
    float p = (float)rand / (float)11;
    for (int i = 0; i < 11; ++i)
      if (p <= awsCum[i]) return awsOrder[i];
    throw std::runtime_error("no city for draw");
  }

  // Node constructors for `count` nodes, appended (ids continue)
  void buildNodes(int count) {
    for (int i = 0; i < count; ++i) {
      HostNode n;
      int32_t rdNode = rd.nextInt();
      if (builder == B_RANDOM) {
        int64_t r = (int64_t)(rdNode >> 16);  // NodeBuilder.java:82-87
        r = r < 0 ? -r : r;
        n.x = (int)(r % 2000 + 1);
        int64_t r2 = (int64_t)(int32_t)((uint32_t)rdNode << 16);  // :90-95
        r2 = r2 < 0 ? -r2 : r2;
        n.y = (int)(r2 % 1112 + 1);
      } else if (builder == B_AWS) {
        int c = awsPick(rdNode);
        n.city = kAwsCities[c].region;
        n.x = kAwsCities[c].x;
        n.y = kAwsCities[c].y;
      }
      if (uniformSpeed)  // Node.java:233-238
        n.speed = rd.nextBoolean() ? (rd.nextInt(67) + 33) / 100.0 : (rd.nextInt(200) + 100) / 100.0;
      if (hasTor) n.extra = rd.nextDouble() < tor ? 500 : 0;  // Node.java:158-160
      nodes.push_back(n);
    }
  }

  // ---- latency ----
  static double gpdInverse(double y) {  // GPD(1.4, -0.3, 0.35), GeneralizedParetoDistribution.java:26-46
    const double shape = 1.4, location = -0.3, scale = 0.35;
    if (y < 0.000001) return location;
    if (y > 0.999999) return std::numeric_limits<double>::infinity();
    return location + scale / shape * (-1 + std::pow(1 - y, -shape));
  }
  void setMeasured(const int* props, const int* vals, int n, std::vector<int16_t>& out) {  // NetworkLatency.java:284-303
    out.assign(100, 0);
    int li = 0, cur = 0, sum = 0;
    for (int i = 0; i < n; i++) {
      if (props[i] == 0) {
        cur = vals[i];
        continue;
      }
      sum += props[i];
      int step = (vals[i] - cur) / props[i];
      for (int ii = 0; ii < props[i]; ii++) {
        cur += step;
        if (li >= 100) throw std::invalid_argument("latency distribution does not sum to 100");
        if (cur > 32767 || cur < -32768) throw std::invalid_argument("latency value out of range");
        out[(size_t)li++] = (int16_t)cur;
      }
    }
    if (sum != 100 || li != 100) throw std::invalid_argument("latency distribution does not sum to 100");
  }
  void setLatencyMeasured(const int* props, const int* vals, int n) {
    setMeasured(props, vals, n, latTab);
    latKind = LAT_DELTA;
    latMax = *std::max_element(latTab.begin(), latTab.end()) + 1000;
    latencySet = true;
  }
  void setLatencyByName(const char* nameOrNull, bool user = true) {
    std::string name = nameOrNull ? nameOrNull : "NetworkLatencyByDistanceWJitter";
    latTab.clear();
    latBase.clear();
    latJit.clear();
    for (int f : {0, 100, 200, 500, 1000, 2000, 4000, 8000}) {
      if (name == "NetworkFixedLatency(" + std::to_string(f) + ")") {
        latKind = LAT_CONST;
        latParam = std::max(1, f);
        latMax = latParam + 1000;
        latencySet = user;
        return;
      }
      if (name == "NetworkUniformLatency(" + std::to_string(f) + ")") {
        latKind = LAT_DELTA;
        int mx = std::max(1, f);
        latTab.resize(100);
        for (int d = 0; d < 100; ++d) latTab[(size_t)d] = (int16_t)(int)((d / 99.0) * mx);
        latMax = mx + 1000;
        latencySet = user;
        return;
      }
    }
    const int md = MAX_DIST;
    if (name == "NetworkLatencyByDistanceWJitter") {  // :49-73
      latKind = LAT_DIST_DELTA;
      latTab.resize((size_t)(md + 1) * 100);
      const double earthPerimeter = 24860;
      const double pointValue = (earthPerimeter / 2) / md;
      latMax = 0;
      for (int dist = 0; dist <= md; ++dist)
        for (int d = 0; d < 100; ++d) {
          double fixedLat = pointValue * dist * 0.022 + 4.862;
          double raw = fixedLat + gpdInverse(d / 100.0);
          int v = (int)(raw / 2);
          latTab[(size_t)dist * 100 + (size_t)d] = (int16_t)v;
          latMax = std::max(latMax, v);
        }
      latMax += 1000;
    } else if (name == "AwsRegionNetworkLatency") {  // :86-152
      latKind = LAT_CITY;
      static const int lat[10][11] = {
          {0, 81, 216, 126, 165, 138, 97, 64, 164, 131, 141}, {0, 0, 182, 181, 232, 195, 167, 13, 88, 80, 75},
          {0, 0, 0, 152, 62, 223, 123, 194, 111, 122, 113},   {0, 0, 0, 0, 97, 133, 35, 184, 259, 254, 264},
          {0, 0, 0, 0, 0, 169, 69, 218, 162, 174, 171},       {0, 0, 0, 0, 0, 0, 105, 210, 282, 269, 271},
          {0, 0, 0, 0, 0, 0, 0, 156, 235, 222, 234},          {0, 0, 0, 0, 0, 0, 0, 0, 101, 78, 87},
          {0, 0, 0, 0, 0, 0, 0, 0, 0, 24, 13},                {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 12}};
      latBase.assign(121, 0);
      int mxb = 0;
      for (int a = 0; a < 11; ++a)
        for (int b = 0; b < 11; ++b)
          if (a != b) {
            int v = lat[std::min(a, b)][std::max(a, b)] / 2;
            latBase[(size_t)a * 11 + (size_t)b] = (int16_t)v;
            mxb = std::max(mxb, v);
          }
      latJit.resize(100);
      for (int d = 0; d < 100; ++d) latJit[(size_t)d] = (int16_t)(int)gpdInverse(d / 100.0);
      latMax = mxb + latJit[99] + 1000;
    } else if (name == "NetworkNoLatency") {
      latKind = LAT_CONST;
      latParam = 1;
      latMax = 1001;
    } else if (name == "EthScanNetworkLatency") {  // :366-383
      static const int p[] = {16, 18, 17, 12, 8, 5, 4, 3, 3, 1, 1, 2, 1, 1, 8};
      static const int v[] = {250, 500, 1000, 1250, 1500, 1750, 2000, 2250, 2500, 2750, 4500, 6000, 8500, 9750, 10000};
      setMeasured(p, v, 15, latTab);
      latKind = LAT_DELTA_2X;
      latMax = 10000 + 2000;
    } else if (name == "IC3NetworkLatency") {  // :399-417
      latKind = LAT_DIST;
      latTab.resize((size_t)md + 1);
      for (int dist = 0; dist <= md; ++dist) {
        double dd = dist;
        double surface = dd * dd * M_PI;
        double totalSurface = 2000 * 1112;
        int position = (int)((surface * 100) / totalSurface);
        int v = position <= 10 ? 92 / 2 : position <= 33 ? 125 / 2 : position <= 50 ? 152 / 2 : position <= 67 ? 200 / 2 : position <= 90 ? 276 / 2 : 350 / 2;
        latTab[(size_t)dist] = (int16_t)v;
      }
      latMax = 175 + 1000;
    } else {
      throw std::invalid_argument("latency '" + name + "' is not available (NetworkLatencyByCity* need CSV data)");
    }
    latencySet = user;
  }
};

}  // namespace wtg
