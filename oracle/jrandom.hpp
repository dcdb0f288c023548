// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into the product library.
//
// Restatement of the JDK primitives the reference's hot path depends on
// (SURVEY.md §8c): java.util.Random (OpenJDK 9+, legacy LCG algorithm),
// Collections.shuffle(List, Random), and an O(log n) jump for the LCG.
// The JDK is a third-party dependency that is not under /root/reference; the
// algorithm below is the published one (java/util/Random.java):
//   seed  = (s ^ 0x5DEECE66D) & (2^48-1)
//   next(b): seed = (seed * 0x5DEECE66D + 0xB) & (2^48-1); return (int)(seed >>> (48-b))
// Call sites in the reference: core/Network.java:55,377,430; core/Node.java:159,237,252;
// protocols/GSFSignature.java:470,618.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace wo {

struct JavaRandom {
  static constexpr uint64_t MULT = 0x5DEECE66DULL;
  static constexpr uint64_t ADD = 0xBULL;
  static constexpr uint64_t MASK = (1ULL << 48) - 1;
  uint64_t seed;
  uint64_t draws = 0;  // number of next() calls, for statistics / parity of stream position

  explicit JavaRandom(int64_t s = 0) { setSeed(s); }
  void setSeed(int64_t s) { seed = (static_cast<uint64_t>(s) ^ MULT) & MASK; }

  int32_t next(int bits) {
    seed = (seed * MULT + ADD) & MASK;
    ++draws;
    return static_cast<int32_t>(static_cast<uint32_t>(seed >> (48 - bits)));
  }
  int32_t nextInt() { return next(32); }

  // java.util.Random.nextInt(int bound)
  int32_t nextInt(int32_t bound) {
    int32_t r = next(31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) {
      r = static_cast<int32_t>((static_cast<int64_t>(bound) * static_cast<int64_t>(r)) >> 31);
    } else {
      // for (int u = r; u - (r = u % bound) + m < 0; u = next(31));   (int overflow intended)
      int32_t u = r;
      for (;;) {
        r = u % bound;
        uint32_t t = static_cast<uint32_t>(u) - static_cast<uint32_t>(r) + static_cast<uint32_t>(m);
        if (static_cast<int32_t>(t) >= 0) break;
        u = next(31);
      }
    }
    return r;
  }
  bool nextBoolean() { return next(1) != 0; }
  double nextDouble() {
    int64_t hi = static_cast<int64_t>(next(26)) << 27;
    int64_t lo = next(27);
    return static_cast<double>(hi + lo) * 0x1.0p-53;
  }
};

// Collections.shuffle(list, rnd): for (i = size; i > 1; i--) swap(list, i-1, rnd.nextInt(i));
template <class T>
inline void javaShuffle(std::vector<T>& v, JavaRandom& rd) {
  for (int32_t i = static_cast<int32_t>(v.size()); i > 1; i--) {
    int32_t j = rd.nextInt(i);
    std::swap(v[i - 1], v[j]);
  }
}
template <class T>
inline void javaShuffle(T* v, int32_t size, JavaRandom& rd) {
  for (int32_t i = size; i > 1; i--) {
    int32_t j = rd.nextInt(i);
    std::swap(v[i - 1], v[j]);
  }
}

// LCG jump: state after n steps = a_n * s + c_n (mod 2^48).
struct LcgJump {
  uint64_t a, c;
};
inline LcgJump lcgJump(uint64_t n) {
  uint64_t a = 1, c = 0;                              // identity
  uint64_t ba = JavaRandom::MULT, bc = JavaRandom::ADD;  // one step
  while (n) {
    if (n & 1) {
      a = (a * ba) & JavaRandom::MASK;
      c = (c * ba + bc) & JavaRandom::MASK;
    }
    bc = (bc * ba + bc) & JavaRandom::MASK;
    ba = (ba * ba) & JavaRandom::MASK;
    n >>= 1;
  }
  return {a, c};
}
inline uint64_t lcgAdvance(uint64_t seed, uint64_t n) {
  LcgJump j = lcgJump(n);
  return (seed * j.a + j.c) & JavaRandom::MASK;
}

}  // namespace wo
