// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// java.util.BitSet value semantics (a set of non-negative ints) as used by
// protocols/GSFSignature.java (cardinality/intersects/equals/or/and/andNot/clone/
// nextSetBit/set(from,to)).  Storage is range-compressed (only the words between the
// lowest and highest touched word are kept) so that large-N oracle runs fit in memory;
// observable behaviour is identical to java.util.BitSet.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace wo {

class JBitSet {
  int64_t lo_ = 0;               // index of first stored word
  std::vector<uint64_t> w_;      // stored words [lo_, lo_+w_.size())

  void ensure(int64_t wfrom, int64_t wto) {  // make [wfrom, wto) addressable
    if (w_.empty()) {
      lo_ = wfrom;
      w_.assign(static_cast<size_t>(wto - wfrom), 0);
      return;
    }
    int64_t hi = lo_ + static_cast<int64_t>(w_.size());
    if (wfrom < lo_) {
      w_.insert(w_.begin(), static_cast<size_t>(lo_ - wfrom), 0);
      lo_ = wfrom;
    }
    if (wto > hi) w_.resize(static_cast<size_t>(wto - lo_), 0);
  }
  uint64_t word(int64_t wi) const {
    int64_t k = wi - lo_;
    return (k < 0 || k >= static_cast<int64_t>(w_.size())) ? 0 : w_[static_cast<size_t>(k)];
  }

 public:
  JBitSet() = default;
  void set(int i) {
    int64_t wi = i >> 6;
    ensure(wi, wi + 1);
    w_[static_cast<size_t>(wi - lo_)] |= 1ULL << (i & 63);
  }
  void clear(int i) {
    int64_t k = (i >> 6) - lo_;
    if (k >= 0 && k < static_cast<int64_t>(w_.size())) w_[static_cast<size_t>(k)] &= ~(1ULL << (i & 63));
  }
  void set(int i, bool v) { v ? set(i) : clear(i); }
  // set(fromIndex inclusive, toIndex exclusive)
  void setRange(int from, int to) {
    if (from >= to) return;
    int64_t wf = from >> 6, wt = (to - 1) >> 6;
    ensure(wf, wt + 1);
    for (int64_t wi = wf; wi <= wt; ++wi) {
      uint64_t m = ~0ULL;
      if (wi == wf) m &= ~0ULL << (from & 63);
      if (wi == wt) m &= ~0ULL >> (63 - ((to - 1) & 63));
      w_[static_cast<size_t>(wi - lo_)] |= m;
    }
  }
  bool get(int i) const { return (word(i >> 6) >> (i & 63)) & 1; }
  int cardinality() const {
    int c = 0;
    for (uint64_t x : w_) c += __builtin_popcountll(x);
    return c;
  }
  bool isEmpty() const {
    for (uint64_t x : w_)
      if (x) return false;
    return true;
  }
  void or_(const JBitSet& o) {
    if (o.w_.empty()) return;
    ensure(o.lo_, o.lo_ + static_cast<int64_t>(o.w_.size()));
    size_t off = static_cast<size_t>(o.lo_ - lo_);
    for (size_t k = 0; k < o.w_.size(); ++k) w_[off + k] |= o.w_[k];
  }
  void and_(const JBitSet& o) {
    for (size_t k = 0; k < w_.size(); ++k) w_[k] &= o.word(lo_ + static_cast<int64_t>(k));
  }
  void andNot(const JBitSet& o) {
    for (size_t k = 0; k < w_.size(); ++k) w_[k] &= ~o.word(lo_ + static_cast<int64_t>(k));
  }
  bool intersects(const JBitSet& o) const {
    int64_t a = std::max(lo_, o.lo_);
    int64_t b = std::min(lo_ + static_cast<int64_t>(w_.size()), o.lo_ + static_cast<int64_t>(o.w_.size()));
    for (int64_t wi = a; wi < b; ++wi)
      if (word(wi) & o.word(wi)) return true;
    return false;
  }
  bool equals(const JBitSet& o) const {
    int64_t a = std::min(lo_, o.lo_);
    int64_t b = std::max(lo_ + static_cast<int64_t>(w_.size()), o.lo_ + static_cast<int64_t>(o.w_.size()));
    if (w_.empty() && o.w_.empty()) return true;
    if (w_.empty()) return o.isEmpty();
    if (o.w_.empty()) return isEmpty();
    for (int64_t wi = a; wi < b; ++wi)
      if (word(wi) != o.word(wi)) return false;
    return true;
  }
  // nextSetBit(fromIndex): -1 if none
  int nextSetBit(int from) const {
    if (w_.empty()) return -1;
    int64_t wi = std::max<int64_t>(from >> 6, lo_);
    int64_t hi = lo_ + static_cast<int64_t>(w_.size());
    if (wi >= hi) return -1;
    uint64_t x = word(wi);
    if (wi == (from >> 6)) x &= ~0ULL << (from & 63);
    for (;;) {
      if (x) return static_cast<int>(wi * 64 + __builtin_ctzll(x));
      if (++wi >= hi) return -1;
      x = word(wi);
    }
  }
  // raw word access for state export (word index in the absolute bit space)
  uint64_t rawWord(int64_t wi) const { return word(wi); }
  size_t storedBytes() const { return w_.size() * 8; }
};

}  // namespace wo
