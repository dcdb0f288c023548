// wittgenstein_b200 — bulk asynchronous copies (the 1-D form of the Tensor Memory Accelerator: `cp.async.bulk`, SASS UBLKCP)
// for the payload snapshots of the tick engine.  One elected lane moves a whole level block global -> shared -> global
// with up to four 2-KiB tiles in flight per warp and no register staging; completion of the loads is tracked by
// mbarriers in shared memory (expect_tx / try_wait.parity), of the stores by bulk async-groups.
#pragma once
#include <stdint.h>

namespace wtg {
#if defined(__CUDACC__)

constexpr int TMA_TILE_BYTES = 2048;  // 16-byte multiples; a level block of l >= 14 is a whole number of tiles at >= 1 KiB
constexpr int TMA_TILES = 2;          // tiles (and mbarriers) per warp: 4 KiB in flight per warp, 96 KiB of shared memory per SM at 3 blocks
constexpr int TMA_MIN_BYTES = 1024;   // below this the lane-strided register copy is as good

struct TmaWarp {  // per-warp context, lives in the CoopWarp of the kernels that copy payloads
  unsigned long long* buf = nullptr;  // [TMA_TILES][TMA_TILE_BYTES / 8] shared memory
  unsigned long long* bar = nullptr;  // [TMA_TILES] mbarriers
  uint32_t phase = 0;                 // bit t: parity the next completion of bar[t] will have
  bool ready = false;                 // mbarriers initialised (by the lane that issues the copies)
};

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarInitFence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WTG_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WTG_DONE_%=;\n"
      "bra WTG_WAIT_%=;\n"
      "WTG_DONE_%=:\n"
      "}\n" ::"r"(smemAddr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(smemDst)), "l"(gsrc),
               "r"(bytes), "r"(smemAddr(bar))
               : "memory");
}
__device__ __forceinline__ void bulkStore(void* gdst, const void* smemSrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smemAddr(smemSrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulkWaitRead0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }  // sources reusable
__device__ __forceinline__ void bulkWaitAll() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }        // stores complete

// one lane: load `bytes` (<= TMA_TILES tiles; a multiple of 16, 16-byte aligned) from global into the warp's tiles and wait
__device__ __forceinline__ void tmaLoadLane(TmaWarp& t, const void* src, uint32_t bytes) {
  if (!t.ready) {
    for (int i = 0; i < TMA_TILES; ++i) mbarInit(t.bar + i, 1);
    mbarInitFence();
    t.ready = true;
  }
  const char* s = (const char*)src;
  const int tiles = (int)((bytes + TMA_TILE_BYTES - 1) / TMA_TILE_BYTES);
#pragma unroll
  for (int i = 0; i < TMA_TILES; ++i)
    if (i < tiles) {  // all loads in flight at once
      const uint32_t sz = bytes - i * TMA_TILE_BYTES < (uint32_t)TMA_TILE_BYTES ? bytes - i * TMA_TILE_BYTES : (uint32_t)TMA_TILE_BYTES;
      mbarExpectTx(t.bar + i, sz);
      bulkLoad(t.buf + (size_t)i * (TMA_TILE_BYTES / 8), s + i * TMA_TILE_BYTES, sz, t.bar + i);
    }
#pragma unroll
  for (int i = 0; i < TMA_TILES; ++i)
    if (i < tiles) {
      mbarWait(t.bar + i, (t.phase >> i) & 1u);
      t.phase ^= 1u << i;
    }
}

#endif
}  // namespace wtg
