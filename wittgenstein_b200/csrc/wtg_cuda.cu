// wittgenstein_b200 — CUDA backend (sm_100a): the kernels of the tick pipeline and the C ABI.
//
// One simulated millisecond = one pass of this kernel sequence over SoA state resident in HBM:
//   k_begin -> k_cond_mark / k_cond_nodes<scan> / k_cond_score / k_cond_nodes<select> (conditional tasks) ->
//   k_dispatch_count -> pair scan A -> k_dispatch_scatter -> k_node_msgs (thread per node) / k_node_tasks (warp per
//   node) -> pair scan B -> k_emit (seed / latency / arrival) -> multisplit (count, column scan, stable scatter
//   into the time ring) -> k_free -> k_end
// All sizes are read from the device control block, so a whole runMs window is enqueued without
// a host round trip.  See DESIGN.md §4 for why this reproduces the reference's sequential order.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "wtg_engine.hpp"

namespace cg = cooperative_groups;
namespace wtg {

#define CUDA_OK(x)                                                                                       \
  do {                                                                                                   \
    cudaError_t e_ = (x);                                                                                \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #x); \
  } while (0)

#if defined(WTG_TMA_SNAPSHOT)
constexpr int NODE_TMA_SMEM = 8 * TMA_TILES * TMA_TILE_BYTES;
#else
constexpr int NODE_TMA_SMEM = 0;
#endif
//  // dynamic shared memory of k_node_tasks (GSF): bulk-copy tiles
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ void b_begin(const Dev& d, int mode, const int VB, const int VG) {  // one warp
  if (threadIdx.x >= 32) return;
  if (d.ffwd && mode == 1) {
    CoopWarp c;
    tickBeginFfwd(d, c);
  } else if (d.farCap > 0 && !d.ffwd) {
    CoopWarp c;
    tickBeginFar(d, c, mode);
  } else if (threadIdx.x == 0) {
    tickBegin(d, mode);
  }
}
__global__ void k_begin(Dev d, int mode) { b_begin(d, mode, blockIdx.x, gridDim.x); }
__device__ __forceinline__ void b_end(const Dev& d, int mode, const int VB, const int VG) {
  if (threadIdx.x == 0) tickEnd(d, mode);
}
__global__ void k_end(Dev d, int mode) { b_end(d, mode, blockIdx.x, gridDim.x); }

// ---- conditional tasks (checkSigs): scan -> score -> select ---------------------------------------
// append node n to a striped list (stripe = global warp index & 63; a stripe receives at most listStripeCap nodes)
__device__ __forceinline__ void listAppend(const Dev& d, bool active, int n, int* cnt, int* list, int VB) {
  unsigned m = __ballot_sync(0xffffffffu, active);
  if (!m) return;
  int lane = threadIdx.x & 31;
  int gw = (VB * blockDim.x + threadIdx.x) >> 5;
  int stripe = gw & (ARENA_STRIPES - 1);
  int base = 0;
  if (lane == 0) base = atomicAdd(&cnt[stripe], __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (active) list[(size_t)stripe * d.listStripeCap + base + __popc(m & ((1u << lane) - 1u))] = n;
}
// same, with a 64-bit payload per entry
__device__ __forceinline__ void listAppendW(const Dev& d, bool active, int n, u64 word, int* cnt, int* list, u64* words, int VB) {
  unsigned m = __ballot_sync(0xffffffffu, active);
  if (!m) return;
  int lane = threadIdx.x & 31;
  int gw = (VB * blockDim.x + threadIdx.x) >> 5;
  int stripe = gw & (ARENA_STRIPES - 1);
  int base = 0;
  if (lane == 0) base = atomicAdd(&cnt[stripe], __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (active) {
    size_t at = (size_t)stripe * d.listStripeCap + base + __popc(m & ((1u << lane) - 1u));
    list[at] = n;
    words[at] = word;
  }
}
// conditional-task bookkeeping, one thread per node -> list of due nodes
__device__ __forceinline__ void b_cond_mark(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  int n = d.n0 + VB * blockDim.x + threadIdx.x;
  bool due = false;
  if (n < d.n0 + d.nLoc) due = d.proto == PROTO_HANDEL ? hCondMark(d, n) : gsfCondMark(d, n);
  listAppend(d, due, n, d.ctl->dueCnt, d.dueList, VB);
}
__global__ void __launch_bounds__(256) k_cond_mark(Dev d) { b_cond_mark(d, blockIdx.x, gridDim.x); }
// one warp per due node, blocks assigned to list stripes
template <int PHASE>
__device__ __forceinline__ void b_cond_nodes(const Dev& d, const int VB, const int VG) {
  extern __shared__ uint32_t keepAll[];
  __shared__ HScratch scratch[8];
  if (d.ctl->error) return;
  const int stripe = VB & (ARENA_STRIPES - 1);
  const int cnt = d.ctl->dueCnt[stripe];
  const int warp = threadIdx.x >> 5;
  const int sub = (VB >> 6) * 8 + warp;
  const int nsub = (VG >> 6) * 8;
  const int* list = d.dueList + (size_t)stripe * d.listStripeCap;
  CoopWarp c;
  for (int t = sub; t < cnt; t += nsub) {
    int n = list[t];
    if (PHASE == 0) {
      if (d.proto == PROTO_HANDEL)
        hCondScanQueue(d, c, n);
      else
        gsfCondScanQueue(d, c, n);
    } else {
      if (d.proto == PROTO_HANDEL)
        hCondSelect(d, c, n, &scratch[warp]);
      else
        gsfCondSelect(d, c, n, keepAll + (size_t)warp * (size_t)(d.qcap / 32));
    }
  }
}
template <int PHASE>
__global__ void __launch_bounds__(256) k_cond_nodes(Dev d) { b_cond_nodes<PHASE>(d, blockIdx.x, gridDim.x); }
// blocks are assigned to arena stripes (blockIdx & 63), so an item is found without walking the 64 counters
__device__ __forceinline__ void b_cond_score(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  const int per = d.workCap / ARENA_STRIPES;
  const int stripe = VB & (ARENA_STRIPES - 1);
  int cnt = d.ctl->workCnt[stripe];
  if (cnt > per) cnt = per;
  const int warpsPerBlock = blockDim.x >> 5;
  const int sub = (VB >> 6) * warpsPerBlock + (threadIdx.x >> 5);
  const int nsub = (VG >> 6) * warpsPerBlock;
  CoopWarp c;
  const uint32_t* wl = d.workList + (size_t)stripe * per;
  for (int t = sub; t < cnt; t += nsub) {
    uint32_t it = wl[t];
    if (d.proto == PROTO_HANDEL)
      hScoreItem(d, c, it);
    else
      gsfScoreItem(d, c, it);
  }
}
__global__ void __launch_bounds__(256) k_cond_score(Dev d) { b_cond_score(d, blockIdx.x, gridDim.x); }
// ---- Handel conditional pass (checkSigs): scan -> score -> select -> draw scan -> pick ---------------------
// does any nextInt(k) of this pass hit java.util.Random's rejection loop?  (probability ~ k / 2^31 per draw)
__device__ __forceinline__ void b_hpick_check(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  for (int n = d.n0 + VB * blockDim.x + threadIdx.x; n < d.n0 + d.nLoc; n += VG * blockDim.x)
    if (d.hCandK[n] > 0 && hCondPick(d, n, (u64)d.hDrawBase[n], false) > d.condDraws[n]) d.ctl->hReject = 1;
}
__global__ void k_hpick_check(Dev d) { b_hpick_check(d, blockIdx.x, gridDim.x); }
__device__ __forceinline__ void b_hpick_apply(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  if (!d.ctl->hReject) {
    for (int n = d.n0 + VB * blockDim.x + threadIdx.x; n < d.n0 + d.nLoc; n += VG * blockDim.x) hCondPick(d, n, (u64)d.hDrawBase[n], true);
  } else if (VB == 0 && threadIdx.x == 0) {  // a rejection shifts every later draw: redo the picks in node order
    u64 idx = 0;
    for (int n = d.n0; n < d.n0 + d.nLoc; ++n) idx += (u64)hCondPick(d, n, idx, true);
  }
}
__global__ void k_hpick_apply(Dev d) { b_hpick_apply(d, blockIdx.x, gridDim.x); }

// ---- dispatch -----------------------------------------------------------------------------
__device__ __forceinline__ void b_dispatch_count(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  int nEv = d.ctl->nEv;
  if (d.allCap > 0) {  // sendAll protocols: a warp per bucket entry
    CoopWarp c;
    int gw = (VB * blockDim.x + threadIdx.x) >> 5, nw = (VG * blockDim.x) >> 5;
    for (int i = gw; i < nEv; i += nw) dispatchCountCoop(d, c, i);
    return;
  }
  for (int i = VB * blockDim.x + threadIdx.x; i < nEv; i += VG * blockDim.x) dispatchCount(d, i);
}
__global__ void k_dispatch_count(Dev d) { b_dispatch_count(d, blockIdx.x, gridDim.x); }
__device__ __forceinline__ void b_dispatch_scatter(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  int nEv = d.ctl->nEv;
  if (d.allCap > 0) {
    CoopWarp c;
    int gw = (VB * blockDim.x + threadIdx.x) >> 5, nw = (VG * blockDim.x) >> 5;
    for (int i = gw; i < nEv; i += nw) dispatchScatterCoop(d, c, i);
    return;
  }
  for (int i = VB * blockDim.x + threadIdx.x; i < nEv; i += VG * blockDim.x) dispatchScatter(d, i);
}
__global__ void k_dispatch_scatter(Dev d) { b_dispatch_scatter(d, blockIdx.x, gridDim.x); }

// ---- handlers: warp per node ----------------------------------------------------------------
// pass 1: one thread per node.  GSF / PingPong: message deliveries (they commute with the node's tasks, see
// nodeProcess); SanFermin: everything (all handlers are scalar).  Leaves nodeTasks[n] = 1 when a warp is needed.
__device__ __forceinline__ void b_node_msgs(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  int n = d.n0 + VB * blockDim.x + threadIdx.x;
  int flag = 0;
  u64 word = ~0ULL;
  if (n < d.n0 + d.nLoc && d.inboxFill[n] > 0) {
    CoopSerial cs;
    if (d.proto == PROTO_SANFERMIN || d.proto == PROTO_CAPPOS)
      nodeProcess(d, cs, n, 0);
    else if (d.proto == PROTO_GSF || d.proto == PROTO_PINGPONG) {
      u64 w = 0;
      int tasks = nodeProcess(d, cs, n, 1, &w);
      flag = tasks > 0 ? 1 : 0;
      if (tasks == 1) word = w;
    }
    else if (d.proto == PROTO_CASPER) {  // an inbox of attestations only is scalar work; blocks and tasks get a warp
      const u64* in = d.inbox + d.inboxOff[n];
      const Ev* bucket = d.buckets + (size_t)(d.ctl->tick & (d.ring - 1)) * (size_t)d.bcap;
      int cnt = d.inboxFill[n];
      bool simple = true;
      for (int r = 0; r < cnt && simple; ++r) {
        const Ev& ev = bucket[inboxEntry(in[r])];
        uint32_t meta = ev.kind == EV_MULTI ? d.rec[ev.aux].meta : ev.meta;
        simple = (ev.kind == EV_MSG || ev.kind == EV_MULTI) && meta == CM_ATT;
      }
      if (simple)
        nodeProcess(d, cs, n, 0);
      else
        flag = 1;
    } else
      flag = 1;
  }
  listAppendW(d, flag != 0, n, word, d.ctl->taskCnt, d.taskList, d.taskWord, VB);
}
__global__ void __launch_bounds__(256) k_node_msgs(Dev d) { b_node_msgs(d, blockIdx.x, gridDim.x); }
// pass 2: one warp per node that has tasks (updateVerifiedSignatures / doCycle / ...), or, for protocols whose
// events do not commute (Handel), all of the node's events in reference order; blocks assigned to list stripes.
__device__ __forceinline__ void b_node_tasks(const Dev& d, const int VB, const int VG) {
  // bulk-copy tiles and their mbarriers, one set per warp (wtg_tma.cuh): payload snapshots of doCycle
  CoopWarp c;
#if defined(WTG_TMA_SNAPSHOT)
  extern __shared__ __align__(128) unsigned long long tmaTiles[];  // [8 warps][TMA_TILES * TMA_TILE_BYTES / 8] (dynamic: NODE_TMA_SMEM)
  __shared__ __align__(8) unsigned long long tmaBars[8][TMA_TILES];
  if (d.proto == PROTO_GSF) {  // the mbarriers are initialised by the first snapshot of the warp (most launches have none)
    const int warp = threadIdx.x >> 5;
    c.tma.buf = tmaTiles + (size_t)warp * (TMA_TILES * TMA_TILE_BYTES / 8);
    c.tma.bar = &tmaBars[warp][0];
  }
#endif
  if (d.ctl->error) return;
  const bool split = d.proto == PROTO_GSF || d.proto == PROTO_PINGPONG;
  const int stripe = VB & (ARENA_STRIPES - 1);
  const int cnt = d.ctl->taskCnt[stripe];
  const int sub = (VB >> 6) * 8 + (threadIdx.x >> 5);
  const int nsub = (VG >> 6) * 8;
  const int* list = d.taskList + (size_t)stripe * d.listStripeCap;
  const u64* words = d.taskWord + (size_t)stripe * d.listStripeCap;
  for (int t = sub; t < cnt; t += nsub) {
    const int n = list[t];
    const u64 w = words[t];
    if (split && w != ~0ULL)
      nodeSingleTask(d, c, n, w);
    else
      nodeProcess(d, c, n, split ? 2 : 0);
  }
#if defined(WTG_TMA_SNAPSHOT)
  if (c.tma.ready && (threadIdx.x & 31) == 0) bulkWaitAll();
#endif  // this lane's bulk stores are complete before the kernel ends
}
// three blocks per SM = 80 registers: measured best (profiles/README.md, round 2: 64 / 80 / 128 registers -> 258 / 251 / 382 ms)
__global__ void __launch_bounds__(256, 3) k_node_tasks(Dev d) { b_node_tasks(d, blockIdx.x, gridDim.x); }
// CasperIMD, after the parallel handler pass (one warp; both are rare): nodes that hit a fork-choice tie run in processing
// order with their exact draw index (randomOnTies), and several blocks created in one millisecond get their ids in
// processing order
__device__ __forceinline__ void b_casper_fixups(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error || VB != 0 || threadIdx.x >= 32) return;
  CoopWarp c;
  if (d.cRandomTies && d.ctl->tieCnt > 0) casperResolveTies(d, c);
  if (d.G == 1 && d.cg->createdThisTick > 1) casperRenumber(d, c);
}
__global__ void k_casper_fixups(Dev d) { b_casper_fixups(d, blockIdx.x, gridDim.x); }
// ---- pair scans ---------------------------------------------------------------------------
__device__ __forceinline__ Pair pairAdd(Pair x, Pair y) {
  Pair r;
  r.a = x.a + y.a;
  r.b = x.b + y.b;
  return r;
}
// exclusive scan of one value per thread across the block; returns the block total in `total`
__device__ __forceinline__ Pair blockExclusive(Pair v, Pair& total) {
  __shared__ Pair warpSums[33];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  Pair inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int ta = __shfl_up_sync(0xffffffffu, inc.a, o), tb = __shfl_up_sync(0xffffffffu, inc.b, o);
    if (lane >= o) {
      inc.a += ta;
      inc.b += tb;
    }
  }
  __syncthreads();  // warpSums may still be read by a previous call
  if (lane == 31) warpSums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    Pair w;
    w.a = lane < nw ? warpSums[lane].a : 0;
    w.b = lane < nw ? warpSums[lane].b : 0;
    Pair wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int ta = __shfl_up_sync(0xffffffffu, wi.a, o), tb = __shfl_up_sync(0xffffffffu, wi.b, o);
      if (lane >= o) {
        wi.a += ta;
        wi.b += tb;
      }
    }
    Pair ex;
    ex.a = wi.a - w.a;
    ex.b = wi.b - w.b;
    warpSums[lane] = ex;
    if (lane == 31) warpSums[32] = wi;
  }
  __syncthreads();
  total = warpSums[32];
  Pair base = warpSums[warp];
  Pair r;
  r.a = base.a + inc.a - v.a;
  r.b = base.b + inc.b - v.b;
  return r;
}

__device__ __forceinline__ void b_scan_partial(const Dev& d, int which, const int VB, const int VG) {
  if (d.ctl->error) return;
  int M = scanCount(d, which);
  int nTiles = (M + SCAN_TILE - 1) / SCAN_TILE;
  for (int tile = VB; tile < nTiles; tile += VG) {
    int j0 = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    Pair s;
    s.a = 0;
    s.b = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
      if (j0 + k < M) s = pairAdd(s, scanLoad(d, which, j0 + k));
    Pair total;
    blockExclusive(s, total);
    if (threadIdx.x == 0) {
      d.scanPartial[2 * tile] = total.a;
      d.scanPartial[2 * tile + 1] = total.b;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partial(Dev d, int which) { b_scan_partial(d, which, blockIdx.x, gridDim.x); }
// second (last) scan kernel: every block first sums the partials of the tiles before its own (a few hundred
// pairs at most), then scans its tile; the block of the last tile also publishes the totals.
__device__ __forceinline__ void b_scan_final(const Dev& d, int which, const int VB, const int VG) {
  if (d.ctl->error) return;
  int M = scanCount(d, which);
  int nTiles = (M + SCAN_TILE - 1) / SCAN_TILE;
  if (nTiles == 0 && VB == 0 && threadIdx.x == 0) {
    Pair z;
    z.a = 0;
    z.b = 0;
    scanTotals(d, which, z);
  }
  for (int tile = VB; tile < nTiles; tile += VG) {
    Pair pre;
    pre.a = 0;
    pre.b = 0;
    for (int t = threadIdx.x; t < tile; t += SCAN_THREADS) {
      pre.a += d.scanPartial[2 * t];
      pre.b += d.scanPartial[2 * t + 1];
    }
    Pair tileBase;
    blockExclusive(pre, tileBase);  // tileBase = sum over all threads
    int j0 = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    Pair v[SCAN_ITEMS];
    Pair s;
    s.a = 0;
    s.b = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
      if (j0 + k < M)
        v[k] = scanLoad(d, which, j0 + k);
      else {
        v[k].a = 0;
        v[k].b = 0;
      }
      s = pairAdd(s, v[k]);
    }
    Pair total;
    Pair ex = blockExclusive(s, total);
    ex = pairAdd(ex, tileBase);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
      if (j0 + k < M) scanStore(d, which, j0 + k, ex);
      ex = pairAdd(ex, v[k]);
    }
    if (tile == nTiles - 1 && threadIdx.x == 0) scanTotals(d, which, pairAdd(tileBase, total));
    __syncthreads();
  }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(Dev d, int which) { b_scan_final(d, which, blockIdx.x, gridDim.x); }

// ---- shuffled multi-sends: optimistic draw indices, checked; re-derived serially when a rejection shifted them ----
__device__ __forceinline__ void b_shuffle_check(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  const int per = d.descCap / ARENA_STRIPES;
  const int stripe = VB & (ARENA_STRIPES - 1);
  int cnt = d.ctl->descCnt[stripe];
  if (cnt > per) cnt = per;
  const int sub = (VB >> 6) * blockDim.x + threadIdx.x;
  const int nsub = (VG >> 6) * blockDim.x;
  for (int j = sub; j < cnt; j += nsub) shuffleCheck(d, stripe * per + j);
}
__global__ void k_shuffle_check(Dev d) { b_shuffle_check(d, blockIdx.x, gridDim.x); }
__device__ __forceinline__ void b_shuffle_serial(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  shuffleSerial(d);
}
__global__ void k_shuffle_serial(Dev d) { b_shuffle_serial(d, blockIdx.x, gridDim.x); }

// ---- node-sharded runs: the two exchanges of a pass (wtg_shard.cuh) --------------------------------------------
// exchange 1, publication: every item of this shard (key, prefix of slots / draws) into every shard's copy of this shard's
// list — peer stores over NVLink; a shard in error still publishes its header (which carries the error) so that the
// others stop at once
__device__ __forceinline__ void b_x1_publish(const Dev& d, const int VB, const int VG) {
  const int n = d.ctl->error ? -1 : d.ctl->nItems;
  for (int i = VB * blockDim.x + threadIdx.x; i <= n; i += VG * blockDim.x) xPublishItem(d, i);
  if (VB == 0 && threadIdx.x == 0) xPublishHeader(d);
}
__global__ void __launch_bounds__(256) k_x1_publish(Dev d) { b_x1_publish(d, blockIdx.x, gridDim.x); }
// signal the end of this shard's publication of `phase` to every shard, then wait for all of theirs (one block:
// spinning must not occupy the machine — shards may share a GPU in the tests)
__device__ __forceinline__ void b_x_sync(const Dev& d, int phase, const int VB, const int VG) {
  if (threadIdx.x == 0) xSignal(d, phase);  // stream order: the publishing kernel has completed
  __syncthreads();
  if (threadIdx.x < d.G && threadIdx.x != d.rank) xWaitOne(d, phase, threadIdx.x);
}
__global__ void k_x_sync(Dev d, int phase) { b_x_sync(d, phase, blockIdx.x, gridDim.x); }
// exchange 1, evaluation: global totals, and for every local item that created something the creation indices / draws
// of the other shards that come first
__device__ __forceinline__ void b_x1_offsets(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  const int n = d.ctl->nItems;
  for (int i = VB * blockDim.x + threadIdx.x; i < n; i += VG * blockDim.x) xOffsets(d, i);
}
__global__ void __launch_bounds__(256) k_x1_offsets(Dev d) { b_x1_offsets(d, blockIdx.x, gridDim.x); }
__device__ __forceinline__ void b_x1_totals(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  xTotals(d);
}
__global__ void k_x1_totals(Dev d) { b_x1_totals(d, blockIdx.x, gridDim.x); }
// exchange 2, after the wait: pooled payloads that arrived in the staging area move into pool slabs (warp per envelope)
__device__ __forceinline__ void b_x2_ingest(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  const int G = d.ctl->totalSlots;
  CoopWarp c;
  const int lane = threadIdx.x & 31;
  const int gw = (VB * blockDim.x + threadIdx.x) >> 5, nw = (VG * blockDim.x) >> 5;
  for (int g0 = gw * 32; g0 < G; g0 += nw * 32) {
    int g = g0 + lane;
    unsigned m = __ballot_sync(0xffffffffu, g < G && xNeedsIngest(d, g));
    while (m) {
      int src = __ffs(m) - 1;
      m &= m - 1;
      xIngest(d, c, g0 + src);
    }
  }
}
__global__ void __launch_bounds__(256) k_x2_ingest(Dev d) { b_x2_ingest(d, blockIdx.x, gridDim.x); }

// ---- emit ------------------------------------------------------------------------------------
__device__ __forceinline__ void b_emit(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  // conditional-task inserts (one per node) ...
  for (int i = VB * blockDim.x + threadIdx.x; i < d.nLoc; i += VG * blockDim.x) emitCond(d, d.n0 + i);
  // ... then the handlers' descriptors, blocks assigned to arena stripes
  const int per = d.descCap / ARENA_STRIPES;
  const int stripe = VB & (ARENA_STRIPES - 1);
  int cnt = d.ctl->descCnt[stripe];
  if (cnt > per) cnt = per;
  const int sub = (VB >> 6) * blockDim.x + threadIdx.x;
  const int nsub = (VG >> 6) * blockDim.x;
  for (int j = sub; j < cnt; j += nsub) emitDesc(d, stripe * per + j);
}
__global__ void k_emit(Dev d) { b_emit(d, blockIdx.x, gridDim.x); }

// sendAll descriptors: one warp each (arrival per destination, stable counting sort by arrival)
__device__ __forceinline__ void b_emit_all(const Dev& d, const int VB, const int VG) {
  __shared__ int hist[4][ALL_HIST];
  if (d.ctl->error) return;
  int cnt = d.ctl->allCnt;
  if (cnt > d.allCap) cnt = d.allCap;
  const int warp = threadIdx.x >> 5;
  if (warp >= 4) return;  // four histograms per block
  const int gw = VB * 4 + warp, nw = VG * 4;
  CoopWarp c;
  for (int j = gw; j < cnt; j += nw) emitAll(d, c, d.allList[j], d.allTmp + (size_t)gw * d.N, hist[warp]);
}
__global__ void __launch_bounds__(128) k_emit_all(Dev d) { b_emit_all(d, blockIdx.x, gridDim.x); }
// node-sharded sendAll (CasperIMD): the descriptors are published to every shard before the envelope exchange ...
__device__ __forceinline__ void b_x_all_publish(const Dev& d, const int VB, const int VG) {
  int cnt = d.ctl->error ? 0 : d.ctl->allCnt;
  if (cnt > d.xAllCap) cnt = 0;  // xPublishAllCount reports the overflow
  for (int j = VB * blockDim.x + threadIdx.x; j < cnt; j += VG * blockDim.x) xPublishAll(d, j);
  if (VB == 0 && threadIdx.x == 0) xPublishAllCount(d);
}
__global__ void __launch_bounds__(128) k_x_all_publish(Dev d) { b_x_all_publish(d, blockIdx.x, gridDim.x); }
// ... and after it every shard builds every sendAll of the pass: the same sorted record in the same slot (replicated records)
__device__ __forceinline__ void b_x_all_build(const Dev& d, const int VB, const int VG) {
  __shared__ int hist[4][ALL_HIST];
  if (d.ctl->error) return;
  const int cnt = xAllTotal(d);
  const int warp = threadIdx.x >> 5;
  if (warp >= 4) return;
  const int gw = VB * 4 + warp, nw = VG * 4;
  CoopWarp c;
  for (int k = gw; k < cnt; k += nw) xBuildAll(d, c, k, d.allTmp + (size_t)gw * d.N, hist[warp]);
}
__global__ void __launch_bounds__(128) k_x_all_build(Dev d) { b_x_all_build(d, blockIdx.x, gridDim.x); }

// ---- multisplit: stable distribution of the new envelopes into the time ring -----------------
// A block of Dev.msWarps warps handles a chunk of msWarps * MS_SUB consecutive envelopes (creation order); warp w owns the w-th
// sub-chunk of MS_SUB envelopes.  count: per-warp histograms over the ring bins in shared memory -> one row of totals per
// chunk.  scan: running offset per bin over the chunks.  scatter: the histograms again, turned into each warp's first
// position per bin (chunk offset + the lower warps' counts); then MS_SUB / 32 rounds of match-any ranked placement per warp.
constexpr int MS_SUB = 256;  // envelopes per warp
constexpr int MS_ROUNDS = MS_SUB / 32;
// warps per block = Dev.msWarps (8 unless the ring is so long that 8 histograms do not fit in shared memory); chunk = msWarps * MS_SUB
__device__ __forceinline__ void msZero(int* hist, int ring, int warps) {
  for (int i = threadIdx.x * 4; i < warps * ring; i += blockDim.x * 4) *reinterpret_cast<int4*>(hist + i) = make_int4(0, 0, 0, 0);
}
// this warp's targets of the chunk (kept in registers) and their histogram
__device__ __forceinline__ void msWarpHistogram(const Dev& d, int* hw, int g0, int G, int tick, int lane, int (&tg)[MS_ROUNDS]) {
#pragma unroll
  for (int r = 0; r < MS_ROUNDS; ++r) {  // all targets are in flight before the first one is used
    int g = g0 + r * 32 + lane;
    tg[r] = g < G ? d.newTarget[g] : -1;
  }
#pragma unroll
  for (int r = 0; r < MS_ROUNDS; ++r)
    if (tg[r] >= 0) atomicAdd(&hw[tg[r] - tick], 1);
}
__device__ __forceinline__ void b_ms_count(const Dev& d, const int VB, const int VG) {
  extern __shared__ int msHist[];  // [msWarps][ring]
  if (d.ctl->error) return;
  const int G = d.ctl->totalSlots, tick = d.ctl->tick, ring = d.ring;
  const int W = d.msWarps, CH = W * MS_SUB;
  const int nChunks = (G + CH - 1) / CH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int ch = VB; ch < nChunks; ch += VG) {
    msZero(msHist, ring, W);
    __syncthreads();
    int tg[MS_ROUNDS];
    msWarpHistogram(d, msHist + warp * ring, ch * CH + warp * MS_SUB, G, tick, lane, tg);
    __syncthreads();
    int* row = d.msCount + (size_t)ch * ring;
    for (int b = threadIdx.x; b < ring; b += blockDim.x) {
      int tot = 0;
      for (int w = 0; w < W; ++w) tot += msHist[w * ring + b];
      row[b] = tot;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_ms_count(Dev d) { b_ms_count(d, blockIdx.x, gridDim.x); }
// per ring bin: running offset over chunks, starting at the bucket's current fill
__device__ __forceinline__ void b_ms_scan(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  int G = d.ctl->totalSlots, tick = d.ctl->tick, ring = d.ring;
  const int CH = d.msWarps * MS_SUB;
  int nChunks = (G + CH - 1) / CH;
  if (nChunks == 0) return;
  for (int b = VB * blockDim.x + threadIdx.x; b < ring; b += VG * blockDim.x) {
    int slot = (tick + b) & (ring - 1);
    int run = d.bucketCount[slot];
    for (int ch0 = 0; ch0 < nChunks; ch0 += 8) {
      int v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = ch0 + k < nChunks ? d.msCount[(size_t)(ch0 + k) * ring + b] : 0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (ch0 + k < nChunks) {
          d.msCount[(size_t)(ch0 + k) * ring + b] = run;
          run += v[k];
        }
    }
    if (run > d.bcap) {
      setError(d, ERR_BUCKET_OVERFLOW, tick + b);
      run = d.bcap;
    }
    d.bucketCount[slot] = run;
  }
}
__global__ void k_ms_scan(Dev d) { b_ms_scan(d, blockIdx.x, gridDim.x); }
__device__ __forceinline__ void b_ms_scatter(const Dev& d, const int VB, const int VG) {
  extern __shared__ int msHist[];
  if (d.ctl->error) return;
  const int G = d.ctl->totalSlots, tick = d.ctl->tick, ring = d.ring;
  const int W = d.msWarps, CH = W * MS_SUB;
  const int nChunks = (G + CH - 1) / CH;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* base = msHist + warp * ring;
  for (int ch = VB; ch < nChunks; ch += VG) {
    msZero(msHist, ring, W);
    __syncthreads();
    const int g0 = ch * CH + warp * MS_SUB;
    int tg[MS_ROUNDS];
    msWarpHistogram(d, base, g0, G, tick, lane, tg);
    __syncthreads();
    const int* row = d.msCount + (size_t)ch * ring;  // first position of this chunk in every bin (after the scan)
    for (int b = threadIdx.x; b < ring; b += blockDim.x) {
      int run = row[b];
      for (int w = 0; w < W; ++w) {
        int c = msHist[w * ring + b];
        msHist[w * ring + b] = run;
        run += c;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < MS_ROUNDS; ++r) {
      int g = g0 + r * 32 + lane;
      int t = tg[r];
      int bin = t >= 0 ? t - tick : -1 - lane;  // unique negative key for lanes without an envelope
      unsigned peers = __match_any_sync(0xffffffffu, bin);
      int rank = __popc(peers & ((1u << lane) - 1u));
      int leader = __ffs(peers) - 1;
      int b0 = 0;
      if (t >= 0 && lane == leader) {
        b0 = base[bin];
        base[bin] = b0 + __popc(peers);
      }
      b0 = __shfl_sync(0xffffffffu, b0, leader);
      if (t >= 0) {
        int pos = b0 + rank;
        if (pos < d.bcap) {
          const int4* src = reinterpret_cast<const int4*>(d.newEv + g);
          int4* dst = reinterpret_cast<int4*>(d.buckets + (size_t)(t & (ring - 1)) * (size_t)d.bcap + pos);
          int4 x = src[0], y = src[1];
          dst[0] = x;
          dst[1] = y;
          if (d.G > 1) d.bucketKey[(size_t)(t & (ring - 1)) * (size_t)d.bcap + pos] = orderKey((unsigned)d.ctl->xseq, (unsigned)g);
        }
        if (d.G > 1) d.newTarget[g] = -1;  // the array is indexed by the global creation index: clean for the next pass
      }
      __syncwarp();
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_ms_scatter(Dev d) { b_ms_scatter(d, blockIdx.x, gridDim.x); }

__device__ __forceinline__ void b_free(const Dev& d, const int VB, const int VG) {
  if (d.ctl->error) return;
  const int per = d.freeCap / ARENA_STRIPES;
  const int stripe = VB & (ARENA_STRIPES - 1);
  int cnt = d.ctl->freeCnt[stripe];
  if (cnt > per) cnt = per;
  const int sub = (VB >> 6) * blockDim.x + threadIdx.x;
  const int nsub = (VG >> 6) * blockDim.x;
  for (int j = sub; j < cnt; j += nsub) freeApply(d, stripe * per + j);
}
__global__ void k_free(Dev d) { b_free(d, blockIdx.x, gridDim.x); }

#if defined(WTG_PERSISTENT_WINDOW)  // experiment kept for reference (profiles/README.md, round 2): slower than the graph at the metric size
// ---- one cooperative kernel per runMs window ------------------------------------------------------------------
// The pipeline's kernels are tiny at most ticks (a launch boundary costs more than the work), so a whole window runs as
// ONE cooperative launch: every stage is a grid-stride loop over the virtual blocks of the stage's stand-alone launch
// configuration, separated by grid-wide barriers instead of kernel boundaries.  Same stage bodies, same order.
struct RunCfg {
  int pre0, count1, post2;  // passes: mode 0 (events at the current time), mode 1 (clock ticks), mode 2 (end of window)
  int sms;
};
#define WTG_STAGE(VGRID, CALL)                                        \
  do {                                                                \
    const int vg_ = (VGRID);                                          \
    for (int vb = blockIdx.x; vb < vg_; vb += gridDim.x) { CALL; }    \
  } while (0)
__global__ void __launch_bounds__(256) k_run(Dev d, RunCfg rc) {
  cg::grid_group grid = cg::this_grid();
  const int total = rc.pre0 + rc.count1 + rc.post2;
  const int wide = rc.sms * 8;
  const int striped = ARENA_STRIPES * 19;
  auto modeOf = [&](int ps) { return ps < rc.pre0 ? 0 : ps < rc.pre0 + rc.count1 ? 1 : 2; };
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) d.ctl->stop = 0;
    __syncthreads();
    b_begin(d, modeOf(0), 0, 1);
  }
  grid.sync();
  for (int ps = 0; ps < total; ++ps) {
    const int mode = modeOf(ps);
    if (d.proto == PROTO_GSF || d.proto == PROTO_HANDEL) {
      WTG_STAGE((d.nLoc + 255) / 256, b_cond_mark(d, vb, vg_));
      grid.sync();
      WTG_STAGE(striped, b_cond_nodes<0>(d, vb, vg_));
      grid.sync();
      WTG_STAGE(striped, b_cond_score(d, vb, vg_));
      grid.sync();
      WTG_STAGE(striped, b_cond_nodes<1>(d, vb, vg_));
      grid.sync();
      if (d.proto == PROTO_HANDEL) {
        WTG_STAGE(wide, b_scan_partial(d, 2, vb, vg_));
        grid.sync();
        WTG_STAGE(wide, b_scan_final(d, 2, vb, vg_));
        grid.sync();
        WTG_STAGE(rc.sms, b_hpick_check(d, vb, vg_));
        grid.sync();
        WTG_STAGE(rc.sms, b_hpick_apply(d, vb, vg_));
        grid.sync();
      }
    }
    if (mode != 2) {
      WTG_STAGE(wide, b_dispatch_count(d, vb, vg_));
      grid.sync();
      WTG_STAGE(wide, b_scan_partial(d, 0, vb, vg_));
      grid.sync();
      WTG_STAGE(wide, b_scan_final(d, 0, vb, vg_));
      grid.sync();
      WTG_STAGE(wide, b_dispatch_scatter(d, vb, vg_));
      grid.sync();
      WTG_STAGE((d.nLoc + 255) / 256, b_node_msgs(d, vb, vg_));
      grid.sync();
      WTG_STAGE(striped, b_node_tasks(d, vb, vg_));
      grid.sync();
    }
    WTG_STAGE(wide, b_scan_partial(d, 1, vb, vg_));
    grid.sync();
    WTG_STAGE(wide, b_scan_final(d, 1, vb, vg_));
    grid.sync();
    if (d.G > 1) {  // node-sharded: exchange 1
      WTG_STAGE(rc.sms * 2, b_x1_publish(d, vb, vg_));
      grid.sync();
      if (blockIdx.x == 0) b_x_sync(d, 0, 0, 1);
      grid.sync();
      if (blockIdx.x == 0 && threadIdx.x == 0) b_x1_totals(d, 0, 1);
      WTG_STAGE(rc.sms * 2, b_x1_offsets(d, vb, vg_));
      grid.sync();
    }
    if (d.shufCap > 0) {
      WTG_STAGE(ARENA_STRIPES * 4, b_shuffle_check(d, vb, vg_));
      grid.sync();
      if (blockIdx.x == 0 && threadIdx.x == 0) b_shuffle_serial(d, 0, 1);
      grid.sync();
    }
    WTG_STAGE(ARENA_STRIPES * 16, b_emit(d, vb, vg_));
    if (d.allCap > 0) {
      grid.sync();
      WTG_STAGE(d.allWarps / 4, b_emit_all(d, vb, vg_));
    }
    grid.sync();
    if (d.G > 1) {  // exchange 2
      if (blockIdx.x == 0) b_x_sync(d, 1, 0, 1);
      grid.sync();
      WTG_STAGE(rc.sms * 4, b_x2_ingest(d, vb, vg_));
      grid.sync();
    }
    WTG_STAGE(rc.sms * 2, b_ms_count(d, vb, vg_));
    if (d.proto == PROTO_GSF || d.proto == PROTO_HANDEL) WTG_STAGE(ARENA_STRIPES * 2, b_free(d, vb, vg_));
    grid.sync();
    WTG_STAGE((d.ring + 255) / 256, b_ms_scan(d, vb, vg_));
    grid.sync();
    WTG_STAGE(rc.sms * 2, b_ms_scatter(d, vb, vg_));
    grid.sync();
    // end of this pass and beginning of the next, by one warp
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      if (threadIdx.x == 0) {
        tickEnd(d, mode);
        d.ctl->passesDone += 1;
        if (d.ctl->error || (d.ffwd && d.ctl->idle)) d.ctl->stop = 1;
      }
      __syncwarp();
      if (ps + 1 < total && !d.ctl->stop) b_begin(d, modeOf(ps + 1), 0, 1);
    }
    grid.sync();
    if (d.ctl->stop) break;
  }
}

#endif

// ---- init kernels ---------------------------------------------------------------------------
__global__ void k_gsf_init_nodes(Dev d) {
  int n = d.n0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (n < d.n0 + d.nLoc) gsfInitNodeBody(d, n);
}
__global__ void k_rng_candidates(Dev d, u64 s0, u64 count, u64 chunk, int maxBound, u64* out, int* outCount, int cap) {
  u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  u64 start = t * chunk;
  if (start >= count) return;
  u64 len = start + chunk <= count ? chunk : count - start;
  rngCandidateChunk(d, s0, start, len, maxBound, out, outCount, cap);
}
template <class PeerT>
__global__ void k_gsf_shuffle(Dev d, int l, u64 s0, const int* liveRank, const u64* rejOrd, int nRej) {
  int n = d.n0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (n < d.n0 + d.nLoc) gsfShuffleLevel<PeerT>(d, n, l, s0, liveRank, rejOrd, nRej);
}

// ------------------------------------------------------------------------------------------------
class CudaBackend : public Backend {
 public:
  cudaStream_t st = nullptr;
  int devId = 0;  // every entry point binds it: callers may drive different networks from different host threads
  void bind() const { cudaSetDevice(devId); }
  int sms = 148;
  int smemOptin = 0;
  cudaGraphExec_t tickGraph = nullptr;
  const void* graphFor = nullptr;
  bool useGraph = true;
  long long launches = 0;      // kernels enqueued (graph replays counted kernel by kernel)
  long long graphKernels = 0;
  cudaEvent_t tm0 = nullptr, tm1 = nullptr;
  // per-kernel profiling
  static constexpr int NK = 16;
  bool profiling = false;
  std::vector<cudaEvent_t> evPool;
  std::vector<int> evKernel;  // kernel id of each event pair
  size_t evUsed = 0;
  double profMs[NK] = {};
  long long profCnt[NK] = {};
  const char* profNames[NK] = {"k_begin", "k_cond_scan", "k_dispatch_count", "k_scan_partial", "k_exchange", "k_scan_final",
                               "k_dispatch_scatter", "k_node", "k_emit", "k_ms_count", "k_ms_scan", "k_ms_scatter", "k_free",
                               "k_end", "k_cond_score", "k_cond_select"};

  explicit CudaBackend(int requested = -1) {
    int dev = 0;
    const char* e = std::getenv("LOCAL_RANK");
    int cnt = 0;
    CUDA_OK(cudaGetDeviceCount(&cnt));
    if (cnt == 0) throw std::runtime_error("no CUDA device");
    if (e) dev = std::atoi(e) % cnt;
    const char* e2 = std::getenv("WTG_DEVICE");
    if (e2) dev = std::atoi(e2) % cnt;
    if (requested >= 0) {  // wtg_create_on: the caller places this network itself
      if (requested >= cnt) throw std::invalid_argument("CUDA device " + std::to_string(requested) + " does not exist");
      dev = requested;
    }
    CUDA_OK(cudaSetDevice(dev));
    devId = dev;
    cudaDeviceProp p;
    CUDA_OK(cudaGetDeviceProperties(&p, dev));
    sms = p.multiProcessorCount;
    CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CUDA_OK(cudaDeviceGetAttribute(&smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    CUDA_OK(cudaFuncSetAttribute(k_ms_count, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
    CUDA_OK(cudaFuncSetAttribute(k_ms_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
    if (NODE_TMA_SMEM > 0) CUDA_OK(cudaFuncSetAttribute(k_node_tasks, cudaFuncAttributeMaxDynamicSharedMemorySize, NODE_TMA_SMEM));
    const char* g = std::getenv("WTG_NO_GRAPH");
    if (g && g[0] == '1') useGraph = false;
#if defined(WTG_PERSISTENT_WINDOW)
    const char* np = std::getenv("WTG_NO_PERSIST");
    if (np && np[0] == '1') usePersistent = false;
#endif
  }
  ~CudaBackend() override {
    for (void* q : ipcOpened) cudaIpcCloseMemHandle(q);
    if (tickGraph) cudaGraphExecDestroy(tickGraph);
    if (st) cudaStreamDestroy(st);
  }
  void* alloc(size_t bytes) override {
    bind();
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) throw std::runtime_error("cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    CUDA_OK(cudaMemsetAsync(p, 0, bytes, st));
    return p;
  }
  void release(void* p) override { cudaFree(p); }
  int deviceId() const override { return devId; }
  // exchange region of a shard: plain cudaMalloc memory (CUDA IPC cannot export pool allocations)
  void exportShared(void* p, unsigned char* handle) override {
    bind();
    std::memset(handle, 0, 128);
    cudaIpcMemHandle_t ih;
    CUDA_OK(cudaIpcGetMemHandle(&ih, p));
    static_assert(sizeof(ih) == 64, "cudaIpcMemHandle_t");
    std::memcpy(handle, &ih, 64);
    long long pid = (long long)getpid();
    std::memcpy(handle + 64, &pid, 8);
    std::memcpy(handle + 72, &p, sizeof(p));
    std::memcpy(handle + 80, &devId, 4);
  }
  std::vector<void*> ipcOpened;
  void* importShared(const unsigned char* handle) override {
    bind();
    long long pid;
    void* p;
    int dev;
    std::memcpy(&pid, handle + 64, 8);
    std::memcpy(&p, handle + 72, sizeof(p));
    std::memcpy(&dev, handle + 80, 4);
    if (pid == (long long)getpid()) {  // a shard of this process: same address space
#if defined(WTG_PERSISTENT_WINDOW)
      if (dev == devId) sharesDevice = true;
#endif
      if (dev != devId) {
        int can = 0;
        CUDA_OK(cudaDeviceCanAccessPeer(&can, devId, dev));
        if (!can) throw std::runtime_error("GPU " + std::to_string(devId) + " cannot access GPU " + std::to_string(dev) + " (peer access)");
        cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_OK(e);
        cudaGetLastError();
      }
      return p;
    }
    cudaIpcMemHandle_t ih;
    std::memcpy(&ih, handle, 64);
    void* q = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&q, ih, cudaIpcMemLazyEnablePeerAccess));
    ipcOpened.push_back(q);
    return q;
  }
  void upload(void* dst, const void* src, size_t bytes) override {
    bind();
    CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
  void download(void* dst, const void* src, size_t bytes) override {
    bind();
    CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
  }
  void sync() override {
    bind();
    CUDA_OK(cudaStreamSynchronize(st));
    drainProfile();
  }
  void timerStart() override {
    bind();
    if (!tm0) {
      CUDA_OK(cudaEventCreate(&tm0));
      CUDA_OK(cudaEventCreate(&tm1));
    }
    CUDA_OK(cudaEventRecord(tm0, st));
  }
  double timerStopMs() override {
    bind();
    CUDA_OK(cudaEventRecord(tm1, st));
    CUDA_OK(cudaEventSynchronize(tm1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, tm0, tm1));
    return (double)ms;
  }
  void profileEnable(bool on) override {
    sync();
    profiling = on;
    if (on) {
      for (int i = 0; i < NK; ++i) {
        profMs[i] = 0;
        profCnt[i] = 0;
      }
    }
  }
  int profileRead(double* ms, long long* cnt, const char** names, int cap) override {
    sync();
    int k = 0;
    for (int i = 0; i < NK && k < cap; ++i, ++k) {
      ms[k] = profMs[i];
      cnt[k] = profCnt[i];
      names[k] = profNames[i];
    }
    return k;
  }
  void drainProfile() {
    for (size_t i = 0; i < evUsed; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, evPool[2 * i], evPool[2 * i + 1]);
      profMs[evKernel[i]] += ms;
      profCnt[evKernel[i]] += 1;
    }
    evUsed = 0;
  }
  void profBegin(int kid) {
    if (!profiling) return;
    if (evUsed * 2 + 2 > evPool.size()) {
      if (evUsed >= 8192) {  // bound the pool: drain what has completed so far
        CUDA_OK(cudaStreamSynchronize(st));
        drainProfile();
      } else {
        cudaEvent_t a, b;
        CUDA_OK(cudaEventCreate(&a));
        CUDA_OK(cudaEventCreate(&b));
        evPool.push_back(a);
        evPool.push_back(b);
        evKernel.push_back(0);
      }
    }
    evKernel[evUsed] = kid;
    CUDA_OK(cudaEventRecord(evPool[2 * evUsed], st));
  }
  void profEnd() {
    if (!profiling) return;
    CUDA_OK(cudaEventRecord(evPool[2 * evUsed + 1], st));
    ++evUsed;
  }

  void enqueueTick(const Dev& d, int mode) {
    const int wide = sms * 8;
    const size_t msSmem = (size_t)d.msWarps * d.ring * sizeof(int);
    // mode 3: the host prepared the control block and the descriptors of sends it injects at the current time
    // (Engine::inject); only the emission half of the pipeline runs
    if (mode != 3) {
      profBegin(0);
      k_begin<<<1, 32, 0, st>>>(d, mode);
      profEnd();
    }
    if (d.proto == PROTO_HANDEL && mode != 3) {
      profBegin(1);
      k_cond_mark<<<(d.nLoc + 255) / 256, 256, 0, st>>>(d);
      k_cond_nodes<0><<<ARENA_STRIPES * 19, 256, 0, st>>>(d);
      profEnd();
      profBegin(14);
      k_cond_score<<<ARENA_STRIPES * 19, 256, 0, st>>>(d);
      profEnd();
      profBegin(15);
      k_cond_nodes<1><<<ARENA_STRIPES * 19, 256, 0, st>>>(d);
      profEnd();
      profBegin(3);
      k_scan_partial<<<wide, SCAN_THREADS, 0, st>>>(d, 2);
      profEnd();
      profBegin(5);
      k_scan_final<<<wide, SCAN_THREADS, 0, st>>>(d, 2);
      profEnd();
      profBegin(15);
      k_hpick_check<<<sms, 256, 0, st>>>(d);
      k_hpick_apply<<<sms, 256, 0, st>>>(d);
      profEnd();
      launches += 7;
    }
    if (d.proto == PROTO_GSF && mode != 3) {
      size_t smem8 = (size_t)8 * (size_t)(d.qcap / 32) * sizeof(uint32_t);
      profBegin(1);
      k_cond_mark<<<(d.nLoc + 255) / 256, 256, 0, st>>>(d);
      k_cond_nodes<0><<<ARENA_STRIPES * 19, 256, 0, st>>>(d);
      profEnd();
      profBegin(14);
      k_cond_score<<<ARENA_STRIPES * 19, 256, 0, st>>>(d);
      profEnd();
      profBegin(15);
      k_cond_nodes<1><<<ARENA_STRIPES * 19, 256, smem8, st>>>(d);
      profEnd();
      launches += 4;
    }
    if (mode != 2 && mode != 3) {
      profBegin(2);
    k_dispatch_count<<<wide, 256, 0, st>>>(d);
    profEnd();
      profBegin(3);
    k_scan_partial<<<wide, SCAN_THREADS, 0, st>>>(d, 0);
    profEnd();
      profBegin(5);
    k_scan_final<<<wide, SCAN_THREADS, 0, st>>>(d, 0);
    profEnd();
      profBegin(6);
    k_dispatch_scatter<<<wide, 256, 0, st>>>(d);
    profEnd();
      profBegin(7);
    k_node_msgs<<<(d.nLoc + 255) / 256, 256, 0, st>>>(d);
    k_node_tasks<<<ARENA_STRIPES * 19, 256, NODE_TMA_SMEM, st>>>(d);
    if (d.proto == PROTO_CASPER) {
      k_casper_fixups<<<1, 32, 0, st>>>(d);
      launches += 1;
    }
    profEnd();
      launches += 6;
    }
    profBegin(3);
    k_scan_partial<<<wide, SCAN_THREADS, 0, st>>>(d, 1);
    profEnd();
    profBegin(5);
    k_scan_final<<<wide, SCAN_THREADS, 0, st>>>(d, 1);
    profEnd();
    if (d.G > 1) {  // node-sharded: exchange 1 (items -> creation / draw offsets over all shards)
      profBegin(4);
      k_x1_publish<<<sms * 2, 256, 0, st>>>(d);
      k_x_sync<<<1, 32, 0, st>>>(d, 0);
      k_x1_totals<<<1, 1, 0, st>>>(d);
      k_x1_offsets<<<sms * 2, 256, 0, st>>>(d);
      profEnd();
      launches += 4;
    }
    profBegin(8);
    if (d.shufCap > 0) {
      k_shuffle_check<<<ARENA_STRIPES * 4, 256, 0, st>>>(d);
      k_shuffle_serial<<<1, 1, 0, st>>>(d);
      launches += 2;
    }
    k_emit<<<ARENA_STRIPES * 16, 256, 0, st>>>(d);
    if (d.allCap > 0) {
      if (d.G > 1)
        k_x_all_publish<<<8, 128, 0, st>>>(d);
      else
        k_emit_all<<<d.allWarps / 4, 128, 0, st>>>(d);
      launches += 1;
    }
    profEnd();
    if (d.G > 1) {  // exchange 2: the envelopes were stored into their destination shards' arrays by k_emit
      profBegin(4);
      k_x_sync<<<1, 32, 0, st>>>(d, 1);
      k_x2_ingest<<<sms * 4, 256, 0, st>>>(d);
      launches += 2;
      if (d.allCap > 0) {
        k_x_all_build<<<d.allWarps / 4, 128, 0, st>>>(d);
        launches += 1;
      }
      profEnd();
    }
    profBegin(9);
    k_ms_count<<<sms * 2, d.msWarps * 32, msSmem, st>>>(d);
    profEnd();
    profBegin(10);
    k_ms_scan<<<(d.ring + 127) / 128, 128, 0, st>>>(d);
    profEnd();
    profBegin(11);
    k_ms_scatter<<<sms * 2, d.msWarps * 32, msSmem, st>>>(d);
    profEnd();
    const bool pooled = d.proto == PROTO_GSF || d.proto == PROTO_HANDEL;  // only these protocols hold pooled payloads
    if (pooled) {
      profBegin(12);
      k_free<<<ARENA_STRIPES * 2, 256, 0, st>>>(d);
      profEnd();
    }
    profBegin(13);
    k_end<<<1, 1, 0, st>>>(d, mode);
    profEnd();
    launches += pooled ? 9 : 8;
  }
#if defined(WTG_PERSISTENT_WINDOW)
  // ---- one cooperative launch per window (k_run) ----
  bool usePersistent = true;
  bool sharesDevice = false;   // another shard of this process lives on this GPU: two cooperative grids that wait for each
                               // other cannot be co-resident, so such shards keep the kernel-per-stage pipeline
  int runGrid = 0;
  size_t runSmem = 0;
  const void* runFor = nullptr;
  long long windows = 0;
  bool persistentOk(const Dev& d) {
    if (!usePersistent || profiling || sharesDevice) return false;
    size_t need = std::max((size_t)8 * d.ring * sizeof(int), (size_t)8 * (size_t)(d.qcap / 32 + 1) * sizeof(uint32_t));
    if (need > 200 * 1024) return false;
    if (runFor != (const void*)d.ctl || runSmem != need) {
      CUDA_OK(cudaFuncSetAttribute(k_run, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
      int per = 0;
      CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_run, 256, need));
      if (per < 1) return false;
      const char* cap = std::getenv("WTG_RUN_BLOCKS_PER_SM");
      if (cap && std::atoi(cap) > 0) per = std::min(per, std::atoi(cap));
      runGrid = per * sms;
      runSmem = need;
      runFor = (const void*)d.ctl;
    }
    return true;
  }
  void runWindow(const Dev& d, int pre0, int count1, int post2) override {
    bind();
    if (pre0 + count1 + post2 <= 0) return;
    if (!persistentOk(d)) {
      if (pre0) tick(d, 0);
      if (count1) ticks(d, count1);
      if (post2) tick(d, 2);
      return;
    }
    RunCfg rc{pre0, count1, post2, sms};
    Dev dd = d;
    void* args[] = {(void*)&dd, (void*)&rc};
    CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_run, dim3(runGrid), dim3(256), args, runSmem, st));
    launches += 1;
    windows += 1;
  }
#endif
  // dynamic shared memory of the multisplit kernels: the attribute is per device and shared by every engine on it (several
  // engines may be driven from concurrent host threads), so it is set once, to the device's opt-in maximum
  void configure(const Dev& d) {
    const size_t msSmem = (size_t)d.msWarps * d.ring * sizeof(int);
    if (msSmem > (size_t)smemOptin) throw std::runtime_error("time ring too large for the multisplit's shared-memory histogram");
  }
  void tick(const Dev& d, int mode) override {
    bind();
    configure(d);
    enqueueTick(d, mode);
    CUDA_OK(cudaGetLastError());
  }
  void ticks(const Dev& d, int count) override {
    bind();
    configure(d);
    if (!useGraph || profiling || count < 4) {
      for (int i = 0; i < count; ++i) enqueueTick(d, 1);
      CUDA_OK(cudaGetLastError());
      return;
    }
    if (!tickGraph || graphFor != (const void*)d.ctl) {
      if (tickGraph) cudaGraphExecDestroy(tickGraph);
      cudaGraph_t g;
      long long before = launches;
      CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      enqueueTick(d, 1);
      CUDA_OK(cudaStreamEndCapture(st, &g));
      graphKernels = launches - before;  // kernels per replay
      launches = before;
      CUDA_OK(cudaGraphInstantiate(&tickGraph, g, 0));
      cudaGraphDestroy(g);
      graphFor = (const void*)d.ctl;
    }
    for (int i = 0; i < count; ++i) CUDA_OK(cudaGraphLaunch(tickGraph, st));
    launches += graphKernels * count;
  }
  void gsfInitNodes(const Dev& d) override {
    bind();
    k_gsf_init_nodes<<<(d.nLoc + 255) / 256, 256, 0, st>>>(d);
    CUDA_OK(cudaGetLastError());
  }
  void rngCandidates(const Dev& d, unsigned long long s0, unsigned long long count, int maxBound,
                     std::vector<unsigned long long>& out) override {
    bind();
    const u64 chunk = 16384;
    u64 threads = (count + chunk - 1) / chunk;
    int cap = (int)std::min<u64>((u64)1 << 26, count / 1024 + (1 << 16));
    u64* dOut = nullptr;
    int* dCnt = nullptr;
    CUDA_OK(cudaMalloc(&dOut, (size_t)cap * sizeof(u64)));
    CUDA_OK(cudaMalloc(&dCnt, sizeof(int)));
    CUDA_OK(cudaMemsetAsync(dCnt, 0, sizeof(int), st));
    u64 blocks = (threads + 127) / 128;
    k_rng_candidates<<<(unsigned)blocks, 128, 0, st>>>(d, s0, count, chunk, maxBound, dOut, dCnt, cap);
    CUDA_OK(cudaGetLastError());
    int n = 0;
    CUDA_OK(cudaMemcpyAsync(&n, dCnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    if (n > cap) {
      cudaFree(dOut);
      cudaFree(dCnt);
      throw std::runtime_error("rng candidate list overflow");
    }
    out.resize((size_t)n);
    if (n) CUDA_OK(cudaMemcpyAsync(out.data(), dOut, (size_t)n * sizeof(u64), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(dOut);
    cudaFree(dCnt);
  }
  void gsfShufflePeers(const Dev& d, unsigned long long s0, const int* liveRank, const unsigned long long* rejOrd, int nRej) override {
    bind();
    for (int l = d.L - 1; l >= 1; --l) {
      if (d.peerBits == 16)
        k_gsf_shuffle<uint16_t><<<(d.nLoc + 127) / 128, 128, 0, st>>>(d, l, s0, liveRank, rejOrd, nRej);
      else
        k_gsf_shuffle<uint32_t><<<(d.nLoc + 127) / 128, 128, 0, st>>>(d, l, s0, liveRank, rejOrd, nRej);
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(st));
  }
};

Backend* makeBackend(int device) { return new CudaBackend(device); }
long long backendLaunches(Backend* b) { return static_cast<CudaBackend*>(b)->launches; }

}  // namespace wtg

#define WTG_API(name) wtg_##name
#include "wtg_capi.inl"
