// nlo_kernels.cu -- sm_100a kernels of the hot path.
//
// The device-resident Gauss-Newton / damped-LM loop (or a single iteration of it) of one or many
// registrations, as two kernels:
//
// gn_iteration_kernel<KIND, LOSS, ST> -- the STREAMING kernel (scans that do not fit shared memory)
//   HBM tiles (tile-interleaved SoA, one contiguous 24 KB run per 256 correspondences)
//     --cp.async.bulk (TMA 1-D, one copy per tile, L2 eviction hint), mbarrier full/empty ring-->
//   shared-memory stages
//     --> 8 warps: residual, analytic Jacobian terms, device-inlined robust loss, 28 (10) fp64
//         register accumulators per thread
//     --> recursive-halving warp reduction -> shared -> per-CTA partial (HBM/L2) + arrival counter
//     --> leader CTA: fixed-order fp64 sum of the partials, rotation to the canonical H|g,
//         [LL-format peer-memory all-reduce over NVLink when the scan is sharded across GPUs],
//         damped 6x6 LDL^T / 3x3 solve, pose update, convergence tests, lambda schedule, trace row
//     --> new state published to the other CTAs as LL words (persistent grid) or left in HBM.
//   Launch shapes (chosen in nlo_api.cu): persistent cooperative grid with the whole loop inside,
//   one CTA per registration with the whole loop inside (batched), or one launch per iteration
//   with a last-CTA-by-ticket finaliser (NCCL flavour, plain assemble calls).
//
// gn_resident_kernel<KIND, LOSS> -- the RESIDENT kernel (latency-bound registrations: every tile is
//   loaded into shared memory once per Solve).  One CTA per SM, thread-block clusters: the CTAs of a
//   cluster pre-reduce over distributed shared memory, the cluster leaders exchange LL-format
//   partials through L2 and hand the totals back over DSMEM, every CTA performs the identical step.
//
// Replaces the per-iteration loops of (paths relative to /root/reference/nonlinear_optimizer/)
//   mahalanobis_distance_minimizer/mahalanobis_distance_minimizer_analytic.cc:92-149
//   mahalanobis_distance_minimizer/mahalanobis_distance_minimizer_analytic_3dof.cc:29-99
//   reprojection_error_minimizer/reprojection_error_minimizer_analytic.cc:26-100
// and their SIMD / thread-pool twins (..._analytic_simd.cc:55-76,114-177).
#include <climits>
#include <cstdio>
#include <cstring>

#include "nlo_device.cuh"

namespace nlo {

// ------------------------------------------------------------------ PTX helpers (mbarrier, TMA)
__device__ __forceinline__ uint32_t SmemAddr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void MbarInit(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void MbarArrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(SmemAddr(bar)) : "memory");
}
__device__ __forceinline__ bool MbarTryWait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(SmemAddr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity) {
  while (!MbarTryWait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared, completion counted on an mbarrier (TMA engine, UBLKCP in SASS).
__device__ __forceinline__ void BulkLoad(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(SmemAddr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(SmemAddr(bar))
      : "memory");
}
// Same, with an explicit L2 eviction policy (createpolicy): tiles that should stay resident in the
// 126 MB L2 across iterations are loaded evict_last, the streamed remainder evict_first.
__device__ __forceinline__ void BulkLoadHint(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(SmemAddr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(SmemAddr(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t PolicyEvictLast() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t PolicyEvictFirst() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void FenceBarrierInit() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ unsigned long long GlobalTimerNs() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Watchdog of a polled wait.  %globaltimer is slow to read (~100 ns, and every polling thread of
// the GPU reading it at once makes it worse), so the spin loops count their polls and look at the
// SM's cycle counter only every 1024th time; the limits are converted from ns at a nominal 2 GHz
// (they are "never hang" limits of tens of seconds, the SM clock varying by 2x does not matter).
struct SpinWatch {
  unsigned int polls = 0;
  long long start = 0;
  // true when the wait has lasted longer than `limit_ns`
  __device__ __forceinline__ bool Expired(unsigned long long limit_ns) {
    if ((++polls & 1023u) != 0u) return false;
    const long long now = clock64();
    if (start == 0) {
      start = now;
      return false;
    }
    return static_cast<unsigned long long>(now - start) > 2ULL * limit_ns;
  }
};

template <int KIND>
struct KindTraits;
template <>
struct KindTraits<kNdt6> {
  static constexpr int kAcc = kAcc6, kStages = 3, kTrace = 36;
};
template <>
struct KindTraits<kNdt3> {
  static constexpr int kAcc = kAcc3, kStages = 3, kTrace = 17;
};
template <>
struct KindTraits<kReproj> {
  static constexpr int kAcc = kAcc6, kStages = 4, kTrace = 36;
};

// Planes of one correspondence for a kind and a storage type.
template <int KIND, typename ST>
__host__ __device__ constexpr int PlanesOf() {
  return KIND == kReproj ? kReprojPlanes : kNdtPlanes;
}

// Shared memory of the reduction / exchange / step phase of an iteration (the same for every kind).
struct ReduceArea {
  double warp_sums[16][kAcc6];    // sums of the 8 (16) consumer warps
  double gather_lanes[8][kAcc6];  // the 8 strided lanes of a cross-CTA / cross-cluster / cross-rank sum
  double total[32];               // reduced (raw, then canonical) sums
  State state;                    // CTA-local copy of the registration state
  unsigned long long seq0;        // peer exchange sequence number at kernel start
  int exchanges;                  // peer exchanges made by this kernel (thread 0's bookkeeping)
  unsigned int halves[kPeerWords];  // streaming kernel: payload halves of the published state (LL words)
  int flag;
  int fail;                       // a polled wait expired
};

// Shared memory of the resident kernel in front of its stages.
struct ResidentSmem {
  ReduceArea red;
  double cluster_part[2][kMaxCluster][kAcc6];  // rank-0 CTA of a cluster: the sums of its CTAs (by iteration parity)
  uint64_t full;  // all tiles of this CTA have landed
};

// ST = storage type of the planes in HBM and in the stages: double (parity mode, 96 B per NDT
// correspondence) or float (fp32 storage, fp64 math: 48 B; twice the stages fit the same smem).
// WG = warp groups (of 8 warps) per CTA = consecutive tiles per ring stage, see the kernel.
template <int KIND, typename ST, int WG = 1>
struct SmemLayout {
  using T = KindTraits<KIND>;
  // bytes per stage: NDT fp64 24 KB (12 planes), NDT fp32 12 KB, PnP 10 KB; two CTAs per SM share
  // the 227 KB
#ifndef NLO_NDT_STAGES
#define NLO_NDT_STAGES 4
#endif
  // (WG = 2: three 48 KB stages -- the depth the streaming loop uses anyway -- leave the L1 ~90 KB instead of ~30)
  static constexpr int kStages = (KIND == kReproj) ? 4 : (sizeof(ST) == 8 ? (WG == 2 ? 3 : NLO_NDT_STAGES) : 8);
  ST stages[kStages][WG][PlanesOf<KIND, ST>()][kTile];
  ReduceArea red;
  uint64_t full[kStages];
  uint64_t empty[kStages];
};

template <int KIND, typename ST = double, int WG = 1>
constexpr size_t SmemBytes() {
  return sizeof(SmemLayout<KIND, ST, WG>);
}

// ------------------------------------------------------------------ thread-block cluster helpers
__device__ __forceinline__ unsigned int ClusterCtaRank() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned int ClusterIdX() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned int ClusterCountX() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned int ClusterSize() {
  unsigned int r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// Hardware barrier of the cluster; release / acquire orders the distributed-shared-memory stores
// made before it against the loads made after it.
__device__ __forceinline__ void ClusterSync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Stores `v` at the address `local` has in the shared memory of CTA `target_rank` of this cluster.
__device__ __forceinline__ void StoreClusterF64(double* local, unsigned int target_rank, double v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(SmemAddr(local)), "r"(target_rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}
__device__ __forceinline__ void StoreClusterU32(int* local, unsigned int target_rank, unsigned int v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(SmemAddr(local)), "r"(target_rank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
}

// ------------------------------------------------------------------ "LL" words
// A double travels as two 8-byte words, each = (tag << 32) | 32 payload bits, so that a word is
// valid the moment its tag matches: one store, one (polled) load, no fence, no separate flag --
// one L2 round trip inside a GPU, one NVLink one-way latency between GPUs.  The two words of a
// double are adjacent (one 16-byte store / load; each half is validated on its own).
// Scope: words exchanged inside one GPU use gpu-scope relaxed accesses, served by the L2; words that
// cross NVLink use system scope (`sys`).
__device__ __forceinline__ void StoreLL(unsigned long long* dst, double v, unsigned int tag, bool sys) {
  const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
  const unsigned long long t = static_cast<unsigned long long>(tag) << 32;
  const unsigned long long lo = t | (bits & 0xffffffffULL), hi = t | (bits >> 32);
  if (sys)
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
  else
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void LoadLL(const unsigned long long* src, unsigned long long& lo,
                                       unsigned long long& hi, bool sys) {
  if (sys)
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
  else
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(src) : "memory");
}

__device__ __forceinline__ unsigned int LoadGpuU32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long LoadGpuU64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Fixed-order sum of `n_src` LL records of NACC doubles (record c at base + c * stride_words):
// thread (j, l8) polls value j of records l8, l8+8, ... -- kGatherDepth loads in flight -- and adds them in
// record order, then the 8 lanes are added in lane order.  total[0..NACC) is written by threads
// 0..NACC-1 (warp 0) after an internal __syncthreads; *fail is set when a record did not show up
// within `timeout_ns`.  Called by all threads of the CTA.  Deliberately small and NOT inlined
// (one copy): the per-iteration code of a latency-bound registration has to stay inside the 32 KB
// instruction cache of the SM, a miss per 128-byte line of cold code costs more than the math.
// Loads of one polling thread in flight.  Measured per iteration, depth 2 / 4 / 6 (B200, ndt6 | ndt3):
// 100 k points (33 records) 7.25 / 7.25 / 7.31 | 6.21 / 6.58 / 6.68 us, 200 k (74 records) 8.05 / 8.05 / 8.32
// | 6.85 / 7.14 / 7.35, 300 k 11.68 / 11.16 / 11.05 | 9.68 / 9.29 / 8.98: more polls in flight delay the
// very stores they wait for at the reference's sizes and only pay off above them.  Backing off between
// polls (__nanosleep 20 / 60 / 200 ns) was measured slower at every size (100 k ndt6: 7.96 / 8.04 / 8.30
// against 7.22 us) and is not kept: the poll that finds the record is the one that matters.
constexpr int kGatherDepth = 2;
template <int NACC>
__device__ __noinline__ void GatherLL(const unsigned long long* base, int stride_words, int n_src,
                                      unsigned int tag, bool sys, unsigned long long timeout_ns,
                                      double (*lanes)[kAcc6], double* total, int* fail) {
  const int tid = threadIdx.x;
  const int j = tid >> 3, l8 = tid & 7;
  if (j < NACC) {
    double s = 0.0;
    SpinWatch watch;
    const unsigned long long* src = base + static_cast<size_t>(l8) * stride_words + 2 * j;
    const size_t step = static_cast<size_t>(8) * stride_words;
    // kGatherDepth records of this lane travel at once; one copy of the spin loop, the registers rotate
    unsigned long long lo[kGatherDepth], hi[kGatherDepth];
#pragma unroll
    for (int u = 0; u < kGatherDepth - 1; ++u) {
      lo[u] = hi[u] = 0ULL;
      if (l8 + 8 * u < n_src) LoadLL(src + u * step, lo[u], hi[u], sys);
    }
    lo[kGatherDepth - 1] = hi[kGatherDepth - 1] = 0ULL;
#pragma unroll 1
    for (int c = l8; c < n_src; c += 8) {
      if (c + 8 * (kGatherDepth - 1) < n_src) LoadLL(src + (kGatherDepth - 1) * step, lo[kGatherDepth - 1], hi[kGatherDepth - 1], sys);
      while (static_cast<unsigned int>(lo[0] >> 32) != tag || static_cast<unsigned int>(hi[0] >> 32) != tag) {
        if (watch.Expired(timeout_ns)) {
          *fail = 1;
          lo[0] = hi[0] = static_cast<unsigned long long>(tag) << 32;
          break;
        }
        LoadLL(src, lo[0], hi[0], sys);
      }
      s += __longlong_as_double(static_cast<long long>((hi[0] << 32) | (lo[0] & 0xffffffffULL)));
      src += step;
#pragma unroll
      for (int u = 0; u + 1 < kGatherDepth; ++u) {
        lo[u] = lo[u + 1];
        hi[u] = hi[u + 1];
      }
    }
    lanes[l8][j] = s;
  }
  __syncthreads();
  if (tid < NACC) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += lanes[w][tid];
    total[tid] = s;
  }
}

// profiling aid (NLO_DEBUG_TIMES=1): CTA 0 / thread 0 stamps the phases of each iteration
// (NLO_DEBUG_TIMES=2: every CTA of the first registration, rows [CTA][iteration][8])
#define NLO_STAMP(slot)                                                                         \
  do {                                                                                          \
    if (p.debug_times != nullptr && blockIdx.y == 0 && threadIdx.x == 0 && it < kDebugIterations && \
        (blockIdx.x == 0 || (p.debug_all_ctas && blockIdx.x < kDebugCtas)))                     \
      p.debug_times[(static_cast<size_t>(blockIdx.x) * kDebugIterations + it) * kDebugSlots + (slot)] = GlobalTimerNs(); \
  } while (0)

enum ReduceOutcome : int {
  kReduceStep = 0,    // red.total holds the canonical sums of the whole registration: step
  kReduceFailed = 2   // a polled wait expired
};

// Everything between the tile loop and the damped step of iteration `it` in the resident kernel.
//  1. CTA sum (fixed order over the 8 warps); the CTAs of a thread-block cluster hand their 28
//     sums to the cluster's rank-0 CTA through distributed shared memory and meet at the hardware
//     cluster barrier; rank 0 adds them in rank order and stores the cluster's partial as "LL"
//     words (32 payload bits + the tag of this iteration per 8-byte store: a word is valid the
//     moment its tag matches -- no counter, no fence).
//  2. Only the rank-0 CTAs poll global memory.  gather_direct (few clusters): every rank-0 CTA
//     gathers all cluster partials and adds them in cluster order.  Otherwise CTA 0 gathers,
//     rotates to the canonical frame and stores the sums (LL again) into the local slot -- or,
//     sharded across GPUs, into the slot of every rank over NVLink -- and the rank-0 CTAs gather
//     those (in rank order: bit-identical sums, and therefore steps, on every GPU).
//  3. Rank 0 hands the totals to the other CTAs of its cluster over distributed shared memory, second
//     cluster barrier; the other CTAs sleep in that hardware barrier meanwhile instead of polling:
//     132 CTAs polling the same few L2 lines serialise in the L2 slices and delay the very stores
//     they wait for (measured: 2.7 us for the all-CTAs gather of 33 partials).
//  4. Every CTA rotates to the canonical frame (if step 2 did not) and steps its own copy of the state.
template <int KIND>
__device__ __forceinline__ int ReduceAndExchange(ResidentSmem& sm, const IterParams& p, const CanonPlan& plan, int it) {
  ReduceArea& red = sm.red;
  constexpr int NACC = KindTraits<KIND>::kAcc;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int problem = blockIdx.y;
  const int grid_x = gridDim.x;
  State& st = red.state;
  NLO_STAMP(8);

  double cta_sum = 0.0;
  if (tid < NACC) {
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) cta_sum += red.warp_sums[w][tid];
  }
  // the cluster this CTA belongs to (1 x 1 x 1 when the launch carries no cluster attribute)
  const unsigned int csize = ClusterSize(), crank = ClusterCtaRank();
  const bool leader = crank == 0;
  const bool grid_exchange = grid_x > 1 && !p.gather_direct;  // CTA 0 gathers on behalf of the grid
  bool canonical_done = false;
  if (grid_x == 1) {
    if (tid < NACC) red.total[tid] = cta_sum;
    NLO_STAMP(3);
  } else {
    const int n_clusters = static_cast<int>(ClusterCountX()), cluster_id = static_cast<int>(ClusterIdX());
    const int parity = it & 1;
    const unsigned int tag = p.tag_base + static_cast<unsigned int>(it) + 1u;
    unsigned long long* row_partials =
        p.ll_partials + static_cast<size_t>(parity * gridDim.y + problem) * n_clusters * (2 * NACC);
    if (csize > 1) {
      if (tid < NACC) StoreClusterF64(&sm.cluster_part[parity][crank][tid], 0u, cta_sum);
      ClusterSync();
      if (leader && tid < NACC) {
        cta_sum = 0.0;
        for (unsigned int r = 0; r < csize; ++r) cta_sum += sm.cluster_part[parity][r][tid];
      }
    }
    if (n_clusters == 1) {
      // one cluster holds the whole registration: its rank-0 CTA already has the sums, nothing goes through L2
      if (leader && tid < NACC) red.total[tid] = cta_sum;
      NLO_STAMP(3);
    } else {
      if (leader && tid < NACC)
        StoreLL(row_partials + static_cast<size_t>(cluster_id) * (2 * NACC) + 2 * tid, cta_sum, tag, false);
      NLO_STAMP(3);
      if (grid_exchange ? blockIdx.x == 0 : leader)
        GatherLL<NACC>(row_partials, 2 * NACC, n_clusters, tag, false, kGridTimeoutNs, red.gather_lanes, red.total,
                       &red.fail);
    }
  }
  NLO_STAMP(4);
  // Exchange of the canonical sums: over NVLink when the scan is sharded across GPUs, through the
  // local slot when CTA 0 gathered on behalf of the grid.
  if (p.use_peer || grid_exchange) {
    const bool pusher = grid_x == 1 || blockIdx.x == 0;
    if (pusher && KIND != kNdt3 && warp == 0)  // raw -> canonical, by warp 0 (which holds red.total)
      CanonicalRotate(red.total, st.R, &red.gather_lanes[0][0], plan, lane);
    canonical_done = true;
    // one exchange per loop iteration in every launch shape that exchanges
    const unsigned long long seq = red.seq0 + static_cast<unsigned long long>(it) + 1ULL;
    unsigned int xtag;
    int nsrc;
    const unsigned long long* src;
    if (p.use_peer) {
      xtag = static_cast<unsigned int>(seq);
      const int xpar = static_cast<int>(seq & 1ULL);
      nsrc = p.peer.nranks;
      src = p.peer.slots[p.peer.rank] + static_cast<size_t>(xpar) * kMaxRanks * kPeerWords;
      if (pusher && tid < NACC) {  // the thread that holds red.total[tid]
        const size_t slot = (static_cast<size_t>(xpar) * kMaxRanks + p.peer.rank) * kPeerWords + 2 * tid;
        for (int r = 0; r < nsrc; ++r) StoreLL(p.peer.slots[r] + slot, red.total[tid], xtag, true);
      }
    } else {
      xtag = p.tag_base + static_cast<unsigned int>(it) + 1u;
      nsrc = 1;
      unsigned long long* slot = p.ll_sums + (static_cast<size_t>(problem) * 2 + (it & 1)) * kPeerWords;
      src = slot;
      if (pusher && tid < NACC) StoreLL(slot + 2 * tid, red.total[tid], xtag, false);
    }
    if (leader) {
      __syncthreads();  // gather_lanes is reused
      GatherLL<NACC>(src, kPeerWords, nsrc, xtag, p.use_peer != 0, p.use_peer ? kPeerTimeoutNs : kGridTimeoutNs,
                     red.gather_lanes, red.total, &red.fail);
    }
    if (tid == 0) {
      red.exchanges += 1;
      if (p.use_peer && red.fail && pusher) *p.peer.error = 1;
    }
  }
  if (csize > 1) {
    // rank 0 -> the other CTAs of the cluster: the totals (and whether a wait expired)
    if (leader && tid < NACC) {
      const double v = red.total[tid];  // written by this thread
      for (unsigned int r = 1; r < csize; ++r) StoreClusterF64(&red.total[tid], r, v);
    }
    if (leader && tid == 0 && red.fail)
      for (unsigned int r = 1; r < csize; ++r) StoreClusterU32(&red.fail, r, 1u);
    ClusterSync();
  }
  NLO_STAMP(9);
  if (!canonical_done && KIND != kNdt3 && warp == 0)
    CanonicalRotate(red.total, st.R, &red.gather_lanes[0][0], plan, lane);
  NLO_STAMP(10);
  return red.fail ? kReduceFailed : kReduceStep;
}

// The damped step of one iteration on the canonical sums in red.total: warp 0 of the CTA, lane 0
// computing (the 6x6 factorisation is one dependent chain), all lanes copying the new state out.
template <int KIND>
__device__ __forceinline__ void StepPhase(ReduceArea& red, const IterParams& p, State* st_global, int it) {
  NLO_STAMP(12);
  const int lane = threadIdx.x & 31;
  const bool writer = (blockIdx.x == 0) || !p.persistent;
  State& st = red.state;
  __syncwarp();
  if (lane == 0) {
    double* trace_row = nullptr;
    if (p.trace != nullptr && writer)
      trace_row = p.trace + (static_cast<size_t>(blockIdx.y) * p.max_iterations + st.iteration) * KindTraits<KIND>::kTrace;
    if (KIND == kNdt3)
      Step3(red.total, &st, p.parameter_tolerance, p.gradient_tolerance, p.max_iterations, trace_row);
    else
      Step6(red.total, &st, p.parameter_tolerance, p.gradient_tolerance, p.max_iterations, trace_row);
  }
  NLO_STAMP(13);
  __syncwarp();
  if (writer) {  // the state in HBM: what the host reads, and what the next launch starts from
    constexpr int kWords = static_cast<int>(sizeof(State) / sizeof(double));
    if (lane < kWords) reinterpret_cast<double*>(st_global)[lane] = reinterpret_cast<const double*>(&st)[lane];
  }
  NLO_STAMP(14);
}

// Reduction, exchange and step of iteration `it`, called by all threads of the CTA after the tile
// loop; returns a ReduceOutcome.
template <int KIND>
__device__ __forceinline__ int PostTileBody(ResidentSmem& sm, const IterParams& p, const CanonPlan& plan,
                                            State* st_global, int it) {
  ReduceArea& red = sm.red;
  const int outcome = ReduceAndExchange<KIND>(sm, p, plan, it);
  NLO_STAMP(11);
  if (outcome == kReduceStep && threadIdx.x < 32) StepPhase<KIND>(red, p, st_global, it);
  return outcome;
}
// ------------------------------------------------------------------ streaming kernel
// The kernel for scans that do not fit the shared memory of the grid: tiles stream through a TMA /
// mbarrier ring, 2 CTAs per SM, 128 registers -- all of which the tile loop uses.  Its
// per-iteration exchange is the leader form: every CTA stores its partial and arrives on a
// counter, CTA 0 sums them in fixed order, [all-reduces over NVLink], steps and publishes the new
// state as LL words that the other CTAs poll.  The cluster / all-gather form of the resident
// kernel was tried here too and is slower for this shape: inlined, its code drives dozens of spills
// into the tile loop (ptxas budgets inlined and directly called code together with the loop);
// called through a pointer, the ABI register saves of 75 000 threads per iteration go straight to
// L2 (the L1 left beside 2 x 104 KB of stages is ~40 KB) and cost 9 - 18 us per iteration
// (1 M points: 30.8 against 22.1 us; 16 M: 238.7 against 220.4 us, same box).
__device__ __noinline__ void Step6Call(const double* __restrict__ sums, State* st, double ptol, double gtol,
                                       int max_iterations, double* trace_row) {
  Step6(sums, st, ptol, gtol, max_iterations, trace_row);
}

// One-shot all-reduce of `total[0..nacc)` over peer-mapped buffers; called by all threads of the
// finalising CTA.  "LL" wire format: every 8-byte word carries 4 bytes of payload and the 4-byte
// sequence number of the exchange, so a word is valid the moment its sequence matches -- one
// NVLink one-way latency, no fences, no separate flag.  A double travels as two words.  Slots are
// double-buffered by sequence parity (a rank can only be one exchange ahead of a peer).  The sum
// runs in rank order 0..n-1 on every rank => bit-identical results everywhere.
__device__ inline void PeerAllReduce(const PeerComm& pc, double* total, int nacc) {
  __shared__ unsigned int halves[kMaxRanks][kPeerWords];
  const int tid = threadIdx.x;
  const unsigned long long seq = *pc.seq + 1;
  const unsigned int seq32 = static_cast<unsigned int>(seq);
  const int parity = static_cast<int>(seq & 1ULL);
  const int words = 2 * nacc;
  // One (rank, word) item per thread, for the stores and for the polls: with the ranks walked one
  // after the other by the same thread every rank costs a system-scope load round trip even when
  // its words are already there (8 ranks: ~6 us of the 8-GPU step); spread over the threads the
  // wait is one round trip after the last rank's words land.
  const int items = pc.nranks * words;
  const int my_slot = (parity * kMaxRanks + pc.rank) * kPeerWords;
  for (int item = tid; item < items; item += blockDim.x) {
    const int r = item / words, w = item - r * words;
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(total[w >> 1]));
    const unsigned long long half = (w & 1) ? (bits >> 32) : (bits & 0xffffffffULL);
    *reinterpret_cast<volatile unsigned long long*>(pc.slots[r] + my_slot + w) =
        (static_cast<unsigned long long>(seq32) << 32) | half;
  }
  SpinWatch watch;
  for (int item = tid; item < items; item += blockDim.x) {
    const int r = item / words, w = item - r * words;
    const volatile unsigned long long* src =
        reinterpret_cast<const volatile unsigned long long*>(pc.slots[pc.rank]) + (parity * kMaxRanks + r) * kPeerWords + w;
    unsigned long long v = *src;
    while (static_cast<unsigned int>(v >> 32) != seq32) {
      if (watch.Expired(kPeerTimeoutNs)) {  // a peer died; fail instead of hanging
        *pc.error = 1;
        break;
      }
      v = *src;
    }
    halves[r][w] = static_cast<unsigned int>(v);
  }
  __syncthreads();
  if (tid < nacc) {
    double s = 0.0;
    for (int r = 0; r < pc.nranks; ++r) {
      const unsigned long long bits = (static_cast<unsigned long long>(halves[r][2 * tid + 1]) << 32) |
                                      static_cast<unsigned long long>(halves[r][2 * tid]);
      s += __longlong_as_double(static_cast<long long>(bits));
    }
    total[tid] = s;
  }
  if (tid == 0) *pc.seq = seq;
  __syncthreads();
}


// WG = 1: 256 threads, two CTAs per SM, one 256-correspondence tile per ring stage.
// WG = 2: 512 threads, ONE CTA per SM, two consecutive tiles (one contiguous 48 KB run) per stage, each
//   of the two warp groups taking one of them.  Same warps, registers and bytes in flight per SM, but a
//   single tile stream per SM: with two CTAs per SM the one placed second finishes its tiles 5 - 8 us
//   after its older neighbour in EVERY iteration of an 8 M-point shard (12 us between the first and the
//   last CTA of the grid; 3 us with one CTA per SM), and everybody waits for the last one.  It also
//   halves the partials, arrivals and pollers of the exchange.  Used for the persistent loop of a
//   streamed scan; batched registrations (tile-aligned, one CTA each) keep WG = 1.
template <int KIND, int LOSS, typename ST, int WG>
__global__ void __launch_bounds__(kThreads * WG, WG == 1 ? 2 : 1) gn_iteration_kernel(const IterParams p) {
  using T = KindTraits<KIND>;
  constexpr int NACC = T::kAcc;
  constexpr int NPLANES = PlanesOf<KIND, ST>();
  constexpr int MAX_STAGES = SmemLayout<KIND, ST, WG>::kStages;
  constexpr int kWarps = kConsumerWarps * WG;
  // ring depth actually used (<= the stages allocated): chosen per launch shape by the host
  const int STAGES = (p.stage_depth > 0 && p.stage_depth < MAX_STAGES) ? p.stage_depth : MAX_STAGES;
  constexpr uint32_t kStageBytes = NPLANES * kTile * sizeof(ST);

  extern __shared__ __align__(128) unsigned char smem_raw[];
  SmemLayout<KIND, ST, WG>& sm = *reinterpret_cast<SmemLayout<KIND, ST, WG>*>(smem_raw);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int sub = (WG == 1) ? 0 : (tid >> 8);    // warp group = tile within a stage
  const int elem = (WG == 1) ? tid : (tid & 255);  // correspondence within the tile
  const int problem = blockIdx.y;
  const int grid_x = gridDim.x;
  State* st_global = p.states + problem;
  State& st = sm.red.state;  // CTA-local copy; in the persistent path every CTA steps it redundantly

  if (tid == 0) {
    if (p.mode != kModeStepOnly) {
      for (int s = 0; s < MAX_STAGES; ++s) {
        MbarInit(&sm.full[s], 1);
        MbarInit(&sm.empty[s], kWarps);
      }
      FenceBarrierInit();
    }
    st = *st_global;
  }
  if (WG == 2 && p.mode != kModeStepOnly) {  // see the tile loop: nothing non-finite may sit in a stage
    ST* flat = &sm.stages[0][0][0][0];
    for (int k = tid; k < MAX_STAGES * WG * NPLANES * kTile; k += kThreads * WG) flat[k] = static_cast<ST>(0);
  }
  __syncthreads();

  const Range range = p.ranges[problem];
  const int64_t tile_lo = range.begin / kTile;
  const int64_t tile_hi = (range.end + kTile - 1) / kTile;
  // tiles of this CTA: tile_lo + blockIdx.x + m * grid_x.  (Static, so that the sums are reproducible.
  // A biased split between the two CTAs of an SM -- the younger one finishes 5 - 8 us later in every
  // iteration of an 8 M-point shard -- was tried and is slower: the SM's bandwidth is shared, the
  // SM is done when both are, 51.6 / 48.4 costs 2 %.)
  // (in units of stages: WG consecutive tiles; the last stage of a range may hold fewer).  Positions
  // inside the range are 32-bit -- a problem holds < 2^31 correspondences -- which spares the tile
  // loop a few registers.
  const int span_tiles = static_cast<int>(tile_hi - tile_lo);
  const int span = (span_tiles + WG - 1) / WG;
  const int my_tiles =
      (span > static_cast<int>(blockIdx.x)) ? (span - static_cast<int>(blockIdx.x) + grid_x - 1) / grid_x : 0;
  const int valid_lo = static_cast<int>(range.begin - tile_lo * kTile);  // first / one-past-last valid
  const int valid_hi = static_cast<int>(range.end - tile_lo * kTile);    // correspondence, from tile_lo
  const bool writer = (blockIdx.x == 0) || !p.persistent;  // who publishes state / trace / sums

  const uint64_t policy_keep = PolicyEvictLast(), policy_stream = PolicyEvictFirst();
  // Ring positions are carried incrementally (no divisions in the tile loop): the consumers'
  // next (stage, phase) and -- thread 0 only -- the producer's, which runs STAGES-1 tiles ahead.
  int c_stage = 0, p_stage = 0;
  uint32_t c_phase = 0, p_phase = 0;
  // When the CTA's share of the scan fits the stage ring and the loop runs in-kernel, the tiles are
  // loaded once and stay resident in shared memory for every later iteration (no HBM/L2 re-read).
  // (never with WG = 2: registrations that small run the resident kernel)
  const bool resident = (WG == 1) && (p.iterations_in_kernel > 1) && (my_tiles <= STAGES);
  // Streaming case: the first tiles of the NEXT iteration (the same tiles, they do not depend on
  // the pose) are requested before this iteration's reduction / step, which hides the pipeline
  // ramp behind the grid-wide exchange.  `prefetched` tiles of the coming iteration are in flight.
  int prefetched = 0;

  for (int it = 0; it < p.iterations_in_kernel; ++it) {
    if (st.done) break;  // uniform: the state only changes behind a __syncthreads
    NLO_STAMP(0);

    if (p.mode != kModeStepOnly) {
      double acc[NACC];
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[k] = 0.0;

      {
        // Producer duty (thread 0): keep STAGES-1 tiles in flight ahead of the tile being consumed.
        auto issue_tile = [&](int m) {
          const int s = p_stage;
          const uint32_t phase = p_phase;
          if (++p_stage == STAGES) { p_stage = 0; p_phase ^= 1u; }
          MbarWait(&sm.empty[s], phase ^ 1u);
          const int first = (static_cast<int>(blockIdx.x) + m * grid_x) * WG;  // tiles before it in the range
          const int64_t tile = tile_lo + first;
          const uint32_t bytes =
              (WG == 1 || span_tiles - first >= WG) ? kStageBytes * WG : kStageBytes * static_cast<uint32_t>(span_tiles - first);
          MbarExpectTx(&sm.full[s], bytes);
          // one contiguous NP x 2 KB run per tile (tile-interleaved layout): a single bulk copy.
          // l2_keep_tiles > 0: the scan is re-read every iteration and is larger than what the L2
          // keeps by itself -- pin the first tiles of every CTA (evict_last, l2_keep_tiles in all),
          // stream the rest (evict_first).
          const ST* src = reinterpret_cast<const ST*>(p.planes[0]) + tile * (NPLANES * kTile);
          if (p.l2_keep_tiles > 0)
            BulkLoadHint(&sm.stages[s][0][0][0], src, bytes, &sm.full[s],
                         static_cast<long long>(m) * grid_x * WG < p.l2_keep_tiles ? policy_keep : policy_stream);
          else
            BulkLoad(&sm.stages[s][0][0][0], src, bytes, &sm.full[s]);
        };
        const bool need_load = !resident || it == 0;
        if (tid == 0 && need_load) {
          for (int m = prefetched; m < STAGES - 1 && m < my_tiles; ++m) issue_tile(m);
        }
        prefetched = 0;
        __syncwarp();
        double R[9], t[3];
        if (KIND == kNdt3) {
#pragma unroll
          for (int k = 0; k < 4; ++k) R[k] = st.R[k];
          t[0] = st.t[0];
          t[1] = st.t[1];
        } else {
#pragma unroll
          for (int k = 0; k < 9; ++k) R[k] = st.R[k];
#pragma unroll
          for (int k = 0; k < 3; ++k) t[k] = st.t[k];
        }
        for (int m = 0; m < my_tiles; ++m) {
          const int s = resident ? m : c_stage;
          const uint32_t phase = c_phase;
          if (!resident && ++c_stage == STAGES) { c_stage = 0; c_phase ^= 1u; }
          if (need_load) {
            if (tid == 0 && m + STAGES - 1 < my_tiles) issue_tile(m + STAGES - 1);
            __syncwarp();
            MbarWait(&sm.full[s], phase);
          }
          // this warp group's tile of the stage (the last stage of the range may not have one)
          // (WG = 2: the last stage of the range may have no tile for the second warp group; it then
          // computes on what the stage held before -- zeros or an earlier tile, finite either way -- and
          // the validity test below, which covers the whole range, masks it out)
          double v[NPLANES];
#pragma unroll
          for (int pl = 0; pl < NPLANES; ++pl) v[pl] = static_cast<double>(sm.stages[s][sub][pl][elem]);
          if (!resident) {
            __syncwarp();
            if (lane == 0) MbarArrive(&sm.empty[s]);
          }
          const int idx = ((static_cast<int>(blockIdx.x) + m * grid_x) * WG + sub) * kTile + elem;
          const bool valid = (idx >= valid_lo) && (idx < valid_hi);
          if (KIND == kNdt6)
            Ndt6Point<LOSS>(v, R, t, p.loss_p0, p.loss_p1, valid, acc);
          else if (KIND == kNdt3)
            Ndt3Point<LOSS>(v, R, t, p.loss_p0, p.loss_p1, valid, acc);
          else
            ReprojPoint<LOSS>(v, R, t, p.intrinsics, p.loss_p0, p.loss_p1, valid, acc);
        }
        // Warp reduction of the NACC accumulators by recursive halving: at offset 16 each lane
        // keeps half of the values and hands the other half to its partner, at offset 8 a
        // quarter, ... so the butterfly costs 16+8+4+2+1 = 31 double shuffles per lane instead of
        // 5 per value (140 for 28 values) -- SHFL issue is shared by the whole SM and was the
        // largest fixed cost of an iteration.  Lane l ends up owning the warp total of value l.
        // The tree is fixed, so the sum is deterministic.
        {
          double v[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = (k < NACC) ? acc[k] : 0.0;
#pragma unroll
          for (int half = 16; half >= 1; half >>= 1) {
            const bool upper = (lane & half) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
              if (i < NACC) {  // compile-time after unrolling: skip slots that are all padding
                const double send = upper ? v[i] : v[i + half];
                const double keep = upper ? v[i + half] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
              }
            }
          }
          if (lane < NACC) sm.red.warp_sums[warp][lane] = v[0];
        }
        if (!resident) {
          if (it + 1 < p.iterations_in_kernel && p.mode == kModeSolve) {
            const int ahead = (my_tiles < STAGES - 1) ? my_tiles : STAGES - 1;
            if (tid == 0)
              for (int m = 0; m < ahead; ++m) issue_tile(m);
            prefetched = ahead;
          }
        }
      }
      NLO_STAMP(1);
      __syncthreads();
      NLO_STAMP(2);

      // CTA sum over the consumer warps, fixed order
      if (tid < NACC) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += sm.red.warp_sums[w][tid];
        sm.red.total[tid] = s;
      }
      if (grid_x > 1) {
        // per-CTA partial -> HBM/L2; double-buffered by iteration parity for the persistent path
        double* partial_base =
            p.partials + static_cast<size_t>((it & 1) * gridDim.y + problem) * grid_x * NACC;
        if (tid < NACC) {
          __stcg(partial_base + static_cast<size_t>(blockIdx.x) * NACC + tid, sm.red.total[tid]);
          // persistent path: thread 0's releasing fence after the bar.sync below covers these
          // stores by cumulativity; the ticket path fences per writer
          if (!p.persistent) __threadfence();
        }
        if (p.persistent) {
          // Persistent grid: every CTA arrives on a counter; CTA 0 (the leader) waits for all
          // partials, reduces them, [all-reduces over NVLink], steps and PUBLISHES the new 160-byte
          // state as 40 "LL" words (32 payload bits + the iteration number in one 8-byte store, so
          // a word is valid the moment its number matches: no fence, no separate flag); the other
          // CTAs poll those words -- one L2 round trip after the leader's stores -- and rebuild
          // the state.
          constexpr int kStateWords = 2 * static_cast<int>(sizeof(State) / sizeof(double));
          unsigned int* counter = reinterpret_cast<unsigned int*>(p.ll_sums + problem * kSyncStride);
          unsigned long long* ll_state = p.ll_sums + problem * kSyncStride + 8;
          const unsigned int want = static_cast<unsigned int>(it) + 1u;
          __syncthreads();
          if (tid == 0) {
            __threadfence();  // releases this CTA's partial (cumulative over the bar.sync above)
            atomicAdd(counter, 1u);
            sm.red.flag = 1;
          }
          SpinWatch watch;
          if (blockIdx.x == 0) {
            if (tid == 0) {
              while (LoadGpuU32(counter) < want * grid_x)
                if (watch.Expired(kGridTimeoutNs)) { sm.red.flag = 0; break; }
              __threadfence();
            }
            __syncthreads();
            if (sm.red.flag == 0) {  // a CTA went missing (cannot happen under a cooperative launch)
              if (tid == 0) { st.status = 2; st.done = 1; *st_global = st; }
              __syncthreads();
              if (tid < kStateWords) {  // still publish, so that the other CTAs leave as well
                const unsigned long long bits = static_cast<unsigned long long>(
                    __double_as_longlong(reinterpret_cast<const double*>(&st)[tid >> 1]));
                __stcg(ll_state + tid, (static_cast<unsigned long long>(want) << 32) |
                                           ((tid & 1) ? (bits >> 32) : (bits & 0xffffffffULL)));
              }
              break;
            }
          } else {
            __syncthreads();  // sm.red.flag = 1 visible
            if (tid < kStateWords) {
              const unsigned long long* src = ll_state + tid;
              unsigned long long w = LoadGpuU64(src);
              while (static_cast<unsigned int>(w >> 32) != want) {
                if (watch.Expired(kGridTimeoutNs)) { sm.red.flag = 0; break; }
                w = LoadGpuU64(src);
              }
              sm.red.halves[tid] = static_cast<unsigned int>(w);
            }
            __syncthreads();
            if (sm.red.flag == 0) {  // the leader went missing
              if (tid == 0) { st.status = 2; st.done = 1; }
              __syncthreads();
              break;
            }
            if (tid < kStateWords / 2)
              reinterpret_cast<double*>(&st)[tid] = __longlong_as_double(static_cast<long long>(
                  (static_cast<unsigned long long>(sm.red.halves[2 * tid + 1]) << 32) | sm.red.halves[2 * tid]));
            __syncthreads();
            continue;
          }
        } else {
          __syncthreads();
          if (tid == 0) {
            const unsigned int ticket = atomicAdd(p.tickets + problem, 1u);
            sm.red.flag = (ticket == static_cast<unsigned int>(grid_x) - 1u) ? 1 : 0;
            if (sm.red.flag) p.tickets[problem] = 0u;  // ready for the next launch
          }
          __syncthreads();
          if (sm.red.flag == 0) return;  // only the last CTA to arrive carries on
          __threadfence();
        }
        NLO_STAMP(3);
        // cross-CTA sum: thread (j, l8) adds CTAs l8, l8+8, ...; then the 8 lanes in order
        const int j = tid >> 3, l8 = tid & 7;
        if (j < NACC) {
          const double* base = partial_base + j;
          double s = 0.0;
          int g = l8;
          // loads are issued 16 at a time (independent addresses, the tail predicated so that it
          // costs one round trip instead of one per element); the adds keep the fixed order
          for (; g < grid_x; g += 128) {
            double v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int c = g + 8 * u;
              v[u] = c < grid_x ? __ldcg(base + static_cast<size_t>(c) * NACC) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) s += v[u];
          }
          sm.red.warp_sums[l8][j] = s;
        }
        __syncthreads();
        if (tid < NACC) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < 8; ++w) s += sm.red.warp_sums[w][tid];
          sm.red.total[tid] = s;
        }
      }
      __syncthreads();

      NLO_STAMP(4);
      // raw -> canonical (needs the R the sums were taken at), by warp 0
      if (KIND != kNdt3 && warp == 0) {
        const CanonPlan plan = MakeCanonPlan(lane);
        CanonicalRotate(sm.red.total, st.R, &sm.red.gather_lanes[0][0], plan, lane);
      }
      __syncthreads();
      if (p.use_peer) PeerAllReduce(p.peer, sm.red.total, NACC);
      if (p.mode == kModeAssemble) {
        if (tid < NACC && writer) p.sums[problem * 32 + tid] = sm.red.total[tid];
        return;
      }
    } else {
      if (tid < NACC) sm.red.total[tid] = p.sums[problem * 32 + tid];
      __syncthreads();
    }

    // ---------------- damped step by warp 0 (redundantly per CTA in the persistent path)
    if (warp == 0) {
      double* trace_row = nullptr;
      if (p.trace != nullptr && writer)
        trace_row = p.trace + (static_cast<size_t>(problem) * p.max_iterations + st.iteration) *
                                  T::kTrace;
      if (lane == 0) {
        if (KIND == kNdt3)
          Step3(sm.red.total, &st, p.parameter_tolerance, p.gradient_tolerance, p.max_iterations,
                trace_row);
        else
          Step6Call(sm.red.total, &st, p.parameter_tolerance, p.gradient_tolerance, p.max_iterations,
                trace_row);
      }
      if (lane == 0 && writer) *st_global = st;
    }
    NLO_STAMP(5);
    __syncthreads();
    if (p.persistent && grid_x > 1 && p.mode == kModeSolve) {  // leader: publish the new state
      constexpr int kStateWords = 2 * static_cast<int>(sizeof(State) / sizeof(double));
      if (tid < kStateWords) {
        const unsigned long long bits = static_cast<unsigned long long>(
            __double_as_longlong(reinterpret_cast<const double*>(&st)[tid >> 1]));
        __stcg(p.ll_sums + problem * kSyncStride + 8 + tid,
               (static_cast<unsigned long long>(it + 1) << 32) |
                   ((tid & 1) ? (bits >> 32) : (bits & 0xffffffffULL)));
      }
    }
    NLO_STAMP(6);
  }
  // a prefetch may still be in flight when the loop ends early: let it land before the CTA exits
  if (prefetched > 0) {
    for (int m = 0; m < prefetched; ++m) {
      MbarWait(&sm.full[c_stage], c_phase);
      if (++c_stage == STAGES) { c_stage = 0; c_phase ^= 1u; }
    }
  }
}

// ------------------------------------------------------------------ resident kernel
// gn_resident_kernel<KIND, LOSS>: the device-resident loop of a registration whose correspondences
// fit the shared memory of the grid (<= resident_stages tiles per CTA: 100 k NDT points are 3 tiles
// on each of 132 - 148 CTAs).  This is the latency-bound regime -- the reference's own problem
// sizes, 9 k - 100 k correspondences -- where an iteration is a few microseconds and everything
// that is not arithmetic shows: the kernel is therefore its own, small piece of code (it stays in
// the instruction cache; the streaming kernel's per-iteration path does not), runs one CTA per SM
// with up to 255 registers per thread so that the reduction, the exchange and the 6x6 step are
// inlined next to the tile loop without spilling, reads the tiles from HBM exactly once per Solve
// (one bulk copy per tile, one mbarrier) and takes two correspondences per thread and step for
// instruction-level parallelism on the fp64 pipe.
template <int KIND, int LOSS>
__global__ void __launch_bounds__(kThreads, 1) gn_resident_kernel(const __grid_constant__ IterParams p) {
  using T = KindTraits<KIND>;
  constexpr int NACC = T::kAcc;
  constexpr int NPLANES = PlanesOf<KIND, double>();
  constexpr uint32_t kStageBytes = NPLANES * kTile * sizeof(double);

  extern __shared__ __align__(128) unsigned char smem_raw[];
  ResidentSmem& sm = *reinterpret_cast<ResidentSmem*>(smem_raw);
  // stages start at the next 128-byte boundary after the header
  double(*stages)[NPLANES][kTile] = reinterpret_cast<double(*)[NPLANES][kTile]>(
      smem_raw + ((sizeof(ResidentSmem) + 127) / 128) * 128);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int problem = blockIdx.y;
  const int grid_x = gridDim.x;
  State* st_global = p.states + problem;
  State& st = sm.red.state;
  const bool writer = blockIdx.x == 0;

  const Range range = p.ranges[problem];
  const int64_t tile_lo = range.begin / kTile;
  const int64_t span = (range.end + kTile - 1) / kTile - tile_lo;
  const int my_tiles =
      (span > blockIdx.x) ? static_cast<int>((span - blockIdx.x + grid_x - 1) / grid_x) : 0;

  if (tid == 0) {
    MbarInit(&sm.full, 1);
    FenceBarrierInit();
    st = *st_global;
    sm.red.fail = 0;
    sm.red.seq0 = p.use_peer ? *p.peer.seq : 0ULL;
    sm.red.exchanges = 0;
    if (my_tiles > 0) {
      MbarExpectTx(&sm.full, kStageBytes * static_cast<uint32_t>(my_tiles));
      for (int m = 0; m < my_tiles; ++m) {
        const int64_t tile = tile_lo + blockIdx.x + static_cast<int64_t>(m) * grid_x;
        BulkLoad(&stages[m][0][0], p.planes[0] + tile * (NPLANES * kTile), kStageBytes, &sm.full);
      }
    }
  }
  __syncthreads();
  // no CTA may store into the shared memory of a cluster peer that has not started yet
  if (ClusterSize() > 1) ClusterSync();
  if (my_tiles > 0) MbarWait(&sm.full, 0);

  // validity of this thread's correspondence in each tile (only the range's first / last tile are ragged)
  auto valid_in = [&](int m) {
    const int64_t idx = (tile_lo + blockIdx.x + static_cast<int64_t>(m) * grid_x) * kTile + tid;
    return idx >= range.begin && idx < range.end;
  };
  auto one_point = [&](const double* v, const double* R, const double* t, bool valid, double* acc) {
    if (KIND == kNdt6)
      Ndt6Point<LOSS>(v, R, t, p.loss_p0, p.loss_p1, valid, acc);
    else if (KIND == kNdt3)
      Ndt3Point<LOSS>(v, R, t, p.loss_p0, p.loss_p1, valid, acc);
    else
      ReprojPoint<LOSS>(v, R, t, p.intrinsics, p.loss_p0, p.loss_p1, valid, acc);
  };

  const CanonPlan canon_plan = MakeCanonPlan(lane);  // what this lane does in the rotation to the canonical frame
  for (int it = 0; it < p.iterations_in_kernel; ++it) {
    if (st.done) break;  // uniform: every CTA steps the same state
    NLO_STAMP(0);
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    double R[9], t[3];
    if (KIND == kNdt3) {
#pragma unroll
      for (int k = 0; k < 4; ++k) R[k] = st.R[k];
      t[0] = st.t[0];
      t[1] = st.t[1];
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = st.R[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] = st.t[k];
    }
    // Two independent correspondences in flight per thread; an odd tile count pairs the last tile
    // with itself, masked out (one loop body: the code has to stay small, see GatherLL).
#pragma unroll 1
    for (int m = 0; m < my_tiles; m += 2) {
      const bool pair = m + 1 < my_tiles;
      const int mb = pair ? m + 1 : m;
      double va[NPLANES], vb[NPLANES];
#pragma unroll
      for (int pl = 0; pl < NPLANES; ++pl) {
        va[pl] = stages[m][pl][tid];
        vb[pl] = stages[mb][pl][tid];
      }
      double accb[NACC];
#pragma unroll
      for (int k = 0; k < NACC; ++k) accb[k] = 0.0;
      one_point(va, R, t, valid_in(m), acc);
      one_point(vb, R, t, pair && valid_in(mb), accb);
#pragma unroll
      for (int k = 0; k < NACC; ++k) acc[k] += accb[k];
    }
    // warp reduction by recursive halving (see the streaming kernel): lane l ends with value l
    {
      double v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = (k < NACC) ? acc[k] : 0.0;
#pragma unroll
      for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          if (i < NACC) {
            const double send = upper ? v[i] : v[i + half];
            const double keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
          }
        }
      }
      if (lane < NACC) sm.red.warp_sums[warp][lane] = v[0];
    }
    NLO_STAMP(1);
    __syncthreads();
    NLO_STAMP(2);
    const int outcome = PostTileBody<KIND>(sm, p, canon_plan, st_global, it);
    if (outcome == kReduceFailed) {
      if (tid == 0) {
        st.status = 2;
        st.done = 1;
        if (writer) *st_global = st;
      }
      __syncthreads();
      break;
    }
    NLO_STAMP(5);
    __syncthreads();
    NLO_STAMP(6);
    if (p.debug_all_ctas && it == 0 && tid == 0 && blockIdx.y == 0 && blockIdx.x < kDebugCtas && p.debug_times != nullptr) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      p.debug_times[(static_cast<size_t>(blockIdx.x) * kDebugIterations) * kDebugSlots + 7] = smid;
    }
  }
  if (p.use_peer && tid == 0 && sm.red.exchanges > 0) {
    const bool grid_exchange = grid_x > 1 && !p.gather_direct;
    if (!grid_exchange || blockIdx.x == 0)
      *p.peer.seq = sm.red.seq0 + static_cast<unsigned long long>(sm.red.exchanges);
  }
}

// ------------------------------------------------------------------ dispatch
struct KernelEntry {
  const void* fn = nullptr;
  size_t smem = 0;
};

template <int KIND, int LOSS>
static KernelEntry EntryOne(bool f32, int wg) {
  KernelEntry e;
  if (f32) {
    if (KIND == kReproj || wg != 1) return e;  // fp32 storage exists for the NDT kinds, one tile per stage
    constexpr int K = (KIND == kReproj ? kNdt6 : KIND);
    e.fn = reinterpret_cast<const void*>(&gn_iteration_kernel<K, LOSS, float, 1>);
    e.smem = SmemBytes<K, float, 1>();
  } else if (wg == 2) {
    e.fn = reinterpret_cast<const void*>(&gn_iteration_kernel<KIND, LOSS, double, 2>);
    e.smem = SmemBytes<KIND, double, 2>();
  } else if (wg == 1) {
    e.fn = reinterpret_cast<const void*>(&gn_iteration_kernel<KIND, LOSS, double, 1>);
    e.smem = SmemBytes<KIND, double, 1>();
  }
  return e;
}

template <int KIND>
static KernelEntry EntryKind(int loss, bool f32, int wg) {
  switch (loss) {
    case kLossNone: return EntryOne<KIND, kLossNone>(f32, wg);
    case kLossExponential: return EntryOne<KIND, kLossExponential>(f32, wg);
    case kLossHuber: return EntryOne<KIND, kLossHuber>(f32, wg);
    case kLossCauchy: return EntryOne<KIND, kLossCauchy>(f32, wg);
  }
  return KernelEntry();
}

static KernelEntry EntryFor(int kind, int loss, bool f32, int wg = 1) {
  switch (kind) {
    case kNdt6: return EntryKind<kNdt6>(loss, f32, wg);
    case kNdt3: return EntryKind<kNdt3>(loss, f32, wg);
    case kReproj: return EntryKind<kReproj>(loss, f32, wg);
  }
  return KernelEntry();
}

// Resident kernel (fp64 storage only): shared memory = header + one stage per tile of the CTA.
constexpr size_t kResidentHeaderBytes = ((sizeof(ResidentSmem) + 127) / 128) * 128;
constexpr size_t kMaxDynamicSmem = 227 * 1024;

static size_t ResidentStageBytes(int kind) {
  return static_cast<size_t>(kind == kReproj ? kReprojPlanes : kNdtPlanes) * kTile * sizeof(double);
}

template <int KIND>
static const void* ResidentKind(int loss) {
  switch (loss) {
    case kLossNone: return reinterpret_cast<const void*>(&gn_resident_kernel<KIND, kLossNone>);
    case kLossExponential: return reinterpret_cast<const void*>(&gn_resident_kernel<KIND, kLossExponential>);
    case kLossHuber: return reinterpret_cast<const void*>(&gn_resident_kernel<KIND, kLossHuber>);
    case kLossCauchy: return reinterpret_cast<const void*>(&gn_resident_kernel<KIND, kLossCauchy>);
  }
  return nullptr;
}

static const void* ResidentFor(int kind, int loss) {
  switch (kind) {
    case kNdt6: return ResidentKind<kNdt6>(loss);
    case kNdt3: return ResidentKind<kNdt3>(loss);
    case kReproj: return ResidentKind<kReproj>(loss);
  }
  return nullptr;
}

int ResidentMaxStages(int kind) {
  return static_cast<int>((kMaxDynamicSmem - kResidentHeaderBytes) / ResidentStageBytes(kind));
}

// Attributes of a launch: a persistent grid is launched cooperatively (all CTAs co-resident, the
// kernel spins on words written by other CTAs), a cluster size > 1 adds the cluster dimension.
static int FillAttributes(cudaLaunchAttribute* attrs, bool cooperative, int cluster) {
  int n = 0;
  if (cooperative) {
    attrs[n].id = cudaLaunchAttributeCooperative;
    attrs[n].val.cooperative = 1;
    ++n;
  }
  if (cluster > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = static_cast<unsigned int>(cluster);
    attrs[n].val.clusterDim.y = 1;
    attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  return n;
}

cudaError_t LaunchIteration(int kind, int loss, const IterParams& p, int grid_x, int num_problems,
                            int cluster, cudaStream_t stream, int warp_groups) {
  const KernelEntry e = EntryFor(kind, loss, p.f32 != 0, warp_groups);
  if (e.fn == nullptr || cluster < 1 || cluster > kMaxCluster || grid_x % cluster != 0) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid_x, num_problems);
  cfg.blockDim = dim3(kThreads * warp_groups);
  cfg.dynamicSmemBytes = e.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  cfg.numAttrs = FillAttributes(attrs, p.persistent != 0, cluster);
  cfg.attrs = attrs;
  IterParams copy = p;
  void* args[] = {&copy};
  return cudaLaunchKernelExC(&cfg, e.fn, args);
}

cudaError_t LaunchResident(int kind, int loss, const IterParams& p, int grid_x, int num_problems, int cluster,
                           int stages, cudaStream_t stream) {
  const void* fn = ResidentFor(kind, loss);
  if (fn == nullptr || p.f32 || cluster < 1 || cluster > kMaxCluster || grid_x % cluster != 0 || stages < 0 ||
      stages > ResidentMaxStages(kind))
    return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid_x, num_problems);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kResidentHeaderBytes + static_cast<size_t>(stages) * ResidentStageBytes(kind);
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  cfg.numAttrs = FillAttributes(attrs, true, cluster);
  cfg.attrs = attrs;
  IterParams copy = p;
  void* args[] = {&copy};
  return cudaLaunchKernelExC(&cfg, fn, args);
}

// CTAs of the iteration kernel (resident: of the resident kernel, one per SM) that can be
// co-resident on the current device when launched in clusters of `cluster` CTAs (the GPCs do not
// all hold a whole number of clusters).
int MaxCoResidentCtas(int kind, int loss, bool f32, int cluster, bool resident) {
  KernelEntry e = EntryFor(kind, loss, f32);
  if (resident) {
    e.fn = f32 ? nullptr : ResidentFor(kind, loss);
    e.smem = kResidentHeaderBytes + static_cast<size_t>(ResidentMaxStages(kind)) * ResidentStageBytes(kind);
  }
  if (e.fn == nullptr) return 0;
  if (cluster <= 1) {
    int per_sm = 0, device = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, e.fn, kThreads, e.smem) != cudaSuccess ||
        cudaGetDevice(&device) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return per_sm * sms;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster * 64);  // any multiple of the cluster size
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = e.smem;
  cudaLaunchAttribute attrs[2];
  cfg.numAttrs = FillAttributes(attrs, false, cluster);
  cfg.attrs = attrs;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, e.fn, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return clusters * cluster;
}

size_t IterationSmemBytes(int kind) {
  switch (kind) {
    case kNdt6: return SmemBytes<kNdt6>();
    case kNdt3: return SmemBytes<kNdt3>();
    default: return SmemBytes<kReproj>();
  }
}

cudaError_t ConfigureKernels() {
  for (int kind = 0; kind < 3; ++kind)
    for (int loss = 0; loss < 4; ++loss)
      for (int variant = 0; variant < 3; ++variant) {  // fp64 one tile per stage, fp32, fp64 two tiles per stage
        const KernelEntry e = EntryFor(kind, loss, variant == 1, variant == 2 ? 2 : 1);
        if (e.fn == nullptr) continue;
        const cudaError_t err = cudaFuncSetAttribute(e.fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     static_cast<int>(e.smem));
        if (err != cudaSuccess) return err;
      }
  for (int kind = 0; kind < 3; ++kind)
    for (int loss = 0; loss < 4; ++loss) {
      const cudaError_t err = cudaFuncSetAttribute(ResidentFor(kind, loss), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(kMaxDynamicSmem));
      if (err != cudaSuccess) return err;
    }
  return cudaSuccess;
}

// ------------------------------------------------------------------ state init / finish
__global__ void init_states_kernel(State* states, const double* poses16, int num_problems,
                                   int kind) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_problems) return;
  const double* P = poses16 + 16 * static_cast<size_t>(i);  // column-major 4x4
  State s;
  for (int k = 0; k < 9; ++k) s.R[k] = 0.0;
  for (int k = 0; k < 4; ++k) s.q[k] = 0.0;
  if (kind == kNdt3) {
    // ..._analytic_3dof.cc:22-24: top-left 2x2 and xy translation, no re-orthonormalisation
    s.t[0] = P[12]; s.t[1] = P[13]; s.t[2] = 0.0;
    s.R[0] = P[0]; s.R[1] = P[4]; s.R[2] = P[1]; s.R[3] = P[5];
  } else {
    // ..._analytic.cc:86-87: Orientation(initial_pose.rotation()) then toRotationMatrix()
    double Rin[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rin[3 * r + c] = P[4 * c + r];
    RotToQuat(Rin, s.q);
    QuatToRot(s.q, s.R);
    s.t[0] = P[12]; s.t[1] = P[13]; s.t[2] = P[14];
  }
  s.lambda = 0.001;
  s.previous_cost = DBL_MAX;
  s.iteration = 0;
  s.done = 0;
  s.status = 0;
  s.pad = 0;
  states[i] = s;
}

// results4: per problem {iterations, status, final_cost, unused}
__global__ void finish_states_kernel(const State* states, double* poses16, double* results4,
                                     int num_problems, int kind) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_problems) return;
  const State& s = states[i];
  double* P = poses16 + 16 * static_cast<size_t>(i);
  if (kind == kNdt3) {
    P[12] = s.t[0]; P[13] = s.t[1];
    P[0] = s.R[0]; P[4] = s.R[1]; P[1] = s.R[2]; P[5] = s.R[3];
  } else {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) P[4 * c + r] = s.R[3 * r + c];
    P[12] = s.t[0]; P[13] = s.t[1]; P[14] = s.t[2];
  }
  results4[4 * i + 0] = static_cast<double>(s.iteration);
  results4[4 * i + 1] = static_cast<double>(s.status);
  results4[4 * i + 2] = s.previous_cost;
  results4[4 * i + 3] = 0.0;
}

// A barrier of the ranks on the device: one exchange of a zero.  Enqueued in front of a sharded
// solve so that the ranks' loops (and the events around them) start together however far apart
// their hosts reached the launch.
__global__ void peer_rendezvous_kernel(const PeerComm pc) {
  __shared__ double zero[1];
  if (threadIdx.x == 0) zero[0] = 0.0;
  __syncthreads();
  PeerAllReduce(pc, zero, 1);
}

cudaError_t LaunchPeerRendezvous(const PeerComm& pc, cudaStream_t stream) {
  peer_rendezvous_kernel<<<1, 32, 0, stream>>>(pc);
  return cudaGetLastError();
}

cudaError_t LaunchInitStates(State* states, const double* poses16, int num_problems, int kind,
                             cudaStream_t stream) {
  const int threads = 128;
  init_states_kernel<<<(num_problems + threads - 1) / threads, threads, 0, stream>>>(
      states, poses16, num_problems, kind);
  return cudaGetLastError();
}

cudaError_t LaunchFinishStates(const State* states, double* poses16, double* results4,
                               int num_problems, int kind, cudaStream_t stream) {
  const int threads = 128;
  finish_states_kernel<<<(num_problems + threads - 1) / threads, threads, 0, stream>>>(
      states, poses16, results4, num_problems, kind);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ pack / unpack (AoS <-> planes)
struct PlanePtrs {
  double* p[kNdtPlanes];
};

// Element (plane k, correspondence i) of an NDT problem stored as ST.
template <typename ST>
__device__ __forceinline__ ST* NdtElem(double* plane0, int k, int64_t i) {
  return reinterpret_cast<ST*>(plane0) + TiledOffset(PlanesOf<kNdt6, ST>(), i) + k * kTile;
}

// Stores the information part of an NDT record from its sqrt_information S (row-major): the 6
// unique entries of L = S^T S, formed in fp64 and rounded once to the storage type.
template <typename ST>
__device__ __forceinline__ void StoreNdtInformation(double* plane0, int64_t i, const double* S) {
  double L[6];
  InformationFromSqrt(S, L);
#pragma unroll
  for (int k = 0; k < 6; ++k) *NdtElem<ST>(plane0, 6 + k, i) = static_cast<ST>(L[k]);
}

template <typename ST>
__global__ void pack_ndt_kernel(const double* __restrict__ point, const double* __restrict__ mean,
                                const double* __restrict__ sqrt_info, int64_t n, PlanePtrs planes,
                                int64_t dst_offset) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t d = dst_offset + i;
#pragma unroll
    for (int k = 0; k < 3; ++k) *NdtElem<ST>(planes.p[0], k, d) = static_cast<ST>(point[3 * i + k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) *NdtElem<ST>(planes.p[0], 3 + k, d) = static_cast<ST>(mean[3 * i + k]);
    double S[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) S[k] = sqrt_info[9 * i + k];
    StoreNdtInformation<ST>(planes.p[0], d, S);
  }
}

__global__ void pack_ndt_from_float_kernel(const float* __restrict__ point, const float* __restrict__ mean,
                                           const float* __restrict__ sqrt_info, int64_t n, PlanePtrs planes,
                                           int64_t dst_offset) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t d = dst_offset + i;
#pragma unroll
    for (int k = 0; k < 3; ++k) *NdtElem<float>(planes.p[0], k, d) = point[3 * i + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) *NdtElem<float>(planes.p[0], 3 + k, d) = mean[3 * i + k];
    double S[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) S[k] = static_cast<double>(sqrt_info[9 * i + k]);
    StoreNdtInformation<float>(planes.p[0], d, S);
  }
}

__global__ void pack_ndt_aos_kernel(const unsigned char* __restrict__ records, int64_t n,
                                    size_t stride_bytes, size_t off_point, size_t off_mean,
                                    size_t off_sqrt, int col_major, PlanePtrs planes, int64_t dst_offset) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned char* rec = records + static_cast<size_t>(i) * stride_bytes;
    const double* pt = reinterpret_cast<const double*>(rec + off_point);
    const double* mu = reinterpret_cast<const double*>(rec + off_mean);
    const double* S = reinterpret_cast<const double*>(rec + off_sqrt);
    const int64_t d = dst_offset + i;
#pragma unroll
    for (int k = 0; k < 3; ++k) *NdtElem<double>(planes.p[0], k, d) = pt[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) *NdtElem<double>(planes.p[0], 3 + k, d) = mu[k];
    double Srow[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) Srow[3 * r + c] = col_major ? S[3 * c + r] : S[3 * r + c];
    StoreNdtInformation<double>(planes.p[0], d, Srow);
  }
}

template <typename ST>
__global__ void unpack_ndt_kernel(PlanePtrs planes, int64_t begin, int64_t end, double* point,
                                  double* mean, double* sqrt_info) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = begin + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < end;
       i += stride) {
    const int64_t o = i - begin;
    for (int k = 0; k < 3; ++k) point[3 * o + k] = static_cast<double>(*NdtElem<ST>(planes.p[0], k, i));
    for (int k = 0; k < 3; ++k) mean[3 * o + k] = static_cast<double>(*NdtElem<ST>(planes.p[0], 3 + k, i));
    // third output: the 6 unique entries of the information matrix S^T S (00 01 02 11 12 22)
    for (int k = 0; k < 6; ++k) sqrt_info[6 * o + k] = static_cast<double>(*NdtElem<ST>(planes.p[0], 6 + k, i));
  }
}

__global__ void pack_reproj_kernel(const double* __restrict__ local_point,
                                   const double* __restrict__ pixel, int64_t n, PlanePtrs planes,
                                   int64_t dst_offset) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t d = TiledOffset(kReprojPlanes, dst_offset + i);
    planes.p[0][d] = local_point[3 * i];
    planes.p[1][d] = local_point[3 * i + 1];
    planes.p[2][d] = local_point[3 * i + 2];
    planes.p[3][d] = pixel[2 * i];
    planes.p[4][d] = pixel[2 * i + 1];
  }
}

static int GridFor(int64_t n, int threads) {
  int64_t blocks = (n + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

static PlanePtrs MakePlanes(double* const* planes, int count) {
  PlanePtrs pp;
  for (int k = 0; k < kNdtPlanes; ++k) pp.p[k] = (k < count) ? planes[k] : nullptr;
  return pp;
}

cudaError_t LaunchPackNdt(const double* point, const double* mean, const double* sqrt_info,
                          int64_t n, double* const planes[kNdtPlanes], int64_t dst_offset,
                          bool f32, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (f32)
    pack_ndt_kernel<float><<<GridFor(n, 256), 256, 0, stream>>>(point, mean, sqrt_info, n,
                                                                MakePlanes(planes, kNdtPlanes), dst_offset);
  else
    pack_ndt_kernel<double><<<GridFor(n, 256), 256, 0, stream>>>(point, mean, sqrt_info, n,
                                                                 MakePlanes(planes, kNdtPlanes), dst_offset);
  return cudaGetLastError();
}

cudaError_t LaunchPackNdt(const float* point, const float* mean, const float* sqrt_info, int64_t n,
                          double* const planes[kNdtPlanes], int64_t dst_offset, bool f32, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (!f32) return cudaErrorInvalidValue;  // float input exists for the fp32 storage mode only
  pack_ndt_from_float_kernel<<<GridFor(n, 256), 256, 0, stream>>>(point, mean, sqrt_info, n,
                                                                  MakePlanes(planes, kNdtPlanes), dst_offset);
  return cudaGetLastError();
}

cudaError_t LaunchPackNdtAos(const unsigned char* records, int64_t n, size_t stride,
                             size_t off_point, size_t off_mean, size_t off_sqrt, int col_major,
                             double* const planes[kNdtPlanes], int64_t dst_offset, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  pack_ndt_aos_kernel<<<GridFor(n, 256), 256, 0, stream>>>(
      records, n, stride, off_point, off_mean, off_sqrt, col_major, MakePlanes(planes, kNdtPlanes), dst_offset);
  return cudaGetLastError();
}

cudaError_t LaunchUnpackNdt(double* const planes[kNdtPlanes], int64_t begin, int64_t end,
                            double* point, double* mean, double* sqrt_info, bool f32,
                            cudaStream_t stream) {
  if (end <= begin) return cudaSuccess;
  if (f32)
    unpack_ndt_kernel<float><<<GridFor(end - begin, 256), 256, 0, stream>>>(
        MakePlanes(planes, kNdtPlanes), begin, end, point, mean, sqrt_info);
  else
    unpack_ndt_kernel<double><<<GridFor(end - begin, 256), 256, 0, stream>>>(
        MakePlanes(planes, kNdtPlanes), begin, end, point, mean, sqrt_info);
  return cudaGetLastError();
}

cudaError_t LaunchPackReproj(const double* local_point, const double* pixel, int64_t n,
                             double* const planes[kReprojPlanes], int64_t dst_offset, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  pack_reproj_kernel<<<GridFor(n, 256), 256, 0, stream>>>(local_point, pixel, n,
                                                          MakePlanes(planes, kReprojPlanes), dst_offset);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ synthetic scan + dense-grid association
__device__ __forceinline__ uint64_t SplitMix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double U01(uint64_t bits) {
  return static_cast<double>(bits >> 11) * (1.0 / 9007199254740992.0);
}

// Room of tests/simple_optimization_test.cc:170-204: floor 7x5 at z=0 and four walls, height 2.5.
// Surfaces are sampled proportionally to area: floor 35, two long walls 17.5 each, two short
// walls 12.5 each (total 95).
__global__ void generate_ndt_kernel(const GenerateParams g) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < g.n;
       i += stride) {
    const uint64_t gid = static_cast<uint64_t>(g.index_offset + i);
    double lx = 0, ly = 0, lz = 0, mean[3] = {0, 0, 0}, S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    {
      uint64_t h = SplitMix64(g.seed ^ SplitMix64(gid));
      const double u0 = U01(h); h = SplitMix64(h);
      const double u1 = U01(h); h = SplitMix64(h);
      const double u2 = U01(h); h = SplitMix64(h);
      const double n0 = U01(h); h = SplitMix64(h);
      const double n1 = U01(h); h = SplitMix64(h);
      const double n2 = U01(h); h = SplitMix64(h);
      const double n3 = U01(h);
      double wx, wy, wz;
      const double a = u0 * 95.0;
      if (a < 35.0) { wx = -3.5 + 7.0 * u1; wy = -2.5 + 5.0 * u2; wz = 0.0; }
      else if (a < 52.5) { wx = -3.5 + 7.0 * u1; wy = -2.5; wz = 2.5 * u2; }
      else if (a < 70.0) { wx = -3.5 + 7.0 * u1; wy = 2.5; wz = 2.5 * u2; }
      else if (a < 82.5) { wx = -3.5; wy = -2.5 + 5.0 * u1; wz = 2.5 * u2; }
      else { wx = 3.5; wy = -2.5 + 5.0 * u1; wz = 2.5 * u2; }
      // Box-Muller noise (3 of 4 values used)
      const double r0 = sqrt(-2.0 * log(fmax(n0, 1e-300))), r1 = sqrt(-2.0 * log(fmax(n2, 1e-300)));
      double s0, c0, s1, c1;
      sincospi(2.0 * n1, &s0, &c0);
      sincospi(2.0 * n3, &s1, &c1);
      wx += g.noise_sigma * r0 * c0;
      wy += g.noise_sigma * r0 * s0;
      wz += g.noise_sigma * r1 * c1;
      // sensor frame: l = R_true^T (w - t_true)
      const double dx = wx - g.t_true[0], dy = wy - g.t_true[1], dz = wz - g.t_true[2];
      lx = g.R_true[0] * dx + g.R_true[3] * dy + g.R_true[6] * dz;
      ly = g.R_true[1] * dx + g.R_true[4] * dy + g.R_true[7] * dz;
      lz = g.R_true[2] * dx + g.R_true[5] * dy + g.R_true[8] * dz;
      // association under the initial pose: the cell of the voxel containing R_init l + t_init,
      // else the nearest valid cell mean within 1.0 m (scan order z, y, x; first strict minimum)
      const double ix = g.R_init[0] * lx + g.R_init[1] * ly + g.R_init[2] * lz + g.t_init[0];
      const double iy = g.R_init[3] * lx + g.R_init[4] * ly + g.R_init[5] * lz + g.t_init[1];
      const double iz = g.R_init[6] * lx + g.R_init[7] * ly + g.R_init[8] * lz + g.t_init[2];
      const int cx = static_cast<int>(floor((ix - g.origin[0]) * g.inv_voxel));
      const int cy = static_cast<int>(floor((iy - g.origin[1]) * g.inv_voxel));
      const int cz = static_cast<int>(floor((iz - g.origin[2]) * g.inv_voxel));
      int cell = -1;
      if (cx >= 0 && cy >= 0 && cz >= 0 && cx < g.dims[0] && cy < g.dims[1] && cz < g.dims[2]) {
        const int own = (cz * g.dims[1] + cy) * g.dims[0] + cx;
        if (g.cell_valid[own]) cell = own;
      }
      if (cell < 0) {
        double best = 1.0;
        for (int oz = -g.reach; oz <= g.reach; ++oz)
          for (int oy = -g.reach; oy <= g.reach; ++oy)
            for (int ox = -g.reach; ox <= g.reach; ++ox) {
              const int x = cx + ox, y = cy + oy, z = cz + oz;
              if (x < 0 || y < 0 || z < 0 || x >= g.dims[0] || y >= g.dims[1] || z >= g.dims[2]) continue;
              const int c = (z * g.dims[1] + y) * g.dims[0] + x;
              if (!g.cell_valid[c]) continue;
              const double ex = ix - g.cell_mean[3 * c], ey = iy - g.cell_mean[3 * c + 1],
                           ez = iz - g.cell_mean[3 * c + 2];
              const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
              if (d2 < best) { best = d2; cell = c; }
            }
      }
      if (cell >= 0) {
        for (int k = 0; k < 3; ++k) mean[k] = g.cell_mean[3 * cell + k];
        for (int k = 0; k < 9; ++k) S[k] = g.cell_sqrt_info[9 * cell + k];
      }
    }
    // a point that never found a valid cell keeps S = 0 and contributes exactly nothing
    const int64_t d = g.dst_offset + i;
    if (g.f32) {
      *NdtElem<float>(g.planes[0], 0, d) = static_cast<float>(lx);
      *NdtElem<float>(g.planes[0], 1, d) = static_cast<float>(ly);
      *NdtElem<float>(g.planes[0], 2, d) = static_cast<float>(lz);
      for (int k = 0; k < 3; ++k) *NdtElem<float>(g.planes[0], 3 + k, d) = static_cast<float>(mean[k]);
      StoreNdtInformation<float>(g.planes[0], d, S);
    } else {
      *NdtElem<double>(g.planes[0], 0, d) = lx;
      *NdtElem<double>(g.planes[0], 1, d) = ly;
      *NdtElem<double>(g.planes[0], 2, d) = lz;
      for (int k = 0; k < 3; ++k) *NdtElem<double>(g.planes[0], 3 + k, d) = mean[k];
      StoreNdtInformation<double>(g.planes[0], d, S);
    }
  }
}

cudaError_t LaunchGenerateNdt(const GenerateParams& p, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  generate_ndt_kernel<<<GridFor(p.n, 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ voxel hash
__device__ __forceinline__ unsigned long long VoxelHashKey(int x, int y, int z) {
  return static_cast<unsigned long long>(x) | (static_cast<unsigned long long>(y) << kHashAxisBits) |
         (static_cast<unsigned long long>(z) << (2 * kHashAxisBits));
}
__device__ __forceinline__ long long VoxelHashHome(unsigned long long key, long long mask) {
  key ^= key >> 30; key *= 0xbf58476d1ce4e5b9ull;
  key ^= key >> 27; key *= 0x94d049bb133111ebull;
  key ^= key >> 31;
  return static_cast<long long>(key) & mask;
}
// Slot of `key`, claiming an empty one on first sight.  *fresh (nullable) = this call claimed it.
__device__ __forceinline__ long long VoxelHashInsert(unsigned long long* keys, long long mask, unsigned long long key,
                                                     bool* fresh) {
  long long s = VoxelHashHome(key, mask);
  while (true) {
    const unsigned long long seen = keys[s];
    if (seen == key) { if (fresh) *fresh = false; return s; }
    if (seen == kHashEmpty) {
      const unsigned long long prev = atomicCAS(keys + s, kHashEmpty, key);
      if (prev == kHashEmpty) { if (fresh) *fresh = true; return s; }
      if (prev == key) { if (fresh) *fresh = false; return s; }
    }
    s = (s + 1) & mask;
  }
}
// Slot of `key`, or -1.  Terminates because the table is never more than half full.
__device__ __forceinline__ long long VoxelHashFind(const unsigned long long* __restrict__ keys, long long mask,
                                                   unsigned long long key) {
  long long s = VoxelHashHome(key, mask);
  while (true) {
    const unsigned long long seen = keys[s];
    if (seen == key) return s;
    if (seen == kHashEmpty) return -1;
    s = (s + 1) & mask;
  }
}
// floor(v) as a voxel index, saturated so that far-away points cannot overflow the int cast.
__device__ __forceinline__ int VoxelIndex(double v) {
  return static_cast<int>(fmin(fmax(floor(v), -1073741824.0), 1073741824.0));
}

// ------------------------------------------------------------------ device NDT matcher
// Restates MatchPointCloud (mahalanobis_distance_minimizer/tests/simple_optimization_test.cc:296-342):
// warp the point by the pose, take the (at most) two nearest valid cell means within the search
// radius (flann radiusSearch with L2_Simple: squared distance < radius), emit one correspondence
// per hit carrying the cell's mean and sqrt_information.  Instead of a KD-tree the dense voxel grid
// (or the voxel hash of a sparse map, probed once per neighbouring voxel in the same z, y, x order)
// is scanned `reach` cells around the point (a mean lies inside its own voxel, so every mean within
// the radius is visited).  A missing neighbour becomes a zero-information record (exact zero
// contribution), which keeps the output a fixed 2 x n layout with no compaction pass.
__global__ void match_ndt_kernel(const MatchParams m) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  unsigned long long local_matched = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < m.n; i += stride) {
    const double lx = m.scan[0][i], ly = m.scan[1][i], lz = m.scan[2][i];
    const double wx = m.R[0] * lx + m.R[1] * ly + m.R[2] * lz + m.t[0];
    const double wy = m.R[3] * lx + m.R[4] * ly + m.R[5] * lz + m.t[1];
    const double wz = m.R[6] * lx + m.R[7] * ly + m.R[8] * lz + m.t[2];
    const int cx = VoxelIndex((wx - m.origin[0]) * m.inv_voxel);
    const int cy = VoxelIndex((wy - m.origin[1]) * m.inv_voxel);
    const int cz = VoxelIndex((wz - m.origin[2]) * m.inv_voxel);
    int best[2] = {-1, -1};
    double best_d2[2] = {m.radius2, m.radius2};
    for (int oz = -m.reach; oz <= m.reach; ++oz) {
      const int z = cz + oz;
      if (z < 0 || z >= m.dims[2]) continue;
      for (int oy = -m.reach; oy <= m.reach; ++oy) {
        const int y = cy + oy;
        if (y < 0 || y >= m.dims[1]) continue;
        for (int ox = -m.reach; ox <= m.reach; ++ox) {
          const int x = cx + ox;
          if (x < 0 || x >= m.dims[0]) continue;
          int c;
          if (m.keys != nullptr) {
            c = static_cast<int>(VoxelHashFind(m.keys, m.hash_mask, VoxelHashKey(x, y, z)));
            if (c < 0) continue;
          } else {
            c = (z * m.dims[1] + y) * m.dims[0] + x;
          }
          if (!m.cell_valid[c]) continue;
          const double ex = wx - m.cell_mean[3 * c], ey = wy - m.cell_mean[3 * c + 1],
                       ez = wz - m.cell_mean[3 * c + 2];
          const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
          if (d2 < best_d2[0]) {
            best_d2[1] = best_d2[0]; best[1] = best[0];
            best_d2[0] = d2; best[0] = c;
          } else if (d2 < best_d2[1]) {
            best_d2[1] = d2; best[1] = c;
          }
        }
      }
    }
    for (int j = 0; j < m.max_neighbors; ++j) {
      const int64_t o = static_cast<int64_t>(j) * m.n + i;
      const int c = best[j];
      *NdtElem<double>(m.planes[0], 0, o) = lx;
      *NdtElem<double>(m.planes[0], 1, o) = ly;
      *NdtElem<double>(m.planes[0], 2, o) = lz;
      double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      double mu[3] = {0, 0, 0};
      if (c >= 0) {
        for (int k = 0; k < 3; ++k) mu[k] = m.cell_mean[3 * c + k];
        for (int k = 0; k < 9; ++k) S[k] = m.cell_sqrt_info[9 * c + k];
        ++local_matched;
      }
      for (int k = 0; k < 3; ++k) *NdtElem<double>(m.planes[0], 3 + k, o) = mu[k];
      StoreNdtInformation<double>(m.planes[0], o, S);
    }
  }
  if (m.matched != nullptr) {
    for (int off = 16; off > 0; off >>= 1) local_matched += __shfl_xor_sync(0xffffffffu, local_matched, off);
    if ((threadIdx.x & 31) == 0 && local_matched) atomicAdd(m.matched, local_matched);
  }
}

cudaError_t LaunchMatchNdt(const MatchParams& p, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  match_ndt_kernel<<<GridFor(p.n, 128), 128, 0, stream>>>(p);
  return cudaGetLastError();
}

// The planar minimizer of the reference processes floor(M / 4) * 4 of the M correspondences it is
// given (..._analytic_3dof.cc:33-36); MatchPointCloud emits them point-major (point 0's hits, nearest
// first, then point 1's, ...; tests/simple_optimization_test.cc:318-340).  The device matcher keeps a
// fixed slot layout (slot j of point i at j * n + i, empty slots with zero information), so the
// M mod 4 hits the reference would drop are the LAST real hits in point-major order: walk back from
// (n - 1, last slot) and zero their information (exact zero contribution).  One thread: it stops
// after at most 3 hits.
__global__ void truncate_ndt3_hits_kernel(PlanePtrs planes, int64_t n, int max_neighbors,
                                          const unsigned long long* matched) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int drop = static_cast<int>(*matched % 4ULL);
  for (int64_t i = n - 1; i >= 0 && drop > 0; --i)
    for (int j = max_neighbors - 1; j >= 0 && drop > 0; --j) {
      const int64_t o = static_cast<int64_t>(j) * n + i;
      // a real hit carries a valid cell's information matrix, whose trace is positive
      const double trace = *NdtElem<double>(planes.p[0], 6, o) + *NdtElem<double>(planes.p[0], 9, o) +
                           *NdtElem<double>(planes.p[0], 11, o);
      if (trace > 0.0) {
        for (int k = 0; k < 6; ++k) *NdtElem<double>(planes.p[0], 6 + k, o) = 0.0;
        --drop;
      }
    }
}
cudaError_t LaunchTruncateNdt3Hits(double* const planes[kNdtPlanes], int64_t n, int max_neighbors,
                                   const unsigned long long* matched, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  truncate_ndt3_hits_kernel<<<1, 32, 0, stream>>>(MakePlanes(planes, kNdtPlanes), n, max_neighbors, matched);
  return cudaGetLastError();
}

__global__ void pack_scan_kernel(const double* __restrict__ xyz, int64_t n, double* px, double* py, double* pz) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    px[i] = xyz[3 * i]; py[i] = xyz[3 * i + 1]; pz[i] = xyz[3 * i + 2];
  }
}
cudaError_t LaunchPackScan(const double* xyz, int64_t n, double* const planes[3], cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  pack_scan_kernel<<<GridFor(n, 256), 256, 0, stream>>>(xyz, n, planes[0], planes[1], planes[2]);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ device NDT map builder
// UpdateNdtMap (tests/simple_optimization_test.cc:236-280) on a dense voxel grid.
__global__ void map_bounds_kernel(const double* __restrict__ xyz, int64_t n, double inv_voxel, int* bounds6) {
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    for (int a = 0; a < 3; ++a) {
      const int k = static_cast<int>(floor(xyz[3 * i + a] * inv_voxel));
      lo[a] = min(lo[a], k);
      hi[a] = max(hi[a], k);
    }
  for (int a = 0; a < 3; ++a) {
    for (int off = 16; off > 0; off >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(bounds6 + a, lo[a]);
      atomicMax(bounds6 + 3 + a, hi[a]);
    }
  }
}

// Sums are taken about the voxel centre: the covariance is shift-invariant, and the reference's
// raw-moment form (moment / n - mean mean^T, :255-259) loses 2 log10(|p| / voxel) digits to
// cancellation for maps far from the origin.
//
// The sums are INTEGERS: an offset d = (p - centre) / voxel lies in [-1/2, 1/2], so d and d d^T are
// accumulated as 64-bit fixed point with `fixed_shift` fractional bits (40 for up to 2 M points:
// a resolution of 1e-12 of the voxel, far below the noise of any scan).  Integer addition is
// associative, so the map -- and with it every registration against it -- is bit-for-bit the same
// on every run, whatever order the atomics land in; fp64 atomics are not.  Contention: scan points
// arrive in sweep order, so the lanes of a warp mostly share a voxel; a warp whose 32 lanes agree
// reduces its ten values by shuffle and issues ten atomics instead of 320.
__device__ __forceinline__ long long ToFixed(double v, int shift) {
  return __double2ll_rn(scalbn(v, shift));
}

__global__ void map_accumulate_kernel(const MapAccumParams m) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  // whole warps iterate together (the tail is masked), so that the shuffles below are converged
  for (int64_t base = first - lane; base < m.n; base += stride) {
    const int64_t i = base + lane;
    const bool live = i < m.n;
    int64_t c = -1;
    long long v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (live) {
      const int ax = static_cast<int>(floor(m.xyz[3 * i] * m.inv_voxel));
      const int ay = static_cast<int>(floor(m.xyz[3 * i + 1] * m.inv_voxel));
      const int az = static_cast<int>(floor(m.xyz[3 * i + 2] * m.inv_voxel));
      const double x = (m.xyz[3 * i] - (ax + 0.5) * m.voxel) * m.inv_voxel;
      const double y = (m.xyz[3 * i + 1] - (ay + 0.5) * m.voxel) * m.inv_voxel;
      const double z = (m.xyz[3 * i + 2] - (az + 0.5) * m.voxel) * m.inv_voxel;
      const int kx = ax - m.kmin[0], ky = ay - m.kmin[1], kz = az - m.kmin[2];
      c = m.keys != nullptr ? VoxelHashInsert(m.keys, m.hash_mask, VoxelHashKey(kx, ky, kz), nullptr)
                            : (static_cast<int64_t>(kz) * m.dims[1] + ky) * m.dims[0] + kx;
      const int sh = m.fixed_shift;
      v[0] = ToFixed(x, sh); v[1] = ToFixed(y, sh); v[2] = ToFixed(z, sh);
      v[3] = ToFixed(x * x, sh); v[4] = ToFixed(x * y, sh); v[5] = ToFixed(x * z, sh);
      v[6] = ToFixed(y * y, sh); v[7] = ToFixed(y * z, sh); v[8] = ToFixed(z * z, sh);
    }
    const int64_t c0 = __shfl_sync(0xffffffffu, c, 0);
    const bool uniform = __all_sync(0xffffffffu, c == c0 && live);
    if (uniform) {
#pragma unroll
      for (int k = 0; k < 9; ++k)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
      if (lane == 0) {
        atomicAdd(m.count + c, 32);
        unsigned long long* s = reinterpret_cast<unsigned long long*>(m.sums) + 9 * c;
#pragma unroll
        for (int k = 0; k < 9; ++k) atomicAdd(s + k, static_cast<unsigned long long>(v[k]));
      }
    } else if (live) {
      atomicAdd(m.count + c, 1);
      unsigned long long* s = reinterpret_cast<unsigned long long*>(m.sums) + 9 * c;
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(s + k, static_cast<unsigned long long>(v[k]));
    }
  }
}

__global__ void map_count_voxels_kernel(const double* __restrict__ xyz, int64_t n, double inv_voxel, int kx0, int ky0,
                                        int kz0, unsigned long long* keys, long long mask,
                                        unsigned long long* distinct) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  unsigned long long fresh_here = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int kx = static_cast<int>(floor(xyz[3 * i] * inv_voxel)) - kx0;
    const int ky = static_cast<int>(floor(xyz[3 * i + 1] * inv_voxel)) - ky0;
    const int kz = static_cast<int>(floor(xyz[3 * i + 2] * inv_voxel)) - kz0;
    bool fresh = false;
    VoxelHashInsert(keys, mask, VoxelHashKey(kx, ky, kz), &fresh);
    fresh_here += fresh ? 1 : 0;
  }
  for (int off = 16; off > 0; off >>= 1) fresh_here += __shfl_xor_sync(0xffffffffu, fresh_here, off);
  if ((threadIdx.x & 31) == 0 && fresh_here) atomicAdd(distinct, fresh_here);
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (row-major a[9]); eigenvalues ascending in
// w, eigenvectors in the COLUMNS of V.  Each eigenvector is sign-normalised so that its
// largest-magnitude component is positive (the decomposition is otherwise unique only up to sign).
__device__ inline void SymmetricEigen3(const double* a_in, double* w, double* V) {
  double a[9];
  for (int i = 0; i < 9; ++i) { a[i] = a_in[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    const double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[3 * p + q];
        if (apq == 0.0) continue;
        const double theta = (a[3 * q + q] - a[3 * p + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[3 * k + p], akq = a[3 * k + q];
          a[3 * k + p] = c * akp - s * akq;
          a[3 * k + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[3 * p + k], aqk = a[3 * q + k];
          a[3 * p + k] = c * apk - s * aqk;
          a[3 * q + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[3 * k + p], vkq = V[3 * k + q];
          V[3 * k + p] = c * vkp - s * vkq;
          V[3 * k + q] = s * vkp + c * vkq;
        }
      }
  }
  w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (w[j] > w[j + 1]) {
        const double tw = w[j]; w[j] = w[j + 1]; w[j + 1] = tw;
        for (int k = 0; k < 3; ++k) { const double tv = V[3 * k + j]; V[3 * k + j] = V[3 * k + j + 1]; V[3 * k + j + 1] = tv; }
      }
  for (int j = 0; j < 3; ++j) {
    int big = 0;
    for (int k = 1; k < 3; ++k) if (fabs(V[3 * k + j]) > fabs(V[3 * big + j])) big = k;
    if (V[3 * big + j] < 0.0) for (int k = 0; k < 3; ++k) V[3 * k + j] = -V[3 * k + j];
  }
}

__global__ void map_finalize_kernel(const MapAccumParams m, int64_t cells, int v_not_transposed, double* cell_mean,
                                    double* cell_sqrt_info, unsigned char* cell_valid) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  double mean[3] = {0, 0, 0}, S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  unsigned char valid = 0;
  const int n = m.count[c];
  if (n >= 5) {  // :250-253
    // voxel indices of this cell, from the slot key or the linear index
    int64_t idx[3];
    if (m.keys != nullptr) {
      const unsigned long long key = m.keys[c];
      const unsigned long long axis_mask = (1ull << kHashAxisBits) - 1ull;
      idx[0] = static_cast<int64_t>(key & axis_mask);
      idx[1] = static_cast<int64_t>((key >> kHashAxisBits) & axis_mask);
      idx[2] = static_cast<int64_t>((key >> (2 * kHashAxisBits)) & axis_mask);
    } else {
      idx[0] = c % m.dims[0];
      idx[1] = (c / m.dims[0]) % m.dims[1];
      idx[2] = c / (static_cast<int64_t>(m.dims[0]) * m.dims[1]);
    }
    const double inv_n = 1.0 / n;
    // fixed-point sums of d = (p - centre) / voxel and d d^T -> metres, metres^2
    const long long* fixed = reinterpret_cast<const long long*>(m.sums) + 9 * c;
    double s[9];
    for (int k = 0; k < 9; ++k)
      s[k] = scalbn(static_cast<double>(fixed[k]), -m.fixed_shift) * (k < 3 ? m.voxel : m.voxel * m.voxel);
    double off[3];  // mean relative to the voxel centre
    for (int a = 0; a < 3; ++a) {
      off[a] = s[a] * inv_n;
      mean[a] = (static_cast<double>(idx[a] + m.kmin[a]) + 0.5) * m.voxel + off[a];
    }
    // moment starts at Identity (types.h:14): cov = (I + sum p p^T) / n - mean mean^T, evaluated
    // about the voxel centre (same value, no cancellation)
    double cov[9];
    cov[0] = (s[3] + 1.0) * inv_n - off[0] * off[0];
    cov[1] = cov[3] = s[4] * inv_n - off[0] * off[1];
    cov[2] = cov[6] = s[5] * inv_n - off[0] * off[2];
    cov[4] = (s[6] + 1.0) * inv_n - off[1] * off[1];
    cov[5] = cov[7] = s[7] * inv_n - off[1] * off[2];
    cov[8] = (s[8] + 1.0) * inv_n - off[2] * off[2];
    double w[3], V[9];
    SymmetricEigen3(cov, w, V);
    if (w[2] >= 0.01) {  // :263
      valid = 1;
      w[0] = fmax(w[0], 0.01 * w[2]);  // :271-272
      w[1] = fmax(w[1], 0.01 * w[2]);
      for (int r = 0; r < 3; ++r) {
        const double d = 1.0 / sqrt(w[r]);
        for (int col = 0; col < 3; ++col)  // diag * V as the reference writes it (:275-276), or diag * V^T
          S[3 * r + col] = d * (v_not_transposed ? V[3 * r + col] : V[3 * col + r]);
      }
    } else {
      mean[0] = mean[1] = mean[2] = 0.0;
    }
  }
  for (int a = 0; a < 3; ++a) cell_mean[3 * c + a] = valid ? mean[a] : 0.0;
  for (int k = 0; k < 9; ++k) cell_sqrt_info[9 * c + k] = S[k];
  cell_valid[c] = valid;
}

cudaError_t LaunchMapBounds(const double* xyz, int64_t n, double inv_voxel, int* bounds6, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  map_bounds_kernel<<<GridFor(n, 256), 256, 0, stream>>>(xyz, n, inv_voxel, bounds6);
  return cudaGetLastError();
}
cudaError_t LaunchMapAccumulate(const MapAccumParams& p, cudaStream_t stream) {
  if (p.n <= 0) return cudaSuccess;
  map_accumulate_kernel<<<GridFor(p.n, 256), 256, 0, stream>>>(p);
  return cudaGetLastError();
}
cudaError_t LaunchMapCountVoxels(const double* xyz, int64_t n, double inv_voxel, const int kmin[3],
                                 unsigned long long* keys, long long hash_mask, unsigned long long* distinct,
                                 cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  map_count_voxels_kernel<<<GridFor(n, 256), 256, 0, stream>>>(xyz, n, inv_voxel, kmin[0], kmin[1], kmin[2], keys,
                                                               hash_mask, distinct);
  return cudaGetLastError();
}
cudaError_t LaunchMapFinalize(const MapAccumParams& p, int64_t cells, int v_not_transposed, double* cell_mean,
                              double* cell_sqrt_info, unsigned char* cell_valid, cudaStream_t stream) {
  if (cells <= 0) return cudaSuccess;
  map_finalize_kernel<<<static_cast<int>((cells + 127) / 128), 128, 0, stream>>>(
      p, cells, v_not_transposed, cell_mean, cell_sqrt_info, cell_valid);
  return cudaGetLastError();
}

}  // namespace nlo
