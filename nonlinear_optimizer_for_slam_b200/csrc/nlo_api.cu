// nlo_api.cu -- the C ABI of include/nlo_cuda.h on top of the kernels in nlo_kernels.cu: contexts,
// problems, one-pass assembly, the device-resident solves and the communicators.
//
// Host-side responsibilities only: device memory for the SoA planes and the per-registration
// state, CUDA-graph capture of the device-resident iteration loop, the optional communicators
// (NCCL through dlopen, or peer memory -- CUDA IPC between processes, direct mapping inside one
// process -- for the fused one-shot all-reduce), and the final pose / iteration-count read-back.
// Ingest lives in nlo_ingest.cu, the map / matcher / outer loop in nlo_map.cu, the multi-device
// dispatcher in nlo_multi.cu.  No arithmetic of the hot path runs on the host and there is no CPU
// fallback.
#include <dlfcn.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "nlo_host.h"

using namespace nlo;

namespace {

constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum

bool LoadNccl(NcclApi* api, std::string* err) {
  if (api->handle != nullptr) return true;
  const char* candidates[] = {"libnccl.so.2", "libnccl.so",
                              "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
  void* h = nullptr;
  for (const char* c : candidates) {
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h != nullptr) break;
  }
  if (h == nullptr) {
    *err = std::string("dlopen(libnccl) failed: ") + dlerror();
    return false;
  }
  api->handle = h;
  api->GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(dlsym(h, "ncclGetUniqueId"));
  api->CommInitRank =
      reinterpret_cast<int (*)(NcclComm*, int, NcclUniqueId, int)>(dlsym(h, "ncclCommInitRank"));
  api->AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, NcclComm,
                                            cudaStream_t)>(dlsym(h, "ncclAllReduce"));
  api->CommDestroy = reinterpret_cast<int (*)(NcclComm)>(dlsym(h, "ncclCommDestroy"));
  api->GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(h, "ncclGetErrorString"));
  if (!api->GetUniqueId || !api->CommInitRank || !api->AllReduce || !api->CommDestroy) {
    *err = "libnccl is missing a required symbol";
    return false;
  }
  return true;
}

int64_t PadToTile(int64_t n) { return ((n + kTile - 1) / kTile) * kTile; }

// The communicator a call on this problem runs under: batched registrations are independent
// (sharded by problem id, no collective), and a suspended communicator is ignored.
int CommFor(const nlo_context* ctx, const nlo_problem* pr) {
  return pr->batched ? static_cast<int>(kCommNone) : ctx->EffectiveComm();
}

IterParams BaseParams(nlo_context* ctx, nlo_problem* pr) {
  IterParams p;
  memset(&p, 0, sizeof(p));
  for (int k = 0; k < pr->num_planes; ++k) p.planes[k] = pr->planes[k];
  p.partials = pr->d_partials;
  p.ll_partials = pr->d_ll_partials;
  p.ll_sums = pr->d_sync;
  p.loss_p0 = ctx->loss_params[0];
  p.loss_p1 = ctx->loss_params[1];
  for (int k = 0; k < 6; ++k) p.intrinsics[k] = pr->intrinsics[k];
  p.parameter_tolerance = 1e-6;
  p.gradient_tolerance = 1e-6;
  p.max_iterations = 1;
  p.iterations_in_kernel = 1;
  p.f32 = pr->f32 ? 1 : 0;
  p.stage_depth = ctx->stage_depth;
  p.debug_times = ctx->d_debug_times;
  p.debug_all_ctas = ctx->debug_all_ctas ? 1 : 0;
  p.use_peer = (CommFor(ctx, pr) == kCommPeer) ? 1 : 0;
  p.peer = ctx->peer;
  return p;
}

bool UsePersistent(const nlo_context* ctx, const nlo_problem* pr) {
  return ctx->use_persistent && CommFor(ctx, pr) != kCommNccl && !pr->batched;
}

int GridFor(nlo_context* ctx, int64_t begin, int64_t end) {
  const int64_t tiles = (end + kTile - 1) / kTile - begin / kTile;
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ctx->grid_single, tiles)));
}

int NcclAllReduceSums(nlo_context* ctx, double* sums) {
  const int rc = ctx->nccl.AllReduce(sums, sums, 32, kNcclFloat64, kNcclSum, ctx->nccl_comm, ctx->stream);
  if (rc != 0)
    return Fail(ctx, NLO_ECOMM,
                std::string("ncclAllReduce: ") + (ctx->nccl.GetErrorString ? ctx->nccl.GetErrorString(rc) : "error"));
  return NLO_OK;
}

int CheckPeerError(nlo_context* ctx, const nlo_problem* pr) {
  if (CommFor(ctx, pr) != kCommPeer) return NLO_OK;
  int err = 0;
  NLO_CUDA(ctx, cudaMemcpy(&err, ctx->d_peer_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (err != 0) {
    // the flag is sticky on the device: clear it so that the next call starts clean
    cudaMemset(ctx->d_peer_error, 0, sizeof(int));
    return Fail(ctx, NLO_ECOMM, "peer all-reduce timed out waiting for a rank");
  }
  return NLO_OK;
}

// Shape of a persistent launch: CTAs per registration (a multiple of the cluster size), CTAs per
// thread-block cluster, and whether every CTA gathers the cluster partials itself.
struct GridPlan {
  int gx = 1;
  int cluster = 1;
  int direct = 0;
};

// CTAs of this kernel variant that are co-resident in clusters of `cluster` (cached per context; a
// device shared by several sub-contexts of one multi-device context is divided between them).
int CoResidentCtas(nlo_context* ctx, const nlo_problem* pr, int kind, int cluster, bool resident) {
  const int key = (((kind * 8 + ctx->loss_kind) * 2 + (pr->f32 ? 1 : 0)) * 16 + cluster) * 2 + (resident ? 1 : 0);
  auto it = ctx->coresident.find(key);
  if (it == ctx->coresident.end())
    it = ctx->coresident.emplace(key, MaxCoResidentCtas(kind, ctx->loss_kind, pr->f32, cluster, resident)).first;
  return it->second / std::max(1, ctx->device_share);
}

// `want` = CTAs per registration the caller would like (before the tile count and the cluster
// granularity are taken into account), `rows` = registrations of the launch (gridDim.y).
GridPlan PlanPersistentGrid(nlo_context* ctx, const nlo_problem* pr, int kind, int64_t tiles, int rows, int want,
                            bool small, bool resident = false) {
  GridPlan plan;
  const int64_t ctas = std::max<int64_t>(1, std::min<int64_t>(tiles, want));
  const int64_t tiles_per_cta = (tiles + ctas - 1) / ctas;
  (void)small;
  for (int cluster = ctx->cluster_small; cluster >= 1; cluster >>= 1) {
    if (cluster > 1 && cluster > ctas) continue;  // no point in clusters of mostly idle CTAs
    const int cap = (CoResidentCtas(ctx, pr, kind, cluster, resident) / std::max(1, rows) / cluster) * cluster;
    if (cap < cluster) continue;
    // idle CTAs in the last cluster are fine; a GPC that does not hold a whole number of clusters
    // strands SMs, which is only acceptable while no CTA gets (noticeably) more tiles for it
    const int gx = static_cast<int>(std::min<int64_t>((ctas + cluster - 1) / cluster * cluster, cap));
    const int64_t with_cluster = (tiles + gx - 1) / gx;
    if (cluster == 1 || with_cluster <= tiles_per_cta + tiles_per_cta / 32) {
      plan.gx = gx;
      plan.cluster = cluster;
      break;
    }
  }
  if (plan.gx < 1) plan.gx = 1;
  if (ctx->d_debug_times != nullptr)
    fprintf(stderr, "[nlo debug] plan: %lld tiles, want %d CTAs x %d rows -> %d CTAs in clusters of %d (%s kernel)\n",
            static_cast<long long>(tiles), want, rows, plan.gx, plan.cluster, resident ? "resident" : "streaming");
  plan.direct = (plan.gx > 1 && plan.gx / plan.cluster <= ctx->direct_max_clusters) ? 1 : 0;
  return plan;
}

// Tag base of the next resident launch on this problem: the solve epoch in bits 16..30, so
// that LL words left by earlier solves never match.  When the epoch wraps the buffers are cleared.
int NextEpoch(nlo_context* ctx, nlo_problem* pr, unsigned int* tag_base) {
  pr->epoch += 1;
  if (pr->epoch > 32767u) {  // 15 bits: bit 31 of a tag marks a failed wait
    pr->epoch = 1;
    NLO_CUDA(ctx, cudaMemsetAsync(pr->d_ll_partials, 0, pr->ll_partials_bytes, ctx->stream));
    NLO_CUDA(ctx, cudaMemsetAsync(pr->d_sync, 0, static_cast<size_t>(pr->num_problems + 1) * kSyncStride * sizeof(unsigned long long),
                                  ctx->stream));
  }
  *tag_base = pr->epoch << 16;
  return NLO_OK;
}

// Enqueue the device-resident loop of one solve on the context stream.
int EnqueueLoop(nlo_context* ctx, nlo_problem* pr, int kind, const nlo_solve_options& opt,
                bool with_trace, int num_problems, int64_t begin_abs, int64_t end_abs) {
  IterParams p = BaseParams(ctx, pr);
  p.ranges = pr->d_ranges;
  p.states = pr->d_states;
  p.tickets = pr->d_tickets;
  p.sums = pr->d_sums;
  p.trace = with_trace ? pr->d_trace : nullptr;
  p.parameter_tolerance = opt.parameter_tolerance;
  p.gradient_tolerance = opt.gradient_tolerance;
  p.max_iterations = opt.max_iterations;
  const int comm = CommFor(ctx, pr);
  const int64_t tiles = (end_abs + kTile - 1) / kTile - begin_abs / kTile;
  const bool in_cta_loop = (comm == kCommNone) && (pr->batched || tiles <= kInCtaTiles);
  const bool tags_fit = opt.max_iterations < 65535;  // resident kernel: the iteration number shares the 32-bit LL tag with the solve epoch
  if (pr->batched && ctx->use_persistent && tiles > kInCtaTiles && 2 * num_problems <= ctx->grid_single) {
    // A small batch: one CTA per registration would leave most SMs idle, so every registration
    // gets G = (2 x SMs) / B CTAs of ONE persistent cooperative grid (gridDim.y = registrations),
    // each registration with its own leader CTA, arrival counter and published state.
    const int gx = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ctx->grid_single / num_problems, tiles)));
    {
      IterParams q = p;
      q.mode = kModeSolve;
      q.persistent = 1;
      q.iterations_in_kernel = opt.max_iterations;
      NLO_CUDA(ctx, cudaMemsetAsync(pr->d_sync, 0, static_cast<size_t>(num_problems) * kSyncStride * sizeof(unsigned long long),
                                    ctx->stream));
      const cudaError_t ce = LaunchIteration(kind, ctx->loss_kind, q, gx, num_problems, 1, ctx->stream);
      if (ce == cudaSuccess) return NLO_OK;
      if (ce != cudaErrorCooperativeLaunchTooLarge && ce != cudaErrorLaunchOutOfResources)
        return Fail(ctx, NLO_ECUDA, std::string("cooperative launch: ") + cudaGetErrorString(ce));
      cudaGetLastError();  // not co-resident right now: fall through to one CTA per registration
    }
  }
  if (in_cta_loop && !pr->batched && ctx->use_persistent && ctx->use_resident && tags_fit && !pr->f32) {
    // a registration of <= 3 tiles: the resident kernel as a single CTA (its step and rotation are
    // the fast, fully inlined ones; nothing is exchanged)
    IterParams q = p;
    q.mode = kModeSolve;
    q.persistent = 1;
    q.iterations_in_kernel = opt.max_iterations;
    q.gather_direct = 1;
    const int rc = NextEpoch(ctx, pr, &q.tag_base);
    if (rc != NLO_OK) return rc;
    const cudaError_t ce = LaunchResident(kind, ctx->loss_kind, q, 1, 1, 1, static_cast<int>(std::max<int64_t>(tiles, 1)), ctx->stream);
    if (ce == cudaSuccess) return NLO_OK;
    if (ce != cudaErrorCooperativeLaunchTooLarge && ce != cudaErrorLaunchOutOfResources)
      return Fail(ctx, NLO_ECUDA, std::string("cooperative launch (resident, one CTA): ") + cudaGetErrorString(ce));
    cudaGetLastError();
  }
  if (in_cta_loop) {
    // whole loop inside one CTA per registration: a single launch
    p.mode = kModeSolve;
    p.iterations_in_kernel = opt.max_iterations;
    NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, 1, num_problems, 1, ctx->stream));
    return NLO_OK;
  }
  const int grid_x = GridFor(ctx, begin_abs, end_abs);
  if (UsePersistent(ctx, pr) && tags_fit && ctx->use_resident && !pr->f32) {
    // Latency-bound registration: if its tiles fit the shared memory of one CTA per SM they are
    // loaded once and the whole loop runs in the resident kernel.
    const GridPlan plan = PlanPersistentGrid(ctx, pr, kind, tiles, 1, ctx->sm_count / std::max(1, ctx->device_share), true, true);
    const int64_t stages = (tiles + plan.gx - 1) / plan.gx;
    if (stages <= ResidentMaxStages(kind)) {
      IterParams q = p;
      q.mode = kModeSolve;
      q.persistent = 1;
      q.iterations_in_kernel = opt.max_iterations;
      q.gather_direct = (comm == kCommNone) ? plan.direct : 0;
      const int rc = NextEpoch(ctx, pr, &q.tag_base);
      if (rc != NLO_OK) return rc;
      const cudaError_t ce = LaunchResident(kind, ctx->loss_kind, q, plan.gx, 1, plan.cluster, static_cast<int>(stages), ctx->stream);
      if (ce == cudaSuccess) return NLO_OK;
      if (ce != cudaErrorCooperativeLaunchTooLarge && ce != cudaErrorLaunchOutOfResources)
        return Fail(ctx, NLO_ECUDA, std::string("cooperative launch (resident): ") + cudaGetErrorString(ce));
      cudaGetLastError();  // not co-resident right now: the streaming kernel below
    }
  }
  if (UsePersistent(ctx, pr)) {
    // persistent cooperative grid: the whole loop in ONE launch.  An empty shard (a rank that owns
    // no points) still launches one CTA: it takes part in the all-reduce with zero sums.
    int gx = grid_x;
    if (tiles < 4LL * ctx->grid_single)
      gx = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(tiles, ctx->grid_small)));
    // One 512-thread CTA per SM with two tiles per ring stage instead of two 256-thread CTAs (see the
    // kernel): a single tile stream per SM, half the partials.  fp64 storage only.
    int wg = 1;
    if (ctx->warp_groups == 2 && !pr->f32 && tiles >= 2) {
      wg = 2;
      const int64_t stages_total = (tiles + 1) / 2;
      gx = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(stages_total, ctx->grid_small)));
    }
    p.mode = kModeSolve;
    p.persistent = 1;
    p.iterations_in_kernel = opt.max_iterations;
    // Streaming a large scan runs fastest with 3 tiles in flight per CTA (measured on B200, fp64 NDT,
    // 64M points: depth 2 / 3 / 4 = 71.3 / 75.9 / 71.3 Gpoints/s); the batched one-CTA-per-registration
    // shape keeps all 4 allocated stages.
    if (ctx->stage_depth == 0 && pr->family == 0 && !pr->f32) p.stage_depth = 3;
    {
      const double tile_mb = static_cast<double>(pr->num_planes) * kTile * (pr->f32 ? 4.0 : 8.0) / 1.0e6;
      const double scan_mb = static_cast<double>(tiles) * tile_mb;
      if (ctx->l2_keep_mb > 0.0 && scan_mb > ctx->l2_policy_min_mb && opt.max_iterations > 1)
        p.l2_keep_tiles = static_cast<long long>(ctx->l2_keep_mb / tile_mb);
    }
    NLO_CUDA(ctx, cudaMemsetAsync(pr->d_sync, 0, kSyncStride * sizeof(unsigned long long), ctx->stream));
    const cudaError_t ce = LaunchIteration(kind, ctx->loss_kind, p, gx, 1, 1, ctx->stream, wg);
    if (ce == cudaSuccess) return NLO_OK;
    if (ce != cudaErrorCooperativeLaunchTooLarge && ce != cudaErrorLaunchOutOfResources)
      return Fail(ctx, NLO_ECUDA, std::string("cooperative launch: ") + cudaGetErrorString(ce));
    // the grid cannot be co-resident right now (e.g. the GPU is shared): one launch per iteration
    cudaGetLastError();
    p.persistent = 0;
    p.iterations_in_kernel = 1;
    p.l2_keep_tiles = 0;
  }
  // One launch per iteration.  A batched problem only gets here when its cooperative launch was
  // refused; its registrations then run one CTA each (the partial buffer is sized for one grid row).
  const int gx_loop = pr->batched ? 1 : grid_x;
  for (int it = 0; it < opt.max_iterations; ++it) {
    if (comm == kCommNccl) {
      p.mode = kModeAssemble;
      NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, gx_loop, num_problems, 1, ctx->stream));
      const int rc = NcclAllReduceSums(ctx, pr->d_sums);
      if (rc != NLO_OK) return rc;
      p.mode = kModeStepOnly;
      NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, 1, num_problems, 1, ctx->stream));
    } else {
      p.mode = kModeSolve;
      NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, gx_loop, num_problems, 1, ctx->stream));
    }
  }
  return NLO_OK;
}

int RunLoop(nlo_context* ctx, nlo_problem* pr, int kind, const nlo_solve_options& opt,
            bool with_trace, int num_problems, int64_t begin_abs, int64_t end_abs) {
  // NCCL calls are left out of graph capture (their capture support depends on the library
  // build); the other paths run as one CUDA graph so the loop needs a single host call.
  const int comm = CommFor(ctx, pr);
  const int64_t tiles_all = (end_abs + kTile - 1) / kTile - begin_abs / kTile;
  const bool single_launch = UsePersistent(ctx, pr) ||
                             ((comm == kCommNone) && (pr->batched || tiles_all <= kInCtaTiles));
  const bool graph = ctx->use_graph && comm != kCommNccl && !single_launch;
  if (!graph) return EnqueueLoop(ctx, pr, kind, opt, with_trace, num_problems, begin_abs, end_abs);
  int64_t ptol_bits, gtol_bits;
  memcpy(&ptol_bits, &opt.parameter_tolerance, 8);
  memcpy(&gtol_bits, &opt.gradient_tolerance, 8);
  const std::array<int64_t, 10> key = {kind, ctx->loss_kind, comm, opt.max_iterations,
                                       with_trace ? 1 : 0, ctx->generation, ptol_bits, gtol_bits,
                                       begin_abs * 4 + num_problems, end_abs};
  auto it = pr->graphs.find(key);
  if (it == pr->graphs.end()) {
    if (pr->graphs.size() > 16) DropGraphs(pr);
    cudaGraph_t graph_obj = nullptr;
    NLO_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = EnqueueLoop(ctx, pr, kind, opt, with_trace, num_problems, begin_abs, end_abs);
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph_obj);
    if (rc != NLO_OK) {
      if (graph_obj) cudaGraphDestroy(graph_obj);
      return rc;
    }
    NLO_CUDA(ctx, e);
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph_obj, 0);
    cudaGraphDestroy(graph_obj);
    NLO_CUDA(ctx, e);
    it = pr->graphs.emplace(key, exec).first;
  }
  NLO_CUDA(ctx, cudaGraphLaunch(it->second, ctx->stream));
  return NLO_OK;
}

}  // namespace

namespace nlo {

int Fail(nlo_context* ctx, int code, const std::string& msg) {
  if (ctx != nullptr) ctx->error = msg;
  return code;
}

// ---- device allocations, optionally guarded (NLO_GUARD=1; see nlo_host.h) ----
namespace {
constexpr size_t kGuardBytes = 64 * 1024;  // more than two tiles of any kind (24 KB), a whole LL partial row
struct GuardedBlock {
  size_t bytes;  // payload
  int device;
};
struct GuardState {
  std::mutex mu;
  std::map<void*, GuardedBlock> live;  // payload pointer -> block
  int64_t checked = 0;                 // allocations whose bands were compared when they were freed
  int64_t corrupted_bytes = 0;         // guard bytes found != 0xFF, freed allocations
  int64_t peak_live = 0;
  bool selftest_pending = false;       // NLO_GUARD=selftest: one byte of the first tail band is overwritten on purpose
};
GuardState& Guards() {
  static GuardState g;
  return g;
}
int GuardMode() {  // 0 off, 1 on, 2 on + the deliberate overrun that proves the check sees one
  static const int mode = [] {
    const char* v = getenv("NLO_GUARD");
    if (v == nullptr || v[0] == '\0' || v[0] == '0') return 0;
    if (strcmp(v, "selftest") == 0) {
      Guards().selftest_pending = true;
      return 2;
    }
    return 1;
  }();
  return mode;
}
// bytes != 0xFF in the two bands around `payload` (on the device that owns it)
int64_t CountGuardDamage(void* payload, const GuardedBlock& b) {
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(b.device);
  cudaDeviceSynchronize();  // the library's streams are non-blocking: a plain cudaMemcpy does not wait for them
  std::vector<unsigned char> host(2 * kGuardBytes);
  unsigned char* base = static_cast<unsigned char*>(payload) - kGuardBytes;
  const size_t tail_offset = kGuardBytes + ((b.bytes + 255) / 256) * 256;
  int64_t bad = 0;
  if (cudaMemcpy(host.data(), base, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(host.data() + kGuardBytes, base + tail_offset, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    bad = -1;
  } else {
    for (unsigned char c : host) bad += (c != 0xFF) ? 1 : 0;
  }
  cudaSetDevice(prev);
  return bad;
}
}  // namespace

cudaError_t DevMallocBytes(void** p, size_t bytes) {
  if (GuardMode() == 0) return cudaMalloc(p, bytes);
  *p = nullptr;
  // [front band][payload, rounded up to 256 B so that the tail band (and TMA sources) stay aligned][tail band]
  const size_t padded = ((bytes + 255) / 256) * 256;
  unsigned char* base = nullptr;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&base), padded + 2 * kGuardBytes);
  if (e != cudaSuccess) return e;
  e = cudaMemset(base, 0xFF, padded + 2 * kGuardBytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(base);
    return e;
  }
  GuardedBlock b{bytes, 0};
  cudaGetDevice(&b.device);
  GuardState& g = Guards();
  {
    std::lock_guard<std::mutex> lock(g.mu);
    if (g.selftest_pending) {
      g.selftest_pending = false;
      cudaMemset(base + kGuardBytes + padded + 5, 0x00, 1);  // the overrun the self-test expects to be reported
      cudaDeviceSynchronize();
    }
    g.live[base + kGuardBytes] = b;
    g.peak_live = std::max<int64_t>(g.peak_live, static_cast<int64_t>(g.live.size()));
  }
  *p = base + kGuardBytes;
  return cudaSuccess;
}

cudaError_t DevFree(void* p) {
  if (p == nullptr || GuardMode() == 0) return cudaFree(p);
  GuardState& g = Guards();
  GuardedBlock b{0, 0};
  {
    std::lock_guard<std::mutex> lock(g.mu);
    auto it = g.live.find(p);
    if (it == g.live.end()) return cudaFree(p);
    b = it->second;
    g.live.erase(it);
  }
  const int64_t bad = CountGuardDamage(p, b);
  {
    std::lock_guard<std::mutex> lock(g.mu);
    g.checked += 1;
    g.corrupted_bytes += (bad < 0) ? 1 : bad;
  }
  if (bad != 0)
    fprintf(stderr, "[nlo guard] %lld damaged guard byte(s) around a %zu-byte device allocation\n",
            static_cast<long long>(bad), b.bytes);
  return cudaFree(static_cast<unsigned char*>(p) - kGuardBytes);
}

int EnsureStaging(nlo_context* ctx, size_t bytes) {
  if (bytes <= ctx->staging_bytes) return NLO_OK;
  if (ctx->staging != nullptr) {
    NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NLO_CUDA(ctx, DevFree(ctx->staging));
    ctx->staging = nullptr;
    ctx->staging_bytes = 0;
  }
  NLO_CUDA(ctx, DevMalloc(&ctx->staging, bytes));
  ctx->staging_bytes = bytes;
  return NLO_OK;
}

void DropGraphs(nlo_problem* pr) {
  for (auto& kv : pr->graphs) cudaGraphExecDestroy(kv.second);
  pr->graphs.clear();
}

void PoseToRt(const double pose[16], double R[9], double t[3]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = pose[4 * c + r];
  t[0] = pose[12];
  t[1] = pose[13];
  t[2] = pose[14];
}

int CreateProblem(nlo_context* ctx, int family, int num_problems, const int64_t* counts,
                  bool batched, nlo_problem** out, bool f32) {
  if (ctx == nullptr || out == nullptr || num_problems < 1) return Fail(ctx, NLO_EINVAL, "bad argument");
  // registrations sit on gridDim.y of the batched launches
  if (num_problems > 65535) return Fail(ctx, NLO_EINVAL, "more than 65535 registrations in one batched problem");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  nlo_problem* pr = new nlo_problem();
  pr->family = family;
  pr->num_planes = (family == 0) ? kNdtPlanes : kReprojPlanes;
  pr->num_problems = num_problems;
  pr->batched = batched;
  pr->f32 = f32;
  int64_t cursor = 0;
  for (int k = 0; k < num_problems; ++k) {
    if (counts[k] < 0) {
      delete pr;
      return Fail(ctx, NLO_EINVAL, "negative count");
    }
    pr->counts.push_back(counts[k]);
    pr->h_ranges.push_back(Range{cursor, cursor + (batched ? counts[k] : 0)});
    cursor += std::max<int64_t>(PadToTile(counts[k]), kTile);
  }
  pr->capacity = cursor;
  const size_t elem = f32 ? sizeof(float) : sizeof(double);
  const size_t plane_bytes = static_cast<size_t>(pr->capacity) * elem;
  auto fail_free = [&](int code, const std::string& msg) {
    nlo_problem_destroy(ctx, pr);
    return Fail(ctx, code, msg);
  };
#define NLO_CUDA_P(expr)                                                               \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      return fail_free((_e == cudaErrorMemoryAllocation) ? NLO_ENOMEM : NLO_ECUDA,     \
                       std::string(#expr) + ": " + cudaGetErrorString(_e));            \
  } while (0)
  NLO_CUDA_P(DevMalloc(&pr->plane_block, plane_bytes * pr->num_planes));
  NLO_CUDA_P(cudaMemsetAsync(pr->plane_block, 0, plane_bytes * pr->num_planes, ctx->stream));
  // tile-interleaved layout: plane k of tile 0 starts at k * 256 (nlo_internal.h TiledOffset)
  for (int k = 0; k < pr->num_planes; ++k)
    pr->planes[k] = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(pr->plane_block) +
                                              static_cast<size_t>(k) * kTile * elem);
  const int slots = num_problems + 1;
  NLO_CUDA_P(DevMalloc(&pr->d_ranges, slots * sizeof(Range)));
  NLO_CUDA_P(DevMalloc(&pr->d_states, slots * sizeof(State)));
  // per-CTA partials [2 parities][grid rows x CTAs per row][28]: a single problem uses up to
  // grid_single CTAs, a small batch num_problems x (grid_single / num_problems) <= grid_single
  NLO_CUDA_P(DevMalloc(&pr->d_partials,
                        2 * static_cast<size_t>(std::max(ctx->grid_single, 1)) * kAcc6 * sizeof(double)));
  // cluster partials of the persistent path as LL words: [2 parities][<= grid_single clusters][28][2]
  pr->ll_partials_bytes = 2 * static_cast<size_t>(std::max(ctx->grid_single, 1) + kMaxCluster) * kAcc6 * 2 * sizeof(unsigned long long);
  NLO_CUDA_P(DevMalloc(&pr->d_ll_partials, pr->ll_partials_bytes));
  NLO_CUDA_P(cudaMemsetAsync(pr->d_ll_partials, 0, pr->ll_partials_bytes, ctx->stream));
  NLO_CUDA_P(DevMalloc(&pr->d_sync, static_cast<size_t>(slots) * kSyncStride * sizeof(unsigned long long)));
  NLO_CUDA_P(cudaMemsetAsync(pr->d_sync, 0, static_cast<size_t>(slots) * kSyncStride * sizeof(unsigned long long), ctx->stream));
  NLO_CUDA_P(DevMalloc(&pr->d_tickets, slots * sizeof(unsigned int)));
  NLO_CUDA_P(cudaMemsetAsync(pr->d_tickets, 0, slots * sizeof(unsigned int), ctx->stream));
  NLO_CUDA_P(DevMalloc(&pr->d_sums, slots * 32 * sizeof(double)));
  NLO_CUDA_P(cudaMemsetAsync(pr->d_sums, 0, slots * 32 * sizeof(double), ctx->stream));
  NLO_CUDA_P(DevMalloc(&pr->d_poses, slots * 16 * sizeof(double)));
  NLO_CUDA_P(DevMalloc(&pr->d_results, slots * 4 * sizeof(double)));
  NLO_CUDA_P(cudaMemcpyAsync(pr->d_ranges, pr->h_ranges.data(), num_problems * sizeof(Range),
                             cudaMemcpyHostToDevice, ctx->stream));
  NLO_CUDA_P(cudaStreamSynchronize(ctx->stream));
#undef NLO_CUDA_P
  *out = pr;
  return NLO_OK;
}

// Grows the device trace buffer of a problem.  cudaMalloc / cudaFree may wait for the device to go
// idle, so a multi-device context calls this for every shard BEFORE it starts the shards' solves:
// on a device shared by several shards the kernels already running would spin on the one whose
// host thread sits in cudaMalloc.
int EnsureTrace(nlo_context* ctx, nlo_problem* pr, size_t need_doubles) {
  if (need_doubles <= pr->trace_doubles) return NLO_OK;
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  if (pr->d_trace) NLO_CUDA(ctx, DevFree(pr->d_trace));
  pr->d_trace = nullptr;
  pr->trace_doubles = 0;
  NLO_CUDA(ctx, DevMalloc(&pr->d_trace, need_doubles * sizeof(double)));
  pr->trace_doubles = need_doubles;
  DropGraphs(pr);
  return NLO_OK;
}

int AssembleImpl(nlo_context* ctx, nlo_problem* pr, int kind, int problem_index, const double pose[16],
                 int64_t begin, int64_t end, double* H, int nh, double* g, int ng, double* cost) {
  // (constructed first: whatever way this call returns, the other shards' threads are not left waiting)
  RendezvousGuard rendezvous(
      (ctx != nullptr && pr != nullptr && CommFor(ctx, pr) == kCommPeer) ? ctx->launch_barrier : nullptr, 2);
  if (ctx == nullptr || pr == nullptr || pose == nullptr || H == nullptr || g == nullptr || cost == nullptr)
    return Fail(ctx, NLO_EINVAL, "null argument");
  if ((kind == kReproj) != (pr->family == 1)) return Fail(ctx, NLO_EINVAL, "problem family mismatch");
  if (problem_index < 0 || problem_index >= pr->num_problems) return Fail(ctx, NLO_EINVAL, "bad problem_index");
  const int64_t count = pr->batched ? pr->counts[problem_index] : pr->n;
  if (begin < 0 || end < begin || end > count) return Fail(ctx, NLO_EINVAL, "bad [begin, end)");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int comm = CommFor(ctx, pr);
  const int slot = pr->num_problems;  // scratch slot
  const int64_t base = pr->h_ranges[problem_index].begin;
  double* hs = ctx->host_small;
  memcpy(hs, pose, 16 * sizeof(double));
  Range* hr = reinterpret_cast<Range*>(hs + 16);
  hr->begin = base + begin;
  hr->end = base + end;
  NLO_CUDA(ctx, cudaMemcpyAsync(pr->d_poses + 16 * slot, hs, 16 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NLO_CUDA(ctx, cudaMemcpyAsync(pr->d_ranges + slot, hr, sizeof(Range), cudaMemcpyHostToDevice, ctx->stream));
  NLO_CUDA(ctx, LaunchInitStates(pr->d_states + slot, pr->d_poses + 16 * slot, 1, kind, ctx->stream));
  IterParams p = BaseParams(ctx, pr);
  p.ranges = pr->d_ranges + slot;
  p.states = pr->d_states + slot;
  p.tickets = pr->d_tickets + slot;
  p.sums = pr->d_sums + 32 * slot;
  p.mode = kModeAssemble;
  const int grid_x = GridFor(ctx, hr->begin, hr->end);
  if (rendezvous.active()) NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  rendezvous.Sync();  // every shard is ready to launch
  NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, grid_x, 1, 1, ctx->stream));
  rendezvous.Sync();  // every shard's kernel is queued: only now may anything wait behind it
  if (comm == kCommNccl) {
    const int rc = NcclAllReduceSums(ctx, pr->d_sums + 32 * slot);
    if (rc != NLO_OK) return rc;
  }
  double* out = hs + 64;
  NLO_CUDA(ctx, cudaMemcpyAsync(out, pr->d_sums + 32 * slot, 32 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(H, out, nh * sizeof(double));
  memcpy(g, out + nh, ng * sizeof(double));
  *cost = out[nh + ng];
  return CheckPeerError(ctx, pr);
}

int SolveImpl(nlo_context* ctx, nlo_problem* pr, int kind, const nlo_solve_options* options,
              double* poses, nlo_solve_result* results, double* trace, bool batched_call) {
  RendezvousGuard rendezvous(
      (ctx != nullptr && pr != nullptr && CommFor(ctx, pr) == kCommPeer) ? ctx->launch_barrier : nullptr, 2);
  if (ctx == nullptr || pr == nullptr || options == nullptr || poses == nullptr || results == nullptr)
    return Fail(ctx, NLO_EINVAL, "null argument");
  if ((kind == kReproj) != (pr->family == 1)) return Fail(ctx, NLO_EINVAL, "problem family mismatch");
  if (options->max_iterations < 0) return Fail(ctx, NLO_EINVAL, "max_iterations < 0");
  if (batched_call != pr->batched) return Fail(ctx, NLO_EINVAL, "batched / single problem mismatch");
  const int B = pr->num_problems;
  for (int k = 0; k < B; ++k) results[k] = nlo_solve_result{0, 0, 0.0, 0.0};
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int trace_width = (kind == kNdt3) ? NLO_TRACE3 : NLO_TRACE6;
  const bool with_trace = (trace != nullptr) && !pr->batched && options->max_iterations > 0;
  if (with_trace) {
    const size_t need = static_cast<size_t>(options->max_iterations) * trace_width;
    const int rc = EnsureTrace(ctx, pr, need);
    if (rc != NLO_OK) return rc;
    NLO_CUDA(ctx, cudaMemsetAsync(pr->d_trace, 0, need * sizeof(double), ctx->stream));
  }
  // ranges of the registrations (single problem: [0, n) resp. floor(n/4)*4 for the 3-DoF path,
  // ..._analytic_3dof.cc:33-36; a shard of a larger scan gets its end from the owner of the shards)
  int64_t begin_abs = 0, end_abs = 0;
  if (!pr->batched) {
    Range r{0, pr->n};
    if (kind == kNdt3) r.end = (pr->ndt3_end_override >= 0) ? std::min(pr->ndt3_end_override, pr->n) : (pr->n / 4) * 4;
    Range* hr = reinterpret_cast<Range*>(ctx->host_small + 32);
    *hr = r;
    NLO_CUDA(ctx, cudaMemcpyAsync(pr->d_ranges, hr, sizeof(Range), cudaMemcpyHostToDevice, ctx->stream));
    begin_abs = r.begin;
    end_abs = r.end;
  } else {
    int64_t max_count = 0;
    for (int64_t c : pr->counts) max_count = std::max(max_count, c);
    end_abs = max_count;
    // per-registration ranges for this kind (3-DoF: floor(n/4)*4 of each registration)
    std::vector<Range> ranges(pr->h_ranges);
    if (kind == kNdt3)
      for (int k = 0; k < B; ++k) ranges[k].end = ranges[k].begin + (pr->counts[k] / 4) * 4;
    NLO_CUDA(ctx, cudaMemcpyAsync(pr->d_ranges, ranges.data(), static_cast<size_t>(B) * sizeof(Range),
                                  cudaMemcpyHostToDevice, ctx->stream));
    NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `ranges` is a local vector
  }
  NLO_CUDA(ctx, cudaMemcpyAsync(pr->d_poses, poses, static_cast<size_t>(B) * 16 * sizeof(double),
                                cudaMemcpyHostToDevice, ctx->stream));
  NLO_CUDA(ctx, LaunchInitStates(pr->d_states, pr->d_poses, B, kind, ctx->stream));
  if (options->max_iterations > 0) {
    if (rendezvous.active()) NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rendezvous.Sync();  // every shard is ready to launch
    // shards on several GPUs: a device-side barrier, so that the loops and the events around them
    // start together (the first exchange would otherwise absorb the hosts' launch skew)
    if (CommFor(ctx, pr) == kCommPeer) NLO_CUDA(ctx, LaunchPeerRendezvous(ctx->peer, ctx->stream));
  }
  NLO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  if (options->max_iterations > 0) {
    const int rc = RunLoop(ctx, pr, kind, *options, with_trace, B, begin_abs, end_abs);
    if (rc != NLO_OK) return rc;
    rendezvous.Sync();  // every shard's loop is queued: only now may anything wait behind it
  }
  NLO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  NLO_CUDA(ctx, LaunchFinishStates(pr->d_states, pr->d_poses, pr->d_results, B, kind, ctx->stream));
  // a one-registration solve reads its results back through the pinned scratch; a batch too large
  // for it through a pageable vector
  const bool small = static_cast<size_t>(B) * 4 + 128 <= static_cast<size_t>(kSmallDoubles);
  std::vector<double> res_big(small ? 0 : static_cast<size_t>(B) * 4);
  double* res = small ? ctx->host_small + 128 : res_big.data();
  NLO_CUDA(ctx, cudaMemcpyAsync(poses, pr->d_poses, static_cast<size_t>(B) * 16 * sizeof(double),
                                cudaMemcpyDeviceToHost, ctx->stream));
  NLO_CUDA(ctx, cudaMemcpyAsync(res, pr->d_results, static_cast<size_t>(B) * 4 * sizeof(double),
                                cudaMemcpyDeviceToHost, ctx->stream));
  if (with_trace)
    NLO_CUDA(ctx, cudaMemcpyAsync(trace, pr->d_trace,
                                  static_cast<size_t>(options->max_iterations) * trace_width * sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->stream));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  NLO_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  if (ctx->d_debug_times != nullptr && options->max_iterations >= 8 && options->max_iterations <= kDebugIterations) {
    std::vector<unsigned long long> ts(kDebugIterations * kDebugSlots);
    cudaMemcpy(ts.data(), ctx->d_debug_times, ts.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    // Phase stamps of CTA 0, averaged over iterations 2..7 (ns).  Stamps 0-6: iteration start, tiles done,
    // CTA sync, [resident: cluster pre-reduction + LL store | streaming: all CTAs arrived], sums gathered,
    // stepped, CTA sync; 8-14 (resident kernel only): sub-phases.  A stamp a launch shape does not take is 0.
    auto span = [&](int from, int to, double* out) {
      double total = 0.0;
      for (int it = 2; it < 8; ++it) {
        const unsigned long long a0 = ts[it * kDebugSlots + from], a1 = ts[it * kDebugSlots + to];
        if (a0 == 0ULL || a1 == 0ULL) return false;
        total += static_cast<double>(a1) - static_cast<double>(a0);
      }
      *out = total / 6.0;
      return true;
    };
    if (ctx->debug_all_ctas && getenv("NLO_DEBUG_FILE") != nullptr) {
      std::vector<unsigned long long> all(static_cast<size_t>(kDebugCtas) * kDebugIterations * kDebugSlots);
      cudaMemcpy(all.data(), ctx->d_debug_times, all.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      if (FILE* f = fopen(getenv("NLO_DEBUG_FILE"), "wb")) {
        fwrite(all.data(), sizeof(unsigned long long), all.size(), f);
        fclose(f);
      }
    }
    std::string line = "[nlo debug] ns/iter:";
    const char* names[6] = {"tiles", "cta-sync", "pre-reduce|arrivals", "gather", "canon+xchg+step", "sync"};
    char buf[96];
    for (int k = 0; k < 6; ++k) {
      double v = 0.0;
      // a missing stamp 3 (single-CTA loop) folds phases 2 and 3 into one
      const bool ok = span(k, k + 1, &v);
      if (ok) snprintf(buf, sizeof(buf), " %s %.0f |", names[k], v);
      else snprintf(buf, sizeof(buf), " %s - |", names[k]);
      line += buf;
    }
    double period = 0.0;
    {
      double total = 0.0;
      for (int it = 2; it < 8; ++it)
        total += static_cast<double>(ts[(it + 1) * kDebugSlots]) - static_cast<double>(ts[it * kDebugSlots]);
      period = total / 6.0;
    }
    snprintf(buf, sizeof(buf), " period %.0f", period);
    line += buf;
    fprintf(stderr, "%s\n", line.c_str());
    double sub = 0.0;
    if (span(2, 8, &sub)) {  // resident kernel: sub-phases after the CTA sync
      const char* sub_names[7] = {"reduce-entry", "totals-in-cluster", "canonical", "reduce-returned", "step-entry",
                                  "stepped", "state-out"};
      std::string sline = "[nlo debug] ns after cta-sync:";
      for (int k = 0; k < 7; ++k) {
        double v = 0.0;
        if (span(2, 8 + k, &v)) {
          snprintf(buf, sizeof(buf), " %s %.0f |", sub_names[k], v);
          sline += buf;
        }
      }
      fprintf(stderr, "%s\n", sline.c_str());
    }
    cudaMemset(ctx->d_debug_times, 0, static_cast<size_t>(kDebugCtas) * kDebugIterations * kDebugSlots * sizeof(unsigned long long));
  }
  int rc_all = NLO_OK;
  for (int k = 0; k < B; ++k) {
    const int device_status = static_cast<int>(res[4 * k + 1]);  // State::status: 1 non-finite, 2 grid wait expired
    results[k].iterations = static_cast<int32_t>(res[4 * k]);
    results[k].status = device_status == 0 ? NLO_OK : (device_status == 2 ? NLO_ETIMEOUT : NLO_ENUMERIC);
    results[k].final_cost = res[4 * k + 2];
    results[k].device_ms = ms;
    if (results[k].status != NLO_OK && rc_all != NLO_ETIMEOUT) rc_all = results[k].status;
  }
  const int pe = CheckPeerError(ctx, pr);
  if (pe != NLO_OK) return pe;
  if (rc_all == NLO_ETIMEOUT)
    return Fail(ctx, rc_all, "a grid-wide wait of the persistent iteration kernel expired (GPU shared or preempted?)");
  if (rc_all != NLO_OK) return Fail(ctx, rc_all, "non-finite value met during the solve");
  return NLO_OK;
}

}  // namespace nlo

extern "C" {

int nlo_abi_version(void) { return NLO_ABI_VERSION; }

int nlo_visible_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return count;
}

int nlo_context_create(int device, nlo_context** out) {
  if (out == nullptr) return NLO_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return NLO_ECUDA;
  if (cudaSetDevice(device) != cudaSuccess) return NLO_ECUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NLO_ECUDA;
  if (prop.major < 10) return NLO_ECUDA;  // sm_100a binary only; there is no fallback path
  nlo_context* ctx = new nlo_context();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->grid_single = 2 * prop.multiProcessorCount;
  const char* env = getenv("NLO_NO_GRAPH");
  ctx->use_graph = !(env != nullptr && env[0] == '1');
  const char* penv = getenv("NLO_NO_PERSISTENT");
  ctx->use_persistent = !(penv != nullptr && penv[0] == '1');
  ctx->grid_small = prop.multiProcessorCount;
  const char* sgenv = getenv("NLO_GRID_SMALL");
  if (sgenv != nullptr && atoi(sgenv) > 0) ctx->grid_small = atoi(sgenv);
  // A scan that is re-read every iteration and does not fit the L2 by itself gets its first 96 MB
  // pinned (evict_last) and the rest streamed (evict_first).  Measured on B200 (scripts/l2_sweep.sh,
  // 3-DoF): 1 M points 23.6 -> 17.5 us / iteration, 2 M 41.3 -> 35.9, 4 M 75.4 -> 70.2; 110 MB thrashes.
  ctx->l2_keep_mb = 96.0;
  ctx->l2_policy_min_mb = 56.0;
  const char* kenv = getenv("NLO_L2_KEEP_MB");
  if (kenv != nullptr) ctx->l2_keep_mb = atof(kenv);
  const char* menv = getenv("NLO_L2_MIN_MB");
  if (menv != nullptr) ctx->l2_policy_min_mb = atof(menv);
  const char* senv = getenv("NLO_STAGE_DEPTH");
  if (senv != nullptr) ctx->stage_depth = atoi(senv);
  // Thread-block clusters of the persistent path (DSMEM pre-reduction of the per-CTA sums): CTAs per
  // cluster for problems that live in shared memory / L2 and for streamed scans, and the number of
  // cluster partials up to which every CTA gathers them itself.
  auto env_int = [](const char* name, int fallback, int lo, int hi) {
    const char* v = getenv(name);
    if (v == nullptr) return fallback;
    return std::max(lo, std::min(hi, atoi(v)));
  };
  auto pow2_floor = [](int v) { int p2 = 1; while (2 * p2 <= v) p2 *= 2; return p2; };
  ctx->cluster_small = pow2_floor(env_int("NLO_CLUSTER", 8, 1, kMaxCluster));
  ctx->direct_max_clusters = env_int("NLO_DIRECT_MAX", 48, 0, 1 << 20);
  ctx->use_resident = env_int("NLO_NO_RESIDENT", 0, 0, 1) == 0;
  ctx->warp_groups = env_int("NLO_WARP_GROUPS", 2, 1, 2);
  const char* tenv = getenv("NLO_INGEST_THREADS");
  if (tenv != nullptr) ctx->ingest_threads = atoi(tenv);
  const char* denv = getenv("NLO_DEBUG_TIMES");
  if (denv != nullptr && (denv[0] == '1' || denv[0] == '2')) {
    // 1: phases of CTA 0 printed to stderr after every solve; 2: stamps of every CTA also written to
    // $NLO_DEBUG_FILE (raw u64 [kDebugCtas][kDebugIterations][8], slot 7 of iteration 0 = SM id)
    ctx->debug_all_ctas = denv[0] == '2';
    const size_t bytes = static_cast<size_t>(kDebugCtas) * kDebugIterations * kDebugSlots * sizeof(unsigned long long);
    DevMalloc(reinterpret_cast<void**>(&ctx->d_debug_times), bytes);
    cudaMemset(ctx->d_debug_times, 0, bytes);
  }
  const char* genv = getenv("NLO_GRID");
  if (genv != nullptr && atoi(genv) > 0) ctx->grid_single = atoi(genv);
  bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess &&
            cudaMallocHost(reinterpret_cast<void**>(&ctx->host_small), kSmallDoubles * sizeof(double)) == cudaSuccess &&
            ConfigureKernels() == cudaSuccess;
  if (!ok) {
    nlo_context_destroy(ctx);
    return NLO_ECUDA;
  }
  *out = ctx;
  return NLO_OK;
}

int nlo_context_create_multi(const int32_t* devices, int32_t num_devices, nlo_context** out) {
  if (out == nullptr) return NLO_EINVAL;
  *out = nullptr;
  if (devices == nullptr || num_devices < 1 || num_devices > kMaxRanks) return NLO_EINVAL;
  std::vector<int> dev(devices, devices + num_devices);
  return multi::CreateContext(dev.data(), num_devices, out);
}

int nlo_context_device_count(const nlo_context* ctx) {
  if (ctx == nullptr) return 0;
  return ctx->IsMulti() ? static_cast<int>(ctx->subs.size()) : 1;
}

int nlo_context_destroy(nlo_context* ctx) {
  if (ctx == nullptr) return NLO_OK;
  if (ctx->IsMulti()) {
    multi::DestroyContext(ctx);
    return NLO_OK;
  }
  cudaSetDevice(ctx->device);
  nlo_comm_destroy(ctx);
  if (ctx->reg_workspace) nlo_problem_destroy(ctx, ctx->reg_workspace);
  ctx->reg_workspace = nullptr;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  FreeIngestRing(ctx);
  FreeIngestPool(ctx);
  if (ctx->staging) DevFree(ctx->staging);
  if (ctx->d_debug_times) DevFree(ctx->d_debug_times);
  if (ctx->host_small) cudaFreeHost(ctx->host_small);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return NLO_OK;
}

int nlo_debug_guard_report(int32_t* enabled, int64_t* allocations_checked, int64_t* allocations_live,
                           int64_t* corrupted_bytes) {
  const int mode = GuardMode();
  if (enabled) *enabled = mode != 0 ? 1 : 0;
  int64_t checked = 0, live = 0, bad = 0;
  if (mode != 0) {
    GuardState& g = Guards();
    std::map<void*, GuardedBlock> snapshot;
    {
      std::lock_guard<std::mutex> lock(g.mu);
      snapshot = g.live;
      checked = g.checked;
      bad = g.corrupted_bytes;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) == cudaSuccess) {
      int prev = 0;
      cudaGetDevice(&prev);
      for (int d = 0; d < count; ++d) {
        cudaSetDevice(d);
        cudaDeviceSynchronize();
      }
      cudaSetDevice(prev);
    }
    for (const auto& kv : snapshot) {
      const int64_t b = CountGuardDamage(kv.first, kv.second);
      bad += (b < 0) ? 1 : b;
      checked += 1;
    }
    live = static_cast<int64_t>(snapshot.size());
  }
  if (allocations_checked) *allocations_checked = checked;
  if (allocations_live) *allocations_live = live;
  if (corrupted_bytes) *corrupted_bytes = bad;
  return NLO_OK;
}

const char* nlo_last_error(const nlo_context* ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int nlo_context_info(const nlo_context* ctx, int* sm_count, int* assemble_grid) {
  if (ctx == nullptr) return NLO_EINVAL;
  const nlo_context* c = ctx->IsMulti() ? ctx->subs[0] : ctx;
  if (sm_count) *sm_count = c->sm_count;
  if (assemble_grid) *assemble_grid = c->grid_single;
  return NLO_OK;
}

int nlo_synchronize(nlo_context* ctx) {
  if (ctx == nullptr) return NLO_EINVAL;
  if (ctx->IsMulti()) return multi::Synchronize(ctx);
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NLO_OK;
}

int nlo_set_loss(nlo_context* ctx, int kind, const double params[2]) {
  if (ctx == nullptr) return NLO_EINVAL;
  double p0 = params ? params[0] : 0.0, p1 = params ? params[1] : 0.0;
  switch (kind) {
    case NLO_LOSS_NONE: break;
    case NLO_LOSS_EXPONENTIAL:  // loss_function.h:24-25: negative parameters are rejected
      if (params == nullptr || p0 < 0.0 || p1 < 0.0) return Fail(ctx, NLO_EINVAL, "c1, c2 should be positive numbers");
      break;
    case NLO_LOSS_HUBER:  // loss_function.h:53-54
      if (params == nullptr || !(p0 > 0.0)) return Fail(ctx, NLO_EINVAL, "threshold value should be larger than zero");
      break;
    case NLO_LOSS_CAUCHY:
      if (params == nullptr || !(p0 > 0.0)) return Fail(ctx, NLO_EINVAL, "c should be larger than zero");
      break;
    default: return Fail(ctx, NLO_EINVAL, "unknown loss kind");
  }
  if (ctx->IsMulti()) return multi::SetLoss(ctx, kind, params);
  if (kind == NLO_LOSS_CAUCHY) p1 = 1.0 / (p0 * p0);  // the kernels multiply by 1 / c^2 instead of dividing per correspondence
  ctx->loss_kind = kind;
  ctx->loss_params[0] = p0;
  ctx->loss_params[1] = p1;
  ctx->generation++;
  return NLO_OK;
}

// Pinned host memory ON THE NUMA NODE OF THE CURRENT DEVICE (best effort): anonymous pages bound to the
// node of the GPU's PCIe root, touched, then registered with CUDA.  With several GPUs per box the
// uploads of the GPUs on the second socket otherwise all cross the inter-socket link (a process
// usually starts on the first socket and cudaMallocHost places the pages where the caller runs).
namespace {
std::mutex g_host_mu;
std::map<void*, size_t> g_host_blocks;  // mmap'ed + registered blocks handed out by nlo_host_alloc
}  // namespace

int nlo_host_alloc(void** ptr, size_t bytes) {
  if (ptr == nullptr) return NLO_EINVAL;
  *ptr = nullptr;
  if (bytes == 0) return NLO_EINVAL;
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) {
    cudaGetLastError();
    return NLO_ECUDA;
  }
  const size_t page = 1u << 21;
  const size_t len = ((bytes + page - 1) / page) * page;
  void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) return NLO_ENOMEM;
#ifdef SYS_mbind
  const int node = DeviceNumaNode(device);
  if (node >= 0 && node < 64) {
    unsigned long mask = 1ul << node;
    syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, &mask, 65ul, 0u);  // refused (cpuset) = default placement
  }
#endif
  madvise(p, len, MADV_HUGEPAGE);
  memset(p, 0, len);  // first touch under the policy
  if (cudaHostRegister(p, len, cudaHostRegisterDefault) != cudaSuccess) {
    cudaGetLastError();
    munmap(p, len);
    return NLO_ENOMEM;
  }
  {
    std::lock_guard<std::mutex> lock(g_host_mu);
    g_host_blocks[p] = len;
  }
  *ptr = p;
  return NLO_OK;
}
int nlo_host_free(void* ptr) {
  if (ptr == nullptr) return NLO_OK;
  size_t len = 0;
  {
    std::lock_guard<std::mutex> lock(g_host_mu);
    auto it = g_host_blocks.find(ptr);
    if (it == g_host_blocks.end()) return NLO_EINVAL;
    len = it->second;
    g_host_blocks.erase(it);
  }
  const bool ok = cudaHostUnregister(ptr) == cudaSuccess;
  munmap(ptr, len);
  return ok ? NLO_OK : NLO_ECUDA;
}

// ---- problems ----
static int CreateAny(nlo_context* ctx, int family, int num_problems, const int64_t* counts, bool batched, bool f32,
                     nlo_problem** problem) {
  if (ctx == nullptr || problem == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  if (ctx->IsMulti()) return multi::Create(ctx, family, num_problems, counts, batched, f32, problem);
  return CreateProblem(ctx, family, num_problems, counts, batched, problem, f32);
}

int nlo_ndt_create(nlo_context* ctx, int64_t capacity, nlo_problem** problem) {
  if (capacity < 0) return Fail(ctx, NLO_EINVAL, "negative capacity");
  return CreateAny(ctx, 0, 1, &capacity, false, false, problem);
}

int nlo_ndt_create_f32(nlo_context* ctx, int64_t capacity, nlo_problem** problem) {
  if (capacity < 0) return Fail(ctx, NLO_EINVAL, "negative capacity");
  return CreateAny(ctx, 0, 1, &capacity, false, true, problem);
}

int nlo_ndt_create_batched(nlo_context* ctx, int32_t num_problems, const int64_t* counts, nlo_problem** problem) {
  if (counts == nullptr) return Fail(ctx, NLO_EINVAL, "null counts");
  return CreateAny(ctx, 0, num_problems, counts, true, false, problem);
}

int nlo_reproj_create(nlo_context* ctx, int64_t capacity, nlo_problem** problem) {
  if (capacity < 0) return Fail(ctx, NLO_EINVAL, "negative capacity");
  return CreateAny(ctx, 1, 1, &capacity, false, false, problem);
}

int nlo_reproj_create_batched(nlo_context* ctx, int32_t num_problems, const int64_t* counts, nlo_problem** problem) {
  if (counts == nullptr) return Fail(ctx, NLO_EINVAL, "null counts");
  return CreateAny(ctx, 1, num_problems, counts, true, false, problem);
}

int nlo_problem_destroy(nlo_context* ctx, nlo_problem* pr) {
  if (pr == nullptr) return NLO_OK;
  if (!pr->shards.empty() || (ctx != nullptr && ctx->IsMulti())) return multi::Destroy(ctx, pr);
  if (ctx != nullptr) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  DropGraphs(pr);
  DevFree(pr->plane_block);
  DevFree(pr->d_ranges);
  DevFree(pr->d_states);
  DevFree(pr->d_partials);
  DevFree(pr->d_tickets);
  DevFree(pr->d_sync);
  DevFree(pr->d_ll_partials);
  DevFree(pr->d_sums);
  DevFree(pr->d_poses);
  DevFree(pr->d_results);
  DevFree(pr->d_trace);
  delete pr;
  return NLO_OK;
}

int64_t nlo_problem_size(const nlo_problem* pr) { return pr ? pr->n : -1; }

// ---- ingest (nlo_ingest.cu) ----
int nlo_ndt_upload(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* point, const double* mean,
                   const double* sqrt_info) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0) return Fail(ctx, NLO_EINVAL, "bad problem");
  if (n < 0 || (n > 0 && (point == nullptr || mean == nullptr || sqrt_info == nullptr)))
    return Fail(ctx, NLO_EINVAL, "null array");
  if (ctx->IsMulti()) return multi::UploadNdt(ctx, pr, n, point, mean, sqrt_info);
  return UploadNdt(ctx, pr, n, point, mean, sqrt_info);
}

int nlo_ndt_upload_f32(nlo_context* ctx, nlo_problem* pr, int64_t n, const float* point, const float* mean,
                       const float* sqrt_info) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0 || pr->batched || !pr->f32)
    return Fail(ctx, NLO_EINVAL, "bad problem (needs an fp32-storage NDT problem)");
  if (n < 0 || n > pr->counts[0] || (n > 0 && (point == nullptr || mean == nullptr || sqrt_info == nullptr)))
    return Fail(ctx, NLO_EINVAL, "bad n / null array");
  if (ctx->IsMulti()) return multi::UploadNdtF32(ctx, pr, n, point, mean, sqrt_info);
  return UploadNdtF32(ctx, pr, n, point, mean, sqrt_info);
}

int nlo_ndt_upload_aos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                       size_t offset_point, size_t offset_mean, size_t offset_sqrt_info, int col_major) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0 || pr->batched || pr->f32)
    return Fail(ctx, NLO_EINVAL, "bad problem (AoS ingest needs a single fp64 NDT problem)");
  if (n < 0 || n > pr->counts[0] || (n > 0 && records == nullptr)) return Fail(ctx, NLO_EINVAL, "bad n / records");
  if (stride % 8 != 0 || offset_point % 8 != 0 || offset_mean % 8 != 0 || offset_sqrt_info % 8 != 0)
    return Fail(ctx, NLO_EINVAL, "record layout must be 8-byte aligned");
  if (offset_point + 24 > stride || offset_mean + 24 > stride || offset_sqrt_info + 72 > stride)
    return Fail(ctx, NLO_EINVAL, "field outside the record");
  if (ctx->IsMulti())
    return multi::UploadNdtAos(ctx, pr, n, records, stride, offset_point, offset_mean, offset_sqrt_info, col_major);
  return UploadNdtAos(ctx, pr, n, records, stride, offset_point, offset_mean, offset_sqrt_info, col_major);
}

int nlo_ndt_generate(nlo_context* ctx, nlo_problem* pr, int64_t n, uint64_t seed, int64_t global_index_offset,
                     double noise_sigma, const double true_pose[16], const double init_pose[16],
                     const double grid_origin[3], const int32_t grid_dims[3], double voxel_size,
                     const double* cell_mean, const double* cell_sqrt_info, const uint8_t* cell_valid) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0 || pr->batched) return Fail(ctx, NLO_EINVAL, "bad problem");
  if (n < 0 || n > pr->counts[0]) return Fail(ctx, NLO_EINVAL, "n exceeds capacity");
  auto fn = ctx->IsMulti() ? multi::Generate : GenerateNdt;
  return fn(ctx, pr, seed, global_index_offset, noise_sigma, true_pose, init_pose, grid_origin, grid_dims,
            voxel_size, cell_mean, cell_sqrt_info, cell_valid, n);
}

int nlo_ndt_generate_batched(nlo_context* ctx, nlo_problem* pr, uint64_t seed, double noise_sigma,
                             const double* true_poses, const double init_pose[16], const double grid_origin[3],
                             const int32_t grid_dims[3], double voxel_size, const double* cell_mean,
                             const double* cell_sqrt_info, const uint8_t* cell_valid) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0 || !pr->batched) return Fail(ctx, NLO_EINVAL, "bad problem");
  auto fn = ctx->IsMulti() ? multi::Generate : GenerateNdt;
  return fn(ctx, pr, seed, 0, noise_sigma, true_poses, init_pose, grid_origin, grid_dims, voxel_size, cell_mean,
            cell_sqrt_info, cell_valid, 0);
}

int nlo_ndt_download(nlo_context* ctx, const nlo_problem* pr, int64_t begin, int64_t end, double* point,
                     double* mean, double* information) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0 || pr->batched) return Fail(ctx, NLO_EINVAL, "bad problem");
  auto fn = ctx->IsMulti() ? multi::Download : DownloadNdt;
  return fn(ctx, pr, 0, begin, end, point, mean, information);
}

int nlo_ndt_download_problem(nlo_context* ctx, const nlo_problem* pr, int32_t problem_index, int64_t begin,
                             int64_t end, double* point, double* mean, double* information) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0) return Fail(ctx, NLO_EINVAL, "bad problem");
  auto fn = ctx->IsMulti() ? multi::Download : DownloadNdt;
  return fn(ctx, pr, problem_index, begin, end, point, mean, information);
}

int nlo_reproj_upload(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* local_point, const double* pixel,
                      const double intrinsics[6]) {
  if (ctx == nullptr || pr == nullptr || pr->family != 1) return Fail(ctx, NLO_EINVAL, "bad problem");
  if (n < 0 || intrinsics == nullptr || (n > 0 && (!local_point || !pixel))) return Fail(ctx, NLO_EINVAL, "bad argument");
  if (ctx->IsMulti()) return multi::UploadReproj(ctx, pr, n, local_point, pixel, intrinsics);
  return UploadReproj(ctx, pr, n, local_point, pixel, intrinsics);
}

int nlo_reproj_upload_aos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                          size_t offset_local_point, size_t offset_pixel, const double intrinsics[6]) {
  if (ctx == nullptr || pr == nullptr || pr->family != 1 || pr->batched)
    return Fail(ctx, NLO_EINVAL, "bad problem (AoS ingest needs a single reprojection problem)");
  if (n < 0 || n > pr->counts[0] || intrinsics == nullptr || (n > 0 && records == nullptr))
    return Fail(ctx, NLO_EINVAL, "bad argument");
  if (stride % 8 != 0 || offset_local_point % 8 != 0 || offset_pixel % 8 != 0)
    return Fail(ctx, NLO_EINVAL, "record layout must be 8-byte aligned");
  if (offset_local_point + 24 > stride || offset_pixel + 16 > stride) return Fail(ctx, NLO_EINVAL, "field outside the record");
  if (ctx->IsMulti()) return multi::UploadReprojAos(ctx, pr, n, records, stride, offset_local_point, offset_pixel, intrinsics);
  return UploadReprojAos(ctx, pr, n, records, stride, offset_local_point, offset_pixel, intrinsics);
}

int nlo_ingest_stats(const nlo_context* ctx, double* total_ms, double* host_gather_ms) {
  if (ctx == nullptr) return NLO_EINVAL;
  // multi-device: the shards are ingested concurrently, the slowest one is what the caller waited for
  double t = ctx->last_ingest_ms, g = ctx->last_ingest_gather_ms;
  for (const nlo_context* s : ctx->subs) {
    t = std::max(t, s->last_ingest_ms);
    g = std::max(g, s->last_ingest_gather_ms);
  }
  if (total_ms) *total_ms = t;
  if (host_gather_ms) *host_gather_ms = g;
  return NLO_OK;
}

// ---- assembly / solves ----
static int AssembleAny(nlo_context* ctx, nlo_problem* pr, int kind, int32_t problem_index, const double pose[16],
                       int64_t begin, int64_t end, double* H, int nh, double* g, int ng, double* cost) {
  if (ctx == nullptr || pr == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  if (ctx->IsMulti()) return multi::Assemble(ctx, pr, kind, problem_index, pose, begin, end, H, nh, g, ng, cost);
  return AssembleImpl(ctx, pr, kind, problem_index, pose, begin, end, H, nh, g, ng, cost);
}
static int SolveAny(nlo_context* ctx, nlo_problem* pr, int kind, const nlo_solve_options* options, double* poses,
                    nlo_solve_result* results, double* trace, bool batched_call) {
  if (ctx == nullptr || pr == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  if (ctx->IsMulti()) return multi::Solve(ctx, pr, kind, options, poses, results, trace, batched_call);
  return SolveImpl(ctx, pr, kind, options, poses, results, trace, batched_call);
}

int nlo_ndt6_assemble(nlo_context* ctx, nlo_problem* pr, int32_t problem_index, const double pose[16], int64_t begin,
                      int64_t end, double H21[21], double g[6], double* cost) {
  return AssembleAny(ctx, pr, kNdt6, problem_index, pose, begin, end, H21, 21, g, 6, cost);
}
int nlo_ndt3_assemble(nlo_context* ctx, nlo_problem* pr, int32_t problem_index, const double pose[16], int64_t begin,
                      int64_t end, double H6[6], double g[3], double* cost) {
  return AssembleAny(ctx, pr, kNdt3, problem_index, pose, begin, end, H6, 6, g, 3, cost);
}
int nlo_reproj_assemble(nlo_context* ctx, nlo_problem* pr, int32_t problem_index, const double pose[16],
                        int64_t begin, int64_t end, double H21[21], double g[6], double* cost) {
  return AssembleAny(ctx, pr, kReproj, problem_index, pose, begin, end, H21, 21, g, 6, cost);
}

int nlo_ndt6_solve(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double pose[16],
                   nlo_solve_result* result, double* trace) {
  return SolveAny(ctx, pr, kNdt6, options, pose, result, trace, false);
}
int nlo_ndt3_solve(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double pose[16],
                   nlo_solve_result* result, double* trace) {
  return SolveAny(ctx, pr, kNdt3, options, pose, result, trace, false);
}
int nlo_reproj_solve(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double pose[16],
                     nlo_solve_result* result, double* trace) {
  return SolveAny(ctx, pr, kReproj, options, pose, result, trace, false);
}
int nlo_ndt6_solve_batched(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double* poses,
                           nlo_solve_result* results) {
  return SolveAny(ctx, pr, kNdt6, options, poses, results, nullptr, true);
}
int nlo_ndt3_solve_batched(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double* poses,
                           nlo_solve_result* results) {
  return SolveAny(ctx, pr, kNdt3, options, poses, results, nullptr, true);
}
int nlo_reproj_solve_batched(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double* poses,
                             nlo_solve_result* results) {
  return SolveAny(ctx, pr, kReproj, options, poses, results, nullptr, true);
}

// ---- communicators (one process per GPU) ----
int nlo_comm_unique_id(nlo_context* ctx, uint8_t id[128]) {
  if (ctx == nullptr || id == nullptr) return NLO_EINVAL;
  if (ctx->IsMulti()) return Fail(ctx, NLO_EINVAL, "a multi-device context shards inside the process; no communicator");
  std::string err;
  if (!LoadNccl(&ctx->nccl, &err)) return Fail(ctx, NLO_ECOMM, err);
  NcclUniqueId uid;
  const int rc = ctx->nccl.GetUniqueId(&uid);
  if (rc != 0) return Fail(ctx, NLO_ECOMM, "ncclGetUniqueId failed");
  memcpy(id, uid.internal, 128);
  return NLO_OK;
}

int nlo_comm_init_nccl(nlo_context* ctx, const uint8_t id[128], int32_t rank, int32_t nranks) {
  if (ctx == nullptr || id == nullptr || nranks < 1 || rank < 0 || rank >= nranks) return Fail(ctx, NLO_EINVAL, "bad rank");
  if (ctx->IsMulti()) return Fail(ctx, NLO_EINVAL, "a multi-device context shards inside the process; no communicator");
  if (ctx->comm_kind != kCommNone) return Fail(ctx, NLO_EINVAL, "a communicator is already attached");
  std::string err;
  if (!LoadNccl(&ctx->nccl, &err)) return Fail(ctx, NLO_ECOMM, err);
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclUniqueId uid;
  memcpy(uid.internal, id, 128);
  const int rc = ctx->nccl.CommInitRank(&ctx->nccl_comm, nranks, uid, rank);
  if (rc != 0)
    return Fail(ctx, NLO_ECOMM, std::string("ncclCommInitRank: ") +
                                    (ctx->nccl.GetErrorString ? ctx->nccl.GetErrorString(rc) : "error"));
  ctx->comm_kind = kCommNccl;
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->generation++;
  return NLO_OK;
}

int nlo_comm_peer_export(nlo_context* ctx, uint8_t handle[64]) {
  if (ctx == nullptr || handle == nullptr) return NLO_EINVAL;
  if (ctx->IsMulti()) return Fail(ctx, NLO_EINVAL, "a multi-device context shards inside the process; no communicator");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->peer_buf == nullptr) {
    NLO_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->peer_buf), kPeerBufBytes));
    NLO_CUDA(ctx, cudaMemset(ctx->peer_buf, 0, kPeerBufBytes));
    NLO_CUDA(ctx, DevMalloc(reinterpret_cast<void**>(&ctx->d_peer_seq), sizeof(unsigned long long)));
    NLO_CUDA(ctx, cudaMemset(ctx->d_peer_seq, 0, sizeof(unsigned long long)));
    NLO_CUDA(ctx, DevMalloc(reinterpret_cast<void**>(&ctx->d_peer_error), sizeof(int)));
    NLO_CUDA(ctx, cudaMemset(ctx->d_peer_error, 0, sizeof(int)));
    NLO_CUDA(ctx, cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  NLO_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->peer_buf));
  memcpy(handle, &h, 64);
  return NLO_OK;
}

int nlo_comm_peer_init(nlo_context* ctx, const uint8_t* handles, int32_t rank, int32_t nranks) {
  if (ctx == nullptr || handles == nullptr || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
    return Fail(ctx, NLO_EINVAL, "bad rank / nranks (max 8)");
  if (ctx->IsMulti()) return Fail(ctx, NLO_EINVAL, "a multi-device context shards inside the process; no communicator");
  if (ctx->comm_kind != kCommNone) return Fail(ctx, NLO_EINVAL, "a communicator is already attached");
  if (ctx->peer_buf == nullptr) return Fail(ctx, NLO_EINVAL, "call nlo_comm_peer_export first");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  PeerComm pc;
  memset(&pc, 0, sizeof(pc));
  pc.rank = rank;
  pc.nranks = nranks;
  for (int r = 0; r < nranks; ++r) {
    unsigned char* base = nullptr;
    if (r == rank) {
      base = ctx->peer_buf;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + 64 * r, 64);
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
        return Fail(ctx, NLO_ECOMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
      ctx->peer_opened[r] = ptr;
      base = static_cast<unsigned char*>(ptr);
    }
    pc.slots[r] = reinterpret_cast<unsigned long long*>(base);
  }
  pc.seq = ctx->d_peer_seq;
  pc.error = ctx->d_peer_error;
  ctx->peer = pc;
  ctx->comm_kind = kCommPeer;
  ctx->rank = rank;
  ctx->nranks = nranks;
  ctx->generation++;
  return NLO_OK;
}

int nlo_problem_set_global_range(nlo_context* ctx, nlo_problem* pr, int64_t global_begin, int64_t global_total) {
  if (ctx == nullptr || pr == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  if (ctx->IsMulti() || !pr->shards.empty()) return Fail(ctx, NLO_EINVAL, "a multi-device context shards by itself");
  if (pr->batched) return Fail(ctx, NLO_EINVAL, "a batched problem is sharded by registration, not by point range");
  if (global_total < 0) {
    pr->ndt3_end_override = -1;
    return NLO_OK;
  }
  if (global_begin < 0 || global_begin > global_total) return Fail(ctx, NLO_EINVAL, "bad global range");
  pr->ndt3_end_override = std::max<int64_t>(0, (global_total / 4) * 4 - global_begin);
  return NLO_OK;
}

int nlo_comm_suspend(nlo_context* ctx, int32_t suspended) {
  if (ctx == nullptr) return NLO_EINVAL;
  if (ctx->IsMulti()) return Fail(ctx, NLO_EINVAL, "a multi-device context shards inside the process; no communicator");
  if ((suspended != 0) != ctx->comm_suspended) {
    ctx->comm_suspended = suspended != 0;
    ctx->generation++;
  }
  return NLO_OK;
}

int nlo_comm_destroy(nlo_context* ctx) {
  if (ctx == nullptr) return NLO_EINVAL;
  if (ctx->IsMulti()) return Fail(ctx, NLO_EINVAL, "a multi-device context shards inside the process; no communicator");
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->comm_kind == kCommNccl && ctx->nccl_comm != nullptr) {
    ctx->nccl.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  if (!ctx->comm_in_process) {
    for (int r = 0; r < kMaxRanks; ++r) {
      if (ctx->peer_opened[r] != nullptr) {
        cudaIpcCloseMemHandle(ctx->peer_opened[r]);
        ctx->peer_opened[r] = nullptr;
      }
    }
  }
  if (ctx->peer_buf) { cudaFree(ctx->peer_buf); ctx->peer_buf = nullptr; }
  if (ctx->d_peer_seq) { DevFree(ctx->d_peer_seq); ctx->d_peer_seq = nullptr; }
  if (ctx->d_peer_error) { DevFree(ctx->d_peer_error); ctx->d_peer_error = nullptr; }
  ctx->comm_kind = kCommNone;
  ctx->comm_in_process = false;
  ctx->comm_suspended = false;
  ctx->rank = 0;
  ctx->nranks = 1;
  ctx->generation++;
  return NLO_OK;
}

}  // extern "C"
