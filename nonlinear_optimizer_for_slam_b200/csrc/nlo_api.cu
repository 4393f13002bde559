// nlo_api.cu -- the C ABI of include/nlo_cuda.h on top of the kernels in nlo_kernels.cu.
//
// Host-side responsibilities only: device memory for the SoA planes and the per-registration
// state, upload/repack, CUDA-graph capture of the device-resident iteration loop, the optional
// communicators (NCCL through dlopen, or CUDA-IPC peer memory for the fused one-shot all-reduce),
// and the final pose / iteration-count read-back.  No arithmetic of the hot path runs on the host
// and there is no CPU fallback.
#include <dlfcn.h>

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/nlo_cuda.h"
#include "nlo_internal.h"

using namespace nlo;

namespace {

// ---- NCCL through dlopen (no link-time dependency; the symbols are only needed multi-GPU) ----
struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
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

enum CommKind { kCommNone = 0, kCommNccl = 1, kCommPeer = 2 };

constexpr size_t kPeerBufBytes = 2 * kMaxRanks * kPeerWords * sizeof(unsigned long long);

constexpr int kInCtaTiles = 3;  // a registration this small runs its whole loop inside one CTA
constexpr int kSmallDoubles = 8192;  // pinned + device scratch for poses / sums / results

}  // namespace

struct nlo_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 0;
  int grid_single = 0;
  int loss_kind = NLO_LOSS_NONE;
  double loss_params[2] = {0.0, 0.0};
  std::string error;
  void* staging = nullptr;
  size_t staging_bytes = 0;
  double* host_small = nullptr;  // pinned
  bool use_graph = true;
  bool use_persistent = true;
  int stage_depth = 0;  // NLO_STAGE_DEPTH: ring depth override (0 = per-shape default)
  double l2_keep_mb = 0.0;     // NLO_L2_KEEP_MB: bytes of a re-read scan pinned in L2 (0 = off)
  double l2_policy_min_mb = 0.0;  // only scans larger than this get an explicit policy
  int grid_small = 0;  // CTAs of the persistent path for L2-resident problems
  // communicator
  int comm_kind = kCommNone;
  int rank = 0, nranks = 1;
  NcclApi nccl;
  NcclComm nccl_comm = nullptr;
  unsigned char* peer_buf = nullptr;  // local exchange buffer (exported over CUDA IPC)
  void* peer_opened[kMaxRanks] = {nullptr};
  PeerComm peer{};
  unsigned long long* d_peer_seq = nullptr;
  int* d_peer_error = nullptr;
  nlo_problem* reg_workspace = nullptr;  // correspondences of nlo_ndt_register: kept across scans,
  int64_t reg_workspace_capacity = 0;    // grown on demand (no per-frame allocation)
  unsigned long long* d_debug_times = nullptr;  // NLO_DEBUG_TIMES=1: [64][8] stamps of the last loop
  int generation = 0;  // bumped whenever cached graphs become stale (loss / comm change)
};

struct nlo_problem {
  int family = 0;  // 0 NDT, 1 reprojection
  int num_planes = 0;
  int64_t capacity = 0;  // padded, per plane
  int64_t n = 0;
  int num_problems = 1;
  bool batched = false;
  bool f32 = false;  // planes stored as float (NDT, single problem only): fp32 storage, fp64 math
  double* plane_block = nullptr;
  double* planes[kNdtPlanes] = {nullptr};
  std::vector<Range> h_ranges;
  std::vector<int64_t> counts;
  Range* d_ranges = nullptr;  // [num_problems + 1]; the last slot is the scratch range for assemble
  State* d_states = nullptr;  // [num_problems + 1]
  double* d_partials = nullptr;
  unsigned int* d_tickets = nullptr;  // [num_problems + 1]
  unsigned long long* d_sync = nullptr;  // [num_problems + 1][kSyncStride] persistent-path counter + LL state
  double* d_sums = nullptr;           // [(num_problems + 1) * 32]
  double* d_poses = nullptr;          // [(num_problems + 1) * 16]
  double* d_results = nullptr;        // [(num_problems + 1) * 4]
  double* d_trace = nullptr;
  size_t trace_doubles = 0;
  double intrinsics[6] = {0, 0, 0, 0, 0, 0};
  std::map<std::array<int64_t, 10>, cudaGraphExec_t> graphs;
};

struct nlo_ndt_map {
  double origin[3] = {0, 0, 0};
  int dims[3] = {0, 0, 0};
  double voxel = 0.0;
  int64_t cells = 0;
  double* d_mean = nullptr;          // [cells][3]
  double* d_sqrt_info = nullptr;     // [cells][9] row-major
  unsigned char* d_valid = nullptr;  // [cells]
  // sparse map: cells = slots of the voxel hash, dims = bounding box in voxels
  unsigned long long* d_keys = nullptr;  // [cells], kHashEmpty = free slot
  long long hash_mask = 0;
};

struct nlo_scan {
  int64_t n = 0;
  double* block = nullptr;
  double* planes[3] = {nullptr, nullptr, nullptr};
  unsigned long long* d_matched = nullptr;
};

namespace {

int Fail(nlo_context* ctx, int code, const std::string& msg) {
  if (ctx != nullptr) ctx->error = msg;
  return code;
}

#define NLO_CUDA(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      return Fail((ctx), (_e == cudaErrorMemoryAllocation) ? NLO_ENOMEM : NLO_ECUDA,     \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
    }                                                                                    \
  } while (0)

int EnsureStaging(nlo_context* ctx, size_t bytes) {
  if (bytes <= ctx->staging_bytes) return NLO_OK;
  if (ctx->staging != nullptr) {
    NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NLO_CUDA(ctx, cudaFree(ctx->staging));
    ctx->staging = nullptr;
    ctx->staging_bytes = 0;
  }
  NLO_CUDA(ctx, cudaMalloc(&ctx->staging, bytes));
  ctx->staging_bytes = bytes;
  return NLO_OK;
}

void DropGraphs(nlo_problem* pr) {
  for (auto& kv : pr->graphs) cudaGraphExecDestroy(kv.second);
  pr->graphs.clear();
}

int64_t PadToTile(int64_t n) { return ((n + kTile - 1) / kTile) * kTile; }

int CreateProblem(nlo_context* ctx, int family, int num_problems, const int64_t* counts,
                  bool batched, nlo_problem** out, bool f32 = false) {
  if (ctx == nullptr || out == nullptr || num_problems < 1) return Fail(ctx, NLO_EINVAL, "bad argument");
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
  NLO_CUDA_P(cudaMalloc(&pr->plane_block, plane_bytes * pr->num_planes));
  NLO_CUDA_P(cudaMemsetAsync(pr->plane_block, 0, plane_bytes * pr->num_planes, ctx->stream));
  // tile-interleaved layout: plane k of tile 0 starts at k * 256 (nlo_internal.h TiledOffset)
  for (int k = 0; k < pr->num_planes; ++k)
    pr->planes[k] = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(pr->plane_block) +
                                              static_cast<size_t>(k) * kTile * elem);
  const int slots = num_problems + 1;
  NLO_CUDA_P(cudaMalloc(&pr->d_ranges, slots * sizeof(Range)));
  NLO_CUDA_P(cudaMalloc(&pr->d_states, slots * sizeof(State)));
  NLO_CUDA_P(cudaMalloc(&pr->d_partials,
                        2 * static_cast<size_t>(std::max(ctx->grid_single, 1)) * kAcc6 * sizeof(double)));
  NLO_CUDA_P(cudaMalloc(&pr->d_sync, static_cast<size_t>(slots) * kSyncStride * sizeof(unsigned long long)));
  NLO_CUDA_P(cudaMemsetAsync(pr->d_sync, 0, static_cast<size_t>(slots) * kSyncStride * sizeof(unsigned long long), ctx->stream));
  NLO_CUDA_P(cudaMalloc(&pr->d_tickets, slots * sizeof(unsigned int)));
  NLO_CUDA_P(cudaMemsetAsync(pr->d_tickets, 0, slots * sizeof(unsigned int), ctx->stream));
  NLO_CUDA_P(cudaMalloc(&pr->d_sums, slots * 32 * sizeof(double)));
  NLO_CUDA_P(cudaMemsetAsync(pr->d_sums, 0, slots * 32 * sizeof(double), ctx->stream));
  NLO_CUDA_P(cudaMalloc(&pr->d_poses, slots * 16 * sizeof(double)));
  NLO_CUDA_P(cudaMalloc(&pr->d_results, slots * 4 * sizeof(double)));
  NLO_CUDA_P(cudaMemcpyAsync(pr->d_ranges, pr->h_ranges.data(), num_problems * sizeof(Range),
                             cudaMemcpyHostToDevice, ctx->stream));
  NLO_CUDA_P(cudaStreamSynchronize(ctx->stream));
#undef NLO_CUDA_P
  *out = pr;
  return NLO_OK;
}

void PoseToRt(const double pose[16], double R[9], double t[3]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R[3 * r + c] = pose[4 * c + r];
  t[0] = pose[12];
  t[1] = pose[13];
  t[2] = pose[14];
}

IterParams BaseParams(nlo_context* ctx, nlo_problem* pr) {
  IterParams p;
  memset(&p, 0, sizeof(p));
  for (int k = 0; k < pr->num_planes; ++k) p.planes[k] = pr->planes[k];
  p.partials = pr->d_partials;
  p.sync_words = pr->d_sync;
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
  p.use_peer = (ctx->comm_kind == kCommPeer) ? 1 : 0;
  p.peer = ctx->peer;
  return p;
}

bool UsePersistent(const nlo_context* ctx, const nlo_problem* pr) {
  return ctx->use_persistent && ctx->comm_kind != kCommNccl && !pr->batched;
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

int CheckPeerError(nlo_context* ctx) {
  if (ctx->comm_kind != kCommPeer) return NLO_OK;
  int err = 0;
  NLO_CUDA(ctx, cudaMemcpy(&err, ctx->d_peer_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (err != 0) return Fail(ctx, NLO_ECOMM, "peer all-reduce timed out waiting for a rank");
  return NLO_OK;
}

int Assemble(nlo_context* ctx, nlo_problem* pr, int kind, int problem_index, const double pose[16],
             int64_t begin, int64_t end, double* H, int nh, double* g, int ng, double* cost) {
  if (ctx == nullptr || pr == nullptr || pose == nullptr || H == nullptr || g == nullptr || cost == nullptr)
    return Fail(ctx, NLO_EINVAL, "null argument");
  if ((kind == kReproj) != (pr->family == 1)) return Fail(ctx, NLO_EINVAL, "problem family mismatch");
  if (problem_index < 0 || problem_index >= pr->num_problems) return Fail(ctx, NLO_EINVAL, "bad problem_index");
  const int64_t count = pr->batched ? pr->counts[problem_index] : pr->n;
  if (begin < 0 || end < begin || end > count) return Fail(ctx, NLO_EINVAL, "bad [begin, end)");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
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
  NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, grid_x, 1, ctx->stream));
  if (ctx->comm_kind == kCommNccl) {
    const int rc = NcclAllReduceSums(ctx, pr->d_sums + 32 * slot);
    if (rc != NLO_OK) return rc;
  }
  double* out = hs + 64;
  NLO_CUDA(ctx, cudaMemcpyAsync(out, pr->d_sums + 32 * slot, 32 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(H, out, nh * sizeof(double));
  memcpy(g, out + nh, ng * sizeof(double));
  *cost = out[nh + ng];
  return CheckPeerError(ctx);
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
  const int64_t tiles = (end_abs + kTile - 1) / kTile - begin_abs / kTile;
  const bool in_cta_loop =
      (ctx->comm_kind == kCommNone) && (pr->batched || tiles <= kInCtaTiles);
  if (pr->batched && ctx->comm_kind == kCommNone && ctx->use_persistent && tiles > kInCtaTiles &&
      2 * num_problems <= ctx->grid_single) {
    // A small batch: one CTA per registration would leave most SMs idle, so every registration
    // gets G = (2 x SMs) / B CTAs of ONE persistent cooperative grid (gridDim.y = registrations),
    // each with its own leader CTA, arrival counter and published state.
    const int gx = static_cast<int>(std::min<int64_t>(ctx->grid_single / num_problems, tiles));
    IterParams q = p;
    q.mode = kModeSolve;
    q.persistent = 1;
    q.iterations_in_kernel = opt.max_iterations;
    NLO_CUDA(ctx, cudaMemsetAsync(pr->d_sync, 0, static_cast<size_t>(num_problems) * kSyncStride * sizeof(unsigned long long),
                                  ctx->stream));
    const cudaError_t ce = LaunchIteration(kind, ctx->loss_kind, q, gx, num_problems, ctx->stream);
    if (ce == cudaSuccess) return NLO_OK;
    if (ce != cudaErrorCooperativeLaunchTooLarge && ce != cudaErrorLaunchOutOfResources)
      return Fail(ctx, NLO_ECUDA, std::string("cooperative launch: ") + cudaGetErrorString(ce));
    cudaGetLastError();  // not co-resident right now: fall through to one CTA per registration
  }
  if (in_cta_loop) {
    // whole loop inside one CTA per registration: a single launch
    p.mode = kModeSolve;
    p.iterations_in_kernel = opt.max_iterations;
    NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, 1, num_problems, ctx->stream));
    return NLO_OK;
  }
  const int grid_x = GridFor(ctx, begin_abs, end_abs);
  if (UsePersistent(ctx, pr)) {
    // persistent cooperative grid: the whole loop in ONE launch, one grid barrier per iteration
    int gx = grid_x;
    if (tiles < 4LL * ctx->grid_single) gx = static_cast<int>(std::min<int64_t>(tiles, ctx->grid_small));
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
    const cudaError_t ce = LaunchIteration(kind, ctx->loss_kind, p, gx, 1, ctx->stream);
    if (ce == cudaSuccess) return NLO_OK;
    if (ce != cudaErrorCooperativeLaunchTooLarge && ce != cudaErrorLaunchOutOfResources)
      return Fail(ctx, NLO_ECUDA, std::string("cooperative launch: ") + cudaGetErrorString(ce));
    // the grid cannot be co-resident right now (e.g. the GPU is shared): one launch per iteration
    cudaGetLastError();
    p.persistent = 0;
    p.iterations_in_kernel = 1;
    p.l2_keep_tiles = 0;
  }
  for (int it = 0; it < opt.max_iterations; ++it) {
    if (ctx->comm_kind == kCommNccl) {
      p.mode = kModeAssemble;
      NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, grid_x, num_problems, ctx->stream));
      const int rc = NcclAllReduceSums(ctx, pr->d_sums);
      if (rc != NLO_OK) return rc;
      p.mode = kModeStepOnly;
      NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, 1, num_problems, ctx->stream));
    } else {
      p.mode = kModeSolve;
      NLO_CUDA(ctx, LaunchIteration(kind, ctx->loss_kind, p, grid_x, num_problems, ctx->stream));
    }
  }
  return NLO_OK;
}

int RunLoop(nlo_context* ctx, nlo_problem* pr, int kind, const nlo_solve_options& opt,
            bool with_trace, int num_problems, int64_t begin_abs, int64_t end_abs) {
  // NCCL calls are left out of graph capture (their capture support depends on the library
  // build); the other paths run as one CUDA graph so the loop needs a single host call.
  const int64_t tiles_all = (end_abs + kTile - 1) / kTile - begin_abs / kTile;
  const bool single_launch = UsePersistent(ctx, pr) ||
                             ((ctx->comm_kind == kCommNone) && (pr->batched || tiles_all <= kInCtaTiles));
  const bool graph = ctx->use_graph && ctx->comm_kind != kCommNccl && !single_launch;
  if (!graph) return EnqueueLoop(ctx, pr, kind, opt, with_trace, num_problems, begin_abs, end_abs);
  int64_t ptol_bits, gtol_bits;
  memcpy(&ptol_bits, &opt.parameter_tolerance, 8);
  memcpy(&gtol_bits, &opt.gradient_tolerance, 8);
  const std::array<int64_t, 10> key = {kind, ctx->loss_kind, ctx->comm_kind, opt.max_iterations,
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

int Solve(nlo_context* ctx, nlo_problem* pr, int kind, const nlo_solve_options* options,
          double* poses, nlo_solve_result* results, double* trace, bool batched_call) {
  if (ctx == nullptr || pr == nullptr || options == nullptr || poses == nullptr || results == nullptr)
    return Fail(ctx, NLO_EINVAL, "null argument");
  if ((kind == kReproj) != (pr->family == 1)) return Fail(ctx, NLO_EINVAL, "problem family mismatch");
  if (options->max_iterations < 0) return Fail(ctx, NLO_EINVAL, "max_iterations < 0");
  if (batched_call != pr->batched) return Fail(ctx, NLO_EINVAL, "batched / single problem mismatch");
  const int B = pr->num_problems;
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int trace_width = (kind == kNdt3) ? NLO_TRACE3 : NLO_TRACE6;
  const bool with_trace = (trace != nullptr) && !pr->batched && options->max_iterations > 0;
  if (with_trace) {
    const size_t need = static_cast<size_t>(options->max_iterations) * trace_width;
    if (need > pr->trace_doubles) {
      if (pr->d_trace) NLO_CUDA(ctx, cudaFree(pr->d_trace));
      pr->d_trace = nullptr;
      NLO_CUDA(ctx, cudaMalloc(&pr->d_trace, need * sizeof(double)));
      pr->trace_doubles = need;
      DropGraphs(pr);
    }
    NLO_CUDA(ctx, cudaMemsetAsync(pr->d_trace, 0, need * sizeof(double), ctx->stream));
  }
  // ranges of the registrations (single problem: [0, n) resp. floor(n/4)*4 for the 3-DoF path,
  // ..._analytic_3dof.cc:33-36)
  int64_t begin_abs = 0, end_abs = 0;
  if (!pr->batched) {
    Range r{0, pr->n};
    if (kind == kNdt3) r.end = (pr->n / 4) * 4;
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
  NLO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  if (options->max_iterations > 0) {
    const int rc = RunLoop(ctx, pr, kind, *options, with_trace, B, begin_abs, end_abs);
    if (rc != NLO_OK) return rc;
  }
  NLO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  NLO_CUDA(ctx, LaunchFinishStates(pr->d_states, pr->d_poses, pr->d_results, B, kind, ctx->stream));
  std::vector<double> res(static_cast<size_t>(B) * 4);
  NLO_CUDA(ctx, cudaMemcpyAsync(poses, pr->d_poses, static_cast<size_t>(B) * 16 * sizeof(double),
                                cudaMemcpyDeviceToHost, ctx->stream));
  NLO_CUDA(ctx, cudaMemcpyAsync(res.data(), pr->d_results, static_cast<size_t>(B) * 4 * sizeof(double),
                                cudaMemcpyDeviceToHost, ctx->stream));
  if (with_trace)
    NLO_CUDA(ctx, cudaMemcpyAsync(trace, pr->d_trace,
                                  static_cast<size_t>(options->max_iterations) * trace_width * sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->stream));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  NLO_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  if (ctx->d_debug_times != nullptr && options->max_iterations >= 8 && options->max_iterations <= 64) {
    std::vector<unsigned long long> ts(64 * 8);
    cudaMemcpy(ts.data(), ctx->d_debug_times, ts.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    // average phase durations over iterations 2..7 (ns): tiles | cta-sync | barrier | x-cta sum | step | sync
    double acc[6] = {0, 0, 0, 0, 0, 0};
    double period = 0;
    for (int it = 2; it < 8; ++it) {
      for (int k = 0; k < 6; ++k) acc[k] += static_cast<double>(ts[it * 8 + k + 1]) - static_cast<double>(ts[it * 8 + k]);
      period += static_cast<double>(ts[(it + 1) * 8]) - static_cast<double>(ts[it * 8]);
    }
    fprintf(stderr, "[nlo debug] ns/iter: tiles %.0f | cta-sync %.0f | barrier %.0f | x-cta-sum %.0f | canon+step %.0f | sync %.0f | period %.0f\n",
            acc[0] / 6, acc[1] / 6, acc[2] / 6, acc[3] / 6, acc[4] / 6, acc[5] / 6, period / 6);
  }
  int rc_all = NLO_OK;
  for (int k = 0; k < B; ++k) {
    results[k].iterations = static_cast<int32_t>(res[4 * k]);
    results[k].status = (res[4 * k + 1] != 0.0) ? NLO_ENUMERIC : NLO_OK;
    results[k].final_cost = res[4 * k + 2];
    results[k].device_ms = ms;
    if (results[k].status != NLO_OK) rc_all = NLO_ENUMERIC;
  }
  const int pe = CheckPeerError(ctx);
  if (pe != NLO_OK) return pe;
  if (rc_all != NLO_OK) return Fail(ctx, rc_all, "non-finite value met during the solve");
  return NLO_OK;
}

}  // namespace

extern "C" {

int nlo_abi_version(void) { return NLO_ABI_VERSION; }

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
  const char* denv = getenv("NLO_DEBUG_TIMES");
  if (denv != nullptr && denv[0] == '1') {
    cudaMalloc(reinterpret_cast<void**>(&ctx->d_debug_times), 64 * 8 * sizeof(unsigned long long));
    cudaMemset(ctx->d_debug_times, 0, 64 * 8 * sizeof(unsigned long long));
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

int nlo_context_destroy(nlo_context* ctx) {
  if (ctx == nullptr) return NLO_OK;
  cudaSetDevice(ctx->device);
  nlo_comm_destroy(ctx);
  if (ctx->reg_workspace) nlo_problem_destroy(ctx, ctx->reg_workspace);
  ctx->reg_workspace = nullptr;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->staging) cudaFree(ctx->staging);
  if (ctx->d_debug_times) cudaFree(ctx->d_debug_times);
  if (ctx->host_small) cudaFreeHost(ctx->host_small);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return NLO_OK;
}

const char* nlo_last_error(const nlo_context* ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int nlo_context_info(const nlo_context* ctx, int* sm_count, int* assemble_grid) {
  if (ctx == nullptr) return NLO_EINVAL;
  if (sm_count) *sm_count = ctx->sm_count;
  if (assemble_grid) *assemble_grid = ctx->grid_single;
  return NLO_OK;
}

int nlo_synchronize(nlo_context* ctx) {
  if (ctx == nullptr) return NLO_EINVAL;
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
      p1 = 1.0 / (p0 * p0);  // the kernels multiply by 1 / c^2 instead of dividing per correspondence
      break;
    default: return Fail(ctx, NLO_EINVAL, "unknown loss kind");
  }
  ctx->loss_kind = kind;
  ctx->loss_params[0] = p0;
  ctx->loss_params[1] = p1;
  ctx->generation++;
  return NLO_OK;
}

int nlo_host_alloc(void** ptr, size_t bytes) {
  if (ptr == nullptr) return NLO_EINVAL;
  return cudaMallocHost(ptr, bytes) == cudaSuccess ? NLO_OK : NLO_ENOMEM;
}
int nlo_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? NLO_OK : NLO_ECUDA; }

int nlo_ndt_create(nlo_context* ctx, int64_t capacity, nlo_problem** problem) {
  if (capacity < 0) return Fail(ctx, NLO_EINVAL, "negative capacity");
  return CreateProblem(ctx, 0, 1, &capacity, false, problem);
}

int nlo_ndt_create_f32(nlo_context* ctx, int64_t capacity, nlo_problem** problem) {
  if (capacity < 0) return Fail(ctx, NLO_EINVAL, "negative capacity");
  return CreateProblem(ctx, 0, 1, &capacity, false, problem, true);
}

int nlo_ndt_create_batched(nlo_context* ctx, int32_t num_problems, const int64_t* counts, nlo_problem** problem) {
  if (counts == nullptr) return Fail(ctx, NLO_EINVAL, "null counts");
  return CreateProblem(ctx, 0, num_problems, counts, true, problem);
}

int nlo_reproj_create(nlo_context* ctx, int64_t capacity, nlo_problem** problem) {
  if (capacity < 0) return Fail(ctx, NLO_EINVAL, "negative capacity");
  return CreateProblem(ctx, 1, 1, &capacity, false, problem);
}

int nlo_reproj_create_batched(nlo_context* ctx, int32_t num_problems, const int64_t* counts, nlo_problem** problem) {
  if (counts == nullptr) return Fail(ctx, NLO_EINVAL, "null counts");
  return CreateProblem(ctx, 1, num_problems, counts, true, problem);
}

int nlo_problem_destroy(nlo_context* ctx, nlo_problem* pr) {
  if (pr == nullptr) return NLO_OK;
  if (ctx != nullptr) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  DropGraphs(pr);
  cudaFree(pr->plane_block);
  cudaFree(pr->d_ranges);
  cudaFree(pr->d_states);
  cudaFree(pr->d_partials);
  cudaFree(pr->d_tickets);
  cudaFree(pr->d_sync);
  cudaFree(pr->d_sums);
  cudaFree(pr->d_poses);
  cudaFree(pr->d_results);
  cudaFree(pr->d_trace);
  delete pr;
  return NLO_OK;
}

int64_t nlo_problem_size(const nlo_problem* pr) { return pr ? pr->n : -1; }

int nlo_ndt_upload(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* point, const double* mean,
                   const double* sqrt_info) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0) return Fail(ctx, NLO_EINVAL, "bad problem");
  if (n < 0 || (n > 0 && (point == nullptr || mean == nullptr || sqrt_info == nullptr)))
    return Fail(ctx, NLO_EINVAL, "null array");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  int64_t total = 0;
  for (int64_t c : pr->counts) total += c;
  if (pr->batched ? (n != total) : (n > pr->counts[0])) return Fail(ctx, NLO_EINVAL, "n does not fit the problem");
  const size_t bytes = static_cast<size_t>(n) * 15 * sizeof(double);
  const size_t extra = pr->batched ? (pr->num_problems + 1) * sizeof(int64_t) : 0;
  int rc = EnsureStaging(ctx, bytes + extra + 256);
  if (rc != NLO_OK) return rc;
  double* s_point = static_cast<double*>(ctx->staging);
  double* s_mean = s_point + 3 * n;
  double* s_sqrt = s_mean + 3 * n;
  if (n > 0) {
    NLO_CUDA(ctx, cudaMemcpyAsync(s_point, point, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NLO_CUDA(ctx, cudaMemcpyAsync(s_mean, mean, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NLO_CUDA(ctx, cudaMemcpyAsync(s_sqrt, sqrt_info, 9 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  if (!pr->batched) {
    NLO_CUDA(ctx, LaunchPackNdt(s_point, s_mean, s_sqrt, n, pr->planes, 0, pr->f32, ctx->stream));
    pr->n = n;
    pr->h_ranges[0] = Range{0, n};
  } else {
    std::vector<int64_t> prefix(pr->num_problems + 1, 0);
    for (int k = 0; k < pr->num_problems; ++k) prefix[k + 1] = prefix[k] + pr->counts[k];
    int64_t* d_prefix = reinterpret_cast<int64_t*>(reinterpret_cast<unsigned char*>(ctx->staging) +
                                                   ((bytes + 255) / 256) * 256);
    NLO_CUDA(ctx, cudaMemcpyAsync(d_prefix, prefix.data(), prefix.size() * sizeof(int64_t),
                                  cudaMemcpyHostToDevice, ctx->stream));
    NLO_CUDA(ctx, LaunchPackNdtBatched(s_point, s_mean, s_sqrt, n, d_prefix, pr->d_ranges,
                                       pr->num_problems, pr->planes, ctx->stream));
    pr->n = n;
    NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // prefix vector goes out of scope
  }
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NLO_OK;
}

int nlo_ndt_upload_f32(nlo_context* ctx, nlo_problem* pr, int64_t n, const float* point, const float* mean,
                       const float* sqrt_info) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0 || pr->batched || !pr->f32)
    return Fail(ctx, NLO_EINVAL, "bad problem (needs an fp32-storage NDT problem)");
  if (n < 0 || n > pr->counts[0] || (n > 0 && (point == nullptr || mean == nullptr || sqrt_info == nullptr)))
    return Fail(ctx, NLO_EINVAL, "bad n / null array");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = EnsureStaging(ctx, static_cast<size_t>(n) * 15 * sizeof(float) + 256);
  if (rc != NLO_OK) return rc;
  float* s_point = static_cast<float*>(ctx->staging);
  float* s_mean = s_point + 3 * n;
  float* s_sqrt = s_mean + 3 * n;
  if (n > 0) {
    NLO_CUDA(ctx, cudaMemcpyAsync(s_point, point, 3 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    NLO_CUDA(ctx, cudaMemcpyAsync(s_mean, mean, 3 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    NLO_CUDA(ctx, cudaMemcpyAsync(s_sqrt, sqrt_info, 9 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    NLO_CUDA(ctx, LaunchPackNdtFromFloat(s_point, s_mean, s_sqrt, n, pr->planes, ctx->stream));
  }
  pr->n = n;
  pr->h_ranges[0] = Range{0, n};
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NLO_OK;
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
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = static_cast<size_t>(n) * stride;
  int rc = EnsureStaging(ctx, bytes + 256);
  if (rc != NLO_OK) return rc;
  if (n > 0) {
    NLO_CUDA(ctx, cudaMemcpyAsync(ctx->staging, records, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NLO_CUDA(ctx, LaunchPackNdtAos(static_cast<const unsigned char*>(ctx->staging), n, stride, offset_point,
                                   offset_mean, offset_sqrt_info, col_major, pr->planes, ctx->stream));
  }
  pr->n = n;
  pr->h_ranges[0] = Range{0, n};
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NLO_OK;
}

namespace {
int GenerateCommon(nlo_context* ctx, nlo_problem* pr, uint64_t seed, int64_t global_index_offset,
                   double noise_sigma, const double* true_poses, const double init_pose[16],
                   const double grid_origin[3], const int32_t grid_dims[3], double voxel_size,
                   const double* cell_mean, const double* cell_sqrt_info, const uint8_t* cell_valid,
                   int64_t n_single) {
  if (!true_poses || !init_pose || !grid_origin || !grid_dims || !cell_mean || !cell_sqrt_info || !cell_valid ||
      !(voxel_size > 0.0))
    return Fail(ctx, NLO_EINVAL, "null / bad grid argument");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t cells = static_cast<size_t>(grid_dims[0]) * grid_dims[1] * grid_dims[2];
  const size_t bytes = cells * (12 * sizeof(double) + 1) + 512;
  int rc = EnsureStaging(ctx, bytes);
  if (rc != NLO_OK) return rc;
  double* d_mean = static_cast<double*>(ctx->staging);
  double* d_sqrt = d_mean + 3 * cells;
  unsigned char* d_valid = reinterpret_cast<unsigned char*>(d_sqrt + 9 * cells);
  NLO_CUDA(ctx, cudaMemcpyAsync(d_mean, cell_mean, 3 * cells * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NLO_CUDA(ctx, cudaMemcpyAsync(d_sqrt, cell_sqrt_info, 9 * cells * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NLO_CUDA(ctx, cudaMemcpyAsync(d_valid, cell_valid, cells, cudaMemcpyHostToDevice, ctx->stream));
  GenerateParams g;
  memset(&g, 0, sizeof(g));
  g.noise_sigma = noise_sigma;
  PoseToRt(init_pose, g.R_init, g.t_init);
  for (int k = 0; k < 3; ++k) {
    g.origin[k] = grid_origin[k];
    g.dims[k] = grid_dims[k];
  }
  g.inv_voxel = 1.0 / voxel_size;
  g.reach = std::min(4, static_cast<int>(std::ceil(1.0 / voxel_size)));
  g.cell_mean = d_mean;
  g.cell_sqrt_info = d_sqrt;
  g.cell_valid = d_valid;
  const int B = pr->batched ? pr->num_problems : 1;
  int64_t total = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t begin = pr->batched ? pr->h_ranges[b].begin : 0;
    const int64_t n = pr->batched ? pr->counts[b] : n_single;
    for (int k = 0; k < pr->num_planes; ++k) g.planes[k] = pr->planes[k];
    g.dst_offset = begin;
    g.f32 = pr->f32 ? 1 : 0;
    g.n = n;
    g.seed = seed + static_cast<uint64_t>(b);
    g.index_offset = pr->batched ? 0 : global_index_offset;
    PoseToRt(true_poses + 16 * static_cast<size_t>(b), g.R_true, g.t_true);
    NLO_CUDA(ctx, LaunchGenerateNdt(g, ctx->stream));
    total += n;
  }
  pr->n = total;
  if (!pr->batched) pr->h_ranges[0] = Range{0, n_single};
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NLO_OK;
}
}  // namespace

int nlo_ndt_generate(nlo_context* ctx, nlo_problem* pr, int64_t n, uint64_t seed, int64_t global_index_offset,
                     double noise_sigma, const double true_pose[16], const double init_pose[16],
                     const double grid_origin[3], const int32_t grid_dims[3], double voxel_size,
                     const double* cell_mean, const double* cell_sqrt_info, const uint8_t* cell_valid) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0 || pr->batched) return Fail(ctx, NLO_EINVAL, "bad problem");
  if (n < 0 || n > pr->counts[0]) return Fail(ctx, NLO_EINVAL, "n exceeds capacity");
  return GenerateCommon(ctx, pr, seed, global_index_offset, noise_sigma, true_pose, init_pose, grid_origin,
                        grid_dims, voxel_size, cell_mean, cell_sqrt_info, cell_valid, n);
}

int nlo_ndt_generate_batched(nlo_context* ctx, nlo_problem* pr, uint64_t seed, double noise_sigma,
                             const double* true_poses, const double init_pose[16], const double grid_origin[3],
                             const int32_t grid_dims[3], double voxel_size, const double* cell_mean,
                             const double* cell_sqrt_info, const uint8_t* cell_valid) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0 || !pr->batched) return Fail(ctx, NLO_EINVAL, "bad problem");
  return GenerateCommon(ctx, pr, seed, 0, noise_sigma, true_poses, init_pose, grid_origin, grid_dims,
                        voxel_size, cell_mean, cell_sqrt_info, cell_valid, 0);
}

int nlo_ndt_download(nlo_context* ctx, const nlo_problem* pr, int64_t begin, int64_t end, double* point,
                     double* mean, double* information) {
  if (ctx == nullptr || pr == nullptr || pr->family != 0 || pr->batched) return Fail(ctx, NLO_EINVAL, "bad problem");
  if (begin < 0 || end < begin || end > pr->n) return Fail(ctx, NLO_EINVAL, "bad [begin, end)");
  if (!point || !mean || !information) return Fail(ctx, NLO_EINVAL, "null array");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = end - begin;
  int rc = EnsureStaging(ctx, static_cast<size_t>(n) * 12 * sizeof(double) + 256);
  if (rc != NLO_OK) return rc;
  double* s_point = static_cast<double*>(ctx->staging);
  double* s_mean = s_point + 3 * n;
  double* s_info = s_mean + 3 * n;
  NLO_CUDA(ctx, LaunchUnpackNdt(const_cast<double* const*>(pr->planes), begin, end, s_point, s_mean, s_info,
                                pr->f32, ctx->stream));
  if (n > 0) {
    NLO_CUDA(ctx, cudaMemcpyAsync(point, s_point, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NLO_CUDA(ctx, cudaMemcpyAsync(mean, s_mean, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NLO_CUDA(ctx, cudaMemcpyAsync(information, s_info, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NLO_OK;
}

int nlo_reproj_upload(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* local_point, const double* pixel,
                      const double intrinsics[6]) {
  if (ctx == nullptr || pr == nullptr || pr->family != 1) return Fail(ctx, NLO_EINVAL, "bad problem");
  int64_t total = 0;
  for (int64_t c : pr->counts) total += c;
  if (n < 0 || (pr->batched ? (n != total) : (n > pr->counts[0])) || intrinsics == nullptr ||
      (n > 0 && (!local_point || !pixel)))
    return Fail(ctx, NLO_EINVAL, "bad argument");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = static_cast<size_t>(n) * 5 * sizeof(double);
  int rc = EnsureStaging(ctx, bytes + (pr->num_problems + 1) * sizeof(int64_t) + 512);
  if (rc != NLO_OK) return rc;
  double* s_point = static_cast<double*>(ctx->staging);
  double* s_pixel = s_point + 3 * n;
  std::vector<int64_t> prefix(pr->num_problems + 1, 0);
  if (n > 0) {
    NLO_CUDA(ctx, cudaMemcpyAsync(s_point, local_point, 3 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NLO_CUDA(ctx, cudaMemcpyAsync(s_pixel, pixel, 2 * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (!pr->batched) {
      NLO_CUDA(ctx, LaunchPackReproj(s_point, s_pixel, n, pr->planes, ctx->stream));
    } else {
      for (int k = 0; k < pr->num_problems; ++k) prefix[k + 1] = prefix[k] + pr->counts[k];
      int64_t* d_prefix = reinterpret_cast<int64_t*>(reinterpret_cast<unsigned char*>(ctx->staging) +
                                                     ((bytes + 255) / 256) * 256);
      NLO_CUDA(ctx, cudaMemcpyAsync(d_prefix, prefix.data(), prefix.size() * sizeof(int64_t),
                                    cudaMemcpyHostToDevice, ctx->stream));
      NLO_CUDA(ctx, LaunchPackReprojBatched(s_point, s_pixel, n, d_prefix, pr->d_ranges, pr->num_problems,
                                            pr->planes, ctx->stream));
      NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  for (int k = 0; k < 6; ++k) pr->intrinsics[k] = intrinsics[k];
  pr->n = n;
  if (!pr->batched) pr->h_ranges[0] = Range{0, n};
  DropGraphs(pr);  // intrinsics are baked into captured launches
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NLO_OK;
}

int nlo_ndt6_assemble(nlo_context* ctx, nlo_problem* pr, int32_t problem_index, const double pose[16], int64_t begin,
                      int64_t end, double H21[21], double g[6], double* cost) {
  return Assemble(ctx, pr, kNdt6, problem_index, pose, begin, end, H21, 21, g, 6, cost);
}
int nlo_ndt3_assemble(nlo_context* ctx, nlo_problem* pr, int32_t problem_index, const double pose[16], int64_t begin,
                      int64_t end, double H6[6], double g[3], double* cost) {
  return Assemble(ctx, pr, kNdt3, problem_index, pose, begin, end, H6, 6, g, 3, cost);
}
int nlo_reproj_assemble(nlo_context* ctx, nlo_problem* pr, int32_t problem_index, const double pose[16],
                        int64_t begin, int64_t end, double H21[21], double g[6], double* cost) {
  return Assemble(ctx, pr, kReproj, problem_index, pose, begin, end, H21, 21, g, 6, cost);
}

int nlo_ndt6_solve(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double pose[16],
                   nlo_solve_result* result, double* trace) {
  return Solve(ctx, pr, kNdt6, options, pose, result, trace, false);
}
int nlo_ndt3_solve(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double pose[16],
                   nlo_solve_result* result, double* trace) {
  return Solve(ctx, pr, kNdt3, options, pose, result, trace, false);
}
int nlo_reproj_solve(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double pose[16],
                     nlo_solve_result* result, double* trace) {
  return Solve(ctx, pr, kReproj, options, pose, result, trace, false);
}
int nlo_ndt6_solve_batched(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double* poses,
                           nlo_solve_result* results) {
  return Solve(ctx, pr, kNdt6, options, poses, results, nullptr, true);
}

int nlo_ndt3_solve_batched(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double* poses,
                           nlo_solve_result* results) {
  return Solve(ctx, pr, kNdt3, options, poses, results, nullptr, true);
}
int nlo_reproj_solve_batched(nlo_context* ctx, nlo_problem* pr, const nlo_solve_options* options, double* poses,
                             nlo_solve_result* results) {
  return Solve(ctx, pr, kReproj, options, poses, results, nullptr, true);
}

// ---- NDT map / scan / matcher / outer registration loop ----
namespace {

constexpr int64_t kMaxMapCells = 1LL << 28;  // dense cells, or occupied voxels of a hashed map

// hash_slots == 0: dense grid of dims cells; otherwise a voxel hash with that many slots (2^k).
int AllocMap(nlo_context* ctx, const double origin[3], const int32_t dims[3], double voxel, int64_t hash_slots,
             nlo_ndt_map** out) {
  if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0 || !(voxel > 0.0)) return Fail(ctx, NLO_EINVAL, "bad grid");
  int64_t cells = hash_slots;
  if (hash_slots == 0) {
    // the product cannot overflow: each factor is first checked against the limit
    if (dims[0] > kMaxMapCells || dims[1] > kMaxMapCells || dims[2] > kMaxMapCells ||
        static_cast<int64_t>(dims[0]) * dims[1] > kMaxMapCells ||
        static_cast<int64_t>(dims[0]) * dims[1] * dims[2] > kMaxMapCells)
      return Fail(ctx, NLO_EINVAL, "dense grid too large (> 2^28 cells); build a hashed map");
    cells = static_cast<int64_t>(dims[0]) * dims[1] * dims[2];
  } else {
    for (int k = 0; k < 3; ++k)
      if (dims[k] > (1 << kHashAxisBits)) return Fail(ctx, NLO_EINVAL, "map spans more than 2^21 voxels on an axis");
  }
  nlo_ndt_map* m = new nlo_ndt_map();
  for (int k = 0; k < 3; ++k) { m->origin[k] = origin[k]; m->dims[k] = dims[k]; }
  m->voxel = voxel;
  m->cells = cells;
  m->hash_mask = hash_slots ? hash_slots - 1 : 0;
  if (cudaMalloc(&m->d_mean, cells * 3 * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&m->d_sqrt_info, cells * 9 * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&m->d_valid, cells) != cudaSuccess ||
      (hash_slots && cudaMalloc(&m->d_keys, cells * sizeof(unsigned long long)) != cudaSuccess)) {
    nlo_ndt_map_destroy(ctx, m);
    return Fail(ctx, NLO_ENOMEM, "cudaMalloc(map) failed");
  }
  *out = m;
  return NLO_OK;
}

int MatchInto(nlo_context* ctx, const nlo_scan* scan, const nlo_ndt_map* map, const double pose[16], double radius,
              int max_neighbors, nlo_problem* pr, unsigned long long* d_matched) {
  MatchParams mp;
  memset(&mp, 0, sizeof(mp));
  for (int k = 0; k < 3; ++k) mp.scan[k] = scan->planes[k];
  mp.n = scan->n;
  for (int k = 0; k < pr->num_planes; ++k) mp.planes[k] = pr->planes[k];
  PoseToRt(pose, mp.R, mp.t);
  for (int k = 0; k < 3; ++k) { mp.origin[k] = map->origin[k]; mp.dims[k] = map->dims[k]; }
  mp.inv_voxel = 1.0 / map->voxel;
  mp.reach = static_cast<int>(std::ceil(radius / map->voxel));
  mp.radius2 = radius * radius;
  mp.max_neighbors = max_neighbors;
  mp.cell_mean = map->d_mean;
  mp.cell_sqrt_info = map->d_sqrt_info;
  mp.cell_valid = map->d_valid;
  mp.keys = map->d_keys;
  mp.hash_mask = map->hash_mask;
  mp.matched = d_matched;
  if (d_matched) NLO_CUDA(ctx, cudaMemsetAsync(d_matched, 0, sizeof(unsigned long long), ctx->stream));
  NLO_CUDA(ctx, LaunchMatchNdt(mp, ctx->stream));
  pr->n = static_cast<int64_t>(max_neighbors) * scan->n;
  pr->h_ranges[0] = Range{0, pr->n};
  return NLO_OK;
}

// Eigen::Quaterniond(Matrix3d) restated on the host for the outer-loop convergence test
void HostRotToQuat(const double* R, double* q) {
  double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) {
    double s = std::sqrt(tr + 1.0);
    q[3] = 0.5 * s; s = 0.5 / s;
    q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * s; s = 0.5 / s;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * s;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * s;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
  }
}

}  // namespace

int nlo_ndt_map_create(nlo_context* ctx, const double grid_origin[3], const int32_t grid_dims[3], double voxel_size,
                       const double* cell_mean, const double* cell_sqrt_info, const uint8_t* cell_valid,
                       nlo_ndt_map** map) {
  if (ctx == nullptr || map == nullptr || !grid_origin || !grid_dims || !cell_mean || !cell_sqrt_info || !cell_valid)
    return Fail(ctx, NLO_EINVAL, "null argument");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  nlo_ndt_map* m = nullptr;
  int rc = AllocMap(ctx, grid_origin, grid_dims, voxel_size, 0, &m);
  if (rc != NLO_OK) return rc;
  cudaError_t e = cudaMemcpyAsync(m->d_mean, cell_mean, m->cells * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_sqrt_info, cell_sqrt_info, m->cells * 9 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_valid, cell_valid, m->cells, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    nlo_ndt_map_destroy(ctx, m);
    return Fail(ctx, NLO_ECUDA, std::string("map upload: ") + cudaGetErrorString(e));
  }
  *map = m;
  return NLO_OK;
}

namespace {

int64_t NextPow2(int64_t v) {
  int64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// layout: 0 dense, 1 hashed, 2 dense unless the bounding box has more than kMaxMapCells voxels
int BuildMap(nlo_context* ctx, int64_t n, const double* points_xyz, double voxel_size, int v_not_transposed,
             int layout, nlo_ndt_map** map) {
  if (ctx == nullptr || map == nullptr || points_xyz == nullptr || n <= 0 || !(voxel_size > 0.0))
    return Fail(ctx, NLO_EINVAL, "bad argument");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t pbytes = static_cast<size_t>(n) * 3 * sizeof(double);
  int rc = EnsureStaging(ctx, pbytes + 256);
  if (rc != NLO_OK) return rc;
  double* d_xyz = static_cast<double*>(ctx->staging);
  int* d_bounds = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ctx->staging) + ((pbytes + 63) / 64) * 64);
  NLO_CUDA(ctx, cudaMemcpyAsync(d_xyz, points_xyz, pbytes, cudaMemcpyHostToDevice, ctx->stream));
  int* hb = reinterpret_cast<int*>(ctx->host_small);
  for (int k = 0; k < 3; ++k) { hb[k] = INT_MAX; hb[3 + k] = INT_MIN; }
  NLO_CUDA(ctx, cudaMemcpyAsync(d_bounds, hb, 6 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  const double inv = 1.0 / voxel_size;
  NLO_CUDA(ctx, LaunchMapBounds(d_xyz, n, inv, d_bounds, ctx->stream));
  NLO_CUDA(ctx, cudaMemcpyAsync(hb, d_bounds, 6 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int32_t dims[3];
  double origin[3];
  int kmin[3];
  double box_cells = 1.0;
  for (int k = 0; k < 3; ++k) {
    const int64_t span = static_cast<int64_t>(hb[3 + k]) - hb[k] + 1;
    if (span > (1LL << 30)) return Fail(ctx, NLO_EINVAL, "points span more than 2^30 voxels on an axis");
    kmin[k] = hb[k];
    dims[k] = static_cast<int32_t>(span);
    origin[k] = hb[k] * voxel_size;
    box_cells *= static_cast<double>(span);
  }
  const bool hashed = layout == 1 || (layout == 2 && box_cells > static_cast<double>(kMaxMapCells));

  int64_t slots = 0;
  if (hashed) {
    for (int k = 0; k < 3; ++k)
      if (dims[k] > (1 << kHashAxisBits)) return Fail(ctx, NLO_EINVAL, "map spans more than 2^21 voxels on an axis");
    // pass 1: distinct occupied voxels, through a scratch key table of >= 2 n slots
    const int64_t scratch_slots = NextPow2(std::max<int64_t>(2 * n, 1024));
    unsigned long long* d_scratch = nullptr;
    if (cudaMalloc(&d_scratch, (scratch_slots + 1) * sizeof(unsigned long long)) != cudaSuccess) {
      cudaGetLastError();
      return Fail(ctx, NLO_ENOMEM, "cudaMalloc(voxel hash scratch) failed");
    }
    unsigned long long* d_distinct = d_scratch + scratch_slots;
    cudaError_t e = cudaMemsetAsync(d_scratch, 0xff, scratch_slots * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_distinct, 0, sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess)
      e = LaunchMapCountVoxels(d_xyz, n, inv, kmin, d_scratch, scratch_slots - 1, d_distinct, ctx->stream);
    unsigned long long distinct = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->host_small, d_distinct, sizeof(distinct), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_scratch);
    if (e != cudaSuccess) return Fail(ctx, NLO_ECUDA, std::string("map build (voxel count): ") + cudaGetErrorString(e));
    memcpy(&distinct, ctx->host_small, sizeof(distinct));
    if (static_cast<int64_t>(distinct) > kMaxMapCells) return Fail(ctx, NLO_EINVAL, "more than 2^28 occupied voxels");
    slots = NextPow2(std::max<int64_t>(2 * static_cast<int64_t>(distinct), 1024));
  }

  nlo_ndt_map* m = nullptr;
  rc = AllocMap(ctx, origin, dims, voxel_size, slots, &m);
  if (rc != NLO_OK) return rc;
  int* d_count = nullptr;
  double* d_sums = nullptr;
  auto cleanup = [&]() { cudaFree(d_count); cudaFree(d_sums); };
  cudaError_t e = cudaMalloc(&d_count, m->cells * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&d_sums, m->cells * 9 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, m->cells * sizeof(int), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_sums, 0, m->cells * 9 * sizeof(double), ctx->stream);
  if (e == cudaSuccess && hashed) e = cudaMemsetAsync(m->d_keys, 0xff, m->cells * sizeof(unsigned long long), ctx->stream);
  MapAccumParams ap;
  memset(&ap, 0, sizeof(ap));
  ap.xyz = d_xyz; ap.n = n; ap.inv_voxel = inv; ap.voxel = voxel_size;
  for (int k = 0; k < 3; ++k) { ap.kmin[k] = kmin[k]; ap.dims[k] = dims[k]; }
  ap.count = d_count; ap.sums = d_sums;
  ap.keys = m->d_keys; ap.hash_mask = m->hash_mask;
  if (e == cudaSuccess) e = LaunchMapAccumulate(ap, ctx->stream);
  if (e == cudaSuccess)
    e = LaunchMapFinalize(ap, m->cells, v_not_transposed, m->d_mean, m->d_sqrt_info, m->d_valid, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess) {
    nlo_ndt_map_destroy(ctx, m);
    return Fail(ctx, e == cudaErrorMemoryAllocation ? NLO_ENOMEM : NLO_ECUDA,
                std::string("map build: ") + cudaGetErrorString(e));
  }
  *map = m;
  return NLO_OK;
}

}  // namespace

int nlo_ndt_map_build(nlo_context* ctx, int64_t n, const double* points_xyz, double voxel_size, int v_not_transposed,
                      nlo_ndt_map** map) {
  return BuildMap(ctx, n, points_xyz, voxel_size, v_not_transposed, 2, map);
}

int nlo_ndt_map_build_hashed(nlo_context* ctx, int64_t n, const double* points_xyz, double voxel_size,
                             int v_not_transposed, nlo_ndt_map** map) {
  return BuildMap(ctx, n, points_xyz, voxel_size, v_not_transposed, 1, map);
}

int nlo_ndt_map_layout(nlo_context* ctx, const nlo_ndt_map* map, int32_t* hashed, int64_t* cells) {
  if (ctx == nullptr || map == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  if (hashed) *hashed = map->d_keys != nullptr ? 1 : 0;
  if (cells) *cells = map->cells;
  return NLO_OK;
}

int nlo_ndt_map_download_keys(nlo_context* ctx, const nlo_ndt_map* map, uint64_t* slot_keys) {
  if (ctx == nullptr || map == nullptr || slot_keys == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  if (map->d_keys == nullptr) return Fail(ctx, NLO_EINVAL, "not a hashed map");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  NLO_CUDA(ctx, cudaMemcpy(slot_keys, map->d_keys, map->cells * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return NLO_OK;
}

int nlo_ndt_map_info(nlo_context* ctx, const nlo_ndt_map* map, double grid_origin[3], int32_t grid_dims[3],
                     double* voxel_size, int64_t* valid_cells) {
  if (ctx == nullptr || map == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  for (int k = 0; k < 3; ++k) {
    if (grid_origin) grid_origin[k] = map->origin[k];
    if (grid_dims) grid_dims[k] = map->dims[k];
  }
  if (voxel_size) *voxel_size = map->voxel;
  if (valid_cells) {
    NLO_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<unsigned char> v(map->cells);
    NLO_CUDA(ctx, cudaMemcpy(v.data(), map->d_valid, map->cells, cudaMemcpyDeviceToHost));
    int64_t c = 0;
    for (unsigned char x : v) c += x ? 1 : 0;
    *valid_cells = c;
  }
  return NLO_OK;
}

int nlo_ndt_map_download(nlo_context* ctx, const nlo_ndt_map* map, double* cell_mean, double* cell_sqrt_info,
                         uint8_t* cell_valid) {
  if (ctx == nullptr || map == nullptr || !cell_mean || !cell_sqrt_info || !cell_valid)
    return Fail(ctx, NLO_EINVAL, "null argument");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  NLO_CUDA(ctx, cudaMemcpy(cell_mean, map->d_mean, map->cells * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  NLO_CUDA(ctx, cudaMemcpy(cell_sqrt_info, map->d_sqrt_info, map->cells * 9 * sizeof(double), cudaMemcpyDeviceToHost));
  NLO_CUDA(ctx, cudaMemcpy(cell_valid, map->d_valid, map->cells, cudaMemcpyDeviceToHost));
  return NLO_OK;
}

int nlo_ndt_map_destroy(nlo_context* ctx, nlo_ndt_map* map) {
  if (map == nullptr) return NLO_OK;
  if (ctx != nullptr) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  cudaFree(map->d_mean);
  cudaFree(map->d_sqrt_info);
  cudaFree(map->d_valid);
  cudaFree(map->d_keys);
  delete map;
  return NLO_OK;
}

int nlo_scan_create(nlo_context* ctx, int64_t n, const double* points_xyz, nlo_scan** scan) {
  if (ctx == nullptr || scan == nullptr || n < 0 || (n > 0 && points_xyz == nullptr)) return Fail(ctx, NLO_EINVAL, "bad argument");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  nlo_scan* sc = new nlo_scan();
  sc->n = n;
  const int64_t cap = std::max<int64_t>(n, 1);
  // one allocation: three planes + the matched-correspondence counter behind them
  if (cudaMalloc(&sc->block, (cap * 3 + 8) * sizeof(double)) != cudaSuccess) {
    nlo_scan_destroy(ctx, sc);
    return Fail(ctx, NLO_ENOMEM, "cudaMalloc(scan) failed");
  }
  sc->d_matched = reinterpret_cast<unsigned long long*>(sc->block + cap * 3);
  for (int k = 0; k < 3; ++k) sc->planes[k] = sc->block + static_cast<size_t>(k) * cap;
  if (n > 0) {
    int rc = EnsureStaging(ctx, static_cast<size_t>(n) * 3 * sizeof(double));
    if (rc != NLO_OK) { nlo_scan_destroy(ctx, sc); return rc; }
    cudaError_t e = cudaMemcpyAsync(ctx->staging, points_xyz, static_cast<size_t>(n) * 3 * sizeof(double),
                                    cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = LaunchPackScan(static_cast<const double*>(ctx->staging), n, sc->planes, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      nlo_scan_destroy(ctx, sc);
      return Fail(ctx, NLO_ECUDA, std::string("scan upload: ") + cudaGetErrorString(e));
    }
  }
  *scan = sc;
  return NLO_OK;
}

int nlo_scan_destroy(nlo_context* ctx, nlo_scan* scan) {
  if (scan == nullptr) return NLO_OK;
  if (ctx != nullptr) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  cudaFree(scan->block);
  delete scan;
  return NLO_OK;
}

int nlo_ndt_match(nlo_context* ctx, const nlo_scan* scan, const nlo_ndt_map* map, const double pose[16], double radius,
                  int32_t max_neighbors, nlo_problem* pr, int64_t* matched) {
  if (ctx == nullptr || scan == nullptr || map == nullptr || pose == nullptr || pr == nullptr)
    return Fail(ctx, NLO_EINVAL, "null argument");
  if (pr->family != 0 || pr->batched || pr->f32) return Fail(ctx, NLO_EINVAL, "problem must be a single fp64 NDT problem");
  if (max_neighbors < 1 || max_neighbors > 2 || !(radius > 0.0)) return Fail(ctx, NLO_EINVAL, "bad radius / max_neighbors");
  if (static_cast<int64_t>(max_neighbors) * scan->n > pr->counts[0]) return Fail(ctx, NLO_EINVAL, "problem capacity too small");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = MatchInto(ctx, scan, map, pose, radius, max_neighbors, pr, scan->d_matched);
  if (rc != NLO_OK) return rc;
  unsigned long long m = 0;
  NLO_CUDA(ctx, cudaMemcpyAsync(ctx->host_small, scan->d_matched, sizeof(m), cudaMemcpyDeviceToHost, ctx->stream));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(&m, ctx->host_small, sizeof(m));
  if (matched) *matched = static_cast<int64_t>(m);
  return NLO_OK;
}

int nlo_ndt_register(nlo_context* ctx, const nlo_scan* scan_in, const nlo_ndt_map* map, const nlo_solve_options* options,
                     double radius, int32_t max_neighbors, int32_t max_outer, int32_t three_dof, double pose[16],
                     nlo_register_result* result) {
  if (ctx == nullptr || scan_in == nullptr || map == nullptr || options == nullptr || pose == nullptr || result == nullptr)
    return Fail(ctx, NLO_EINVAL, "null argument");
  if (max_neighbors < 1 || max_neighbors > 2 || !(radius > 0.0) || max_outer < 0) return Fail(ctx, NLO_EINVAL, "bad argument");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  nlo_scan* scan = const_cast<nlo_scan*>(scan_in);
  const int64_t need = static_cast<int64_t>(max_neighbors) * scan->n;
  if (ctx->reg_workspace == nullptr || ctx->reg_workspace_capacity < need) {
    if (ctx->reg_workspace) nlo_problem_destroy(ctx, ctx->reg_workspace);
    ctx->reg_workspace = nullptr;
    const int64_t cap = std::max<int64_t>(need + need / 2, 4096);
    int rc = nlo_ndt_create(ctx, cap, &ctx->reg_workspace);
    if (rc != NLO_OK) return rc;
    ctx->reg_workspace_capacity = cap;
  }
  nlo_problem* pr = ctx->reg_workspace;
  memset(result, 0, sizeof(*result));
  cudaEvent_t e0, e1;
  NLO_CUDA(ctx, cudaEventCreate(&e0));
  NLO_CUDA(ctx, cudaEventCreate(&e1));
  NLO_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  int rc = NLO_OK;
  for (int outer = 0; outer < max_outer; ++outer) {
    double last[16];
    memcpy(last, pose, sizeof(last));
    rc = MatchInto(ctx, scan, map, pose, radius, max_neighbors, pr, scan->d_matched);
    if (rc != NLO_OK) break;
    nlo_solve_result sr;
    rc = Solve(ctx, pr, three_dof ? kNdt3 : kNdt6, options, pose, &sr, nullptr, false);
    if (rc != NLO_OK) break;
    result->outer_iterations = outer + 1;
    result->inner_iterations += sr.iterations;
    if (sr.iterations > 0 || outer == 0) result->final_cost = sr.final_cost;
    // :495-499  pose_diff = last^-1 * current; stop on |dt| < 1e-5 and |dq.vec| < 1e-5
    double Rl[9], tl[3], Rc[9], tc[3], Rd[9], td[3], q[4];
    PoseToRt(last, Rl, tl);
    PoseToRt(pose, Rc, tc);
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        Rd[3 * r + c] = Rl[r] * Rc[c] + Rl[3 + r] * Rc[3 + c] + Rl[6 + r] * Rc[6 + c];
      td[r] = Rl[r] * (tc[0] - tl[0]) + Rl[3 + r] * (tc[1] - tl[1]) + Rl[6 + r] * (tc[2] - tl[2]);
    }
    HostRotToQuat(Rd, q);
    const double dt = std::sqrt(td[0] * td[0] + td[1] * td[1] + td[2] * td[2]);
    const double dq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (dt < 1e-5 && dq < 1e-5) break;
  }
  cudaEventRecord(e1, ctx->stream);
  unsigned long long m = 0;
  cudaMemcpyAsync(ctx->host_small, scan->d_matched, sizeof(m), cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  memcpy(&m, ctx->host_small, sizeof(m));
  result->matched = static_cast<int64_t>(m);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  result->device_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  result->status = rc;
  return rc;
}

// ---- communicators ----
int nlo_comm_unique_id(nlo_context* ctx, uint8_t id[128]) {
  if (ctx == nullptr || id == nullptr) return NLO_EINVAL;
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
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->peer_buf == nullptr) {
    NLO_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->peer_buf), kPeerBufBytes));
    NLO_CUDA(ctx, cudaMemset(ctx->peer_buf, 0, kPeerBufBytes));
    NLO_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_peer_seq), sizeof(unsigned long long)));
    NLO_CUDA(ctx, cudaMemset(ctx->d_peer_seq, 0, sizeof(unsigned long long)));
    NLO_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&ctx->d_peer_error), sizeof(int)));
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

int nlo_comm_destroy(nlo_context* ctx) {
  if (ctx == nullptr) return NLO_EINVAL;
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->comm_kind == kCommNccl && ctx->nccl_comm != nullptr) {
    ctx->nccl.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  for (int r = 0; r < kMaxRanks; ++r) {
    if (ctx->peer_opened[r] != nullptr) {
      cudaIpcCloseMemHandle(ctx->peer_opened[r]);
      ctx->peer_opened[r] = nullptr;
    }
  }
  if (ctx->peer_buf) { cudaFree(ctx->peer_buf); ctx->peer_buf = nullptr; }
  if (ctx->d_peer_seq) { cudaFree(ctx->d_peer_seq); ctx->d_peer_seq = nullptr; }
  if (ctx->d_peer_error) { cudaFree(ctx->d_peer_error); ctx->d_peer_error = nullptr; }
  ctx->comm_kind = kCommNone;
  ctx->rank = 0;
  ctx->nranks = 1;
  ctx->generation++;
  return NLO_OK;
}

}  // extern "C"
