// nlo_multi.cu -- one context spanning several devices of THIS process (nlo_context_create_multi).
//
// The reference's data-parallel mechanism lives inside Solve: the correspondence vector is split
// into per-thread ranges, every thread assembles its partial H | g | cost, and the partials are
// summed (mahalanobis_distance_minimizer_analytic.cc:59-73,104-119; ..._analytic_simd.cc:55-76).
// The multi-device context is the same thing with B200s in place of threads, behind the same
// calls: a single problem is split by contiguous point range over the devices at upload, every
// device runs the persistent iteration kernel on its shard, the 28 (10) doubles are summed each
// iteration by the one-shot peer-memory all-reduce fused into that kernel (the exchange buffers
// of the other devices are mapped directly -- cudaDeviceEnablePeerAccess, no IPC, no second
// process), and every device applies the identical step.  A batched problem is split by
// registration id; its shards never communicate.
//
// One persistent host thread per device issues that device's calls, so uploads overlap on the
// PCIe links and the spinning iteration kernels of all devices are launched together.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "nlo_host.h"

namespace nlo {

// ------------------------------------------------------------------ worker thread
Worker::Worker() : thread_([this]() { Loop(); }) {}

Worker::~Worker() {
  {
    std::lock_guard<std::mutex> lock(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  thread_.join();
}

void Worker::Post(std::function<void()> job) {
  {
    std::lock_guard<std::mutex> lock(mu_);
    job_ = std::move(job);
    has_job_ = true;
    busy_ = true;
  }
  cv_.notify_all();
}

void Worker::Wait() {
  std::unique_lock<std::mutex> lock(mu_);
  cv_.wait(lock, [this]() { return !busy_; });
}

void Worker::Loop() {
  std::unique_lock<std::mutex> lock(mu_);
  while (true) {
    cv_.wait(lock, [this]() { return has_job_ || stop_; });
    if (stop_) return;
    std::function<void()> job = std::move(job_);
    has_job_ = false;
    lock.unlock();
    job();
    lock.lock();
    busy_ = false;
    cv_.notify_all();
  }
}

namespace multi {

namespace {

int NumDevices(const nlo_context* ctx) { return static_cast<int>(ctx->subs.size()); }

// Runs fn(r) for every device on that device's worker thread; returns the first failure and
// copies its message to the outer context.
template <typename Fn>
int ForEach(nlo_context* ctx, Fn fn) {
  const int D = NumDevices(ctx);
  std::vector<int> rc(static_cast<size_t>(D), NLO_OK);
  if (D == 1) {
    rc[0] = fn(0);
  } else {
    for (int r = 0; r < D; ++r) ctx->workers[static_cast<size_t>(r)]->Post([&, r]() { rc[static_cast<size_t>(r)] = fn(r); });
    for (int r = 0; r < D; ++r) ctx->workers[static_cast<size_t>(r)]->Wait();
  }
  for (int r = 0; r < D; ++r)
    if (rc[static_cast<size_t>(r)] != NLO_OK) {
      ctx->error = "device " + std::to_string(ctx->subs[static_cast<size_t>(r)]->device) + ": " +
                   ctx->subs[static_cast<size_t>(r)]->error;
      return rc[static_cast<size_t>(r)];
    }
  return NLO_OK;
}

// [begin, end) of the n points owned by shard r: floor(n / D) each, the remainder on the last
// device (the same rule as sharding.point_range).
void PointRange(int64_t n, int r, int D, int64_t* begin, int64_t* end) {
  const int64_t per = n / D;
  *begin = r * per;
  *end = (r == D - 1) ? n : *begin + per;
}

// block partition of B registrations: sizes differ by at most one (sharding.problem_partition)
void ProblemRange(int B, int r, int D, int* begin, int* end) {
  const int base = B / D, extra = B % D;
  *begin = r * base + std::min(r, extra);
  *end = *begin + base + (r < extra ? 1 : 0);
}

void SetPointSplit(nlo_problem* pr, int64_t n, int D) {
  pr->shard_begin.assign(static_cast<size_t>(D) + 1, 0);
  for (int r = 0; r < D; ++r) {
    int64_t b, e;
    PointRange(n, r, D, &b, &e);
    pr->shard_begin[static_cast<size_t>(r)] = b;
    pr->shard_begin[static_cast<size_t>(r) + 1] = e;
  }
  pr->n = n;
}

int CheckSharded(nlo_context* ctx, const nlo_problem* pr) {
  if (pr == nullptr || static_cast<int>(pr->shards.size()) != NumDevices(ctx))
    return Fail(ctx, NLO_EINVAL, "the problem was not created on this multi-device context");
  return NLO_OK;
}

// source offset (in correspondences) of the first registration of shard r of a batched problem
int64_t BatchedSourceOffset(const nlo_problem* pr, int r) {
  int64_t off = 0;
  for (int64_t k = 0; k < pr->shard_begin[static_cast<size_t>(r)]; ++k) off += pr->counts[static_cast<size_t>(k)];
  return off;
}
int64_t BatchedShardCount(const nlo_problem* pr, int r) {
  int64_t c = 0;
  for (int64_t k = pr->shard_begin[static_cast<size_t>(r)]; k < pr->shard_begin[static_cast<size_t>(r) + 1]; ++k)
    c += pr->counts[static_cast<size_t>(k)];
  return c;
}

// Generic upload splitter: `one(r, sub, shard, first, count)` ingests source records
// [first, first + count) into shard r.
template <typename One>
int SplitUpload(nlo_context* ctx, nlo_problem* pr, int64_t n, One one) {
  int rc = CheckSharded(ctx, pr);
  if (rc != NLO_OK) return rc;
  const int D = NumDevices(ctx);
  if (!pr->batched) {
    if (n > pr->counts[0]) return Fail(ctx, NLO_EINVAL, "n does not fit the problem");
    SetPointSplit(pr, n, D);
    return ForEach(ctx, [&](int r) {
      const int64_t b = pr->shard_begin[static_cast<size_t>(r)], e = pr->shard_begin[static_cast<size_t>(r) + 1];
      return one(r, ctx->subs[static_cast<size_t>(r)], pr->shards[static_cast<size_t>(r)], b, e - b);
    });
  }
  int64_t total = 0;
  for (int64_t c : pr->counts) total += c;
  if (n != total) return Fail(ctx, NLO_EINVAL, "n does not fit the problem");
  pr->n = n;
  return ForEach(ctx, [&](int r) {
    nlo_problem* shard = pr->shards[static_cast<size_t>(r)];
    if (shard == nullptr) return static_cast<int>(NLO_OK);
    return one(r, ctx->subs[static_cast<size_t>(r)], shard, BatchedSourceOffset(pr, r), BatchedShardCount(pr, r));
  });
}

}  // namespace

int CreateContext(const int* devices, int n, nlo_context** out) {
  nlo_context* ctx = new nlo_context();
  ctx->device = devices[0];
  auto fail = [&](int code, const std::string& msg) {
    fprintf(stderr, "nlo_context_create_multi: %s\n", msg.c_str());
    DestroyContext(ctx);
    return code;
  };
  for (int r = 0; r < n; ++r) {
    nlo_context* sub = nullptr;
    const int rc = nlo_context_create(devices[r], &sub);
    if (rc != NLO_OK) return fail(rc, "device " + std::to_string(devices[r]) + " is not a usable sm_100 GPU");
    ctx->subs.push_back(sub);
  }
  // A device listed m times (tests on a one-GPU box) runs m iteration kernels at once; they spin on
  // each other, so all of them must be resident together: each gets 1 / m of the CTA slots.
  for (int r = 0; r < n; ++r) {
    int m = 0;
    for (int s = 0; s < n; ++s) m += devices[s] == devices[r] ? 1 : 0;
    nlo_context* sub = ctx->subs[static_cast<size_t>(r)];
    sub->grid_single = std::max(1, sub->grid_single / m);
    sub->grid_small = std::max(1, sub->grid_small / m);
    sub->device_share = m;
    if (m > 1) {  // several cooperative cluster launches do not reliably co-schedule on one device
      sub->cluster_small = 1;
    }
  }
  bool shared = false;
  for (nlo_context* sub : ctx->subs) shared = shared || sub->device_share > 1;
  if (shared) {
    ctx->launch_barrier = new HostBarrier(n);
    for (nlo_context* sub : ctx->subs) sub->launch_barrier = ctx->launch_barrier;
  }
  ctx->sm_count = ctx->subs[0]->sm_count;
  ctx->grid_single = ctx->subs[0]->grid_single;
  // the host threads of one upload are shared between the devices
  const int hw = std::max(1, static_cast<int>(std::thread::hardware_concurrency()));
  for (nlo_context* sub : ctx->subs)
    if (sub->ingest_threads == 0) sub->ingest_threads = std::max(1, std::min(hw, 32) / n);
  if (n > 1) {
    // exchange buffers, mapped directly by the peers
    for (int r = 0; r < n; ++r) {
      nlo_context* sub = ctx->subs[static_cast<size_t>(r)];
      cudaError_t e = cudaSetDevice(sub->device);
      if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&sub->peer_buf), kPeerBufBytes);
      if (e == cudaSuccess) e = cudaMemset(sub->peer_buf, 0, kPeerBufBytes);
      if (e == cudaSuccess) e = DevMalloc(reinterpret_cast<void**>(&sub->d_peer_seq), sizeof(unsigned long long));
      if (e == cudaSuccess) e = cudaMemset(sub->d_peer_seq, 0, sizeof(unsigned long long));
      if (e == cudaSuccess) e = DevMalloc(reinterpret_cast<void**>(&sub->d_peer_error), sizeof(int));
      if (e == cudaSuccess) e = cudaMemset(sub->d_peer_error, 0, sizeof(int));
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) return fail(NLO_ECUDA, std::string("exchange buffer: ") + cudaGetErrorString(e));
    }
    for (int r = 0; r < n; ++r)
      for (int s = 0; s < n; ++s) {
        const int a = ctx->subs[static_cast<size_t>(r)]->device, b = ctx->subs[static_cast<size_t>(s)]->device;
        if (a == b) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, a, b);
        if (!can) return fail(NLO_ECOMM, "device " + std::to_string(a) + " cannot map device " + std::to_string(b));
        cudaSetDevice(a);
        const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(NLO_ECOMM, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
    for (int r = 0; r < n; ++r) {
      nlo_context* sub = ctx->subs[static_cast<size_t>(r)];
      PeerComm pc;
      memset(&pc, 0, sizeof(pc));
      pc.rank = r;
      pc.nranks = n;
      for (int s = 0; s < n; ++s)
        pc.slots[s] = reinterpret_cast<unsigned long long*>(ctx->subs[static_cast<size_t>(s)]->peer_buf);
      pc.seq = sub->d_peer_seq;
      pc.error = sub->d_peer_error;
      sub->peer = pc;
      sub->comm_kind = kCommPeer;
      sub->comm_in_process = true;
      sub->rank = r;
      sub->nranks = n;
      sub->generation++;
    }
    for (int r = 0; r < n; ++r) ctx->workers.push_back(new Worker());
  }
  *out = ctx;
  return NLO_OK;
}

void DestroyContext(nlo_context* ctx) {
  for (Worker* w : ctx->workers) delete w;
  ctx->workers.clear();
  // every kernel of every device must have ended before any exchange buffer goes away
  for (nlo_context* sub : ctx->subs) {
    cudaSetDevice(sub->device);
    if (sub->stream) cudaStreamSynchronize(sub->stream);
  }
  for (nlo_context* sub : ctx->subs) nlo_context_destroy(sub);
  ctx->subs.clear();
  delete ctx->launch_barrier;
  delete ctx;
}

int SetLoss(nlo_context* ctx, int kind, const double params[2]) {
  for (nlo_context* sub : ctx->subs) {
    const int rc = nlo_set_loss(sub, kind, params);
    if (rc != NLO_OK) {
      ctx->error = sub->error;
      return rc;
    }
  }
  ctx->loss_kind = kind;
  return NLO_OK;
}

int Synchronize(nlo_context* ctx) {
  for (nlo_context* sub : ctx->subs) {
    const int rc = nlo_synchronize(sub);
    if (rc != NLO_OK) {
      ctx->error = sub->error;
      return rc;
    }
  }
  return NLO_OK;
}

int Create(nlo_context* ctx, int family, int num_problems, const int64_t* counts, bool batched, bool f32,
           nlo_problem** out) {
  if (num_problems < 1) return Fail(ctx, NLO_EINVAL, "bad argument");
  const int D = NumDevices(ctx);
  nlo_problem* pr = new nlo_problem();
  pr->family = family;
  pr->num_planes = (family == 0) ? kNdtPlanes : kReprojPlanes;
  pr->num_problems = num_problems;
  pr->batched = batched;
  pr->f32 = f32;
  for (int k = 0; k < num_problems; ++k) {
    if (counts[k] < 0) {
      delete pr;
      return Fail(ctx, NLO_EINVAL, "negative count");
    }
    pr->counts.push_back(counts[k]);
  }
  pr->shards.assign(static_cast<size_t>(D), nullptr);
  pr->shard_begin.assign(static_cast<size_t>(D) + 1, 0);
  int rc = NLO_OK;
  for (int r = 0; r < D && rc == NLO_OK; ++r) {
    nlo_context* sub = ctx->subs[static_cast<size_t>(r)];
    if (!batched) {
      // a shard holds floor(n / D) points, the last one up to D - 1 more
      const int64_t cap = counts[0] / D + D;
      rc = CreateProblem(sub, family, 1, &cap, false, &pr->shards[static_cast<size_t>(r)], f32);
    } else {
      int b, e;
      ProblemRange(num_problems, r, D, &b, &e);
      pr->shard_begin[static_cast<size_t>(r)] = b;
      pr->shard_begin[static_cast<size_t>(r) + 1] = e;
      if (e > b) rc = CreateProblem(sub, family, e - b, counts + b, true, &pr->shards[static_cast<size_t>(r)], f32);
    }
    if (rc != NLO_OK) ctx->error = sub->error;
  }
  if (rc != NLO_OK) {
    Destroy(ctx, pr);
    return rc;
  }
  *out = pr;
  return NLO_OK;
}

int Destroy(nlo_context* ctx, nlo_problem* pr) {
  if (pr == nullptr) return NLO_OK;
  for (size_t r = 0; r < pr->shards.size(); ++r) {
    nlo_context* sub = (ctx != nullptr && r < ctx->subs.size()) ? ctx->subs[r] : nullptr;
    if (pr->shards[r] != nullptr) nlo_problem_destroy(sub, pr->shards[r]);
  }
  pr->shards.clear();
  delete pr;
  return NLO_OK;
}

int UploadNdt(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* point, const double* mean,
              const double* sqrt_info) {
  return SplitUpload(ctx, pr, n, [&](int, nlo_context* sub, nlo_problem* shard, int64_t first, int64_t count) {
    return nlo::UploadNdt(sub, shard, count, point + 3 * first, mean + 3 * first, sqrt_info + 9 * first);
  });
}

int UploadNdtF32(nlo_context* ctx, nlo_problem* pr, int64_t n, const float* point, const float* mean,
                 const float* sqrt_info) {
  return SplitUpload(ctx, pr, n, [&](int, nlo_context* sub, nlo_problem* shard, int64_t first, int64_t count) {
    return nlo::UploadNdtF32(sub, shard, count, point + 3 * first, mean + 3 * first, sqrt_info + 9 * first);
  });
}

int UploadNdtAos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                 size_t offset_point, size_t offset_mean, size_t offset_sqrt_info, int col_major) {
  const unsigned char* base = static_cast<const unsigned char*>(records);
  return SplitUpload(ctx, pr, n, [&](int, nlo_context* sub, nlo_problem* shard, int64_t first, int64_t count) {
    return nlo::UploadNdtAos(sub, shard, count, base + static_cast<size_t>(first) * stride, stride, offset_point,
                             offset_mean, offset_sqrt_info, col_major);
  });
}

int UploadReproj(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* local_point, const double* pixel,
                 const double intrinsics[6]) {
  for (int k = 0; k < 6; ++k) pr->intrinsics[k] = intrinsics[k];
  return SplitUpload(ctx, pr, n, [&](int, nlo_context* sub, nlo_problem* shard, int64_t first, int64_t count) {
    return nlo::UploadReproj(sub, shard, count, local_point + 3 * first, pixel + 2 * first, intrinsics);
  });
}

int UploadReprojAos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                    size_t offset_local_point, size_t offset_pixel, const double intrinsics[6]) {
  const unsigned char* base = static_cast<const unsigned char*>(records);
  for (int k = 0; k < 6; ++k) pr->intrinsics[k] = intrinsics[k];
  return SplitUpload(ctx, pr, n, [&](int, nlo_context* sub, nlo_problem* shard, int64_t first, int64_t count) {
    return nlo::UploadReprojAos(sub, shard, count, base + static_cast<size_t>(first) * stride, stride,
                                offset_local_point, offset_pixel, intrinsics);
  });
}

int Generate(nlo_context* ctx, nlo_problem* pr, uint64_t seed, int64_t global_index_offset, double noise_sigma,
             const double* true_poses, const double init_pose[16], const double grid_origin[3],
             const int32_t grid_dims[3], double voxel_size, const double* cell_mean, const double* cell_sqrt_info,
             const uint8_t* cell_valid, int64_t n_single) {
  int rc = CheckSharded(ctx, pr);
  if (rc != NLO_OK) return rc;
  const int D = NumDevices(ctx);
  if (!pr->batched) SetPointSplit(pr, n_single, D);
  rc = ForEach(ctx, [&](int r) {
    nlo_context* sub = ctx->subs[static_cast<size_t>(r)];
    nlo_problem* shard = pr->shards[static_cast<size_t>(r)];
    if (shard == nullptr) return static_cast<int>(NLO_OK);
    const int64_t b = pr->shard_begin[static_cast<size_t>(r)], e = pr->shard_begin[static_cast<size_t>(r) + 1];
    if (!pr->batched)  // the same counter-based stream as one device would produce: point i = PRNG(seed, offset + i)
      return GenerateNdt(sub, shard, seed, global_index_offset + b, noise_sigma, true_poses, init_pose, grid_origin,
                         grid_dims, voxel_size, cell_mean, cell_sqrt_info, cell_valid, e - b);
    return GenerateNdt(sub, shard, seed + static_cast<uint64_t>(b), 0, noise_sigma, true_poses + 16 * b, init_pose,
                       grid_origin, grid_dims, voxel_size, cell_mean, cell_sqrt_info, cell_valid, 0);
  });
  if (rc == NLO_OK && pr->batched) {
    pr->n = 0;
    for (int64_t c : pr->counts) pr->n += c;
  }
  return rc;
}

int Download(nlo_context* ctx, const nlo_problem* pr, int32_t problem_index, int64_t begin, int64_t end,
             double* point, double* mean, double* information) {
  int rc = CheckSharded(ctx, pr);
  if (rc != NLO_OK) return rc;
  const int D = NumDevices(ctx);
  if (pr->batched) {
    if (problem_index < 0 || problem_index >= pr->num_problems) return Fail(ctx, NLO_EINVAL, "bad problem_index");
    for (int r = 0; r < D; ++r)
      if (problem_index >= pr->shard_begin[static_cast<size_t>(r)] && problem_index < pr->shard_begin[static_cast<size_t>(r) + 1]) {
        nlo_context* sub = ctx->subs[static_cast<size_t>(r)];
        rc = DownloadNdt(sub, pr->shards[static_cast<size_t>(r)],
                         problem_index - static_cast<int32_t>(pr->shard_begin[static_cast<size_t>(r)]), begin, end, point,
                         mean, information);
        if (rc != NLO_OK) ctx->error = sub->error;
        return rc;
      }
    return Fail(ctx, NLO_EINVAL, "bad problem_index");
  }
  if (begin < 0 || end < begin || end > pr->n) return Fail(ctx, NLO_EINVAL, "bad [begin, end)");
  for (int r = 0; r < D; ++r) {
    const int64_t b = std::max(begin, pr->shard_begin[static_cast<size_t>(r)]);
    const int64_t e = std::min(end, pr->shard_begin[static_cast<size_t>(r) + 1]);
    if (e <= b) continue;
    nlo_context* sub = ctx->subs[static_cast<size_t>(r)];
    const int64_t sb = pr->shard_begin[static_cast<size_t>(r)];
    rc = DownloadNdt(sub, pr->shards[static_cast<size_t>(r)], 0, b - sb, e - sb, point + 3 * (b - begin),
                     mean + 3 * (b - begin), information + 6 * (b - begin));
    if (rc != NLO_OK) {
      ctx->error = sub->error;
      return rc;
    }
  }
  return NLO_OK;
}

int Assemble(nlo_context* ctx, nlo_problem* pr, int kind, int problem_index, const double pose[16], int64_t begin,
             int64_t end, double* H, int nh, double* g, int ng, double* cost) {
  int rc = CheckSharded(ctx, pr);
  if (rc != NLO_OK) return rc;
  if (pose == nullptr || H == nullptr || g == nullptr || cost == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  const int D = NumDevices(ctx);
  if (pr->batched) {
    if (problem_index < 0 || problem_index >= pr->num_problems) return Fail(ctx, NLO_EINVAL, "bad problem_index");
    for (int r = 0; r < D; ++r)
      if (problem_index >= pr->shard_begin[static_cast<size_t>(r)] && problem_index < pr->shard_begin[static_cast<size_t>(r) + 1]) {
        nlo_context* sub = ctx->subs[static_cast<size_t>(r)];
        rc = AssembleImpl(sub, pr->shards[static_cast<size_t>(r)], kind,
                          problem_index - static_cast<int>(pr->shard_begin[static_cast<size_t>(r)]), pose, begin, end, H,
                          nh, g, ng, cost);
        if (rc != NLO_OK) ctx->error = sub->error;
        return rc;
      }
    return Fail(ctx, NLO_EINVAL, "bad problem_index");
  }
  if (problem_index != 0) return Fail(ctx, NLO_EINVAL, "bad problem_index");
  if (begin < 0 || end < begin || end > pr->n) return Fail(ctx, NLO_EINVAL, "bad [begin, end)");
  // every device assembles its part of [begin, end) (possibly empty) and takes part in the all-reduce
  std::vector<std::array<double, 32>> out(static_cast<size_t>(D));
  rc = ForEach(ctx, [&](int r) {
    const int64_t sb = pr->shard_begin[static_cast<size_t>(r)], se = pr->shard_begin[static_cast<size_t>(r) + 1];
    const int64_t b = std::min(std::max(begin, sb), se) - sb;
    const int64_t e = std::max(std::min(end, se), sb + b) - sb;
    double* o = out[static_cast<size_t>(r)].data();
    return AssembleImpl(ctx->subs[static_cast<size_t>(r)], pr->shards[static_cast<size_t>(r)], kind, 0, pose, b, e, o,
                        nh, o + nh, ng, o + nh + ng);
  });
  if (rc != NLO_OK) return rc;
  memcpy(H, out[0].data(), static_cast<size_t>(nh) * sizeof(double));
  memcpy(g, out[0].data() + nh, static_cast<size_t>(ng) * sizeof(double));
  *cost = out[0][static_cast<size_t>(nh + ng)];
  return NLO_OK;
}

int Solve(nlo_context* ctx, nlo_problem* pr, int kind, const nlo_solve_options* options, double* poses,
          nlo_solve_result* results, double* trace, bool batched_call) {
  int rc = CheckSharded(ctx, pr);
  if (rc != NLO_OK) return rc;
  if (options == nullptr || poses == nullptr || results == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  if (batched_call != pr->batched) return Fail(ctx, NLO_EINVAL, "batched / single problem mismatch");
  const int D = NumDevices(ctx);
  if (pr->batched) {
    // registrations are independent: every device solves its block of them, nothing is exchanged
    rc = ForEach(ctx, [&](int r) {
      nlo_problem* shard = pr->shards[static_cast<size_t>(r)];
      if (shard == nullptr) return static_cast<int>(NLO_OK);
      const int64_t b = pr->shard_begin[static_cast<size_t>(r)];
      return SolveImpl(ctx->subs[static_cast<size_t>(r)], shard, kind, options, poses + 16 * b, results + b, nullptr, true);
    });
    double ms = 0.0;  // the batch is done when the slowest device is
    for (int k = 0; k < pr->num_problems; ++k) ms = std::max(ms, results[k].device_ms);
    for (int k = 0; k < pr->num_problems; ++k) results[k].device_ms = ms;
    return rc;
  }
  // The planar minimizer drops the last n mod 4 correspondences of the WHOLE list
  // (..._analytic_3dof.cc:33-36): only the shards that reach past floor(n / 4) * 4 are cut.
  const int64_t global_end = (kind == kNdt3) ? (pr->n / 4) * 4 : pr->n;
  std::vector<std::array<double, 16>> pose(static_cast<size_t>(D));
  std::vector<nlo_solve_result> res(static_cast<size_t>(D));
  for (int r = 0; r < D; ++r) memcpy(pose[static_cast<size_t>(r)].data(), poses, 16 * sizeof(double));
  if (trace != nullptr && options->max_iterations > 0) {  // allocate before any shard's kernel runs (see EnsureTrace)
    const size_t need = static_cast<size_t>(options->max_iterations) * ((kind == kNdt3) ? NLO_TRACE3 : NLO_TRACE6);
    rc = EnsureTrace(ctx->subs[0], pr->shards[0], need);
    if (rc != NLO_OK) return Fail(ctx, rc, ctx->subs[0]->error);
  }
  rc = ForEach(ctx, [&](int r) {
    nlo_problem* shard = pr->shards[static_cast<size_t>(r)];
    const int64_t sb = pr->shard_begin[static_cast<size_t>(r)];
    shard->ndt3_end_override = std::max<int64_t>(0, std::min(global_end - sb, shard->n));
    const int code = SolveImpl(ctx->subs[static_cast<size_t>(r)], shard, kind, options, pose[static_cast<size_t>(r)].data(),
                               &res[static_cast<size_t>(r)], r == 0 ? trace : nullptr, false);
    shard->ndt3_end_override = -1;
    return code;
  });
  results[0] = res[0];
  memcpy(poses, pose[0].data(), 16 * sizeof(double));
  if (rc != NLO_OK) return rc;
  for (int r = 1; r < D; ++r) {
    results[0].device_ms = std::max(results[0].device_ms, res[static_cast<size_t>(r)].device_ms);
    // no pose is ever broadcast: every device applied the same step to the same sums
    if (res[static_cast<size_t>(r)].iterations != res[0].iterations ||
        memcmp(pose[static_cast<size_t>(r)].data(), pose[0].data(), 16 * sizeof(double)) != 0)
      return Fail(ctx, NLO_ECOMM, "the devices of a sharded solve ended in different states");
  }
  return NLO_OK;
}

}  // namespace multi
}  // namespace nlo
