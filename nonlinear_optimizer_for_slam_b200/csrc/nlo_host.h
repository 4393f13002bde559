// nlo_host.h -- host-side structures shared by the translation units behind the C ABI
// (nlo_api.cu: contexts, problems, assemble / solve, communicators; nlo_ingest.cu: host -> device
// ingest; nlo_map.cu: NDT map, scan, matcher, outer registration loop; nlo_multi.cu: one context
// spanning several devices of the process).  Nothing here is visible through include/nlo_cuda.h.
#ifndef NLO_HOST_H_
#define NLO_HOST_H_

#include <array>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nlo_cuda.h"
#include "nlo_internal.h"

namespace nlo {

// ---- NCCL through dlopen (no link-time dependency; the symbols are only needed multi-process) ----
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

enum CommKind { kCommNone = 0, kCommNccl = 1, kCommPeer = 2 };

constexpr size_t kPeerBufBytes = 2 * kMaxRanks * kPeerWords * sizeof(unsigned long long);
constexpr int kInCtaTiles = 3;       // a registration this small runs its whole loop inside one CTA
constexpr int kSmallDoubles = 8192;  // pinned + device scratch for poses / sums / results

// Host -> device ingest pipeline of one context (nlo_ingest.cu): a ring of pinned chunks that host
// threads fill (gather of the hot scalars out of the caller's records, or a plain copy out of
// pageable arrays) while the previous chunk travels over PCIe and is repacked on the device.
struct IngestRing {
  static constexpr int kSlots = 3;
  size_t chunk_bytes = 0;  // capacity of every slot (host and device)
  unsigned char* host[kSlots] = {nullptr, nullptr, nullptr};   // pinned
  unsigned char* device[kSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t done[kSlots] = {nullptr, nullptr, nullptr};      // device finished reading slot k
  bool used[kSlots] = {false, false, false};
};

// Rendezvous of the host threads that drive the devices of a multi-device context.  Needed when one
// device carries several shards (a test configuration: the device list names a device twice): the
// shards' kernels spin on each other, so no host thread may queue work BEHIND its spinning kernel
// (the read-back of the result, an event) before the kernels of all shards are in their queues --
// work of different streams can share a hardware queue, and an entry that waits for a spinning
// kernel then holds back the very kernel it spins for.
class HostBarrier {
 public:
  explicit HostBarrier(int participants) : n_(participants) {}
  void ArriveAndWait() {
    std::unique_lock<std::mutex> lock(mu_);
    const long generation = generation_;
    if (++count_ == n_) {
      count_ = 0;
      ++generation_;
      cv_.notify_all();
    } else {
      cv_.wait(lock, [&] { return generation_ != generation; });
    }
  }
  void ArriveNoWait() {
    std::lock_guard<std::mutex> lock(mu_);
    if (++count_ == n_) {
      count_ = 0;
      ++generation_;
      cv_.notify_all();
    }
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int n_;
  int count_ = 0;
  long generation_ = 0;
};

// The `phases` rendezvous points of one call; whatever is left when the call returns early is
// arrived at without waiting, so the other threads never hang on a failed shard.
class RendezvousGuard {
 public:
  RendezvousGuard(HostBarrier* barrier, int phases) : barrier_(barrier), left_(barrier ? phases : 0) {}
  ~RendezvousGuard() {
    while (left_-- > 0) barrier_->ArriveNoWait();
  }
  bool active() const { return left_ > 0; }
  void Sync() {
    if (left_ > 0) {
      --left_;
      barrier_->ArriveAndWait();
    }
  }

 private:
  HostBarrier* barrier_;
  int left_;
};

// A persistent worker thread (one per device of a multi-device context).
class Worker {
 public:
  Worker();
  ~Worker();
  void Post(std::function<void()> job);
  void Wait();

 private:
  void Loop();
  std::mutex mu_;
  std::condition_variable cv_;
  std::function<void()> job_;
  bool has_job_ = false, busy_ = false, stop_ = false;
  std::thread thread_;
};

}  // namespace nlo

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
  int stage_depth = 0;            // NLO_STAGE_DEPTH: ring depth override (0 = per-shape default)
  double l2_keep_mb = 0.0;        // NLO_L2_KEEP_MB: bytes of a re-read scan pinned in L2 (0 = off)
  double l2_policy_min_mb = 0.0;  // only scans larger than this get an explicit policy
  int grid_small = 0;             // CTAs of the persistent path for L2-resident problems
  int cluster_small = 8;          // NLO_CLUSTER: CTAs per thread-block cluster, problems resident in smem / L2
  int direct_max_clusters = 48;   // NLO_DIRECT_MAX: up to this many cluster partials every CTA gathers them itself
  int warp_groups = 2;            // NLO_WARP_GROUPS: 2 = the persistent streaming loop runs one 512-thread CTA per SM
  bool use_resident = true;       // NLO_NO_RESIDENT=1: latency-bound registrations also run the streaming kernel
  int device_share = 1;           // sub-contexts of one multi-device context that sit on this device
  std::map<int, int> coresident;  // (kernel variant, cluster size) -> co-resident CTAs on this device
  // communicator
  int comm_kind = nlo::kCommNone;
  bool comm_suspended = false;  // nlo_comm_suspend: calls behave as if no communicator were attached
  bool comm_in_process = false; // peer buffers of the other devices are mapped directly (multi-device context)
  int rank = 0, nranks = 1;
  nlo::NcclApi nccl;
  nlo::NcclComm nccl_comm = nullptr;
  unsigned char* peer_buf = nullptr;  // local exchange buffer (exported over CUDA IPC / mapped by peers)
  void* peer_opened[nlo::kMaxRanks] = {nullptr};
  nlo::PeerComm peer{};
  unsigned long long* d_peer_seq = nullptr;
  int* d_peer_error = nullptr;
  nlo_problem* reg_workspace = nullptr;  // correspondences of nlo_ndt_register: kept across scans,
  int64_t reg_workspace_capacity = 0;    // grown on demand (no per-frame allocation)
  unsigned long long* d_debug_times = nullptr;  // NLO_DEBUG_TIMES=1|2: [kDebugCtas][64][8] stamps of the last loop
  bool debug_all_ctas = false;
  int generation = 0;  // bumped whenever cached graphs become stale (loss / comm change)
  // ingest pipeline
  nlo::IngestRing ring;
  int ingest_threads = 0;  // host threads one upload may use (NLO_INGEST_THREADS; 0 = decide per call)
  std::vector<nlo::Worker*> ingest_pool;  // persistent gather threads (created on the first threaded upload)
  int numa_node = -2;                     // NUMA node of the device's PCIe root (-1 unknown, -2 not looked up yet)
  double last_ingest_ms = 0.0, last_ingest_gather_ms = 0.0;  // wall time of the last upload / its host gather
  // multi-device context (nlo_context_create_multi): this object is then only a dispatcher
  std::vector<nlo_context*> subs;
  std::vector<nlo::Worker*> workers;

  nlo::HostBarrier* launch_barrier = nullptr;  // sub-context of a multi-device context that shares devices: see HostBarrier

  bool IsMulti() const { return !subs.empty(); }
  int EffectiveComm() const { return comm_suspended ? nlo::kCommNone : comm_kind; }
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
  double* planes[nlo::kNdtPlanes] = {nullptr};
  std::vector<nlo::Range> h_ranges;
  std::vector<int64_t> counts;
  nlo::Range* d_ranges = nullptr;  // [num_problems + 1]; the last slot is the scratch range for assemble
  nlo::State* d_states = nullptr;  // [num_problems + 1]
  double* d_partials = nullptr;
  unsigned int* d_tickets = nullptr;     // [num_problems + 1]
  unsigned long long* d_sync = nullptr;  // [num_problems + 1][kSyncStride] persistent path: canonical sums as LL words
  unsigned long long* d_ll_partials = nullptr;  // persistent path: cluster partials as LL words
  size_t ll_partials_bytes = 0;
  unsigned int epoch = 0;                // persistent launches made on this problem (mod 65535): upper half of the LL tags
  double* d_sums = nullptr;              // [(num_problems + 1) * 32]
  double* d_poses = nullptr;             // [(num_problems + 1) * 16]
  double* d_results = nullptr;           // [(num_problems + 1) * 4]
  double* d_trace = nullptr;
  size_t trace_doubles = 0;
  double intrinsics[6] = {0, 0, 0, 0, 0, 0};
  std::map<std::array<int64_t, 10>, cudaGraphExec_t> graphs;
  // 3-DoF solve of a shard of a larger scan: the reference drops the last n mod 4 correspondences of
  // the WHOLE list (..._analytic_3dof.cc:33-36), so the owner of the shards fixes the end of every
  // shard's range here (-1 = this problem is the whole list: floor(n / 4) * 4).
  int64_t ndt3_end_override = -1;
  // multi-device problem: one shard per device (point ranges for a single problem, problem-id
  // ranges for a batched one); this object is then only a dispatcher
  std::vector<nlo_problem*> shards;
  std::vector<int64_t> shard_begin;  // [shards + 1]: first point (single) / first problem id (batched) of shard r
};

struct nlo_ndt_map {
  nlo_context* owner = nullptr;  // the (sub-)context whose device holds the tables
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

namespace nlo {

int Fail(nlo_context* ctx, int code, const std::string& msg);

#define NLO_CUDA(ctx, expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      return ::nlo::Fail((ctx), (_e == cudaErrorMemoryAllocation) ? NLO_ENOMEM : NLO_ECUDA, \
                         std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    }                                                                                    \
  } while (0)

// Device allocations of the library (nlo_api.cu).  Plain cudaMalloc / cudaFree unless NLO_GUARD=1 is in
// the environment -- the memory-safety check that stands in for compute-sanitizer where that tool is
// not available: every allocation then sits between two 64 KB guard bands filled with 0xFF, and its
// payload is pre-filled with 0xFF too (as doubles NaN, as LL tags a value no iteration uses, as
// counters / indices absurd), so a write past either end of a buffer is found when the bands are
// compared (at cudaFree time and by nlo_debug_guard_report), and a read of memory nothing wrote, or
// just outside a buffer, poisons the result the parity tests look at.
cudaError_t DevMallocBytes(void** p, size_t bytes);
template <typename T>
cudaError_t DevMalloc(T** p, size_t bytes) {
  return DevMallocBytes(reinterpret_cast<void**>(p), bytes);
}
cudaError_t DevFree(void* p);

int EnsureStaging(nlo_context* ctx, size_t bytes);
void DropGraphs(nlo_problem* pr);
void PoseToRt(const double pose[16], double R[9], double t[3]);

// nlo_api.cu (single-device implementations the multi-device dispatcher calls per shard)
int CreateProblem(nlo_context* ctx, int family, int num_problems, const int64_t* counts, bool batched,
                  nlo_problem** out, bool f32 = false);
int AssembleImpl(nlo_context* ctx, nlo_problem* pr, int kind, int problem_index, const double pose[16],
                 int64_t begin, int64_t end, double* H, int nh, double* g, int ng, double* cost);
int EnsureTrace(nlo_context* ctx, nlo_problem* pr, size_t need_doubles);
int SolveImpl(nlo_context* ctx, nlo_problem* pr, int kind, const nlo_solve_options* options, double* poses,
              nlo_solve_result* results, double* trace, bool batched_call);

// nlo_ingest.cu
int DeviceNumaNode(int device);  // NUMA node of the device's PCIe root, -1 if unknown
int UploadNdt(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* point, const double* mean,
              const double* sqrt_info);
int UploadNdtF32(nlo_context* ctx, nlo_problem* pr, int64_t n, const float* point, const float* mean,
                 const float* sqrt_info);
int UploadNdtAos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                 size_t offset_point, size_t offset_mean, size_t offset_sqrt_info, int col_major);
int UploadReproj(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* local_point, const double* pixel,
                 const double intrinsics[6]);
int UploadReprojAos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                    size_t offset_local_point, size_t offset_pixel, const double intrinsics[6]);
int GenerateNdt(nlo_context* ctx, nlo_problem* pr, uint64_t seed, int64_t global_index_offset, double noise_sigma,
                const double* true_poses, const double init_pose[16], const double grid_origin[3],
                const int32_t grid_dims[3], double voxel_size, const double* cell_mean,
                const double* cell_sqrt_info, const uint8_t* cell_valid, int64_t n_single);
int DownloadNdt(nlo_context* ctx, const nlo_problem* pr, int32_t problem_index, int64_t begin, int64_t end,
                double* point, double* mean, double* information);
void FreeIngestRing(nlo_context* ctx);
void FreeIngestPool(nlo_context* ctx);

// nlo_multi.cu
namespace multi {
int CreateContext(const int* devices, int n, nlo_context** out);
void DestroyContext(nlo_context* ctx);
int SetLoss(nlo_context* ctx, int kind, const double params[2]);
int Synchronize(nlo_context* ctx);
int Create(nlo_context* ctx, int family, int num_problems, const int64_t* counts, bool batched, bool f32,
           nlo_problem** out);
int Destroy(nlo_context* ctx, nlo_problem* pr);
int UploadNdt(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* point, const double* mean,
              const double* sqrt_info);
int UploadNdtF32(nlo_context* ctx, nlo_problem* pr, int64_t n, const float* point, const float* mean,
                 const float* sqrt_info);
int UploadNdtAos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                 size_t offset_point, size_t offset_mean, size_t offset_sqrt_info, int col_major);
int UploadReproj(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* local_point, const double* pixel,
                 const double intrinsics[6]);
int UploadReprojAos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                    size_t offset_local_point, size_t offset_pixel, const double intrinsics[6]);
int Generate(nlo_context* ctx, nlo_problem* pr, uint64_t seed, int64_t global_index_offset, double noise_sigma,
             const double* true_poses, const double init_pose[16], const double grid_origin[3],
             const int32_t grid_dims[3], double voxel_size, const double* cell_mean, const double* cell_sqrt_info,
             const uint8_t* cell_valid, int64_t n_single);
int Download(nlo_context* ctx, const nlo_problem* pr, int32_t problem_index, int64_t begin, int64_t end,
             double* point, double* mean, double* information);
int Assemble(nlo_context* ctx, nlo_problem* pr, int kind, int problem_index, const double pose[16], int64_t begin,
             int64_t end, double* H, int nh, double* g, int ng, double* cost);
int Solve(nlo_context* ctx, nlo_problem* pr, int kind, const nlo_solve_options* options, double* poses,
          nlo_solve_result* results, double* trace, bool batched_call);
}  // namespace multi

}  // namespace nlo

#endif  // NLO_HOST_H_
