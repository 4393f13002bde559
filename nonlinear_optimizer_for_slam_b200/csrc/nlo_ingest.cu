// nlo_ingest.cu -- host -> device ingest of correspondences (and the synthetic generator / the
// download used by tests).
//
// Every upload runs through ONE chunked pipeline (PipelinedIngest): the source is cut into chunks
// of a few MB; host threads fill a ring of pinned chunks -- gathering the 15 hot doubles out of the
// reference's 304-byte `Correspondence` records (types.h:11-26), or copying slices of pageable
// arrays -- while the previous chunk travels over PCIe (cudaMemcpyAsync) and is repacked into the
// tile-interleaved planes by a gather kernel that also forms S^T S.  Device staging is O(chunk), not
// O(problem); the PCIe link carries 120 bytes per NDT correspondence instead of the 304-byte record;
// the caller's memory may be pageable.  Arrays that already live in pinned memory skip the host
// copy: their slices go straight to the device stage.
//
// This replaces the per-Solve AoS -> SoA conversion of the reference's SIMD minimizers
// (mahalanobis_distance_minimizer_analytic_simd.cc:19-28, ..._simd_various.cc:1252-1266); like
// there, the host only MOVES data -- all arithmetic (S^T S included) stays on the device.
#include <emmintrin.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "nlo_host.h"

namespace nlo {

namespace {

using Clock = std::chrono::steady_clock;
double MsSince(Clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

// One contiguous run of source records and where it lands in the planes.
struct Segment {
  int64_t src_begin, dst_begin, count;
};

}  // namespace

// ---- NUMA placement (best effort): pinned chunks and gather threads near the GPU's PCIe root ----
int DeviceNumaNode(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return -1;
  for (char* c = bus; *c; ++c) *c = static_cast<char>(tolower(*c));
  const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
  FILE* f = fopen(path.c_str(), "r");
  if (f == nullptr) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}

namespace {

// CPUs of a NUMA node that this process may run on (empty if unknown).
std::vector<int> NodeCpus(int node) {
  std::vector<int> cpus;
  if (node < 0) return cpus;
  const std::string path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
  FILE* f = fopen(path.c_str(), "r");
  if (f == nullptr) return cpus;
  char buf[4096] = {0};
  const size_t got = fread(buf, 1, sizeof(buf) - 1, f);
  fclose(f);
  buf[got] = 0;
  cpu_set_t allowed;
  CPU_ZERO(&allowed);
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return cpus;
  const char* p = buf;
  while (*p) {
    char* end = nullptr;
    long a = strtol(p, &end, 10);
    if (end == p) break;
    long b = a;
    p = end;
    if (*p == '-') {
      b = strtol(p + 1, &end, 10);
      p = end;
    }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
      if (CPU_ISSET(c, &allowed)) cpus.push_back(static_cast<int>(c));
    while (*p == ',' || *p == '\n' || *p == ' ') ++p;
  }
  return cpus;
}

void PreferNode(int node) {
#ifdef SYS_set_mempolicy
  if (node < 0 || node >= 64) return;
  unsigned long mask = 1ul << node;
  syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, 65ul);
#else
  (void)node;
#endif
}
void DefaultMemPolicy() {
#ifdef SYS_set_mempolicy
  syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
#endif
}

int EnsureRing(nlo_context* ctx, size_t chunk_bytes) {
  IngestRing& r = ctx->ring;
  if (chunk_bytes <= r.chunk_bytes) return NLO_OK;
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  FreeIngestRing(ctx);
  if (ctx->numa_node == -2) ctx->numa_node = DeviceNumaNode(ctx->device);
  PreferNode(ctx->numa_node);
  cudaError_t e = cudaSuccess;
  for (int k = 0; k < IngestRing::kSlots && e == cudaSuccess; ++k) {
    e = cudaHostAlloc(reinterpret_cast<void**>(&r.host[k]), chunk_bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) memset(r.host[k], 0, chunk_bytes);  // first touch under the preferred-node policy
    if (e == cudaSuccess) e = DevMalloc(reinterpret_cast<void**>(&r.device[k]), chunk_bytes);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.done[k], cudaEventDisableTiming);
  }
  DefaultMemPolicy();
  if (e != cudaSuccess) {
    FreeIngestRing(ctx);
    cudaGetLastError();
    return Fail(ctx, e == cudaErrorMemoryAllocation ? NLO_ENOMEM : NLO_ECUDA,
                std::string("ingest ring allocation: ") + cudaGetErrorString(e));
  }
  r.chunk_bytes = chunk_bytes;
  return NLO_OK;
}

bool IsPinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Persistent gather threads of a context, bound (once) to the CPUs of the device's NUMA node; returns
// how many of the `want` helpers exist.
int EnsureIngestPool(nlo_context* ctx, int want) {
  if (static_cast<int>(ctx->ingest_pool.size()) >= want) return want;
  if (ctx->numa_node == -2) ctx->numa_node = DeviceNumaNode(ctx->device);
  const std::vector<int> cpus = NodeCpus(ctx->numa_node);
  while (static_cast<int>(ctx->ingest_pool.size()) < want) {
    Worker* w = nullptr;
    try {
      w = new Worker();
    } catch (...) {  // thread limit of the process / cgroup: gather with the helpers there are
      break;
    }
    if (!cpus.empty()) {
      w->Post([cpus]() {
        cpu_set_t set;
        CPU_ZERO(&set);
        for (int c : cpus) CPU_SET(c, &set);
        sched_setaffinity(0, sizeof(set), &set);
      });
      w->Wait();
    }
    ctx->ingest_pool.push_back(w);
  }
  return std::min(want, static_cast<int>(ctx->ingest_pool.size()));
}

int ThreadsFor(const nlo_context* ctx, int64_t n) {
  int hw = static_cast<int>(std::thread::hardware_concurrency());
  if (hw < 1) hw = 1;
  int cap = ctx->ingest_threads > 0 ? ctx->ingest_threads : std::min(hw, 16);
  const int64_t by_size = n / 8192;  // a thread is worth waking for a few thousand records
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(cap, by_size)));
}

// The pipeline.  fill(host_slot, src_first, count, part, parts) writes records [src_first, src_first +
// count) of the source into the pinned slot (called concurrently for part = 0..parts-1; each call
// handles its own share of the records); host_fill == false skips the ring's host side (the source
// is pinned: `copy` moves slices straight from the caller's memory).  copy(device_slot, host_slot,
// src_first, count) enqueues the H2D transfer(s) of one chunk, consume(device_slot, src_first, count)
// the device repack.  rec_bytes = bytes of one record in a slot.
template <typename Fill, typename Copy, typename Consume>
int PipelinedIngest(nlo_context* ctx, int64_t n, size_t rec_bytes, bool host_fill, Fill fill, Copy copy,
                    Consume consume) {
  const auto t_begin = Clock::now();
  ctx->last_ingest_ms = 0.0;
  ctx->last_ingest_gather_ms = 0.0;
  if (n <= 0) return NLO_OK;
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  // chunk: a quarter of the upload, at least 8192 records, at most 16 MB
  const int64_t max_records = std::max<int64_t>(kTile, static_cast<int64_t>((16u << 20) / rec_bytes) / kTile * kTile);
  int64_t chunk = ((n + 3) / 4 + kTile - 1) / kTile * kTile;
  chunk = std::min(std::max<int64_t>(chunk, 8192), max_records);
  const int64_t num_chunks = (n + chunk - 1) / chunk;
  int rc = EnsureRing(ctx, static_cast<size_t>(chunk) * rec_bytes);
  if (rc != NLO_OK) return rc;
  IngestRing& ring = ctx->ring;
  for (int k = 0; k < IngestRing::kSlots; ++k) ring.used[k] = false;

  const int T = host_fill ? ThreadsFor(ctx, n) : 1;
  // Host gather: every chunk is cut into blocks that the threads CLAIM (an atomic cursor per chunk), so
  // a helper that wakes late simply takes fewer blocks instead of holding up the chunk.  The helpers
  // are persistent workers of the context (EnsureIngestPool): an upload of the reference's size
  // (100 k records, 0.3 ms of PCIe time) cannot afford to create a dozen threads.
  // next[c] = next unclaimed block of chunk c, filled[c] = blocks written; free_upto = chunks whose
  // slot may be overwritten
  int per_thread = 4;
  if (const char* v = getenv("NLO_INGEST_BLOCKS")) per_thread = std::max(1, std::min(16, atoi(v)));
  const int blocks = (T > 1) ? static_cast<int>(std::max<int64_t>(T, std::min<int64_t>(static_cast<int64_t>(per_thread) * T, chunk / 1024))) : 1;
  std::vector<std::atomic<int>> next(static_cast<size_t>(num_chunks)), filled(static_cast<size_t>(num_chunks));
  for (auto& f : next) f.store(0, std::memory_order_relaxed);
  for (auto& f : filled) f.store(0, std::memory_order_relaxed);
  std::atomic<int64_t> free_upto{0};
  std::atomic<bool> abort{false};
  auto fill_blocks = [&](int64_t c) {
    const int64_t first = c * chunk;
    const int64_t count = std::min(chunk, n - first);
    unsigned char* host = ring.host[c % IngestRing::kSlots];
    for (;;) {
      const int b = next[static_cast<size_t>(c)].fetch_add(1, std::memory_order_relaxed);
      if (b >= blocks) return;
      fill(host, first, count, b, blocks);
      filled[static_cast<size_t>(c)].fetch_add(1, std::memory_order_release);
    }
  };
  int helpers = 0;
  if (T > 1) {
    helpers = EnsureIngestPool(ctx, T - 1);
    for (int t = 0; t < helpers; ++t) {
      ctx->ingest_pool[static_cast<size_t>(t)]->Post([&]() {
        for (int64_t c = 0; c < num_chunks; ++c) {
          while (free_upto.load(std::memory_order_acquire) <= c) {
            if (abort.load(std::memory_order_relaxed)) return;
            std::this_thread::yield();
          }
          fill_blocks(c);
        }
      });
    }
  }
  auto join_all = [&]() {
    abort.store(true);
    for (int t = 0; t < helpers; ++t) ctx->ingest_pool[static_cast<size_t>(t)]->Wait();
    helpers = 0;
  };
  double gather_ms = 0.0;
  cudaError_t e = cudaSuccess;
  for (int64_t c = 0; c < num_chunks && e == cudaSuccess; ++c) {
    const int slot = static_cast<int>(c % IngestRing::kSlots);
    const int64_t first = c * chunk;
    const int64_t count = std::min(chunk, n - first);
    if (host_fill) {
      if (ring.used[slot]) e = cudaEventSynchronize(ring.done[slot]);  // the device is done with this slot
      if (e != cudaSuccess) break;
      const auto t0 = Clock::now();
      free_upto.store(c + 1, std::memory_order_release);
      fill_blocks(c);
      while (filled[static_cast<size_t>(c)].load(std::memory_order_acquire) < blocks) std::this_thread::yield();
      gather_ms += MsSince(t0);
    }
    e = copy(ring.device[slot], ring.host[slot], first, count);
    if (e == cudaSuccess) e = consume(ring.device[slot], first, count);
    if (e == cudaSuccess && host_fill) {
      e = cudaEventRecord(ring.done[slot], ctx->stream);
      ring.used[slot] = true;
    }
  }
  join_all();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return Fail(ctx, NLO_ECUDA, std::string("ingest: ") + cudaGetErrorString(e));
  ctx->last_ingest_ms = MsSince(t_begin);
  ctx->last_ingest_gather_ms = gather_ms;
  return NLO_OK;
}

// share [lo, hi) of `count` records for part `part` of `parts`
inline void Share(int64_t count, int part, int parts, int64_t* lo, int64_t* hi) {
  *lo = count * part / parts;
  *hi = count * (part + 1) / parts;
}

inline void StreamStore(double* dst, double v) {
  long long bits;
  memcpy(&bits, &v, 8);
  _mm_stream_si64(reinterpret_cast<long long*>(dst), bits);  // the pinned chunk is write-only for the CPU
}

// Destination runs of the source range [first, first + count) (`segments` is sorted by src_begin).
template <typename Fn>
cudaError_t ForEachRun(const std::vector<Segment>& segments, int64_t first, int64_t count, Fn fn) {
  const int64_t last = first + count;
  // first segment that ends after `first`
  size_t lo = 0, hi = segments.size();
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (segments[mid].src_begin + segments[mid].count <= first) lo = mid + 1; else hi = mid;
  }
  for (size_t s = lo; s < segments.size() && segments[s].src_begin < last; ++s) {
    const int64_t a = std::max(first, segments[s].src_begin);
    const int64_t b = std::min(last, segments[s].src_begin + segments[s].count);
    if (b <= a) continue;
    const cudaError_t e = fn(a - first, segments[s].dst_begin + (a - segments[s].src_begin), b - a);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

std::vector<Segment> SegmentsOf(const nlo_problem* pr, int64_t n) {
  std::vector<Segment> segs;
  if (!pr->batched) {
    segs.push_back(Segment{0, 0, n});
    return segs;
  }
  int64_t cursor = 0;
  for (int k = 0; k < pr->num_problems; ++k) {
    segs.push_back(Segment{cursor, pr->h_ranges[k].begin, pr->counts[k]});
    cursor += pr->counts[k];
  }
  return segs;
}

int CheckCount(nlo_context* ctx, const nlo_problem* pr, int64_t n) {
  int64_t total = 0;
  for (int64_t c : pr->counts) total += c;
  if (pr->batched ? (n != total) : (n > pr->counts[0])) return Fail(ctx, NLO_EINVAL, "n does not fit the problem");
  return NLO_OK;
}

}  // namespace

void FreeIngestPool(nlo_context* ctx) {
  for (Worker* w : ctx->ingest_pool) delete w;
  ctx->ingest_pool.clear();
}

void FreeIngestRing(nlo_context* ctx) {
  IngestRing& r = ctx->ring;
  for (int k = 0; k < IngestRing::kSlots; ++k) {
    if (r.host[k]) cudaFreeHost(r.host[k]);
    if (r.device[k]) DevFree(r.device[k]);
    if (r.done[k]) cudaEventDestroy(r.done[k]);
    r.host[k] = nullptr;
    r.device[k] = nullptr;
    r.done[k] = nullptr;
    r.used[k] = false;
  }
  r.chunk_bytes = 0;
}

// SoA arrays (point[3n], mean[3n], sqrt_info[9n]) of element type T (double, or float for the fp32
// storage mode).  Slot layout of a chunk of c records: [point 3c][mean 3c][sqrt_info 9c].
template <typename T>
static int UploadNdtSoa(nlo_context* ctx, nlo_problem* pr, int64_t n, const T* point, const T* mean,
                        const T* sqrt_info) {
  int rc = CheckCount(ctx, pr, n);
  if (rc != NLO_OK) return rc;
  const std::vector<Segment> segs = SegmentsOf(pr, n);
  const bool pinned = n > 0 && IsPinned(point) && IsPinned(mean) && IsPinned(sqrt_info);
  auto fill = [&](unsigned char* host, int64_t first, int64_t count, int part, int parts) {
    int64_t lo, hi;
    Share(count, part, parts, &lo, &hi);
    T* h = reinterpret_cast<T*>(host);
    memcpy(h + 3 * lo, point + 3 * (first + lo), static_cast<size_t>(hi - lo) * 3 * sizeof(T));
    memcpy(h + 3 * count + 3 * lo, mean + 3 * (first + lo), static_cast<size_t>(hi - lo) * 3 * sizeof(T));
    memcpy(h + 6 * count + 9 * lo, sqrt_info + 9 * (first + lo), static_cast<size_t>(hi - lo) * 9 * sizeof(T));
  };
  auto copy = [&](unsigned char* dev, unsigned char* host, int64_t first, int64_t count) {
    if (!pinned)
      return cudaMemcpyAsync(dev, host, static_cast<size_t>(count) * 15 * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    T* d = reinterpret_cast<T*>(dev);
    cudaError_t e = cudaMemcpyAsync(d, point + 3 * first, static_cast<size_t>(count) * 3 * sizeof(T),
                                    cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d + 3 * count, mean + 3 * first, static_cast<size_t>(count) * 3 * sizeof(T),
                          cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d + 6 * count, sqrt_info + 9 * first, static_cast<size_t>(count) * 9 * sizeof(T),
                          cudaMemcpyHostToDevice, ctx->stream);
    return e;
  };
  auto consume = [&](unsigned char* dev, int64_t first, int64_t count) {
    const T* d = reinterpret_cast<const T*>(dev);
    return ForEachRun(segs, first, count, [&](int64_t off, int64_t dst, int64_t run) {
      return LaunchPackNdt(d + 3 * off, d + 3 * count + 3 * off, d + 6 * count + 9 * off, run, pr->planes, dst,
                           pr->f32, ctx->stream);
    });
  };
  rc = PipelinedIngest(ctx, n, 15 * sizeof(T), !pinned, fill, copy, consume);
  if (rc != NLO_OK) return rc;
  pr->n = n;
  if (!pr->batched) pr->h_ranges[0] = Range{0, n};
  return NLO_OK;
}

int UploadNdt(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* point, const double* mean,
              const double* sqrt_info) {
  return UploadNdtSoa<double>(ctx, pr, n, point, mean, sqrt_info);
}

int UploadNdtF32(nlo_context* ctx, nlo_problem* pr, int64_t n, const float* point, const float* mean,
                 const float* sqrt_info) {
  return UploadNdtSoa<float>(ctx, pr, n, point, mean, sqrt_info);
}

// The reference's AoS records in place.  Host threads gather point | mean | sqrt_information (15
// doubles) of every record into 120-byte slot records; the device kernel transposes a column-major
// S, forms S^T S and writes the planes.
int UploadNdtAos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                 size_t off_point, size_t off_mean, size_t off_sqrt, int col_major) {
  const unsigned char* base = static_cast<const unsigned char*>(records);
  auto fill = [&](unsigned char* host, int64_t first, int64_t count, int part, int parts) {
    int64_t lo, hi;
    Share(count, part, parts, &lo, &hi);
    double* out = reinterpret_cast<double*>(host) + 15 * lo;
    const unsigned char* rec = base + static_cast<size_t>(first + lo) * stride;
    for (int64_t i = lo; i < hi; ++i, rec += stride, out += 15) {
      const double* p = reinterpret_cast<const double*>(rec + off_point);
      const double* m = reinterpret_cast<const double*>(rec + off_mean);
      const double* s = reinterpret_cast<const double*>(rec + off_sqrt);
      StreamStore(out + 0, p[0]); StreamStore(out + 1, p[1]); StreamStore(out + 2, p[2]);
      StreamStore(out + 3, m[0]); StreamStore(out + 4, m[1]); StreamStore(out + 5, m[2]);
      for (int k = 0; k < 9; ++k) StreamStore(out + 6 + k, s[k]);
    }
    _mm_sfence();
  };
  auto copy = [&](unsigned char* dev, unsigned char* host, int64_t, int64_t count) {
    return cudaMemcpyAsync(dev, host, static_cast<size_t>(count) * 120, cudaMemcpyHostToDevice, ctx->stream);
  };
  auto consume = [&](unsigned char* dev, int64_t first, int64_t count) {
    return LaunchPackNdtAos(dev, count, 120, 0, 24, 48, col_major, pr->planes, first, ctx->stream);
  };
  const int rc = PipelinedIngest(ctx, n, 120, true, fill, copy, consume);
  if (rc != NLO_OK) return rc;
  pr->n = n;
  pr->h_ranges[0] = Range{0, n};
  return NLO_OK;
}

int UploadReproj(nlo_context* ctx, nlo_problem* pr, int64_t n, const double* local_point, const double* pixel,
                 const double intrinsics[6]) {
  int rc = CheckCount(ctx, pr, n);
  if (rc != NLO_OK) return rc;
  const std::vector<Segment> segs = SegmentsOf(pr, n);
  const bool pinned = n > 0 && IsPinned(local_point) && IsPinned(pixel);
  auto fill = [&](unsigned char* host, int64_t first, int64_t count, int part, int parts) {
    int64_t lo, hi;
    Share(count, part, parts, &lo, &hi);
    double* h = reinterpret_cast<double*>(host);
    memcpy(h + 3 * lo, local_point + 3 * (first + lo), static_cast<size_t>(hi - lo) * 24);
    memcpy(h + 3 * count + 2 * lo, pixel + 2 * (first + lo), static_cast<size_t>(hi - lo) * 16);
  };
  auto copy = [&](unsigned char* dev, unsigned char* host, int64_t first, int64_t count) {
    if (!pinned) return cudaMemcpyAsync(dev, host, static_cast<size_t>(count) * 40, cudaMemcpyHostToDevice, ctx->stream);
    double* d = reinterpret_cast<double*>(dev);
    cudaError_t e = cudaMemcpyAsync(d, local_point + 3 * first, static_cast<size_t>(count) * 24, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d + 3 * count, pixel + 2 * first, static_cast<size_t>(count) * 16, cudaMemcpyHostToDevice, ctx->stream);
    return e;
  };
  auto consume = [&](unsigned char* dev, int64_t first, int64_t count) {
    const double* d = reinterpret_cast<const double*>(dev);
    return ForEachRun(segs, first, count, [&](int64_t off, int64_t dst, int64_t run) {
      return LaunchPackReproj(d + 3 * off, d + 3 * count + 2 * off, run, pr->planes, dst, ctx->stream);
    });
  };
  rc = PipelinedIngest(ctx, n, 40, !pinned, fill, copy, consume);
  if (rc != NLO_OK) return rc;
  for (int k = 0; k < 6; ++k) pr->intrinsics[k] = intrinsics[k];
  pr->n = n;
  if (!pr->batched) pr->h_ranges[0] = Range{0, n};
  DropGraphs(pr);  // intrinsics are baked into captured launches
  return NLO_OK;
}

// reprojection_error_minimizer/types.h:14-28 records in place: local_point (3 doubles) and pixel (2).
int UploadReprojAos(nlo_context* ctx, nlo_problem* pr, int64_t n, const void* records, size_t stride,
                    size_t off_point, size_t off_pixel, const double intrinsics[6]) {
  const unsigned char* base = static_cast<const unsigned char*>(records);
  auto fill = [&](unsigned char* host, int64_t first, int64_t count, int part, int parts) {
    int64_t lo, hi;
    Share(count, part, parts, &lo, &hi);
    double* h = reinterpret_cast<double*>(host);
    const unsigned char* rec = base + static_cast<size_t>(first + lo) * stride;
    for (int64_t i = lo; i < hi; ++i, rec += stride) {
      const double* X = reinterpret_cast<const double*>(rec + off_point);
      const double* px = reinterpret_cast<const double*>(rec + off_pixel);
      h[3 * i] = X[0]; h[3 * i + 1] = X[1]; h[3 * i + 2] = X[2];
      h[3 * count + 2 * i] = px[0]; h[3 * count + 2 * i + 1] = px[1];
    }
  };
  auto copy = [&](unsigned char* dev, unsigned char* host, int64_t, int64_t count) {
    return cudaMemcpyAsync(dev, host, static_cast<size_t>(count) * 40, cudaMemcpyHostToDevice, ctx->stream);
  };
  auto consume = [&](unsigned char* dev, int64_t first, int64_t count) {
    const double* d = reinterpret_cast<const double*>(dev);
    return LaunchPackReproj(d, d + 3 * count, count, pr->planes, first, ctx->stream);
  };
  const int rc = PipelinedIngest(ctx, n, 40, true, fill, copy, consume);
  if (rc != NLO_OK) return rc;
  for (int k = 0; k < 6; ++k) pr->intrinsics[k] = intrinsics[k];
  pr->n = n;
  pr->h_ranges[0] = Range{0, n};
  DropGraphs(pr);
  return NLO_OK;
}

int GenerateNdt(nlo_context* ctx, nlo_problem* pr, uint64_t seed, int64_t global_index_offset,
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

int DownloadNdt(nlo_context* ctx, const nlo_problem* pr, int32_t problem_index, int64_t begin, int64_t end,
                double* point, double* mean, double* information) {
  if (problem_index < 0 || problem_index >= pr->num_problems) return Fail(ctx, NLO_EINVAL, "bad problem_index");
  const int64_t count = pr->batched ? pr->counts[problem_index] : pr->n;
  if (begin < 0 || end < begin || end > count) return Fail(ctx, NLO_EINVAL, "bad [begin, end)");
  if (!point || !mean || !information) return Fail(ctx, NLO_EINVAL, "null array");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = end - begin;
  const int64_t base = pr->batched ? pr->h_ranges[problem_index].begin : 0;
  int rc = EnsureStaging(ctx, static_cast<size_t>(n) * 12 * sizeof(double) + 256);
  if (rc != NLO_OK) return rc;
  double* s_point = static_cast<double*>(ctx->staging);
  double* s_mean = s_point + 3 * n;
  double* s_info = s_mean + 3 * n;
  NLO_CUDA(ctx, LaunchUnpackNdt(const_cast<double* const*>(pr->planes), base + begin, base + end, s_point, s_mean,
                                s_info, pr->f32, ctx->stream));
  if (n > 0) {
    NLO_CUDA(ctx, cudaMemcpyAsync(point, s_point, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NLO_CUDA(ctx, cudaMemcpyAsync(mean, s_mean, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    NLO_CUDA(ctx, cudaMemcpyAsync(information, s_info, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return NLO_OK;
}

}  // namespace nlo
