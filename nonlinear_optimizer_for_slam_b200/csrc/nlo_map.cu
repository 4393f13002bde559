// nlo_map.cu -- host side of the widened rows of the scope table: the device
// NDT map (dense voxel grid or voxel hash), the scan, the matcher and the outer registration loop.
// The reference keeps these in its test mains (mahalanobis_distance_minimizer/tests/
// simple_optimization_test.cc:236-342,473-505); they are inside every timing it publishes.
//
// On a multi-device context these calls run on the context's first device (a registration frame is
// tens of thousands of points: one B200 is already latency-bound on it).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>

#include "nlo_host.h"

using namespace nlo;

namespace {

// A multi-device context runs these calls on its first device, with that device's in-process
// communicator suspended (the other devices do not take part) and the error message copied back.
struct FirstDevice {
  nlo_context* outer;
  nlo_context* sub;
  bool was_suspended;
  explicit FirstDevice(nlo_context* o) : outer(o), sub(o->subs[0]), was_suspended(sub->comm_suspended) {
    if (!was_suspended) {
      sub->comm_suspended = true;
      sub->generation++;
    }
  }
  ~FirstDevice() {
    if (!was_suspended) {
      sub->comm_suspended = false;
      sub->generation++;
    }
    outer->error = sub->error;
  }
};
#define NLO_ON_FIRST_DEVICE(ctx, call)            \
  if ((ctx) != nullptr && (ctx)->IsMulti()) {     \
    FirstDevice first(ctx);                       \
    nlo_context* sub = first.sub;                 \
    return call;                                  \
  }

struct EventPair {
  cudaEvent_t begin = nullptr, end = nullptr;
  ~EventPair() {
    if (begin) cudaEventDestroy(begin);
    if (end) cudaEventDestroy(end);
  }
};

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
  m->owner = ctx;
  for (int k = 0; k < 3; ++k) { m->origin[k] = origin[k]; m->dims[k] = dims[k]; }
  m->voxel = voxel;
  m->cells = cells;
  m->hash_mask = hash_slots ? hash_slots - 1 : 0;
  if (DevMalloc(&m->d_mean, cells * 3 * sizeof(double)) != cudaSuccess ||
      DevMalloc(&m->d_sqrt_info, cells * 9 * sizeof(double)) != cudaSuccess ||
      DevMalloc(&m->d_valid, cells) != cudaSuccess ||
      (hash_slots && DevMalloc(&m->d_keys, cells * sizeof(unsigned long long)) != cudaSuccess)) {
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

extern "C" {

int nlo_ndt_map_create(nlo_context* ctx, const double grid_origin[3], const int32_t grid_dims[3], double voxel_size,
                       const double* cell_mean, const double* cell_sqrt_info, const uint8_t* cell_valid,
                       nlo_ndt_map** map) {
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_map_create(sub, grid_origin, grid_dims, voxel_size, cell_mean, cell_sqrt_info, cell_valid, map));
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

}  // extern "C"

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
    if (DevMalloc(&d_scratch, (scratch_slots + 1) * sizeof(unsigned long long)) != cudaSuccess) {
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
    DevFree(d_scratch);
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
  auto cleanup = [&]() { DevFree(d_count); DevFree(d_sums); };
  cudaError_t e = DevMalloc(&d_count, m->cells * sizeof(int));
  if (e == cudaSuccess) e = DevMalloc(&d_sums, m->cells * 9 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, m->cells * sizeof(int), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_sums, 0, m->cells * 9 * sizeof(double), ctx->stream);
  if (e == cudaSuccess && hashed) e = cudaMemsetAsync(m->d_keys, 0xff, m->cells * sizeof(unsigned long long), ctx->stream);
  MapAccumParams ap;
  memset(&ap, 0, sizeof(ap));
  ap.xyz = d_xyz; ap.n = n; ap.inv_voxel = inv; ap.voxel = voxel_size;
  for (int k = 0; k < 3; ++k) { ap.kmin[k] = kmin[k]; ap.dims[k] = dims[k]; }
  ap.count = d_count; ap.sums = d_sums;
  {
    int bits = 0;  // of the point count: |d| <= 1/2 summed n times must stay below 2^62
    while ((static_cast<int64_t>(1) << bits) <= n) ++bits;
    ap.fixed_shift = std::min(40, 61 - bits);
  }
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

extern "C" {

int nlo_ndt_map_build(nlo_context* ctx, int64_t n, const double* points_xyz, double voxel_size, int v_not_transposed,
                      nlo_ndt_map** map) {
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_map_build(sub, n, points_xyz, voxel_size, v_not_transposed, map));
  return BuildMap(ctx, n, points_xyz, voxel_size, v_not_transposed, 2, map);
}

int nlo_ndt_map_build_hashed(nlo_context* ctx, int64_t n, const double* points_xyz, double voxel_size,
                             int v_not_transposed, nlo_ndt_map** map) {
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_map_build_hashed(sub, n, points_xyz, voxel_size, v_not_transposed, map));
  return BuildMap(ctx, n, points_xyz, voxel_size, v_not_transposed, 1, map);
}

int nlo_ndt_map_layout(nlo_context* ctx, const nlo_ndt_map* map, int32_t* hashed, int64_t* cells) {
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_map_layout(sub, map, hashed, cells));
  if (ctx == nullptr || map == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  if (hashed) *hashed = map->d_keys != nullptr ? 1 : 0;
  if (cells) *cells = map->cells;
  return NLO_OK;
}

int nlo_ndt_map_download_keys(nlo_context* ctx, const nlo_ndt_map* map, uint64_t* slot_keys) {
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_map_download_keys(sub, map, slot_keys));
  if (ctx == nullptr || map == nullptr || slot_keys == nullptr) return Fail(ctx, NLO_EINVAL, "null argument");
  if (map->d_keys == nullptr) return Fail(ctx, NLO_EINVAL, "not a hashed map");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  NLO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  NLO_CUDA(ctx, cudaMemcpy(slot_keys, map->d_keys, map->cells * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return NLO_OK;
}

int nlo_ndt_map_info(nlo_context* ctx, const nlo_ndt_map* map, double grid_origin[3], int32_t grid_dims[3],
                     double* voxel_size, int64_t* valid_cells) {
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_map_info(sub, map, grid_origin, grid_dims, voxel_size, valid_cells));
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
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_map_download(sub, map, cell_mean, cell_sqrt_info, cell_valid));
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
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_map_destroy(sub, map));
  if (map == nullptr) return NLO_OK;
  if (ctx != nullptr) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  DevFree(map->d_mean);
  DevFree(map->d_sqrt_info);
  DevFree(map->d_valid);
  DevFree(map->d_keys);
  delete map;
  return NLO_OK;
}

int nlo_scan_create(nlo_context* ctx, int64_t n, const double* points_xyz, nlo_scan** scan) {
  NLO_ON_FIRST_DEVICE(ctx, nlo_scan_create(sub, n, points_xyz, scan));
  if (ctx == nullptr || scan == nullptr || n < 0 || (n > 0 && points_xyz == nullptr)) return Fail(ctx, NLO_EINVAL, "bad argument");
  NLO_CUDA(ctx, cudaSetDevice(ctx->device));
  nlo_scan* sc = new nlo_scan();
  sc->n = n;
  const int64_t cap = std::max<int64_t>(n, 1);
  // one allocation: three planes + the matched-correspondence counter behind them
  if (DevMalloc(&sc->block, (cap * 3 + 8) * sizeof(double)) != cudaSuccess) {
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
  NLO_ON_FIRST_DEVICE(ctx, nlo_scan_destroy(sub, scan));
  if (scan == nullptr) return NLO_OK;
  if (ctx != nullptr) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  DevFree(scan->block);
  delete scan;
  return NLO_OK;
}

int nlo_ndt_match(nlo_context* ctx, const nlo_scan* scan, const nlo_ndt_map* map, const double pose[16], double radius,
                  int32_t max_neighbors, nlo_problem* pr, int64_t* matched) {
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_match(sub, scan, map, pose, radius, max_neighbors, pr, matched));
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
  NLO_ON_FIRST_DEVICE(ctx, nlo_ndt_register(sub, scan_in, map, options, radius, max_neighbors, max_outer, three_dof, pose, result));
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
  EventPair ev;
  NLO_CUDA(ctx, cudaEventCreate(&ev.begin));
  NLO_CUDA(ctx, cudaEventCreate(&ev.end));
  NLO_CUDA(ctx, cudaEventRecord(ev.begin, ctx->stream));
  int rc = NLO_OK;
  for (int outer = 0; outer < max_outer; ++outer) {
    double last[16];
    memcpy(last, pose, sizeof(last));
    rc = MatchInto(ctx, scan, map, pose, radius, max_neighbors, pr, scan->d_matched);
    if (rc != NLO_OK) break;
    if (three_dof) {
      // The planar minimizer processes floor(M / 4) * 4 of the M correspondences MatchPointCloud
      // returned (..._analytic_3dof.cc:33-36), i.e. it drops the last M mod 4 REAL hits of the
      // point-major list.  Here the list is a fixed 2 x n slot layout with zero-information fillers,
      // so those hits are found on the device and zeroed (an exact zero contribution), and the solve
      // then runs over every slot.
      cudaError_t ce = LaunchTruncateNdt3Hits(pr->planes, scan->n, max_neighbors, scan->d_matched, ctx->stream);
      if (ce != cudaSuccess) {
        rc = Fail(ctx, NLO_ECUDA, std::string("truncate hits: ") + cudaGetErrorString(ce));
        break;
      }
      pr->ndt3_end_override = pr->n;
    }
    nlo_solve_result sr{};
    rc = SolveImpl(ctx, pr, three_dof ? kNdt3 : kNdt6, options, pose, &sr, nullptr, false);
    pr->ndt3_end_override = -1;
    if (rc != NLO_OK) break;
    result->outer_iterations = outer + 1;
    result->inner_iterations += sr.iterations;
    if (sr.iterations > 0 || outer == 0) result->final_cost = sr.final_cost;
    // :495-499  pose_diff = current^-1 * last; stop on |dt| < 1e-5 and |dq.vec| < 1e-5 (the norms of
    // last^-1 * current, computed here, are the same)
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
  cudaEventRecord(ev.end, ctx->stream);
  unsigned long long m = 0;
  cudaMemcpyAsync(ctx->host_small, scan->d_matched, sizeof(m), cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  memcpy(&m, ctx->host_small, sizeof(m));
  result->matched = static_cast<int64_t>(m);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev.begin, ev.end);
  result->device_ms = ms;
  result->status = rc;
  return rc;
}

}  // extern "C"
