// nlo_internal.h -- structures shared by the kernels (nlo_kernels.cu) and the C ABI (nlo_api.cu).
#ifndef NLO_INTERNAL_H_
#define NLO_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>

namespace nlo {

enum Kind : int { kNdt6 = 0, kNdt3 = 1, kReproj = 2 };

constexpr int kTile = 256;            // correspondences per pipeline stage = consumer threads
constexpr int kConsumerWarps = kTile / 32;
constexpr int kThreads = kTile;       // thread 0 doubles as the TMA producer
constexpr int kNdtPlanes = 12;        // x y z | mx my mz | L00 L01 L02 L11 L12 L22 with L = S^T S, formed once
                                      // at ingest (both NDT minimizers need S only through S^T S); stored as
                                      // double (96 B per correspondence) or, opt-in, as float (48 B)
constexpr int kReprojPlanes = 5;      // X Y Z | u v
constexpr int kAcc6 = 28;             // 21 H + 6 g + cost
constexpr int kAcc3 = 10;             // 6 H + 3 g + cost
constexpr int kMaxRanks = 8;
constexpr int kPeerWords = 64;        // 8-byte words per (parity, source rank) slot of the peer exchange
constexpr int kSyncStride = 2 * kPeerWords;  // 8-byte words per registration of IterParams::ll_sums: [2 parities][kPeerWords]
constexpr int kMaxCluster = 8;        // CTAs per thread-block cluster of the persistent path (portable limit)
constexpr int kDebugIterations = 64;  // iterations stamped in IterParams::debug_times
constexpr int kDebugSlots = 16;       // stamps per iteration: 0-6 phases of the kernel, 7 SM id, 8-14 sub-phases
constexpr int kDebugCtas = 320;       // CTAs stamped when every CTA is (NLO_DEBUG_TIMES=2)
// Spin-wait limits (a wait that expires ends the solve with status 2 / NLO_ETIMEOUT or NLO_ECOMM, never a
// hang).  The leader CTA may legitimately sit in the peer exchange for as long as another rank needs to
// reach its own Solve (ingest skew, lazy module load, a time-sliced GPU), and the other CTAs of the grid
// wait for the leader meanwhile, so the intra-grid limit has to be the longer one.
constexpr unsigned long long kPeerTimeoutNs = 20000000000ULL;  // 20 s
constexpr unsigned long long kGridTimeoutNs = 60000000000ULL;  // 60 s

// Per-registration optimisation state, resident in HBM for the whole solve.
//   6-DoF / reprojection: t[3], q[4] (x,y,z,w), R = row-major rotation of q.
//   3-DoF: t[0..1], R[0..3] = row-major 2x2 (the reference keeps an Isometry2d, no quaternion).
struct State {
  double t[3];
  double q[4];
  double R[9];
  double lambda;
  double previous_cost;
  int iteration;
  int done;
  int status;
  int pad;
};

// HBM layout of the correspondences: tile-interleaved SoA ("AoSoA").  Tile T (256 consecutive
// correspondences) stores its NP planes back to back, [plane][256] doubles, so a whole tile is ONE
// contiguous run of NP * 2 KB that a single TMA bulk copy moves into a shared-memory stage
// unchanged.  planes[k] points at plane k of tile 0; element i of plane k lives at
// planes[k][TiledOffset(NP, i)].
__host__ __device__ inline int64_t TiledOffset(int nplanes, int64_t i) {
  return (i >> 8) * (static_cast<int64_t>(nplanes) << 8) + (i & 255);
}

struct Range {
  int64_t begin;  // absolute index into the planes
  int64_t end;
};

// One-shot all-reduce over peer-mapped buffers (NVLink / NVSwitch).
//   slots[r] points into rank r's exchange buffer: [2 parities][kMaxRanks sources][kPeerWords] of
//   8-byte words, each = (sequence number << 32) | 32 payload bits
struct PeerComm {
  int rank;
  int nranks;
  unsigned long long* slots[kMaxRanks];
  unsigned long long* seq;  // local, device: exchange sequence number (monotonic)
  int* error;               // local, device: set to 1 if a wait timed out
};

enum Mode : int {
  kModeSolve = 0,     // assemble + (last CTA) reduce + damped step + state update
  kModeAssemble = 1,  // assemble + reduce, canonical sums written to sums_out, no step
  kModeStepOnly = 2   // no assembly: read reduced canonical sums from sums_in and step (NCCL path)
};

struct IterParams {
  const double* planes[kNdtPlanes];
  const Range* ranges;  // [num_problems]
  State* states;        // [num_problems]
  double* partials;     // [2 parities][num_problems][grid.x][nacc]
  unsigned int* tickets;  // [num_problems] last-CTA election of the one-iteration-per-launch path
  // Persistent path ("LL" words: 32 payload bits + a 32-bit tag per 8-byte store, valid the moment the tag
  // matches -- no fence, no flag, no counter):
  unsigned long long* ll_partials;  // [2 parities][num_problems][clusters per registration][nacc][2 words]: the
                                    // raw sums of one thread-block cluster, written by its rank-0 CTA
  unsigned long long* ll_sums;      // [num_problems][2 parities][kPeerWords]: the canonical sums of the whole
                                    // registration, written by CTA 0 when the CTAs do not gather the partials
                                    // themselves (gather_direct == 0, single GPU)
  unsigned int tag_base;            // tag of iteration `it` = tag_base + it + 1 (solve epoch in the upper 16 bits)
  int gather_direct;                // every CTA gathers the cluster partials and steps on its own (few clusters);
                                    // 0: CTA 0 gathers, [pushes to the peers], every CTA gathers ll_sums / peer slots
  double* sums;           // [num_problems][32] canonical H|g|cost (out for assemble, in for step-only)
  double* trace;          // nullable: [num_problems][max_iterations][trace_width]
  double loss_p0, loss_p1;
  double intrinsics[6];
  double parameter_tolerance, gradient_tolerance;
  int max_iterations;
  int iterations_in_kernel;  // > 1 only when grid.x == 1 (whole loop inside one CTA)
  int mode;
  int use_peer;
  unsigned long long* debug_times;  // nullable: [CTA][iterations][8] globaltimer stamps (profiling aid; CTA 0 only
                                    // unless debug_all_ctas)
  int debug_all_ctas;
  long long l2_keep_tiles;  // > 0: tiles [0, l2_keep_tiles) of the range are loaded evict_last, the rest evict_first
  int stage_depth;  // > 0: use only this many of the allocated shared-memory stages
  int f32;  // NDT planes stored as float (fp32 storage, fp64 math); planes[0] then points at float data
  int persistent;  // cooperative launch: the whole loop in one grid, grid barrier per iteration
  PeerComm peer;
};

// Launchers (nlo_kernels.cu).  grid_x CTAs per registration, num_problems registrations.
// `cluster` CTAs per thread-block cluster (1 = none; grid_x must be a multiple of it).
// warp_groups: 1 = 256 threads, one tile per ring stage (2 CTAs/SM); 2 = 512 threads, two tiles per stage (1 CTA/SM).
cudaError_t LaunchIteration(int kind, int loss, const IterParams& p, int grid_x, int num_problems,
                            int cluster, cudaStream_t stream, int warp_groups = 1);
int MaxCoResidentCtas(int kind, int loss, bool f32, int cluster, bool resident);
// Resident kernel: the whole registration lives in the shared memory of the grid, `stages` tiles per CTA
// (<= ResidentMaxStages(kind)); always a persistent cooperative launch of the complete loop.
cudaError_t LaunchResident(int kind, int loss, const IterParams& p, int grid_x, int num_problems, int cluster,
                           int stages, cudaStream_t stream);
int ResidentMaxStages(int kind);
cudaError_t ConfigureKernels();
size_t IterationSmemBytes(int kind);

// Small utility kernels.
cudaError_t LaunchPeerRendezvous(const PeerComm& pc, cudaStream_t stream);
cudaError_t LaunchInitStates(State* states, const double* poses16, int num_problems, int kind,
                             cudaStream_t stream);
cudaError_t LaunchFinishStates(const State* states, double* poses16, double* results4,
                               int num_problems, int kind, cudaStream_t stream);
cudaError_t LaunchPackNdt(const double* point, const double* mean, const double* sqrt_info,
                          int64_t n, double* const planes[kNdtPlanes], int64_t dst_offset,
                          bool f32, cudaStream_t stream);
// float input (fp32 storage mode only: f32 must be true)
cudaError_t LaunchPackNdt(const float* point, const float* mean, const float* sqrt_info, int64_t n,
                          double* const planes[kNdtPlanes], int64_t dst_offset, bool f32, cudaStream_t stream);
cudaError_t LaunchPackNdtAos(const unsigned char* records, int64_t n, size_t stride,
                             size_t off_point, size_t off_mean, size_t off_sqrt, int col_major,
                             double* const planes[kNdtPlanes], int64_t dst_offset, cudaStream_t stream);
cudaError_t LaunchUnpackNdt(double* const planes[kNdtPlanes], int64_t begin, int64_t end,
                            double* point, double* mean, double* sqrt_info, bool f32,
                            cudaStream_t stream);
cudaError_t LaunchPackReproj(const double* local_point, const double* pixel, int64_t n,
                             double* const planes[kReprojPlanes], int64_t dst_offset, cudaStream_t stream);

struct GenerateParams {
  double* planes[kNdtPlanes];
  int64_t dst_offset;  // first correspondence index written (tile-aligned for batched problems)
  int f32;             // planes hold floats
  int64_t n;
  uint64_t seed;
  int64_t index_offset;
  double noise_sigma;
  double R_true[9], t_true[3];  // sensor -> world (row-major R)
  double R_init[9], t_init[3];
  double origin[3];
  int dims[3];
  double inv_voxel;
  int reach;  // neighbour cells scanned for the nearest-mean fallback: ceil(1 m / voxel)
  const double* cell_mean;
  const double* cell_sqrt_info;
  const unsigned char* cell_valid;
};
cudaError_t LaunchGenerateNdt(const GenerateParams& p, cudaStream_t stream);

// Device NDT matcher (MatchPointCloud of the reference's test mains): for every scan point warped
// by the pose, the <= max_neighbors nearest valid cell means within `radius`.
struct MatchParams {
  const double* scan[3];        // x y z planes of the scan (sensor frame)
  int64_t n;                    // scan points
  double* planes[kNdtPlanes];   // output correspondences: neighbour slot j of point i at j * n + i
  double R[9], t[3];            // pose (row-major R)
  double origin[3];
  int dims[3];
  double inv_voxel;
  int reach;
  double radius2;
  int max_neighbors;            // 1 or 2
  const double* cell_mean;
  const double* cell_sqrt_info;
  const unsigned char* cell_valid;
  unsigned long long* matched;  // device counter of real (non-empty) correspondences
  const unsigned long long* keys;  // hashed map: slot keys (nullptr = dense grid, cell = linear index)
  long long hash_mask;             // slots - 1 (slots is a power of two)
};
cudaError_t LaunchMatchNdt(const MatchParams& p, cudaStream_t stream);
// Zeroes the last (matched mod 4) real hits in point-major order (the planar minimizer's truncation).
cudaError_t LaunchTruncateNdt3Hits(double* const planes[kNdtPlanes], int64_t n, int max_neighbors,
                                   const unsigned long long* matched, cudaStream_t stream);
cudaError_t LaunchPackScan(const double* xyz, int64_t n, double* const planes[3], cudaStream_t stream);

// Voxel hash of the sparse map: the reference keeps its NDT map in a std::unordered_map keyed by
// a pairing of the voxel indices (tests/simple_optimization_test.cc:282-294); on the device the
// table is open-addressing with linear probing over 64-bit keys x | y << 21 | z << 42 (indices
// relative to the map's lowest voxel), at most half full.
constexpr unsigned long long kHashEmpty = ~0ull;
constexpr int kHashAxisBits = 21;

// Device NDT map builder (UpdateNdtMap of the reference's test mains) on a dense voxel grid or,
// when `keys` is set, on the voxel hash.
struct MapAccumParams {
  const double* xyz;  // interleaved points
  int64_t n;
  double inv_voxel;
  int kmin[3];
  int dims[3];
  double voxel;
  int* count;       // [cells]
  double* sums;     // [cells][9] 64-bit FIXED-POINT words (fixed_shift fractional bits): sum d (3) | moment d d^T:
                    // xx xy xz yy yz zz (6), d = (p - voxel centre) / voxel in [-1/2, 1/2]
  int fixed_shift;  // min(40, 61 - bits(n)): n points cannot overflow 63 bits
  unsigned long long* keys;  // hashed: [slots], kHashEmpty-initialised; cells = slots
  long long hash_mask;
};
// Inserts every point's voxel key into `keys` (scratch table, >= 2 n slots) and counts the
// distinct voxels, so that the final table can be sized from the occupancy instead of from n.
cudaError_t LaunchMapCountVoxels(const double* xyz, int64_t n, double inv_voxel, const int kmin[3],
                                 unsigned long long* keys, long long hash_mask, unsigned long long* distinct,
                                 cudaStream_t stream);
cudaError_t LaunchMapBounds(const double* xyz, int64_t n, double inv_voxel, int* bounds6, cudaStream_t stream);
cudaError_t LaunchMapAccumulate(const MapAccumParams& p, cudaStream_t stream);
// `p` as given to LaunchMapAccumulate (the voxel centres are rebuilt from kmin / dims / keys).
cudaError_t LaunchMapFinalize(const MapAccumParams& p, int64_t cells, int v_not_transposed, double* cell_mean,
                              double* cell_sqrt_info, unsigned char* cell_valid, cudaStream_t stream);

}  // namespace nlo

#endif  // NLO_INTERNAL_H_
