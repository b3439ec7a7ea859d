// Internal declarations shared by the CUDA translation units of libloamlivox_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/loamlivox_b200.h"

#define LL_WARP 32
#define LL_KNN 5

#define LL_CUDA(ctx, expr)                                                                        \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
      return LL_ERR_CUDA;                                                                         \
    }                                                                                             \
  } while (0)
#define LL_TRY(expr) do { int _s = (expr); if (_s != LL_OK) return _s; } while (0)

// ------------------------------------------------------------------------------------------------ device arrays
struct DevBuf {  // grow-only device allocation
  void* p = nullptr; size_t cap = 0;
  size_t floor = 0;   // smallest size ever allocated: a caller that knows how far the buffer will grow (the streaming mapper) sets it once, then nothing is
                      // reallocated in steady state (a cudaMalloc + cudaFree pair was measured at 100-800 ms on the GPU boxes: profiles/r2/bench_c3_*)
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    // first allocation: a little slack; regrowth: at least double, so that a buffer following a growing map is reallocated O(log n) times
    // (cudaFree synchronises the device and showed up as 10-250 ms spikes in the streaming mapper)
    size_t want = bytes + bytes / 8 + 256;
    if (want < ((size_t)8 << 20)) want = (size_t)8 << 20;   // floor: a (re)allocation costs 60-90 ms on this platform, 8 MB of a 180 GB HBM costs nothing
    if (p && want < 2 * cap) want = 2 * cap;
    if (want < floor) want = floor;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  cudaError_t reserve_floor(size_t bytes) { if (bytes > floor) floor = bytes; return reserve(floor); }   // allocate the floor now
  template <typename T> T* as() const { return (T*)p; }
};

// ------------------------------------------------------------------------------------------------ map index
// "Bucket tree": map points sorted along a 63-bit Hilbert curve (isotropic cells), cut into leaf buckets of 32 points
// (512 B = one coalesced warp load), with a 32-ary tree of axis-aligned boxes above them (a node record = its 32 children's
// boxes, 1 KB).  Search is exact: a box gives a true lower bound of the fp32 distance.  lo[l] holds level l's node records;
// hi[0] the base of the node array (top level first).
#define LL_MAX_LEVELS 8
struct BucketTree {
  int n = 0;             // points given (the count of finite ones stays on the device)
  int n_pad = 0;         // padded to a multiple of 32
  int n_levels = 0;      // number of box levels (>= 1 when n > 0)
  int level_count[LL_MAX_LEVELS] = {0};   // boxes per level (unpadded)
  float4* pts = nullptr;                  // [n_pad] x,y,z, w = __int_as_float(original index); pad = +inf
  float4* lo[LL_MAX_LEVELS] = {nullptr};  // [level_count padded to 32] box minima (w unused)
  float4* hi[LL_MAX_LEVELS] = {nullptr};
  float4* src = nullptr;                  // [n_src] the cloud as given to ll_map_build (x,y,z,intensity)
  int n_src = 0;
  float* d_bbox = nullptr;                // device: min xyz, max xyz of the finite points (6 floats inside `storage`)
  DevBuf storage;                          // one allocation backing all of the above
};

// Regular grid of cubic cells over the bounding box of the whole map (shard.cu): cell of a point = clamp(floor((p - origin) * inv_cell), 0, dims - 1)
// per axis, i.e. the border cells extend outwards without bound and every point of space has exactly one cell.
struct ShardGrid { float origin[3] = {0, 0, 0}; float cell = 1.f, inv_cell = 1.f; int dims[3] = {1, 1, 1}; };
__host__ __device__ inline int shard_cell_index(const ShardGrid& g, float x, float y, float z) {
  const float c[3] = {x, y, z}; int ci[3];
  for (int k = 0; k < 3; k++) { int v = (int)floorf((c[k] - g.origin[k]) * g.inv_cell); ci[k] = v < 0 ? 0 : (v > g.dims[k] - 1 ? g.dims[k] - 1 : v); }
  return (ci[2] * g.dims[1] + ci[1]) * g.dims[0] + ci[0];
}

struct ll_map {
  int device = 0;
  BucketTree corner, surf;
  // sharding (world == 1: everything owned).  A sharded snapshot indexes only the points of the cells this rank owns plus the halo of the match
  // gates around them; the owner table (rank per cell) lives on the device for the query side and on the host for ll_map_shard_info.
  int rank = 0, world = 1; float cell_size = 0.f;
  ShardGrid grid; float halo[2] = {0.f, 0.f}; DevBuf shard_owner; std::vector<int> h_owner; long long shard_total[2] = {0, 0};
};

// per-scan feature-extraction state (device)
struct ExtractState {
  int n = 0;
  float4* raw = nullptr;          // [n] x,y,z,reflectivity
  int* pt_type = nullptr; int* pt_label = nullptr;
  float* curvature = nullptr; float* view_angle = nullptr; float* depth_sq2 = nullptr; float* time_stamp = nullptr;
  float* polar_dis_sq2 = nullptr; int8_t* polar_dir = nullptr; uint8_t* self_mask = nullptr; uint8_t* cand = nullptr;
  int* cand_idx = nullptr; int* d_num_cand = nullptr;
  int* split_idx = nullptr;        // [<= n]
  int* scan_first = nullptr; int* scan_last = nullptr;
  double* d_time = nullptr;        // m_current_time of this scan (device scalar read by ex_point_kernel)
  int* d_meta = nullptr;           // [0]=n_split [1]=n_scans [2]=clutter_size
  double first_receive_time = -1, current_time = 0, last_maximum_time_stamp = 0;
};

// Solver / registration device state shared between kernels (lives in global memory)
struct RegDevState;  // defined in solve.cuh

struct ll_ctx {
  int device = 0;
  ll_config cfg;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  cudaEvent_t evp[5 * 16 + 2] = {nullptr};   // per-ICP-iteration phase events
  std::string err;
  uint64_t launches = 0;
  int hook_slots = 0;      // slot count left behind by ll_build_blocks for the parity hooks
  int last_nc = 0, last_ns = 0;   // features of the last registration (ll_last_features_dev)
  float last_full_min_t = 10000.f, last_full_max_t = -10000.f;   // find_min_max_intensity over the last front end's full cloud (laser_mapping.hpp:1336)
  int num_sms = 0;
  int knn_tma = 0;         // LL_KNN_TMA=1 in the environment at ll_ctx_create: the variant of knn_blocks_kernel that stages leaf buckets with cp.async.bulk (measurement only)
  // arenas
  DevBuf scratch;      // CUB temp storage
  DevBuf stage_in;     // raw uploads (PCL32 or XYZI16)
  ll_point_layout layout = {16, 0, 4, 8, 12, LL_I_FLOAT32};   // LL_FMT_STRIDED records
  DevBuf extract_buf;  // ExtractState arrays
  DevBuf feat_buf;     // feature clouds / voxel-grid temporaries
  DevBuf reg_buf;      // registration arrays
  void* pinned = nullptr; size_t pinned_cap = 0;   // pinned host staging for small D2H/H2D control blocks
  ExtractState ex;
  RegDevState* d_reg = nullptr;   // device
  struct SolveSync* d_sync = nullptr;   // device: exchange rows of the solver kernels (solve.cu)
  // multi-GPU
  int rank = 0, world = 1;
  cudaStream_t stream3 = nullptr; cudaEvent_t ev_fork3 = nullptr, ev_join3 = nullptr; DevBuf scratch3;   // second side stream: the petal bookkeeping of the extractor (whole_frame front end)
  cudaStream_t stream2 = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_it[2] = {nullptr, nullptr}; DevBuf scratch2, scratch_fe;   // side stream of the per-scan front end; the front end's own scratch arenas (fixed once the scan size is
                                 // known: the captured graph must never see them reallocated by the map refresh or the registration)
  // The per-scan front end (extract + get_features + 4 VoxelGrids + count read-back) replayed as ONE CUDA graph: ~60 launches whose
  // enqueue cost on the host (CUB dispatch included) was longer than their execution.  Re-captured when the shape or any buffer changes.
  struct FrontGraph { cudaGraphExec_t exec = nullptr; size_t n = 0; ll_pipeline_cfg pc; void* bufs[6] = {nullptr}; bool warm = false; uint64_t launches = 0; } fg;
  int reg_deblur = 0;                      // if_motion_deblur of the registration whose blocks are on the device
  int solve_world = 1;                     // world size the solver kernels all-reduce over (1 unless the map in use is sharded)
  void* comm_local = nullptr;              // this rank's staging slot (device memory, IPC-exported)
  void* comm_peers[8] = {nullptr};         // mapped peer slots (index = rank)
  void set_error(const std::string& s) { err = s; }
};

inline int ll_div_up(int a, int b) { return (a + b - 1) / b; }

// Residual-block cap (point_cloud_registration.hpp:232-238,339-345,434-458): the reference draws from a std::random_device-seeded mt19937
// (include/tools/tools_random.hpp:18-25).  The rule is kept, the numbers come from a counter-based generator keyed on the caller's seed
// (ll_reg_state::rng_seed), the ICP iteration, a stream (0 corner pre-skip, 1 surface pre-skip, 2 drop) and the feature / slot index:
// splitmix64 finaliser, top 24 bits -> float in [0,1).  Exported as ll_cap_uniform so that a test can pin it.
__host__ __device__ inline float ll_cap_uniform_f(int seed, int icp_iter, int stream, int index) {
  unsigned long long z = (unsigned long long)(unsigned)seed * 0x9E3779B97F4A7C15ull + (((unsigned long long)(unsigned)icp_iter << 40) | ((unsigned long long)(unsigned)stream << 32) | (unsigned long long)(unsigned)index);
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ull; z ^= z >> 27; z *= 0x94d049bb133111ebull; z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}
