// gb_internal.cuh -- shared declarations of libglim_b200.so (not part of the public boundary).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/glim_b200.h"

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void gb_set_error(const char* fmt, ...);
#define GB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      gb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return e__ == cudaErrorMemoryAllocation ? GB_ERR_OUT_OF_MEMORY : GB_ERR_CUDA;           \
    }                                                                                         \
  } while (0)
#define GB_CHECK(st)                 \
  do {                               \
    gb_status s__ = (st);            \
    if (s__ != GB_OK) return s__;    \
  } while (0)
#define GB_REQUIRE(cond, msg)                       \
  do {                                              \
    if (!(cond)) {                                  \
      gb_set_error("invalid argument: %s", msg);    \
      return GB_ERR_INVALID_ARGUMENT;               \
    }                                               \
  } while (0)

// ---------------------------------------------------------------------------------------------
// HBM data layout (DESIGN.md "Data layout")
// ---------------------------------------------------------------------------------------------
// Source cloud: three planes, 36 B / point, every warp access fully coalesced:
//   p0[i] = {x, y, z, c00}   p1[i] = {c01, c02, c11, c12}   p2[i] = c22
// Voxel map: open-addressing table of 16-byte buckets {cx, cy, cz, voxel index (-1 = empty)} and
// 48-byte voxel records (3 x float4): {mx, my, mz, c00} {c01, c02, c11, c12} {c22, num_points, 0, 0}.
struct gb_cloud {
  int device;       // clouds / voxel maps do NOT keep their creating context: they outlive it when frames migrate between threads
  size_t n;
  float4* p0;
  float4* p1;
  float* p2;
  float4* normals;  // {nx, ny, nz, 0} or nullptr
  // Points are stored in Morton order of their 1/16 m cell (gather locality of the sweep kernel: lanes of a warp
  // then hit the same few voxels).  perm[j] = original index of stored point j, inv_perm = its inverse; nullptr = identity.
  int* perm;
  int* inv_perm;
  void* base;       // one allocation
  size_t bytes;
};

struct gb_voxelmap {
  int device;
  float resolution, inv_res;
  int max_scan;
  int num_voxels, num_buckets;
  int num_dropped_points;
  int4* buckets;
  float4* voxels;   // 3 float4 per voxel
  void* base;
  size_t bytes;
};

// device-side factor descriptor (80 B)
struct FactorDesc {
  const float4* p0;
  const float4* p1;
  const float* p2;
  const float4* normals;  // source normals, non-null only when surface validation is on for this factor
  const int4* buckets;
  const float4* voxels;
  uint32_t mask;
  int max_scan;
  float inv_res;
  int n;
  int pair;
  int flags;
  int num_tiles;
  int chunk;       // contiguous items: points per item of THIS factor (the last factors of a sweep get smaller items: tail tapering)
};
static_assert(sizeof(FactorDesc) == 80, "FactorDesc size");

struct gb_factor {
  gb_ctx* ctx;
  const gb_voxelmap* target;
  const gb_cloud* source;
  int flags;
  gb_sweep* single;  // lazily created 1-factor sweep
  float inlier_frac; // inlier fraction of the last linearization (< 0: unknown) -- sizes the work items of the next sweep
  uint64_t id;       // process-wide unique
  std::vector<gb_sweep*> users;  // sweeps (of any context) that reference this factor; guarded by the registry mutex
};

#define GB_MAX_PEERS 8
// device-resident parameter block of the fused result exchange (one per step parity)
struct PeerPush {
  int world;                      // 0 = disabled
  float* base[GB_MAX_PEERS];      // every rank's slab buffer of the current step parity, as mapped on THIS device
  const int* pair_ptr;            // CSR over global pair ids -> local factor indices (empty for pairs owned elsewhere)
  const int* pair_factors;
  unsigned* pair_done;            // per-pair tickets, self-cleaning
};

struct gb_peer_slab {
  gb_ctx* ctx;
  size_t num_pairs;
  int world, rank;
  size_t buf_floats;              // floats per buffer
  char* local;                    // cudaMalloc: [buffer 0][buffer 1][flags: world x u32, padded]
  char* peer[GB_MAX_PEERS];       // every rank's allocation as mapped here (peer[rank] == local)
  bool opened[GB_MAX_PEERS];
  unsigned step;                  // last launched step
  int parity;                     // buffer written by the NEXT launch
  int completed_parity;           // buffer completed by the last signal_wait
  int* d_timeout;
  float* h_pinned;                // num_pairs x GB_SLAB_STRIDE, for the fetches
  bool connected;
  // deferred exchange (default): the sweep stores finished pair rows into the LOCAL buffer only; the exchange kernel that
  // follows it copies this rank's rows to every peer (one CTA per peer) before it publishes the completion flags
  bool deferred;
  int* d_my_pairs;                // pair ids owned by the attached sweep
  int num_my_pairs;
};

#define GB_ACC_STRIDE 32      // doubles per factor in the accumulation buffer (29 used)
#define GB_OUT_DOUBLES 122    // gb_linearized6

struct gb_sweep {
  gb_ctx* ctx;
  size_t F;
  std::vector<gb_factor*> factors;
  FactorDesc* d_descs;
  int2* d_tiles;          // work items {factor, first point}, factor-major
  double* d_poses;        // F x 16 (T_lin)
  double* d_poses_eval;   // F x 16 (error mode)
  double* d_accum;        // F x acc_slots x GB_ACC_STRIDE, zero between sweeps (self-cleaning)
  int acc_slots;          // power of two: copies of each factor's accumulator (spreads same-address atomics of few-factor sweeps)
  unsigned* d_done;       // F tickets, zero between sweeps
  unsigned long long* d_tile_ctr;  // dynamic tile queue head, monotonic across launches
  unsigned long long ctr_base;     // value of the counter at the start of the next launch
  double* d_out;          // F x 122
  double* h_poses_eval;   // pinned
  double* h_out;          // pinned
  float* d_slab;
  size_t num_pairs;
  gb_peer_slab* peer;             // fused exchange target (or nullptr)
  int* d_pair_ptr;                // CSR pair -> factors (device), built when a peer slab is attached
  int* d_pair_factors;
  unsigned* d_pair_done;
  PeerPush* d_peer_tables;        // [2]: one per step parity
  std::vector<int> h_pair;        // pair id per factor
  int num_tiles, tile_size, grid;   // work items, points per item, CTAs
  int kernel_version;               // 5 = small sweeps (one wave of strided items); 3 = large sweeps (queue of contiguous items); 4 = bulk-async (TMA) staged experiment (GB_KERNEL=3/4/5 forces one)
  bool any_sv;                      // some factor of the sweep has surface validation on
  int pipe;                         // v5 software-pipelining variant (GB_PIPE; A/B only)
  int strided;                      // v5, about one item per warp: item j of a factor owns the rows j, j + J, ... (see k_vgicp_sweep5)
  bool calibrated;                  // strided: the item table has been re-sized from measured inlier fractions
  int capacity;                     // CTAs of a full grid
  FactorDesc* h_descs;              // pinned copies (re-uploaded when the item table is re-sized)
  int2* h_tiles;
  size_t tiles_cap;
  int stage_points;                 // v4: points per shared-memory stage (128: 2 CTAs / SM; 64: 3 CTAs / SM)
  uint64_t point_factors, algorithmic_bytes;
  uint64_t key;           // cache key
  bool stale;             // a factor of this sweep was destroyed: it can no longer be launched
  double* h_pose_slot[2]; // pinned pose staging, double buffered (no stream sync in gb_sweep_set_poses)
  cudaEvent_t pose_ev[2]; // recorded after the H2D that read the slot
  int pose_slot;
  cudaGraphExec_t graph_exec;       // small sweeps: poses H2D -> kernel -> records D2H as ONE graph launch (gb_factor_set_linearize)
  int graph_state;                  // 0 = not built, 1 = valid, -1 = capture failed (plain launches from then on)
  void* pool_d; size_t pool_d_cap;  // the blocks this sweep took from its context's pool
  void* pool_h; size_t pool_h_cap;
};

struct gb_pool_block { void* d; size_t d_cap; void* h; size_t h_cap; };

struct gb_ctx {
  int device;
  cudaStream_t stream;
  bool own_stream;
  int num_sms;
  void* scratch;
  size_t scratch_cap;
  void* pinned;
  size_t pinned_cap;
  uint64_t launches;
  std::atomic<int> refs;   // owner + live factors / sweeps / peer slabs; the context is torn down when the last one lets go
  std::vector<gb_sweep*> sweep_cache;
  std::vector<gb_pool_block> pool;  // device + pinned blocks of retired sweeps, reused by the next gb_sweep_create
  // A context may be driven from more than one host thread (a frame cloned by the odometry thread is later used by the
  // sub-mapping thread): every entry point that touches the stream, the scratch arena or the caches takes this lock.
  std::recursive_mutex mu;
};
#define GB_LOCK(ctx) std::lock_guard<std::recursive_mutex> gb_lock__((ctx)->mu)

// Device blocks of clouds / voxel maps come from a process-wide pool (a frame costs one cloud + two maps = five allocations;
// cudaMalloc / cudaFree are 50-200 us each and cudaFree synchronises the device).  gb_dev_free waits for the streams of
// every live context of the device (what the implicit synchronisation of cudaFree used to guarantee) and keeps the block.
cudaError_t gb_dev_malloc(int device, size_t bytes, void** out);
void gb_dev_free(int device, void* p);
gb_status gb_ctx_scratch(gb_ctx* ctx, size_t bytes, void** out);  // device scratch, valid until the next call
gb_status gb_ctx_pinned(gb_ctx* ctx, size_t bytes, void** out);   // pinned host staging, same lifetime rule

// kernel launchers (gb_kernels_*.cu)
enum { GB_MODE_LINEARIZE = 0, GB_MODE_ERROR = 1 };
gb_status gb_launch_sweep(gb_sweep* s, int mode);
size_t gb_sweep4_smem_bytes(int stage_points);
gb_status gb_launch_peer_signal_wait(gb_peer_slab* ps);
gb_status gb_launch_overlap(gb_ctx* ctx, int num_targets, const FactorDesc* d_descs, const double* d_poses, int n, int* d_count);
gb_status gb_cloud_reorder_impl(gb_ctx* ctx, gb_cloud* c, const void* staged /* device copy of the planes in original order */, size_t b0, size_t b1, size_t b2, size_t b3);
size_t gb_cloud_reorder_scratch_bytes(size_t n, size_t staged_bytes);
gb_status gb_voxelmap_build_impl(gb_ctx* ctx, const gb_cloud* cloud, float resolution, int init_buckets, int max_scan, double drop_rate, gb_voxelmap* out);
gb_status gb_covariances_impl(gb_ctx* ctx, size_t n, const double* xyzw, const int32_t* neighbors, int kc, int k, double* normals4, double* cov4x4);
gb_status gb_preprocess_impl(gb_ctx* ctx, size_t n, const double* xyzw, const double* times, const double* intensities, const gb_preprocess_params* P, gb_preprocessed* out, gb_cloud* cloud_out);
gb_status gb_merge_frames_impl(gb_ctx* ctx, int K, const gb_cloud* const* frames, const double* poses, double resolution, int target, unsigned long long seed, double* out_xyzw, double* out_cov4x4, size_t* num_out, gb_cloud* cloud_out);
gb_status gb_find_neighbors_pyramid_impl(gb_ctx* ctx, size_t n, const double* xyzw, int k, int32_t* neighbors);
gb_status gb_find_neighbors_impl(gb_ctx* ctx, size_t n, const double* xyzw, int k, int32_t* neighbors);
gb_status gb_voxelgrid_sampling_impl(gb_ctx* ctx, size_t n, const double* xyzw, const double* times, const double* intensities, double resolution, double* out_xyzw, double* out_times, double* out_intensities, size_t* num_out);

// ---------------------------------------------------------------------------------------------
// device helpers shared by kernels
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#include "gb_vgicp_math.cuh"  // gb_coord, gb_hash (shared with the host-compiled CPU test of the kernel arithmetic)
// packed 3 x 21-bit voxel key (ascending key = canonical voxel order); false if out of range
#define GB_KEY_OFFSET (1 << 20)
__device__ __forceinline__ bool gb_pack_key(int x, int y, int z, unsigned long long* key) {
  if (x < -GB_KEY_OFFSET || x >= GB_KEY_OFFSET || y < -GB_KEY_OFFSET || y >= GB_KEY_OFFSET || z < -GB_KEY_OFFSET || z >= GB_KEY_OFFSET) return false;
  *key = ((unsigned long long)(x + GB_KEY_OFFSET) << 42) | ((unsigned long long)(y + GB_KEY_OFFSET) << 21) | (unsigned long long)(z + GB_KEY_OFFSET);
  return true;
}
__device__ __forceinline__ void gb_unpack_key(unsigned long long key, int& x, int& y, int& z) {
  x = (int)((key >> 42) & 0x1FFFFF) - GB_KEY_OFFSET;
  y = (int)((key >> 21) & 0x1FFFFF) - GB_KEY_OFFSET;
  z = (int)(key & 0x1FFFFF) - GB_KEY_OFFSET;
}
// linear-probing lookup
__device__ __forceinline__ int gb_lookup(const int4* __restrict__ buckets, uint32_t mask, int max_scan, int cx, int cy, int cz) {
  const uint32_t h = gb_hash(cx, cy, cz);
  for (int i = 0; i < max_scan; i++) {
    const int4 b = __ldg(&buckets[(h + (uint32_t)i) & mask]);
    if (b.w < 0) return -1;
    if (b.x == cx && b.y == cy && b.z == cz) return b.w;
  }
  return -1;
}
#endif
