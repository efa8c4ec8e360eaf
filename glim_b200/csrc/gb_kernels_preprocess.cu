// gb_kernels_preprocess.cu -- per-frame preprocess on the GPU (sm_100a): k-NN, covariance / normal
// estimation, voxel-grid downsampling.
//
// Replaces, at GLIM's call sites:
//   CloudPreprocessor::find_neighbors            src/glim/preprocess/cloud_preprocessor.cpp:190-221  (gtsam_points::KdTree::knn_search)
//   CloudCovarianceEstimation::estimate (PLANE)  src/glim/common/cloud_covariance_estimation.cpp:43-122, :181-196
//   gtsam_points::voxelgrid_sampling             src/glim/preprocess/cloud_preprocessor.cpp:108
// Oracles: go_knn_bruteforce, go_covariance_estimate, go_voxelgrid_sampling (oracle/glim_oracle.c).
// All three work in fp64 like the reference's host code (Vector4d / Matrix4d).
#include "gb_internal.cuh"

#include <cub/cub.cuh>
#include <cmath>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace {

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// exact k-NN, brute force over shared-memory tiles.  One query per thread, candidates streamed
// in ascending index with a strict '<' insertion, so ties resolve to the lower index exactly as
// in the oracle.  The query itself is a candidate (distance 0), as in the reference (:196).
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) k_knn_bruteforce(int n, const double4* __restrict__ pts, int* __restrict__ neighbors) {
  constexpr int TILE = 256;
  __shared__ double sx[TILE], sy[TILE], sz[TILE];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double px = 0, py = 0, pz = 0;
  if (i < n) { const double4 p = pts[i]; px = p.x; py = p.y; pz = p.z; }
  double bd[K];
  int bi[K];
#pragma unroll
  for (int k = 0; k < K; k++) { bd[k] = 1e300; bi[k] = i; }
  int cnt = 0;
  for (int base = 0; base < n; base += TILE) {
    __syncthreads();
    for (int t = threadIdx.x; t < TILE; t += blockDim.x) {
      const int j = base + t;
      if (j < n) { const double4 q = pts[j]; sx[t] = q.x; sy[t] = q.y; sz[t] = q.z; }
    }
    __syncthreads();
    const int m = min(TILE, n - base);
    if (i < n) {
      for (int t = 0; t < m; t++) {
        const double dx = px - sx[t], dy = py - sy[t], dz = pz - sz[t];
        // no FMA contraction: bit-identical to the oracle's (dx*dx + dy*dy) + dz*dz
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (d < bd[K - 1]) {
          // insert (d, base+t) keeping ascending distance; equal distances stay behind earlier ones
          double cd = d;
          int ci = base + t;
#pragma unroll
          for (int k = 0; k < K; k++) {
            if (cd < bd[k]) {
              const double td = bd[k]; const int ti = bi[k];
              bd[k] = cd; bi[k] = ci; cd = td; ci = ti;
            }
          }
          cnt++;
        }
      }
    }
  }
  if (i < n) {
    const int found = min(cnt, K);
#pragma unroll
    for (int k = 0; k < K; k++) neighbors[(size_t)i * K + k] = k < found ? bi[k] : i;
  }
}

// ---------------------------------------------------------------------------------------------
// exact k-NN on a uniform grid.  Points are sorted by the packed key of their cell (x major, z minor), so the cells of one
// (x, y) column with consecutive z are one contiguous key range: a ring of the search cube costs (2r+1)^2 binary searches.
// One query per thread, in sorted order (spatially coherent warps).  After ring r every unseen point lies outside the cube
// of cells [c-r, c+r]^3, i.e. at least m*h away (m = distance from the query to the nearest cube face, in cells), so the
// search stops as soon as the k-th best distance is within that bound -- the result is EXACT, with the same
// (distance, index) tie rule and the same un-contracted fp64 distance as the brute-force kernel and the oracle.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kKnnInvalid = ~0ull;

__global__ void k_knn_keys(int n, const double4* __restrict__ pts, double inv_h, unsigned long long* __restrict__ keys, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = pts[i];
  unsigned long long key = kKnnInvalid;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    const double fx = floor(p.x * inv_h), fy = floor(p.y * inv_h), fz = floor(p.z * inv_h);
    if (fabs(fx) < 1048575.0 && fabs(fy) < 1048575.0 && fabs(fz) < 1048575.0) {
      unsigned long long k;
      if (gb_pack_key((int)fx, (int)fy, (int)fz, &k)) key = k;
    }
  }
  keys[i] = key;
  idx[i] = i;
}
__global__ void k_knn_gather(int n, const int* __restrict__ idx_s, const double4* __restrict__ pts, double4* __restrict__ pts_s) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) pts_s[s] = pts[idx_s[s]];
}
__global__ void k_count_heads(int n, const unsigned long long* __restrict__ keys_s, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int head = 0;
  if (i < n) { const unsigned long long k = keys_s[i]; head = (k != kKnnInvalid && (i == 0 || keys_s[i - 1] != k)) ? 1 : 0; }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) head += __shfl_xor_sync(0xffffffffu, head, o);
  if ((threadIdx.x & 31) == 0 && head) atomicAdd(count, head);
}

__device__ __forceinline__ int knn_lower_bound(const unsigned long long* __restrict__ keys, int n, unsigned long long key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(&keys[mid]) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <int K>
__global__ void __launch_bounds__(128) k_knn_grid(int n, const double4* __restrict__ pts_s, const unsigned long long* __restrict__ keys_s, const int* __restrict__ idx_s,
                                                  double inv_h, double h, int3 cmin, int3 cmax, int max_ring, int* __restrict__ neighbors) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int self = idx_s[t];
  const unsigned long long key = keys_s[t];
  double bd[K];
  int bi[K];
#pragma unroll
  for (int k = 0; k < K; k++) { bd[k] = 1e300; bi[k] = 0x7fffffff; }
  int cnt = 0;
  if (key != kKnnInvalid) {
    const double4 p = pts_s[t];
    int cx, cy, cz;
    gb_unpack_key(key, cx, cy, cz);
    const double ux = p.x * inv_h - (double)cx, uy = p.y * inv_h - (double)cy, uz = p.z * inv_h - (double)cz;  // in [0, 1)
    // the cube must eventually cover every occupied cell
    const int rmax = max(max(max(cx - cmin.x, cmax.x - cx), max(cy - cmin.y, cmax.y - cy)), max(cz - cmin.z, cmax.z - cz));
    // ring r costs (2r+1)^2 range searches: an isolated point (a far return with no neighbours nearby) would walk thousands of
    // empty rings.  After max_ring rings (chosen on the host so that the ring walk costs about as much as one pass over all
    // points) without a proof of completeness the query falls back to a scan of ALL points (exact, O(N)).
    bool complete = false;
    for (int r = 0; r <= min(rmax, max_ring); r++) {
      for (int dx = -r; dx <= r; dx++) {
        const int x = cx + dx;
        if (x < cmin.x || x > cmax.x) continue;
        for (int dy = -r; dy <= r; dy++) {
          const int y = cy + dy;
          if (y < cmin.y || y > cmax.y) continue;
          const bool edge = (abs(dx) == r) || (abs(dy) == r);
          // edge columns contribute their whole z range; interior columns only the two new cells at z = cz -+ r
          const int nseg = edge ? 1 : 2;
          for (int sgi = 0; sgi < nseg; sgi++) {
            int z0, z1;
            if (edge) { z0 = cz - r; z1 = cz + r; } else { z0 = z1 = (sgi == 0) ? cz - r : cz + r; }
            z0 = max(z0, cmin.z); z1 = min(z1, cmax.z);
            if (z0 > z1) continue;
            unsigned long long klo, khi;
            gb_pack_key(x, y, z0, &klo);
            gb_pack_key(x, y, z1, &khi);
            for (int s = knn_lower_bound(keys_s, n, klo); s < n; s++) {
              if (__ldg(&keys_s[s]) > khi) break;
              const double4 q = pts_s[s];
              const double ex = p.x - q.x, ey = p.y - q.y, ez = p.z - q.z;
              const double d = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
              const int j = __ldg(&idx_s[s]);
              if (d < bd[K - 1] || (d == bd[K - 1] && j < bi[K - 1])) {
                double cd = d;
                int ci = j;
#pragma unroll
                for (int k = 0; k < K; k++) {
                  if (cd < bd[k] || (cd == bd[k] && ci < bi[k])) {
                    const double td = bd[k]; const int ti = bi[k];
                    bd[k] = cd; bi[k] = ci; cd = td; ci = ti;
                  }
                }
                cnt++;
              }
            }
          }
        }
      }
      if (r == rmax) { complete = true; break; }  // the cube covers every occupied cell
      if (cnt >= K) {
        const double m = fmin(fmin(fmin(ux, 1.0 - ux), fmin(uy, 1.0 - uy)), fmin(uz, 1.0 - uz)) + (double)r;
        const double bound = m * h * (1.0 - 1e-12);
        if (bd[K - 1] <= bound * bound) { complete = true; break; }
      }
    }
    if (!complete) {
#pragma unroll
      for (int k = 0; k < K; k++) { bd[k] = 1e300; bi[k] = 0x7fffffff; }
      cnt = 0;
      for (int s = 0; s < n; s++) {
        if (__ldg(&keys_s[s]) == kKnnInvalid) break;  // invalid points sort last
        const double4 q = pts_s[s];
        const double ex = p.x - q.x, ey = p.y - q.y, ez = p.z - q.z;
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
        const int j = __ldg(&idx_s[s]);
        if (d < bd[K - 1] || (d == bd[K - 1] && j < bi[K - 1])) {
          double cd = d;
          int ci = j;
#pragma unroll
          for (int k = 0; k < K; k++) {
            if (cd < bd[k] || (cd == bd[k] && ci < bi[k])) {
              const double td = bd[k]; const int ti = bi[k];
              bd[k] = cd; bi[k] = ci; cd = td; ci = ti;
            }
          }
          cnt++;
        }
      }
    }
  }
  const int found = min(cnt, K);
#pragma unroll
  for (int k = 0; k < K; k++) neighbors[(size_t)self * K + k] = k < found ? bi[k] : self;
}

// ---------------------------------------------------------------------------------------------
// closed-form symmetric 3x3 eigen decomposition (same published algorithm as the oracle's
// go_eigen_sym3_direct / Eigen's computeDirect): evals ascending, V[r*3+k] = k-th eigenvector.
// ---------------------------------------------------------------------------------------------
__device__ inline void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ inline void extract_kernel(const double* m, double* res, double* rep) {
  int i0 = 0;
  double best = fabs(m[0]);
  if (fabs(m[4]) > best) { best = fabs(m[4]); i0 = 1; }
  if (fabs(m[8]) > best) { best = fabs(m[8]); i0 = 2; }
  double c1v[3], c2v[3];
  const int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
  for (int r = 0; r < 3; r++) { rep[r] = m[r * 3 + i0]; c1v[r] = m[r * 3 + i1]; c2v[r] = m[r * 3 + i2]; }
  double c0[3], c1[3];
  cross3(rep, c1v, c0);
  cross3(rep, c2v, c1);
  const double n0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
  const double n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
  if (n0 > n1) { const double s = 1.0 / sqrt(n0); for (int k = 0; k < 3; k++) res[k] = c0[k] * s; }
  else if (n1 > 0.0) { const double s = 1.0 / sqrt(n1); for (int k = 0; k < 3; k++) res[k] = c1[k] * s; }
  else { res[0] = 1; res[1] = 0; res[2] = 0; }
}
__device__ void eigen_sym3_direct(const double* A, double* evals, double* V) {
  const double shift = (A[0] + A[4] + A[8]) / 3.0;
  double m[9];
  for (int k = 0; k < 9; k++) m[k] = A[k];
  m[0] -= shift; m[4] -= shift; m[8] -= shift;
  double scale = 0.0;
  for (int k = 0; k < 9; k++) scale = fmax(scale, fabs(m[k]));
  if (scale > 0.0) for (int k = 0; k < 9; k++) m[k] /= scale;
  const double m00 = m[0], m11 = m[4], m22 = m[8], m10 = m[3], m20 = m[6], m21 = m[7];
  const double c0 = m00 * m11 * m22 + 2.0 * m10 * m20 * m21 - m00 * m21 * m21 - m11 * m20 * m20 - m22 * m10 * m10;
  const double c1 = m00 * m11 - m10 * m10 + m00 * m22 - m20 * m20 + m11 * m22 - m21 * m21;
  const double c2 = m00 + m11 + m22;
  const double c2_3 = c2 / 3.0;
  double a_3 = (c2 * c2_3 - c1) / 3.0;
  if (a_3 < 0.0) a_3 = 0.0;
  const double half_b = 0.5 * (c0 + c2_3 * (2.0 * c2_3 * c2_3 - c1));
  double qq = a_3 * a_3 * a_3 - half_b * half_b;
  if (qq < 0.0) qq = 0.0;
  const double rho = sqrt(a_3);
  const double theta = atan2(sqrt(qq), half_b) / 3.0;
  const double ct = cos(theta), st = sin(theta);
  const double s3 = 1.7320508075688772935;
  double ev[3];
  ev[0] = c2_3 - rho * (ct + s3 * st);
  ev[1] = c2_3 - rho * (ct - s3 * st);
  ev[2] = c2_3 + 2.0 * rho * ct;
  const double eps = 2.220446049250313e-16;
  if ((ev[2] - ev[0]) <= eps) {
    for (int k = 0; k < 9; k++) V[k] = 0.0;
    V[0] = V[4] = V[8] = 1.0;
  } else {
    double d0 = ev[2] - ev[1], d1 = ev[1] - ev[0];
    int k = 0, l = 2;
    if (d0 > d1) { const double t = d0; d0 = d1; d1 = t; k = 2; l = 0; }
    double tmp[9], vk[3], vl[3], rep[3];
    for (int e = 0; e < 9; e++) tmp[e] = m[e];
    tmp[0] -= ev[k]; tmp[4] -= ev[k]; tmp[8] -= ev[k];
    extract_kernel(tmp, vk, rep);
    if (d0 <= 2.0 * eps * d1) {
      const double dp = vk[0] * rep[0] + vk[1] * rep[1] + vk[2] * rep[2];
      for (int r = 0; r < 3; r++) vl[r] = rep[r] - dp * vk[r];
      const double nl = sqrt(vl[0] * vl[0] + vl[1] * vl[1] + vl[2] * vl[2]);
      if (nl > 0) for (int r = 0; r < 3; r++) vl[r] /= nl;
    } else {
      double dummy[3];
      for (int e = 0; e < 9; e++) tmp[e] = m[e];
      tmp[0] -= ev[l]; tmp[4] -= ev[l]; tmp[8] -= ev[l];
      extract_kernel(tmp, vl, dummy);
    }
    double v0[3], v1[3], v2[3];
    for (int r = 0; r < 3; r++) { v0[r] = (k == 0) ? vk[r] : vl[r]; v2[r] = (k == 0) ? vl[r] : vk[r]; }
    cross3(v2, v0, v1);
    const double n1 = sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
    if (n1 > 0) for (int r = 0; r < 3; r++) v1[r] /= n1;
    for (int r = 0; r < 3; r++) { V[r * 3 + 0] = v0[r]; V[r * 3 + 1] = v1[r]; V[r * 3 + 2] = v2[r]; }
  }
  for (int k = 0; k < 3; k++) evals[k] = ev[k] * scale + shift;
}

// CloudCovarianceEstimation::estimate, calc_cov (cloud_covariance_estimation.cpp:80-102).
// The reference materialises pt_cross = p p^T per point (:58-63) and gathers 128-byte rows; here the
// outer products are formed in registers from the gathered neighbour points (same sums, same order).
__global__ void __launch_bounds__(128) k_covariances(int n, const double4* __restrict__ pts, const int* __restrict__ neighbors, int kc, int k, double4* __restrict__ normals, double* __restrict__ covs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double S[4] = {0, 0, 0, 0};
  double X[16];
  for (int e = 0; e < 16; e++) X[e] = 0.0;
  const size_t begin = (size_t)kc * (size_t)i;
  for (int j = 0; j < k; j++) {
    const double4 q = pts[neighbors[begin + j]];
    const double p[4] = {q.x, q.y, q.z, q.w};
    for (int r = 0; r < 4; r++) S[r] += p[r];
    for (int c = 0; c < 4; c++)
      for (int r = 0; r < 4; r++) X[c * 4 + r] += p[r] * p[c];
  }
  double mean[4], A[9];
  for (int r = 0; r < 4; r++) mean[r] = S[r] / k;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) A[r * 3 + c] = (X[c * 4 + r] - mean[r] * S[c]) / k;
  double evals[3], V[9];
  eigen_sym3_direct(A, evals, V);
  const double values[3] = {1e-3, 1.0, 1.0};
  double* C = covs + 16 * (size_t)i;
  for (int e = 0; e < 16; e++) C[e] = 0.0;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int e = 0; e < 3; e++) s += V[r * 3 + e] * values[e] * V[c * 3 + e];
      C[c * 4 + r] = s;
    }
  const double4 p = pts[i];
  double nx = V[0], ny = V[3], nz = V[6];
  if (p.x * nx + p.y * ny + p.z * nz > 0.0) { nx = -nx; ny = -ny; nz = -nz; }
  normals[i] = make_double4(nx, ny, nz, 0.0);
}

// ---------------------------------------------------------------------------------------------
// voxel-grid downsampling (fp64 coordinates)
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kInvalidKey = ~0ull;

__global__ void k_grid_keys(int n, const double4* __restrict__ pts, double inv_res, unsigned long long* __restrict__ keys, int* __restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 p = pts[i];
  unsigned long long key = kInvalidKey;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    const double fx = floor(p.x * inv_res), fy = floor(p.y * inv_res), fz = floor(p.z * inv_res);
    if (fabs(fx) < 2e9 && fabs(fy) < 2e9 && fabs(fz) < 2e9) {
      unsigned long long k;
      if (gb_pack_key((int)fx, (int)fy, (int)fz, &k)) key = k;
    }
  }
  keys[i] = key;
  idx[i] = i;
}
__global__ void k_grid_flags(int n, const unsigned long long* __restrict__ keys, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  flags[i] = (k != kInvalidKey && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}
__global__ void k_grid_starts(int n, const unsigned long long* __restrict__ keys, const int* __restrict__ flags, const int* __restrict__ pos, int* __restrict__ starts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i]) starts[pos[i] - 1] = i;
  const bool valid = keys[i] != kInvalidKey;
  const bool next_valid = (i + 1 < n) && keys[i + 1] != kInvalidKey;
  if (valid && !next_valid) starts[pos[i]] = i + 1;
}
__global__ void k_grid_means(int V, const int* __restrict__ starts, const int* __restrict__ idx, const double4* __restrict__ pts, const double* __restrict__ times, const double* __restrict__ intens,
                             double4* __restrict__ out_pts, double* __restrict__ out_times, double* __restrict__ out_intens) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int b = starts[v], e = starts[v + 1];
  double sx = 0, sy = 0, sz = 0, sw = 0, st = 0, si = 0;
  for (int s = b; s < e; s++) {
    const int i = idx[s];
    const double4 p = pts[i];
    sx += p.x; sy += p.y; sz += p.z; sw += p.w;
    if (times) st += times[i];
    if (intens) si += intens[i];
  }
  const int cnt = e - b;
  out_pts[v] = make_double4(sx / cnt, sy / cnt, sz / cnt, sw / cnt);
  if (times) out_times[v] = st / cnt;
  if (intens) out_intens[v] = si / cnt;
}

}  // namespace

template <int K>
static void launch_knn(bool grid, int n, const double4* d_pts, const double4* d_pts_s, const unsigned long long* d_keys_s, const int* d_idx_s, double inv_h, double h, int3 cmin, int3 cmax, int* d_nb, cudaStream_t st) {
  const int tb = 128, gb = (n + tb - 1) / tb;
  // (4/3) R^3 range searches of log2(n) steps each ~ n  =>  R = cbrt(0.75 n / log2 n)
  const int max_ring = std::min(64, std::max(6, (int)cbrt(0.75 * (double)n / std::max(1.0, log2((double)n)))));
  if (grid) k_knn_grid<K><<<gb, tb, 0, st>>>(n, d_pts_s, d_keys_s, d_idx_s, inv_h, h, cmin, cmax, max_ring, d_nb);
  else k_knn_bruteforce<K><<<gb, tb, 0, st>>>(n, d_pts, d_nb);
}

gb_status gb_find_neighbors_impl(gb_ctx* ctx, size_t n_, const double* xyzw, int k, int32_t* neighbors) {
  const int n = (int)n_;
  if (n == 0) return GB_OK;
  cudaStream_t st = ctx->stream;
  const char* mode = getenv("GB_KNN");
  const bool grid = mode ? (strcmp(mode, "grid") == 0) : (n >= 4096);
  size_t cub_sort = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_sort, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, n, 0, 64, st);
  const size_t pts_b = align_up(sizeof(double4) * (size_t)n, 256), nb_b = align_up(sizeof(int) * (size_t)n * k, 256);
  const size_t key_b = align_up(sizeof(unsigned long long) * (size_t)n, 256), idx_b = align_up(sizeof(int) * (size_t)n, 256), cub_b = align_up(cub_sort, 256);
  char* base = nullptr;
  GB_CHECK(gb_ctx_scratch(ctx, 2 * pts_b + nb_b + 2 * key_b + 2 * idx_b + cub_b + 256, (void**)&base));
  char* p = base;
  double4* d_pts = (double4*)p; p += pts_b;
  double4* d_pts_s = (double4*)p; p += pts_b;
  int* d_nb = (int*)p; p += nb_b;
  unsigned long long* d_keys = (unsigned long long*)p; p += key_b;
  unsigned long long* d_keys_s = (unsigned long long*)p; p += key_b;
  int* d_idx = (int*)p; p += idx_b;
  int* d_idx_s = (int*)p; p += idx_b;
  void* d_cub = p; p += cub_b;
  int* d_count = (int*)p;
  GB_CUDA(cudaMemcpyAsync(d_pts, xyzw, sizeof(double4) * (size_t)n, cudaMemcpyHostToDevice, st));
  double h = 0.25, inv_h = 4.0;
  int3 cmin = make_int3(0, 0, 0), cmax = make_int3(0, 0, 0);
  if (grid) {
    const int tb = 256, gb = (n + tb - 1) / tb;
    // cell size: measure the occupancy at 0.25 m, then aim for ~4 points per occupied cell (points lie on surfaces:
    // points per cell grows with h^2); a second pass re-sorts at the chosen size
    for (int pass = 0; pass < 2; pass++) {
      inv_h = 1.0 / h;
      k_knn_keys<<<gb, tb, 0, st>>>(n, d_pts, inv_h, d_keys, d_idx);
      size_t tmp = cub_b;
      GB_CUDA(cub::DeviceRadixSort::SortPairs(d_cub, tmp, d_keys, d_keys_s, d_idx, d_idx_s, n, 0, 64, st));
      ctx->launches += 2;
      if (pass == 1) break;
      int occ = 0;
      GB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), st));
      k_count_heads<<<gb, tb, 0, st>>>(n, d_keys_s, d_count);
      GB_CUDA(cudaMemcpyAsync(&occ, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
      GB_CUDA(cudaStreamSynchronize(st));
      ctx->launches++;
      const double per_cell = occ > 0 ? (double)n / occ : 1.0;
      double h2 = h * sqrt(4.0 / per_cell);
      h2 = fmin(4.0, fmax(0.02, h2));
      if (fabs(h2 - h) < 0.1 * h) break;  // close enough: keep the first sort
      h = h2;
    }
    inv_h = 1.0 / h;
    k_knn_gather<<<gb, tb, 0, st>>>(n, d_idx_s, d_pts, d_pts_s);
    ctx->launches++;
    // bounding box of the occupied cells (host: one pass over the caller's array, same floor(p * inv_h) as the kernel)
    bool any = false;
    for (int i = 0; i < n; i++) {
      const double* q = xyzw + 4 * (size_t)i;
      if (!(std::isfinite(q[0]) && std::isfinite(q[1]) && std::isfinite(q[2]))) continue;
      const double fx = floor(q[0] * inv_h), fy = floor(q[1] * inv_h), fz = floor(q[2] * inv_h);
      if (!(fabs(fx) < 1048575.0 && fabs(fy) < 1048575.0 && fabs(fz) < 1048575.0)) continue;
      const int x = (int)fx, y = (int)fy, z = (int)fz;
      if (!any) { cmin = cmax = make_int3(x, y, z); any = true; }
      cmin.x = std::min(cmin.x, x); cmin.y = std::min(cmin.y, y); cmin.z = std::min(cmin.z, z);
      cmax.x = std::max(cmax.x, x); cmax.y = std::max(cmax.y, y); cmax.z = std::max(cmax.z, z);
    }
  }
  switch (k) {
#define GB_KNN_CASE(K) case K: launch_knn<K>(grid, n, d_pts, d_pts_s, d_keys_s, d_idx_s, inv_h, h, cmin, cmax, d_nb, st); break;
    GB_KNN_CASE(1) GB_KNN_CASE(2) GB_KNN_CASE(3) GB_KNN_CASE(4) GB_KNN_CASE(5) GB_KNN_CASE(6) GB_KNN_CASE(7) GB_KNN_CASE(8)
    GB_KNN_CASE(9) GB_KNN_CASE(10) GB_KNN_CASE(12) GB_KNN_CASE(15) GB_KNN_CASE(16) GB_KNN_CASE(20) GB_KNN_CASE(24) GB_KNN_CASE(32)
#undef GB_KNN_CASE
    default:
      gb_set_error("k = %d is not an instantiated neighbour count (1-10, 12, 15, 16, 20, 24, 32)", k);
      return GB_ERR_INVALID_ARGUMENT;
  }
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  GB_CUDA(cudaMemcpyAsync(neighbors, d_nb, sizeof(int) * (size_t)n * k, cudaMemcpyDeviceToHost, st));
  GB_CUDA(cudaStreamSynchronize(st));
  return GB_OK;
}

gb_status gb_covariances_impl(gb_ctx* ctx, size_t n_, const double* xyzw, const int32_t* neighbors, int kc, int k, double* normals4, double* cov4x4) {
  const int n = (int)n_;
  if (n == 0) return GB_OK;  // cloud_covariance_estimation.cpp:49-51
  cudaStream_t st = ctx->stream;
  const size_t pts_b = align_up(sizeof(double4) * (size_t)n, 256), nb_b = align_up(sizeof(int) * (size_t)n * kc, 256);
  const size_t nrm_b = pts_b, cov_b = align_up(sizeof(double) * 16 * (size_t)n, 256);
  char* base = nullptr;
  GB_CHECK(gb_ctx_scratch(ctx, pts_b + nb_b + nrm_b + cov_b, (void**)&base));
  double4* d_pts = (double4*)base;
  int* d_nb = (int*)(base + pts_b);
  double4* d_nrm = (double4*)(base + pts_b + nb_b);
  double* d_cov = (double*)(base + pts_b + nb_b + nrm_b);
  GB_CUDA(cudaMemcpyAsync(d_pts, xyzw, sizeof(double4) * (size_t)n, cudaMemcpyHostToDevice, st));
  GB_CUDA(cudaMemcpyAsync(d_nb, neighbors, sizeof(int) * (size_t)n * kc, cudaMemcpyHostToDevice, st));
  k_covariances<<<(n + 127) / 128, 128, 0, st>>>(n, d_pts, d_nb, kc, k, d_nrm, d_cov);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  GB_CUDA(cudaMemcpyAsync(normals4, d_nrm, sizeof(double4) * (size_t)n, cudaMemcpyDeviceToHost, st));
  GB_CUDA(cudaMemcpyAsync(cov4x4, d_cov, sizeof(double) * 16 * (size_t)n, cudaMemcpyDeviceToHost, st));
  GB_CUDA(cudaStreamSynchronize(st));
  return GB_OK;
}

gb_status gb_voxelgrid_sampling_impl(gb_ctx* ctx, size_t n_, const double* xyzw, const double* times, const double* intensities, double resolution, double* out_xyzw, double* out_times, double* out_intensities, size_t* num_out) {
  const int n = (int)n_;
  *num_out = 0;
  if (n == 0) return GB_OK;
  cudaStream_t st = ctx->stream;
  size_t cub_sort = 0, cub_scan = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_sort, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, n, 0, 64, st);
  cub::DeviceScan::InclusiveSum(nullptr, cub_scan, (int*)nullptr, (int*)nullptr, n, st);
  const size_t cub_b = align_up(cub_sort > cub_scan ? cub_sort : cub_scan, 256);
  const size_t pts_b = align_up(sizeof(double4) * (size_t)n, 256), d_b = align_up(sizeof(double) * (size_t)n, 256);
  const size_t key_b = align_up(sizeof(unsigned long long) * (size_t)n, 256), i_b = align_up(sizeof(int) * (size_t)(n + 1), 256);
  char* base = nullptr;
  GB_CHECK(gb_ctx_scratch(ctx, cub_b + 2 * pts_b + 4 * d_b + 2 * key_b + 5 * i_b, (void**)&base));
  char* p = base;
  void* d_cub = p; p += cub_b;
  double4* d_pts = (double4*)p; p += pts_b;
  double4* d_opts = (double4*)p; p += pts_b;
  double* d_t = (double*)p; p += d_b;
  double* d_i = (double*)p; p += d_b;
  double* d_ot = (double*)p; p += d_b;
  double* d_oi = (double*)p; p += d_b;
  unsigned long long* d_keys = (unsigned long long*)p; p += key_b;
  unsigned long long* d_keys_s = (unsigned long long*)p; p += key_b;
  int* d_idx = (int*)p; p += i_b;
  int* d_idx_s = (int*)p; p += i_b;
  int* d_flags = (int*)p; p += i_b;
  int* d_pos = (int*)p; p += i_b;
  int* d_starts = (int*)p; p += i_b;
  GB_CUDA(cudaMemcpyAsync(d_pts, xyzw, sizeof(double4) * (size_t)n, cudaMemcpyHostToDevice, st));
  if (times) GB_CUDA(cudaMemcpyAsync(d_t, times, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  if (intensities) GB_CUDA(cudaMemcpyAsync(d_i, intensities, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
  const int tb = 256, gb = (n + tb - 1) / tb;
  k_grid_keys<<<gb, tb, 0, st>>>(n, d_pts, 1.0 / resolution, d_keys, d_idx);
  size_t tmp = cub_b;
  GB_CUDA(cub::DeviceRadixSort::SortPairs(d_cub, tmp, d_keys, d_keys_s, d_idx, d_idx_s, n, 0, 64, st));
  k_grid_flags<<<gb, tb, 0, st>>>(n, d_keys_s, d_flags);
  tmp = cub_b;
  GB_CUDA(cub::DeviceScan::InclusiveSum(d_cub, tmp, d_flags, d_pos, n, st));
  int V = 0;
  GB_CUDA(cudaMemcpyAsync(&V, d_pos + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  GB_CUDA(cudaStreamSynchronize(st));
  ctx->launches += 4;
  if (V > 0) {
    k_grid_starts<<<gb, tb, 0, st>>>(n, d_keys_s, d_flags, d_pos, d_starts);
    k_grid_means<<<(V + 127) / 128, 128, 0, st>>>(V, d_starts, d_idx_s, d_pts, times ? d_t : nullptr, intensities ? d_i : nullptr, d_opts, d_ot, d_oi);
    GB_CUDA(cudaGetLastError());
    ctx->launches += 2;
    GB_CUDA(cudaMemcpyAsync(out_xyzw, d_opts, sizeof(double4) * (size_t)V, cudaMemcpyDeviceToHost, st));
    if (times && out_times) GB_CUDA(cudaMemcpyAsync(out_times, d_ot, sizeof(double) * (size_t)V, cudaMemcpyDeviceToHost, st));
    if (intensities && out_intensities) GB_CUDA(cudaMemcpyAsync(out_intensities, d_oi, sizeof(double) * (size_t)V, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
  }
  *num_out = (size_t)V;
  return GB_OK;
}
