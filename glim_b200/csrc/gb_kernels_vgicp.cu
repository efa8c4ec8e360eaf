// gb_kernels_vgicp.cu -- the fused VGICP linearize / error sweep and the overlap kernel (sm_100a).
//
// Replaces gtsam_points::IntegratedVGICPFactorGPU::linearize / ::error as driven by
// NonlinearFactorSetGPU (GLIM call sites: src/glim/odometry/odometry_estimation_gpu.cpp:144,161,383-386;
// src/glim/mapping/sub_mapping.cpp:307; src/glim/mapping/global_mapping.cpp:466) and
// gtsam_points::overlap_gpu (odometry_estimation_gpu.cpp:231,248).  Math: SURVEY.md Appendix A;
// oracle: go_vgicp_linearize_gpumap / go_vgicp_error_gpumap / go_overlap_gpumap (oracle/glim_oracle.c).
//
// One launch covers a whole factor set.  Work unit = an item of up to `chunk` consecutive source
// points of one factor; items are laid out factor-major and drawn from a global queue by the warps of
// a persistent grid, so at any instant the grid works on a window of a few consecutive factors whose
// source cloud and voxel table stay L2 resident.  Per inlier (one lane):
//   q = R a + t -> voxel coord -> hash probe (16-byte buckets) -> 48-byte voxel record ->
//   S = C_B + R C_A R^T, M = S^-1 (symmetric 3x3) -> accumulate the 21 unique entries of
//   H_tt = J_t^T M J_t (J_t = [-hat(q) | I]), the 6 of b_t = J_t^T M r, the error r^T M r and the
//   inlier count: 29 registers.
// Per item: transposing warp reduce-scatter (31 shuffles for 32 values) and 29 fp64 atomics into the
// factor's accumulator.  The warp that retires a factor's last item runs the fp64 epilogue:
// H_ts = -H_tt Ad, H_ss = Ad^T H_tt Ad, b_s = -Ad^T b_t with Ad = AdjointMap(delta),
// writes the 122-double record (and adds it to the pair slab when one is attached), and re-zeroes
// the accumulator for the next sweep.
#include "gb_internal.cuh"

#include <string.h>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct PoseF {
  float r00, r01, r02, r10, r11, r12, r20, r21, r22, tx, ty, tz;
};

__device__ __forceinline__ PoseF load_pose(const float* s) {
  PoseF P;
  P.r00 = s[0]; P.r01 = s[1]; P.r02 = s[2];
  P.r10 = s[3]; P.r11 = s[4]; P.r12 = s[5];
  P.r20 = s[6]; P.r21 = s[7]; P.r22 = s[8];
  P.tx = s[9]; P.ty = s[10]; P.tz = s[11];
  return P;
}

// q = R a + t with the canonical FMA order (bit-exact with transform_f32 in the oracle)
__device__ __forceinline__ void transform(const PoseF& P, float ax, float ay, float az, float& qx, float& qy, float& qz) {
  qx = fmaf(P.r00, ax, fmaf(P.r01, ay, fmaf(P.r02, az, P.tx)));
  qy = fmaf(P.r10, ax, fmaf(P.r11, ay, fmaf(P.r12, az, P.ty)));
  qz = fmaf(P.r20, ax, fmaf(P.r21, ay, fmaf(P.r22, az, P.tz)));
}

// M = (C_B + R C_A R^T)^-1, symmetric 3x3 (xx xy xz yy yz zz)
__device__ __forceinline__ void fused_mahalanobis(
  const PoseF& P, float cxx, float cxy, float cxz, float cyy, float cyz, float czz,  // C_A
  float bxx, float bxy, float bxz, float byy, float byz, float bzz,                  // C_B
  float& mxx, float& mxy, float& mxz, float& myy, float& myz, float& mzz) {
  // T = R * C_A
  const float t00 = P.r00 * cxx + P.r01 * cxy + P.r02 * cxz;
  const float t01 = P.r00 * cxy + P.r01 * cyy + P.r02 * cyz;
  const float t02 = P.r00 * cxz + P.r01 * cyz + P.r02 * czz;
  const float t10 = P.r10 * cxx + P.r11 * cxy + P.r12 * cxz;
  const float t11 = P.r10 * cxy + P.r11 * cyy + P.r12 * cyz;
  const float t12 = P.r10 * cxz + P.r11 * cyz + P.r12 * czz;
  const float t20 = P.r20 * cxx + P.r21 * cxy + P.r22 * cxz;
  const float t21 = P.r20 * cxy + P.r21 * cyy + P.r22 * cyz;
  const float t22 = P.r20 * cxz + P.r21 * cyz + P.r22 * czz;
  // S = C_B + T R^T (upper triangle)
  const float sxx = bxx + (t00 * P.r00 + t01 * P.r01 + t02 * P.r02);
  const float sxy = bxy + (t00 * P.r10 + t01 * P.r11 + t02 * P.r12);
  const float sxz = bxz + (t00 * P.r20 + t01 * P.r21 + t02 * P.r22);
  const float syy = byy + (t10 * P.r10 + t11 * P.r11 + t12 * P.r12);
  const float syz = byz + (t10 * P.r20 + t11 * P.r21 + t12 * P.r22);
  const float szz = bzz + (t20 * P.r20 + t21 * P.r21 + t22 * P.r22);
  // inverse by cofactors
  const float c00 = syy * szz - syz * syz;
  const float c01 = sxz * syz - sxy * szz;
  const float c02 = sxy * syz - sxz * syy;
  const float c11 = sxx * szz - sxz * sxz;
  const float c12 = sxy * sxz - sxx * syz;
  const float c22 = sxx * syy - sxy * sxy;
  const float det = sxx * c00 + sxy * c01 + sxz * c02;
  const float id = __fdividef(1.0f, det);
  mxx = c00 * id; mxy = c01 * id; mxz = c02 * id; myy = c11 * id; myz = c12 * id; mzz = c22 * id;
}

// Transposing warp reduction: on return lane l holds sum over the warp of v[l] (in v[0]).
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int j = 0; j < step; j++) {
      const float send = upper ? v[j] : v[j + step];
      const float keep = upper ? v[j + step] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}

// fp64 epilogue of one factor, executed by warp 0 of the CTA that retired the factor's last tile.
// slab row element e (see GB_SLAB_STRIDE in include/glim_b200.h) -> index in the 122-double record
__device__ __forceinline__ int slab_to_record(int e) {
  if (e < 21 || (e >= 57 && e < 78)) {  // upper triangles of H_tt / H_ss, row-major (i <= j)
    int u = e < 21 ? e : e - 57, i = 0;
    while (u >= 6 - i) { u -= 6 - i; i++; }
    const int j = i + u;
    return (e < 21 ? 0 : 36) + j * 6 + i;
  }
  if (e < 57) return 72 + (e - 21);   // H_ts, column-major as in the record
  if (e < 84) return 108 + (e - 78);  // b_t
  if (e < 90) return 114 + (e - 84);  // b_s
  return 120 + (e - 90);              // error, num_inliers
}

// Fused result exchange: executed by the warp that completed factor f.  If f was the LAST factor of its pair, sum the
// pair's records (fp64, fixed order -> deterministic), and store the fp32 row into every rank's slab (128-bit stores;
// peers are reached through their IPC-mapped addresses over NVLink).
__device__ __noinline__ void pair_push(int f, const FactorDesc& D, const double* __restrict__ out, const PeerPush* __restrict__ peer_tab, float* row /* 96 floats of shared memory */) {
  const PeerPush& peer = *peer_tab;
  const int lane = threadIdx.x & 31;
  __threadfence();  // this factor's record is visible before the ticket
  __syncwarp();
  int last = 0;
  const int pb = peer.pair_ptr[D.pair], pe = peer.pair_ptr[D.pair + 1];
  if (lane == 0) {
    const unsigned t = atomicAdd(&peer.pair_done[D.pair], 1u);
    last = (t == (unsigned)(pe - pb) - 1u);
    if (last) peer.pair_done[D.pair] = 0u;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  for (int e = lane; e < GB_SLAB_STRIDE; e += 32) {
    double s = 0.0;
    if (e < 92) {
      const int r = slab_to_record(e);
      for (int k = pb; k < pe; k++) s += __ldcg(&out[(size_t)peer.pair_factors[k] * GB_OUT_DOUBLES + r]);
    }
    row[e] = (float)s;
  }
  __syncwarp();
  if (lane < GB_SLAB_STRIDE / 4) {
    const float4 v = reinterpret_cast<const float4*>(row)[lane];
    for (int p = 0; p < peer.world; p++) reinterpret_cast<float4*>(peer.base[p] + (size_t)D.pair * GB_SLAB_STRIDE)[lane] = v;
  }
  (void)f;
}

__device__ void factor_epilogue(int f, const FactorDesc& D, const double* __restrict__ poses, double* __restrict__ accum, int acc_slots, double* __restrict__ out, float* __restrict__ slab, double* sm /* >= 36+36+36+32 doubles */) {
  const int lane = threadIdx.x & 31;
  double* A = sm;            // 32 accumulators
  double* H = sm + 32;       // 6x6 H_tt, row-major (symmetric)
  double* Ad = sm + 68;      // 6x6 adjoint, row-major
  double* X = sm + 104;      // H_tt * Ad, row-major
  // the accumulators were produced by L2 atomics of other CTAs: read them past L1
  // (a factor's accumulator is replicated over acc_slots copies so that the items of a sweep with FEW factors do not all
  //  serialise on the same 29 addresses in L2; summed here in slot order)
  if (lane < 29) {
    double a = 0.0;
    for (int sl = 0; sl < acc_slots; sl++) {
      double* slot = &accum[((size_t)f * acc_slots + sl) * GB_ACC_STRIDE + lane];
      a += __ldcg(slot);
      *slot = 0.0;  // re-zero for the next sweep (self-cleaning; nobody touches this factor again in this launch)
    }
    A[lane] = a;
  }
  __syncwarp();
  // unpack upper triangle
  if (lane == 0) {
    int k = 0;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) { H[i * 6 + j] = A[k]; H[j * 6 + i] = A[k]; k++; }
    // Ad = [[R, 0], [hat(t) R, R]] from the fp32-cast pose the kernel used
    const double* T = poses + (size_t)f * 16;
    double R[9], t[3];
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) R[r * 3 + c] = (double)(float)T[c * 4 + r]; t[r] = (double)(float)T[12 + r]; }
    const double ht[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
    for (int i = 0; i < 36; i++) Ad[i] = 0.0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        Ad[i * 6 + j] = R[i * 3 + j];
        Ad[(i + 3) * 6 + (j + 3)] = R[i * 3 + j];
        Ad[(i + 3) * 6 + j] = ht[i * 3 + 0] * R[0 * 3 + j] + ht[i * 3 + 1] * R[1 * 3 + j] + ht[i * 3 + 2] * R[2 * 3 + j];
      }
  }
  __syncwarp();
  for (int e = lane; e < 36; e += 32) {
    const int i = e / 6, j = e % 6;
    double s = 0;
    for (int k = 0; k < 6; k++) s += H[i * 6 + k] * Ad[k * 6 + j];
    X[e] = s;
  }
  __syncwarp();
  double* o = out + (size_t)f * GB_OUT_DOUBLES;
  float* srow = slab ? slab + (size_t)D.pair * GB_SLAB_STRIDE : nullptr;
  for (int e = lane; e < 36; e += 32) {
    const int i = e / 6, j = e % 6;  // output element (row i, col j), stored column-major
    double ss = 0;
    for (int k = 0; k < 6; k++) ss += Ad[k * 6 + i] * X[k * 6 + j];
    o[j * 6 + i] = H[i * 6 + j];             // H_tt
    o[36 + j * 6 + i] = ss;                  // H_ss = Ad^T H_tt Ad
    o[72 + j * 6 + i] = -X[i * 6 + j];       // H_ts = -H_tt Ad
    if (srow) {
      atomicAdd(&srow[21 + j * 6 + i], (float)(-X[i * 6 + j]));
      if (j >= i) {
        const int u = i * 6 - (i * (i - 1)) / 2 + (j - i);  // index in the row-major upper triangle
        atomicAdd(&srow[u], (float)H[i * 6 + j]);
        atomicAdd(&srow[57 + u], (float)ss);
      }
    }
  }
  if (lane < 6) {
    double bs = 0;
    for (int k = 0; k < 6; k++) bs += Ad[k * 6 + lane] * A[21 + k];
    o[108 + lane] = A[21 + lane];
    o[114 + lane] = -bs;
    if (srow) { atomicAdd(&srow[78 + lane], (float)A[21 + lane]); atomicAdd(&srow[84 + lane], (float)(-bs)); }
  }
  if (lane == 6) { o[120] = A[27]; if (srow) atomicAdd(&srow[90], (float)A[27]); }
  if (lane == 7) { o[121] = A[28]; if (srow) atomicAdd(&srow[91], (float)A[28]); }
}

// ---------------------------------------------------------------------------------------------
// The sweep kernel.  Persistent grid of independent WARPS: every warp starts on the item with its own
// index and then draws further work items from a global queue (one atomic per item; the counter is
// monotonic across launches so it never needs a reset; sweeps with no more items than warps use no
// queue at all).  An item = up to `chunk` consecutive source points of one factor, processed in rounds of
// kSubMax points with two warp-local phases:
//   A (lookup, all lanes busy): 16-byte read of (x, y, z, c00), transform, voxel coordinate, hash
//     probe (4 points per lane in flight); hits are COMPACTED into the warp's shared-memory queue as (point, voxel) pairs with
//     ballot + popc, in point order (deterministic).
//   B (derivatives, dense): lanes walk the queue, so the warp stays full whatever the inlier rate;
//     only hits pay for the remaining 20 bytes of the source point, the 48-byte voxel record and
//     the ~180-instruction Mahalanobis / Hessian update.
// This is the reference's lookup-pass / compaction / derivative-pass structure, but the inlier list
// lives in shared memory for the lifetime of one round instead of making a round trip through HBM.
// There is no block-level synchronisation anywhere: the item ends with a transposing warp
// reduce-scatter (31 shuffles) and 29 fp64 atomics into the factor's accumulator; the warp that
// retires a factor's last item runs its fp64 epilogue.
// ---------------------------------------------------------------------------------------------
constexpr int kSubMax = 512;   // queue capacity per warp (points per round)
constexpr int kLookupUnroll = 4;

__device__ __forceinline__ PoseF pose_from_colmajor(const double* __restrict__ T) {
  PoseF P;
  P.r00 = (float)T[0]; P.r01 = (float)T[4]; P.r02 = (float)T[8];  P.tx = (float)T[12];
  P.r10 = (float)T[1]; P.r11 = (float)T[5]; P.r12 = (float)T[9];  P.ty = (float)T[13];
  P.r20 = (float)T[2]; P.r21 = (float)T[6]; P.r22 = (float)T[10]; P.tz = (float)T[14];
  return P;
}

template <int MODE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_vgicp_sweep(
  const FactorDesc* __restrict__ descs, const double* __restrict__ poses, const double* __restrict__ poses_eval,
  const int2* __restrict__ items, int num_items, int chunk, int static_first,
  unsigned long long* __restrict__ item_ctr, unsigned long long ctr_base,
  double* __restrict__ accum, int acc_slots, unsigned* __restrict__ done, double* __restrict__ out, float* __restrict__ slab, const PeerPush* __restrict__ peer) {
  __shared__ __align__(16) uint2 s_q[kWarps][kSubMax];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint2* __restrict__ q = s_q[warp];

  // first item: static (warp id) -- thousands of warps hammering one atomic at kernel start cost ~10 us, which is most of
  // a small odometry sweep; further items (only when there are more items than warps) come from the global queue
  const int total_warps = static_first ? gridDim.x * kWarps : 0;
  const bool dynamic = num_items > total_warps;
  int item = blockIdx.x * kWarps + warp;
  if (!static_first) {
    if (lane == 0) item = (int)(atomicAdd(item_ctr, 1ull) - ctr_base);
    item = __shfl_sync(0xffffffffu, item, 0);
  }

  while (item < num_items) {
    int next_item = 0x7fffffff;
    if (dynamic && lane == 0) next_item = (int)(atomicAdd(item_ctr, 1ull) - ctr_base) + total_warps;  // latency hidden behind the item
    const int2 it = __ldg(&items[item]);
    const int f = it.x;
    const FactorDesc D = descs[f];
    // pose -> fp32 row-major R | t  (Isometry3f cast of the reference GPU factor)
    const PoseF P = pose_from_colmajor(poses + (size_t)f * 16);
    PoseF Pe = P;
    if (MODE == GB_MODE_ERROR) Pe = pose_from_colmajor(poses_eval + (size_t)f * 16);
    const int item_end = min(it.y + chunk, D.n);

    float acc[32];
#pragma unroll
    for (int k = 0; k < 32; k++) acc[k] = 0.f;

    for (int wb = it.y; wb < item_end; wb += kSubMax) {
      const int we = min(wb + kSubMax, item_end);
      // ---------------- phase A: lookup + compaction ----------------
      int nq = 0;  // warp-uniform queue length
      for (int i0 = wb; i0 < we; i0 += 32 * kLookupUnroll) {
        int cx[kLookupUnroll], cy[kLookupUnroll], cz[kLookupUnroll];
        uint32_t h[kLookupUnroll];
        int4 b[kLookupUnroll], b1[kLookupUnroll];
#pragma unroll
        for (int u = 0; u < kLookupUnroll; u++) {
          const int i = i0 + u * 32 + lane;
          const float4 a0 = __ldg(&D.p0[min(i, we - 1)]);
          float qx, qy, qz;
          transform(P, a0.x, a0.y, a0.z, qx, qy, qz);
          cx[u] = gb_coord(qx, D.inv_res); cy[u] = gb_coord(qy, D.inv_res); cz[u] = gb_coord(qz, D.inv_res);
          h[u] = gb_hash(cx[u], cy[u], cz[u]);
          // the first two probe slots travel together (adjacent 16-byte buckets, one round trip): measured +5 % over
          // fetching the second slot on demand even with tables at load <= 1/8
          b[u] = __ldg(&D.buckets[h[u] & D.mask]);
          b1[u] = __ldg(&D.buckets[(h[u] + 1u) & D.mask]);
        }
#pragma unroll
        for (int u = 0; u < kLookupUnroll; u++) {
          const int i = i0 + u * 32 + lane;
          int v = -1;
          if (b[u].w >= 0) {
            if (b[u].x == cx[u] && b[u].y == cy[u] && b[u].z == cz[u]) {
              v = b[u].w;
            } else if (D.max_scan > 1 && b1[u].w >= 0) {
              if (b1[u].x == cx[u] && b1[u].y == cy[u] && b1[u].z == cz[u]) {
                v = b1[u].w;
              } else {
                for (int k = 2; k < D.max_scan; k++) {  // rare: longer collision chain
                  const int4 bb = __ldg(&D.buckets[(h[u] + (uint32_t)k) & D.mask]);
                  if (bb.w < 0) break;
                  if (bb.x == cx[u] && bb.y == cy[u] && bb.z == cz[u]) { v = bb.w; break; }
                }
              }
            }
          }
          if (i >= we) v = -1;
          const unsigned m = __ballot_sync(0xffffffffu, v >= 0);
          if (v >= 0) q[nq + __popc(m & lt_mask)] = make_uint2((unsigned)i, (unsigned)v);
          nq += __popc(m);
        }
      }
      __syncwarp();

      // ---------------- phase B: dense derivative pass over the round's inliers ----------------
#pragma unroll 2
      for (int k = lane; k < nq; k += 32) {
        const uint2 e = q[k];
        const int i = (int)e.x;
        const float4 a0 = __ldg(&D.p0[i]);
        const float4 a1 = __ldg(&D.p1[i]);
        const float a2 = __ldg(&D.p2[i]);
        const float4 v0 = __ldg(&D.voxels[3 * (size_t)e.y + 0]);
        const float4 v1 = __ldg(&D.voxels[3 * (size_t)e.y + 1]);
        const float4 v2 = __ldg(&D.voxels[3 * (size_t)e.y + 2]);
        float qx, qy, qz;
        transform(Pe, a0.x, a0.y, a0.z, qx, qy, qz);
        float mxx, mxy, mxz, myy, myz, mzz;
        fused_mahalanobis(Pe, a0.w, a1.x, a1.y, a1.z, a1.w, a2, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, mxx, mxy, mxz, myy, myz, mzz);
        const float rx = v0.x - qx, ry = v0.y - qy, rz = v0.z - qz;
        const float wx = mxx * rx + mxy * ry + mxz * rz;
        const float wy = mxy * rx + myy * ry + myz * rz;
        const float wz = mxz * rx + myz * ry + mzz * rz;
        acc[27] += rx * wx + ry * wy + rz * wz;
        acc[28] += 1.0f;
        if (MODE == GB_MODE_LINEARIZE) {
          // G = hat(q) M   (rows: rotation, cols: translation block of H_tt)
          const float g00 = qy * mxz - qz * mxy, g01 = qy * myz - qz * myy, g02 = qy * mzz - qz * myz;
          const float g10 = qz * mxx - qx * mxz, g11 = qz * mxy - qx * myz, g12 = qz * mxz - qx * mzz;
          const float g20 = qx * mxy - qy * mxx, g21 = qx * myy - qy * mxy, g22 = qx * myz - qy * mxz;
          // H_rr = G hat(q)^T : row i = q x g_i   (upper triangle)
          acc[0] += qy * g02 - qz * g01;
          acc[1] += qz * g00 - qx * g02;
          acc[2] += qx * g01 - qy * g00;
          acc[3] += g00; acc[4] += g01; acc[5] += g02;
          acc[6] += qz * g10 - qx * g12;
          acc[7] += qx * g11 - qy * g10;
          acc[8] += g10; acc[9] += g11; acc[10] += g12;
          acc[11] += qx * g21 - qy * g20;
          acc[12] += g20; acc[13] += g21; acc[14] += g22;
          acc[15] += mxx; acc[16] += mxy; acc[17] += mxz; acc[18] += myy; acc[19] += myz; acc[20] += mzz;
          // b_t = [q x w ; w]
          acc[21] += qy * wz - qz * wy;
          acc[22] += qz * wx - qx * wz;
          acc[23] += qx * wy - qy * wx;
          acc[24] += wx; acc[25] += wy; acc[26] += wz;
        }
      }
      __syncwarp();  // the queue is overwritten by the next round
    }

    // ---------------- item reduction: warp -> 29 fp64 atomics ----------------
    double* __restrict__ my_acc = accum + ((size_t)f * acc_slots + (size_t)(item & (acc_slots - 1))) * GB_ACC_STRIDE;
    if (MODE == GB_MODE_LINEARIZE) {
      const float r = warp_reduce_scatter32(acc, lane);
      if (lane < 29) atomicAdd(&my_acc[lane], (double)r);
    } else {
      float e = acc[27], n = acc[28];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) { e += __shfl_xor_sync(0xffffffffu, e, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
      if (lane == 27) atomicAdd(&my_acc[27], (double)e);
      if (lane == 28) atomicAdd(&my_acc[28], (double)n);
    }
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      const unsigned ticket = atomicAdd(&done[f], 1u);
      last = (ticket == (unsigned)D.num_tiles - 1u);
      if (last) done[f] = 0u;  // self-cleaning
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      factor_epilogue(f, D, MODE == GB_MODE_ERROR ? poses_eval : poses, accum, acc_slots, out, slab, reinterpret_cast<double*>(q));
      __syncwarp();
      if (MODE == GB_MODE_LINEARIZE && peer != nullptr) pair_push(f, D, out, peer, reinterpret_cast<float*>(q) + 512);
      __syncwarp();
    }
    item = __shfl_sync(0xffffffffu, next_item, 0);
  }
}

// overlap: one thread per source point, count points that hit an occupied voxel of any target
__global__ void __launch_bounds__(256) k_overlap(int num_targets, const FactorDesc* __restrict__ descs, const double* __restrict__ poses, int n, int* __restrict__ count) {
  extern __shared__ float s_poses[];  // num_targets x 12
  for (int k = threadIdx.x; k < num_targets * 12; k += blockDim.x) {
    const int t = k / 12, e = k % 12;
    const int r = e < 9 ? e / 3 : e - 9, c = e < 9 ? e % 3 : 3;
    s_poses[k] = (float)poses[(size_t)t * 16 + c * 4 + r];
  }
  __syncthreads();
  int local = 0;
  const FactorDesc D0 = descs[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 a0 = __ldg(&D0.p0[i]);
    for (int t = 0; t < num_targets; t++) {
      const PoseF P = load_pose(s_poses + 12 * t);
      float qx, qy, qz;
      transform(P, a0.x, a0.y, a0.z, qx, qy, qz);
      const float inv_res = descs[t].inv_res;
      const int v = gb_lookup(descs[t].buckets, descs[t].mask, descs[t].max_scan, gb_coord(qx, inv_res), gb_coord(qy, inv_res), gb_coord(qz, inv_res));
      if (v >= 0) { local++; break; }
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

}  // namespace

template <int MODE, int MINB>
static void launch_variant(gb_sweep* s, const double* poses_eval, float* slab) {
  // the table for the buffer of the current step parity (both were written to the device when the slab was attached)
  const PeerPush* pp = (MODE == GB_MODE_LINEARIZE && s->peer) ? s->d_peer_tables + s->peer->parity : nullptr;
  k_vgicp_sweep<MODE, MINB><<<s->grid, kThreads, 0, s->ctx->stream>>>(s->d_descs, s->d_poses, poses_eval, s->d_tiles, s->num_tiles, s->tile_size, s->static_first, s->d_tile_ctr, s->ctr_base, s->d_accum, s->acc_slots, s->d_done, s->d_out, slab, pp);
}

// completion flags of the fused exchange: one thread per rank publishes this rank's step to that peer, then waits for the
// peer's flag.  The preceding sweep kernel has completed (stream order), so all its peer stores have been performed.
struct PeerFlags { unsigned* flags[GB_MAX_PEERS]; };
__global__ void k_peer_signal_wait(PeerFlags pf, int world, int rank, unsigned step, int* timeout) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();
  volatile unsigned* remote = pf.flags[t] + rank;
  *remote = step;
  __threadfence_system();
  volatile unsigned* mine = pf.flags[rank] + t;
  const long long t0 = clock64();
  while ((int)(*mine - step) < 0) {
    __nanosleep(200);
    if (clock64() - t0 > 20000000000ll) { *timeout = 1; break; }  // ~10 s: a peer died; do not hang the GPU
  }
  __threadfence_system();
}

gb_status gb_launch_sweep(gb_sweep* s, int mode) {
  if (s->num_tiles == 0) return GB_OK;
  gb_ctx* ctx = s->ctx;
  // register-budget variant: 2 CTAs/SM (124 regs, no spills), 3 (80 regs), 4 (64 regs); chosen when the sweep is created
  const int v = s->min_blocks;
  if (mode == GB_MODE_LINEARIZE) {
    if (v >= 4) launch_variant<GB_MODE_LINEARIZE, 4>(s, nullptr, s->d_slab);
    else if (v == 3) launch_variant<GB_MODE_LINEARIZE, 3>(s, nullptr, s->d_slab);
    else launch_variant<GB_MODE_LINEARIZE, 2>(s, nullptr, s->d_slab);
  } else {
    if (v >= 4) launch_variant<GB_MODE_ERROR, 4>(s, s->d_poses_eval, nullptr);
    else if (v == 3) launch_variant<GB_MODE_ERROR, 3>(s, s->d_poses_eval, nullptr);
    else launch_variant<GB_MODE_ERROR, 2>(s, s->d_poses_eval, nullptr);
  }
  GB_CUDA(cudaGetLastError());
  // every processed item draws exactly one ticket from the queue (when the queue is in use at all)
  if (!s->static_first) s->ctr_base += (unsigned long long)s->num_tiles + (unsigned long long)s->grid * kWarps;
  else if (s->num_tiles > s->grid * kWarps) s->ctr_base += (unsigned long long)s->num_tiles;
  ctx->launches++;
  return GB_OK;
}

gb_status gb_launch_peer_signal_wait(gb_peer_slab* ps) {
  PeerFlags pf;
  memset(&pf, 0, sizeof(pf));
  for (int p = 0; p < ps->world; p++) pf.flags[p] = reinterpret_cast<unsigned*>(ps->peer[p] + 2 * ps->buf_floats * sizeof(float));
  k_peer_signal_wait<<<1, 32, 0, ps->ctx->stream>>>(pf, ps->world, ps->rank, ps->step, ps->d_timeout);
  GB_CUDA(cudaGetLastError());
  ps->ctx->launches++;
  return GB_OK;
}

gb_status gb_launch_overlap(gb_ctx* ctx, int num_targets, const FactorDesc* d_descs, const double* d_poses, int n, int* d_count) {
  if (n <= 0 || num_targets <= 0) return GB_OK;
  const int grid = min((n + 255) / 256, ctx->num_sms * 8);
  k_overlap<<<grid, 256, (size_t)num_targets * 12 * sizeof(float), ctx->stream>>>(num_targets, d_descs, d_poses, n, d_count);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}
