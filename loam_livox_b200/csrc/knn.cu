// Map index ("bucket tree") build and the exact 5-NN search + residual-block construction kernels.
//
// Replaces, on the reference side:
//   pcl::KdTreeFLANN::setInputCloud            /root/reference/source/laser_mapping.hpp:544-545,
//                                              /root/reference/source/point_cloud_registration.hpp:596-597   (K5)
//   pointAssociateToMap + nearestKSearch(k=5)  /root/reference/source/point_cloud_registration.hpp:247-249,349-351,622-661 (K6)
//   gates + functor constructors               /root/reference/source/point_cloud_registration.hpp:254-331,353-431,
//                                              /root/reference/source/ceres_icp.hpp:246-260,314-336          (K7)
//
// Compiled with -fmad=false: the float distance (FLANN L2_Simple<float>: ((dx*dx)+dy*dy)+dz*dz), the fp64
// transform and the fp64 line/plane geometry must round exactly like the scalar CPU code.
#include <cub/cub.cuh>
#include "common.cuh"
#include "kernels.cuh"
#include "exact_math.cuh"

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------ build
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// bbox[0..2] = min (ordered-int encoding), bbox[3..5] = max, bbox[6] = number of finite points
__global__ void bbox_kernel(const float4* __restrict__ src, int n, int* __restrict__ bbox) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  int cnt = 0;
  for (; i < n; i += gridDim.x * blockDim.x) {
    float4 p = src[i];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
      hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
      cnt++;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; k++) { lo[k] = fminf(lo[k], __shfl_xor_sync(FULL, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(FULL, hi[k], o)); }
    cnt += __shfl_xor_sync(FULL, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { atomicMin(&bbox[k], f2ord(lo[k])); atomicMax(&bbox[3 + k], f2ord(hi[k])); }
    atomicAdd(&bbox[6], cnt);
  }
}
__global__ void bbox_init_kernel(int* bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = f2ord(INFINITY);
  else if (threadIdx.x < 6) bbox[threadIdx.x] = f2ord(-INFINITY);
  else if (threadIdx.x == 6) bbox[6] = 0;
}

__device__ __forceinline__ unsigned long long spread21(unsigned v) {
  unsigned long long x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
__global__ void morton_kernel(const float4* __restrict__ src, int n, const int* __restrict__ bbox, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = src[i];
  unsigned long long key = ~0ull;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    float lo[3] = {ord2f(bbox[0]), ord2f(bbox[1]), ord2f(bbox[2])}, hi[3] = {ord2f(bbox[3]), ord2f(bbox[4]), ord2f(bbox[5])};
    float c[3] = {p.x, p.y, p.z}; unsigned q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float ext = hi[k] - lo[k];
      float u = ext > 0.f ? (c[k] - lo[k]) / ext : 0.f;
      u = fminf(fmaxf(u, 0.f), 1.f);
      q[k] = (unsigned)fminf(u * 2097152.0f, 2097151.0f);
    }
    key = spread21(q[0]) | (spread21(q[1]) << 1) | (spread21(q[2]) << 2);
  }
  keys[i] = key; vals[i] = i;
}
__global__ void gather_kernel(const float4* __restrict__ src, const int* __restrict__ order, int n_valid, int n_pad, float4* __restrict__ pts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  if (i < n_valid) { int j = order[i]; float4 p = src[j]; p.w = __int_as_float(j); pts[i] = p; }
  else pts[i] = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7fffffff));
}
// One warp per box: level 0 boxes bound 32 points, higher levels bound 32 child boxes. Pads are neutral (+inf / -inf).
__global__ void box_kernel(const float4* __restrict__ child_lo, const float4* __restrict__ child_hi, int n_child, int n_box_pad, int n_box, float4* __restrict__ lo, float4* __restrict__ hi) {
  int box = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (box >= n_box_pad) return;
  float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
  int c = box * 32 + lane;
  if (box < n_box && c < n_child) {
    float4 a = child_lo[c]; float4 b = child_hi ? child_hi[c] : a;
    l[0] = a.x; l[1] = a.y; l[2] = a.z; h[0] = b.x; h[1] = b.y; h[2] = b.z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 3; k++) { l[k] = fminf(l[k], __shfl_xor_sync(FULL, l[k], o)); h[k] = fmaxf(h[k], __shfl_xor_sync(FULL, h[k], o)); }
  if (lane == 0) { lo[box] = make_float4(l[0], l[1], l[2], 0.f); hi[box] = make_float4(h[0], h[1], h[2], 0.f); }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int build_bucket_tree(ll_ctx* ctx, const float4* d_src, int n_src, BucketTree* t) {
  *t = BucketTree();
  t->n_src = n_src;
  cudaStream_t s = ctx->stream;
  // scratch: bbox(8 ints) | keys | keys_out | vals | vals_out | cub temp
  size_t temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, n_src > 0 ? n_src : 1, 0, 63, s);
  size_t off_bbox = 0, off_k0 = align256(64), off_k1 = off_k0 + align256((size_t)n_src * 8), off_v0 = off_k1 + align256((size_t)n_src * 8),
         off_v1 = off_v0 + align256((size_t)n_src * 4), off_tmp = off_v1 + align256((size_t)n_src * 4);
  LL_CUDA(ctx, ctx->scratch.reserve(off_tmp + temp_bytes + 256));
  char* base = ctx->scratch.as<char>();
  int* bbox = (int*)(base + off_bbox);
  unsigned long long* k0 = (unsigned long long*)(base + off_k0); unsigned long long* k1 = (unsigned long long*)(base + off_k1);
  int* v0 = (int*)(base + off_v0); int* v1 = (int*)(base + off_v1);
  int n_valid = 0;
  bbox_init_kernel<<<1, 32, 0, s>>>(bbox); ctx->launches++;
  if (n_src > 0) {
    int grid = min(ll_div_up(n_src, 256), ctx->num_sms * 8);
    bbox_kernel<<<grid, 256, 0, s>>>(d_src, n_src, bbox); ctx->launches++;
    morton_kernel<<<ll_div_up(n_src, 256), 256, 0, s>>>(d_src, n_src, bbox, k0, v0); ctx->launches++;
    LL_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + off_tmp, temp_bytes, k0, k1, v0, v1, n_src, 0, 63, s)); ctx->launches += 8;
    LL_CUDA(ctx, cudaMemcpyAsync(&n_valid, bbox + 6, sizeof(int), cudaMemcpyDeviceToHost, s));
    LL_CUDA(ctx, cudaStreamSynchronize(s));
  }
  t->n = n_valid; t->n_pad = ll_div_up(n_valid > 0 ? n_valid : 1, 32) * 32;
  // level sizes
  int cnt = t->n_pad / 32; t->n_levels = 0;
  for (;;) { t->level_count[t->n_levels++] = cnt; if (cnt <= 32 || t->n_levels == LL_MAX_LEVELS) break; cnt = ll_div_up(cnt, 32); }
  if (t->level_count[t->n_levels - 1] > 32) { ctx->set_error("map too large for LL_MAX_LEVELS"); return LL_ERR_CAPACITY; }
  size_t bytes = align256((size_t)t->n_pad * 16) + align256((size_t)n_src * 16);
  for (int l = 0; l < t->n_levels; l++) bytes += 2 * align256((size_t)ll_div_up(t->level_count[l], 32) * 32 * 16);
  LL_CUDA(ctx, t->storage.reserve(bytes));
  char* p = t->storage.as<char>();
  t->pts = (float4*)p; p += align256((size_t)t->n_pad * 16);
  for (int l = 0; l < t->n_levels; l++) { size_t b = align256((size_t)ll_div_up(t->level_count[l], 32) * 32 * 16); t->lo[l] = (float4*)p; p += b; t->hi[l] = (float4*)p; p += b; }
  t->src = (float4*)p;
  if (n_src > 0) LL_CUDA(ctx, cudaMemcpyAsync(t->src, d_src, (size_t)n_src * 16, cudaMemcpyDeviceToDevice, s));
  gather_kernel<<<ll_div_up(t->n_pad, 256), 256, 0, s>>>(d_src, v1, n_valid, t->n_pad, t->pts); ctx->launches++;
  for (int l = 0; l < t->n_levels; l++) {
    int n_box = t->level_count[l], n_box_pad = ll_div_up(n_box, 32) * 32;
    const float4* clo = l == 0 ? t->pts : t->lo[l - 1]; const float4* chi = l == 0 ? nullptr : t->hi[l - 1];
    int n_child = l == 0 ? n_valid : t->level_count[l - 1];
    box_kernel<<<ll_div_up(n_box_pad * 32, 256), 256, 0, s>>>(clo, chi, n_child, n_box_pad, n_box, t->lo[l], t->hi[l]); ctx->launches++;
  }
  LL_CUDA(ctx, cudaGetLastError());
  return LL_OK;
}

TreeView make_view(const BucketTree& t) {
  TreeView v; v.pts = t.pts; v.n = t.n; v.n_levels = t.n_levels;
  for (int l = 0; l < LL_MAX_LEVELS; l++) { v.lo[l] = t.lo[l]; v.hi[l] = t.hi[l]; }
  return v;
}

// ------------------------------------------------------------------------------------------------ search
struct Top5 { float d[LL_KNN]; int id[LL_KNN]; int pos[LL_KNN]; };

__device__ __forceinline__ bool lex_less(float d, int id, float d2, int id2) { return d < d2 || (d == d2 && id < id2); }

__device__ __forceinline__ void top5_insert(Top5& t, float d, int id, int pos) {
  t.d[4] = d; t.id[4] = id; t.pos[4] = pos;
#pragma unroll
  for (int j = 4; j > 0; --j) {
    if (lex_less(t.d[j], t.id[j], t.d[j - 1], t.id[j - 1])) {
      float td = t.d[j]; t.d[j] = t.d[j - 1]; t.d[j - 1] = td;
      int ti = t.id[j]; t.id[j] = t.id[j - 1]; t.id[j - 1] = ti;
      int tp = t.pos[j]; t.pos[j] = t.pos[j - 1]; t.pos[j - 1] = tp;
    }
  }
}

// FLANN L2_Simple<float>: result += diff*diff, x then y then z, float accumulation, no contraction.
__device__ __forceinline__ float dist2_exact(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// Lower bound of dist2_exact over every point inside the box: same operation sequence on the per-axis gaps, and
// round-to-nearest is monotone, so lb <= d2 holds bit-wise (the search stays exact).
__device__ __forceinline__ float box_lb(const float4& lo, const float4& hi, float qx, float qy, float qz) {
  float ex = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.f);
  float ey = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.f);
  float ez = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.f);
  return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
}

__device__ __forceinline__ void scan_bucket(const TreeView& tv, int leaf, float qx, float qy, float qz, Top5& t, int lane) {
  const float4 p = __ldg(&tv.pts[leaf * 32 + lane]);   // 512 B coalesced: the whole bucket in one warp load
  const float d = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
  const int id = __float_as_int(p.w);
  bool c = d < INFINITY && lex_less(d, id, t.d[4], t.id[4]);
  while (__any_sync(FULL, c)) {
    unsigned key = c ? __float_as_uint(d) : 0xffffffffu;
    unsigned mn = __reduce_min_sync(FULL, key);
    bool tie = c && key == mn;
    unsigned idk = tie ? (unsigned)id : 0xffffffffu;
    unsigned mid = __reduce_min_sync(FULL, idk);
    unsigned who = __ballot_sync(FULL, tie && idk == mid);
    int sel = __ffs(who) - 1;
    top5_insert(t, __uint_as_float(mn), (int)mid, leaf * 32 + sel);
    if (lane == sel) c = false;
    c = c && lex_less(d, id, t.d[4], t.id[4]);
  }
}

template <int L>
__device__ __forceinline__ void visit(const TreeView& tv, int group, float qx, float qy, float qz, Top5& t, int lane) {
  const float4 lo = __ldg(&tv.lo[L][group * 32 + lane]);
  const float4 hi = __ldg(&tv.hi[L][group * 32 + lane]);
  const float lb = box_lb(lo, hi, qx, qy, qz);
  bool todo = lb < INFINITY;
  for (;;) {
    bool act = todo && lb <= t.d[4];
    unsigned key = act ? __float_as_uint(lb) : 0xffffffffu;
    unsigned mn = __reduce_min_sync(FULL, key);
    if (mn == 0xffffffffu) break;
    unsigned who = __ballot_sync(FULL, act && key == mn);
    int sel = __ffs(who) - 1;
    if (lane == sel) todo = false;
    int child = group * 32 + sel;
    if constexpr (L == 0) scan_bucket(tv, child, qx, qy, qz, t, lane);
    else visit<L - 1>(tv, child, qx, qy, qz, t, lane);
  }
}

__device__ __forceinline__ void warp_knn5(const TreeView& tv, float qx, float qy, float qz, Top5& t, int lane) {
#pragma unroll
  for (int j = 0; j < LL_KNN; j++) { t.d[j] = INFINITY; t.id[j] = 0x7fffffff; t.pos[j] = -1; }
  if (tv.n <= 0) return;
  switch (tv.n_levels) {
    case 1: visit<0>(tv, 0, qx, qy, qz, t, lane); break;
    case 2: visit<1>(tv, 0, qx, qy, qz, t, lane); break;
    case 3: visit<2>(tv, 0, qx, qy, qz, t, lane); break;
    case 4: visit<3>(tv, 0, qx, qy, qz, t, lane); break;
    case 5: visit<4>(tv, 0, qx, qy, qz, t, lane); break;
    default: visit<5>(tv, 0, qx, qy, qz, t, lane); break;
  }
}

// Parity hook (ll_knn): world-frame queries, one warp each.
__global__ void __launch_bounds__(256) knn_query_kernel(TreeView tv, const float4* __restrict__ q, int nq, int* __restrict__ idx5, float* __restrict__ d5) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nq) return;
  float4 p = __ldg(&q[w]);
  Top5 t; warp_knn5(tv, p.x, p.y, p.z, t, lane);
  if (lane < LL_KNN) {
    float d = t.d[0]; int id = t.id[0];
#pragma unroll
    for (int j = 1; j < LL_KNN; j++) if (lane == j) { d = t.d[j]; id = t.id[j]; }
    idx5[w * LL_KNN + lane] = (id == 0x7fffffff) ? -1 : id; d5[w * LL_KNN + lane] = d;
  }
}

__device__ __forceinline__ unsigned cell_owner(float x, float y, float z, float inv_cell, int world) {
  int ix = (int)floorf(x * inv_cell), iy = (int)floorf(y * inv_cell), iz = (int)floorf(z * inv_cell);
  unsigned h = (unsigned)ix * 73856093u ^ (unsigned)iy * 19349663u ^ (unsigned)iz * 83492791u;
  return h % (unsigned)world;
}

// K6 + K7 fused: one warp per scan feature. Slot i < n_corner is a corner feature, the rest are surface features.
// Writes one residual-block slot per feature: blk_a[slot] = (a.x, a.y, a.z, type) with type 0 invalid / 1 line / 2 plane,
// blk_v[slot*3..] = unit line direction or (un-normalised) plane normal, in fp64.
__global__ void __launch_bounds__(256) knn_blocks_kernel(KnnBlocksArgs a) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int M = a.n_corner + a.n_surf;
  if (w >= M) return;
  const bool is_corner = w < a.n_corner;
  const float4 f = __ldg(&a.feat[w]);
  // pointAssociateToMap (non-deblur branch): p_w = q_curr * p + t_curr in fp64, stored as fp32
  const double* qc = a.pose;      // q_curr w,x,y,z
  const double* tc = a.pose + 4;  // t_curr
  double wx, wy, wz; qrot_d(qc, (double)f.x, (double)f.y, (double)f.z, wx, wy, wz);
  const float qx = (float)(wx + tc[0]), qy = (float)(wy + tc[1]), qz = (float)(wz + tc[2]);
  int type = 0; double ax = 0, ay = 0, az = 0, vx = 0, vy = 0, vz = 0;
  bool finite_in = isfinite(f.x) && isfinite(f.y) && isfinite(f.z);
  bool owned = true;
  if (a.world > 1) owned = cell_owner(qx, qy, qz, a.inv_cell, a.world) == (unsigned)a.rank;
  if (owned && (finite_in || !is_corner)) {
    const TreeView& tv = is_corner ? a.corner : a.surf;
    Top5 t; warp_knn5(tv, qx, qy, qz, t, lane);
    if (a.knn_idx && lane < LL_KNN) {
      float d = t.d[0]; int id = t.id[0];
#pragma unroll
      for (int j = 1; j < LL_KNN; j++) if (lane == j) { d = t.d[j]; id = t.id[j]; }
      a.knn_idx[w * LL_KNN + lane] = (id == 0x7fffffff) ? -1 : id; a.knn_d[w * LL_KNN + lane] = d;
    }
    const bool found5 = t.pos[4] >= 0;
    if (is_corner) {
      if (found5 && (double)t.d[4] < a.max_dis_line) {
        if (a.icp_line) {
          float4 p1 = __ldg(&tv.pts[t.pos[0]]), p2 = __ldg(&tv.pts[t.pos[1]]);
          double d0 = (double)p1.x - (double)p2.x, d1 = (double)p1.y - (double)p2.y, d2 = (double)p1.z - (double)p2.z;
          double dn = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
          if (!(dn < 0.0001)) {
            // ceres_icp_point2line ctor: unit_vec_ab = (b - a) / |b - a|
            double u0 = (double)p2.x - (double)p1.x, u1 = (double)p2.y - (double)p1.y, u2 = (double)p2.z - (double)p1.z;
            double n = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
            vx = u0 / n; vy = u1 / n; vz = u2 / n; ax = p1.x; ay = p1.y; az = p1.z; type = 1;
            if (lane == 0) atomicAdd(a.corner_avail, 1);
          }
        }
      }
    } else {
      if (found5 && (double)t.d[4] < a.max_dis_plane) {
        if (a.icp_plane) {
          float4 pa = __ldg(&tv.pts[t.pos[0]]), pb = __ldg(&tv.pts[t.pos[2]]), pc = __ldg(&tv.pts[t.pos[4]]);
          double b0 = (double)pb.x - (double)pa.x, b1 = (double)pb.y - (double)pa.y, b2 = (double)pb.z - (double)pa.z;
          double nb = sqrt(b0 * b0 + b1 * b1 + b2 * b2); b0 = b0 / nb; b1 = b1 / nb; b2 = b2 / nb;
          double c0 = (double)pc.x - (double)pa.x, c1 = (double)pc.y - (double)pa.y, c2 = (double)pc.z - (double)pa.z;
          double nc = sqrt(c0 * c0 + c1 * c1 + c2 * c2); c0 = c0 / nc; c1 = c1 / nc; c2 = c2 / nc;
          vx = b1 * c2 - b2 * c1; vy = b2 * c0 - b0 * c2; vz = b0 * c1 - b1 * c0;   // NOT re-normalised (ceres_icp.hpp:334)
          ax = pa.x; ay = pa.y; az = pa.z; type = 2;
        }
        if (lane == 0) atomicAdd(a.surf_avail, 1);
      }
    }
  }
  if (lane == 0) {
    a.blk_a[w] = make_float4((float)ax, (float)ay, (float)az, __int_as_float(type));
    a.blk_v[(size_t)w * 3 + 0] = vx; a.blk_v[(size_t)w * 3 + 1] = vy; a.blk_v[(size_t)w * 3 + 2] = vz;
  }
}

int launch_knn_query(ll_ctx* ctx, const BucketTree& t, const float4* d_q, int nq, int* d_idx, float* d_d) {
  if (nq == 0) return LL_OK;
  knn_query_kernel<<<ll_div_up(nq * 32, 256), 256, 0, ctx->stream>>>(make_view(t), d_q, nq, d_idx, d_d); ctx->launches++;
  LL_CUDA(ctx, cudaGetLastError());
  return LL_OK;
}
int launch_knn_blocks(ll_ctx* ctx, const KnnBlocksArgs& a) {
  int M = a.n_corner + a.n_surf;
  if (M == 0) return LL_OK;
  knn_blocks_kernel<<<ll_div_up(M * 32, 256), 256, 0, ctx->stream>>>(a); ctx->launches++;
  LL_CUDA(ctx, cudaGetLastError());
  return LL_OK;
}
