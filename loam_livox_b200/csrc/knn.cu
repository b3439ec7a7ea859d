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
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "kernels.cuh"
#include "exact_math.cuh"

#define FULL 0xffffffffu
#ifndef LL_GROUP
#define LL_GROUP 32
#endif
#define BUCKET LL_GROUP  // points per leaf bucket (one per lane of a query group)
#define FANOUT LL_GROUP  // children per node; a node record holds its 8 children's boxes, child c = rec[2c] (lo.xyz) + rec[2c+1] (hi.xyz): 256 B
#define NODE_F4 (2 * LL_GROUP)

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
// 63-bit Hilbert index (Skilling's transpose form, 21 bits per axis) on an ISOTROPIC grid (one scale for all axes):
// consecutive runs along a Hilbert curve are connected blobs, so fixed-size buckets get tight, nearly cubic boxes.
__global__ void hilbert_kernel(const float4* __restrict__ src, int n, const int* __restrict__ bbox, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = src[i];
  unsigned long long key = ~0ull;
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    const float lo[3] = {ord2f(bbox[0]), ord2f(bbox[1]), ord2f(bbox[2])}, hi[3] = {ord2f(bbox[3]), ord2f(bbox[4]), ord2f(bbox[5])};
    const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    const float c[3] = {p.x, p.y, p.z}; unsigned X[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float u = ext > 0.f ? (c[k] - lo[k]) / ext : 0.f;
      u = fminf(fmaxf(u, 0.f), 1.f);
      X[k] = (unsigned)fminf(u * 2097152.0f, 2097151.0f);
    }
    for (unsigned Q = 1u << 20; Q > 1; Q >>= 1) {
      const unsigned P = Q - 1;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (X[k] & Q) X[0] ^= P;
        else { unsigned t = (X[0] ^ X[k]) & P; X[0] ^= t; X[k] ^= t; }
      }
    }
    X[1] ^= X[0]; X[2] ^= X[1];
    unsigned t = 0;
    for (unsigned Q = 1u << 20; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    key = (spread21(X[0]) << 2) | (spread21(X[1]) << 1) | spread21(X[2]);
  }
  keys[i] = key; vals[i] = i;
}
__global__ void gather_kernel(const float4* __restrict__ src, const int* __restrict__ order, int n_valid, int n_pad, float4* __restrict__ pts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  if (i < n_valid) { int j = order[i]; float4 p = src[j]; p.w = __int_as_float(j); pts[i] = p; }
  else pts[i] = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7fffffff));
}
// One thread per (node, child).  Level 0: the children are buckets of 8 points.  Level l > 0: the children are level l-1 nodes
// (box = union of that node's 8 child boxes).  Missing children get the neutral box (+inf, -inf).
__global__ void node_kernel(const float4* __restrict__ pts, int n_valid, const float4* __restrict__ child_nodes, int n_child, int n_nodes, float4* __restrict__ nodes) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = g / FANOUT, c = g % FANOUT;
  if (j >= n_nodes) return;
  float l0 = INFINITY, l1 = INFINITY, l2 = INFINITY, h0 = -INFINITY, h1 = -INFINITY, h2 = -INFINITY;
  const int ci = j * FANOUT + c;
  if (ci < n_child) {
    if (child_nodes == nullptr) {
      for (int k = 0; k < BUCKET; k++) {
        const int pi = ci * BUCKET + k;
        if (pi < n_valid) { const float4 p = pts[pi]; l0 = fminf(l0, p.x); l1 = fminf(l1, p.y); l2 = fminf(l2, p.z); h0 = fmaxf(h0, p.x); h1 = fmaxf(h1, p.y); h2 = fmaxf(h2, p.z); }
      }
    } else {
      const float4* r = child_nodes + (size_t)ci * NODE_F4;
      for (int k = 0; k < FANOUT; k++) {
        const float4 a = r[2 * k], b = r[2 * k + 1];
        l0 = fminf(l0, a.x); l1 = fminf(l1, a.y); l2 = fminf(l2, a.z); h0 = fmaxf(h0, b.x); h1 = fmaxf(h1, b.y); h2 = fmaxf(h2, b.z);
      }
    }
  }
  float4* o = nodes + (size_t)j * NODE_F4;
  o[2 * c] = make_float4(l0, l1, l2, 0.f); o[2 * c + 1] = make_float4(h0, h1, h2, 0.f);
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int build_bucket_tree(ll_ctx* ctx, const float4* d_src, int n_src, BucketTree* t) {
  { DevBuf keep = t->storage; *t = BucketTree(); t->storage = keep; }   // re-indexing in place (ll_map_rebuild) reuses the allocation
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
    hilbert_kernel<<<ll_div_up(n_src, 256), 256, 0, s>>>(d_src, n_src, bbox, k0, v0); ctx->launches++;
    LL_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + off_tmp, temp_bytes, k0, k1, v0, v1, n_src, 0, 63, s)); ctx->launches += 8;
    int hb[8];
    LL_CUDA(ctx, cudaMemcpyAsync(hb, bbox, 7 * sizeof(int), cudaMemcpyDeviceToHost, s));
    LL_CUDA(ctx, cudaStreamSynchronize(s));
    n_valid = hb[6];
    for (int k = 0; k < 6; k++) { int v = hb[k]; v = v >= 0 ? v : v ^ 0x7fffffff; memcpy(&t->bbox[k], &v, 4); }
  }
  t->n = n_valid; t->n_pad = ll_div_up(n_valid > 0 ? n_valid : 1, BUCKET) * BUCKET;
  // level sizes: level 0 nodes have buckets as children; the top level has exactly one node
  int cnt = ll_div_up(t->n_pad / BUCKET, FANOUT); t->n_levels = 0;
  for (;;) { t->level_count[t->n_levels++] = cnt; if (cnt <= 1) break; if (t->n_levels == LL_MAX_LEVELS) { ctx->set_error("map too large for LL_MAX_LEVELS"); return LL_ERR_CAPACITY; } cnt = ll_div_up(cnt, FANOUT); }
  // node storage is laid out top level first
  size_t node_total = 0; for (int l = 0; l < t->n_levels; l++) node_total += (size_t)t->level_count[l];
  size_t bytes = align256((size_t)t->n_pad * 16) + align256((size_t)n_src * 16) + align256(node_total * NODE_F4 * 16);
  LL_CUDA(ctx, t->storage.reserve(bytes));
  char* p = t->storage.as<char>();
  t->pts = (float4*)p; p += align256((size_t)t->n_pad * 16);
  float4* nodes = (float4*)p; p += align256(node_total * NODE_F4 * 16);
  { size_t off = 0; for (int l = t->n_levels - 1; l >= 0; l--) { t->lo[l] = nodes + off * NODE_F4; off += (size_t)t->level_count[l]; } }
  t->hi[0] = nodes;   // base of the node array (top level first)
  t->src = (float4*)p;
  if (n_src > 0) LL_CUDA(ctx, cudaMemcpyAsync(t->src, d_src, (size_t)n_src * 16, cudaMemcpyDeviceToDevice, s));
  gather_kernel<<<ll_div_up(t->n_pad, 256), 256, 0, s>>>(d_src, v1, n_valid, t->n_pad, t->pts); ctx->launches++;
  for (int l = 0; l < t->n_levels; l++) {
    const int n_child = l == 0 ? t->n_pad / BUCKET : t->level_count[l - 1];
    node_kernel<<<ll_div_up(t->level_count[l] * FANOUT, 128), 128, 0, s>>>(t->pts, n_valid, l == 0 ? nullptr : t->lo[l - 1], n_child, t->level_count[l], t->lo[l]); ctx->launches++;
  }
  LL_CUDA(ctx, cudaGetLastError());
  return LL_OK;
}

TreeView make_view(const BucketTree& t) {
  TreeView v; v.pts = t.pts; v.src = t.src; v.n = t.n; v.n_levels = t.n_levels;
  for (int l = 0; l < LL_MAX_LEVELS; l++) v.nodes[l] = t.lo[l];
  for (int k = 0; k < 6; k++) v.bbox[k] = t.bbox[k];
  return v;
}

// ------------------------------------------------------------------------------------------------ search
// LL_GROUP lanes per query (32: one warp per query).  A step of a query's best-first search pops one item from the group's stack
// (shared memory) and handles it with all lanes at once: a NODE -> lane c tests child c's box (two coalesced 16-B loads per lane,
// 1 KB per group) and the qualifying children are pushed far-to-near in one shot (ballot + popc); a BUCKET -> lane c takes
// point c (one 16-B load, 512 B per group) and the candidates are merged into the group's top-5.  The top-5 and the stack pointer are
// replicated in the lanes.  Exact: a box gives a true lower bound of the fp32 distance and (d2, index) is a total order.
#define GROUP LL_GROUP
#define KNN_THREADS 256
#define GROUPS_PER_CTA (KNN_THREADS / GROUP)
#define STACK_CAP (LL_GROUP == 32 ? 160 : 64)
#define ITEM_BUCKET 0x80000000u

__device__ __forceinline__ bool lex_less(float d, int id, float d2, int id2) { return d < d2 || (d == d2 && id < id2); }

// FLANN L2_Simple<float>: result += diff*diff, x then y then z, float accumulation, no contraction.
__device__ __forceinline__ float dist2_exact(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
// Lower bound of dist2_exact over every point inside the box: same operation sequence on the per-axis gaps, and
// round-to-nearest is monotone, so lb <= d2 holds bit-wise (the search stays exact).
__device__ __forceinline__ float box_lb(float lox, float loy, float loz, float hix, float hiy, float hiz, float qx, float qy, float qz) {
  float ex = fmaxf(fmaxf(__fsub_rn(lox, qx), __fsub_rn(qx, hix)), 0.f);
  float ey = fmaxf(fmaxf(__fsub_rn(loy, qy), __fsub_rn(qy, hiy)), 0.f);
  float ez = fmaxf(fmaxf(__fsub_rn(loz, qz), __fsub_rn(qz, hiz)), 0.f);
  return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
}

struct Top5 { float d[LL_KNN]; int id[LL_KNN]; };

__device__ __forceinline__ void top5_insert(Top5& t, float d, int id) {
  t.d[4] = d; t.id[4] = id;
#pragma unroll
  for (int j = 4; j > 0; --j) {
    if (lex_less(t.d[j], t.id[j], t.d[j - 1], t.id[j - 1])) {
      float td = t.d[j]; t.d[j] = t.d[j - 1]; t.d[j - 1] = td;
      int ti = t.id[j]; t.id[j] = t.id[j - 1]; t.id[j - 1] = ti;
    }
  }
}

struct GroupStack { unsigned item[STACK_CAP]; float lb[STACK_CAP]; };

// Merge this step's candidates (one per lane, flag c) into the group's top-5.  Warp-collective.  A candidate whose index is already
// in the list is ignored, so seeding the list with real points (below) can never create duplicates.
__device__ __forceinline__ void merge_candidates(Top5& t, float d, int id, bool c) {
  c = c && d < INFINITY && lex_less(d, id, t.d[4], t.id[4]) && id != t.id[0] && id != t.id[1] && id != t.id[2] && id != t.id[3];
  while (__any_sync(FULL, c)) {
    float md; int mi;   // group minimum of (d, id) among the remaining candidates
    if (GROUP == 32) {   // whole-warp group: two REDUX instructions (non-negative floats order like their bit patterns)
      const unsigned key = c ? __float_as_uint(d) : 0xffffffffu;
      const unsigned mn = __reduce_min_sync(FULL, key);
      mi = (int)__reduce_min_sync(FULL, (c && key == mn) ? (unsigned)id : 0x7fffffffu);
      md = mn == 0xffffffffu ? INFINITY : __uint_as_float(mn);
    } else {
      md = c ? d : INFINITY; mi = c ? id : 0x7fffffff;
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(FULL, md, o, GROUP); const int oi = __shfl_xor_sync(FULL, mi, o, GROUP);
        if (lex_less(od, oi, md, mi)) { md = od; mi = oi; }
      }
    }
    if (md < INFINITY && lex_less(md, mi, t.d[4], t.id[4])) top5_insert(t, md, mi);
    if (c && id == mi && d == md) c = false;
    c = c && lex_less(d, id, t.d[4], t.id[4]);
  }
}

// All 32 lanes of the warp must call this together (width-GROUP shuffles); `active` is uniform inside a group.
// seed_ids: the query's 5 neighbours of the previous ICP iteration (or null / -1): their distances to the moved query seed the
// list, so the bound is tight from the first step.  Without seeds a greedy walk (child with the smallest farthest-corner distance)
// reaches a bucket next to the query and its points seed the list.
__device__ __forceinline__ void group_knn5(const TreeView& tv, GroupStack& st, bool active, float qx, float qy, float qz, Top5& t, const int* seed_ids) {
  const int gl = threadIdx.x & (GROUP - 1);   // lane inside the group
#pragma unroll
  for (int j = 0; j < LL_KNN; j++) { t.d[j] = INFINITY; t.id[j] = 0x7fffffff; }
  const bool go = active && tv.n > 0;
  // ---- seeds
  int sid = -1;
  if (go && seed_ids && gl < LL_KNN) sid = seed_ids[gl];
  const bool seeded = __shfl_sync(FULL, sid, (threadIdx.x & 31) & ~(GROUP - 1), 32) >= 0;   // group-uniform: lane 0 of the group has a seed
  if (__any_sync(FULL, go && seeded)) {
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sid >= 0) p = __ldg(tv.src + sid);
    merge_candidates(t, sid >= 0 ? dist2_exact(qx, qy, qz, p.x, p.y, p.z) : INFINITY, sid, sid >= 0);
  }
  if (__any_sync(FULL, go && !seeded)) {
    const bool walk = go && !seeded;
    int idx = 0;
    const int max_levels = __reduce_max_sync(FULL, tv.n_levels);   // corner / surface groups of one warp may use different trees
    for (int lv = max_levels - 1; lv >= 0; lv--) {
      float md = INFINITY;
      if (walk && lv < tv.n_levels) {
        const float4* r = tv.nodes[lv] + (size_t)idx * NODE_F4 + 2 * gl; const float4 A = __ldg(r), B = __ldg(r + 1);
        if (A.x <= B.x) {   // a real child (the neutral box has lo = +inf > hi)
          const float ex = fmaxf(fabsf(qx - A.x), fabsf(qx - B.x)), ey = fmaxf(fabsf(qy - A.y), fabsf(qy - B.y)), ez = fmaxf(fabsf(qz - A.z), fabsf(qz - B.z));
          md = ex * ex + ey * ey + ez * ez;
        }
      }
      float mm = md;
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) mm = fminf(mm, __shfl_xor_sync(FULL, mm, o, GROUP));
      const unsigned gmask = (GROUP == 32 ? 0xffffffffu : ((1u << GROUP) - 1u)) << ((threadIdx.x & 31) & ~(GROUP - 1));
      const unsigned who = __ballot_sync(FULL, walk && md == mm && md < INFINITY) & gmask;
      if (walk && who) idx = idx * FANOUT + ((__ffs(who) - 1) & (GROUP - 1));
    }
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (walk) p = __ldg(tv.pts + (size_t)idx * BUCKET + gl);
    merge_candidates(t, walk ? dist2_exact(qx, qy, qz, p.x, p.y, p.z) : INFINITY, __float_as_int(p.w), walk);
  }
  // ---- exact best-first search, pruned by the (already tight) 5th distance
  int sp = 0;
  if (go) { if (gl == 0) { st.item[0] = (unsigned)(tv.n_levels - 1) << 26; st.lb[0] = 0.f; } sp = 1; }
  __syncwarp();
  while (__any_sync(FULL, sp > 0)) {
    // ---- pop (skip items that the shrinking bound has made useless)
    bool have = false; unsigned item = 0;
    while (sp > 0) { sp--; if (st.lb[sp] <= t.d[4]) { item = st.item[sp]; have = true; break; } }
    const bool is_bucket = have && (item & ITEM_BUCKET);
    const bool is_node = have && !is_bucket;
    const int lv = (int)((item >> 26) & 31u);
    const int idx = (int)(is_bucket ? (item & 0x7fffffffu) : (item & 0x03ffffffu));
    // ---- one batch of loads per step: child box (2 x 16 B) or point (16 B)
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
    if (is_node) { const float4* r = tv.nodes[lv] + (size_t)idx * NODE_F4 + 2 * gl; A = __ldg(r); B = __ldg(r + 1); }
    else if (is_bucket) A = __ldg(tv.pts + (size_t)idx * BUCKET + gl);
    __syncwarp();   // stack reads above are complete before anyone pushes below
    if (__any_sync(FULL, is_node)) {
      const float lb = is_node ? box_lb(A.x, A.y, A.z, B.x, B.y, B.z, qx, qy, qz) : INFINITY;
      const bool q = is_node && lb < INFINITY && lb <= t.d[4];
      // push every qualifying child in one shot; the nearest one goes on top of the stack (it is popped next), the others in lane order
      const unsigned wq = __ballot_sync(FULL, q);
      const unsigned gmask = (GROUP == 32 ? 0xffffffffu : ((1u << GROUP) - 1u)) << ((threadIdx.x & 31) & ~(GROUP - 1));
      const unsigned gq = wq & gmask;
      const int nq = __popc(gq);
      float mlb = q ? lb : INFINITY;
#pragma unroll
      for (int o = GROUP / 2; o > 0; o >>= 1) mlb = fminf(mlb, __shfl_xor_sync(FULL, mlb, o, GROUP));
      const unsigned near = __ballot_sync(FULL, q && lb == mlb) & gmask;
      const int near_lane = __ffs(near) - 1;                       // warp lane of the nearest qualifying child (or -1)
      const int me = threadIdx.x & 31;
      if (q) {
        const unsigned below = gq & ((1u << me) - 1u);
        int pos = __popc(below); if (near_lane >= 0 && near_lane < me) pos--;   // rank among the non-nearest
        if (me == near_lane) pos = nq - 1;
        st.item[sp + pos] = (lv == 0 ? ITEM_BUCKET : ((unsigned)(lv - 1) << 26)) | (unsigned)(idx * FANOUT + gl); st.lb[sp + pos] = lb;
      }
      if (is_node) sp += nq;
    }
    if (__any_sync(FULL, is_bucket)) merge_candidates(t, is_bucket ? dist2_exact(qx, qy, qz, A.x, A.y, A.z) : INFINITY, __float_as_int(A.w), is_bucket);
    __syncwarp();   // pushes are visible before the next pop
  }
}

// Parity hook (ll_knn): world-frame queries in caller order.
__global__ void __launch_bounds__(KNN_THREADS) knn_query_kernel(TreeView tv, const float4* __restrict__ q, int nq, int* __restrict__ idx5, float* __restrict__ d5) {
  __shared__ GroupStack stacks[GROUPS_PER_CTA];
  const int g = blockIdx.x * GROUPS_PER_CTA + (threadIdx.x / GROUP), gl = threadIdx.x & (GROUP - 1);
  const bool have = g < nq;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f); if (have) p = __ldg(&q[g]);
  const bool active = have && isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
  Top5 t; group_knn5(tv, stacks[threadIdx.x / GROUP], active, p.x, p.y, p.z, t, nullptr);
  if (have && gl < LL_KNN) {
    float d = t.d[0]; int id = t.id[0];
#pragma unroll
    for (int j = 1; j < LL_KNN; j++) if (gl == j) { d = t.d[j]; id = t.id[j]; }
    idx5[g * LL_KNN + gl] = (id == 0x7fffffff) ? -1 : id; d5[g * LL_KNN + gl] = d;
  }
}


// Spatial sort key of a feature at the current pose: class bit (corner/surface) | 21-bit Hilbert index (7 bits per axis) inside the tree's box.
// Neighbouring queries then sit in the same warp / CTA, walk the same tree nodes and buckets, and hit them in L1.
__global__ void query_key_kernel(KnnBlocksArgs a, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = a.n_corner + a.n_surf;
  if (i >= M) return;
  const bool is_corner = i < a.n_corner;
  const float4 f = __ldg(&a.feat[i]);
  double wx, wy, wz; qrot_d(a.pose, (double)f.x, (double)f.y, (double)f.z, wx, wy, wz);
  const float c[3] = {(float)(wx + a.pose[4]), (float)(wy + a.pose[5]), (float)(wz + a.pose[6])};
  const float* bb = is_corner ? a.corner.bbox : a.surf.bbox;
  const float ext = fmaxf(fmaxf(bb[3] - bb[0], bb[4] - bb[1]), bb[5] - bb[2]);
  unsigned X[3]; bool ok = true;
#pragma unroll
  for (int k = 0; k < 3; k++) { if (!isfinite(c[k])) ok = false; float u = ext > 0.f ? (c[k] - bb[k]) / ext : 0.f; u = fminf(fmaxf(u, 0.f), 1.f); X[k] = (unsigned)fminf(u * 128.0f, 127.0f); }   // 7 bits per axis: 21-bit curve + class bit = 3 radix passes, and tiles of ~0.3 m are fine enough
  for (unsigned Q = 1u << 6; Q > 1; Q >>= 1) {
    const unsigned P = Q - 1;
#pragma unroll
    for (int k = 0; k < 3; k++) { if (X[k] & Q) X[0] ^= P; else { unsigned tt = (X[0] ^ X[k]) & P; X[0] ^= tt; X[k] ^= tt; } }
  }
  X[1] ^= X[0]; X[2] ^= X[1];
  unsigned tt = 0;
  for (unsigned Q = 1u << 6; Q > 1; Q >>= 1) if (X[2] & Q) tt ^= Q - 1;
  X[0] ^= tt; X[1] ^= tt; X[2] ^= tt;
  unsigned h = 0;
#pragma unroll
  for (int b = 6; b >= 0; b--) h = (h << 3) | (((X[0] >> b) & 1u) << 2) | (((X[1] >> b) & 1u) << 1) | ((X[2] >> b) & 1u);
  if (!ok) h = 0x1fffffu;
  keys[i] = (is_corner ? 0u : 0x200000u) | h; vals[i] = i;
}

// Gates + functor constructors (K7) for one feature whose 5 nearest neighbours are in `t`; writes the residual-block slot `w`.
__device__ __forceinline__ void emit_block(const KnnBlocksArgs& a, const TreeView& tv, bool is_corner, bool active, int w, const Top5& t) {
  int type = 0; double ax = 0, ay = 0, az = 0, vx = 0, vy = 0, vz = 0;
  if (active) {
    if (a.seed_ids) {
#pragma unroll
      for (int k = 0; k < LL_KNN; k++) a.seed_ids[(size_t)w * LL_KNN + k] = (t.id[k] == 0x7fffffff) ? -1 : t.id[k];
    }
    if (a.knn_d) {
#pragma unroll
      for (int k = 0; k < LL_KNN; k++) a.knn_d[w * LL_KNN + k] = t.d[k];
    }
    const bool found5 = t.id[4] != 0x7fffffff;
    if (is_corner) {
      if (found5 && (double)t.d[4] < a.max_dis_line) {
        if (a.icp_line) {
          const float4 p1 = __ldg(&tv.src[t.id[0]]), p2 = __ldg(&tv.src[t.id[1]]);
          double d0 = (double)p1.x - (double)p2.x, d1 = (double)p1.y - (double)p2.y, d2 = (double)p1.z - (double)p2.z;
          double dn = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
          if (!(dn < 0.0001)) {
            // ceres_icp_point2line ctor: unit_vec_ab = (b - a) / |b - a|
            double u0 = (double)p2.x - (double)p1.x, u1 = (double)p2.y - (double)p1.y, u2 = (double)p2.z - (double)p1.z;
            double n = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
            vx = u0 / n; vy = u1 / n; vz = u2 / n; ax = p1.x; ay = p1.y; az = p1.z; type = 1;
            atomicAdd(a.corner_avail, 1);
          }
        }
      }
    } else {
      if (found5 && (double)t.d[4] < a.max_dis_plane) {
        if (a.icp_plane) {
          const float4 pa = __ldg(&tv.src[t.id[0]]), pb = __ldg(&tv.src[t.id[2]]), pc = __ldg(&tv.src[t.id[4]]);
          double b0 = (double)pb.x - (double)pa.x, b1 = (double)pb.y - (double)pa.y, b2 = (double)pb.z - (double)pa.z;
          double nb = sqrt(b0 * b0 + b1 * b1 + b2 * b2); b0 = b0 / nb; b1 = b1 / nb; b2 = b2 / nb;
          double c0 = (double)pc.x - (double)pa.x, c1 = (double)pc.y - (double)pa.y, c2 = (double)pc.z - (double)pa.z;
          double nc = sqrt(c0 * c0 + c1 * c1 + c2 * c2); c0 = c0 / nc; c1 = c1 / nc; c2 = c2 / nc;
          vx = b1 * c2 - b2 * c1; vy = b2 * c0 - b0 * c2; vz = b0 * c1 - b1 * c0;   // NOT re-normalised (ceres_icp.hpp:334)
          ax = pa.x; ay = pa.y; az = pa.z; type = 2;
        }
        atomicAdd(a.surf_avail, 1);
      }
    }
  }
  if (type != 0) atomicAdd(a.n_blocks, 1);   // residual_block_ids.size() of this ICP iteration (the cap's drop rule needs it, :436)
  a.blk_a[w] = make_float4((float)ax, (float)ay, (float)az, __int_as_float(type));
  a.blk_v[(size_t)w * 3 + 0] = vx; a.blk_v[(size_t)w * 3 + 1] = vy; a.blk_v[(size_t)w * 3 + 2] = vz;
}

// K6 + K7 fused, one query group (LL_GROUP lanes) per scan feature, features taken in spatially sorted order (perm).
// Writes one residual-block slot per feature (indexed by the ORIGINAL feature order): blk_a[slot] = (a.x, a.y, a.z, type) with
// type 0 invalid / 1 line / 2 plane, blk_v[slot*3..] = unit line direction or (un-normalised) plane normal, in fp64.
__global__ void __launch_bounds__(KNN_THREADS) knn_blocks_kernel(KnnBlocksArgs a) {
  if (a.st->icp_done) return;   // launched ahead of the termination test by the host: the ICP loop has already ended
  __shared__ GroupStack stacks[GROUPS_PER_CTA];
  const int j = blockIdx.x * GROUPS_PER_CTA + (threadIdx.x / GROUP);
  const int gl = threadIdx.x & (GROUP - 1);
  const int M = a.n_corner + a.n_surf;
  const bool have = j < M;
  const int w = have ? (a.perm ? a.perm[j] : j) : 0;      // original feature index
  const bool is_corner = w < a.n_corner;
  const TreeView& tv = is_corner ? a.corner : a.surf;
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f); if (have) f = __ldg(&a.feat[w]);
  // pointAssociateToMap (non-deblur branch): p_w = q_curr * p + t_curr in fp64, stored as fp32
  const double* qc = a.pose; const double* tc = a.pose + 4;
  double wx, wy, wz; qrot_d(qc, (double)f.x, (double)f.y, (double)f.z, wx, wy, wz);
  float qx = (float)(wx + tc[0]), qy = (float)(wy + tc[1]), qz = (float)(wz + tc[2]);
  if (a.deblur) {
    // pointAssociateToMap, deblur branch (:627-654, if_undistore_in_matching = 1): Rodrigues-interpolated increment applied before q_last
    const RegDevState* st = a.st;
    const double is = (double)refine_blur_f(f.w, (float)st->min_ts, (float)st->max_ts);
    if (is != 1.0) {
      const double th = st->interp_theta * is; double sn, cs; sincos(th, &sn, &cs); const double oc = 1.0 - cs;
      const double* H = st->interp_hat; const double* H2 = st->interp_hat_sq;
      const double px = (double)f.x, py = (double)f.y, pz = (double)f.z;
      double rx = ((1.0 + sn * H[0] + oc * H2[0]) * px + (sn * H[1] + oc * H2[1]) * py) + (sn * H[2] + oc * H2[2]) * pz;
      double ry = ((sn * H[3] + oc * H2[3]) * px + (1.0 + sn * H[4] + oc * H2[4]) * py) + (sn * H[5] + oc * H2[5]) * pz;
      double rz = ((sn * H[6] + oc * H2[6]) * px + (sn * H[7] + oc * H2[7]) * py) + (1.0 + sn * H[8] + oc * H2[8]) * pz;
      rx += st->x[4] * (is * 1.0); ry += st->x[5] * (is * 1.0); rz += st->x[6] * (is * 1.0);
      double ox, oy, oz; qrot_d(st->pose_last, rx, ry, rz, ox, oy, oz);
      qx = (float)(ox + st->pose_last[4]); qy = (float)(oy + st->pose_last[5]); qz = (float)(oz + st->pose_last[6]);
    }
  }
  // Non-finite features: the corner loop skips them explicitly (:242-245); the surface loop does not (:347-351), but a NaN query makes every
  // squared distance NaN and `NaN < 50.0` (:353) rejects the match, so the block set is the same.
  const bool finite_in = isfinite(f.x) && isfinite(f.y) && isfinite(f.z);
  bool owned = true;
  if (a.world > 1) owned = finite_in && __ldg(&a.shard_owner[shard_cell_index(a.grid, qx, qy, qz)]) == a.rank;   // shard.cu: every point of space has exactly one owner
  // Residual-block cap, pre-skip (:232-238 corners, :339-345 surfaces): with N features of this class and N > 2 cap, a feature is skipped when
  // rand * N > 2 cap (float arithmetic, like m_rand_float), before it is transformed or searched.  Drawn per (seed, ICP iteration, class, index).
  bool skipped = false;
  { const int ncls = is_corner ? a.n_corner : a.n_surf;
    if (have && ncls > 2 * a.cap) skipped = ll_cap_uniform_f(a.rng_seed, a.st->icp_iter, is_corner ? 0 : 1, is_corner ? w : w - a.n_corner) * (float)ncls > (float)(2 * a.cap); }
  const bool active = have && owned && finite_in && !skipped;
  Top5 t;
  group_knn5(tv, stacks[threadIdx.x / GROUP], active, qx, qy, qz, t, (a.seed_ids && have) ? a.seed_ids + (size_t)w * LL_KNN : nullptr);

  if (!have || gl != 0) return;
  emit_block(a, tv, is_corner, active, w, t);
}

int launch_knn_query(ll_ctx* ctx, const BucketTree& t, const float4* d_q, int nq, int* d_idx, float* d_d) {
  if (nq == 0) return LL_OK;
  knn_query_kernel<<<ll_div_up(nq, GROUPS_PER_CTA), KNN_THREADS, 0, ctx->stream>>>(make_view(t), d_q, nq, d_idx, d_d); ctx->launches++;
  LL_CUDA(ctx, cudaGetLastError());
  return LL_OK;
}
// Sorts the features spatially (once per registration, at the initial pose): perm[] lists corner features then surface features.
int launch_query_sort(ll_ctx* ctx, const KnnBlocksArgs& a, int* d_perm) {
  const int M = a.n_corner + a.n_surf;
  if (M == 0) return LL_OK;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (unsigned*)nullptr, (unsigned*)nullptr, (int*)nullptr, (int*)nullptr, M, 0, 22, ctx->stream);
  size_t o_k0 = 0, o_k1 = align256((size_t)M * 4), o_v0 = o_k1 + align256((size_t)M * 4), o_t = o_v0 + align256((size_t)M * 4);
  LL_CUDA(ctx, ctx->scratch.reserve(o_t + tmp + 256));
  char* base = ctx->scratch.as<char>();
  query_key_kernel<<<ll_div_up(M, 256), 256, 0, ctx->stream>>>(a, (unsigned*)(base + o_k0), (int*)(base + o_v0));
  LL_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + o_t, tmp, (unsigned*)(base + o_k0), (unsigned*)(base + o_k1), (int*)(base + o_v0), d_perm, M, 0, 22, ctx->stream));
  ctx->launches += 4;
  return LL_OK;
}
int launch_knn_blocks(ll_ctx* ctx, const KnnBlocksArgs& a) {
  int M = a.n_corner + a.n_surf;
  if (M == 0) return LL_OK;
  knn_blocks_kernel<<<ll_div_up(M, GROUPS_PER_CTA), KNN_THREADS, 0, ctx->stream>>>(a);
  ctx->launches++;
  LL_CUDA(ctx, cudaGetLastError());
  return LL_OK;
}
