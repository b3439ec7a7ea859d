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
// Hilbert index (Skilling's transpose form, `bits` per axis, <= 21) on an ISOTROPIC grid (one scale for all axes): consecutive runs along a
// Hilbert curve are connected blobs, so fixed-size buckets get tight, nearly cubic boxes.  The key only decides WHICH points share a bucket (the
// boxes are computed from the points themselves, so the search is exact for any key): its width follows the map size (about 8 cells per bucket
// side at the finest level), which is what bounds the number of radix passes of the sort below.
template <typename KeyT>
__global__ void hilbert_kernel(const float4* __restrict__ src, int n, const int* __restrict__ bbox, int bits, KeyT* __restrict__ keys, int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = src[i];
  unsigned long long key = 1ull << (3 * bits);   // non-finite points: one bit above every real key -> they sort to the end
  if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
    const float lo[3] = {ord2f(bbox[0]), ord2f(bbox[1]), ord2f(bbox[2])}, hi[3] = {ord2f(bbox[3]), ord2f(bbox[4]), ord2f(bbox[5])};
    const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    const float c[3] = {p.x, p.y, p.z}; unsigned X[3];
    const float scale = (float)(1u << bits), top = scale - 1.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float u = ext > 0.f ? (c[k] - lo[k]) / ext : 0.f;
      u = fminf(fmaxf(u, 0.f), 1.f);
      X[k] = (unsigned)fminf(u * scale, top);
    }
    for (unsigned Q = 1u << (bits - 1); Q > 1; Q >>= 1) {
      const unsigned P = Q - 1;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (X[k] & Q) X[0] ^= P;
        else { unsigned t = (X[0] ^ X[k]) & P; X[0] ^= t; X[k] ^= t; }
      }
    }
    X[1] ^= X[0]; X[2] ^= X[1];
    unsigned t = 0;
    for (unsigned Q = 1u << (bits - 1); Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    key = (spread21(X[0]) << 2) | (spread21(X[1]) << 1) | spread21(X[2]);   // < 2^(3 bits)
  }
  keys[i] = (KeyT)key;
  vals[i] = i;
}
// n_valid (= bbox[6], the number of finite points) is read on the device: the host never waits for it.
__global__ void gather_kernel(const float4* __restrict__ src, const int* __restrict__ order, const int* __restrict__ bbox, int n_pad, float4* __restrict__ pts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const int n_valid = bbox[6];
  if (i < n_valid) { int j = order[i]; float4 p = src[j]; p.w = __int_as_float(j); pts[i] = p; }
  else pts[i] = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7fffffff));
}
// One thread per (node, child).  Level 0: the children are buckets of 32 points (pad points are +inf and do not count).  Level l > 0: the children
// are level l-1 nodes (box = union of that node's child boxes).  Missing / empty children get the neutral box (+inf, -inf).
__global__ void node_kernel(const float4* __restrict__ pts, const float4* __restrict__ child_nodes, int n_child, int n_nodes, float4* __restrict__ nodes) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = g / FANOUT, c = g % FANOUT;
  if (j >= n_nodes) return;
  float l0 = INFINITY, l1 = INFINITY, l2 = INFINITY, h0 = -INFINITY, h1 = -INFINITY, h2 = -INFINITY;
  const int ci = j * FANOUT + c;
  if (ci < n_child) {
    if (child_nodes == nullptr) {
      for (int k = 0; k < BUCKET; k++) {
        const float4 p = pts[ci * BUCKET + k];
        if (p.x < INFINITY) { l0 = fminf(l0, p.x); l1 = fminf(l1, p.y); l2 = fminf(l2, p.z); h0 = fmaxf(h0, p.x); h1 = fmaxf(h1, p.y); h2 = fmaxf(h2, p.z); }
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
__global__ void bbox_publish_kernel(const int* __restrict__ bbox, float* __restrict__ out6) {   // decoded box for the query-sort kernel
  if (threadIdx.x < 6) out6[threadIdx.x] = ord2f(bbox[threadIdx.x]);
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Everything is enqueued on the context's stream and nothing is read back: the layout is sized from n_src (the non-finite points -- there are
// usually none -- only leave a few all-pad buckets with neutral boxes at the end), the count of finite points and the box stay on the device.
int build_bucket_tree(ll_ctx* ctx, const float4* d_src, int n_src, BucketTree* t) { return build_bucket_tree_on(ctx, ctx->stream, ctx->scratch, d_src, n_src, t); }
// Same on an explicit stream with its own scratch arena: the corner and the surface index of one map are built side by side (every kernel of a
// 100k-point build is far too small to fill the GPU).
int build_bucket_tree_on(ll_ctx* ctx, cudaStream_t s, DevBuf& scratch, const float4* d_src, int n_src, BucketTree* t) {
  { DevBuf keep = t->storage; *t = BucketTree(); t->storage = keep; }   // re-indexing in place (ll_map_rebuild) reuses the allocation
  t->n_src = n_src;
  // key width: lidar maps are surfaces, so a grid of 2^b cells per axis has ~4^b occupied cells; b = ceil(log2(n) / 2) - 1 keeps the occupied
  // cells well below a bucket's 32 points (5M points: b = 11, 34 sorted bits = 5 radix passes instead of 8; 30k points: b = 7, 3 passes)
  int lg = 0; while ((1ll << lg) < (long long)(n_src > 0 ? n_src : 1)) lg++;
  int bits = (lg + 1) / 2 - 1; if (bits < 5) bits = 5; if (bits > 21) bits = 21;
  const int key_bits = 3 * bits, sort_bits = key_bits + 1;   // one more bit: non-finite points get the key 2^key_bits and sort behind every real one
  const bool k32 = sort_bits <= 32;
  const size_t ksz = k32 ? 4 : 8;
  // scratch: bbox(8 ints) | keys | keys_out | vals | vals_out | cub temp
  size_t temp_bytes = 0;
  if (k32) cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned*)nullptr, (unsigned*)nullptr, (int*)nullptr, (int*)nullptr, n_src > 0 ? n_src : 1, 0, sort_bits, s);
  else cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, n_src > 0 ? n_src : 1, 0, sort_bits, s);
  size_t off_bbox = 0, off_k0 = align256(64), off_k1 = off_k0 + align256((size_t)n_src * ksz), off_v0 = off_k1 + align256((size_t)n_src * ksz),
         off_v1 = off_v0 + align256((size_t)n_src * 4), off_tmp = off_v1 + align256((size_t)n_src * 4);
  LL_CUDA(ctx, scratch.reserve(off_tmp + temp_bytes + 256));
  char* base = scratch.as<char>();
  int* bbox = (int*)(base + off_bbox);
  int* v0 = (int*)(base + off_v0); int* v1 = (int*)(base + off_v1);
  t->n = n_src; t->n_pad = ll_div_up(n_src > 0 ? n_src : 1, BUCKET) * BUCKET;
  // level sizes: level 0 nodes have buckets as children; the top level has exactly one node
  int cnt = ll_div_up(t->n_pad / BUCKET, FANOUT); t->n_levels = 0;
  for (;;) { t->level_count[t->n_levels++] = cnt; if (cnt <= 1) break; if (t->n_levels == LL_MAX_LEVELS) { ctx->set_error("map too large for LL_MAX_LEVELS"); return LL_ERR_CAPACITY; } cnt = ll_div_up(cnt, FANOUT); }
  // node storage is laid out top level first
  size_t node_total = 0; for (int l = 0; l < t->n_levels; l++) node_total += (size_t)t->level_count[l];
  size_t bytes = align256((size_t)t->n_pad * 16) + align256((size_t)n_src * 16) + align256(node_total * NODE_F4 * 16) + 256;
  LL_CUDA(ctx, t->storage.reserve(bytes));
  char* p = t->storage.as<char>();
  t->pts = (float4*)p; p += align256((size_t)t->n_pad * 16);
  float4* nodes = (float4*)p; p += align256(node_total * NODE_F4 * 16);
  { size_t off = 0; for (int l = t->n_levels - 1; l >= 0; l--) { t->lo[l] = nodes + off * NODE_F4; off += (size_t)t->level_count[l]; } }
  t->hi[0] = nodes;   // base of the node array (top level first)
  t->src = (float4*)p; p += align256((size_t)n_src * 16);
  t->d_bbox = (float*)p;
  bbox_init_kernel<<<1, 32, 0, s>>>(bbox); ctx->launches++;
  if (n_src > 0) {
    int grid = min(ll_div_up(n_src, 256), ctx->num_sms * 8);
    bbox_kernel<<<grid, 256, 0, s>>>(d_src, n_src, bbox); ctx->launches++;
    if (k32) {
      unsigned* k0 = (unsigned*)(base + off_k0); unsigned* k1 = (unsigned*)(base + off_k1);
      hilbert_kernel<unsigned><<<ll_div_up(n_src, 256), 256, 0, s>>>(d_src, n_src, bbox, bits, k0, v0); ctx->launches++;
      LL_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + off_tmp, temp_bytes, k0, k1, v0, v1, n_src, 0, sort_bits, s));
    } else {
      unsigned long long* k0 = (unsigned long long*)(base + off_k0); unsigned long long* k1 = (unsigned long long*)(base + off_k1);
      hilbert_kernel<unsigned long long><<<ll_div_up(n_src, 256), 256, 0, s>>>(d_src, n_src, bbox, bits, k0, v0); ctx->launches++;
      LL_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + off_tmp, temp_bytes, k0, k1, v0, v1, n_src, 0, sort_bits, s));
    }
    ctx->launches += 3 + (sort_bits + 7) / 8;
    LL_CUDA(ctx, cudaMemcpyAsync(t->src, d_src, (size_t)n_src * 16, cudaMemcpyDeviceToDevice, s));
  }
  bbox_publish_kernel<<<1, 32, 0, s>>>(bbox, t->d_bbox); ctx->launches++;
  gather_kernel<<<ll_div_up(t->n_pad, 256), 256, 0, s>>>(d_src, v1, bbox, t->n_pad, t->pts); ctx->launches++;
  for (int l = 0; l < t->n_levels; l++) {
    const int n_child = l == 0 ? t->n_pad / BUCKET : t->level_count[l - 1];
    node_kernel<<<ll_div_up(t->level_count[l] * FANOUT, 128), 128, 0, s>>>(t->pts, l == 0 ? nullptr : t->lo[l - 1], n_child, t->level_count[l], t->lo[l]); ctx->launches++;
  }
  LL_CUDA(ctx, cudaGetLastError());
  return LL_OK;
}

TreeView make_view(const BucketTree& t) {
  TreeView v; v.pts = t.pts; v.src = t.src; v.n = t.n; v.n_levels = t.n_levels;
  for (int l = 0; l < LL_MAX_LEVELS; l++) v.nodes[l] = t.lo[l];
  v.bbox = t.d_bbox;
  return v;
}

// ------------------------------------------------------------------------------------------------ search
// One warp per query.  Depth-first over the 32-ary box tree, nearest child first at every node: for the node being expanded at level l, lane c
// holds the lower bound of child c (registers -> a 128-byte row of shared memory per level), a 32-bit mask says which children are still to
// be visited, and "pick" = one REDUX.MIN over the bounds of the remaining children that can still beat the current 5th distance.  The bound is
// tested when a child is PICKED, against the distance list as it is then -- nothing stale is ever expanded, and once the nearest remaining
// child of a node fails the test the whole node is finished.  A bucket = 32 points, one per lane (one coalesced 512-byte load).
// The 5 best (d2, index) pairs live in lanes 0..4, sorted; an insertion is a ballot + two SHFL.UP.  Exact: a box gives a true lower bound of the
// fp32 distance, (d2, index) is a total order, and ties on d2 keep the smaller index, so the result does not depend on the visiting order.
#define KNN_THREADS 256
#define WARPS_PER_CTA (KNN_THREADS / 32)
#define KNN_LEVELS LL_MAX_LEVELS   // 32^8 buckets: more than any map that fits the HBM

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

// The query's 5 best so far: lane j < 5 holds the j-th smallest (d, id); d5 / id5 = lane 4's entry, known to every lane.
struct LaneTop { float d; int id; float d5; int id5; };
__device__ __forceinline__ void top_init(LaneTop& t) { t.d = INFINITY; t.id = 0x7fffffff; t.d5 = INFINITY; t.id5 = 0x7fffffff; }
// (md, mi) is warp-uniform and lex-less than the 5th entry.  An index that is already in the list is ignored (seeds are real points of the map).
__device__ __forceinline__ void top_insert(LaneTop& t, float md, int mi, int lane) {
  const bool mine = lane < LL_KNN;
  if (__ballot_sync(FULL, mine && t.id == mi)) return;
  const int pos = __popc(__ballot_sync(FULL, mine && !lex_less(md, mi, t.d, t.id)));   // entries that stay in front of the newcomer
  const float ud = __shfl_up_sync(FULL, t.d, 1); const int ui = __shfl_up_sync(FULL, t.id, 1);
  if (mine && lane >= pos) { t.d = lane == pos ? md : ud; t.id = lane == pos ? mi : ui; }
  t.d5 = __shfl_sync(FULL, t.d, LL_KNN - 1); t.id5 = __shfl_sync(FULL, t.id, LL_KNN - 1);
}
// Merge this step's candidates (one per lane, flag c): repeatedly take the smallest one that still beats the 5th entry (<= 5 rounds).
__device__ __forceinline__ void top_merge(LaneTop& t, float d, int id, bool c, int lane) {
  c = c && lex_less(d, id, t.d5, t.id5);   // NaN and +inf distances fail here
  while (__any_sync(FULL, c)) {
    const unsigned key = c ? __float_as_uint(d) : 0xffffffffu;   // non-negative floats order like their bit patterns
    const unsigned mn = __reduce_min_sync(FULL, key);
    const int mi = (int)__reduce_min_sync(FULL, (c && key == mn) ? (unsigned)id : 0x7fffffffu);
    top_insert(t, __uint_as_float(mn), mi, lane);
    if (id == mi && key == mn) c = false;
    c = c && lex_less(d, id, t.d5, t.id5);
  }
}

struct WarpWalk { float lb[KNN_LEVELS][32]; };   // per warp: the child bounds of the node open at every level

// ---- optional TMA staging of leaf buckets (LL_KNN_TMA=1; north_star's "TMA/shared-memory staging of KD-tree leaf buckets") --------------------
// When a level-0 node is opened, the two nearest qualifying buckets are fetched with cp.async.bulk (512 B each) into a warp-private double buffer
// in shared memory, completion on an mbarrier; the pick that reaches such a bucket waits on the barrier and reads its point from shared memory
// instead of issuing the load itself.  Same visiting order, same results.  Measured against the plain variant in profiles/r2.
struct __align__(16) WarpStage { float4 pts[2][32]; unsigned long long bar[2]; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// All 32 lanes of the warp call this together; `active` is warp-uniform.
// seed_ids: the query's 5 neighbours of the previous ICP iteration (or null / -1): their distances to the moved query seed the list, so the
// bound is tight before the first box is opened.
template <bool TMA>
__device__ __forceinline__ void warp_knn5(const TreeView& tv, WarpWalk& ws, WarpStage* stg, bool active, float qx, float qy, float qz, LaneTop& t, const int* seed_ids) {
  const int lane = threadIdx.x & 31;
  top_init(t);
  if (!(active && tv.n > 0)) return;
  int pre_child[2] = {-1, -1}; unsigned pre_phase[2] = {0u, 0u};   // TMA: which bucket sits (or is landing) in each stage, and the barrier's next parity
  if (TMA) { if (lane == 0) { mbar_init(&stg->bar[0], 1); mbar_init(&stg->bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); } __syncwarp(); }
  if (seed_ids) {
    int sid = -1;
    if (lane < LL_KNN) sid = seed_ids[lane];
    if (__any_sync(FULL, sid >= 0)) {
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sid >= 0) p = __ldg(tv.src + sid);
      top_merge(t, sid >= 0 ? dist2_exact(qx, qy, qz, p.x, p.y, p.z) : INFINITY, sid, sid >= 0, lane);
    }
  }
  const int top = tv.n_levels - 1;
  int l = top, idx = 0;          // the node open at level l is node `idx` of that level
  unsigned remv = 0;             // lane k keeps the mask of the children still to visit of the node open at level k
  bool open_node = true;
  for (;;) {
    if (open_node) {             // expand node idx of level l: one box per lane (2 x 16-byte loads, 1 KB per node)
      const float4* r = tv.nodes[l] + (size_t)idx * NODE_F4 + 2 * lane;
      const float4 A = __ldg(r), B = __ldg(r + 1);
      const float lb = box_lb(A.x, A.y, A.z, B.x, B.y, B.z, qx, qy, qz);   // a missing child has the neutral box (+inf, -inf): lb = +inf
      ws.lb[l][lane] = lb;
      const unsigned m = __ballot_sync(FULL, lb <= t.d5 && lb < INFINITY);
      if (lane == l) remv = m;
      open_node = false;   // (every lane only ever reads back its own slot of the row: no synchronisation needed)
      if (TMA && l == 0 && m) {   // stage the two nearest qualifying buckets
        const unsigned k1 = ((m >> lane) & 1u) ? __float_as_uint(lb) : 0xffffffffu;
        const unsigned m1 = __reduce_min_sync(FULL, k1);
        const int c1 = __ffs(__ballot_sync(FULL, k1 == m1)) - 1;
        const unsigned k2 = lane == c1 ? 0xffffffffu : k1;
        const unsigned m2 = __reduce_min_sync(FULL, k2);
        const int c2 = m2 == 0xffffffffu ? -1 : __ffs(__ballot_sync(FULL, k2 == m2)) - 1;
        const int want[2] = {idx * FANOUT + c1, c2 >= 0 ? idx * FANOUT + c2 : -1};
#pragma unroll
        for (int b = 0; b < 2; b++) {
          if (pre_child[b] >= 0) { mbar_wait(&stg->bar[b], pre_phase[b]); pre_phase[b] ^= 1u; pre_child[b] = -1; }   // a staged bucket that was pruned: let its copy land first
          if (want[b] >= 0) {
            __syncwarp(); fence_proxy_async();   // every lane is done reading the stage (generic proxy) before the async proxy overwrites it
            if (lane == 0) { mbar_expect_tx(&stg->bar[b], 512u); tma_bulk_g2s(stg->pts[b], tv.pts + (size_t)want[b] * BUCKET, 512u, &stg->bar[b]); }
            pre_child[b] = want[b];
          }
        }
      }
    }
    // pick the nearest remaining child of the node open at level l that can still hold a neighbour
    const float lb = ws.lb[l][lane];
    const unsigned rem = __shfl_sync(FULL, remv, l);
    const bool ok = ((rem >> lane) & 1u) && lb <= t.d5;
    const unsigned key = ok ? __float_as_uint(lb) : 0xffffffffu;
    const unsigned mn = __reduce_min_sync(FULL, key);
    if (mn == 0xffffffffu) {     // nothing left here: back to the parent (whose bounds are still in its row)
      if (l == top) break;
      l++; idx >>= 5; continue;
    }
    const unsigned who = __ballot_sync(FULL, key == mn);
    const int c = __ffs(who) - 1;
    const unsigned okm = __ballot_sync(FULL, ok);
    if (lane == l) remv = okm & ~(1u << c);   // (okm is a subset of rem) the children that failed the test now can never pass it later
    const int child = idx * FANOUT + c;
    if (l == 0) {                // a bucket: one point per lane
      float4 P;
      if (TMA && (child == pre_child[0] || child == pre_child[1])) {
        const int b = child == pre_child[0] ? 0 : 1;
        mbar_wait(&stg->bar[b], pre_phase[b]); pre_phase[b] ^= 1u; pre_child[b] = -1;
        P = stg->pts[b][lane];
      } else P = __ldg(tv.pts + (size_t)child * BUCKET + lane);
      top_merge(t, dist2_exact(qx, qy, qz, P.x, P.y, P.z), __float_as_int(P.w), true, lane);   // pad points are +inf: never candidates
    } else { l--; idx = child; open_node = true; }
  }
  if (TMA) {   // no copy may still be in flight when the CTA's shared memory goes away
#pragma unroll
    for (int b = 0; b < 2; b++) if (pre_child[b] >= 0) mbar_wait(&stg->bar[b], pre_phase[b]);
  }
}

#ifdef LL_KNN_R1
#include "knn_r1.cuh"
#endif

// Parity hook (ll_knn): world-frame queries in caller order.
__global__ void __launch_bounds__(KNN_THREADS) knn_query_kernel(TreeView tv, const float4* __restrict__ q, int nq, int* __restrict__ idx5, float* __restrict__ d5) {
  __shared__ WarpWalk walks[WARPS_PER_CTA];
  const int g = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const bool have = g < nq;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f); if (have) p = __ldg(&q[g]);
  const bool active = have && isfinite(p.x) && isfinite(p.y) && isfinite(p.z);
  LaneTop t;
#ifdef LL_KNN_R1
  __shared__ r1::GroupStack stacks[WARPS_PER_CTA];
  r1::warp_knn5_r1(tv, stacks[threadIdx.x >> 5], active, p.x, p.y, p.z, t, nullptr);
#else
  warp_knn5<false>(tv, walks[threadIdx.x >> 5], nullptr, active, p.x, p.y, p.z, t, nullptr);
#endif
  if (have && lane < LL_KNN) { idx5[g * LL_KNN + lane] = (t.id == 0x7fffffff) ? -1 : t.id; d5[g * LL_KNN + lane] = t.d; }
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
  const float* bb = is_corner ? a.corner.bbox : a.surf.bbox;   // device memory: the box never visits the host
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

// Gates + functor constructors (K7) for one feature whose 5 nearest neighbours are id[0..4] (d4 = the 5th squared distance); lane 0 writes slot `w`.
__device__ __forceinline__ void emit_block(const KnnBlocksArgs& a, const TreeView& tv, bool is_corner, bool active, int w, int id0, int id1, int id2, int id4, float d4) {
  int type = 0; double ax = 0, ay = 0, az = 0, vx = 0, vy = 0, vz = 0;
  if (active) {
    const bool found5 = id4 != 0x7fffffff;
    if (is_corner) {
      if (found5 && (double)d4 < a.max_dis_line) {
        if (a.icp_line) {
          const float4 p1 = __ldg(&tv.src[id0]), p2 = __ldg(&tv.src[id1]);
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
      if (found5 && (double)d4 < a.max_dis_plane) {
        if (a.icp_plane) {
          const float4 pa = __ldg(&tv.src[id0]), pb = __ldg(&tv.src[id2]), pc = __ldg(&tv.src[id4]);
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
template <bool TMA>
__global__ void __launch_bounds__(KNN_THREADS) knn_blocks_kernel(KnnBlocksArgs a) {
  if (a.st->icp_done) return;   // launched ahead of the termination test by the host: the ICP loop has already ended
  __shared__ WarpWalk walks[WARPS_PER_CTA];
  __shared__ WarpStage stages[TMA ? WARPS_PER_CTA : 1];
  const int j = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
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
  LaneTop t;
#ifdef LL_KNN_R1
  __shared__ r1::GroupStack stacks[WARPS_PER_CTA];
  r1::warp_knn5_r1(tv, stacks[threadIdx.x >> 5], active, qx, qy, qz, t, (a.seed_ids && have) ? a.seed_ids + (size_t)w * LL_KNN : nullptr);
#else
  warp_knn5<TMA>(tv, walks[threadIdx.x >> 5], TMA ? &stages[threadIdx.x >> 5] : nullptr, active, qx, qy, qz, t, (a.seed_ids && have) ? a.seed_ids + (size_t)w * LL_KNN : nullptr);
#endif
  if (!have) return;
  if (active && lane < LL_KNN) {   // the neighbours seed the next ICP iteration's search (5 lanes, one 20-byte row)
    if (a.seed_ids) a.seed_ids[(size_t)w * LL_KNN + lane] = (t.id == 0x7fffffff) ? -1 : t.id;
    if (a.knn_d) a.knn_d[(size_t)w * LL_KNN + lane] = t.d;
  }
  const int id0 = __shfl_sync(FULL, t.id, 0), id1 = __shfl_sync(FULL, t.id, 1), id2 = __shfl_sync(FULL, t.id, 2);
  if (lane != 0) return;
  emit_block(a, tv, is_corner, active, w, id0, id1, id2, t.id5, t.d5);
}

int launch_knn_query(ll_ctx* ctx, const BucketTree& t, const float4* d_q, int nq, int* d_idx, float* d_d) {
  if (nq == 0) return LL_OK;
  knn_query_kernel<<<ll_div_up(nq, WARPS_PER_CTA), KNN_THREADS, 0, ctx->stream>>>(make_view(t), d_q, nq, d_idx, d_d); ctx->launches++;
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
  if (ctx->knn_tma) knn_blocks_kernel<true><<<ll_div_up(M, WARPS_PER_CTA), KNN_THREADS, 0, ctx->stream>>>(a);
  else knn_blocks_kernel<false><<<ll_div_up(M, WARPS_PER_CTA), KNN_THREADS, 0, ctx->stream>>>(a);
  ctx->launches++;
  LL_CUDA(ctx, cudaGetLastError());
  return LL_OK;
}
