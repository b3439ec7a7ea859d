// Round 1's search (warp per query, best-first with a shared-memory stack, replicated insertion-sorted top-5), kept behind -DLL_KNN_R1 as the
// measured baseline of the round-2 rewrite in knn.cu (and as a known-good fallback): same inputs, same results.  Included from knn.cu only.
#pragma once
namespace r1 {
// ------------------------------------------------------------------------------------------------ search
// LL_GROUP lanes per query (32: one warp per query).  A step of a query's best-first search pops one item from the group's stack
// (shared memory) and handles it with all lanes at once: a NODE -> lane c tests child c's box (two coalesced 16-B loads per lane,
// 1 KB per group) and the qualifying children are pushed far-to-near in one shot (ballot + popc); a BUCKET -> lane c takes
// point c (one 16-B load, 512 B per group) and the candidates are merged into the group's top-5.  The top-5 and the stack pointer are
// replicated in the lanes.  Exact: a box gives a true lower bound of the fp32 distance and (d2, index) is a total order.
#define GROUP 32
#define GROUPS_PER_CTA (KNN_THREADS / GROUP)
#define STACK_CAP 160
#define ITEM_BUCKET 0x80000000u


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

// adapter to the lane-distributed result the kernels of round 2 consume
__device__ __forceinline__ void warp_knn5_r1(const TreeView& tv, GroupStack& st, bool active, float qx, float qy, float qz, LaneTop& t, const int* seed_ids) {
  Top5 o; group_knn5(tv, st, active, qx, qy, qz, o, seed_ids);
  const int lane = threadIdx.x & 31;
  t.d = INFINITY; t.id = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < LL_KNN; j++) if (lane == j) { t.d = o.d[j]; t.id = o.id[j]; }
  t.d5 = o.d[LL_KNN - 1]; t.id5 = o.id[LL_KNN - 1];
}
}  // namespace r1
