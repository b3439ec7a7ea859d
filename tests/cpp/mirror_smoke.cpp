// Compiled, linked against libloamlivox_b200.so and run by tests/test_cabi.py (no GPU needed to build; without a GPU the run stops at the context).
// Written the way laser_mapping.hpp would use the mirror: the reference's class names over the C-ABI.
#include <cstdio>
#include "loamlivox_b200.hpp"

int main() {
  ll_reg_state y; ll_reg_state_yaml(&y, 0);
  if (y.maximum_allow_residual_block != 200 || y.max_final_cost != 2.0) { std::printf("yaml defaults wrong\n"); return 2; }
  if (!(ll_cap_uniform(1, 2, 1, 3) >= 0.f && ll_cap_uniform(1, 2, 1, 3) < 1.f)) return 3;
  int dims[3] = {4, 2, 1}, counts[8] = {5, 5, 5, 5, 5, 5, 5, 5}, owner[8];
  if (ll_shard_plan(counts, dims, 2, owner) != LL_OK || owner[0] != 0) return 4;
  try {
    ll200::Context ctx(0);
    ll200::Livox_laser laser(ctx);
    ll200::PointCloud scan(16);
    for (size_t i = 0; i < scan.size(); i++) { scan[i].x = 1.f + 0.01f * i; scan[i].y = 0.1f * i; scan[i].z = 0.f; scan[i].intensity = 50.f; }
    const int petals = laser.extract_laser_features(scan, 100.0);
    std::printf("GPU run: %d petals\n", petals);
  } catch (const ll200::Error& e) {
    if (e.status != LL_ERR_CUDA) { std::printf("unexpected status %d\n", e.status); return 5; }
    std::printf("no GPU: context refused (LL_ERR_CUDA), as it must -- there is no CPU fallback\n");
  }
  return 0;
}
