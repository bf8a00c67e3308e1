// tests/cpp/shared_headers_capi.cc — host build of the product's shared headers (az_rng.h,
// az_math.h, az_connect4.h) behind a C ABI so pytest can compare them with the real
// libstdc++/glibc/reference on a CPU. Test scaffolding: the product never links this.
#include <cstring>

#include "az_rng.h"

using namespace b2az;
extern "C" {
void* azp_rng_new(uint64_t seed, int use_stream, uint64_t stream) {
  auto* r = new Pcg32;
  if (use_stream) pcg32_seed_stream(*r, seed, stream); else pcg32_seed(*r, seed);
  return r;
}
void azp_rng_free(void* r) { delete static_cast<Pcg32*>(r); }
uint32_t azp_rng_u32(void* r) { return pcg32_next(*static_cast<Pcg32*>(r)); }
void azp_rng_shuffle(void* r, uint32_t n, uint32_t* inout) { rng_shuffle(*static_cast<Pcg32*>(r), inout, n); }
void azp_rng_shuffle_discard(void* r, uint32_t n) { rng_shuffle_discard(*static_cast<Pcg32*>(r), n); }
float azp_rng_uniform01(void* r) { return rng_uniform01(*static_cast<Pcg32*>(r)); }
void azp_rng_gamma(void* r, float alpha, uint32_t n, float* out) {
  GammaDist g;
  gamma_init(g, alpha);
  for (uint32_t i = 0; i < n; ++i) out[i] = gamma_draw(*static_cast<Pcg32*>(r), g);
}
float azp_rng_gumbel(void* r) { return rng_gumbel(*static_cast<Pcg32*>(r)); }
float azp_logf(float x) { return az_logf(x); }
float azp_expf(float x) { return az_expf(x); }
float azp_powf(float x, float y) { return az_powf(x, y); }
}
