/* A plain-C consumer of libjxb.so: what a maintainer binding the boundary from another language links against.
 * Host-only calls (key algebra) must work anywhere; device calls must fail loudly without a GPU. */
#include <stdio.h>
#include <string.h>

#include "jxb.h"

int main(void) {
  if (jxb_version() != JXB_VERSION) return 1;
  /* jax.random.split(PRNGKey(0)) under both stream layouts (tests/golden/threefry_kat.json) */
  const uint32_t key[2] = {0u, 0u};
  uint32_t out[4];
  if (jxb_prng_split(JXB_RNG_LEGACY, key, 2, out) != JXB_OK) return 2;
  if (out[0] != 4146024105u || out[1] != 967050713u || out[2] != 2718843009u || out[3] != 1272950319u) return 3;
  if (jxb_prng_split(JXB_RNG_PARTITIONABLE, key, 2, out) != JXB_OK) return 4;
  if (out[0] != 1797259609u || out[1] != 2579123966u || out[2] != 928981903u || out[3] != 3453687069u) return 5;
  float u = 0.f;
  if (jxb_prng_uniform(JXB_RNG_LEGACY, key, 1, 0.0f, 1.0f, &u) != JXB_OK) return 6;
  if (u < 0.41845702f || u > 0.41845704f) return 7;
  jxb_engine* eng = NULL;
  const int rc = jxb_engine_create(0, &eng);
  if (rc == JXB_OK) {                 /* a GPU is present: the engine works, nothing more to check here */
    int sms = 0;
    if (jxb_engine_sm_count(eng, &sms) != JXB_OK || sms <= 0) return 8;
    jxb_engine_destroy(eng);
    printf("abi ok (device with %d SMs)\n", sms);
    return 0;
  }
  if (rc != JXB_ERR_NO_DEVICE || strstr(jxb_last_error(), "no CPU fallback") == NULL) return 9;
  printf("abi ok (no device: %s)\n", jxb_last_error());
  return 0;
}
