/* capi_smoke.c -- a plain C host of the drop-in boundary: compiled against include/fg_b200.h ONLY (no Python, no
 * torch, no C++), linked to libfg_b200.so.  Runs two adversarial.lua loop bodies (fg_train_step) at batch 16 on
 * host buffers, then the same step composed from the L-net calls, and prints the statistics a host would feed
 * into optim.ConfusionMatrix / OPTSTATE.  Exit code 0 = every call succeeded and the numbers are sane.
 *   gcc -std=c99 -Iinclude tests/capi_smoke.c -Lface_generator_b200 -lfg_b200 -Wl,-rpath,face_generator_b200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fg_b200.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != FG_OK) {                                                              \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, fg_last_error());          \
      return 2;                                                                      \
    }                                                                                \
  } while (0)

static unsigned long long rng_state = 88172645463325252ULL;
static float urand(void) { /* xorshift64: U[0,1) */
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return (float)((rng_state >> 40) * (1.0 / 16777216.0));
}
static float nrand(void) { /* Box-Muller */
  float u = urand() + 1e-7f, v = urand();
  return sqrtf(-2.0f * logf(u)) * cosf(6.2831853f * v);
}

int main(void) {
  const int B = 16, C = 3;
  fg_ctx* ctx = NULL;
  CHECK(fg_create(&ctx, 0, B, C));
  printf("%s\n", fg_version());
  const long nG = (long)fg_param_count(FG_NET_G, C), nD = (long)fg_param_count(FG_NET_D, C);
  float* pG = (float*)malloc(sizeof(float) * nG);
  float* pD = (float*)malloc(sizeof(float) * nD);
  /* NN_UTILS.initializeWeights scale (utils/nn_utils.lua:17-29) would give sigmoid(0) everywhere; use 0.05 */
  for (long i = 0; i < nG; ++i) pG[i] = 0.05f * nrand();
  for (long i = 0; i < nD; ++i) pD[i] = 0.05f * nrand();
  CHECK(fg_set_params(ctx, FG_NET_G, pG));
  CHECK(fg_set_params(ctx, FG_NET_D, pD));
  float* real = (float*)malloc(sizeof(float) * (B / 2) * C * 1024);
  float* noiseD = (float*)malloc(sizeof(float) * (B / 2) * 100);
  float* noiseG = (float*)malloc(sizeof(float) * B * 100);
  fg_hyper h;
  fg_hyper_default(&h);
  fg_step_stats st;
  for (int it = 1; it <= 2; ++it) {
    for (int i = 0; i < (B / 2) * C * 1024; ++i) real[i] = urand();
    for (int i = 0; i < (B / 2) * 100; ++i) noiseD[i] = 2.0f * urand() - 1.0f;
    for (int i = 0; i < B * 100; ++i) noiseG[i] = 2.0f * urand() - 1.0f;
    CHECK(fg_train_step(ctx, &h, B, real, noiseD, noiseG, NULL, NULL, (unsigned long long)it, &st));
    printf("step %d: loss_D %.6f loss_G %.6f conf [%d %d %d %d] trained_D %d t_D %d t_G %d acc_D %.3f\n", it, st.loss_D,
           st.loss_G, st.conf[0], st.conf[1], st.conf[2], st.conf[3], st.trained_D, st.t_D, st.t_G, st.acc_D);
    if (!(st.loss_D > 0 && st.loss_D < 50 && st.loss_G > 0 && st.loss_G < 50)) return 3;
    if (st.conf[0] + st.conf[1] + st.conf[2] + st.conf[3] != B || st.t_D != it || st.t_G != it || st.trained_D != 1) return 4;
  }
  /* L-net level: MODEL_G:forward / MODEL_D:forward / :backward (what b200.FusedG / b200.FusedD call) */
  float* img = (float*)malloc(sizeof(float) * B * C * 1024);
  float* out = (float*)malloc(sizeof(float) * B);
  float* dout = (float*)malloc(sizeof(float) * B);
  float* dimg = (float*)malloc(sizeof(float) * B * C * 1024);
  CHECK(fg_G_forward(ctx, noiseG, B, 1, img));
  CHECK(fg_D_forward(ctx, img, B, 1, NULL, 7ULL, out));
  float loss = 0.f;
  for (int i = 0; i < B; ++i) dout[i] = 1.0f; /* targets = Y_NOT_GENERATOR */
  CHECK(fg_bce_forward(ctx, out, dout, B, &loss));
  CHECK(fg_bce_backward(ctx, out, dout, B, dout));
  CHECK(fg_zero_grads(ctx, FG_NET_G));
  CHECK(fg_D_backward(ctx, dout, 0, dimg));
  CHECK(fg_G_backward(ctx, dimg, NULL));
  CHECK(fg_get_grads(ctx, FG_NET_G, pG));
  double gn = 0;
  for (long i = 0; i < nG; ++i) gn += (double)pG[i] * pG[i];
  printf("L-net: BCE %.6f, |grad G| %.6e, image range [%.3f, %.3f], kernels launched %lld\n", loss, sqrt(gn), img[0], img[1],
         (long long)fg_kernel_launches(ctx));
  if (!(loss > 0 && gn > 0 && isfinite(gn))) return 5;
  for (int i = 0; i < B * C * 1024; ++i)
    if (!(img[i] >= 0.f && img[i] <= 1.f)) return 6;
  /* error behaviour: an odd batch is refused with a message, nothing crashes */
  if (fg_train_step(ctx, &h, 7, real, noiseD, noiseG, NULL, NULL, 1ULL, &st) != FG_ERR_INVALID) return 7;
  printf("expected error: %s\n", fg_last_error());
  CHECK(fg_destroy(ctx));
  free(pG); free(pD); free(real); free(noiseD); free(noiseG); free(img); free(out); free(dout); free(dimg);
  printf("capi_smoke OK\n");
  return 0;
}
