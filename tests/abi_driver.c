/*
 * abi_driver.c -- plain C (gcc, no Python, no Lua) driver that replays ONE EPOCH of adversarial.train
 * (/root/reference/adversarial.lua:27-292) through include/catgen.h in the call order the Lua loop has, one C-ABI call per
 * nn.Module method the closures invoke:
 *
 *   adversarial.lua:51-68   epoch loop, tail batch, <4 abort        -> the for loop below
 *   :223-238                B/2 real + B/2 fake (separate G forward) -> cg_G_forward + memcpy
 *   :72-112  fevalD         zero grads, D fwd, BCE, D bwd, penalty+clamp, confusion -> cg_model_zero_grads, cg_D_forward, cg_bce,
 *                                                                     cg_D_backward, cg_penalty_clamp
 *   :245     optim.adam(fevalD, PARAMETERS_D, OPTSTATE.adam.D)       -> cg_adam_step(t, 0, cfg)
 *   :171-215 fevalG_on_D    zero grads, G fwd, D fwd, BCE vs 1, D bwd (gradInput), G bwd, clamp -> cg_G_forward, cg_D_forward, cg_bce,
 *                                                                     cg_D_backward, cg_G_backward, cg_penalty_clamp
 *   :262     optim.adam(fevalG_on_D, PARAMETERS_G, ...)              -> cg_adam_step(t, 1, cfg)
 *
 * This is what a LuaJIT-FFI shim does (cat-generator_b200/lua/), written in the one language that can be executed in this image.
 * The same epoch is then run through the fused entry point cg_train_step on a second (G, D) pair created from the same seeds with
 * the same data and noise: both paths must report the same losses (first step: to rounding; later steps: the trajectory floor,
 * profiles/r01_parity_noise_floor.txt) and the same confusion counts within a few flips.  Exit code 0 and a final line
 * "ABI_DRIVER_OK" mean every call succeeded and the two paths agreed.
 *
 *   gcc -std=c99 -O2 -I include tests/abi_driver.c -o /tmp/abi_driver -L cat-generator_b200 -lcatgen -Wl,-rpath,$PWD/cat-generator_b200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "catgen.h"

#define CHECK(call) do { int s_ = (call); if (s_ != CG_OK) { fprintf(stderr, "FAILED %s -> %d: %s\n", #call, s_, cg_last_error()); return 2; } } while (0)

static unsigned long long lcg_state = 88172645463325252ULL;
static float urand(void) {   /* xorshift64*: uniform [0,1) */
  lcg_state ^= lcg_state >> 12; lcg_state ^= lcg_state << 25; lcg_state ^= lcg_state >> 27;
  return (float)((lcg_state * 2685821657736338717ULL) >> 40) / 16777216.0f;
}
static void fill_uniform(float* a, long n, float lo, float hi) { for (long i = 0; i < n; ++i) a[i] = lo + (hi - lo) * urand(); }

typedef struct { long c[2][2]; } confusion;   /* optim.ConfusionMatrix({"0","1"}) (train.lua:188): [prediction][target] */
static void confusion_add(confusion* m, const float* out, int B) {   /* adversarial.lua:101-106 */
  for (int i = 0; i < B; ++i) m->c[out[i] > 0.5f ? 1 : 0][i < B / 2 ? 1 : 0]++;
}
static double total_valid(const confusion* m) { long t = m->c[0][0] + m->c[0][1] + m->c[1][0] + m->c[1][1]; return t ? (double)(m->c[0][0] + m->c[1][1]) / t : 0.0; }

int main(int argc, char** argv) {
  const int C = 3, NZ = 100, IMG = 3 * 32 * 32;
  const int B = argc > 1 ? atoi(argv[1]) : 8, N_epoch = argc > 2 ? atoi(argv[2]) : 26, NDATA = 64;
  cg_step_cfg cfg = {B, 1, 1, 0.f, 1e-4f, 0.f, 0.f, 1.f, 5.f, 1e-3f, 0.9f, 0.999f, 1e-8f};   /* train.lua:26-36 + optim.adam defaults */
  CHECK(cg_init(0));
  printf("%s\n", cg_version());
  cg_model *G[2], *D[2]; cg_trainer* T[2];
  for (int v = 0; v < 2; ++v) {   /* two identical pairs: v = 0 per-module calls, v = 1 fused step */
    CHECK(cg_model_create(&G[v], CG_G32UPC, C, NZ, 1)); CHECK(cg_model_create(&D[v], CG_D32_ST3, C, NZ, 2));
    CHECK(cg_trainer_create(&T[v], G[v], D[v]));
  }
  int64_t npG = 0, npD = 0; CHECK(cg_model_nparams(G[0], &npG)); CHECK(cg_model_nparams(D[0], &npD));
  printf("getParameters(): G %lld, D %lld\n", (long long)npG, (long long)npD);
  if (npG != 5191687 || npD != 6664777) { fprintf(stderr, "parameter counts differ from models.lua\n"); return 3; }

  float* data = (float*)malloc(sizeof(float) * (size_t)NDATA * IMG);
  fill_uniform(data, (long)NDATA * IMG, 0.f, 1.f);   /* dataset.lua:131-149: images in [0,1] */
  float *inputs = malloc(sizeof(float) * (size_t)B * IMG), *samples = malloc(sizeof(float) * (size_t)B * IMG), *gimg = malloc(sizeof(float) * (size_t)B * IMG);
  float *zD = malloc(sizeof(float) * (size_t)B * NZ), *zG = malloc(sizeof(float) * (size_t)B * NZ), *real = malloc(sizeof(float) * (size_t)B * IMG);
  float *targets = malloc(sizeof(float) * B), *ones = malloc(sizeof(float) * B), *out = malloc(sizeof(float) * B), *df = malloc(sizeof(float) * B), *dout = malloc(sizeof(float) * B);
  confusion conf[2]; memset(conf, 0, sizeof(conf));
  int steps = 0; double worst_first = 0, worst_later = 0;

  for (int t = 1; t <= N_epoch; t += B / 2) {                      /* adversarial.lua:51 */
    int thisB = B < N_epoch - t + 1 ? B : N_epoch - t + 1;         /* :53 */
    if (thisB < 4) { printf("<trainer> batch of %d examples skipped (adversarial.lua:65-68)\n", thisB); break; }
    thisB -= thisB % 2;
    const int hB = thisB / 2;
    cfg.B = thisB;
    for (int i = 0; i < hB; ++i) memcpy(real + (size_t)i * IMG, data + (size_t)((int)(urand() * NDATA) % NDATA) * IMG, sizeof(float) * IMG);   /* :225-230 */
    fill_uniform(zD, (long)hB * NZ, -1.f, 1.f); fill_uniform(zG, (long)thisB * NZ, -1.f, 1.f);   /* nn_utils.lua:35-39 */
    for (int i = 0; i < thisB; ++i) { targets[i] = i < hB ? 1.f : 0.f; ones[i] = 1.f; }          /* train.lua:70-71 */

    /* ---- v = 0: the closures, one call per module method */
    float lossD, penD, lossG, penG;
    memcpy(inputs, real, sizeof(float) * (size_t)hB * IMG);
    CHECK(cg_G_forward(G[0], zD, hB, inputs + (size_t)hB * IMG));                                /* :233 createImages */
    CHECK(cg_model_zero_grads(D[0]));                                                            /* :81 */
    CHECK(cg_D_forward(D[0], inputs, thisB, out, NULL));                                         /* :84 */
    CHECK(cg_bce(out, targets, thisB, &lossD, df));                                              /* :85,88 */
    CHECK(cg_D_backward(D[0], df, NULL));                                                        /* :89 */
    CHECK(cg_penalty_clamp(D[0], cfg.D_L1, cfg.D_L1, cfg.D_L2, cfg.D_clamp, &penD));             /* :92-112 */
    confusion_add(&conf[0], out, thisB);
    CHECK(cg_adam_step(T[0], 0, &cfg));                                                          /* :245 */
    CHECK(cg_model_zero_grads(G[0]));                                                            /* :177 */
    CHECK(cg_G_forward(G[0], zG, thisB, samples));                                               /* :185 */
    CHECK(cg_D_forward(D[0], samples, thisB, out, NULL));                                        /* :187 */
    CHECK(cg_bce(out, ones, thisB, &lossG, df));                                                 /* :188,191 */
    CHECK(cg_D_backward(D[0], df, gimg));                                                        /* :192-193 MODEL_D.modules[1].gradInput */
    CHECK(cg_G_backward(G[0], gimg, NULL));                                                      /* :197 */
    CHECK(cg_penalty_clamp(G[0], cfg.G_L1, cfg.G_L2, cfg.G_L2, cfg.G_clamp, &penG));             /* :201-212 */
    CHECK(cg_adam_step(T[0], 1, &cfg));                                                          /* :262 */

    /* ---- v = 1: the same step through the fused entry point */
    float fD, fG;
    CHECK(cg_train_step(T[1], &cfg, real, zD, zG, &fD, &fG, dout));
    confusion_add(&conf[1], dout, thisB);

    double eD = fabs((double)(lossD + penD) - fD), eG = fabs((double)(lossG + penG) - fG);
    printf("step %2d B=%2d  per-module lossD %.6f lossG %.6f | fused lossD %.6f lossG %.6f | diff %.2e %.2e\n", steps, thisB, lossD + penD, lossG + penG, fD, fG, eD, eG);
    if (!isfinite(fD) || !isfinite(fG) || !isfinite(lossD) || !isfinite(lossG)) { fprintf(stderr, "non-finite loss\n"); return 4; }
    if (steps == 0) { worst_first = eD > eG ? eD : eG; } else { if (eD > worst_later) worst_later = eD; if (eG > worst_later) worst_later = eG; }
    ++steps;
  }
  CHECK(cg_sync());
  printf("<trainer> steps %d, totalValid per-module %.4f fused %.4f\n", steps, total_valid(&conf[0]), total_valid(&conf[1]));
  /* parameters after the epoch: both pairs took the same updates wherever the gradient is not sign-noise */
  float *p0 = malloc(sizeof(float) * npD), *p1 = malloc(sizeof(float) * npD);
  CHECK(cg_model_get_params(D[0], p0)); CHECK(cg_model_get_params(D[1], p1));
  long far = 0; for (long i = 0; i < npD; ++i) if (fabsf(p0[i] - p1[i]) > 0.5e-3f * steps) ++far;
  printf("D parameters further apart than steps*lr/2: %.4f %%\n", 100.0 * far / npD);
  for (int v = 0; v < 2; ++v) { CHECK(cg_trainer_free(T[v])); CHECK(cg_model_free(G[v])); CHECK(cg_model_free(D[v])); }
  cg_shutdown();
  if (worst_first > 1e-4) { fprintf(stderr, "first step differs by %.3e between the per-module sequence and cg_train_step\n", worst_first); return 5; }
  if (worst_later > 2e-2) { fprintf(stderr, "later steps differ by %.3e\n", worst_later); return 6; }
  if ((double)far / npD > 0.05) { fprintf(stderr, "parameter trajectories diverged\n"); return 7; }
  printf("ABI_DRIVER_OK\n");
  return 0;
}
