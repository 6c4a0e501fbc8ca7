/* Plain-C caller of the drop-in boundary: compiled as C99 against include/sdso_b200.h and linked with libsdso_b200.so, so a
 * signature that drifts between the header and the library breaks this build (ctypes argtypes are hand-declared and cannot see
 * that). Flow: ctx_create -> make_images (2 frames) -> tracker_set_ref -> track, i.e. FrameHessian::makeImages,
 * CoarseTracker::setCoarseTrackingRef / makeCoarseDepthL0 and trackNewestCoarse through the ABI.
 *
 * usage: abi_smoke <input.bin> <variant>
 *   input.bin: int32 w, h, n; float K[4], baseline; float img0[w*h], img1[w*h]; float uvidw[n*4]; double T0[12]
 *   prints:    "T <12 doubles>", "aff <2>", "res <5>", "ok <0|1>" with 17 significant digits, or "nodevice" (exit 3) when
 *              sdso_ctx_create reports SDSO_E_NODEVICE / SDSO_E_CUDA (there is no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "sdso_b200.h"

static int die(sdso_ctx* ctx, const char* what, int rc) {
  fprintf(stderr, "%s failed: %d (%s)\n", what, rc, ctx ? sdso_last_error(ctx) : "-");
  return 1;
}

int main(int argc, char** argv) {
  int32_t hdr[3];
  float K[4], baseline;
  FILE* f;
  size_t npx;
  float *img0, *img1, *uvidw;
  double T[12], aff[2] = {0.0, 0.0}, minRes[5], lastRes[5], flow[3];
  int iters[5], ok = 0, f0 = -1, f1 = -1, rc, i, variant;
  sdso_ctx* ctx = NULL;
  sdso_settings S;
  if (argc < 3) { fprintf(stderr, "usage: %s input.bin variant\n", argv[0]); return 2; }
  variant = atoi(argv[2]);
  f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  if (fread(hdr, sizeof(int32_t), 3, f) != 3 || fread(K, sizeof(float), 4, f) != 4 || fread(&baseline, sizeof(float), 1, f) != 1) return 2;
  npx = (size_t)hdr[0] * (size_t)hdr[1];
  img0 = (float*)malloc(npx * sizeof(float)); img1 = (float*)malloc(npx * sizeof(float)); uvidw = (float*)malloc((size_t)hdr[2] * 4 * sizeof(float));
  if (fread(img0, sizeof(float), npx, f) != npx || fread(img1, sizeof(float), npx, f) != npx ||
      fread(uvidw, sizeof(float), (size_t)hdr[2] * 4, f) != (size_t)hdr[2] * 4 || fread(T, sizeof(double), 12, f) != 12) return 2;
  fclose(f);
  sdso_default_settings(&S);
  rc = sdso_ctx_create(&ctx, 0, hdr[0], hdr[1], K, baseline, &S);
  if (rc == SDSO_E_NODEVICE || rc == SDSO_E_CUDA) { printf("nodevice\n"); return 3; }
  if (rc != SDSO_OK) return die(ctx, "sdso_ctx_create", rc);
  if ((rc = sdso_frame_create(ctx, &f0)) != SDSO_OK || (rc = sdso_frame_create(ctx, &f1)) != SDSO_OK) return die(ctx, "sdso_frame_create", rc);
  if ((rc = sdso_make_images(ctx, f0, img0, 1.0f, 1)) != SDSO_OK) return die(ctx, "sdso_make_images", rc);
  if ((rc = sdso_make_images(ctx, f1, img1, 1.0f, 1)) != SDSO_OK) return die(ctx, "sdso_make_images", rc);
  if ((rc = sdso_tracker_make_k(ctx, K)) != SDSO_OK) return die(ctx, "sdso_tracker_make_k", rc);
  if ((rc = sdso_tracker_set_ref(ctx, f0, uvidw, hdr[2], aff)) != SDSO_OK) return die(ctx, "sdso_tracker_set_ref", rc);
  for (i = 0; i < 5; i++) minRes[i] = NAN;
  rc = sdso_track(ctx, f1, T, aff, sdso_pyr_levels(ctx) - 1, minRes, variant, lastRes, flow, iters, &ok);
  if (rc != SDSO_OK) return die(ctx, "sdso_track", rc);
  printf("T");
  for (i = 0; i < 12; i++) printf(" %.17g", T[i]);
  printf("\naff %.17g %.17g\nres", aff[0], aff[1]);
  for (i = 0; i < 5; i++) printf(" %.17g", lastRes[i]);
  printf("\nok %d\nlaunches %llu\n", ok, (unsigned long long)sdso_launch_count(ctx));
  sdso_ctx_destroy(ctx);
  free(img0); free(img1); free(uvidw);
  return 0;
}
