// TEST INFRASTRUCTURE ONLY. Thin driver around the reference's vendored tinyexr (Extern/tinyexr), compiled where it lies into
// oracle/_ref/libtinyexr_ref.so: SaveEXR exactly as Utilities::saveExr calls it (Src/Utilities.cpp:258-271) and LoadEXR, so the
// tests can hold the product's EXR writer to what the reference's own library reads and writes.
#include <stdlib.h>
#include <string.h>

#include "tinyexr.h"

extern "C" {
int ref_save_exr(const char* path, int width, int height, const float* rgba) { return SaveEXR(rgba, width, height, 4, 0, path, nullptr); }
// decodes into out (RGBA, 4 floats per texel); returns 0 and the size, or tinyexr's error code
int ref_load_exr(const char* path, float* out, unsigned long long capacity_floats, int* width, int* height) {
  float* img = nullptr;
  const char* err = nullptr;
  int rc = LoadEXR(&img, width, height, path, &err);
  if (rc != TINYEXR_SUCCESS) { if (err) FreeEXRErrorMessage(err); return rc; }
  const unsigned long long n = 4ull * (unsigned long long)(*width) * (unsigned long long)(*height);
  if (n <= capacity_floats) memcpy(out, img, n * sizeof(float));
  free(img);
  return n <= capacity_floats ? 0 : -100;
}
}
