/* ORACLE — TEST INFRASTRUCTURE ONLY. The REFERENCE's image IO library, compiled from the headers where they lie
 * (/root/reference/Extern/stb/stb_image.h, stb_image_write.h; never copied) into oracle/_ref/libstb_ref.so by
 * oracle/Makefile's `ref` target. Exposes the three calls Src/Utilities.cpp makes on the IBL path (loadHdri :189-213,
 * saveHdri :244-255) and the 8-bit decode it uses for textures (:104-150), for tests/test_hdr_cache.py to hold
 * althea_host_save_hdri / althea_host_load_hdri to. */
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include "stb_image.h"
#include "stb_image_write.h"

float* ref_stbi_loadf_from_memory(const unsigned char* buf, int len, int* w, int* h) {
  int original = 0;
  return stbi_loadf_from_memory(buf, len, w, h, &original, 4); /* Utilities.cpp:198-204 */
}
unsigned char* ref_stbi_load_from_memory(const unsigned char* buf, int len, int* w, int* h) {
  int original = 0;
  return stbi_load_from_memory(buf, len, w, h, &original, 4); /* Utilities.cpp:110-116 */
}
int ref_stbi_write_hdr(const char* path, int w, int h, const float* rgba) { return stbi_write_hdr(path, w, h, 4, rgba); }
void ref_stbi_free(void* p) { stbi_image_free(p); }
