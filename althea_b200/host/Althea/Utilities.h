// Host-side mirror of the image-file IO the IBL cache goes through: Utilities::loadHdri / saveHdri
// (Include/Althea/Utilities.h:29-53, Src/Utilities.cpp:189-255). The reference delegates to stb_image.h's stbi_loadf and
// stb_image_write.h's stbi_write_hdr; this header reads and writes the same Radiance RGBE container without them:
//   * encode: per texel the largest of r, g, b fixes a shared exponent (frexp), the three mantissas are TRUNCATED to
//     8 bits, values below 1e-32 become four zero bytes; alpha is not stored;
//   * container: "#?RADIANCE" text header, "-Y <h> +X <w>", then per row either flat RGBE quadruples (width < 8 or
//     >= 32768) or a 2,2,hi,lo marker followed by the four byte planes of the row, each run-length coded (a run is three
//     or more equal bytes, at most 127 per packet; anything else goes out as literals, at most 128 per packet);
//   * decode: mantissa * 2^(e - 136), e == 0 means zero, alpha comes back as 1.
// Files written here are byte-identical to stbi_write_hdr's apart from the one comment line of the header, and
// decoding agrees bit for bit with stbi_loadf (tests/test_hdr_cache.py, against oracle/_ref's build of the reference's stb).
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace AltheaEngine {

class Utilities {
public:
  struct ImageFile { // Include/Althea/Utilities.h:29-35
    int width;
    int height;
    int channels;
    int bytesPerChannel;
    std::vector<std::byte> data;
  };

  // one texel, linear rgb -> r, g, b mantissas + shared exponent
  static void linearToRgbe(const float* linear, uint8_t* rgbe) {
    const float top = std::fmax(linear[0], std::fmax(linear[1], linear[2]));
    if (top < 1e-32f) {
      rgbe[0] = rgbe[1] = rgbe[2] = rgbe[3] = 0;
      return;
    }
    int exponent;
    const float normalize = (float)std::frexp(top, &exponent) * 256.0f / top;
    rgbe[0] = (uint8_t)(linear[0] * normalize);
    rgbe[1] = (uint8_t)(linear[1] * normalize);
    rgbe[2] = (uint8_t)(linear[2] * normalize);
    rgbe[3] = (uint8_t)(exponent + 128);
  }

  static void rgbeToLinear(const uint8_t* rgbe, float* rgba) {
    if (rgbe[3] != 0) {
      const float scale = (float)std::ldexp(1.0f, (int)rgbe[3] - 136);
      rgba[0] = rgbe[0] * scale;
      rgba[1] = rgbe[1] * scale;
      rgba[2] = rgbe[2] * scale;
    } else {
      rgba[0] = rgba[1] = rgba[2] = 0.0f;
    }
    rgba[3] = 1.0f;
  }

  // The whole file as bytes. `channels` floats per texel (3 or 4), row 0 first.
  static std::vector<uint8_t> encodeHdri(int width, int height, int channels, const float* texels) {
    if (width <= 0 || height <= 0 || !texels || (channels != 3 && channels != 4))
      throw std::runtime_error("encodeHdri: invalid image");
    std::vector<uint8_t> out;
    const std::string head = "#?RADIANCE\n# Written by althea_b200 (stb_image_write.h layout)\nFORMAT=32-bit_rle_rgbe\n"
                             "EXPOSURE=          1.0000000000000\n\n-Y " + std::to_string(height) + " +X " + std::to_string(width) + "\n";
    out.insert(out.end(), head.begin(), head.end());
    std::vector<uint8_t> planes((size_t)width * 4);
    for (int y = 0; y < height; ++y) {
      const float* row = texels + (size_t)y * width * channels;
      if (width < 8 || width >= 32768) {
        for (int x = 0; x < width; ++x) {
          uint8_t q[4];
          linearToRgbe(row + (size_t)x * channels, q);
          out.insert(out.end(), q, q + 4);
        }
        continue;
      }
      for (int x = 0; x < width; ++x) {
        uint8_t q[4];
        linearToRgbe(row + (size_t)x * channels, q);
        for (int c = 0; c < 4; ++c) planes[(size_t)c * width + x] = q[c];
      }
      const uint8_t marker[4] = {2, 2, (uint8_t)((width >> 8) & 0xff), (uint8_t)(width & 0xff)};
      out.insert(out.end(), marker, marker + 4);
      for (int c = 0; c < 4; ++c) packPlane(planes.data() + (size_t)c * width, width, out);
    }
    return out;
  }

  // Decodes a file image; returns false on anything stbi_loadf would reject. rgba: 4 floats per texel, alpha 1.
  static bool decodeHdri(const uint8_t* file, size_t size, int& width, int& height, std::vector<float>& rgba) {
    size_t at = 0;
    auto line = [&](std::string& s) {
      s.clear();
      if (at >= size) return false;
      while (at < size && file[at] != '\n') s.push_back((char)file[at++]);
      if (at < size) ++at;
      return true;
    };
    std::string s;
    if (!line(s) || (s != "#?RADIANCE" && s != "#?RGBE")) return false;
    bool format = false;
    for (;;) {
      if (!line(s)) return false;
      if (s.empty()) break;
      if (s == "FORMAT=32-bit_rle_rgbe") format = true;
    }
    if (!format || !line(s)) return false;
    long h = 0, w = 0;
    char tail = 0;
    if (std::sscanf(s.c_str(), "-Y %ld +X %ld%c", &h, &w, &tail) != 2 || h <= 0 || w <= 0 || h > (1 << 24) || w > (1 << 24))
      return false;
    width = (int)w;
    height = (int)h;
    rgba.assign((size_t)w * h * 4, 0.0f);
    const bool flatFile = w < 8 || w >= 32768 || at + 4 > size || file[at] != 2 || file[at + 1] != 2 || (file[at + 2] & 0x80);
    if (flatFile) {
      if (size - at < (size_t)w * h * 4) return false;
      for (size_t i = 0; i < (size_t)w * h; ++i) rgbeToLinear(file + at + 4 * i, rgba.data() + 4 * i);
      return true;
    }
    std::vector<uint8_t> planes((size_t)w * 4);
    for (long y = 0; y < h; ++y) {
      if (at + 4 > size || file[at] != 2 || file[at + 1] != 2 || (((int)file[at + 2] << 8) | file[at + 3]) != w) return false;
      at += 4;
      for (int c = 0; c < 4; ++c) {
        long x = 0;
        while (x < w) {
          if (at >= size) return false;
          int count = file[at++];
          if (count > 128) {
            count -= 128;
            if (count > w - x || at >= size) return false;
            std::memset(planes.data() + (size_t)c * w + x, file[at++], (size_t)count);
          } else {
            if (count == 0 || count > w - x || at + (size_t)count > size) return false;
            std::memcpy(planes.data() + (size_t)c * w + x, file + at, (size_t)count);
            at += (size_t)count;
          }
          x += count;
        }
      }
      for (long x = 0; x < w; ++x) {
        const uint8_t q[4] = {planes[x], planes[(size_t)w + x], planes[(size_t)2 * w + x], planes[(size_t)3 * w + x]};
        rgbeToLinear(q, rgba.data() + ((size_t)y * w + x) * 4);
      }
    }
    return true;
  }

  // Src/Utilities.cpp:189-213: four float channels, alpha = 1
  static void loadHdri(const std::string& path, ImageFile& result) {
    std::vector<uint8_t> file = readFile(path);
    std::vector<float> rgba;
    if (!decodeHdri(file.data(), file.size(), result.width, result.height, rgba))
      throw std::runtime_error("Failed to load HDR image: " + path);
    result.channels = 4;
    result.bytesPerChannel = 4;
    result.data.resize(rgba.size() * sizeof(float));
    std::memcpy(result.data.data(), rgba.data(), result.data.size());
  }

  // Src/Utilities.cpp:244-255: `data` holds width * height RGBA32F texels (the reference takes a gsl::span of bytes)
  static void saveHdri(const std::string& path, int width, int height, const std::byte* data, size_t bytes) {
    if (bytes < (size_t)width * height * 16) throw std::runtime_error("saveHdri: buffer smaller than the image");
    const std::vector<uint8_t> file = encodeHdri(width, height, 4, reinterpret_cast<const float*>(data));
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Failed to open file for writing: " + path);
    const size_t n = std::fwrite(file.data(), 1, file.size(), f);
    std::fclose(f);
    if (n != file.size()) throw std::runtime_error("Failed to write file: " + path);
  }

  // Src/Utilities.cpp:258-271: `data` holds width * height RGBA32F texels; tinyexr's SaveEXR(data, w, h, 4, /*fp16*/0, path): an
  // OpenEXR 2 scan-line file with the four channels A, B, G, R as 32-bit FLOAT, increasing-y line order, data window
  // (0, 0)-(w-1, h-1). Written here uncompressed (tinyexr deflates blocks of 16 lines; every OpenEXR reader, tinyexr's LoadEXR
  // included, reads either): the decoded image is the same bits.
  static void saveExr(const std::string& path, int width, int height, const std::byte* data, size_t bytes) {
    if (width <= 0 || height <= 0 || bytes < (size_t)width * height * 16) throw std::runtime_error("saveExr: buffer smaller than the image");
    const float* rgba = reinterpret_cast<const float*>(data);
    std::vector<uint8_t> f;
    auto u8 = [&](uint8_t v) { f.push_back(v); };
    auto i32 = [&](int32_t v) { for (int k = 0; k < 4; ++k) f.push_back((uint8_t)((uint32_t)v >> (8 * k))); };
    auto f32 = [&](float v) { uint32_t b; std::memcpy(&b, &v, 4); i32((int32_t)b); };
    auto str = [&](const char* z) { while (*z) f.push_back((uint8_t)*z++); f.push_back(0); };
    auto attr = [&](const char* name, const char* type, int32_t size) { str(name); str(type); i32(size); };
    i32(20000630); // magic
    i32(2);        // version 2, single-part scan lines
    attr("channels", "chlist", 4 * 18 + 1);
    for (const char* c : {"A", "B", "G", "R"}) { str(c); i32(2 /*FLOAT*/); u8(0); u8(0); u8(0); u8(0); i32(1); i32(1); }
    u8(0);
    attr("compression", "compression", 1); u8(0 /*NO_COMPRESSION*/);
    attr("dataWindow", "box2i", 16); i32(0); i32(0); i32(width - 1); i32(height - 1);
    attr("displayWindow", "box2i", 16); i32(0); i32(0); i32(width - 1); i32(height - 1);
    attr("lineOrder", "lineOrder", 1); u8(0 /*INCREASING_Y*/);
    attr("pixelAspectRatio", "float", 4); f32(1.0f);
    attr("screenWindowCenter", "v2f", 8); f32(0.0f); f32(0.0f);
    attr("screenWindowWidth", "float", 4); f32(1.0f);
    u8(0); // end of header
    const size_t lineBytes = (size_t)width * 16, block = 8 + lineBytes;
    const uint64_t first = f.size() + (uint64_t)height * 8;
    for (int y = 0; y < height; ++y) { // offset table: one block per scan line
      const uint64_t o = first + (uint64_t)y * block;
      for (int k = 0; k < 8; ++k) f.push_back((uint8_t)(o >> (8 * k)));
    }
    f.reserve(f.size() + (size_t)height * block);
    for (int y = 0; y < height; ++y) {
      i32(y);
      i32((int32_t)lineBytes);
      for (int c : {3, 2, 1, 0}) // channel-planar lines in the order of the channel list: A, B, G, R
        for (int x = 0; x < width; ++x) f32(rgba[((size_t)y * width + x) * 4 + c]);
    }
    FILE* out = std::fopen(path.c_str(), "wb");
    if (!out) throw std::runtime_error("Failed to open file for writing: " + path);
    const size_t n = std::fwrite(f.data(), 1, f.size(), out);
    std::fclose(out);
    if (n != f.size()) throw std::runtime_error("Failed to write file: " + path);
  }

  static std::vector<uint8_t> readFile(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Failed to open file: " + path);
    std::vector<uint8_t> buf;
    uint8_t chunk[1 << 16];
    size_t n;
    while ((n = std::fread(chunk, 1, sizeof chunk, f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    std::fclose(f);
    return buf;
  }

private:
  // one byte plane of a row -> packets
  static void packPlane(const uint8_t* p, int n, std::vector<uint8_t>& out) {
    int x = 0;
    while (x < n) {
      int run = x;  // first position from x where three equal bytes start, or n if there is none
      while (run + 2 < n && !(p[run] == p[run + 1] && p[run] == p[run + 2])) ++run;
      const bool found = run + 2 < n;
      if (!found) run = n;
      while (x < run) {
        const int len = run - x > 128 ? 128 : run - x;
        out.push_back((uint8_t)len);
        out.insert(out.end(), p + x, p + x + len);
        x += len;
      }
      if (found) {
        int end = run;
        while (end < n && p[end] == p[x]) ++end;
        while (x < end) {
          const int len = end - x > 127 ? 127 : end - x;
          out.push_back((uint8_t)(len + 128));
          out.push_back(p[x]);
          x += len;
        }
      }
    }
  }
};

} // namespace AltheaEngine
