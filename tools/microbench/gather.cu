// Divergent 32-byte record gather: LDG.256 vs two 128-bit texture fetches vs a mix (tuning aid for the SSAO proxy march).
// Every lane reads records at hash-random positions inside a (2R)^2 window around its CTA's tile of a 3841 x 2161 record grid,
// as the SSAO march does; reported: records per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int GW = 3841, GH = 2161;
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

struct Rec { uint32_t w[8]; };
__device__ __forceinline__ Rec ldg256(const void* p) {
  Rec r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
  return r;
}

template <int MODE> __global__ void __launch_bounds__(256, 4) gather(const Rec* recs, cudaTextureObject_t tex, uint32_t* out, int taps, int R) {
  const int tx = blockIdx.x * 16 + (threadIdx.x & 15), ty = blockIdx.y * 16 + (threadIdx.x >> 4);
  uint32_t s = hash(tx * 9781u + ty * 6271u + 1u), acc = 0;
  for (int t = 0; t < taps; ++t) {
    s = s * 1664525u + 1013904223u;
    const int dx = (int)((s >> 8) % (2u * R)) - R, dy = (int)((s >> 20) % (2u * R)) - R;
    const int x = min(max(tx + dx, 0), GW - 1), y = min(max(ty + dy, 0), GH - 1);
    const int idx = y * GW + x;
    const bool useTex = MODE == 1 || (MODE == 2 && (t & 1)) || (MODE == 3 && (t % 3 == 2));
    if (MODE == 4) { // 16-byte records: the same grid at half the bytes (one LDG.128 per tap)
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(recs) + idx);
      acc += a.x ^ a.y ^ a.z ^ a.w;
    } else if (useTex) {
      uint4 a = tex1Dfetch<uint4>(tex, 2 * idx), b = tex1Dfetch<uint4>(tex, 2 * idx + 1);
      acc += a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    } else {
      Rec r = ldg256(recs + idx);
      acc += r.w[0] ^ r.w[1] ^ r.w[2] ^ r.w[3] ^ r.w[4] ^ r.w[5] ^ r.w[6] ^ r.w[7];
    }
  }
  out[ty * 3840 + tx] = acc;
}

template <int MODE> static void run(const char* name, const Rec* recs, cudaTextureObject_t tex, uint32_t* out, int R, int sms, double ghz) {
  const int taps = 128;
  dim3 grid(3840 / 16, 2160 / 16);
  gather<MODE><<<grid, 256>>>(recs, tex, out, 8, R);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  gather<MODE><<<grid, 256>>>(recs, tex, out, taps, R);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double recsTotal = 3840.0 * 2160.0 * taps;
  printf("{\"mode\": \"%s\", \"R\": %d, \"ms\": %.3f, \"records_per_clk_per_sm\": %.3f, \"Grec_per_s\": %.1f}\n", name, R, ms,
         recsTotal / (ms * 1e-3 * ghz * 1e9) / sms, recsTotal / ms * 1e-6);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double ghz = clk * 1e-6;
  Rec* recs;
  const size_t n = (size_t)GW * GH;
  cudaMalloc(&recs, n * sizeof(Rec));
  cudaMemset(recs, 1, n * sizeof(Rec));
  uint32_t* out;
  cudaMalloc(&out, 3840 * 2160 * 4);
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeLinear;
  rd.res.linear.devPtr = recs;
  rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
  rd.res.linear.sizeInBytes = n * sizeof(Rec);
  cudaTextureDesc td = {};
  td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex = 0;
  cudaError_t e = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  if (e != cudaSuccess) { printf("texture: %s\n", cudaGetErrorString(e)); return 1; }
  for (int R : {8, 32, 64, 96}) {
    run<0>("ldg256", recs, tex, out, R, p.multiProcessorCount, ghz);
    run<1>("tex2x128", recs, tex, out, R, p.multiProcessorCount, ghz);
    run<2>("mix 1:1", recs, tex, out, R, p.multiProcessorCount, ghz);
    run<3>("mix 2:1", recs, tex, out, R, p.multiProcessorCount, ghz);
    run<4>("ldg128 (16-byte records)", recs, tex, out, R, p.multiProcessorCount, ghz);
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
