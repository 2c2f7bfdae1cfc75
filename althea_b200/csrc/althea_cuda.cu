// The extern "C" shim: context, resource table (imported / wrapped / owned device memory), and the entry points that
// turn the reference's parameter blocks into kernel launches. See include/althea_cuda.h for what each one replaces.
// No CPU fallback anywhere: a compute entry point either launches sm_100a kernels or returns an error.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/althea_cuda.h"
#include "launchers.h"
#include "raster_launchers.h"

namespace {

thread_local std::string g_createError;

enum class ResKind { Image, Buffer, Semaphore };

struct Resource {
  ResKind kind = ResKind::Buffer;
  void* dptr = nullptr;
  size_t bytes = 0;
  uint32_t format = 0, w = 0, h = 0, mips = 1, layers = 1;
  size_t pitch0 = 0; // row pitch of level 0 (bytes)
  bool owned = false;
  cudaExternalMemory_t extMem = nullptr;
  cudaExternalSemaphore_t extSem = nullptr;
  bool timeline = false;
};

struct TimingEntry {
  const char* name;
  cudaEvent_t start, stop;
};

} // namespace

struct RasterScratch;
static void freeRasterScratch(RasterScratch* r);

struct althea_cuda_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint32_t flags = 0;
  std::string lastError;
  std::unordered_map<uint64_t, Resource> resources;
  uint64_t nextHandle = 1;
  uint64_t launches = 0;
  uint32_t scissorY0 = 0, scissorY1 = 0; // rows of the final image this ctx shades; y1 == 0 => whole frame
  // Engine-internal scratch of the per-frame stages, one set per CUDA stream the stages are called on: the reference keeps
  // MAX_FRAMES_IN_FLIGHT = 2 frames in flight (Include/Althea/Library.h:3) and althea_sync.cuda_stream lets a host do the same
  // here, so two frames' scratch must not alias. A set is written and consumed inside ONE stage call, in stream order.
  struct Scratch {
    void* ao = nullptr; size_t aoBytes = 0;             // SSAO occluded-ray counts (when the caller passes no ao_counts image)
    void* position = nullptr; size_t positionBytes = 0; // mode D: positions reconstructed from depth
    void* quad = nullptr; size_t quadBytes = 0;         // SSAO position-quad proxy, (W+1) x (H+1) x 32 B
    void* plane = nullptr; size_t planeBytes = 0;       // SSAO plane records (three levels, padded), tile hand-over list, reciprocal depths
    void* depthPad = nullptr; size_t depthPadBytes = 0; // SSR padded depth, (W+2) x (H+2) floats
    void* ssrPlane = nullptr;                           // SSR plane records, kSsrPlaneStride x kSsrPlaneRows x 16 B
    void* ssrHits = nullptr; size_t ssrHitsBytes = 0;   // SSR hit list: a counter, then 12 floats per pixel of the launch (worst case: every pixel hits)
  };
  std::map<cudaStream_t, Scratch> scratch;
  struct RasterScratch* raster = nullptr; // scratch of the rasterising producers (draw_gbuffer / draw_shadow_cubes)
  unsigned long long* gatherCounter = nullptr; // device counters (4) of the ALTHEA_CTX_SSAO_COUNT_TAPS diagnostic
  // SSAO ray-direction table (FrameParams::ssaoDirs): read-only once built, shared by every stream; rebuilt when a larger frame arrives
  void* ssaoDirs = nullptr; int ssaoDirRow = 0, ssaoDirRows = 0; bool ssaoDirsParity = false; // (which build's arithmetic filled it)
  bool ssaoDirTable = true; // ALTHEA_SSAO_DIR_TABLE=0 in the environment: hash inline instead (A/B switch)
  // timing
  bool timing = false;
  std::vector<TimingEntry> pending;
  std::vector<cudaEvent_t> eventPool;
  std::map<std::string, std::pair<double, uint32_t>> totals;
  std::vector<const char*> totalNames;
};

namespace {

int fail(althea_cuda_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->lastError = buf;
  else g_createError = buf;
  return code;
}

#define CUDA_TRY(ctx, expr)                                                                              \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess) {                                                                             \
      cudaGetLastError(); /* the runtime's sticky last error must not surface in a later, unrelated call */ \
      return fail(ctx, ALTHEA_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));                 \
    }                                                                                                    \
  } while (0)

size_t bytesPerTexel(uint32_t fmt) {
  switch (fmt) {
  case ALTHEA_FORMAT_R8G8B8A8_UNORM: return 4;
  case ALTHEA_FORMAT_R16G16B16A16_SFLOAT: return 8;
  case ALTHEA_FORMAT_R32_SFLOAT:
  case ALTHEA_FORMAT_D32_SFLOAT: return 4;
  case ALTHEA_FORMAT_R32G32B32A32_SFLOAT: return 16;
  case ALTHEA_FORMAT_R8_UINT: return 1;
  default: return 0;
  }
}
inline uint32_t mipDim(uint32_t d, uint32_t k) { uint32_t v = d >> k; return v ? v : 1u; }
size_t chainBytes(uint32_t fmt, uint32_t w, uint32_t h, uint32_t mips) {
  size_t n = 0, bpp = bytesPerTexel(fmt);
  for (uint32_t k = 0; k < mips; ++k) n += (size_t)mipDim(w, k) * mipDim(h, k) * bpp;
  return n;
}

Resource* find(althea_cuda_ctx* ctx, uint64_t handle, ResKind kind) {
  auto it = ctx->resources.find(handle);
  if (it == ctx->resources.end() || it->second.kind != kind) return nullptr;
  return &it->second;
}

// view of (level, layer) of an image
bool levelView(const Resource& r, uint32_t level, uint32_t layer, ImgView* out) {
  if (level >= r.mips || layer >= r.layers) return false;
  size_t bpp = bytesPerTexel(r.format);
  size_t layerBytes = (r.mips == 1) ? r.pitch0 * r.h : chainBytes(r.format, r.w, r.h, r.mips);
  size_t off = layerBytes * layer;
  for (uint32_t k = 0; k < level; ++k) off += (size_t)mipDim(r.w, k) * mipDim(r.h, k) * bpp;
  out->ptr = static_cast<const char*>(r.dptr) + off;
  out->w = (int)mipDim(r.w, level);
  out->h = (int)mipDim(r.h, level);
  out->pitch = (int)((r.mips == 1) ? r.pitch0 : (size_t)out->w * bpp);
  return true;
}
bool chainView(const Resource& r, uint32_t layer, ChainView* out) {
  if (r.mips > (uint32_t)kMaxMips) return false;
  out->mips = (int)r.mips;
  for (uint32_t k = 0; k < r.mips; ++k)
    if (!levelView(r, k, layer, &out->level[k])) return false;
  return true;
}

int getImage(althea_cuda_ctx* ctx, uint64_t handle, uint32_t format, const char* what, Resource** out, bool optional = false) {
  *out = nullptr;
  if (handle == 0) {
    if (optional) return ALTHEA_OK;
    return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "%s: image handle is 0", what);
  }
  Resource* r = find(ctx, handle, ResKind::Image);
  if (!r) return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "%s: handle %llu is not a live image", what, (unsigned long long)handle);
  uint32_t f = r->format == ALTHEA_FORMAT_D32_SFLOAT ? (uint32_t)ALTHEA_FORMAT_R32_SFLOAT : r->format;
  if (format && f != format) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "%s: expected VkFormat %u, image has %u", what, format, r->format);
  *out = r;
  return ALTHEA_OK;
}

// ---- stream / semaphore / timing plumbing ------------------------------------------------------------------------
struct Scope {
  althea_cuda_ctx* ctx;
  cudaStream_t stream;
};

cudaStream_t workStream(althea_cuda_ctx* ctx, const althea_sync* sync) { return (sync && sync->cuda_stream) ? (cudaStream_t)sync->cuda_stream : ctx->stream; }
// grows one scratch allocation of the calling stream's set; contents are not preserved
int growScratchBuf(althea_cuda_ctx* ctx, void** ptr, size_t* have, size_t need, const char* what) {
  if (*have >= need) return ALTHEA_OK;
  if (*ptr) { cudaDeviceSynchronize(); cudaFree(*ptr); *ptr = nullptr; *have = 0; }
  cudaError_t e = cudaMalloc(ptr, need);
  if (e != cudaSuccess) { cudaGetLastError(); *ptr = nullptr; return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "cudaMalloc(%s %zu): %s", what, need, cudaGetErrorString(e)); }
  *have = need;
  return ALTHEA_OK;
}
int beginWork(althea_cuda_ctx* ctx, const althea_sync* sync, cudaStream_t* stream) {
  *stream = (sync && sync->cuda_stream) ? (cudaStream_t)sync->cuda_stream : ctx->stream;
  // both handles are checked before anything is enqueued: a stage never waits on a semaphore it will not signal
  if (sync && sync->signal_sem && !find(ctx, sync->signal_sem, ResKind::Semaphore))
    return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "signal_sem %llu is not a live semaphore", (unsigned long long)sync->signal_sem);
  if (sync && sync->wait_sem) {
    Resource* s = find(ctx, sync->wait_sem, ResKind::Semaphore);
    if (!s) return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "wait_sem %llu is not a live semaphore", (unsigned long long)sync->wait_sem);
    cudaExternalSemaphoreWaitParams wp;
    memset(&wp, 0, sizeof wp);
    wp.params.fence.value = sync->wait_value;
    CUDA_TRY(ctx, cudaWaitExternalSemaphoresAsync(&s->extSem, &wp, 1, *stream));
  }
  return ALTHEA_OK;
}
int endWork(althea_cuda_ctx* ctx, const althea_sync* sync, cudaStream_t stream) {
  CUDA_TRY(ctx, cudaGetLastError());
  if (sync && sync->signal_sem) {
    Resource* s = find(ctx, sync->signal_sem, ResKind::Semaphore);
    if (!s) return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "signal_sem %llu is not a live semaphore", (unsigned long long)sync->signal_sem);
    cudaExternalSemaphoreSignalParams sp;
    memset(&sp, 0, sizeof sp);
    sp.params.fence.value = sync->signal_value;
    CUDA_TRY(ctx, cudaSignalExternalSemaphoresAsync(&s->extSem, &sp, 1, stream));
  }
  return ALTHEA_OK;
}

// A stage that fails AFTER beginWork has enqueued the wait still signals its semaphore when it returns: the host's timeline must
// not be left waiting for a value that never arrives (the status and message of the failure are what the caller gets).
struct WorkGuard {
  althea_cuda_ctx* ctx;
  const althea_sync* sync;
  cudaStream_t stream;
  bool armed;
  WorkGuard(althea_cuda_ctx* c, const althea_sync* s, cudaStream_t st) : ctx(c), sync(s), stream(st), armed(true) {}
  int finish() { armed = false; return endWork(ctx, sync, stream); }
  ~WorkGuard() {
    if (!armed || !sync || !sync->signal_sem) return;
    if (Resource* s = find(ctx, sync->signal_sem, ResKind::Semaphore)) {
      cudaExternalSemaphoreSignalParams sp;
      memset(&sp, 0, sizeof sp);
      sp.params.fence.value = sync->signal_value;
      cudaSignalExternalSemaphoresAsync(&s->extSem, &sp, 1, stream);
      cudaGetLastError();
    }
  }
};

cudaEvent_t takeEvent(althea_cuda_ctx* ctx) {
  if (!ctx->eventPool.empty()) {
    cudaEvent_t e = ctx->eventPool.back();
    ctx->eventPool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

// launches `fn` as one counted, optionally event-timed kernel launch
template <typename F> void timedLaunch(althea_cuda_ctx* ctx, const char* name, cudaStream_t stream, F&& fn) {
  if (ctx->timing) {
    TimingEntry t{name, takeEvent(ctx), takeEvent(ctx)};
    cudaEventRecord(t.start, stream);
    fn();
    cudaEventRecord(t.stop, stream);
    ctx->pending.push_back(t);
  } else {
    fn();
  }
  ctx->launches += 1;
}

void drainTimings(althea_cuda_ctx* ctx) {
  for (TimingEntry& t : ctx->pending) {
    cudaEventSynchronize(t.stop);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, t.start, t.stop);
    auto it = ctx->totals.find(t.name);
    if (it == ctx->totals.end()) {
      ctx->totals[t.name] = std::make_pair((double)ms, 1u);
      ctx->totalNames.push_back(t.name);
    } else {
      it->second.first += ms;
      it->second.second += 1;
    }
    ctx->eventPool.push_back(t.start);
    ctx->eventPool.push_back(t.stop);
  }
  ctx->pending.clear();
}

// Row ranges [lo, hi) of every reflection mip that must be present so that rows [y0, y1) of the FINAL image come out the
// same as in a whole-frame run: the deferred pass fetches the mips trilinearly at the pixel's own uv (DeferredPass.frag:29-31)
// and level L is a 7-tap blur of level L-1 along y (L odd) or x (L even) (ReflectionBuffer.cpp:224-278). Conservative:
// a few rows more than strictly needed never change a value, they are only recomputed.
void bandRows(uint32_t w, uint32_t h, uint32_t mips, uint32_t y0, uint32_t y1, uint32_t* lo, uint32_t* hi) {
  auto clampRow = [](double v, uint32_t n) { return (uint32_t)(v < 0.0 ? 0.0 : (v > (double)n ? (double)n : v)); };
  for (uint32_t L = 0; L < mips; ++L) {
    const double hL = (double)mipDim(h, L);
    lo[L] = clampRow(floor((double)y0 * hL / (double)h) - 2.0, mipDim(h, L));
    hi[L] = clampRow(ceil((double)y1 * hL / (double)h) + 2.0, mipDim(h, L));
  }
  for (uint32_t L = mips - 1; L >= 1; --L) {
    const double hS = (double)mipDim(h, L - 1), hD = (double)mipDim(h, L), wD = (double)mipDim(w, L);
    const double off = (L & 1u) ? 5.176470588235294 * hS / wD : 0.0; // offsets are divided by the WIDTH on both axes (:41)
    const uint32_t sLo = clampRow(floor((double)lo[L] * hS / hD - 0.5 - off) - 1.0, mipDim(h, L - 1));
    const uint32_t sHi = clampRow(ceil((double)(hi[L] ? hi[L] - 1 : 0) * hS / hD - 0.5 + off) + 3.0, mipDim(h, L - 1));
    if (hi[L] > lo[L]) {
      if (sLo < lo[L - 1]) lo[L - 1] = sLo;
      if (sHi > hi[L - 1]) hi[L - 1] = sHi;
    }
  }
}

// projection * view in the oracle's op order (column by column, summed left to right, no contraction: this TU's host
// code is compiled with -ffp-contract=off via -Xcompiler)
void matmul44(const float* A, const float* B, float* R) {
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      float s = A[0 * 4 + r] * B[c * 4 + 0];
      s = s + A[1 * 4 + r] * B[c * 4 + 1];
      s = s + A[2 * 4 + r] * B[c * 4 + 2];
      s = s + A[3 * 4 + r] * B[c * 4 + 3];
      R[c * 4 + r] = s;
    }
}

int fillFrameParams(althea_cuda_ctx* ctx, const althea_global_uniforms* u, const althea_gbuffer* gb, const althea_ibl* ibl,
                    uint64_t lightsBuf, uint64_t shadow, uint64_t reflection, bool needPosition, FrameParams* P) {
  if (!u || !gb || !ibl) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "uniforms, gbuffer and ibl must be non-null");
  memset(P, 0, sizeof *P);
  P->g = *u;
  matmul44(u->projection, u->view, P->projView);
  {
    const float* ip = u->inverseProjection;
    const float* iv = u->inverseView;
    double dx[3], dy[3], dc[3]; // (inverseProjection * (x, y, 2, 1)).xyz = dx x + dy y + dc
    for (int r = 0; r < 3; ++r) { dx[r] = ip[0 + r]; dy[r] = ip[4 + r]; dc[r] = 2.0 * ip[8 + r] + ip[12 + r]; }
    double Wx[3], Wy[3], Wc[3];
    for (int r = 0; r < 3; ++r) {
      Wx[r] = iv[0 + r] * dx[0] + iv[4 + r] * dx[1] + iv[8 + r] * dx[2];
      Wy[r] = iv[0 + r] * dy[0] + iv[4 + r] * dy[1] + iv[8 + r] * dy[2];
      Wc[r] = iv[0 + r] * dc[0] + iv[4 + r] * dc[1] + iv[8 + r] * dc[2];
    }
    double s0 = 0, su = 0, sv = 0; // x = 2u - 1, y = 2v - 1
    for (int r = 0; r < 3; ++r) {
      const double w0 = Wc[r] - Wx[r] - Wy[r], wu = 2.0 * Wx[r], wv = 2.0 * Wy[r];
      P->ssrW0[r] = (float)w0; P->ssrWu[r] = (float)wu; P->ssrWv[r] = (float)wv;
      s0 += w0 * iv[8 + r]; su += wu * iv[8 + r]; sv += wv * iv[8 + r];
    }
    P->ssrS[0] = (float)s0; P->ssrS[1] = (float)su; P->ssrS[2] = (float)sv;
  }
  Resource *depth, *position, *normal, *albedo, *mro, *env, *pre, *irr, *lut, *refl, *sh;
  int rc;
  if ((rc = getImage(ctx, gb->normal, ALTHEA_FORMAT_R16G16B16A16_SFLOAT, "gbuffer.normal", &normal))) return rc;
  if ((rc = getImage(ctx, gb->albedo, ALTHEA_FORMAT_R8G8B8A8_UNORM, "gbuffer.albedo", &albedo))) return rc;
  if ((rc = getImage(ctx, gb->mro, ALTHEA_FORMAT_R8G8B8A8_UNORM, "gbuffer.mro", &mro))) return rc;
  // SSR needs the depth image. The lighting pass needs positions: the legacy RGBA32F attachment when the host has one
  // (mode P), else they are reconstructed from depth into engine scratch (mode D: what today's
  // GBufferResources provides, Src/DeferredRendering.cpp:42-99 has no position attachment)
  if ((rc = getImage(ctx, gb->position, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, "gbuffer.position", &position, true))) return rc;
  if ((rc = getImage(ctx, gb->depth, ALTHEA_FORMAT_R32_SFLOAT, "gbuffer.depth", &depth, needPosition && position != nullptr))) return rc;
  if ((rc = getImage(ctx, ibl->env, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, "ibl.env", &env))) return rc;
  if ((rc = getImage(ctx, ibl->prefiltered, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, "ibl.prefiltered", &pre))) return rc;
  if ((rc = getImage(ctx, ibl->irradiance, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, "ibl.irradiance", &irr))) return rc;
  if ((rc = getImage(ctx, ibl->brdf_lut, ALTHEA_FORMAT_R8G8B8A8_UNORM, "ibl.brdf_lut", &lut))) return rc;
  if ((rc = getImage(ctx, reflection, ALTHEA_FORMAT_R16G16B16A16_SFLOAT, "reflection", &refl))) return rc;
  if ((rc = getImage(ctx, shadow, ALTHEA_FORMAT_R32_SFLOAT, "shadow_cube_array", &sh, true))) return rc;
  P->W = (int)normal->w;
  P->H = (int)normal->h;
  P->y0 = 0;
  P->y1 = P->H;
  P->Wf = (float)P->W;
  P->Hf = (float)P->H;
  { // SSAO ray-depth proxy: the view ray through texel (x, y), scaled to view z = -1, is affine in (x, y) for a perspective camera
    const float* ip = u->inverseProjection;
    const float* iv = u->inverseView;
    auto dirV = [&](double nx, double ny, double* d) {
      double q[4];
      for (int r = 0; r < 4; ++r) q[r] = ip[0 + r] * nx + ip[4 + r] * ny + ip[8 + r] * 1.0 + ip[12 + r];
      for (int r = 0; r < 3; ++r) d[r] = q[r] / -q[2];
    };
    double d00[3], d10[3], d01[3];
    dirV(-1.0, -1.0, d00); dirV(1.0, -1.0, d10); dirV(-1.0, 1.0, d01);
    double vc[3], vx[3], vy[3];
    for (int r = 0; r < 3; ++r) {
      const double gx = 0.5 * (d10[r] - d00[r]), gy = 0.5 * (d01[r] - d00[r]); // per unit of ndc
      vx[r] = 2.0 * gx / P->W;
      vy[r] = 2.0 * gy / P->H;
      vc[r] = d00[r] + gx / P->W + gy / P->H; // ndc of texel (0, 0)'s centre is (1/W - 1, 1/H - 1)
    }
    for (int r = 0; r < 3; ++r) {
      P->ssaoDc[r] = (float)(iv[0 + r] * vc[0] + iv[4 + r] * vc[1] + iv[8 + r] * vc[2]);
      P->ssaoDx[r] = (float)(iv[0 + r] * vx[0] + iv[4 + r] * vx[1] + iv[8 + r] * vx[2]);
      P->ssaoDy[r] = (float)(iv[0 + r] * vy[0] + iv[4 + r] * vy[1] + iv[8 + r] * vy[2]);
      P->ssaoFwd[r] = -iv[8 + r];
      P->ssaoCam[r] = iv[12 + r];
    }
    { // world origin in model coordinates: solve [Dc Dx Dy] o = -cam (Cramer, double)
      const double m[3][3] = {{P->ssaoDc[0], P->ssaoDx[0], P->ssaoDy[0]}, {P->ssaoDc[1], P->ssaoDx[1], P->ssaoDy[1]}, {P->ssaoDc[2], P->ssaoDx[2], P->ssaoDy[2]}};
      const double rhs[3] = {-(double)P->ssaoCam[0], -(double)P->ssaoCam[1], -(double)P->ssaoCam[2]};
      auto det3 = [](const double a[3][3]) {
        return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
      };
      const double d = det3(m);
      for (int c = 0; c < 3; ++c) {
        double mc[3][3];
        for (int r = 0; r < 3; ++r)
          for (int k = 0; k < 3; ++k) mc[r][k] = k == c ? rhs[r] : m[r][k];
        P->ssaoOrigin[c] = d != 0.0 ? (float)(det3(mc) / d) : 0.0f; // a degenerate camera only costs speed: E0 flags the records
      }
    }
    auto dotf = [](const float* a, const float* b) { return (double)a[0] * b[0] + (double)a[1] * b[1] + (double)a[2] * b[2]; };
    P->ssaoGram[0] = (float)dotf(P->ssaoDc, P->ssaoDc); P->ssaoGram[1] = (float)dotf(P->ssaoDx, P->ssaoDx); P->ssaoGram[2] = (float)dotf(P->ssaoDy, P->ssaoDy);
    { // constants of the SSAO coarse sign test (frame_kernels.cu, ssao_cull_kernel)
      auto l1 = [](const float* a) { return fabs((double)a[0]) + fabs((double)a[1]) + fabs((double)a[2]); };
      double dmax = 0.0; // |D|_1 is convex in (x, y): its maximum over the image sits at a corner
      for (int c = 0; c < 4; ++c) {
        const double cx = (c & 1) ? P->W - 1 : 0, cy = (c & 2) ? P->H - 1 : 0;
        double v = 0.0;
        for (int r = 0; r < 3; ++r) v += fabs((double)P->ssaoDc[r] + (double)P->ssaoDx[r] * cx + (double)P->ssaoDy[r] * cy);
        dmax = std::max(dmax, v);
      }
      P->ssaoDmax1 = (float)(dmax * 1.0001);
      P->ssaoDmag = (float)((l1(P->ssaoDc) + l1(P->ssaoDx) * P->W + l1(P->ssaoDy) * P->H) * 1.0001);
      P->ssaoCamL1 = (float)(l1(P->ssaoCam) * 1.0001);
      P->ssaoFocalPx = (float)std::max(0.5 * P->W * fabs((double)u->projection[0]), 0.5 * P->H * fabs((double)u->projection[5]));
    }
    P->ssaoGram[3] = (float)(2.0 * dotf(P->ssaoDc, P->ssaoDx)); P->ssaoGram[4] = (float)(2.0 * dotf(P->ssaoDc, P->ssaoDy)); P->ssaoGram[5] = (float)(2.0 * dotf(P->ssaoDx, P->ssaoDy));
  }
  if (ctx->scissorY1) {
    if (ctx->scissorY1 > normal->h) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "scissor rows [%u, %u) exceed the frame height %u", ctx->scissorY0, ctx->scissorY1, normal->h);
    P->y0 = (int)ctx->scissorY0;
    P->y1 = (int)ctx->scissorY1;
  }
  auto sameSize = [&](Resource* r) { return !r || (r->w == normal->w && r->h == normal->h); };
  if (!sameSize(albedo) || !sameSize(mro) || !sameSize(depth) || !sameSize(position) || !sameSize(refl))
    return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "G-buffer and reflection images must all be %ux%u", normal->w, normal->h);
  levelView(*normal, 0, 0, &P->normal);
  levelView(*albedo, 0, 0, &P->albedo);
  levelView(*mro, 0, 0, &P->mro);
  if (depth) levelView(*depth, 0, 0, &P->depth);
  if (position) levelView(*position, 0, 0, &P->position);
  levelView(*env, 0, 0, &P->env);
  levelView(*irr, 0, 0, &P->irr);
  levelView(*lut, 0, 0, &P->lut);
  if (!chainView(*pre, 0, &P->pre) || !chainView(*refl, 0, &P->refl)) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "too many mip levels");
  if (u->lightCount < 0) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "lightCount < 0");
  if (u->lightCount > 0) {
    Resource* lb = find(ctx, lightsBuf, ResKind::Buffer);
    if (!lb) return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "lights_buf %llu is not a live buffer", (unsigned long long)lightsBuf);
    if (lb->bytes < (size_t)u->lightCount * sizeof(althea_point_light))
      return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "lights_buf holds %zu bytes, lightCount=%d needs %zu", lb->bytes, u->lightCount,
                  (size_t)u->lightCount * sizeof(althea_point_light));
    P->lights = static_cast<const float*>(lb->dptr);
    if (sh) {
      if (sh->layers < 6u * (uint32_t)u->lightCount || sh->w != sh->h)
        return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "shadow_cube_array needs square faces and >= %d layers (has %u)", 6 * u->lightCount, sh->layers);
      levelView(*sh, 0, 0, &P->shadow);
      ImgView l1;
      P->shadowLayerStride = sh->layers > 1 && levelView(*sh, 0, 1, &l1) ? (size_t)((const char*)l1.ptr - (const char*)P->shadow.ptr) : 0;
      P->shadowRes = (int)sh->w;
    }
  }
  return ALTHEA_OK;
}

} // namespace

extern "C" {

int althea_cuda_abi_version(void) { return ALTHEA_CUDA_ABI_VERSION; }

int althea_cuda_create(althea_cuda_ctx** out_ctx, int cuda_device, const uint8_t vk_device_uuid[16]) {
  if (!out_ctx) return fail(nullptr, ALTHEA_ERR_INVALID_ARGUMENT, "out_ctx is null");
  *out_ctx = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, ALTHEA_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(e));
  if (cuda_device < 0 || cuda_device >= count) return fail(nullptr, ALTHEA_ERR_INVALID_ARGUMENT, "cuda_device %d out of range [0,%d)", cuda_device, count);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, cuda_device)) != cudaSuccess) return fail(nullptr, ALTHEA_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, ALTHEA_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library carries sm_100a code only", cuda_device, prop.major, prop.minor);
  if (vk_device_uuid && memcmp(vk_device_uuid, prop.uuid.bytes, 16) != 0)
    return fail(nullptr, ALTHEA_ERR_INVALID_ARGUMENT, "CUDA device %d is not the Vulkan device (deviceUUID mismatch)", cuda_device);
  if ((e = cudaSetDevice(cuda_device)) != cudaSuccess) return fail(nullptr, ALTHEA_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  althea_cuda_ctx* ctx = new althea_cuda_ctx();
  ctx->device = cuda_device;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, ALTHEA_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
  }
  if (const char* t = getenv("ALTHEA_SSAO_DIR_TABLE")) ctx->ssaoDirTable = atoi(t) != 0; // tuning / A-B: 0 = hash the ray directions inline
  *out_ctx = ctx;
  return ALTHEA_OK;
}

void althea_cuda_destroy(althea_cuda_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  drainTimings(ctx);
  for (auto& kv : ctx->resources) {
    Resource& r = kv.second;
    if (r.extSem) cudaDestroyExternalSemaphore(r.extSem);
    if (r.extMem) cudaDestroyExternalMemory(r.extMem);
    else if (r.owned && r.dptr) cudaFree(r.dptr);
  }
  for (cudaEvent_t e : ctx->eventPool) cudaEventDestroy(e);
  for (auto& kv : ctx->scratch)
    for (void* p : {kv.second.ao, kv.second.position, kv.second.quad, kv.second.plane, kv.second.depthPad, kv.second.ssrPlane, kv.second.ssrHits})
      if (p) cudaFree(p);
  if (ctx->gatherCounter) cudaFree(ctx->gatherCounter);
  if (ctx->ssaoDirs) cudaFree(ctx->ssaoDirs);
  freeRasterScratch(ctx->raster);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* althea_cuda_last_error(const althea_cuda_ctx* ctx) { return ctx ? ctx->lastError.c_str() : g_createError.c_str(); }

int althea_cuda_set_flags(althea_cuda_ctx* ctx, uint32_t flags) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  ctx->flags = flags;
  return ALTHEA_OK;
}

int althea_cuda_set_scissor_rows(althea_cuda_ctx* ctx, uint32_t y0, uint32_t y1) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (y1 != 0 && y1 <= y0) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "scissor rows [%u, %u) are empty", y0, y1);
  ctx->scissorY0 = y1 ? y0 : 0;
  ctx->scissorY1 = y1;
  return ALTHEA_OK;
}
int althea_cuda_band_rows(uint32_t w, uint32_t h, uint32_t mips, uint32_t y0, uint32_t y1, uint32_t* out_lo, uint32_t* out_hi) {
  if (!w || !h || !mips || mips > (uint32_t)kMaxMips || y1 <= y0 || y1 > h || !out_lo || !out_hi) return ALTHEA_ERR_INVALID_ARGUMENT;
  bandRows(w, h, mips, y0, y1, out_lo, out_hi);
  return ALTHEA_OK;
}

int althea_cuda_enable_timing(althea_cuda_ctx* ctx, int enable) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  ctx->timing = enable != 0;
  return ALTHEA_OK;
}
int althea_cuda_get_timings(althea_cuda_ctx* ctx, const char** names, float* total_ms, uint32_t* launches, int cap) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  drainTimings(ctx);
  int n = 0;
  for (const char* name : ctx->totalNames) {
    if (n >= cap) break;
    auto& t = ctx->totals[name];
    if (names) names[n] = name;
    if (total_ms) total_ms[n] = (float)t.first;
    if (launches) launches[n] = t.second;
    ++n;
  }
  return n;
}
int althea_cuda_reset_timings(althea_cuda_ctx* ctx) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  drainTimings(ctx);
  ctx->totals.clear();
  ctx->totalNames.clear();
  return ALTHEA_OK;
}
uint64_t althea_cuda_launch_count(const althea_cuda_ctx* ctx) { return ctx ? ctx->launches : 0; }

size_t althea_cuda_image_bytes(uint32_t vk_format, uint32_t w, uint32_t h, uint32_t mips, uint32_t layers) {
  if (!bytesPerTexel(vk_format) || !w || !h || !mips || !layers) return 0;
  return chainBytes(vk_format, w, h, mips) * layers;
}

static int registerImage(althea_cuda_ctx* ctx, Resource& r, size_t pitch, size_t available, uint64_t* out_handle) {
  size_t bpp = bytesPerTexel(r.format);
  if (!bpp) return fail(ctx, ALTHEA_ERR_UNSUPPORTED, "VkFormat %u is not on the deferred path", r.format);
  if (!r.w || !r.h || !r.mips || !r.layers) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "image extent/mips/layers must be non-zero");
  if (r.mips > (uint32_t)kMaxMips) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "mips %u > %d", r.mips, kMaxMips);
  size_t tight = (size_t)r.w * bpp;
  if (pitch == 0) pitch = tight;
  if (pitch < tight || (pitch % bpp) != 0) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "pitch %zu invalid for width %u x %zu B", pitch, r.w, bpp);
  if (r.mips > 1 && pitch != tight) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "mip chains must be tightly packed (pitch %zu != %zu)", pitch, tight);
  r.pitch0 = pitch;
  r.bytes = (r.mips == 1 ? pitch * r.h : chainBytes(r.format, r.w, r.h, r.mips)) * r.layers;
  if (available && available < r.bytes) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "image needs %zu bytes, only %zu provided", r.bytes, available);
  if (((uintptr_t)r.dptr % 16) != 0 || (pitch % (bpp < 16 ? bpp : 16)) != 0)
    return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "image base must be 16-byte aligned");
  r.kind = ResKind::Image;
  uint64_t h = ctx->nextHandle++;
  ctx->resources[h] = r;
  *out_handle = h;
  return ALTHEA_OK;
}

int althea_cuda_import_image(althea_cuda_ctx* ctx, int fd, uint64_t alloc_size, uint64_t offset, uint32_t vk_format, uint32_t w, uint32_t h,
                             uint32_t mips, uint32_t layers, uint32_t flags, uint64_t pitch, uint64_t* out_handle) {
  if (!ctx || !out_handle) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (fd < 0) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "import_image: fd %d is not a file descriptor", fd);
  if (alloc_size == 0 || offset >= alloc_size) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "import_image: offset %llu does not lie inside the allocation of %llu bytes", (unsigned long long)offset, (unsigned long long)alloc_size);
  if (!w || !h || !mips || !layers) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "import_image: empty extent %ux%u, %u mips, %u layers", w, h, mips, layers);
  if (flags & ALTHEA_IMAGE_OPTIMAL_TILING)
    return fail(ctx, ALTHEA_ERR_UNSUPPORTED,
                "optimal-tiled images cannot be mapped as linear memory; create the image with VK_IMAGE_TILING_LINEAR or copy it to an "
                "exportable buffer (INTEGRATION.md)");
  cudaExternalMemoryHandleDesc hd;
  memset(&hd, 0, sizeof hd);
  hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
  hd.handle.fd = fd;
  hd.size = alloc_size;
  Resource r;
  CUDA_TRY(ctx, cudaImportExternalMemory(&r.extMem, &hd)); // on success CUDA owns the fd
  cudaExternalMemoryBufferDesc bd;
  memset(&bd, 0, sizeof bd);
  bd.offset = offset;
  bd.size = alloc_size - offset;
  cudaError_t e = cudaExternalMemoryGetMappedBuffer(&r.dptr, r.extMem, &bd);
  if (e != cudaSuccess) {
    cudaDestroyExternalMemory(r.extMem);
    return fail(ctx, ALTHEA_ERR_CUDA, "cudaExternalMemoryGetMappedBuffer: %s", cudaGetErrorString(e));
  }
  r.format = vk_format; r.w = w; r.h = h; r.mips = mips; r.layers = layers;
  int rc = registerImage(ctx, r, (size_t)pitch, (size_t)(alloc_size - offset), out_handle);
  if (rc != ALTHEA_OK) cudaDestroyExternalMemory(r.extMem);
  return rc;
}

int althea_cuda_import_buffer(althea_cuda_ctx* ctx, int fd, uint64_t size, uint64_t offset, uint64_t* out_handle) {
  if (!ctx || !out_handle) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (fd < 0) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "import_buffer: fd %d is not a file descriptor", fd);
  if (size == 0 || offset >= size) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "import_buffer: offset %llu does not lie inside the allocation of %llu bytes", (unsigned long long)offset, (unsigned long long)size);
  cudaExternalMemoryHandleDesc hd;
  memset(&hd, 0, sizeof hd);
  hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
  hd.handle.fd = fd;
  hd.size = size;
  Resource r;
  CUDA_TRY(ctx, cudaImportExternalMemory(&r.extMem, &hd));
  cudaExternalMemoryBufferDesc bd;
  memset(&bd, 0, sizeof bd);
  bd.offset = offset;
  bd.size = size - offset;
  cudaError_t e = cudaExternalMemoryGetMappedBuffer(&r.dptr, r.extMem, &bd);
  if (e != cudaSuccess) {
    cudaDestroyExternalMemory(r.extMem);
    return fail(ctx, ALTHEA_ERR_CUDA, "cudaExternalMemoryGetMappedBuffer: %s", cudaGetErrorString(e));
  }
  r.kind = ResKind::Buffer;
  r.bytes = (size_t)(size - offset);
  uint64_t h = ctx->nextHandle++;
  ctx->resources[h] = r;
  *out_handle = h;
  return ALTHEA_OK;
}

int althea_cuda_import_semaphore(althea_cuda_ctx* ctx, int fd, int is_timeline, uint64_t* out_handle) {
  if (!ctx || !out_handle) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (fd < 0) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "import_semaphore: fd %d is not a file descriptor", fd);
  cudaExternalSemaphoreHandleDesc sd;
  memset(&sd, 0, sizeof sd);
  sd.type = is_timeline ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
  sd.handle.fd = fd;
  Resource r;
  CUDA_TRY(ctx, cudaImportExternalSemaphore(&r.extSem, &sd));
  r.kind = ResKind::Semaphore;
  r.timeline = is_timeline != 0;
  uint64_t h = ctx->nextHandle++;
  ctx->resources[h] = r;
  *out_handle = h;
  return ALTHEA_OK;
}

int althea_cuda_wrap_linear_image(althea_cuda_ctx* ctx, void* dptr, size_t pitch, uint32_t vk_format, uint32_t w, uint32_t h, uint32_t mips,
                                  uint32_t layers, uint64_t* out_handle) {
  if (!ctx || !out_handle) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (!dptr) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "dptr is null");
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, dptr);
  if (e != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
    cudaGetLastError();
    return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "dptr %p is not device memory", dptr);
  }
  Resource r;
  r.dptr = dptr;
  r.format = vk_format; r.w = w; r.h = h; r.mips = mips; r.layers = layers;
  return registerImage(ctx, r, pitch, 0, out_handle);
}

int althea_cuda_wrap_buffer(althea_cuda_ctx* ctx, void* dptr, size_t size, uint64_t* out_handle) {
  if (!ctx || !out_handle) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (!dptr || !size) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "dptr/size is null");
  Resource r;
  r.kind = ResKind::Buffer;
  r.dptr = dptr;
  r.bytes = size;
  uint64_t h = ctx->nextHandle++;
  ctx->resources[h] = r;
  *out_handle = h;
  return ALTHEA_OK;
}

int althea_cuda_create_image(althea_cuda_ctx* ctx, uint32_t vk_format, uint32_t w, uint32_t h, uint32_t mips, uint32_t layers, uint64_t* out_handle) {
  if (!ctx || !out_handle) return ALTHEA_ERR_INVALID_ARGUMENT;
  size_t bytes = althea_cuda_image_bytes(vk_format, w, h, mips, layers);
  if (!bytes) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "bad image description (format %u, %ux%u, mips %u, layers %u)", vk_format, w, h, mips, layers);
  Resource r;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMalloc(&r.dptr, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  }
  r.owned = true;
  r.format = vk_format; r.w = w; r.h = h; r.mips = mips; r.layers = layers;
  int rc = registerImage(ctx, r, 0, bytes, out_handle);
  if (rc != ALTHEA_OK) cudaFree(r.dptr);
  return rc;
}

int althea_cuda_create_buffer(althea_cuda_ctx* ctx, size_t size, uint64_t* out_handle) {
  if (!ctx || !out_handle || !size) return ALTHEA_ERR_INVALID_ARGUMENT;
  Resource r;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMalloc(&r.dptr, size);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu): %s", size, cudaGetErrorString(e));
  }
  r.kind = ResKind::Buffer;
  r.owned = true;
  r.bytes = size;
  uint64_t h = ctx->nextHandle++;
  ctx->resources[h] = r;
  *out_handle = h;
  return ALTHEA_OK;
}

static Resource* findMem(althea_cuda_ctx* ctx, uint64_t handle) {
  auto it = ctx->resources.find(handle);
  if (it == ctx->resources.end() || it->second.kind == ResKind::Semaphore) return nullptr;
  return &it->second;
}

int althea_cuda_upload(althea_cuda_ctx* ctx, uint64_t handle, const void* host, size_t bytes, void* stream) {
  if (!ctx || !host) return ALTHEA_ERR_INVALID_ARGUMENT;
  Resource* r = findMem(ctx, handle);
  if (!r) return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "upload: handle %llu is not live memory", (unsigned long long)handle);
  if (bytes > r->bytes) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "upload of %zu bytes into a %zu-byte resource", bytes, r->bytes);
  CUDA_TRY(ctx, cudaMemcpyAsync(r->dptr, host, bytes, cudaMemcpyHostToDevice, stream ? (cudaStream_t)stream : ctx->stream));
  return ALTHEA_OK;
}
int althea_cuda_download(althea_cuda_ctx* ctx, uint64_t handle, void* host, size_t bytes, void* stream) {
  if (!ctx || !host) return ALTHEA_ERR_INVALID_ARGUMENT;
  Resource* r = findMem(ctx, handle);
  if (!r) return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "download: handle %llu is not live memory", (unsigned long long)handle);
  if (bytes > r->bytes) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "download of %zu bytes from a %zu-byte resource", bytes, r->bytes);
  CUDA_TRY(ctx, cudaMemcpyAsync(host, r->dptr, bytes, cudaMemcpyDeviceToHost, stream ? (cudaStream_t)stream : ctx->stream));
  return ALTHEA_OK;
}
int althea_cuda_device_pointer(althea_cuda_ctx* ctx, uint64_t handle, void** out_dptr, size_t* out_bytes) {
  if (!ctx || !out_dptr) return ALTHEA_ERR_INVALID_ARGUMENT;
  Resource* r = findMem(ctx, handle);
  if (!r) return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "handle %llu is not live memory", (unsigned long long)handle);
  *out_dptr = r->dptr;
  if (out_bytes) *out_bytes = r->bytes;
  return ALTHEA_OK;
}

int althea_cuda_release(althea_cuda_ctx* ctx, uint64_t handle) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  auto it = ctx->resources.find(handle);
  if (it == ctx->resources.end()) return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "release: handle %llu is not live", (unsigned long long)handle);
  Resource& r = it->second;
  if (r.extSem) cudaDestroyExternalSemaphore(r.extSem);
  if (r.extMem) cudaDestroyExternalMemory(r.extMem);
  else if (r.owned && r.dptr) {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(r.dptr);
  }
  ctx->resources.erase(it);
  return ALTHEA_OK;
}

int althea_cuda_synchronize(althea_cuda_ctx* ctx, void* stream) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  CUDA_TRY(ctx, cudaStreamSynchronize(stream ? (cudaStream_t)stream : ctx->stream));
  return ALTHEA_OK;
}

// ---- per-frame stages -----------------------------------------------------------------------------------------------
int althea_cuda_ssr_capture(althea_cuda_ctx* ctx, const althea_global_uniforms* uniforms, const althea_gbuffer* gbuffer, const althea_ibl* ibl,
                            uint64_t lights_buf, uint64_t shadow_cube_array, uint64_t reflection, const althea_sync* sync) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  FrameParams P;
  int rc = fillFrameParams(ctx, uniforms, gbuffer, ibl, lights_buf, shadow_cube_array, reflection, /*needPosition=*/false, &P);
  if (rc) return rc;
  if (ctx->scissorY1 && !(ctx->flags & ALTHEA_CTX_BAND_EXCHANGE_HALO)) { // the glossy mips of the band need a halo of mip-0 rows: recomputed locally
    Resource* refl = find(ctx, reflection, ResKind::Image);                 // (DESIGN.md 6), unless the host exchanges them between the ranks
    uint32_t lo[kMaxMips], hi[kMaxMips];
    bandRows(refl->w, refl->h, refl->mips, ctx->scissorY0, ctx->scissorY1, lo, hi);
    P.y0 = (int)lo[0];
    P.y1 = (int)hi[0];
  }
  althea_cuda_ctx::Scratch& S = ctx->scratch[workStream(ctx, sync)];
  { // the padded depth covers the whole frame even under a scissor: a ray may leave the band. The fast build's plain march reads
    // it as one 16-byte record per bilinear footprint instead ((W + 1) x (H + 1) records in the same scratch allocation)
    const bool parityMath = ctx->flags & ALTHEA_CTX_PARITY_MATH;
    const bool quads = !(ctx->flags & ALTHEA_CTX_SSR_PLANE_SKIP) && (parityMath ? althea_parity::ssr_march_reads_depth_quads() : althea_fast::ssr_march_reads_depth_quads());
    const size_t need = quads ? ((size_t)P.W + 1) * ((size_t)P.H + 1) * sizeof(float4) : ((size_t)P.W + 2) * ((size_t)P.H + 2) * sizeof(float);
    if ((rc = growScratchBuf(ctx, &S.depthPad, &S.depthPadBytes, need, "ssr padded depth"))) return rc;
    P.depthPadRow = P.W + 2;
    P.depthPad = static_cast<const float*>(S.depthPad);
    P.depthPadOrigin = P.depthPad + P.depthPadRow + 1;
    P.depthQuadOrigin = nullptr;
    P.depthQuadRow = P.W + 1;
    if (quads) P.depthQuadOrigin = static_cast<const float4*>(S.depthPad) + ((size_t)P.depthQuadRow + 1);
  }
  P.ssrHits = nullptr;
  P.ssrHitCount = nullptr;
  if (!(ctx->flags & ALTHEA_CTX_SSR_PLANE_SKIP)) { // the march's hit list (the plane-skip variant shades in place)
    const size_t cap = (size_t)P.W * (size_t)(P.y1 - P.y0);
    if ((rc = growScratchBuf(ctx, &S.ssrHits, &S.ssrHitsBytes, 256 + cap * 12 * sizeof(float), "ssr hit list"))) return rc;
    P.ssrHitCount = static_cast<unsigned*>(S.ssrHits);
    P.ssrHits = reinterpret_cast<float*>(static_cast<char*>(S.ssrHits) + 256);
    P.ssrHitCap = (unsigned)cap;
  }
  P.ssrPlanes = nullptr;
  if (ctx->flags & ALTHEA_CTX_SSR_PLANE_SKIP) { // sign test over plane records of the depth buffer: the smallest blocks that cover the frame
    int shift = 3;
    while (((P.W + (1 << shift) - 1) >> shift) > kSsrPlaneStride - 1 || ((P.H + (1 << shift) - 1) >> shift) > kSsrPlaneRows - 1) ++shift;
    size_t have = S.ssrPlane ? (size_t)kSsrPlaneStride * kSsrPlaneRows * 16 : 0;
    if ((rc = growScratchBuf(ctx, &S.ssrPlane, &have, (size_t)kSsrPlaneStride * kSsrPlaneRows * 16, "ssr plane records"))) return rc;
    P.ssrPlanes = static_cast<const float4*>(S.ssrPlane);
    if (const char* e = getenv("ALTHEA_SSR_PLANE_SHIFT")) shift = std::max(shift, atoi(e)); // tuning: coarser blocks
    P.ssrPlaneShift = shift;
  }
  if (P.ssrPlanes && (ctx->flags & ALTHEA_CTX_SSAO_COUNT_TAPS)) { // diagnostics: the skip kernel's tap counters (read with althea_cuda_diag_ssao_cull)
    if (!ctx->gatherCounter) {
      cudaError_t e = cudaMalloc(&ctx->gatherCounter, 80 * sizeof(unsigned long long));
      if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "cudaMalloc(gather counter): %s", cudaGetErrorString(e)); }
    }
    P.gatherCounter = ctx->gatherCounter;
  }
  cudaStream_t stream;
  if ((rc = beginWork(ctx, sync, &stream))) return rc;
  if (P.gatherCounter) cudaMemsetAsync(P.gatherCounter, 0, 80 * sizeof(unsigned long long), stream);
  const bool parity = ctx->flags & ALTHEA_CTX_PARITY_MATH;
  timedLaunch(ctx, "ssr_depth_pad", stream, [&] { parity ? althea_parity::launch_ssr_depth_pad(P, stream) : althea_fast::launch_ssr_depth_pad(P, stream); });
  if (P.ssrPlanes) timedLaunch(ctx, "ssr_planes", stream, [&] { parity ? althea_parity::launch_ssr_planes(P, stream) : althea_fast::launch_ssr_planes(P, stream); });
  timedLaunch(ctx, "ssr_capture", stream, [&] { parity ? althea_parity::launch_ssr_capture(P, stream) : althea_fast::launch_ssr_capture(P, stream); });
  if (P.ssrHits) timedLaunch(ctx, "ssr_shade_hits", stream, [&] { parity ? althea_parity::launch_ssr_shade_hits(P, stream) : althea_fast::launch_ssr_shade_hits(P, stream); });
  return endWork(ctx, sync, stream);
}

int althea_cuda_glossy_convolve(althea_cuda_ctx* ctx, uint64_t reflection, const althea_sync* sync) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  Resource* refl;
  int rc = getImage(ctx, reflection, ALTHEA_FORMAT_R16G16B16A16_SFLOAT, "reflection", &refl);
  if (rc) return rc;
  // everything that can fail is checked before beginWork enqueues the wait on the caller's semaphore: an error return after it
  // would leave signal_sem unsignalled and the Vulkan side waiting for a timeline value that never arrives
  uint32_t lo[kMaxMips], hi[kMaxMips];
  if (ctx->scissorY1) {
    if (ctx->scissorY1 > refl->h) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "scissor rows exceed the frame height %u", refl->h);
    bandRows(refl->w, refl->h, refl->mips, ctx->scissorY0, ctx->scissorY1, lo, hi);
  } else {
    for (uint32_t level = 0; level < refl->mips; ++level) { lo[level] = 0; hi[level] = mipDim(refl->h, level); }
  }
  cudaStream_t stream;
  if ((rc = beginWork(ctx, sync, &stream))) return rc;
  const bool parity = ctx->flags & ALTHEA_CTX_PARITY_MATH;
  for (uint32_t level = 1; level < refl->mips; ++level) { // ReflectionBuffer.cpp:224-278
    if (hi[level] <= lo[level]) continue;
    ConvolveParams C;
    levelView(*refl, level - 1, 0, &C.src);
    levelView(*refl, level, 0, &C.dst);
    C.vertical = (int)(level & 1u);
    C.y0 = (int)lo[level];
    C.y1 = (int)hi[level];
    timedLaunch(ctx, "glossy_convolve", stream, [&] { parity ? althea_parity::launch_glossy_convolve(C, stream) : althea_fast::launch_glossy_convolve(C, stream); });
  }
  return endWork(ctx, sync, stream);
}

int althea_cuda_deferred_shade(althea_cuda_ctx* ctx, const althea_global_uniforms* uniforms, const althea_gbuffer* gbuffer, const althea_ibl* ibl,
                               uint64_t lights_buf, uint64_t shadow_cube_array, uint64_t reflection, uint64_t out_color, uint64_t ao_counts,
                               uint32_t flags, const althea_sync* sync) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  FrameParams P;
  int rc = fillFrameParams(ctx, uniforms, gbuffer, ibl, lights_buf, shadow_cube_array, reflection, /*needPosition=*/true, &P);
  if (rc) return rc;
  Resource *out = nullptr, *ao;
  const bool aoOnly = flags & ALTHEA_SHADE_AO_ONLY;
  if (aoOnly && (flags & (ALTHEA_SHADE_NO_SSAO | ALTHEA_SHADE_AO_FROM_IMAGE))) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "AO_ONLY excludes NO_SSAO and AO_FROM_IMAGE");
  if (!aoOnly) {
    if ((rc = getImage(ctx, out_color, 0, "out_color", &out))) return rc;
    if (out->format != ALTHEA_FORMAT_R16G16B16A16_SFLOAT && out->format != ALTHEA_FORMAT_R32G32B32A32_SFLOAT)
      return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "out_color must be RGBA16F or RGBA32F (has VkFormat %u)", out->format);
    if ((int)out->w != P.W || (int)out->h != P.H) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "out_color must be %dx%d", P.W, P.H);
    levelView(*out, 0, 0, &P.out);
    P.outIsF32 = out->format == ALTHEA_FORMAT_R32G32B32A32_SFLOAT;
  }
  P.flags = flags;
  if ((rc = getImage(ctx, ao_counts, ALTHEA_FORMAT_R8_UINT, "ao_counts", &ao, true))) return rc;
  if ((flags & (ALTHEA_SHADE_AO_FROM_IMAGE | ALTHEA_SHADE_AO_ONLY)) && !ao) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "AO_FROM_IMAGE / AO_ONLY need ao_counts");
  const bool needAo = !(flags & ALTHEA_SHADE_NO_SSAO);
  althea_cuda_ctx::Scratch& S = ctx->scratch[workStream(ctx, sync)];
  if (needAo) {
    if (ao) {
      if ((int)ao->w != P.W || (int)ao->h != P.H) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "ao_counts must be %dx%d", P.W, P.H);
      levelView(*ao, 0, 0, &P.ao);
    } else {
      if ((rc = growScratchBuf(ctx, &S.ao, &S.aoBytes, (size_t)P.W * P.H, "ao scratch"))) return rc;
      P.ao.ptr = S.ao; P.ao.w = P.W; P.ao.h = P.H; P.ao.pitch = P.W;
    }
  }
  const bool reconstruct = P.position.ptr == nullptr; // mode D
  if (reconstruct) {
    if ((rc = growScratchBuf(ctx, &S.position, &S.positionBytes, (size_t)P.W * P.H * 16, "position scratch"))) return rc;
    P.position.ptr = S.position; P.position.w = P.W; P.position.h = P.H; P.position.pitch = P.W * 16;
  }
  const bool parity = ctx->flags & ALTHEA_CTX_PARITY_MATH;
  const bool computeAo = needAo && !(flags & ALTHEA_SHADE_AO_FROM_IMAGE);
  const bool exactTaps = ctx->flags & ALTHEA_CTX_SSAO_EXACT_TAPS;
  if (computeAo && !exactTaps) {
    if ((rc = growScratchBuf(ctx, &S.quad, &S.quadBytes, ((size_t)P.W + 1) * 32 * ((size_t)P.H + 1), "ssao quad scratch"))) return rc;
    P.quads = S.quad;
    P.quadRow = P.W + 1;
    P.quadKind = (ctx->flags & ALTHEA_CTX_SSAO_RAY_DEPTH_PROXY) ? 1 : 0;
    P.quadPitch = ((size_t)P.W + 1) * (P.quadKind ? 16 : 32);
    P.quadsOrigin = static_cast<const char*>(S.quad) + ((size_t)P.quadRow + 1) * (P.quadKind ? 16 : 32);
  }
  // the coarse sign test needs a perspective camera's G-buffer model (it degrades to the march tile by tile when the
  // positions do not fit it) and the position records; the ray-depth records keep their own march
  const bool cull = computeAo && !exactTaps && !(ctx->flags & (ALTHEA_CTX_SSAO_NO_CULL | ALTHEA_CTX_SSAO_RAY_DEPTH_PROXY));
  P.ssaoTileList = nullptr;
  P.ssaoPlaneStats = nullptr;
  P.ssaoRecip = nullptr;
  if (cull) {
    size_t need = 0, off[3];
    for (int l = 0; l < 3; ++l) {
      const int S = 8 << l;
      P.ssaoPlaneNx[l] = (P.W + S - 1) / S;
      P.ssaoPlaneNy[l] = (P.H + S - 1) / S;
      P.ssaoPlaneRow[l] = P.ssaoPlaneNx[l] + 2 * kSsaoPlanePad + 1;
      off[l] = need;
      need += (size_t)P.ssaoPlaneRow[l] * (P.ssaoPlaneNy[l] + 2 * kSsaoPlanePad + 1) * 16;
    }
    const size_t listOff = need; // two words of record statistics, then the tile list
    need += ((size_t)((P.W + 15) / 16) * ((P.H + 15) / 16) + 3) * sizeof(unsigned);
    const size_t recipOff = (need + 255) & ~(size_t)255;
    need = recipOff + (size_t)P.W * P.H * sizeof(float);
    if ((rc = growScratchBuf(ctx, &S.plane, &S.planeBytes, need, "ssao plane scratch"))) return rc;
    for (int l = 0; l < 3; ++l)
      P.ssaoPlanes[l] = reinterpret_cast<const float4*>(static_cast<const char*>(S.plane) + off[l]) + ((size_t)kSsaoPlanePad * P.ssaoPlaneRow[l] + kSsaoPlanePad);
    P.ssaoPlaneStats = reinterpret_cast<unsigned*>(static_cast<char*>(S.plane) + listOff);
    P.ssaoTileList = P.ssaoPlaneStats + 2;
    P.ssaoRecip = reinterpret_cast<float*>(static_cast<char*>(S.plane) + recipOff);
  }
  if (computeAo && !exactTaps && (ctx->flags & ALTHEA_CTX_SSAO_COUNT_TAPS)) {
    if (!ctx->gatherCounter) {
      cudaError_t e = cudaMalloc(&ctx->gatherCounter, 80 * sizeof(unsigned long long));
      if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "cudaMalloc(gather counter): %s", cudaGetErrorString(e)); }
    }
    P.gatherCounter = ctx->gatherCounter;
  }
  P.ssaoDirs = nullptr;
  P.ssaoDirRow = 0;
  if (cull && ctx->ssaoDirTable) {
    // seeds reach (W - 1 + 3 * 23, H - 1 + 3 * 23); rows padded to whole 128-byte lines. Built on the ctx's own stream and waited
    // for (once per frame size): the table is shared by every stream a host renders on, and a rebuild must not free entries a
    // frame in flight still reads
    const int row = (P.W + 72 + 7) & ~7, rows = P.H + 72;
    if (row > ctx->ssaoDirRow || rows > ctx->ssaoDirRows || parity != ctx->ssaoDirsParity) { // (the two builds normalise differently)
      const int nrow = std::max(row, ctx->ssaoDirRow), nrows = std::max(rows, ctx->ssaoDirRows);
      if (ctx->ssaoDirs) { // a frame in flight on another stream may still read the old table
        CUDA_TRY(ctx, cudaDeviceSynchronize());
        cudaFree(ctx->ssaoDirs);
      }
      ctx->ssaoDirs = nullptr; ctx->ssaoDirRow = ctx->ssaoDirRows = 0;
      cudaError_t e = cudaMalloc(&ctx->ssaoDirs, (size_t)nrow * nrows * 16);
      if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "cudaMalloc(ssao direction table, %zu bytes): %s", (size_t)nrow * nrows * 16, cudaGetErrorString(e)); }
      parity ? althea_parity::launch_ssao_dirs(static_cast<float4*>(ctx->ssaoDirs), nrow, nrows, ctx->stream)
             : althea_fast::launch_ssao_dirs(static_cast<float4*>(ctx->ssaoDirs), nrow, nrows, ctx->stream);
      ctx->launches += 1;
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      ctx->ssaoDirRow = nrow; ctx->ssaoDirRows = nrows; ctx->ssaoDirsParity = parity;
    }
    P.ssaoDirs = static_cast<const float4*>(ctx->ssaoDirs);
    P.ssaoDirRow = ctx->ssaoDirRow;
  }
  cudaStream_t stream;
  if ((rc = beginWork(ctx, sync, &stream))) return rc;
  if (P.gatherCounter) cudaMemsetAsync(P.gatherCounter, 0, 80 * sizeof(unsigned long long), stream);
  if (reconstruct)
    timedLaunch(ctx, "reconstruct_position", stream, [&] { parity ? althea_parity::launch_reconstruct_position(P, stream) : althea_fast::launch_reconstruct_position(P, stream); });
  if (computeAo) {
    if (exactTaps) {
      timedLaunch(ctx, "ssao_exact", stream, [&] { parity ? althea_parity::launch_ssao_exact(P, stream) : althea_fast::launch_ssao_exact(P, stream); });
    } else {
      timedLaunch(ctx, "ssao_quads", stream, [&] { parity ? althea_parity::launch_ssao_quads(P, stream) : althea_fast::launch_ssao_quads(P, stream); });
      if (cull) {
        for (bool coarsest : {true, false})
          timedLaunch(ctx, "ssao_planes", stream, [&] { parity ? althea_parity::launch_ssao_planes(P, stream, coarsest) : althea_fast::launch_ssao_planes(P, stream, coarsest); });
        timedLaunch(ctx, "ssao_cull", stream, [&] { parity ? althea_parity::launch_ssao_cull(P, stream) : althea_fast::launch_ssao_cull(P, stream); });
      }
      timedLaunch(ctx, "ssao", stream, [&] { parity ? althea_parity::launch_ssao(P, stream) : althea_fast::launch_ssao(P, stream); });
    }
  }
  if (!aoOnly)
    timedLaunch(ctx, "deferred_shade", stream, [&] { parity ? althea_parity::launch_deferred_shade(P, stream) : althea_fast::launch_deferred_shade(P, stream); });
  return endWork(ctx, sync, stream);
}

// ---- IBL precompute -------------------------------------------------------------------------------------------------
int althea_cuda_generate_mips(althea_cuda_ctx* ctx, uint64_t image, const althea_sync* sync) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  Resource* img;
  int rc = getImage(ctx, image, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, "image", &img);
  if (rc) return rc;
  cudaStream_t stream;
  if ((rc = beginWork(ctx, sync, &stream))) return rc;
  for (uint32_t layer = 0; layer < img->layers; ++layer)
    for (uint32_t level = 1; level < img->mips; ++level) {
      MipGenParams M;
      levelView(*img, level - 1, layer, &M.src);
      levelView(*img, level, layer, &M.dst);
      M.channels = 4;
      timedLaunch(ctx, "mip_downsample", stream, [&] { althea_iblk::launch_mip_downsample(M, stream); });
    }
  return endWork(ctx, sync, stream);
}

int althea_cuda_ibl_precompute(althea_cuda_ctx* ctx, uint64_t env_with_mips, const althea_ibl_precompute_desc* desc, uint64_t out_irradiance,
                               uint64_t out_prefiltered, const althea_sync* sync) {
  if (!ctx || !desc) return ALTHEA_ERR_INVALID_ARGUMENT;
  Resource *env, *irr, *pre;
  int rc;
  if ((rc = getImage(ctx, env_with_mips, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, "env_with_mips", &env))) return rc;
  if ((rc = getImage(ctx, out_irradiance, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, "out_irradiance", &irr, true))) return rc;
  if ((rc = getImage(ctx, out_prefiltered, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, "out_prefiltered", &pre, true))) return rc;
  if (desc->layout > ALTHEA_IBL_LAYOUT_CUBE || desc->sequence > ALTHEA_IBL_SEQ_HAMMERSLEY) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "bad layout/sequence");
  const bool cube = desc->layout == ALTHEA_IBL_LAYOUT_CUBE;
  if (cube && ((irr && irr->layers != 6) || (pre && pre->layers != 6))) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "cube outputs need 6 layers");
  IblParams I;
  memset(&I, 0, sizeof I);
  if (!chainView(*env, 0, &I.env)) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "env has too many mips");
  I.layout = (int)desc->layout;
  I.sequence = (int)desc->sequence;
  I.numSamples = desc->prefilter_samples ? (int)desc->prefilter_samples : 10000;
  I.thetaSamples = desc->theta_samples ? (int)desc->theta_samples : 300;
  // GenIrradianceMap.comp:119-124 with width/height = the env map's (ImageBasedLighting.cpp:328-333)
  I.phiSamples = (int)((float)env->h * (float)I.thetaSamples / (float)env->w);
  I.mip = log2f((float)env->w / (float)I.thetaSamples);
  cudaStream_t stream;
  if ((rc = beginWork(ctx, sync, &stream))) return rc;
  const uint32_t layers = cube ? 6u : 1u;
  I.faces = (int)layers; // all faces of a level in one launch (blockIdx.y): small cube levels do not fill the GPU face by face
  if (irr) {
    for (uint32_t face = 0; face < layers; ++face) levelView(*irr, 0, face, &I.out[face]);
    timedLaunch(ctx, "ibl_irradiance", stream, [&] { althea_iblk::launch_ibl_irradiance(I, stream); });
  }
  if (pre)
    for (uint32_t level = 0; level < pre->mips; ++level) {
      // equirect (reference): roughness = i/4 for the 5 images (ImageBasedLighting.cpp:384); cube: k/(n-1)
      I.roughness = cube ? (pre->mips > 1 ? (float)level / (float)(pre->mips - 1) : 0.0f) : (float)level / 4.0f;
      for (uint32_t face = 0; face < layers; ++face) levelView(*pre, level, face, &I.out[face]);
      timedLaunch(ctx, "ibl_prefilter", stream, [&] { althea_iblk::launch_ibl_prefilter(I, stream); });
    }
  return endWork(ctx, sync, stream);
}

int althea_cuda_brdf_lut(althea_cuda_ctx* ctx, uint32_t samples, uint64_t out_lut, const althea_sync* sync) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  Resource* lut;
  int rc = getImage(ctx, out_lut, 0, "out_lut", &lut);
  if (rc) return rc;
  if (lut->format != ALTHEA_FORMAT_R8G8B8A8_UNORM && lut->format != ALTHEA_FORMAT_R32G32B32A32_SFLOAT)
    return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "out_lut must be RGBA8 or RGBA32F");
  if (lut->w != lut->h) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "out_lut must be square");
  LutParams L;
  levelView(*lut, 0, 0, &L.out);
  L.outIsF32 = lut->format == ALTHEA_FORMAT_R32G32B32A32_SFLOAT;
  L.samples = samples ? (int)samples : 1024;
  cudaStream_t stream;
  if ((rc = beginWork(ctx, sync, &stream))) return rc;
  timedLaunch(ctx, "brdf_lut", stream, [&] { althea_iblk::launch_brdf_lut(L, stream); });
  return endWork(ctx, sync, stream);
}

} // extern "C"

// ---- diagnostics: the machine's rate for the SSAO march's access pattern --------------------------------------------------
// Every lane reads 32-byte records (one 256-bit load each, as ssao_kernel does) at hash-random positions inside a (2R)^2 window
// around its 16 x 16 tile of a (w+1) x (h+1) record grid; the measured records/s is the ceiling bench.py holds the SSAO march
// against (the stage is bound by the L1 data pipe: one wavefront per distinct 128-byte line).
namespace {
__device__ __forceinline__ uint32_t diagHash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void __launch_bounds__(256, 4) diag_gather_kernel(const void* recs, uint32_t* sink, int gw, int gh, int w, int h, int taps, int radius) {
  const int tx = blockIdx.x * 16 + (threadIdx.x & 15), ty = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (tx >= w || ty >= h) return;
  uint32_t s = diagHash((uint32_t)tx * 9781u + (uint32_t)ty * 6271u + 1u), acc = 0;
  for (int t = 0; t < taps; ++t) {
    s = s * 1664525u + 1013904223u;
    const int dx = (int)((s >> 8) % (2u * (uint32_t)radius)) - radius, dy = (int)((s >> 20) % (2u * (uint32_t)radius)) - radius;
    const int x = min(max(tx + dx, 0), gw - 1), y = min(max(ty + dy, 0), gh - 1);
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "l"(static_cast<const char*>(recs) + ((size_t)y * gw + x) * 32));
    acc += r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
  }
  if (acc == 0x9e3779b9u) sink[0] = acc; // keeps the loads alive; practically never taken
}
} // namespace

extern "C" {
int althea_cuda_diag_ssao_gathers(althea_cuda_ctx* ctx, uint64_t* out_records) {
  if (!ctx || !out_records) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (!ctx->gatherCounter) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "no SSAO launch has run with ALTHEA_CTX_SSAO_COUNT_TAPS set");
  CUDA_TRY(ctx, cudaDeviceSynchronize());
  unsigned long long v = 0;
  CUDA_TRY(ctx, cudaMemcpy(&v, ctx->gatherCounter, sizeof v, cudaMemcpyDeviceToHost));
  *out_records = v;
  return ALTHEA_OK;
}
int althea_cuda_diag_ssao_exact_fallbacks(althea_cuda_ctx* ctx, uint64_t* out_taps) {
  if (!ctx || !out_taps) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (!ctx->gatherCounter) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "no SSAO launch has run with ALTHEA_CTX_SSAO_COUNT_TAPS set");
  CUDA_TRY(ctx, cudaDeviceSynchronize());
  unsigned long long v = 0;
  CUDA_TRY(ctx, cudaMemcpy(&v, ctx->gatherCounter + 1, sizeof v, cudaMemcpyDeviceToHost));
  *out_taps = v;
  return ALTHEA_OK;
}

// undocumented tuning aid: the per-level / per-tap histogram the counting cull kernel fills (76 values after the four counters)
int althea_cuda_diag_ssao_cull_histogram(althea_cuda_ctx* ctx, uint64_t* out76) {
  if (!ctx || !out76 || !ctx->gatherCounter) return ALTHEA_ERR_INVALID_ARGUMENT;
  CUDA_TRY(ctx, cudaDeviceSynchronize());
  CUDA_TRY(ctx, cudaMemcpy(out76, ctx->gatherCounter + 4, 76 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return ALTHEA_OK;
}
int althea_cuda_diag_ssao_cull(althea_cuda_ctx* ctx, uint64_t out_counts[4]) {
  if (!ctx || !out_counts) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (!ctx->gatherCounter) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "no SSAO launch has run with ALTHEA_CTX_SSAO_COUNT_TAPS set");
  CUDA_TRY(ctx, cudaDeviceSynchronize());
  unsigned long long v[4] = {0, 0, 0, 0};
  CUDA_TRY(ctx, cudaMemcpy(v, ctx->gatherCounter, sizeof v, cudaMemcpyDeviceToHost));
  for (int k = 0; k < 4; ++k) out_counts[k] = v[k];
  return ALTHEA_OK;
}

int althea_cuda_diag_gather_ceiling(althea_cuda_ctx* ctx, uint32_t w, uint32_t h, uint32_t radius, uint32_t taps_per_pixel, double* out_records_per_second) {
  if (!ctx || !out_records_per_second || !w || !h || !radius || !taps_per_pixel) return ALTHEA_ERR_INVALID_ARGUMENT;
  const int gw = (int)w + 1, gh = (int)h + 1;
  const size_t need = (size_t)gw * gh * 32;
  void* recs = nullptr;
  uint32_t* sink = nullptr;
  cudaError_t e = cudaMalloc(&recs, need);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "cudaMalloc(gather records %zu): %s", need, cudaGetErrorString(e)); }
  e = cudaMalloc(&sink, sizeof(uint32_t));
  if (e != cudaSuccess) { cudaGetLastError(); cudaFree(recs); return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "cudaMalloc(sink): %s", cudaGetErrorString(e)); }
  cudaMemsetAsync(recs, 1, need, ctx->stream);
  const dim3 grid((w + 15) / 16, (h + 15) / 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  diag_gather_kernel<<<grid, 256, 0, ctx->stream>>>(recs, sink, gw, gh, (int)w, (int)h, 8, (int)radius); // warm-up
  cudaEventRecord(e0, ctx->stream);
  diag_gather_kernel<<<grid, 256, 0, ctx->stream>>>(recs, sink, gw, gh, (int)w, (int)h, (int)taps_per_pixel, (int)radius);
  cudaEventRecord(e1, ctx->stream);
  ctx->launches += 2;
  e = cudaEventSynchronize(e1);
  float ms = 0.0f;
  if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(recs);
  cudaFree(sink);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, ALTHEA_ERR_CUDA, "gather ceiling: %s", cudaGetErrorString(e)); }
  *out_records_per_second = (double)w * h * taps_per_pixel / ((double)ms * 1e-3);
  return ALTHEA_OK;
}
} // extern "C"

// ---- rasterising producers: G-buffer pass and omni shadow cubes (raster_kernels.cu) -----------------------------------------
struct RasterScratch {
  void* recs = nullptr; size_t recsBytes = 0;
  void* work = nullptr; size_t workBytes = 0;
  void* vis = nullptr; size_t visBytes = 0;
  void* openBound = nullptr; size_t openBoundBytes = 0; // alpha blending: per-pixel peel bound (4 B/px)
  void* layerTri = nullptr; size_t layerTriBytes = 0;   // and the translucent layers' triangle ordinals (allocated on first need)
  void* prims = nullptr; size_t primsBytes = 0;
  void* views = nullptr; size_t viewsBytes = 0;
  uint32_t* counters = nullptr;
  unsigned int* minAlphaDev = nullptr;
  bool srgbUploaded = false;
  int sms = 148;
  std::unordered_map<uint64_t, unsigned> minAlpha; // per base-colour image handle
};
static void freeRasterScratch(RasterScratch* r) {
  if (!r) return;
  for (void* p : {r->recs, r->work, r->vis, r->openBound, r->layerTri, r->prims, r->views, (void*)r->counters, (void*)r->minAlphaDev})
    if (p) cudaFree(p);
  delete r;
}

namespace {
int growScratch(althea_cuda_ctx* ctx, void** p, size_t* have, size_t need, const char* what) {
  if (*have >= need) return ALTHEA_OK;
  if (*p) { cudaDeviceSynchronize(); cudaFree(*p); *p = nullptr; *have = 0; }
  cudaError_t e = cudaMalloc(p, need);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "cudaMalloc(%s, %zu bytes): %s", what, need, cudaGetErrorString(e)); }
  *have = need;
  return ALTHEA_OK;
}

int rasterScratch(althea_cuda_ctx* ctx, RasterScratch** out) {
  if (!ctx->raster) {
    ctx->raster = new RasterScratch();
    cudaDeviceGetAttribute(&ctx->raster->sms, cudaDevAttrMultiProcessorCount, ctx->device);
    CUDA_TRY(ctx, cudaMalloc(&ctx->raster->counters, 8 * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->raster->minAlphaDev, sizeof(unsigned int)));
  }
  *out = ctx->raster;
  return ALTHEA_OK;
}

int textureOf(althea_cuda_ctx* ctx, const althea_texture_ref& ref, const char* what, RasterTex* out) {
  memset(out, 0, sizeof *out);
  if (!ref.image) return ALTHEA_OK;
  Resource* r;
  int rc = getImage(ctx, ref.image, ALTHEA_FORMAT_R8G8B8A8_UNORM, what, &r);
  if (rc) return rc;
  if (r->mips == 1 && r->pitch0 != (size_t)r->w * 4) return fail(ctx, ALTHEA_ERR_UNSUPPORTED, "%s: textures must be tightly packed", what);
  if (r->mips > (uint32_t)kMaxMips + 2) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "%s: too many mip levels", what);
  out->texels = static_cast<const uint8_t*>(r->dptr);
  out->w = (int)r->w; out->h = (int)r->h; out->mips = (int)r->mips;
  out->sampler = ref.sampler;
  return ALTHEA_OK;
}

// device-side primitive table + the draw-ordered triangle count
int buildPrims(althea_cuda_ctx* ctx, RasterScratch* R, const althea_primitive* prims, uint32_t n, cudaStream_t stream, uint32_t* triTotal) {
  if (!prims && n) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "primitives is null");
  std::vector<RasterPrim> host(n);
  uint64_t tris = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const althea_primitive& p = prims[i];
    RasterPrim& d = host[i];
    memset(&d, 0, sizeof d);
    Resource* vb = find(ctx, p.vertices, ResKind::Buffer);
    Resource* ib = find(ctx, p.indices, ResKind::Buffer);
    if (!vb || !ib) return fail(ctx, ALTHEA_ERR_BAD_HANDLE, "primitive %u: vertices/indices are not live buffers", i);
    if (p.index_count % 3u) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "primitive %u: index_count %u is not a triangle list", i, p.index_count);
    if ((size_t)p.index_count * 4 > ib->bytes) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "primitive %u: index buffer holds %zu bytes, index_count=%u", i, ib->bytes, p.index_count);
    if (vb->bytes < sizeof(althea_vertex)) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "primitive %u: empty vertex buffer", i);
    d.verts = static_cast<const althea_vertex*>(vb->dptr);
    d.idx = static_cast<const uint32_t*>(ib->dptr);
    d.triCount = p.index_count / 3u;
    d.vertCount = (uint32_t)std::min<size_t>(vb->bytes / sizeof(althea_vertex), 0xffffffffu);
    d.triOffset = (uint32_t)tris;
    tris += d.triCount;
    memcpy(d.model, p.model, sizeof d.model);
    d.frontCW = p.front_face_clockwise ? 1u : 0u;
    const althea_material& m = p.material;
    memcpy(d.mat.baseColorFactor, m.baseColorFactor, sizeof d.mat.baseColorFactor);
    if (m.baseTextureCoordinateIndex < 0 || m.baseTextureCoordinateIndex > 3 || m.metallicRoughnessTextureCoordinateIndex < 0 || m.metallicRoughnessTextureCoordinateIndex > 3)
      return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "primitive %u: texture coordinate index out of [0, 3]", i);
    d.mat.baseUv = m.baseTextureCoordinateIndex;
    d.mat.mrUv = m.metallicRoughnessTextureCoordinateIndex;
    d.mat.normalScale = m.normalScale; d.mat.metallicFactor = m.metallicFactor; d.mat.roughnessFactor = m.roughnessFactor; d.mat.alphaCutoff = m.alphaCutoff;
    int rc;
    if ((rc = textureOf(ctx, m.baseTexture, "material.baseTexture", &d.mat.base))) return rc;
    if ((rc = textureOf(ctx, m.normalTexture, "material.normalTexture", &d.mat.normal))) return rc;
    if ((rc = textureOf(ctx, m.metallicRoughnessTexture, "material.metallicRoughnessTexture", &d.mat.mr))) return rc;
    // can the fragment's alpha ever fall below the cutoff? (filtered alpha is a convex combination of texel alphas)
    float minA = 1.0f;
    if (d.mat.base.texels) {
      auto it = R->minAlpha.find(m.baseTexture.image);
      if (it == R->minAlpha.end()) {
        unsigned init = 255u, got = 255u;
        CUDA_TRY(ctx, cudaMemcpyAsync(R->minAlphaDev, &init, sizeof init, cudaMemcpyHostToDevice, stream));
        althea_raster::launch_texture_min_alpha(reinterpret_cast<const uint32_t*>(d.mat.base.texels), (size_t)d.mat.base.w * d.mat.base.h, R->minAlphaDev, stream);
        ctx->launches += 1;
        CUDA_TRY(ctx, cudaMemcpyAsync(&got, R->minAlphaDev, sizeof got, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(stream));
        it = R->minAlpha.emplace(m.baseTexture.image, got).first;
      }
      minA = (float)it->second / 255.0f;
    }
    d.opaque = (m.alphaCutoff <= 0.0f || minA * m.baseColorFactor[3] * (1.0f - 1e-5f) >= m.alphaCutoff) ? 1u : 0u;
  }
  if (tris > 0x7fffffffull) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "too many triangles (%llu)", (unsigned long long)tris);
  *triTotal = (uint32_t)tris;
  int rc = growScratch(ctx, &R->prims, &R->primsBytes, (n ? n : 1u) * sizeof(RasterPrim), "raster primitive table");
  if (rc) return rc;
  if (n) {
    CUDA_TRY(ctx, cudaMemcpyAsync(R->prims, host.data(), n * sizeof(RasterPrim), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(stream)); // `host` is pageable and dies with this scope
  }
  if (!R->srgbUploaded) {
    float table[256];
    for (int i = 0; i < 256; ++i) {
      const double c = i / 255.0;
      table[i] = (float)(c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4));
    }
    althea_raster::upload_srgb_table(table, stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(stream));
    R->srgbUploaded = true;
  }
  return ALTHEA_OK;
}

// setup then fill, no host round trip in between: the fill's persistent grid reads the work count on the device. Record and
// tile-work lists are sized by demand, not by the worst case (every triangle visible in every view): if one is too small the
// setup raises the pass's overflow flag, keeps counting what it would have needed (counters[3], [4]) and the fill does nothing;
// the caller checks once per draw call (rasterOverflow) and redoes the call with larger lists.
int runRaster(althea_cuda_ctx* ctx, RasterScratch* R, RasterJob& J, cudaStream_t stream) {
  if (!J.triTotal || !J.nViews) return ALTHEA_OK;
  const unsigned long long pairs = (unsigned long long)J.triTotal * (unsigned)J.nViews;
  size_t recStart = (size_t)1 << 21, workStart = (size_t)1 << 22; // 2 M records (256 MB) and 4 M tile items to start with
  if (const char* e = getenv("ALTHEA_RASTER_INITIAL_LISTS")) { // test hook: tiny lists force the overflow-and-redo path
    const long v = atol(e);
    if (v > 0) recStart = workStart = (size_t)v;
  }
  const size_t recWant = (size_t)std::min<unsigned long long>(pairs, recStart);
  int rc;
  if (R->recsBytes < recWant * sizeof(RasterRecord) && (rc = growScratch(ctx, &R->recs, &R->recsBytes, recWant * sizeof(RasterRecord), "raster records"))) return rc;
  if (R->workBytes < workStart * sizeof(uint2) && (rc = growScratch(ctx, &R->work, &R->workBytes, workStart * sizeof(uint2), "raster tile work list"))) return rc;
  J.recs = static_cast<RasterRecord*>(R->recs);
  J.recCap = (uint32_t)std::min<size_t>(R->recsBytes / sizeof(RasterRecord), 0xffffffffu);
  J.work = static_cast<uint2*>(R->work);
  J.workCap = (uint32_t)std::min<size_t>(R->workBytes / sizeof(uint2), 0xffffffffu);
  J.counters = R->counters;
  CUDA_TRY(ctx, cudaMemsetAsync(R->counters, 0, 3 * sizeof(uint32_t), stream)); // [3], [4] stay: sticky across the passes of a call
  timedLaunch(ctx, "raster_setup", stream, [&] { althea_raster::launch_raster_setup(J, stream); });
  timedLaunch(ctx, "raster_fill", stream, [&] { althea_raster::launch_raster_fill(J, R->sms, stream); });
  return ALTHEA_OK;
}
// after the last pass of a draw call: grows whichever list overflowed (one host sync); *again = the call must be redone
int rasterOverflow(althea_cuda_ctx* ctx, RasterScratch* R, cudaStream_t stream, bool* again, uint32_t* openPixels = nullptr) {
  uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CUDA_TRY(ctx, cudaMemcpyAsync(c, R->counters, sizeof c, cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(stream));
  *again = c[3] != 0 || c[4] != 0;
  if (openPixels) *openPixels = c[5];
  int rc;
  if (c[3] && (rc = growScratch(ctx, &R->work, &R->workBytes, ((size_t)c[3] + (c[3] >> 2) + 1024) * sizeof(uint2), "raster tile work list"))) return rc;
  if (c[4] && (rc = growScratch(ctx, &R->recs, &R->recsBytes, ((size_t)c[4] + (c[4] >> 2) + 1024) * sizeof(RasterRecord), "raster records"))) return rc;
  return ALTHEA_OK;
}
} // namespace

extern "C" {
int althea_cuda_draw_gbuffer(althea_cuda_ctx* ctx, const althea_global_uniforms* uniforms, const althea_primitive* primitives, uint32_t primitive_count,
                             const althea_gbuffer* gbuffer, const althea_sync* sync) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (!uniforms || !gbuffer) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "uniforms and gbuffer must be non-null");
  Resource *depth, *position, *normal, *albedo, *mro;
  int rc;
  if ((rc = getImage(ctx, gbuffer->depth, ALTHEA_FORMAT_R32_SFLOAT, "gbuffer.depth", &depth, true))) return rc;
  if ((rc = getImage(ctx, gbuffer->position, ALTHEA_FORMAT_R32G32B32A32_SFLOAT, "gbuffer.position", &position, true))) return rc;
  if ((rc = getImage(ctx, gbuffer->normal, ALTHEA_FORMAT_R16G16B16A16_SFLOAT, "gbuffer.normal", &normal, true))) return rc;
  if ((rc = getImage(ctx, gbuffer->albedo, ALTHEA_FORMAT_R8G8B8A8_UNORM, "gbuffer.albedo", &albedo, true))) return rc;
  if ((rc = getImage(ctx, gbuffer->mro, ALTHEA_FORMAT_R8G8B8A8_UNORM, "gbuffer.mro", &mro, true))) return rc;
  Resource* all[5] = {depth, position, normal, albedo, mro};
  Resource* first = nullptr;
  for (Resource* r : all) {
    if (!r) continue;
    if (!first) first = r;
    if (r->w != first->w || r->h != first->h) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "G-buffer attachments must all be %ux%u", first->w, first->h);
  }
  if (!first) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "gbuffer has no attachment to write");
  if (first->w > 65535u || first->h > 65535u) return fail(ctx, ALTHEA_ERR_UNSUPPORTED, "frame larger than 65535 pixels on a side");
  RasterScratch* R;
  if ((rc = rasterScratch(ctx, &R))) return rc;
  cudaStream_t stream;
  if ((rc = beginWork(ctx, sync, &stream))) return rc;
  WorkGuard work(ctx, sync, stream);
  RasterJob J;
  memset(&J, 0, sizeof J);
  if ((rc = buildPrims(ctx, R, primitives, primitive_count, stream, &J.triTotal))) return rc;
  J.prims = static_cast<const RasterPrim*>(R->prims);
  J.nPrims = (int)primitive_count;
  J.W = (int)first->w;
  J.H = (int)first->h;
  J.mode = RASTER_MODE_GBUFFER;
  J.tile = 64;
  RasterView view;
  memset(&view, 0, sizeof view);
  matmul44(uniforms->projection, uniforms->view, view.a); // Gltf.vert:55 evaluates (projection * view) * worldPos
  if ((rc = growScratch(ctx, &R->views, &R->viewsBytes, 8 * sizeof(RasterView), "raster views"))) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(R->views, &view, sizeof view, cudaMemcpyHostToDevice, stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(stream));
  J.views = static_cast<const RasterView*>(R->views);
  J.nViews = 1;
  const size_t px = (size_t)J.W * J.H;
  if ((rc = growScratch(ctx, &R->vis, &R->visBytes, px * sizeof(unsigned long long), "visibility buffer"))) return rc;
  J.vis = static_cast<unsigned long long*>(R->vis);
  if ((rc = growScratch(ctx, &R->openBound, &R->openBoundBytes, px * sizeof(uint32_t), "peel bounds"))) return rc;
  J.openBound = static_cast<uint32_t*>(R->openBound);
  ImgView v;
  if (depth) { levelView(*depth, 0, 0, &v); J.outDepth = static_cast<float*>(const_cast<void*>(v.ptr)); J.pitchDepth = v.pitch; }
  if (position) { levelView(*position, 0, 0, &v); J.outPosition = static_cast<float4*>(const_cast<void*>(v.ptr)); J.pitchPosition = v.pitch; }
  if (normal) { levelView(*normal, 0, 0, &v); J.outNormal = static_cast<uint2*>(const_cast<void*>(v.ptr)); J.pitchNormal = v.pitch; }
  if (albedo) { levelView(*albedo, 0, 0, &v); J.outAlbedo = static_cast<uint32_t*>(const_cast<void*>(v.ptr)); J.pitchAlbedo = v.pitch; }
  if (mro) { levelView(*mro, 0, 0, &v); J.outMro = static_cast<uint32_t*>(const_cast<void*>(v.ptr)); J.pitchMro = v.pitch; }
  uint32_t open = 0;
  for (int attempt = 0;; ++attempt) {
    CUDA_TRY(ctx, cudaMemsetAsync(R->counters, 0, 8 * sizeof(uint32_t), stream));
    timedLaunch(ctx, "raster_clear", stream, [&] { althea_raster::launch_raster_clear(J.vis, nullptr, px, stream); });
    J.bound = nullptr;
    if ((rc = runRaster(ctx, R, J, stream))) return rc;
    timedLaunch(ctx, "gbuffer_resolve", stream, [&] { althea_raster::launch_gbuffer_resolve(J, stream); });
    bool again = false;
    if ((rc = rasterOverflow(ctx, R, stream, &again, &open))) return rc;
    if (!again) break;
    if (attempt >= 2) return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "raster lists overflowed three times");
  }
  // Translucent depth winners (alpha < 1 after the cutoff): the attachments are alpha-blended in draw order
  // (Src/GraphicsPipeline.cpp:138-154), so what was drawn under them matters. Peel it, one layer per pass, until every open
  // pixel reaches an opaque fragment or the clear colour (at most kPeelLayers layers; deeper ones are dropped).
  constexpr int kPeelLayers = 4;
  if (open) {
    if ((rc = growScratch(ctx, &R->layerTri, &R->layerTriBytes, px * sizeof(uint32_t) * kPeelLayers, "translucent layer stack"))) return rc;
    J.layerTri = static_cast<uint32_t*>(R->layerTri);
    timedLaunch(ctx, "gbuffer_peel", stream, [&] { althea_raster::launch_gbuffer_peel(J, 0, 0, stream); });
    for (int pass = 1; pass < kPeelLayers && open; ++pass) {
      timedLaunch(ctx, "raster_clear", stream, [&] { althea_raster::launch_raster_clear(J.vis, nullptr, px, stream); });
      J.bound = J.openBound;
      if ((rc = runRaster(ctx, R, J, stream))) return rc; // same triangles as pass 0: the lists are large enough
      CUDA_TRY(ctx, cudaMemsetAsync(R->counters + 5, 0, sizeof(uint32_t), stream));
      timedLaunch(ctx, "gbuffer_peel", stream, [&] { althea_raster::launch_gbuffer_peel(J, pass, pass == kPeelLayers - 1, stream); });
      bool again = false;
      if ((rc = rasterOverflow(ctx, R, stream, &again, &open))) return rc;
    }
  }
  return work.finish();
}

int althea_cuda_draw_shadow_cubes(althea_cuda_ctx* ctx, uint64_t lights_buf, uint32_t light_count, const althea_point_light_constants* constants,
                                  const althea_primitive* primitives, uint32_t primitive_count, uint64_t shadow_cube_array, const althea_sync* sync) {
  if (!ctx) return ALTHEA_ERR_INVALID_ARGUMENT;
  if (!constants) return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "constants must be non-null");
  Resource* sh;
  int rc = getImage(ctx, shadow_cube_array, ALTHEA_FORMAT_R32_SFLOAT, "shadow_cube_array", &sh);
  if (rc) return rc;
  if (sh->w != sh->h || sh->layers < 6u * light_count || sh->mips != 1 || sh->pitch0 != (size_t)sh->w * 4)
    return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "shadow_cube_array needs square, tightly packed, single-mip faces and >= %u layers (has %u)", 6u * light_count, sh->layers);
  if (sh->w > 65535u) return fail(ctx, ALTHEA_ERR_UNSUPPORTED, "shadow faces larger than 65535 pixels");
  Resource* lb = light_count ? find(ctx, lights_buf, ResKind::Buffer) : nullptr;
  if (light_count && (!lb || lb->bytes < (size_t)light_count * sizeof(althea_point_light)))
    return fail(ctx, ALTHEA_ERR_INVALID_ARGUMENT, "lights_buf must hold %u althea_point_light records", light_count);
  RasterScratch* R;
  if ((rc = rasterScratch(ctx, &R))) return rc;
  cudaStream_t stream;
  if ((rc = beginWork(ctx, sync, &stream))) return rc;
  WorkGuard work(ctx, sync, stream);
  std::vector<althea_point_light> lights(light_count);
  if (light_count) {
    CUDA_TRY(ctx, cudaMemcpyAsync(lights.data(), lb->dptr, light_count * sizeof(althea_point_light), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(stream));
  }
  RasterJob J;
  memset(&J, 0, sizeof J);
  if ((rc = buildPrims(ctx, R, primitives, primitive_count, stream, &J.triTotal))) return rc;
  J.prims = static_cast<const RasterPrim*>(R->prims);
  J.nPrims = (int)primitive_count;
  J.W = J.H = (int)sh->w;
  J.mode = RASTER_MODE_SHADOW;
  J.tile = 32;
  J.nViews = 6;
  if ((rc = growScratch(ctx, &R->views, &R->viewsBytes, 8 * sizeof(RasterView), "raster views"))) return rc;
  J.views = static_cast<const RasterView*>(R->views);
  ImgView l0, l1;
  levelView(*sh, 0, 0, &l0);
  J.shadowLayerStride = sh->layers > 1 && levelView(*sh, 0, 1, &l1) ? (size_t)((const char*)l1.ptr - (const char*)l0.ptr) / sizeof(float) : (size_t)sh->w * sh->h;
  const size_t facePx = (size_t)sh->w * sh->h;
  // the views of every light go up in one copy; each light's pass reads its six
  std::vector<RasterView> views((size_t)6 * (light_count ? light_count : 1u));
  memset(views.data(), 0, views.size() * sizeof(RasterView));
  for (uint32_t l = 0; l < light_count; ++l)
    for (int f = 0; f < 6; ++f) { // ShadowMapBindless.vert:44-47: csPos = views[gl_ViewIndex] * (worldPos - light), gl_Position = projection * csPos
      RasterView& v = views[6 * l + f];
      memcpy(v.a, constants->views[f], sizeof v.a);
      memcpy(v.b, constants->projection, sizeof v.b);
      memcpy(v.off, lights[l].position, sizeof v.off);
      v.hasB = 1;
    }
  if ((rc = growScratch(ctx, &R->views, &R->viewsBytes, views.size() * sizeof(RasterView), "raster views"))) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(R->views, views.data(), views.size() * sizeof(RasterView), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(stream));
  { // which face looks along which axis: a view matrix maps world to view space and the camera looks down -z of view space, so its
    // forward direction in world space is minus the third row of the rotation, -(a[2], a[6], a[10])
    bool ok = true;
    int seen = 0;
    for (int f = 0; f < 6 && ok; ++f) {
      const float* a = constants->views[f];
      const float fwd[3] = {-a[2], -a[6], -a[10]};
      int axis = -1;
      for (int j = 0; j < 3; ++j)
        if (fabsf(fwd[j]) > 0.9999f) axis = 2 * j + (fwd[j] < 0.0f ? 1 : 0);
      // a plain rigid camera at the origin with the reference's 90 degree, aspect 1 projection: otherwise no shortcut
      ok = axis >= 0 && !(seen & (1 << axis)) && a[12] == 0.0f && a[13] == 0.0f && a[14] == 0.0f && a[3] == 0.0f && a[7] == 0.0f && a[11] == 0.0f && a[15] == 1.0f;
      if (ok) { J.faceOfAxis[axis] = f; seen |= 1 << axis; }
    }
    const float* pr = constants->projection;
    ok = ok && seen == 63 && fabsf(fabsf(pr[0]) - 1.0f) < 1e-5f && fabsf(fabsf(pr[5]) - 1.0f) < 1e-5f && pr[11] == -1.0f && pr[15] == 0.0f &&
         pr[1] == 0.0f && pr[2] == 0.0f && pr[3] == 0.0f && pr[4] == 0.0f && pr[6] == 0.0f && pr[7] == 0.0f && pr[8] == 0.0f && pr[9] == 0.0f && pr[12] == 0.0f && pr[13] == 0.0f;
    J.cubeFaces = ok ? 1 : 0;
  }
  // ONE pass for every light: the layers of the cube array are contiguous (layer = 6 * light + face = the view index), so a
  // triangle's vertices are fetched and transformed to world space once and tested against all 6 * light_count faces
  J.views = static_cast<const RasterView*>(R->views);
  J.nViews = (int)(6u * light_count);
  J.shadowBase = static_cast<float*>(const_cast<void*>(l0.ptr));
  for (int attempt = 0; light_count; ++attempt) {
    CUDA_TRY(ctx, cudaMemsetAsync(R->counters, 0, 8 * sizeof(uint32_t), stream));
    timedLaunch(ctx, "raster_clear", stream, [&] { althea_raster::launch_raster_clear(nullptr, J.shadowBase, J.shadowLayerStride * (6u * light_count - 1u) + facePx, stream); });
    if ((rc = runRaster(ctx, R, J, stream))) return rc;
    bool again = false;
    if ((rc = rasterOverflow(ctx, R, stream, &again))) return rc;
    if (!again) break;
    if (attempt >= 2) return fail(ctx, ALTHEA_ERR_OUT_OF_MEMORY, "raster lists overflowed three times");
  }
  return work.finish();
}
} // extern "C"
