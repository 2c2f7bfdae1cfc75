// Host-callable launchers of raster_kernels.cu (namespace althea_raster).
#pragma once
#include <cuda_runtime.h>

#include "raster_params.h"

namespace althea_raster {
void upload_srgb_table(const float* table256, cudaStream_t s);
void launch_raster_clear(unsigned long long* vis, float* depth, size_t n, cudaStream_t s);
void launch_raster_setup(const RasterJob& J, cudaStream_t s);
void launch_raster_fill(const RasterJob& J, int sms, cudaStream_t s);
void launch_gbuffer_resolve(const RasterJob& J, cudaStream_t s);
void launch_gbuffer_peel(const RasterJob& J, int pass, int lastPass, cudaStream_t s);
void launch_texture_min_alpha(const uint32_t* texels, size_t n, unsigned int* out, cudaStream_t s);
} // namespace althea_raster
