// Host-callable launchers of the sm_100a kernels. frame_kernels.cu is compiled twice (namespaces althea_fast and
// althea_parity, see device_math.cuh); ibl_kernels.cu once (namespace althea_iblk).
#pragma once
#include <cuda_runtime.h>

#include "params.h"

#define ALTHEA_DECLARE_FRAME_LAUNCHERS(ns)                                   \
  namespace ns {                                                             \
  void launch_ssr_capture(const FrameParams& P, cudaStream_t s);             \
  bool ssr_march_reads_depth_quads();                                        \
  void launch_ssr_depth_pad(const FrameParams& P, cudaStream_t s);           \
  void launch_ssr_planes(const FrameParams& P, cudaStream_t s);              \
  void launch_ssr_shade_hits(const FrameParams& P, cudaStream_t s);          \
  void launch_reconstruct_position(const FrameParams& P, cudaStream_t s);    \
  void launch_glossy_convolve(const ConvolveParams& C, cudaStream_t s);      \
  void launch_ssao(const FrameParams& P, cudaStream_t s);                    \
  void launch_ssao_quads(const FrameParams& P, cudaStream_t s);              \
  void launch_ssao_planes(const FrameParams& P, cudaStream_t s, bool coarsest); \
  void launch_ssao_cull(const FrameParams& P, cudaStream_t s);               \
  void launch_ssao_dirs(float4* dirs, int row, int rows, cudaStream_t s);    \
  void launch_ssao_exact(const FrameParams& P, cudaStream_t s);              \
  void launch_deferred_shade(const FrameParams& P, cudaStream_t s);          \
  }
ALTHEA_DECLARE_FRAME_LAUNCHERS(althea_fast)
ALTHEA_DECLARE_FRAME_LAUNCHERS(althea_parity)

namespace althea_iblk {
void launch_mip_downsample(const MipGenParams& M, cudaStream_t s);
void launch_ibl_irradiance(const IblParams& I, cudaStream_t s);
void launch_ibl_prefilter(const IblParams& I, cudaStream_t s);
void launch_brdf_lut(const LutParams& L, cudaStream_t s);
} // namespace althea_iblk
