// region_kernels.cuh -- launch wrappers shared by the region stage (region_hist.cu, region_stage.cu).
#pragma once
#include "common.cuh"

namespace vsb {

// AppearanceDescriptor3D::AddFeatures for every region of a frame at once (region_hist.cu): acc [n_regions][bins] u64
// fixed point 2^-26, cnt [n_regions] pixel counts.
int launch_region_hist(const uint8_t* dev_bgr, int row_stride_bytes, const int* dev_region_ids, int w, int h, int n_regions,
                       int lum_bins, int color_bins, unsigned long long* acc, unsigned* cnt, cudaStream_t s);

}  // namespace vsb
