// TEST INFRASTRUCTURE ONLY.  C wrapper around the REFERENCE's own ColorHistogram (segmentation/histograms.h,
// compiled unmodified from /root/reference by `make -C oracle _ref`), used to pin the oracle's restatement of the
// region appearance descriptor (oracle/vso_region.cpp) against the reference itself.
#include <stdint.h>

#include "segmentation/histograms.h"

extern "C" {

// AppearanceDescriptor3D semantics (region_descriptor.cpp:91-127): a ColorHistogram(lum, col) fed with
// AddPixelInterpolated for n Lab pixels in order, then NormalizeToOne; writes all bins and returns WeightSum().
// Dense storage: the public interface cannot enumerate a sparse histogram (GetBinValue of an absent bin is
// undefined); the per-bin arithmetic is the same, and the sparse path is exercised through ChiSquareDist below.
double ref_color_hist(const uint8_t* lab, int n, int lum_bins, int color_bins, float* bins_out) {
  segmentation::ColorHistogram h(lum_bins, color_bins, false);
  for (int i = 0; i < n; ++i) h.AddPixelInterpolated(lab + 3 * i);
  h.NormalizeToOne();
  const int total = lum_bins * color_bins * color_bins;
  for (int b = 0; b < total; ++b) bins_out[b] = h.GetBinValue(b);
  return h.WeightSum();
}

// ChiSquareDist between the descriptors of two pixel sets (both built as above).
float ref_color_hist_chisquare(const uint8_t* lab_a, int na, const uint8_t* lab_b, int nb, int lum_bins, int color_bins, int sparse) {
  segmentation::ColorHistogram a(lum_bins, color_bins, sparse != 0), b(lum_bins, color_bins, sparse != 0);
  for (int i = 0; i < na; ++i) a.AddPixelInterpolated(lab_a + 3 * i);
  for (int i = 0; i < nb; ++i) b.AddPixelInterpolated(lab_b + 3 * i);
  a.NormalizeToOne();
  b.NormalizeToOne();
  return a.ChiSquareDist(b);
}

}  // extern "C"
