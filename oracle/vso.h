/* vso.h -- C ABI of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a dependency-free C++17 restatement of the reference's dense
 * over-segmentation path (videosegmentation/video_segment @ c930c455).  Every
 * function cites the reference file:line it follows (paths relative to the
 * reference root).  PINNING: the reference ships no tests, golden vectors or
 * fixtures for this path and its executables cannot be
 * built in this image (OpenCV 2.4 / FFmpeg 2.2 / glog / gflags / boost /
 * protobuf are absent), but its over-segmentation library compiles UNMODIFIED
 * into oracle/_ref against small stand-ins for those libraries (Makefile target
 * _ref, oracle/ref_shim/): libref_results.so is the reference's whole
 * DenseSegmentation::ProcessFrame stream, and tests/test_oracle_cpu.py holds this
 * restatement identical to it in every field of every frame result (17 cases),
 * plus the bilateral filter, the merge and the colour histogram in isolation.
 * Reference-generated golden digests: tests/golden/reference_results.json.
 * Third-party arithmetic (cv::Mat::convertTo, cv::copyMakeBorder, cv::minMaxLoc,
 * 8-bit cv::cvtColor BGR2Lab) is pinned by cv2 golden vectors (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (video_segment_b200/) never links or imports it.
 */
#ifndef VSO_H_
#define VSO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors DenseSegmentationOptions (segmentation/dense_segmentation.h:42-95). */
typedef struct vso_dense_opts {
  int32_t presmoothing;               /* 0 none, 1 gaussian(unsupported), 2 bilateral */
  float frac_min_region_size;         /* 0.01 */
  int32_t chunk_size;                 /* 20 */
  float chunk_overlap_ratio;          /* 0.2 */
  int32_t num_constraint_frames;      /* 1 */
  int32_t enforce_n4_connectivity;    /* 1 */
  int32_t enforce_spatial_connectedness; /* 1 */
  int32_t color_distance;             /* 0 L1, 1 L2 */
  int32_t num_threads;                /* oracle only: 1 = serial, >1 = reference-style
                                         threading (thread per frame edge build,
                                         row blocks for the bilateral filter) */
} vso_dense_opts;

/* Same layout as vsb200_frame_result (include/vsb200.h) so tests can compare
 * the two field by field.  All pointers are owned by the handle and valid until
 * the next pop / destroy. */
typedef struct vso_frame_result {
  int32_t width, height, chunk_id, chunk_size, overlap_start, hierarchy_frame_idx;
  int32_t connectedness;              /* 1 = N4_CONNECT, 2 = N8_CONNECT */
  int32_t n_regions;
  const int32_t* region_id;           /* [n_regions] */
  const int32_t* interval_offset;     /* [n_regions + 1] */
  const int32_t* intervals;           /* [3 * n_intervals]: y, left_x, right_x */
  const float* shape_moments;         /* [6 * n_regions]: size mean_x mean_y xx xy yy */
  int32_t n_compound;                 /* > 0 only on the first frame of a chunk */
  const int32_t* compound;            /* [4 * n_compound]: id size start_frame end_frame */
  const int32_t* neighbor_offset;     /* [n_compound + 1] */
  const int32_t* neighbor_id;
  int64_t pts;
} vso_frame_result;

typedef struct vso_dense vso_dense;

void vso_default_opts(vso_dense_opts* o);

/* ---- stage functions (kernel-level oracles) ---- */

/* cv::Mat::convertTo(CV_32FC3, 1/255)  -- dense_segmentation.cpp:180-181 */
void vso_convert_u8_to_f32(const uint8_t* bgr, int w, int h, int row_stride, float* out);
/* imagefilter::BilateralFilter(in, sigma_space, sigma_color) -- image_filter.cpp:184-277.
 * lut_out (nullable) receives the 12288-entry exp LUT, scale_out the LUT scale. */
void vso_bilateral(const float* in, int w, int h, float sigma_space, float sigma_color,
                   float* out, int num_threads, float* lut_out, float* scale_out);
/* Full PreprocessFeatures (dense_segmentation.cpp:164-198). */
void vso_preprocess(const uint8_t* bgr, int w, int h, int row_stride, int presmoothing,
                    float* out, int num_threads);
/* Spatial edge weights in direction-planar layout out[d][y][x], d = R,B,BL,BR
 * (dense_segmentation_graph.h:956-1000); missing edges hold -1. */
void vso_spatial_weights(const float* img, int w, int h, int l1, float* out);
/* Temporal edge weights out[d][y][x], d = TL,T,TR,L,C,R,BL,B,BR relative to the
 * (flow displaced, clamped) centre in prev (dense_segmentation_graph.h:1002-1142);
 * flow nullable; missing edges hold -1. */
void vso_temporal_weights(const float* curr, const float* prev, const float* flow,
                          int w, int h, int l1, float* out);
/* Bucket index (segmentation_graph.h:158-162,336). */
int vso_bucket_index(float weight);

/* ---- streaming engine = DenseSegmentation (dense_segmentation.cpp) ---- */
int vso_dense_create(const vso_dense_opts* o, int w, int h, int use_flow, vso_dense** out);
int vso_dense_push(vso_dense*, const uint8_t* bgr, int row_stride, const float* flow_xy,
                   int flow_row_stride_bytes, int64_t pts, int* n_ready);
int vso_dense_flush(vso_dense*, int* n_ready);
int vso_dense_pop(vso_dense*, vso_frame_result* out);
void vso_dense_destroy(vso_dense*);

/* Debug taps on the most recently segmented chunk (valid until next chunk):
 * number of graph slots, per-node union-find labels right after SegmentGraph +
 * FlattenUnionFind (before N4 / connectedness), and the per-slot id images after
 * EnforceN4Connectivity.  Labels are reference representative ids (opaque). */
int vso_dense_last_chunk_slots(vso_dense*);
const int32_t* vso_dense_last_chunk_node_labels(vso_dense*);   /* [slots * w * h] */
const int32_t* vso_dense_last_chunk_id_images(vso_dense*);     /* [slots * w * h], -1 on virtual slots */
/* Merge statistics of the last chunk: regular, small-region, forced merges. */
void vso_dense_last_chunk_merge_stats(vso_dense*, int64_t stats[3]);
/* Stage wall-times accumulated since creation, seconds:
 * [0] preprocess [1] edge build [2] segment graph [3] obtain results [4] shaping */
void vso_dense_stage_seconds(vso_dense*, double out[5]);

/* ---- whole-graph helper for merge-kernel parity: segment one chunk given the
 * smoothed float frames (no constraints) and return node labels. ---- */
int vso_segment_chunk_labels(const float* frames, int w, int h, int t, int l1,
                             int min_region_size, int32_t* labels_out);

/* ---- region stage, appearance descriptor (vso_region.cpp) ---- */
/* cv::cvtColor(CV_BGR2Lab) on 8-bit data (region_descriptor.cpp:73); lab_out is dense [h][w][3]. */
void vso_bgr2lab(const uint8_t* bgr, int w, int h, int row_stride, uint8_t* lab_out);
/* AppearanceDescriptor3D::AddFeatures (region_descriptor.cpp:97-111) for every region of one frame. */
void vso_region_hist_add(const uint8_t* lab, const int32_t* ids, int w, int h, int n_regions, int lum_bins,
                         int color_bins, int exact, double* hist, double* weight_sum);
/* ColorHistogram::NormalizeToOne (histograms.cpp:340-360). */
void vso_hist_normalize(const double* hist, const double* weight_sum, int n_regions, int total_bins, int exact, float* out);
/* ColorHistogram::ChiSquareDist (histograms.cpp:391-407) for region pairs [2 * n_pairs]. */
void vso_hist_chisquare(const float* hist, int total_bins, const int32_t* pairs, int n_pairs, float* out);

/* Test tap: the chunk hand-over state after the last chunk boundary -- the two overlap frames' region-id maps
 * (int32 [2][h][w]), state = {max_region_id_, id of the chunk they constrain, frames output so far}
 * (dense_segmentation.cpp:300-328,360-365).  Returns 1 before the first boundary. */
int vso_dense_last_overlap_state(vso_dense* d, const int32_t** maps, int32_t state[3]);

/* Hierarchical region stage (vso_hier.cpp): RegionSegmentation::ProcessFrame fed with the over-segmentation of the dense
 * stage, frame by frame (segmentation/region_segmentation.cpp:97-205).  Results pop as flat int32 records in the layout
 * of oracle/ref_hier_wrap.cpp (ref_hier_pop). */
void* vso_hier_create(int width, int height, int use_flow, int chunk_set_size, int chunk_set_overlap, int constraint_chunks,
                      int min_region_num, int max_region_num, float level_cutoff_fraction, float small_region_penalizer);
int vso_hier_push(void* h, const vso_frame_result* overseg, const uint8_t* bgr, const float* flow_xy_or_null);
int vso_hier_flush(void* h);
long long vso_hier_pop(void* h, const int32_t** out);
void vso_hier_destroy(void* h);

#ifdef __cplusplus
}
#endif
#endif  /* VSO_H_ */
