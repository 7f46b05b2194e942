/* vsb200.h -- C ABI of the B200-native dense video over-segmentation path.
 *
 * Drop-in boundary for the reference's DenseSegmentationUnit plug point
 * (videosegmentation/video_segment @ c930c455; paths below are relative to the
 * reference root).  The reference has no C ABI: these entry points are what a
 * VideoUnit adapter binds.  Conventions mirror the reference:
 *   - setup returns a status (<-> OpenStreams() bool + LOG(ERROR),
 *     segmentation/segmentation_unit.cpp:58-116); 0 = ok;
 *   - invariant violations abort (<-> CHECK);
 *   - one handle is driven by one thread (<-> one thread per VideoUnit,
 *     video_framework/video_pipeline.cpp:82-135);
 *   - input buffers are HOST pointers owned by the caller and may be released as
 *     soon as push returns; results are owned by the handle until the next pop.
 * No torch / C++ types cross this boundary.  There is no CPU fallback: every
 * entry point fails with VSB200_ERR_NO_DEVICE when no sm_100 GPU is usable.
 */
#ifndef VSB200_H_
#define VSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSB200_OK 0
#define VSB200_ERR_INVALID 1      /* bad argument / option (<-> OpenStreams false) */
#define VSB200_ERR_NO_DEVICE 2    /* CUDA device missing or kernel image not loadable */
#define VSB200_ERR_CUDA 3         /* CUDA runtime error (message via vsb200_last_error) */
#define VSB200_ERR_EMPTY 4        /* pop with nothing ready */
#define VSB200_ERR_UNSUPPORTED 5  /* option the reference has but this path does not build */
#define VSB200_ERR_CAPACITY 6     /* a result table would outgrow its hard limit (message names it) */

/* Mirrors DenseSegmentationOptions (segmentation/dense_segmentation.h:42-95) plus the
 * gflags that override it (segmentation/dense_segmentation.cpp:39-46,55-101). */
typedef struct vsb200_dense_opts {
  int32_t presmoothing;                  /* 0 none, 1 gaussian (unsupported), 2 bilateral (default) */
  float frac_min_region_size;            /* 0.01 */
  int32_t chunk_size;                    /* 20, >= 3 */
  float chunk_overlap_ratio;             /* 0.2 */
  int32_t num_constraint_frames;         /* 1 */
  int32_t two_stage_oversegment;         /* 0 (1 unsupported) */
  int32_t thin_structure_suppression;    /* 0 (1 unsupported: "does not work correctly" in the reference) */
  int32_t enforce_n4_connectivity;       /* 1 */
  int32_t enforce_spatial_connectedness; /* 1 */
  int32_t color_distance;                /* 0 L1, 1 L2 (default) */
  int32_t compute_vectorization;         /* 0 (1 unsupported) */
  int32_t device;                        /* CUDA device ordinal */
  int32_t want_id_maps;                  /* 1: every frame result also carries a host int32 id map */
} vsb200_dense_opts;

/* One frame's SegmentationDesc (segment_util/segmentation.proto:55-172) as flat arrays.
 * Layout is shared with the oracle's vso_frame_result. */
typedef struct vsb200_frame_result {
  int32_t width, height, chunk_id, chunk_size, overlap_start, hierarchy_frame_idx;
  int32_t connectedness;               /* 1 = N4_CONNECT, 2 = N8_CONNECT */
  int32_t n_regions;
  const int32_t* region_id;            /* [n_regions], Region2D.id */
  const int32_t* interval_offset;      /* [n_regions + 1] into intervals */
  const int32_t* intervals;            /* [3 * n_intervals]: y, left_x, right_x */
  const float* shape_moments;          /* [6 * n_regions]: size mean_x mean_y xx xy yy */
  int32_t n_compound;                  /* hierarchy level 0; > 0 only on the first frame of a chunk */
  const int32_t* compound;             /* [4 * n_compound]: id size start_frame end_frame */
  const int32_t* neighbor_offset;      /* [n_compound + 1] */
  const int32_t* neighbor_id;          /* sorted ascending per region when ids are constrained */
  int64_t pts;
} vsb200_frame_result;

typedef struct vsb200_dense vsb200_dense;

/* Fills the reference defaults. */
void vsb200_dense_default_opts(vsb200_dense_opts* o);
/* Last error text of the calling thread ("" if none). */
const char* vsb200_last_error(void);
/* Number of usable sm_100 devices (0 if none); never throws. */
int vsb200_device_count(void);

/* ---- streaming engine: replaces DenseSegmentationUnit::OpenStreams / ProcessFrame /
 *      PostProcess (segmentation/segmentation_unit.cpp:58-178) and everything below it
 *      (DenseSegmentation::ProcessFrame, dense_segmentation.cpp:108-162). ---- */
int vsb200_dense_create(const vsb200_dense_opts* o, int width, int height, int use_flow,
                        vsb200_dense** out);
/* One BGR24 frame (VideoFrame bytes, width_step = row_stride_bytes) and, when created with
 * use_flow, the backward flow (interleaved float x,y; NULL on frame 0).  *n_ready = number of
 * frames whose results became available (a whole chunk at a time). */
int vsb200_dense_push(vsb200_dense*, const uint8_t* bgr, int row_stride_bytes,
                      const float* flow_xy, int flow_row_stride_bytes, int64_t pts, int* n_ready);
/* Same as push for a frame that is already resident in device memory (on-device decoder,
 * bench.py's HBM-resident leg); no flow. */
int vsb200_dense_push_device(vsb200_dense*, const uint8_t* dev_bgr, int row_stride_bytes, int64_t pts,
                             int* n_ready);
/* == ProcessFrame(flush = true) / PostProcess (segmentation_unit.cpp:154-161). */
int vsb200_dense_flush(vsb200_dense*, int* n_ready);
/* Results in input order. */
int vsb200_dense_pop(vsb200_dense*, vsb200_frame_result* out);
/* Host id map (int32 [height * width]) of the most recently popped frame, or NULL
 * unless opts.want_id_maps. */
const int32_t* vsb200_dense_last_id_map(vsb200_dense*);
/* Serialises the most recently popped frame as a segmentation.SegmentationDesc protobuf
 * message (proto2 wire format, segment_util/segmentation.proto:55-172).  Returns the
 * byte count; copies at most cap bytes into buf. */
size_t vsb200_dense_last_proto(vsb200_dense*, uint8_t* buf, size_t cap);
/* The same encoder over caller-held result arrays (host only): protobuf's canonical serialisation of the frame's
 * SegmentationDesc, what SerializeToString writes (segment_util/segmentation_io.cpp:73-78).  Returns the size,
 * copying min(size, cap) bytes. */
size_t vsb200_encode_frame_proto(const vsb200_frame_result* r, uint8_t* buf, size_t cap);
/* Per-stage device/host milliseconds accumulated since creation:
 * [0] h2d+preprocess [1] edge build [2] sort [3] merge [4] labels+n4+rle [5] host shaping
 * [6] neighbours; and counters [7] kernels launched [8] merge rounds. */
void vsb200_dense_stats(vsb200_dense*, double out[9]);
/* Profiling taps (<-> the per-unit timers of video_framework/video_unit.cpp:181-217): CUDA-event
 * timing of every steady-state edge-build launch on the engine's stream; io_stats returns
 * [0] host->device bytes [1] device->host bytes [2] edge-build ms total [3] edge-build launches. */
void vsb200_dense_set_profiling(vsb200_dense*, int time_edge_kernel);
void vsb200_dense_io_stats(vsb200_dense*, double out[4]);
void vsb200_dense_destroy(vsb200_dense*);

/* Multi-GPU seam.  export_halo: region-id maps of the two overlap
 * frames a group holds right after a chunk boundary (overlap_segmentations_,
 * dense_segmentation.cpp:300-315), copied into caller-provided DEVICE buffers (int32 [h*w] each)
 * so the caller can ncclSend/ncclRecv them without host staging, plus the chain state
 * [0] max region id (max_region_id_, :360-365) [1] id of the chunk the maps constrain
 * [2] frames output so far.
 * import_halo: the successor's side.  Called on a fresh handle before its first push, it puts the
 * engine into the predecessor's post-boundary state; the first push must then be the frame of
 * dev_id_map_last (the predecessor's last pushed frame).  Results continue exactly as if one
 * handle had processed the whole sequence (pipelined seam: exact semantics). */
int vsb200_dense_export_halo(vsb200_dense*, int32_t* dev_id_map_prev_out, int32_t* dev_id_map_last_out,
                             int32_t chain_state[3]);
int vsb200_dense_import_halo(vsb200_dense*, const int32_t* dev_id_map_prev,
                             const int32_t* dev_id_map_last, const int32_t chain_state[3]);

/* ---- kernel-level entry points (device pointers; used by the parity tests, bench.py and
 *      the DenseSegGraphInterface adapter).  stream is a cudaStream_t (NULL = default). ---- */

/* u8 BGR -> f32/255 (+ bilateral): DenseSegmentation::PreprocessFeatures
 * (dense_segmentation.cpp:164-198), imagefilter::BilateralFilter (image_filter.cpp:184-277).
 * scratch: >= vsb200_preprocess_scratch_bytes(). */
size_t vsb200_preprocess_scratch_bytes(void);
int vsb200_preprocess(const uint8_t* dev_bgr, int row_stride_bytes, int width, int height,
                      int presmoothing, float* dev_out, void* dev_scratch, void* stream);
/* Spatio-temporal edge weights of one frame: AddSpatialEdgesImpl + AddTemporal[Flow]EdgesImpl
 * (dense_segmentation_graph.h:956-1142) with ColorDiff3L2 / L1 (pixel_distance.h:141-157).
 * spatial_out: [h][w][4] (R, B, BL, BR); temporal_out: [h][w][9] (TL,T,TR,L,C,R,BL,B,BR about
 * the flow-displaced clamped centre); missing edges hold -1.  prev / temporal_out / flow may
 * be NULL (first frame of a chunk: spatial only). */
int vsb200_edge_build(const float* dev_curr, const float* dev_prev, const float* dev_flow,
                      int width, int height, int l1, float* dev_spatial_out,
                      float* dev_temporal_out, void* stream);
/* Bucket index of a weight, FastSegmentationGraph::AddEdge (segmentation_graph.h:158-162). */
int vsb200_bucket_index(float weight);
/* Stable bucket sort of the edge lists of one chunk graph (replaces the insert-time bucketing
 * of segmentation_graph.h:158-162,367-374): keys are the 2048 weight buckets, order inside a
 * bucket is (bucket list, anchor pixel, direction).  seg_ptrs[q] = device weights of bucket
 * list q (NULL = absent), q even spatial ([h][w][4]), q odd temporal ([h][w][9]).
 * codes_out: >= total valid edges uint32 edge codes (rank of the edge in (list, pixel, direction) order: (q/2)*13N + (q odd ? 4N : 0) + pixel*nd + dir);
 * bucket_start_out: device uint32/uint64 [2049]. */
int vsb200_sort_edges(const float* const* host_seg_ptrs, int num_lists, int width, int height,
                      uint32_t* dev_codes_out, uint64_t* dev_bucket_start_out, void* dev_scratch,
                      size_t scratch_bytes, void* stream);
size_t vsb200_sort_scratch_bytes(int num_lists, int width, int height);
/* Whole-chunk over-segmentation of `slots` smoothed frames without constraints (graph build,
 * sort, FastSegmentationGraph::SegmentGraph, segmentation_graph.h:339-463, flatten): writes the
 * per-voxel region label (representative node id) to dev_labels_out [slots][h][w]. */
int vsb200_segment_chunk(const float* dev_frames, int width, int height, int slots, int l1,
                         int min_region_size, int32_t* dev_labels_out, double* stats4, void* stream);

/* K11 + K10 (csrc/shape.cu): N4 connected components of every label in every frame of a label volume and their shape
 * moments -- ConnectedComponents(raster, N4_CONNECT) (segment_util/segmentation_util.cpp:1007-1101) and
 * ShapeMomentsFromRasterization (:652-693) for all regions of all frames at once, as
 * DenseSegmentationGraph::EnforceSpatialConnectedness (dense_segmentation_graph.h:666-777) needs them.
 * dev_labels [slices][h][w].  Components are numbered in the order of their first scan interval over the volume.
 * dev_component_out (may be NULL) [slices][h][w]: component number of every pixel.  host_records_out (may be NULL): up
 * to record_cap records of 10 words: first interval, #intervals, label, slice, area (int32), mean_x, mean_y, xx, xy, yy
 * (float bits).  *n_components_out = number of components. */
int vsb200_label_components(const int32_t* dev_labels, int width, int height, int slices, int32_t* dev_component_out,
                            int32_t* host_records_out, int record_cap, int* n_components_out, void* stream);

/* ---- region stage, appearance descriptor (csrc/region_hist.cu) ---- */

/* cv::cvtColor(CV_BGR2Lab) on 8-bit data, AppearanceExtractor (segmentation/region_descriptor.cpp:59-89, :73).
 * dev_lab_out is dense [h][w][3].  Bit identical to OpenCV's integer path (cv2 4.13 over the whole cube). */
int vsb200_bgr2lab(const uint8_t* dev_bgr, int row_stride_bytes, int width, int height, uint8_t* dev_lab_out,
                   void* stream);
/* Per-region Lab histograms, AppearanceDescriptor3D::AddFeatures (region_descriptor.cpp:97-111) ->
 * ColorHistogram::AddPixelInterpolated / AddValueInterpolated (segmentation/histograms.cpp:140-211), for all
 * regions of a frame at once: reset once per chunk set, add once per frame (BGR frame + its int32 region-id
 * map, ids outside [0, n_regions) are skipped), finish = NormalizeToOne (histograms.cpp:340-360) into
 * dev_hist_out [n_regions][lum_bins * color_bins * color_bins] (bin = l * color_bins^2 + a * color_bins + b)
 * and the weight sums (pixel counts) into dev_weight_sum_out [n_regions] (may be NULL).
 * RegionSegmentationOptions defaults: luminance_bins 10, color_bins 20 (region_segmentation.h:60-61). */
size_t vsb200_region_hist_scratch_bytes(int n_regions, int lum_bins, int color_bins);
int vsb200_region_hist_reset(void* dev_scratch, int n_regions, int lum_bins, int color_bins, void* stream);
int vsb200_region_hist_add(const uint8_t* dev_bgr, int row_stride_bytes, const int32_t* dev_region_ids,
                           int width, int height, int n_regions, int lum_bins, int color_bins,
                           void* dev_scratch, void* stream);
int vsb200_region_hist_finish(const void* dev_scratch, int n_regions, int lum_bins, int color_bins,
                              float* dev_hist_out, float* dev_weight_sum_out, void* stream);
/* AppearanceDescriptor3D::RegionDistance = ColorHistogram::ChiSquareDist (histograms.cpp:391-407) for region
 * pairs dev_pairs [2 * n_pairs] (e.g. the neighbour pairs of the over-segmentation) -> dev_out [n_pairs]. */
int vsb200_hist_chisquare(const float* dev_hist, int total_bins, const int32_t* dev_pairs, int n_pairs,
                          float* dev_out, void* stream);

/* ---- frame-group sharding over the GPUs of a box (csrc/shard.cu) ----
 * One handle per rank.  NCCL is bound at run time (libnccl.so.2).  Replaces nothing in the reference (it has no
 * multi-device path); the exchanged state is DenseSegmentation's chunk hand-over: overlap_segmentations_
 * (dense_segmentation.cpp:300-328) and max_region_id_ (:360-365). */
typedef struct vsb200_shard vsb200_shard;
/* rank 0: ncclGetUniqueId; the 128 bytes go to the other ranks by any side channel. */
int vsb200_shard_unique_id(uint8_t id_out[128]);
int vsb200_shard_create(const uint8_t id[128], int rank, int world, int device, int width, int height, vsb200_shard** out);
/* C1 + C2 at a group boundary (d stands right after a chunk boundary): overlap id maps to rank + 1 / from rank - 1
 * (ncclSend / ncclRecv), all-gather of the region-id counts; id_offsets_out[world + 1] = exclusive prefix, last = total. */
int vsb200_shard_exchange(vsb200_shard* s, vsb200_dense* d, int64_t* id_offsets_out, int* have_pred);
/* the predecessor's two overlap id maps on the device ([2][h][w] int32), NULL on rank 0 / before an exchange */
const int32_t* vsb200_shard_pred_maps(vsb200_shard* s);
/* Seam vote on the device: own id map of the shared frame vs the predecessor's -> dev_table_out[n_ids]
 * (local id -> predecessor id with the largest overlap, ids born later -> id_offset + id). */
int vsb200_shard_vote(vsb200_shard* s, const int32_t* dev_own_first_map, int n_ids, int64_t id_offset, int32_t* dev_table_out);
/* dev_ids[i] = dev_table[dev_ids[i]] for ids in [0, n_ids) */
int vsb200_shard_relabel(vsb200_shard* s, int32_t* dev_ids, size_t n, const int32_t* dev_table, int n_ids);
/* out: exchange ms on the device (sum), exchanges, kernel launches */
void vsb200_shard_stats(vsb200_shard* s, double out[3]);
void vsb200_shard_destroy(vsb200_shard* s);

/* ---- hierarchical region stage (SURVEY section 8a config 3 + 8f N1; csrc/region_stage.cu) ----
 * Plug point: RegionSegmentationUnit (segmentation/segmentation_unit.cpp:180-331) -> RegionSegmentation::ProcessFrame
 * (segmentation/region_segmentation.cpp:97-205).  One handle per unit, driven by one thread. */
typedef struct vsb200_region_opts {       /* RegionSegmentationOptions, segmentation/region_segmentation.h:41-82 */
  int32_t min_region_num;                 /* 10 */
  int32_t max_region_num;                 /* 10000 */
  float level_cutoff_fraction;            /* 0.8 */
  float small_region_penalizer;           /* 0.25 */
  int32_t luminance_bins, color_bins, flow_bins;                       /* 10, 20, 16 */
  int32_t chunk_set_size, chunk_set_overlap, constraint_chunks;        /* 6, 2, 1 */
  int32_t save_descriptors;               /* 0; 1 -> VSB200_ERR_UNSUPPORTED */
  int32_t use_appearance, use_flow, use_size_penalizer;                /* 1, 1 (a flow stream is present), 1 */
  int32_t compute_vectorization;          /* 0; 1 -> VSB200_ERR_UNSUPPORTED */
  int32_t device;
} vsb200_region_opts;
typedef struct vsb200_region vsb200_region;

void vsb200_region_default_opts(vsb200_region_opts* o);
/* RegionSegmentation::RegionSegmentation (region_segmentation.cpp:47-95); its CHECKs become VSB200_ERR_INVALID. */
int vsb200_region_create(const vsb200_region_opts* o, int width, int height, vsb200_region** out);
/* ProcessFrame(false, desc, features, results): one frame of the over-segmentation (the arrays a dense handle popped;
 * n_compound > 0 marks the first frame of a dense chunk), its BGR24 host frame and, for flow streams, the frame's flow
 * field (NULL on the first frame).  *n_ready = hierarchical frame results that became ready. */
int vsb200_region_push(vsb200_region* r, const vsb200_frame_result* overseg, const uint8_t* bgr, int row_stride_bytes,
                       const float* flow_xy, int flow_row_stride_bytes, int* n_ready);
/* ProcessFrame(true, NULL, NULL, results) -- RegionSegmentationUnit::PostProcess. */
int vsb200_region_flush(vsb200_region* r, int* n_ready);
/* Next result as one flat int32 record (floats as bits), valid until the next pop; returns its length in words, 0 if
 * nothing is ready:  width height chunk_id chunk_size overlap_start hierarchy_frame_idx n_regions n_levels |
 * per region: id n_intervals (y left_x right_x)* 6 x shape moment | per hierarchy level: n_compound, per compound:
 * id size parent_id start_frame end_frame n_neighbors n_children neighbor_id* child_id*   (SegmentationDesc,
 * segment_util/segmentation.proto:55-172). */
long long vsb200_region_pop(vsb200_region* r, const int32_t** record);
/* out[0] = kernel launches so far, out[1] = chunk sets segmented */
void vsb200_region_stats(vsb200_region* r, double out[2]);
void vsb200_region_destroy(vsb200_region* r);

/* ---- result container (csrc/pb_io.cpp, host only) ----
 * The reference's segmentation file, segment_util/segmentation_io.cpp: "HEAD" int32 n, n x int32 | per chunk "CHNK"
 * int32 chunk_id, int32 n_frames, n x int64 absolute frame offsets, n x int64 pts, int64 offset of the next header,
 * then per frame "SEGD" int32 size, bytes | "TERM" int32 n_chunks.  Byte identical to SegmentationWriter for the
 * same sequence of calls.  Frame payloads are opaque: the proto2 wire bytes of vsb200_dense_last_proto, or the
 * stripped format below. */
typedef struct vsb200_seg_writer vsb200_seg_writer;
/* SegmentationWriter::OpenFile(header_entries) (segmentation_io.cpp:46-71); SegmentationWriterUnit passes {1, 0}
 * (segmentation_unit.cpp:366-369). */
int vsb200_seg_writer_open(const char* filename, const int32_t* header_entries, int n_entries, vsb200_seg_writer** out);
/* SegmentationWriter::AddSegmentationDataToChunk(data, pts) (:80-88). */
int vsb200_seg_writer_add(vsb200_seg_writer*, const uint8_t* data, size_t size, int64_t pts);
/* SegmentationWriter::AddSegmentationToChunk(desc, pts) (:73-78) for the frame most recently popped from `dense`:
 * its SegmentationDesc is serialised straight from the result arrays, no message objects in between. */
int vsb200_seg_writer_add_last_frame(vsb200_seg_writer*, vsb200_dense* dense, int64_t pts);
/* SegmentationWriter::WriteChunk (:90-143). */
int vsb200_seg_writer_write_chunk(vsb200_seg_writer*);
/* SegmentationWriter::WriteTermHeaderAndClose (:145-155): writes the pending chunk, "TERM", closes and frees. */
int vsb200_seg_writer_close(vsb200_seg_writer*);

typedef struct vsb200_seg_reader vsb200_seg_reader;
/* SegmentationReader::OpenFileAndReadHeaders (:166-229): walks the chunk headers, collects offsets and pts. */
int vsb200_seg_reader_open(const char* filename, vsb200_seg_reader** out);
int vsb200_seg_reader_num_frames(const vsb200_seg_reader*);
int vsb200_seg_reader_num_header_flags(const vsb200_seg_reader*);
const int32_t* vsb200_seg_reader_header_flags(const vsb200_seg_reader*);
const int64_t* vsb200_seg_reader_time_stamps(const vsb200_seg_reader*);
/* SeekToFrame + ReadNextFrameBinary (:253-274): returns the payload size of `frame`, copying min(size, cap) bytes;
 * 0 on a parse error. */
size_t vsb200_seg_reader_read(vsb200_seg_reader*, int frame, uint8_t* buf, size_t cap);
/* Same with a status: VSB200_OK and *size_out = payload size (0 is a legitimately empty frame), VSB200_ERR_INVALID on a
 * parse error / truncated file. */
int vsb200_seg_reader_read_frame(vsb200_seg_reader*, int frame, uint8_t* buf, size_t cap, size_t* size_out);
void vsb200_seg_reader_close(vsb200_seg_reader*);

/* StripToEssentials(desc, save_vectorization = false, save_shape_moments, &binary) (segmentation_io.cpp:311-443) from
 * the arrays of a frame result: returns the size, copying min(size, cap) bytes.  The vectorised variant the writer
 * unit uses needs compute_vectorization (not built). */
size_t vsb200_strip_to_essentials(const vsb200_frame_result* r, int save_shape_moments, uint8_t* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif  /* VSB200_H_ */
