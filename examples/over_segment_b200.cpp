// over_segment_b200.cpp -- the dense half of the reference's seg_tree_sample (`seg_tree_sample --over_segment`,
// seg_tree_sample/seg_tree.cpp:194-217,300-316: DenseSegmentationUnit -> SegmentationWriterUnit) on a B200, in the
// reference's own language: frames go through segmentation::B200DenseSegmentation (video_segment_b200/host, the class
// with DenseSegmentation's interface) and the SegmentationDesc stream is written with the reference's container layout
// through the C ABI's writer.  Decode is out of scope: the input is raw BGR24,
//     int32 width, int32 height, int32 frames, then frames x height x width x 3 bytes,
// which `python -c "import numpy, cv2; ..."` or ffmpeg -pix_fmt bgr24 produce.
// Usage: over_segment_b200 <in.bgr> <out.pb> [chunk_size]
// Builds against the reference tree (see oracle/Makefile, target _ref/over_segment_b200); there is no CPU fallback:
// without an sm_100 device the first frame aborts with the C ABI's error text.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <memory>
#include <vector>

#include <opencv2/core/core.hpp>

#include "b200_dense_segmentation.h"
#include "vsb200.h"

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <in.bgr> <out.pb> [chunk_size]\n", argv[0]);
    return 2;
  }
  FILE* in = fopen(argv[1], "rb");
  int32_t head[3];
  if (!in || fread(head, 4, 3, in) != 3 || head[0] < 2 || head[1] < 2 || head[2] < 0) {
    fprintf(stderr, "cannot read the header of %s\n", argv[1]);
    return 2;
  }
  const int width = head[0], height = head[1], frames = head[2];

  segmentation::DenseSegmentationOptions options;          // the reference's defaults (dense_segmentation.h:42-95)
  if (argc > 3) options.chunk_size = atoi(argv[3]);
  segmentation::B200DenseSegmentation dense(options, width, height);

  vsb200_seg_writer* writer = nullptr;
  const int32_t header_entries[2] = {1, 0};                // SegmentationWriterUnit::OpenStreams (segmentation_unit.cpp:366-369)
  if (vsb200_seg_writer_open(argv[2], header_entries, 2, &writer) != VSB200_OK) {
    fprintf(stderr, "%s\n", vsb200_last_error());
    return 2;
  }

  std::vector<uint8_t> frame((size_t)width * height * 3);
  std::vector<uint8_t> wire;
  long long written = 0;
  auto output = [&](std::vector<std::unique_ptr<segmentation::SegmentationDesc>>& results) {
    for (auto& desc : results) {
      // SegmentationWriter::AddSegmentationToChunk: with real protobuf this is desc->SerializeToString(&data); the
      // same bytes come from the C ABI's encoder over the message's fields, which keeps this file free of a protobuf
      // runtime dependency.
      std::vector<int32_t> region_id, interval_offset(1, 0), intervals, compound, neighbor_offset(1, 0), neighbor_id;
      std::vector<float> moments;
      for (const auto& r : desc->region()) {
        region_id.push_back(r.id());
        for (const auto& s : r.raster().scan_inter()) {
          intervals.push_back(s.y());
          intervals.push_back(s.left_x());
          intervals.push_back(s.right_x());
        }
        interval_offset.push_back((int32_t)(intervals.size() / 3));
        const auto& m = r.shape_moments();
        const float v[6] = {m.size(), m.mean_x(), m.mean_y(), m.moment_xx(), m.moment_xy(), m.moment_yy()};
        moments.insert(moments.end(), v, v + 6);
      }
      if (desc->hierarchy_size() > 0) {
        for (const auto& c : desc->hierarchy(0).region()) {
          const int32_t v[4] = {c.id(), c.size(), c.start_frame(), c.end_frame()};
          compound.insert(compound.end(), v, v + 4);
          for (int k = 0; k < c.neighbor_id_size(); ++k) neighbor_id.push_back(c.neighbor_id(k));
          neighbor_offset.push_back((int32_t)neighbor_id.size());
        }
      }
      vsb200_frame_result fr;
      fr.width = desc->frame_width(); fr.height = desc->frame_height(); fr.chunk_id = desc->chunk_id();
      fr.chunk_size = desc->chunk_size(); fr.overlap_start = desc->overlap_start();
      fr.hierarchy_frame_idx = desc->hierarchy_frame_idx(); fr.connectedness = (int32_t)desc->connectedness();
      fr.n_regions = (int32_t)region_id.size();
      fr.region_id = region_id.data(); fr.interval_offset = interval_offset.data(); fr.intervals = intervals.data();
      fr.shape_moments = moments.data();
      fr.n_compound = (int32_t)(compound.size() / 4);
      fr.compound = compound.data(); fr.neighbor_offset = neighbor_offset.data(); fr.neighbor_id = neighbor_id.data();
      fr.pts = written;
      wire.resize(vsb200_encode_frame_proto(&fr, nullptr, 0));
      vsb200_encode_frame_proto(&fr, wire.data(), wire.size());
      vsb200_seg_writer_add(writer, wire.data(), wire.size(), written);
      ++written;
    }
  };

  std::vector<std::unique_ptr<segmentation::SegmentationDesc>> results;
  for (int k = 0; k < frames; ++k) {
    if (fread(frame.data(), 1, frame.size(), in) != frame.size()) {
      fprintf(stderr, "%s is truncated at frame %d\n", argv[1], k);
      return 2;
    }
    std::vector<cv::Mat> features(1, cv::Mat(height, width, CV_8UC3, frame.data(), (size_t)width * 3));
    dense.ProcessFrame(false, &features, nullptr, &results);
    output(results);
  }
  dense.ProcessFrame(true, nullptr, nullptr, &results);
  output(results);
  fclose(in);
  if (vsb200_seg_writer_close(writer) != VSB200_OK) {
    fprintf(stderr, "%s\n", vsb200_last_error());
    return 2;
  }
  fprintf(stderr, "wrote %lld frames to %s (%lld kernel launches)\n", written, argv[2], dense.KernelLaunches());
  return written == frames ? 0 : 1;
}
