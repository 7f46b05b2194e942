// pb_io.cu -- the reference's segmentation container on the result stream.  Host code only (compiled
// by nvcc with the rest of the library so that libvsb200.so stays one object): writer and reader of the
// HEAD / CHNK / SEGD / TERM file of segment_util/segmentation_io.cpp, byte identical to SegmentationWriter for the
// same call sequence, and StripToEssentials over the arrays of a frame result.  Frame payloads come straight from
// the engine's wire encoder (vsb200_dense_last_proto): no SegmentationDesc objects are built on the way to disk.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/vsb200.h"
#include "common.cuh"

using vsb::set_error;

struct vsb200_seg_writer {
  FILE* f = nullptr;
  int32_t num_chunks = 0;                     // num_chunks_
  int64_t curr_offset = 0;                    // curr_offset_
  std::vector<std::string> chunk_buffer;      // chunk_buffer_
  std::vector<int64_t> file_offsets;          // file_offsets_
  std::vector<int64_t> time_stamps;           // time_stamps_
  int64_t total_frames = 0;
  bool failed = false;
  void put(const void* p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) failed = true;
  }
};

struct vsb200_seg_reader {
  FILE* f = nullptr;
  std::vector<int64_t> file_offsets, time_stamps;
  std::vector<int32_t> header_flags;
};

extern "C" {

int vsb200_seg_writer_open(const char* filename, const int32_t* header_entries, int n_entries, vsb200_seg_writer** out) {
  if (!filename || !out || n_entries < 0 || (n_entries > 0 && !header_entries)) { set_error("seg_writer_open: bad arguments"); return VSB200_ERR_INVALID; }
  FILE* f = fopen(filename, "wb");
  if (!f) { set_error("Could not open %s to write!", filename); return VSB200_ERR_INVALID; }      // segmentation_io.cpp:54-57
  vsb200_seg_writer* w = new vsb200_seg_writer;
  w->f = f;
  // header: "HEAD", entry count, entries (:62-68)
  w->put("HEAD", 4);
  const int32_t n = n_entries;
  w->put(&n, 4);
  w->put(header_entries, (size_t)n_entries * 4);
  w->curr_offset = 4 + 4 + (int64_t)n_entries * 4;
  *out = w;
  return VSB200_OK;
}

int vsb200_seg_writer_add(vsb200_seg_writer* w, const uint8_t* data, size_t size, int64_t pts) {
  if (!w || (!data && size)) return VSB200_ERR_INVALID;
  if (size > 0x7fffffffu) { set_error("seg_writer_add: frame larger than the int32 size field"); return VSB200_ERR_INVALID; }
  // offsets are relative to the chunk's payload until WriteChunk adds the header size (:80-88)
  w->file_offsets.push_back(w->curr_offset);
  w->chunk_buffer.emplace_back((const char*)data, size);
  w->curr_offset += (int64_t)size + 4 + 4;
  w->time_stamps.push_back(pts);
  return VSB200_OK;
}

int vsb200_seg_writer_add_last_frame(vsb200_seg_writer* w, vsb200_dense* dense, int64_t pts) {
  if (!w || !dense) return VSB200_ERR_INVALID;
  const size_t n = vsb200_dense_last_proto(dense, nullptr, 0);
  if (n == 0) { set_error("seg_writer_add_last_frame: no popped frame"); return VSB200_ERR_EMPTY; }
  std::string buf(n, '\0');
  vsb200_dense_last_proto(dense, (uint8_t*)&buf[0], n);
  return vsb200_seg_writer_add(w, (const uint8_t*)buf.data(), n, pts);
}

int vsb200_seg_writer_write_chunk(vsb200_seg_writer* w) {
  if (!w) return VSB200_ERR_INVALID;
  const int32_t num_frames = (int32_t)w->file_offsets.size();
  const int32_t chunk_id = w->num_chunks++;
  w->put("CHNK", 4);
  w->put(&chunk_id, 4);
  w->put(&num_frames, 4);
  // header size: tag, two int32, offsets and pts, offset of the next header (:101-105)
  const int64_t size_of_header = 4 + 2 * 4 + (int64_t)num_frames * 2 * 8 + 8;
  w->curr_offset += size_of_header;
  for (auto& o : w->file_offsets) o += size_of_header;
  w->put(w->file_offsets.data(), (size_t)num_frames * 8);
  w->put(w->time_stamps.data(), (size_t)num_frames * 8);
  w->put(&w->curr_offset, 8);
  for (const std::string& frame : w->chunk_buffer) {
    w->put("SEGD", 4);
    const int32_t frame_size = (int32_t)frame.size();
    w->put(&frame_size, 4);
    w->put(frame.data(), frame.size());
  }
  w->total_frames += num_frames;
  w->chunk_buffer.clear();
  w->file_offsets.clear();
  w->time_stamps.clear();
  if (w->failed) { set_error("seg_writer: write failed"); return VSB200_ERR_INVALID; }
  return VSB200_OK;
}

int vsb200_seg_writer_close(vsb200_seg_writer* w) {
  if (!w) return VSB200_ERR_INVALID;
  if (!w->chunk_buffer.empty()) vsb200_seg_writer_write_chunk(w);     // :146-148
  w->put("TERM", 4);
  w->put(&w->num_chunks, 4);
  const bool ok = fclose(w->f) == 0 && !w->failed;
  delete w;
  if (!ok) { set_error("seg_writer: write failed"); return VSB200_ERR_INVALID; }
  return VSB200_OK;
}

int vsb200_seg_reader_open(const char* filename, vsb200_seg_reader** out) {
  if (!filename || !out) return VSB200_ERR_INVALID;
  FILE* f = fopen(filename, "rb");
  if (!f) { set_error("Could not open segmentation file %s", filename); return VSB200_ERR_INVALID; }
  vsb200_seg_reader* r = new vsb200_seg_reader;
  r->f = f;
  long file_size = 0;
  if (fseek(f, 0, SEEK_END) == 0) file_size = ftell(f);
  fseek(f, 0, SEEK_SET);
  int32_t prev_header_id = -1;
  bool ok = true, said = false;
  while (ok) {                                                          // :178-226
    char tag[5] = {0, 0, 0, 0, 0};
    if (fread(tag, 1, 4, f) != 4) { ok = false; break; }
    if (strcmp(tag, "TERM") == 0) break;
    if (strcmp(tag, "HEAD") == 0) {
      int32_t n = 0;
      ok = fread(&n, 4, 1, f) == 1 && n >= 0 && n < (1 << 20);
      if (ok) {
        r->header_flags.resize(n);
        ok = n == 0 || fread(r->header_flags.data(), 4, n, f) == (size_t)n;
      }
      continue;
    }
    if (strcmp(tag, "CHNK") != 0) { set_error("Parsing error, expected chunk header at current offset. Found: %s", tag); ok = false; said = true; break; }
    int32_t header_id = 0, n = 0;
    ok = fread(&header_id, 4, 1, f) == 1 && header_id == prev_header_id + 1 && fread(&n, 4, 1, f) == 1 && n >= 0;
    // a chunk header lists n offsets and n time stamps (16 bytes per frame): a count the file cannot hold is corruption,
    // not a reason to let std::vector throw through the C boundary
    if (ok && (long long)n * 16 > (long long)file_size) ok = false;
    if (!ok) break;
    prev_header_id = header_id;
    const size_t base = r->file_offsets.size();
    r->file_offsets.resize(base + n);
    r->time_stamps.resize(base + n);
    int64_t next_header_pos = 0;
    ok = (n == 0 || (fread(r->file_offsets.data() + base, 8, n, f) == (size_t)n && fread(r->time_stamps.data() + base, 8, n, f) == (size_t)n)) &&
         fread(&next_header_pos, 8, 1, f) == 1 && fseek(f, (long)next_header_pos, SEEK_SET) == 0;
  }
  if (!ok) {
    fclose(f);
    delete r;
    if (!said) set_error("segmentation file %s is truncated or malformed", filename);
    return VSB200_ERR_INVALID;
  }
  *out = r;
  return VSB200_OK;
}

int vsb200_seg_reader_num_frames(const vsb200_seg_reader* r) { return r ? (int)r->file_offsets.size() : 0; }
int vsb200_seg_reader_num_header_flags(const vsb200_seg_reader* r) { return r ? (int)r->header_flags.size() : 0; }
const int32_t* vsb200_seg_reader_header_flags(const vsb200_seg_reader* r) { return r ? r->header_flags.data() : nullptr; }
const int64_t* vsb200_seg_reader_time_stamps(const vsb200_seg_reader* r) { return r ? r->time_stamps.data() : nullptr; }

size_t vsb200_seg_reader_read(vsb200_seg_reader* r, int frame, uint8_t* buf, size_t cap) {
  if (!r || frame < 0 || frame >= (int)r->file_offsets.size()) return 0;
  char tag[5] = {0, 0, 0, 0, 0};
  int32_t size = 0;
  if (fseek(r->f, (long)r->file_offsets[frame], SEEK_SET) != 0 || fread(tag, 1, 4, r->f) != 4 || strcmp(tag, "SEGD") != 0 ||
      fread(&size, 4, 1, r->f) != 1 || size < 0) {
    set_error("Expecting segmentation header. Error parsing file.");      // :258-261
    return 0;
  }
  const size_t n = (size_t)size < cap ? (size_t)size : cap;
  if (buf && n && fread(buf, 1, n, r->f) != n) return 0;
  return (size_t)size;
}

int vsb200_seg_reader_read_frame(vsb200_seg_reader* r, int frame, uint8_t* buf, size_t cap, size_t* size_out) {
  if (size_out) *size_out = 0;
  if (!r || !size_out || frame < 0 || frame >= (int)r->file_offsets.size()) { set_error("reader: no such frame"); return VSB200_ERR_INVALID; }
  char tag[5] = {0, 0, 0, 0, 0};
  int32_t size = 0;
  if (fseek(r->f, (long)r->file_offsets[frame], SEEK_SET) != 0 || fread(tag, 1, 4, r->f) != 4 || strcmp(tag, "SEGD") != 0 ||
      fread(&size, 4, 1, r->f) != 1 || size < 0) {
    set_error("Expecting segmentation header. Error parsing file.");      // :258-261
    return VSB200_ERR_INVALID;
  }
  const size_t n = (size_t)size < cap ? (size_t)size : cap;
  if (buf && n && fread(buf, 1, n, r->f) != n) { set_error("segmentation file is truncated inside frame %d", frame); return VSB200_ERR_INVALID; }
  *size_out = (size_t)size;
  return VSB200_OK;
}

void vsb200_seg_reader_close(vsb200_seg_reader* r) {
  if (!r) return;
  fclose(r->f);
  delete r;
}

size_t vsb200_strip_to_essentials(const vsb200_frame_result* r, int save_shape_moments, uint8_t* buf, size_t cap) {
  if (!r) return 0;
  std::string s;
  auto put32 = [&s](int32_t v) { s.append((const char*)&v, 4); };
  auto put16 = [&s](int16_t v) { s.append((const char*)&v, 2); };
  put32(r->width);                                                       // segmentation_io.cpp:317-321
  put32(r->height);
  put32(r->n_regions);                                                   // :339-340 (no vectorisation block)
  for (int k = 0; k < r->n_regions; ++k) {
    put32(r->region_id[k]);
    put32(r->interval_offset[k + 1] - r->interval_offset[k]);            // :363-375: int16 y, left, right
    for (int i = r->interval_offset[k]; i < r->interval_offset[k + 1]; ++i) {
      put16((int16_t)r->intervals[3 * i]);
      put16((int16_t)r->intervals[3 * i + 1]);
      put16((int16_t)r->intervals[3 * i + 2]);
    }
    if (save_shape_moments) {                                            // :378-391: floats truncated to int
      for (int m = 0; m < 6; ++m) put32((int32_t)r->shape_moments[6 * k + m]);
    }
  }
  const int32_t hierarchy_size = r->n_compound > 0 ? 1 : 0;              // :395-437
  put32(hierarchy_size);
  if (hierarchy_size) {
    put32(r->n_compound);
    for (int k = 0; k < r->n_compound; ++k) {
      put32(r->compound[4 * k]);          // id
      put32(r->compound[4 * k + 1]);      // size
      put32(-1);                          // parent_id: unset on the top (only) level -> proto default -1
      put32(0);                           // no children in the over-segmentation
      put32(r->compound[4 * k + 2]);      // start_frame
      put32(r->compound[4 * k + 3]);      // end_frame
    }
  }
  if (buf && cap) memcpy(buf, s.data(), s.size() < cap ? s.size() : cap);
  return s.size();
}

}  // extern "C"
