// TEST INFRASTRUCTURE ONLY.  Stand-in for the protoc output of the reference's segment_util/segmentation.proto
// (protoc / libprotobuf are not installed): plain value classes with protobuf's generated accessor names (x(), set_x(),
// has_x(), clear_x(), mutable_x(), add_x(), x_size(), x(i), CopyFrom, MergeFrom, Swap, Clear), field for field after
// segmentation.proto:55-172.  In-memory only: no wire format here (the wire encoder under test is the product's own).
#ifndef VSO_REF_SHIM_SEGMENTATION_PB_H_
#define VSO_REF_SHIM_SEGMENTATION_PB_H_
#include <cstdlib>
#include <memory>
#include <string>
#include <utility>
#include <google/protobuf/repeated_field.h>

#define VSO_PB_SCALAR(type, name, def)                                  \
 public:                                                                \
  type name() const { return name##_; }                                 \
  void set_##name(type v) { name##_ = v; has_##name##_ = true; }        \
  bool has_##name() const { return has_##name##_; }                     \
  void clear_##name() { name##_ = def; has_##name##_ = false; }         \
 private:                                                               \
  type name##_ = def;                                                   \
  bool has_##name##_ = false;

#define VSO_PB_REPEATED_SCALAR(type, name)                                            \
 public:                                                                              \
  type name(int i) const { return name##_.Get(i); }                                   \
  void set_##name(int i, type v) { name##_.Set(i, v); }                               \
  void add_##name(type v) { name##_.Add(v); }                                         \
  int name##_size() const { return name##_.size(); }                                  \
  void clear_##name() { name##_.Clear(); }                                            \
  const ::google::protobuf::RepeatedField<type>& name() const { return name##_; }     \
  ::google::protobuf::RepeatedField<type>* mutable_##name() { return &name##_; }      \
 private:                                                                             \
  ::google::protobuf::RepeatedField<type> name##_;

#define VSO_PB_REPEATED_MSG(type, name)                                               \
 public:                                                                              \
  const type& name(int i) const { return name##_.Get(i); }                            \
  type* mutable_##name(int i) { return name##_.Mutable(i); }                          \
  type* add_##name() { return name##_.Add(); }                                        \
  int name##_size() const { return name##_.size(); }                                  \
  void clear_##name() { name##_.Clear(); }                                            \
  const ::google::protobuf::RepeatedPtrField<type>& name() const { return name##_; }  \
  ::google::protobuf::RepeatedPtrField<type>* mutable_##name() { return &name##_; }   \
 private:                                                                             \
  ::google::protobuf::RepeatedPtrField<type> name##_;

// optional sub-message: held by value-semantics pointer; name() of an unset field is the default instance
#define VSO_PB_MSG(type, name)                                                        \
 public:                                                                              \
  const type& name() const { static const type d; return name##_ ? *name##_ : d; }    \
  type* mutable_##name() { if (!name##_) name##_.reset(new type()); return name##_.get(); } \
  bool has_##name() const { return (bool)name##_; }                                   \
  void clear_##name() { name##_.reset(); }                                            \
 private:                                                                             \
  ::vso_pb::Opt<type> name##_;

// every message: protobuf's value semantics
#define VSO_PB_MESSAGE(cls)                                                           \
 public:                                                                              \
  cls() {}                                                                            \
  void CopyFrom(const cls& o) { if (this != &o) *this = o; }                          \
  void Swap(cls* o) { cls t(*o); *o = *this; *this = t; }                             \
  void Clear() { *this = cls(); }                                                     \
  static const cls& default_instance() { static const cls d; return d; }

namespace vso_pb {
// deep-copying optional holder
template <class T> class Opt {
 public:
  Opt() {}
  Opt(const Opt& o) : p_(o.p_ ? new T(*o.p_) : nullptr) {}
  Opt& operator=(const Opt& o) { if (this != &o) p_.reset(o.p_ ? new T(*o.p_) : nullptr); return *this; }
  explicit operator bool() const { return (bool)p_; }
  T& operator*() const { return *p_; }
  T* get() const { return p_.get(); }
  void reset(T* t = nullptr) { p_.reset(t); }
 private:
  std::unique_ptr<T> p_;
};
}  // namespace vso_pb

namespace segmentation {

class RegionFeatures {
  VSO_PB_MESSAGE(RegionFeatures)
  VSO_PB_SCALAR(unsigned int, id, 0)
 public:
  void MergeFrom(const RegionFeatures& o) { if (o.has_id()) set_id(o.id()); }
};

class SegmentationDesc_Rasterization_ScanInterval {
  VSO_PB_MESSAGE(SegmentationDesc_Rasterization_ScanInterval)
  VSO_PB_SCALAR(int, y, 0)
  VSO_PB_SCALAR(int, left_x, 0)
  VSO_PB_SCALAR(int, right_x, 0)
};

class SegmentationDesc_Rasterization {
  VSO_PB_MESSAGE(SegmentationDesc_Rasterization)
  VSO_PB_REPEATED_MSG(SegmentationDesc_Rasterization_ScanInterval, scan_inter)
 public:
  typedef SegmentationDesc_Rasterization_ScanInterval ScanInterval;
  void MergeFrom(const SegmentationDesc_Rasterization& o) { scan_inter_.MergeFrom(o.scan_inter_); }
};

class SegmentationDesc_ShapeMoments {
  VSO_PB_MESSAGE(SegmentationDesc_ShapeMoments)
  VSO_PB_SCALAR(float, size, 0)
  VSO_PB_SCALAR(float, mean_x, 0)
  VSO_PB_SCALAR(float, mean_y, 0)
  VSO_PB_SCALAR(float, moment_xx, 0)
  VSO_PB_SCALAR(float, moment_xy, 0)
  VSO_PB_SCALAR(float, moment_yy, 0)
};

class SegmentationDesc_VectorMesh {
  VSO_PB_MESSAGE(SegmentationDesc_VectorMesh)
  VSO_PB_REPEATED_SCALAR(float, coord)
};

class SegmentationDesc_Polygon {
  VSO_PB_MESSAGE(SegmentationDesc_Polygon)
  VSO_PB_REPEATED_SCALAR(int, coord_idx)
  VSO_PB_SCALAR(bool, hole, false)
};

class SegmentationDesc_Vectorization {
  VSO_PB_MESSAGE(SegmentationDesc_Vectorization)
  VSO_PB_REPEATED_MSG(SegmentationDesc_Polygon, polygon)
};

class SegmentationDesc_Region2D {
  VSO_PB_MESSAGE(SegmentationDesc_Region2D)
  VSO_PB_SCALAR(int, id, 0)
  VSO_PB_MSG(SegmentationDesc_Rasterization, raster)
  VSO_PB_MSG(SegmentationDesc_ShapeMoments, shape_moments)
  VSO_PB_MSG(SegmentationDesc_Vectorization, vectorization)
};

class SegmentationDesc_CompoundRegion {
  VSO_PB_MESSAGE(SegmentationDesc_CompoundRegion)
  VSO_PB_SCALAR(int, id, 0)
  VSO_PB_SCALAR(int, size, 0)
  VSO_PB_REPEATED_SCALAR(int, neighbor_id)
  VSO_PB_SCALAR(int, parent_id, -1)
  VSO_PB_REPEATED_SCALAR(int, child_id)
  VSO_PB_SCALAR(int, start_frame, 0)
  VSO_PB_SCALAR(int, end_frame, 0)
};

class SegmentationDesc_HierarchyLevel {
  VSO_PB_MESSAGE(SegmentationDesc_HierarchyLevel)
  VSO_PB_REPEATED_MSG(SegmentationDesc_CompoundRegion, region)
 public:
  void MergeFrom(const SegmentationDesc_HierarchyLevel& o) { region_.MergeFrom(o.region_); }
};

enum SegmentationDesc_Connectedness { SegmentationDesc_Connectedness_N4_CONNECT = 1, SegmentationDesc_Connectedness_N8_CONNECT = 2 };

class SegmentationDesc {
  VSO_PB_MESSAGE(SegmentationDesc)
 public:
  typedef SegmentationDesc_Rasterization Rasterization;
  typedef SegmentationDesc_ShapeMoments ShapeMoments;
  typedef SegmentationDesc_VectorMesh VectorMesh;
  typedef SegmentationDesc_Polygon Polygon;
  typedef SegmentationDesc_Vectorization Vectorization;
  typedef SegmentationDesc_Region2D Region2D;
  typedef SegmentationDesc_CompoundRegion CompoundRegion;
  typedef SegmentationDesc_HierarchyLevel HierarchyLevel;
  typedef SegmentationDesc_Connectedness Connectedness;
  static const Connectedness N4_CONNECT = SegmentationDesc_Connectedness_N4_CONNECT;
  static const Connectedness N8_CONNECT = SegmentationDesc_Connectedness_N8_CONNECT;
  VSO_PB_REPEATED_MSG(SegmentationDesc_Region2D, region)
  VSO_PB_REPEATED_MSG(SegmentationDesc_HierarchyLevel, hierarchy)
  VSO_PB_SCALAR(int, frame_width, 0)
  VSO_PB_SCALAR(int, frame_height, 0)
  VSO_PB_SCALAR(int, chunk_size, 0)
  VSO_PB_SCALAR(int, overlap_start, 0)
  VSO_PB_SCALAR(int, chunk_id, -1)
  VSO_PB_SCALAR(int, hierarchy_frame_idx, 0)
  VSO_PB_REPEATED_MSG(RegionFeatures, features)
  VSO_PB_MSG(SegmentationDesc_VectorMesh, vector_mesh)
  VSO_PB_SCALAR(SegmentationDesc_Connectedness, connectedness, SegmentationDesc_Connectedness_N4_CONNECT)
  VSO_PB_SCALAR(bool, rasterization_removed, false)
 public:
  // No wire format in this stand-in (header comment): the reference's file IO is exercised through its byte-level
  // entry points (AddSegmentationDataToChunk / ReadNextFrameBinary); the message-level ones abort if reached.
  bool SerializeToString(std::string*) const { std::abort(); }
  bool ParseFromString(const std::string&) { std::abort(); }
};

}  // namespace segmentation
#endif
