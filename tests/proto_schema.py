"""Builds the segmentation.SegmentationDesc message class at run time (no protoc in the image)
from a hand-written FileDescriptorProto that restates segment_util/segmentation.proto:55-172
(only the fields the dense path emits plus their neighbours)."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_T = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=_T.LABEL_OPTIONAL, type_name=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    return f


def segmentation_desc_class():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "segmentation_restated.proto"
    fd.package = "segmentation"
    fd.syntax = "proto2"
    sd = fd.message_type.add()
    sd.name = "SegmentationDesc"
    ras = sd.nested_type.add(); ras.name = "Rasterization"
    si = ras.nested_type.add(); si.name = "ScanInterval"
    _field(si, "y", 1, _T.TYPE_INT32, _T.LABEL_REQUIRED)
    _field(si, "left_x", 2, _T.TYPE_INT32, _T.LABEL_REQUIRED)
    _field(si, "right_x", 3, _T.TYPE_INT32, _T.LABEL_REQUIRED)
    _field(ras, "scan_inter", 1, _T.TYPE_MESSAGE, _T.LABEL_REPEATED, ".segmentation.SegmentationDesc.Rasterization.ScanInterval")
    sm = sd.nested_type.add(); sm.name = "ShapeMoments"
    for i, n in enumerate(["size", "mean_x", "mean_y", "moment_xx", "moment_xy", "moment_yy"]):
        _field(sm, n, i + 1, _T.TYPE_FLOAT)
    r2 = sd.nested_type.add(); r2.name = "Region2D"
    _field(r2, "id", 1, _T.TYPE_INT32, _T.LABEL_REQUIRED)
    _field(r2, "raster", 3, _T.TYPE_MESSAGE, type_name=".segmentation.SegmentationDesc.Rasterization")
    _field(r2, "shape_moments", 5, _T.TYPE_MESSAGE, type_name=".segmentation.SegmentationDesc.ShapeMoments")
    cr = sd.nested_type.add(); cr.name = "CompoundRegion"
    _field(cr, "id", 1, _T.TYPE_INT32, _T.LABEL_REQUIRED)
    _field(cr, "size", 2, _T.TYPE_INT32, _T.LABEL_REQUIRED)
    _field(cr, "neighbor_id", 3, _T.TYPE_INT32, _T.LABEL_REPEATED)
    f = _field(cr, "parent_id", 4, _T.TYPE_INT32); f.default_value = "-1"
    _field(cr, "child_id", 5, _T.TYPE_INT32, _T.LABEL_REPEATED)
    _field(cr, "start_frame", 6, _T.TYPE_INT32)
    _field(cr, "end_frame", 7, _T.TYPE_INT32)
    hl = sd.nested_type.add(); hl.name = "HierarchyLevel"
    _field(hl, "region", 2, _T.TYPE_MESSAGE, _T.LABEL_REPEATED, ".segmentation.SegmentationDesc.CompoundRegion")
    _field(sd, "region", 2, _T.TYPE_MESSAGE, _T.LABEL_REPEATED, ".segmentation.SegmentationDesc.Region2D")
    _field(sd, "hierarchy", 3, _T.TYPE_MESSAGE, _T.LABEL_REPEATED, ".segmentation.SegmentationDesc.HierarchyLevel")
    _field(sd, "frame_width", 4, _T.TYPE_INT32)
    _field(sd, "frame_height", 5, _T.TYPE_INT32)
    _field(sd, "chunk_size", 6, _T.TYPE_INT32)
    _field(sd, "overlap_start", 7, _T.TYPE_INT32)
    f = _field(sd, "chunk_id", 8, _T.TYPE_INT32); f.default_value = "-1"
    _field(sd, "hierarchy_frame_idx", 9, _T.TYPE_INT32)
    _field(sd, "connectedness", 12, _T.TYPE_INT32)   # enum Connectedness {N4_CONNECT = 1, N8_CONNECT = 2}: same wire type
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("segmentation.SegmentationDesc"))
