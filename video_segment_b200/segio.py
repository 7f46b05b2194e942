"""Host-side mirror of the reference's segmentation container classes (segment_util/segmentation_io.h:59-191) over
the C ABI (csrc/pb_io.cu): SegmentationWriter / SegmentationReader with the reference's method names in snake case,
and strip_to_essentials.  File layout and call semantics are the reference's; payloads are opaque bytes (the proto2
wire bytes the engine encodes straight from its result arrays, or the stripped format)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

from ._lib import FrameResult, check, lib


class SegmentationWriter:
    """segmentation_io.h:59-106."""

    def __init__(self, filename: str):
        self.filename = filename
        self._h = C.c_void_p()

    def open_file(self, header_entries: Sequence[int] = ()) -> bool:
        arr = (C.c_int32 * max(1, len(header_entries)))(*header_entries)
        rc = lib().vsb200_seg_writer_open(self.filename.encode(), arr, len(header_entries), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
        return rc == 0

    def add_segmentation_data_to_chunk(self, data: bytes, pts: int = 0) -> None:
        check(lib().vsb200_seg_writer_add(self._h, data, len(data), pts), "vsb200_seg_writer_add")

    def add_segmentation_to_chunk(self, unit, pts: int = 0) -> None:
        """The frame most recently popped from a DenseSegmentationUnit, serialised from its result arrays."""
        check(lib().vsb200_seg_writer_add_last_frame(self._h, unit._h, pts), "vsb200_seg_writer_add_last_frame")

    def write_chunk(self) -> None:
        check(lib().vsb200_seg_writer_write_chunk(self._h), "vsb200_seg_writer_write_chunk")

    def write_term_header_and_close(self) -> None:
        if self._h:
            h, self._h = self._h, C.c_void_p()
            check(lib().vsb200_seg_writer_close(h), "vsb200_seg_writer_close")


class SegmentationReader:
    """segmentation_io.h:108-170 (binary access; parsing the payload is the caller's protobuf)."""

    def __init__(self, filename: str):
        self.filename = filename
        self._h = C.c_void_p()
        self._curr = 0

    def open_file_and_read_headers(self) -> bool:
        rc = lib().vsb200_seg_reader_open(self.filename.encode(), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
        return rc == 0

    def num_frames(self) -> int:
        return lib().vsb200_seg_reader_num_frames(self._h)

    def remaining_frames(self) -> int:
        return self.num_frames() - self._curr

    def get_header_flags(self) -> List[int]:
        n = lib().vsb200_seg_reader_num_header_flags(self._h)
        p = lib().vsb200_seg_reader_header_flags(self._h)
        return [p[i] for i in range(n)]

    def time_stamps(self) -> List[int]:
        p = lib().vsb200_seg_reader_time_stamps(self._h)
        return [p[i] for i in range(self.num_frames())]

    def seek_to_frame(self, frame: int) -> None:
        if not 0 <= frame < self.num_frames():
            raise IndexError("Requested frame out of bound.")
        self._curr = frame

    def read_next_frame_binary(self) -> Optional[bytes]:
        n = C.c_size_t()
        if lib().vsb200_seg_reader_read_frame(self._h, self._curr, None, 0, C.byref(n)) != 0:
            return None                           # parse error (vsb200_last_error says which)
        buf = C.create_string_buffer(max(1, n.value))
        got = C.c_size_t()
        if lib().vsb200_seg_reader_read_frame(self._h, self._curr, buf, n.value, C.byref(got)) != 0 or got.value != n.value:
            return None
        self._curr += 1
        return buf.raw[:n.value]                  # b"" is a legitimately empty frame

    def close_file(self) -> None:
        if self._h:
            lib().vsb200_seg_reader_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close_file()
        except Exception:
            pass


def strip_to_essentials(result: FrameResult, save_shape_moments: bool = False) -> bytes:
    """StripToEssentials(desc, save_vectorization=False, save_shape_moments) (segmentation_io.cpp:311-443) from a frame
    result structure (as popped from the engine)."""
    n = lib().vsb200_strip_to_essentials(C.byref(result), int(save_shape_moments), None, 0)
    buf = C.create_string_buffer(max(1, n))
    lib().vsb200_strip_to_essentials(C.byref(result), int(save_shape_moments), buf, n)
    return buf.raw[:n]


def encode_frame_proto(result: FrameResult) -> bytes:
    """Wire bytes of the frame's SegmentationDesc (== SegmentationDesc::SerializeToString) from a frame result."""
    n = lib().vsb200_encode_frame_proto(C.byref(result), None, 0)
    buf = C.create_string_buffer(max(1, n))
    lib().vsb200_encode_frame_proto(C.byref(result), buf, n)
    return buf.raw[:n]
