"""Generates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference and Python cv2 exist; the GPU box has neither the reference nor this need).

  real_clip_136x240x24.npz : 24 frames of the reference's only fixture
      (video_example/test_video.MOV, 272x480) decoded with cv2 and area-downsampled x0.5,
      uint8 BGR -- real-video statistics for the parity tests at a size the oracle
      finishes in seconds.
  cv2_thirdparty.npz : known answers for the third-party arithmetic on the path
      (SURVEY.md section 8c): cv::Mat::convertTo(CV_32FC3, 1/255), cv::copyMakeBorder
      (BORDER_REPLICATE, 4 px) and cv::minMaxLoc on a real frame, produced by cv2 4.13.
  oracle_pins.json : checksums of the oracle's own output on the fixtures (regression pin
      of the restatement; the reference itself ships no golden vectors).
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
MOV = "/root/reference/video_example/test_video.MOV"


def main():
    cap = cv2.VideoCapture(MOV)
    frames_full, frames = [], []
    for _ in range(24):
        ok, f = cap.read()
        assert ok
        frames_full.append(f)
        frames.append(cv2.resize(f, (136, 240), interpolation=cv2.INTER_AREA))
    clip = np.stack(frames).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "real_clip_136x240x24.npz"), frames=clip)

    src = np.ascontiguousarray(frames_full[0][100:164, 60:156])   # 64 x 96 crop of a real frame
    # cv2's Python binding does not expose Mat::convertTo; cv2.dnn.blobFromImage runs the same
    # 8U -> 32F conversion + float scale and is bit-identical to float32(u8) * float32(1/255)
    # for all 256 levels (checked below), which is what the oracle and the CUDA kernel compute.
    blob = cv2.dnn.blobFromImage(src, scalefactor=1.0 / 255.0, swapRB=False, crop=False)  # NCHW float32
    conv = np.ascontiguousarray(np.transpose(blob[0], (1, 2, 0)))
    lv = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2)
    lvb = cv2.dnn.blobFromImage(lv, scalefactor=1.0 / 255.0, swapRB=False, crop=False)
    assert np.array_equal(np.transpose(lvb[0], (1, 2, 0)), lv.astype(np.float32) * np.float32(1.0 / 255.0))
    border = cv2.copyMakeBorder(conv, 4, 4, 4, 4, cv2.BORDER_REPLICATE)
    mn, mx, _, _ = cv2.minMaxLoc(conv.reshape(conv.shape[0], -1))
    np.savez_compressed(os.path.join(HERE, "cv2_thirdparty.npz"), src=src, convert=conv, border=border,
                        minmax=np.array([mn, mx], np.float64))

    import oracle_binding as ob
    pins = {}
    o = ob.OracleDense(136, 240)
    res = []
    for f in clip:
        res += o.push(f)
    res += o.flush()
    h = hashlib.sha256()
    for r in res:
        h.update(ob.id_map_from_result(r).tobytes())
    pins["real_clip_id_maps_sha256"] = h.hexdigest()
    pins["real_clip_regions_per_frame"] = [int(r["region_id"].size) for r in res]
    sm = ob.preprocess(clip[0])
    pins["real_clip_frame0_smoothed_sha256"] = hashlib.sha256(sm.tobytes()).hexdigest()
    with open(os.path.join(HERE, "oracle_pins.json"), "w") as fh:
        json.dump(pins, fh, indent=1)
    print("wrote fixtures", clip.shape, conv.shape, pins["real_clip_regions_per_frame"][:5])


if __name__ == "__main__":
    main()
