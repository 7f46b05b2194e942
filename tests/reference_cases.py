"""The clips / option sets on which the oracle is pinned to the compiled reference (oracle/_ref/libref_results.so):
shared by tests/golden/make_reference_golden.py (writes the reference's digests) and tests/test_oracle_cpu.py."""
from __future__ import annotations

import hashlib
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name -> (clip spec, flow sigma or None, options).  Chunk sizes keep the reference's own precondition of 2 overlap
# frames (dense_segmentation.cpp:58-62); between them the cases cover unconstrained and constrained chunks, a flush
# inside the first chunk, both colour distances, every result-shaping switch, flow-displaced temporal edges.
CASES = {
    "real_default_chunk20": (("real", 24), None, dict()),
    "real_chunk8": (("real", 24), None, dict(chunk_size=8)),
    "real_chunk5_overlap04": (("real", 14), None, dict(chunk_size=5, chunk_overlap_ratio=0.4)),
    "real_l1": (("real", 12), None, dict(chunk_size=8, color_distance=0)),
    "real_l1_single_chunk_plain": (("real", 10), None, dict(color_distance=0, enforce_n4_connectivity=0, enforce_spatial_connectedness=0)),
    "real_single_chunk": (("real", 12), None, dict()),
    "real_no_n4": (("real", 12), None, dict(chunk_size=8, enforce_n4_connectivity=0)),
    "real_no_connectedness": (("real", 12), None, dict(chunk_size=8, enforce_spatial_connectedness=0)),
    "real_no_presmoothing": (("real", 12), None, dict(chunk_size=8, presmoothing=0)),
    "real_min_region_003": (("real", 12), None, dict(chunk_size=8, frac_min_region_size=0.03)),
    "real_two_constraint_frames": (("real", 24), None, dict(chunk_size=10, chunk_overlap_ratio=0.3, num_constraint_frames=2)),
    "real_flow": (("real", 24), 2.0, dict(chunk_size=10)),
    "synth_chunk9": (("synth", 7, 96, 72, 22), None, dict(chunk_size=9)),
    "synth_flow": (("synth", 7, 96, 72, 22), 4.0, dict(chunk_size=9)),
    "synth_three_frames": (("synth", 7, 96, 72, 3), None, dict(chunk_size=9)),
    "synth_one_frame": (("synth", 7, 96, 72, 1), None, dict(chunk_size=9)),
    "synth_exact_chunk": (("synth", 11, 64, 48, 17), None, dict(chunk_size=8)),
}


def load_case(name):
    from video_segment_b200.synth import synth_clip
    spec, flow_sigma, opts = CASES[name]
    if spec[0] == "real":
        clip = np.load(os.path.join(_ROOT, "tests", "golden", "real_clip_136x240x24.npz"))["frames"][:spec[1]]
    else:
        clip = synth_clip(*spec[1:])
    flows = None
    if flow_sigma is not None:
        t, h, w, _ = clip.shape
        flows = np.random.default_rng(3).normal(0, flow_sigma, (t, h, w, 2)).astype(np.float32)
    return np.ascontiguousarray(clip), flows, opts


def run_stream(engine_cls, clip, flows, opts):
    """Streams a clip through an engine with the OracleDense call shape; returns the per-frame result dicts."""
    t, h, w, _ = clip.shape
    e = engine_cls(w, h, use_flow=flows is not None, **opts)
    out = []
    for k, f in enumerate(clip):
        out += e.push(f, None if flows is None or k == 0 else flows[k])
    out += e.flush()
    e.close()
    return out


_FIELDS = ("region_id", "interval_offset", "intervals", "shape_moments", "compound", "neighbor_offset", "neighbor_id")
_SCALARS = ("width", "height", "chunk_id", "chunk_size", "overlap_start", "hierarchy_frame_idx", "connectedness")


def digest(results) -> str:
    """SHA-256 over every field of every frame result (header scalars, region ids, scan intervals, shape moments as
    float bits, hierarchy level 0 with neighbours)."""
    h = hashlib.sha256()
    for r in results:
        h.update(np.asarray([r[k] for k in _SCALARS], np.int32).tobytes())
        for k in _FIELDS:
            a = np.ascontiguousarray(r[k])
            h.update(np.asarray(a.shape, np.int32).tobytes())
            h.update(a.tobytes())
    return h.hexdigest()


def first_difference(ref, mine):
    """None if two result streams agree in every field, else a short description of the first mismatch."""
    if len(ref) != len(mine):
        return f"{len(ref)} vs {len(mine)} frames"
    for k, (a, b) in enumerate(zip(ref, mine)):
        for key in _SCALARS:
            if a[key] != b[key]:
                return f"frame {k} {key}: {a[key]} vs {b[key]}"
        for key in _FIELDS:
            x, y = a[key], b[key]
            if x.shape != y.shape or not np.array_equal(x.view(np.uint32) if x.dtype.kind == "f" else x,
                                                        y.view(np.uint32) if y.dtype.kind == "f" else y):
                return f"frame {k} {key}: shapes {x.shape} {y.shape}"
    return None


# ---- randomised cases (differential testing of the oracle against the compiled reference) ----

def random_case(seed):
    """A small random clip (noise / gradient + noise / moving blocks / speckles), random options inside the reference's
    own preconditions (>= 2 overlap frames; bilateral filter needs >= 8 rows), optional random flow."""
    rng = np.random.default_rng(seed)
    w = int(rng.integers(9, 56)); h = int(rng.integers(8, 44)); t = int(rng.integers(1, 30))
    kind = rng.integers(0, 4)
    if kind == 0:      # pure noise
        clip = rng.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
    elif kind == 1:    # low-amplitude noise over a gradient (many force-merge buckets)
        gx = np.linspace(0, 200, w)[None, None, :, None]
        clip = np.clip(gx + rng.normal(0, 6, (t, h, w, 3)), 0, 255).astype(np.uint8)
    elif kind == 2:    # moving blocks
        clip = np.zeros((t, h, w, 3), np.uint8) + rng.integers(0, 256, 3, dtype=np.uint8)
        for k in range(t):
            for b in range(4):
                x0 = (b * 7 + k * (b + 1)) % max(1, w - 5); y0 = (b * 5 + k) % max(1, h - 5)
                clip[k, y0:y0 + 5, x0:x0 + 6] = [40 * b + 30, 255 - 50 * b, 90 + 30 * b]
        clip = np.clip(clip + rng.normal(0, 3, clip.shape), 0, 255).astype(np.uint8)
    else:              # flat with a few speckles
        clip = np.full((t, h, w, 3), 128, np.uint8)
        idx = rng.integers(0, clip.size, clip.size // 50)
        clip.reshape(-1)[idx] = rng.integers(0, 256, idx.size, dtype=np.uint8)
    chunk = int(rng.integers(3, 13))
    ratio = float(rng.choice([0.2, 0.3, 0.4, 0.5, 0.7]))
    if min(int(ratio * chunk + 0.5), 2) < 2 or min(int(ratio * chunk + 0.5), 2) >= chunk:
        chunk, ratio = 8, 0.2
    opts = dict(chunk_size=chunk, chunk_overlap_ratio=ratio,
                num_constraint_frames=int(rng.integers(1, 3)),
                enforce_n4_connectivity=int(rng.integers(0, 2)), enforce_spatial_connectedness=int(rng.integers(0, 2)),
                color_distance=int(rng.integers(0, 2)), presmoothing=int(rng.choice([0, 2])),
                frac_min_region_size=float(rng.choice([0.01, 0.05, 0.1, 0.2])))
    if opts["presmoothing"] == 2 and h < 8:
        opts["presmoothing"] = 0
    flows = None
    if rng.integers(0, 2):
        flows = rng.normal(0, float(rng.choice([0.5, 3.0, 20.0])), (t, h, w, 2)).astype(np.float32)
    return np.ascontiguousarray(clip), flows, opts
