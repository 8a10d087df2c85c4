"""Generates the committed golden vectors under tests/golden/ from the UNMODIFIED reference voldata
sources (oracle/_ref/libvoldata_ref.so, built by oracle/Makefile from /root/reference) and from the
reference's data assets. Run in the build container only (needs /root/reference):

    PYTHONPATH=. python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.binding import VoldataRef  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def synth_cases():
    """Deterministic small dense grids covering the edge cases of grid_brick.cpp."""
    rng = np.random.default_rng(1234)
    cases = {}
    v = (rng.random((20, 33, 70)) * 255).astype(np.uint8)
    v[rng.random(v.shape) < 0.5] = 0
    cases["ragged_70x33x20"] = (v, 0.0, 1.0)
    v = np.zeros((72, 80, 96), np.uint8)
    v[10:30, 30:50, 20:40] = (rng.random((20, 20, 20)) * 255).astype(np.uint8)
    v[60:64, 5:9, 70:72] = 200
    cases["sparse_negmin_96x80x72"] = (v, -1.5, 7.25)     # negative minorant: sign-extension quirk of encode_range
    v = np.full((40, 40, 40), 128, np.uint8)
    v[10, 10, 10] = 129
    cases["fp16_collapse_40"] = (v, 1000.0, 1000.5)       # fp16 range collapse: value_norm divides by zero
    cases["all_empty_16"] = (np.zeros((16, 16, 16), np.uint8), 0.0, 1.0)
    cases["single_voxel_1x1x4"] = (np.array([1, 2, 5, 10], np.uint8).reshape(4, 1, 1), 0.0, 10.0)
    v = (rng.random((64, 64, 64)) * 255).astype(np.uint8)
    cases["full_64"] = (v, 0.25, 3.0)
    z, y, x = np.mgrid[0:48, 0:56, 0:130].astype(np.float32)
    blob = np.exp(-(((x - 60) / 30) ** 2 + ((y - 28) / 14) ** 2 + ((z - 24) / 12) ** 2))
    cases["smooth_blob_130x56x48"] = ((np.clip(blob - 0.2, 0, 1) * 255).astype(np.uint8), 0.0, 2.0)
    return cases


def main():
    ref = VoldataRef()
    out = {}
    for name, (vox, lo, hi) in synth_cases().items():
        g, dec = ref.brick_build(vox, lo, hi, decode=True)
        out[name + ".n_bricks"] = np.array(g.n_bricks, np.uint32)
        out[name + ".atlas_dim"] = np.array(g.atlas_dim, np.uint32)
        out[name + ".brick_count"] = np.array([g.brick_count], np.uint64)
        out[name + ".indirection"] = g.indirection
        out[name + ".range"] = g.range
        out[name + ".atlas_sha1"] = np.frombuffer(hashlib.sha1(g.atlas.tobytes()).digest(), np.uint8)
        for i in range(3):
            out[name + f".mip{i}"] = g.mips[i]
        out[name + ".decode_sha1"] = np.frombuffer(hashlib.sha1(dec.tobytes()).digest(), np.uint8)
    # half conversion: every fp16 tie point and neighbours, plus specials
    fl = []
    for h in range(0, 0x7c00):
        a = np.float64(np.array([h], np.uint16).view(np.float16)[0])
        b = np.float64(np.array([h + 1], np.uint16).view(np.float16)[0]) if h + 1 < 0x7c00 else 65520.0
        mid = np.float32((a + b) / 2)
        fl += [mid, np.nextafter(mid, np.float32(0)), np.nextafter(mid, np.float32(1e9))]
    fl += [0.0, -0.0, 1e-10, -1e-10, 5.96e-8, 2.98e-8, 2.9802322e-8, 65504.0, 65519.9, 65520.0, 1e9, np.inf, -np.inf]
    fl = np.array(fl, np.float32)
    fl = np.concatenate([fl, -fl])
    out["half.inputs"] = fl
    out["half.outputs"] = np.array([ref.to_half(f) for f in fl], np.uint16)
    # DenseGrid(float*) quantiser
    rng = np.random.default_rng(99)
    d = (rng.random((9, 13, 21)).astype(np.float32) * 3 - 0.5)
    q, mm = ref.dense_from_float(d)
    out["dense.input"], out["dense.u8"], out["dense.minmax"] = d, q, np.array(mm, np.float32)
    d2 = -rng.random((4, 5, 6)).astype(np.float32)       # all-negative: max stays FLT_MIN (grid_dense.cpp:61)
    q2, mm2 = ref.dense_from_float(d2)
    out["dense_neg.input"], out["dense_neg.u8"], out["dense_neg.minmax"] = d2, q2, np.array(mm2, np.float32)
    np.savez_compressed(os.path.join(HERE, "voldata_golden.npz"), **out)
    # files WRITTEN by the reference's own cereal serialisation (serialization.cpp:36-43,66-80) for the host / Python readers
    os.makedirs(os.path.join(HERE, "ref_written"), exist_ok=True)
    vox = (np.random.default_rng(77).random((6, 5, 7)) * 255).astype(np.uint8)
    ref.dense_write(vox, -1.25, 3.5, os.path.join(HERE, "ref_written", "ref_7x5x6.dense"))
    vb, lo, hi = synth_cases()["ragged_70x33x20"]
    ref.brick_roundtrip_write(vb, lo, hi, os.path.join(HERE, "ref_written", "ref_ragged_70x33x20.brick"))
    np.savez_compressed(os.path.join(HERE, "ref_written", "expected.npz"), dense_vox=vox, dense_minmax=np.array([-1.25, 3.5], np.float32))
    # TransferFunction::colormap with the reference's tinycolormap: every type at 256 bins, Turbo / Viridis also at 7 and 1000
    cm = {f"type{t}_256": ref.colormap_lut(t, 256) for t in range(14)}
    for t in (3, 9):
        cm[f"type{t}_7"], cm[f"type{t}_1000"] = ref.colormap_lut(t, 7), ref.colormap_lut(t, 1000)
    np.savez_compressed(os.path.join(HERE, "colormap_golden.npz"), **cm)
    # smoke.brick as parsed by the reference's own cereal loader: hashes of every buffer
    s = ref.brick_load(os.path.join(HERE, "assets", "smoke.brick"))
    np.savez_compressed(
        os.path.join(HERE, "smoke_brick_golden.npz"),
        n_bricks=np.array(s.n_bricks, np.uint32), atlas_dim=np.array(s.atlas_dim, np.uint32),
        brick_count=np.array([s.brick_count], np.uint64), min_maj=np.array(s.min_maj, np.float32), transform=s.transform,
        indirection_sha1=np.frombuffer(hashlib.sha1(s.indirection.tobytes()).digest(), np.uint8),
        range_sha1=np.frombuffer(hashlib.sha1(s.range.tobytes()).digest(), np.uint8),
        atlas_sha1=np.frombuffer(hashlib.sha1(s.atlas.tobytes()).digest(), np.uint8),
        mips_sha1=np.stack([np.frombuffer(hashlib.sha1(m.tobytes()).digest(), np.uint8) for m in s.mips]),
        decode_sha1=np.frombuffer(hashlib.sha1(s.decode_all().tobytes()).digest(), np.uint8),
    )
    # the loose end-to-end anchor: imgs/example.jpg (README command, 4096 spp) box-filtered to 64x64
    try:
        import cv2
        img = cv2.imread("/root/reference/imgs/example.jpg")[:, :, ::-1]
        small = cv2.resize(img, (64, 64), interpolation=cv2.INTER_AREA)
        np.save(os.path.join(HERE, "example_64x64_rgb8.npy"), small)
    except Exception as e:  # pragma: no cover
        print("skipping example.jpg golden:", e)
    print("golden vectors written")


if __name__ == "__main__":
    main()
