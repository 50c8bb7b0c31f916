"""Size-reduced fixtures of BASELINE.json configs 1-3 from the *whole* unmodified reference pipeline.

    bash oracle/build_ref_full.sh                                   # builds oracle/_ref/poppy_ref_full (reference TUs + vendored OpenCV)
    cd /root/reference/src && for c in 1 2 3; do \
        /root/repo/oracle/_ref/poppy_ref_full dump $c /root/reference/images /root/repo/oracle/_ref/full/c$c $([ $c = 3 ] && echo some || echo all);
        /root/repo/oracle/_ref/poppy_ref_full raw $c /root/reference/images /root/repo/oracle/_ref/full/c$c; done
    python tests/golden/make_golden_full.py                         # -> tests/golden/full/c{1,2,3}.npz

The dumps under oracle/_ref/full/ (git-ignored, shipped to the GPU box) hold what poppy::morph<Sink>() handed to
morph_images() at src/poppy.hpp:215 - corrected1/2, gabor2 and the point sets of the first call (after blur_margin,
extractor, matcher / face landmarks / autoalign, gabor_filter, add_corners), the per-frame shape/mask ratios, every
frame's morphedPoints and checksum, and the frames themselves. The committed .npz keep the inputs (configs 1-2; config 3
keeps only the points: its 1080p images live in the dump), all morphed points, all frame checksums and a few frames."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
DUMPS = os.path.join(ROOT, "oracle", "_ref", "full")


def load_dump(d):
    meta = {}
    for line in open(os.path.join(d, "meta.txt")):
        k, _, v = line.strip().partition(" ")
        meta[k] = v
    w, h, n, N = int(meta["width"]), int(meta["height"]), int(meta["points"]), int(meta["frames"])
    rd = lambda name, dt: np.fromfile(os.path.join(d, name), dtype=dt)
    out = dict(meta=meta, w=w, h=h, n=n, frames=N, levels=int(meta["levels"]),
               pts1=rd("pts1.f32", np.float32).reshape(n, 2), pts2=rd("pts2.f32", np.float32).reshape(n, 2),
               ratios=rd("ratios.f64", np.float64).reshape(N, 3), morphed=rd("morphed.f32", np.float32).reshape(N, n, 2),
               hashes=rd("hashes.u64", np.uint64), kept=[int(k) for k in meta["kept_frames"].split()])
    out["image"] = lambda name: rd(name + ".u8", np.uint8).reshape(h, w, 3)
    out["gabor2"] = lambda: rd("gabor2.f32", np.float32).reshape(h, w, 3)
    out["frame"] = lambda j: rd("frame_%04d.u8" % j, np.uint8).reshape(h, w, 3)
    return out


def main():
    os.makedirs(os.path.join(HERE, "full"), exist_ok=True)
    for c in (1, 2, 3):
        d = load_dump(os.path.join(DUMPS, f"c{c}"))
        N = d["frames"]
        keep = sorted({0, 1, N // 2, N - 1})
        z = dict(config=c, width=d["w"], height=d["h"], levels=d["levels"], n_frames=N, pts1=d["pts1"], pts2=d["pts2"],
                 ratios=d["ratios"], morphed=d["morphed"], hashes=d["hashes"], cpu_features=d["meta"].get("cpu_features", ""),
                 image_a=d["meta"]["image_a"], image_b=d["meta"]["image_b"])
        if c != 3:
            z.update(corrected1=d["image"]("corrected1"), corrected2=d["image"]("corrected2"), gabor2=d["gabor2"](),
                     kept=np.array(keep), kept_frames=np.stack([d["frame"](j) for j in keep]))
        np.savez_compressed(os.path.join(HERE, "full", f"c{c}.npz"), **z)
        print(c, d["w"], d["h"], d["n"], N, os.path.getsize(os.path.join(HERE, "full", f"c{c}.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
