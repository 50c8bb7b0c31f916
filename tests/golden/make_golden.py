"""Regenerates tests/golden/*.npz from the *reference library* (oracle/_ref/libpoppy_ref.so = unmodified reference
code + vendored OpenCV 4.6.0). Run in the build container, where /root/reference exists:

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

Each frame fixture stores the inputs and every stage boundary of poppy::morph_images();
the chain fixture stores the frames of the reference frame loop; the topology fixture stores point sets and the
triangle index lists produced by cv::Subdiv2D 4.6.0 + get_triangle_indices. The files pin the CPU restatement
(oracle/poppy_oracle.cpp), the host topology stage and the CUDA path where the reference library is unavailable."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from poppy_b200 import synth  # noqa: E402

FRAME_CASES = {
    # name: (inputs, shape ratio, mask ratio, pyramid levels)
    "noise_96x72_s037_L6": (synth.make_inputs(96, 72, 40, 6.0, seed=21), 0.37, 0.37, 6),
    "noise_61x45_s050_L64": (synth.make_inputs(61, 45, 16, 5.0, seed=22), 0.5, 0.5, 64),
    "noise_128x80_s090_L4": (synth.make_inputs(128, 80, 60, 8.0, seed=23), 0.9, 0.9, 4),
    "shapes_112x96_s000_L64": (synth.shape_inputs(112, 96, 32, seed=24), 0.0, 0.0, 64),
    "shapes_112x96_s025_L5": (synth.shape_inputs(112, 96, 32, seed=24), 0.25, 0.25, 5),
    "shapes_112x96_s100_L64": (synth.shape_inputs(112, 96, 32, seed=24), 1.0, 1.0, 64),
    "blocks_90x70_s010_L3": (synth.block_inputs(90, 70, 24, seed=51), 0.1, 0.1, 3),
    "blocks_90x70_s095_L64": (synth.block_inputs(90, 70, 24, seed=52), 0.95, 0.95, 64),
}


def main():
    assert ref.available(), "build oracle/_ref first"
    for name, (inp, s, m, levels) in FRAME_CASES.items():
        st = ref.stages(inp.bgr1, inp.bgr2, inp.gabor2, inp.pts1, inp.pts2, s, m, levels)
        direct, mp = ref.morph_images(inp.bgr1, inp.bgr2, inp.gabor2, inp.pts1, inp.pts2, s, m, levels)
        assert (direct == st.dst).all() and (mp == st.morphed_points).all()
        sharpened = int((np.abs(st.dst.astype(int) - np.clip(np.rint(st.lap_blend * 255), 0, 255)) > 0).any(axis=2).sum())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), bgr1=inp.bgr1, bgr2=inp.bgr2, gabor2=inp.gabor2,
                            pts1=inp.pts1, pts2=inp.pts2, shape=s, mask_ratio=m, levels=levels,
                            **{k: getattr(st, k) for k in st.__dataclass_fields__})
        print(name, "T=%d" % len(st.tri_idx), "unsharp-branch pixels ~%d" % sharpened)

    inp = synth.shape_inputs(80, 64, 24, seed=31)
    frames, pts = ref.chain(inp.bgr1, inp.bgr2, inp.gabor2, inp.pts1, inp.pts2, 12, 64)
    np.savez_compressed(os.path.join(HERE, "chain_shapes_80x64_N12_L64.npz"), bgr1=inp.bgr1, bgr2=inp.bgr2,
                        gabor2=inp.gabor2, pts1=inp.pts1, pts2=inp.pts2, n_frames=12, levels=64, frames=frames,
                        points=pts)
    print("chain", frames.shape)

    rng = np.random.default_rng(41)
    topo = {}
    specs = [("uniform", 300, 211, 180), ("grid", 97, 71, 120), ("dupes", 160, 120, 90), ("lattice", 400, 300, 200),
             ("tiny", 16, 12, 5), ("corners_only", 50, 40, 0)]
    for i, (kind, w, h, n) in enumerate(specs):
        if kind == "uniform":
            p = np.stack([rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)], 1)
        elif kind == "grid":
            p = np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1)
        elif kind == "dupes":
            p = np.stack([rng.uniform(-4, w + 0.5, n), rng.uniform(-4, h + 0.5, n)], 1)
            p[p[:, 0] >= w, 0] = w - 1
            p[p[:, 1] >= h, 1] = h - 1
            p[::4] = p[1::4][: len(p[::4])]
        elif kind == "lattice":
            p = np.stack([rng.integers(0, 9, n) * (w // 9), rng.integers(0, 9, n) * (h // 9)], 1)
        else:
            p = np.stack([rng.uniform(1, w - 2, n), rng.uniform(1, h - 2, n)], 1).reshape(-1, 2)
        p = np.concatenate([p, [[0, 0], [w - 1, 0], [0, h - 1], [w - 1, h - 1]]]).astype(np.float32)
        topo[f"pts_{i}"] = p
        topo[f"size_{i}"] = np.array([w, h])
        topo[f"tri_{i}"] = ref.triangulate(w, h, p)
        print("topology", kind, len(p), "->", len(topo[f"tri_{i}"]))
    np.savez_compressed(os.path.join(HERE, "topology.npz"), n_cases=len(specs), **topo)


if __name__ == "__main__":
    main()
