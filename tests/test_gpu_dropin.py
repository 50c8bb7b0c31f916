"""The compiled drop-in (integration/algo_b200.cpp: poppy::morph_images with the reference's cv::Mat signature, built
against the reference's own src/algo.hpp) under the reference's own frame loop: oracle/_ref/poppy_dropin runs
poppy::morph<Sink>() (reference src/poppy.hpp:46-248: extractor, matcher, gabor_filter, the chain recurrence) twice from
the dumped union images - once with the stock morph_images body, once with the stub - and compares every frame."""
import glob
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "poppy_dropin")
DUMPS = os.path.join(ROOT, "oracle", "_ref", "full")


def _env():
    env = dict(os.environ)
    libs = [os.path.join(ROOT, "poppy_b200")]
    for p in sys.path:                                   # libcudart of the CUDA wheels, as the Python path loads it
        libs += glob.glob(os.path.join(p, "nvidia", "cuda_runtime", "lib"))
    libs += ["/usr/local/cuda/lib64"]
    env["LD_LIBRARY_PATH"] = ":".join(libs + [env.get("LD_LIBRARY_PATH", "")])
    return env


def _replay(config, frames):
    if not os.path.exists(BIN) or not os.path.exists(os.path.join(DUMPS, f"c{config}", "meta.txt")):
        pytest.skip("oracle/_ref/poppy_dropin or the reference dump is not shipped (oracle/build_ref_full.sh)")
    r = subprocess.run([BIN, "replay", os.path.join(DUMPS, f"c{config}"), "both", str(frames)], env=_env(),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
    return json.loads(lines[-1]), r.returncode


def test_dropin_under_the_reference_frame_loop_config1(native_lib):
    """square -> circle, 512x512, 193 matcher points, 60 chained frames, 64 levels: all of configs[0]."""
    out, rc = _replay(1, 60)
    assert out["frames"] == 60 and out["differing_bytes"] == 0 and rc == 0, out


def test_dropin_under_the_reference_frame_loop_config3_prefix(native_lib):
    """cat -> dog --autoalign at 1920x1080: the first 24 frames of a 24-frame run of the reference loop (the reference
    needs ~0.5 s per 1080p frame; the whole 120-frame chain is covered by tests/test_gpu_configs.py)."""
    out, rc = _replay(3, 24)
    assert out["frames"] == 24 and out["differing_bytes"] == 0 and rc == 0, out


def _full(config, frames):
    d = os.path.join(DUMPS, f"c{config}")
    if not os.path.exists(BIN) or not os.path.exists(os.path.join(d, "raw.txt")):
        pytest.skip("oracle/_ref/poppy_dropin or the raw inputs of the dump are not shipped (oracle/build_ref_full.sh, `raw`)")
    r = subprocess.run([BIN, "full", d, str(frames)], env=_env(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=1200)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, (r.returncode, r.stdout[-500:], r.stderr[-1500:])
    return json.loads(lines[-1]), r.returncode


def test_whole_pipeline_with_all_three_gpu_entry_points_config1(native_lib):
    """The reference's whole pipeline from the raw images (blur_margin, extractor, matcher, gabor_filter, frame loop) with
    blur_margin + gabor_filter (integration/util_b200.cpp) + morph_images (integration/algo_b200.cpp) served by the library:
    the union canvases and all 60 frames equal the all-reference run."""
    out, rc = _full(1, 60)
    assert out["frames"] == 60 and out["differing_bytes"] == 0 and out["canvases_differing"] == 0 and rc == 0, out
    assert out["b200_blur_margin_calls"] == 2 and out["b200_gabor_filter_calls"] == 1, out


def test_whole_pipeline_with_all_three_gpu_entry_points_config3_prefix(native_lib):
    """cat -> dog at 1920x1080 with real margins (1080-wide inputs on a 1920-wide canvas), 16 frames."""
    out, rc = _full(3, 16)
    assert out["frames"] == 16 and out["differing_bytes"] == 0 and out["canvases_differing"] == 0 and rc == 0, out
