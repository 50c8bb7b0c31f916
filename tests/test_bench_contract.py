"""bench.py bookkeeping that needs no GPU: the SURVEY.md section 8(d) byte model and the roofline objects."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_match_the_survey_table():
    b = _bench()
    assert b.algorithmic_bytes_per_frame(1920, 1080, 6) == 138_905_040
    assert b.algorithmic_bytes_per_frame(3840, 2160, 6) == 555_588_480
    assert b.algorithmic_bytes_per_frame(7680, 4320, 6) == 2_222_322_240


def test_dominant_kernel_roofline_fields():
    b = _bench()
    kern = [{"kernel": "blend_collapse", "ms_per_step": 57.2, "share": 0.30, "launches_per_step": 133, "us_per_launch": 430.0},
            {"kernel": "raster_warp", "ms_per_step": 50.2, "share": 0.26, "launches_per_step": 19, "us_per_launch": 2642.0}]
    traffic = {"source": "test", "bytes_per_frame": {"blend_collapse": 4.2e8, "raster_warp": 7.6e7}}
    d = b.dominant_kernel_roofline(kern, 3840, 2160, 600, 32, 6532.2, traffic, 6)
    assert d["kernel"] == "blend_collapse" and d["unit"] == "GB/s"
    assert 0 < d["frac"] < 1 and abs(d["achieved"] / d["peak"] - d["frac"]) < 1e-12
    # 7 launches of a 32-frame chunk rated together; traffic scaled to the same frames
    frames = 600 * 7 / 133
    assert abs(d["traffic"] - 4.2e8 * frames) < 1e3
    assert d["algorithmic_bytes_per_launch"] > 0.9 * d["traffic"] * 0.9
    d2 = b.dominant_kernel_roofline(kern[1:], 3840, 2160, 600, 32, 6532.2, traffic, 6)
    assert d2["kernel"] == "raster_warp" and d2["algorithmic_bytes_per_launch"] == int((8 * 3840 * 2160 + 8 * 3840 * 2160 / 32) * 600 / 19)


def test_committed_traffic_file_is_well_formed():
    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert set(t["bytes_per_frame"]) >= {"raster_warp", "pyr_down", "blend_collapse", "unsharp_store"}
    total = sum(t["bytes_per_frame"].values())
    assert 555_588_480 < total < 2 * 555_588_480          # above the algorithmic model, well under 2x


def test_issue_roofline_object():
    b = _bench()
    t = b.ncu_traffic()
    r = b.issue_roofline(t, 3000.0, {"sm_mhz": 1965.0})
    assert r["warp_instructions_per_frame"] == sum(t["warp_instructions_per_frame"].values())
    assert abs(r["roofline_frames_per_s"] * r["frac"] - 3000.0) < 1e-6
    assert 0.3 < r["frac"] < 1.0
