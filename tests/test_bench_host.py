"""Host-side pieces of bench.py that do not need a GPU: the nvidia-smi clock sampler's parsing / windowing and the rule that
`roofline.traffic` is only ever read from THIS round's committed ncu summary."""
import datetime
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


class _DoneProc:
    def terminate(self):
        pass

    def wait(self, timeout=None):
        return 0

    def kill(self):
        pass


def _sampler_with(tmp_path, lines, gpu=0):
    s = bench.ClockSampler(gpu)
    p = tmp_path / "smi.csv"
    p.write_text("\n".join(lines) + "\n")
    s.proc, s.path = _DoneProc(), str(p)
    return s


def _line(ts, idx, sm, power, cap="Not Active"):
    return "%s, %d, %d, 1965, %.2f, Not Active, Not Active, Not Active, %s" % (ts, idx, sm, power, cap)


def test_clock_sampler_keeps_the_samples_inside_the_timed_region(tmp_path):
    t = datetime.datetime(2026, 1, 2, 3, 4, 5)
    f = lambda dt: (t + datetime.timedelta(milliseconds=dt)).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]  # noqa: E731
    lines = [_line(f(0), 0, 1965, 250.0), _line(f(200), 0, 1400, 980.0, "Active"), _line(f(400), 0, 1320, 990.0, "Active"),
             _line(f(400), 1, 700, 100.0),            # another GPU's row is ignored
             "garbage line", _line(f(600), 0, 1965, 300.0)]
    got = _sampler_with(tmp_path, lines).stop(t + datetime.timedelta(milliseconds=100), t + datetime.timedelta(milliseconds=500))
    assert got["samples"] == 2 and got["sm_mhz"] == 1360.0 and got["sm_max_mhz"] == 1965.0
    assert got["reasons"] == ["sw_power_cap"] and got["window"] == "timed region" and got["power_w_max"] == 990.0


def test_clock_sampler_falls_back_to_samples_under_load_for_a_short_region(tmp_path):
    t = datetime.datetime(2026, 1, 2, 3, 4, 5)
    f = lambda dt: (t + datetime.timedelta(milliseconds=dt)).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]  # noqa: E731
    lines = [_line(f(0), 0, 1965, 200.0), _line(f(200), 0, 1500, 900.0), _line(f(400), 0, 1965, 210.0)]
    got = _sampler_with(tmp_path, lines).stop(t + datetime.timedelta(milliseconds=250), t + datetime.timedelta(milliseconds=300))
    assert got["samples"] == 1 and got["sm_mhz"] == 1500.0 and got["window"].startswith("warm-up")


def test_clock_sampler_without_nvidia_smi():
    got = bench.ClockSampler(0).stop()
    assert got["sm_mhz"] is None and got["reasons"] == ["nvidia-smi unavailable"]


def test_traffic_is_read_from_this_rounds_capture_only(monkeypatch):
    got = bench.profile_of_largest_gemm("mistral", "f16f8", True)
    assert got is not None and got["profile_round"] == bench.PROFILE_ROUND and got["traffic"] > 3.6e9
    assert "f16f8_wide" in got["note"]
    assert bench.profile_of_largest_gemm("xlmr", "f16f8", True) is None      # only the Mistral shape has a capture
    monkeypatch.setattr(bench, "PROFILE_ROUND", "r9")
    assert bench.profile_of_largest_gemm("mistral", "f16f8", True) is None   # a capture of another round is refused
