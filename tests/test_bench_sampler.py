"""bench.py's clock sampler against a stand-in nvidia-smi (host logic; no GPU): the poller is up before the load starts,
samples are attributed to the load window / the timed region, throttle reasons are collected, and a missing nvidia-smi
yields an explicit 'unavailable' record instead of an exception."""
import importlib.util
import os
import stat
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("kta_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


FAKE = """#!/bin/bash
sleep 0.1
i=0
while true; do
  if [ $i -lt 3 ]; then clk=1200; else clk=1965; fi
  echo "0, $clk, 1965, 480.5, 0x0000000000000004, Not Active, Not Active, Not Active, Active"
  i=$((i+1))
  sleep 0.02
done
"""


def test_sampler_windows_and_reasons(tmp_path, monkeypatch):
    exe = tmp_path / "nvidia-smi"
    exe.write_text(FAKE)
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    bench = load_bench()
    s = bench.ClockSampler(0)
    s.start()
    s.wait_first()
    assert s.rows, "the poller delivers a line before the load starts"
    time.sleep(0.12)                       # "warm-up": a few samples, the early 1200 MHz ones among them
    t0 = time.perf_counter()
    time.sleep(0.10)                       # "timed region"
    t1 = time.perf_counter()
    out = s.stop(t0, t1)
    assert out["sm_max_mhz"] == 1965 and out["sm_mhz"] == 1965          # the median over the load window
    assert out["reasons"] == ["sw_power_cap"]
    assert 3 <= out["samples_in_region"] <= 8 and out["samples"] > out["samples_in_region"]


def test_sampler_without_nvidia_smi(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", str(tmp_path))
    bench = load_bench()
    s = bench.ClockSampler(0)
    s.start()
    s.wait_first(timeout=0.05)
    out = s.stop(0.0, 1.0)
    assert out["sm_mhz"] is None and out["reasons"] == ["nvidia-smi unavailable"]
