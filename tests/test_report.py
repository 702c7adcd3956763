"""Report formatter (src/main.rs:123-179) pinned to demo_output.png, and the C++ CLI host end to end."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI_DIR = os.path.join(ROOT, "kafka_topic_analyzer_b200", "csrc", "cli")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build():
    from kafka_topic_analyzer_b200 import build
    build()
    subprocess.run(["make", "-C", CLI_DIR], check=True, capture_output=True)


def test_report_matches_demo_output_png():
    _build()
    demo = json.load(open(os.path.join(GOLD, "demo_output.json")))
    rows = demo["rows"]
    total = sum(r["total"] for r in rows)
    inp = [demo["topic"], demo["scanning_took_s"], total, demo["earliest_message_s"], 0, demo["latest_message_s"],
           demo["largest_message"], demo["smallest_message"], demo["topic_size"], 0, 0, len(rows)]
    for r in rows:
        inp += [r["P"], r["start_offset"], r["end_offset"], r["total"], r["alive"], r["tombstones"], r["key_null"],
                r["key_non_null"], r["k_bytes"], r["v_bytes"]]
    out = subprocess.run([os.path.join(CLI_DIR, "report_golden")], input=" ".join(map(str, inp)), text=True,
                         capture_output=True, check=True).stdout
    lines = out.splitlines()
    # header block, literally as in the screenshot
    for want in ["=" * 120, "Calculating statistics...", "Topic global.trv_bulk.partner_import", "Scanning took: 416 seconds",
                 "Estimated Msg/s: 590221", "-" * 120, "Earliest Message: 2018-01-31 17:23:13 UTC",
                 "Latest Message: 2018-04-13 14:29:52 UTC", "Largest Message: 750 bytes", "Smallest Message: 139 bytes",
                 "Topic Size: 66434997213 bytes", "| K = Key, V = Value, P = Partition, Tmb = Tombstone(s), Sz = Size"]:
        assert want in lines, want
    # table rows: from the Total column on, character for character as in the screenshot (the first three
    # columns' headers changed after the screenshot was taken: "|< OS" → "< OS", src/main.rs:150)
    tail0 = "| 25056009 | 25056009 | 0   | 0.0000 | 0      | 25056009 | 6778805354 | 225504081 | 6553301273 | 9      | 261    | 270    |"
    tail8 = "| 20021871 | 20021871 | 0   | 0.0000 | 0      | 20021871 | 5432377054 | 180196839 | 5252180215 | 9      | 262    | 271    |"
    head = "| Total    | Alive    | Tmb | DR     | K Null | K !Null  | P-Bytes    | K-Bytes   | V-Bytes    | A K-Sz | A V-Sz | A M-Sz |"
    assert any(l.startswith("| P ") and l.endswith(head) for l in lines)
    assert any(l.startswith("| 0 | 0    | 112298537 ") and l.endswith(tail0) for l in lines)
    assert any(l.startswith("| 8 | 0    | 112332976 ") and l.endswith(tail8) for l in lines)
    seps = [l for l in lines if l.startswith("+-")]
    assert len(seps) == len(rows) + 2 and len(set(seps)) == 1     # a separator after every row


def test_chrono_display_with_nanoseconds():
    _build()
    # earliest_message keeps the construction clock (with ns) when no record is earlier (metric.rs:39,66-68)
    inp = "t 1 0 1700000000 123456789 0 0 0 0 1 7 0"
    out = subprocess.run([os.path.join(CLI_DIR, "report_golden")], input=inp, text=True, capture_output=True, check=True).stdout
    assert "Earliest Message: 2023-11-14 22:13:20.123456789 UTC" in out
    assert "Latest Message: 1970-01-01 00:00:00 UTC" in out
    assert "Alive keys: 7" in out
    out = subprocess.run([os.path.join(CLI_DIR, "report_golden")], input=inp.replace("123456789", "500000000"), text=True,
                         capture_output=True, check=True).stdout
    assert "Earliest Message: 2023-11-14 22:13:20.500 UTC" in out


def test_cli_requires_reference_flags():
    _build()
    exe = os.path.join(CLI_DIR, "kafka-topic-analyzer")
    r = subprocess.run([exe, "-t", "x"], capture_output=True, text=True)
    assert r.returncode == 2 and "--bootstrap-server <BOOTSTRAP_SERVER>" in r.stderr
    r = subprocess.run([exe, "--version"], capture_output=True, text=True)
    assert r.stdout.strip() == "Kafka Topic Analyzer 0.4.1"        # src/main.rs:35


@pytest.mark.gpu
@pytest.mark.parametrize("feed", ["push", "batch", "device"])
def test_cli_end_to_end_vs_oracle(feed):
    """configs[0] through the C++ host: 4 partitions, 100k messages, -c; the printed table equals the oracle."""
    from kafka_topic_analyzer_b200 import synth
    from parity import oracle_for
    _build()
    exe = os.path.join(CLI_DIR, "kafka-topic-analyzer")
    r = subprocess.run([exe, "-t", "demo", "-b", "localhost:9092", "-c", "--feed", feed, "--synthetic",
                        "n=100000,partitions=4,tombstone_per_10k=2000,distinct_keys=5000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    spec = synth.make_spec(100000, 4, tombstone_per_10k=2000, distinct_keys=5000)
    t = synth.fill_host(spec)
    o = oracle_for(t, count_alive_keys=True)
    lines = r.stdout.splitlines()
    assert "Subscribing to demo" in lines and "Starting message consumption..." in lines
    assert "Alive keys: %d" % o.scalar("sum_all_alive") in lines
    assert "Topic Size: %d bytes" % o.scalar("overall_size") in lines
    assert "Largest Message: %d bytes" % o.scalar("largest_message") in lines
    rows = [l for l in lines if l.startswith("| ") and l[2].isdigit()]
    assert len(rows) == 4
    for l in rows:
        c = [x.strip() for x in l.strip("|").split("|")]
        p = int(c[0])
        assert int(c[2]) == 25000
        assert [int(c[3]), int(c[4]), int(c[5])] == [o.counter("total", p), o.counter("alive", p), o.counter("tombstones", p)]
        assert c[6] == "%.4f" % o.dirty_ratio(p)
        assert [int(c[7]), int(c[8])] == [o.counter("key_null", p), o.counter("key_non_null", p)]
        assert [int(c[10]), int(c[11])] == [o.counter("key_size_sum", p), o.counter("value_size_sum", p)]
        assert int(c[9]) == int(c[10]) + int(c[11])
        assert [int(c[12]), int(c[13]), int(c[14])] == [o.avg("key_size_avg", p), o.avg("value_size_avg", p), o.avg("message_size_avg", p)]
