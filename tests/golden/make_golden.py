#!/usr/bin/env python3
"""Writes the golden fixtures of tests/golden/.  Nothing here runs the reference (no Rust toolchain
exists in this image); the vectors come from

  fnv_kat.json      following /root/reference/src/fnv32.rs:92-101 by hand (pure-Python loop below,
                    independent of oracle/ and of the CUDA code), next to standard FNV-1a-32 so that a
                    test can assert the two DIFFER (the reference multiplies by 0x811c9dc5, fnv32.rs:97);
  demo_output.json  hand transcription of /root/reference/demo_output.png (README.md:27-28), the only
                    real output of the reference that ships with it: 10 table rows printed by
                    src/main.rs:153-171 + the header block of src/main.rs:125-137.

Run:  python tests/golden/make_golden.py
"""
import calendar
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_fnv(b: bytes) -> int:          # src/fnv32.rs:79-81, 92-101
    h = 0x811C9DC5
    for x in b:
        h ^= x
        h = (h * 0x811C9DC5) & 0xFFFFFFFF
    return h


def std_fnv1a(b: bytes) -> int:        # the published FNV-1a-32 (prime 0x01000193), for contrast only
    h = 0x811C9DC5
    for x in b:
        h ^= x
        h = (h * 0x01000193) & 0xFFFFFFFF
    return h


KEYS = [b"", b"a", b"b", b"foobar", b"\x00", b"\xff", b"key-0", b"key-1", bytes(range(16)), b"k" * 64,
        b"\x00" * 16, b"The quick brown fox jumps over the lazy dog", bytes(range(256))]

# columns: P, <OS, >OS, Total, Alive, Tmb, DR, K Null, K !Null, P-Bytes, K-Bytes, V-Bytes, A K-Sz, A V-Sz, A M-Sz
ROWS = [
    [0, 0, 112298537, 25056009, 25056009, 0, "0.0000", 0, 25056009, 6778805354, 225504081, 6553301273, 9, 261, 270],
    [1, 0, 112244988, 25063295, 25063295, 0, "0.0000", 0, 25063295, 6780199421, 225569655, 6554629766, 9, 261, 270],
    [2, 0, 112295570, 25056714, 25056714, 0, "0.0000", 0, 25056714, 6777635839, 225510426, 6552125413, 9, 261, 270],
    [3, 0, 112275362, 25058243, 25058243, 0, "0.0000", 0, 25058243, 6778031556, 225524187, 6552507369, 9, 261, 270],
    [4, 0, 112315450, 25062939, 25062939, 0, "0.0000", 0, 25062939, 6780416185, 225566451, 6554849734, 9, 261, 270],
    [5, 0, 112267563, 25063360, 25063360, 0, "0.0000", 0, 25063360, 6779370776, 225570240, 6553800536, 9, 261, 270],
    [6, 0, 112262485, 25043793, 25043793, 0, "0.0000", 0, 25043793, 6774475467, 225394137, 6549081330, 9, 261, 270],
    [7, 0, 112147975, 25038860, 25038860, 0, "0.0000", 0, 25038860, 6772769509, 225349740, 6547419769, 9, 261, 270],
    [8, 0, 112332976, 20021871, 20021871, 0, "0.0000", 0, 20021871, 5432377054, 180196839, 5252180215, 9, 262, 271],
    [9, 0, 112279184, 25067204, 25067204, 0, "0.0000", 0, 25067204, 6780916052, 225604836, 6555311216, 9, 261, 270],
]


def main():
    kat = [{"key_hex": k.hex(), "reference_fnv32": ref_fnv(k), "standard_fnv1a32": std_fnv1a(k)} for k in KEYS]
    with open(os.path.join(HERE, "fnv_kat.json"), "w") as f:
        json.dump({"source": "/root/reference/src/fnv32.rs:79-101 followed by hand (make_golden.py)", "vectors": kat},
                  f, indent=1)

    cols = ["P", "start_offset", "end_offset", "total", "alive", "tombstones", "dirty_ratio", "key_null",
            "key_non_null", "p_bytes", "k_bytes", "v_bytes", "key_size_avg", "value_size_avg", "message_size_avg"]
    rows = [dict(zip(cols, r)) for r in ROWS]
    # internal consistency of the transcription (any OCR slip shows up here)
    assert sum(r["total"] for r in rows) == 245532288
    assert sum(r["p_bytes"] for r in rows) == 66434997213
    assert 245532288 // 416 == 590221
    for r in rows:
        assert r["k_bytes"] + r["v_bytes"] == r["p_bytes"], r
        assert r["k_bytes"] // r["alive"] == r["key_size_avg"], r
        assert r["v_bytes"] // r["alive"] == r["value_size_avg"], r
        assert r["p_bytes"] // r["alive"] == r["message_size_avg"], r
    demo = {
        "source": "/root/reference/demo_output.png (README.md:27-28), printed by src/main.rs:123-179",
        "topic": "global.trv_bulk.partner_import",
        "scanning_took_s": 416,
        "estimated_msg_s": 590221,
        "earliest_message": "2018-01-31 17:23:13 UTC",
        "earliest_message_s": calendar.timegm((2018, 1, 31, 17, 23, 13)),
        "latest_message": "2018-04-13 14:29:52 UTC",
        "latest_message_s": calendar.timegm((2018, 4, 13, 14, 29, 52)),
        "largest_message": 750,
        "smallest_message": 139,
        "topic_size": 66434997213,
        "rows": rows,
    }
    with open(os.path.join(HERE, "demo_output.json"), "w") as f:
        json.dump(demo, f, indent=1)
    print("wrote fnv_kat.json (%d vectors), demo_output.json (%d rows)" % (len(kat), len(rows)))


if __name__ == "__main__":
    main()
