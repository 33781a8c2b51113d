#!/usr/bin/env python
"""Regenerate tests/golden/*.  TEST INFRASTRUCTURE.

The reference (Rust) cannot be built or run in this environment, so there is no reference-generated output to freeze.
What is frozen here instead:

* `extractor_kats.json` — the reference's OWN known-answer vectors for the extractor (inputs and expected items copied from
  the assertions of crates/matchy-extractor/src/lib.rs:1922-3626; each entry cites its test), in one machine-readable file
  that both the oracle test and the GPU test read;
* `cfgN.ndjson` / `cfgN.counters.json` — for every BASELINE config, the sorted `matchy match` NDJSON and the WorkerStats counters
  of a 1 MiB slice of the deterministic synthetic stream over the 1 %-scale database, as computed by the oracle
  (oracle/oracle.cpp) when this file was generated.  They pin the oracle, the generators, the .mxy writer and the device
  path against drift: any of them changing shows up as a diff of committed text.

usage: python tests/golden/make_golden.py   (rewrites the files next to it)"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SLICE = 1024 * 1024
SCALE = 0.01

# (test name in lib.rs, input, extractor flags, expected [(type name, text)] in any order)
EXTRACTOR_KATS = [
    ("test_bitcoin_legacy_extraction :3240", "Send to 1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa for payment", 0xE0, [["Bitcoin", "1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa"]]),
    ("test_bitcoin_p2sh_extraction :3260", "Payment to 3Cbq7aT1tY8kMxWLbitaG7yT6bPbKChq64 confirmed", 0xE0, [["Bitcoin", "3Cbq7aT1tY8kMxWLbitaG7yT6bPbKChq64"]]),
    ("test_bitcoin_bech32_extraction :3280", "Withdraw to bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq", 0xE0, [["Bitcoin", "bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq"]]),
    ("test_bitcoin_reject_invalid_checksum :3300", "Fake address 1A1zP1eP5QGefi2DMPTfTL5SLmv7Divf00 is invalid", 0xE0, []),
    ("test_bitcoin_reject_too_short :3324", "Short address 1A1zP1eP is invalid", 0xE0, []),
    ("test_ethereum_extraction_lowercase :3343", "Send to 0x5aeda56215b167893e80b4fe645ba6d5bab767de", 0xE0, [["Ethereum", "0x5aeda56215b167893e80b4fe645ba6d5bab767de"]]),
    ("test_ethereum_extraction_checksummed :3363", "Send to 0x5aAeb6053F3E94C9b9A09f33669435E7Ef1BeAed", 0xE0, [["Ethereum", "0x5aAeb6053F3E94C9b9A09f33669435E7Ef1BeAed"]]),
    ("test_ethereum_reject_invalid_checksum :3383", "Bad address 0x5aAeb6053f3e94c9b9a09f33669435e7ef1beaed", 0xE0, []),
    ("test_ethereum_reject_wrong_length :3407", "Short address 0x5aeda56215b167893e80b4fe645ba6d5bab7", 0xE0, []),
    ("test_ethereum_reject_non_hex :3426", "Invalid 0x5aeda56215b167893e80b4fe645ba6d5bab767dg", 0xE0, []),
    ("test_monero_reject_wrong_prefix :3478", "Fake 1AdUndXHHZ6cfufTMvppY6JwXNouMBzSkbLYfpAV5Usx3skxNgYeYTRj5UzqtReoS44qo9mtmXCqY45DJ852K5Jv2684Rge", 0x80, []),
    ("test_monero_reject_too_short :3501", "Short 4AdUndXHHZ6cfufTMvppY6JwXNouMBzSkbLYfpAV5Usx", 0x80, []),
    ("test_crypto_mixed_with_other_types :3520", "Transaction from 192.168.1.1 to bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq via example.com", 0xFF,
     [["IPv4", "192.168.1.1"], ["Bitcoin", "bc1qar0srrr7xfkvy5l643lydnw9re59gtzzwf5mdq"], ["Domain", "example.com"]]),
    ("test_ethereum_in_log_line :3588", "2025-01-15 10:32:45 Transaction to=0x5aeda56215b167893e80b4fe645ba6d5bab767de value=1000000000000000000", 0xE0,
     [["Ethereum", "0x5aeda56215b167893e80b4fe645ba6d5bab767de"]]),
    ("test_bitcoin_chunk_extraction :3608", "Line1: 1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa\nLine2: 3Cbq7aT1tY8kMxWLbitaG7yT6bPbKChq64\n", 0xE0,
     [["Bitcoin", "1A1zP1eP5QGefi2DMPTfTL5SLmv7DivfNa"], ["Bitcoin", "3Cbq7aT1tY8kMxWLbitaG7yT6bPbKChq64"]]),
    ("doc example :405", "test@example.com\n192.168.1.1\nmalware.com", 0x1F, [["Email", "test@example.com"], ["Domain", "example.com"], ["IPv4", "192.168.1.1"], ["Domain", "malware.com"]]),
    ("worker smoke processing/mod.rs:614", "Connection from 1.2.3.4\n", 0x1F, [["IPv4", "1.2.3.4"]]),
]


def main():
    import __graft_entry__ as g
    g.build()
    import oracle_lib as O
    from matchy_b200 import synth
    with open(os.path.join(HERE, "extractor_kats.json"), "w") as f:
        json.dump([{"test": t, "input": i, "flags": fl, "expect": e} for t, i, fl, e in EXTRACTOR_KATS], f, indent=1)
    for cfg in (1, 2, 3, 4, 5):
        db = synth.build_db(cfg, SCALE)
        log = synth.gen_log(cfg, SLICE, SCALE).tobytes()
        o = O.Oracle(db)
        recs, cnt = o.scan(log, chunk_size=128 * 1024)
        lines = sorted(o.ndjson(log, "golden.log").splitlines())
        with open(os.path.join(HERE, "cfg%d.ndjson" % cfg), "wb") as f:
            f.write(b"\n".join(lines) + (b"\n" if lines else b""))
        with open(os.path.join(HERE, "cfg%d.counters.json" % cfg), "w") as f:
            json.dump({"slice_bytes": SLICE, "db_scale": SCALE, "db_bytes": len(db), "records": len(recs),
                       "counters": dict(zip(["lines", "bytes", "candidates", "matches"] + ["type%d" % k for k in range(12)], cnt))}, f, indent=1)
        print("cfg", cfg, len(lines), "lines of NDJSON")


if __name__ == "__main__":
    main()
