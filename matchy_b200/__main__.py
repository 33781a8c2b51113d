"""`python -m matchy_b200 match DB.mxy LOG [LOG...]` — the `matchy match` output contract on the B200 scan path.

Prints one JSON object per match, in the reference's parallel-mode format (bin/match_processor/parallel.rs:297-369: keys
sorted, `timestamp` "0.000", `source` = the file path), and with `--stats` the WorkerStats counters on stderr.  Only what the
scan path needs: no follow mode, no gzip, no `--format`; the CLI proper is out of scope."""
import argparse
import json
import sys
import time

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m matchy_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    m = sub.add_parser("match", help="scan log files against a .mxy database, NDJSON per match on stdout")
    m.add_argument("database")
    m.add_argument("inputs", nargs="+")
    m.add_argument("--device", type=int, default=0)
    m.add_argument("--extractors", default="", help="comma list; '-crypto' drops the Bitcoin/Ethereum/Monero extractors (match_cmd.rs)")
    m.add_argument("--stats", action="store_true")
    m.add_argument("--chunk-mb", type=int, default=512)
    a = ap.parse_args(argv)
    from . import Engine, RecordFormatter
    db = open(a.database, "rb").read()
    eng = Engine(a.device, chunk_bytes=a.chunk_mb << 20)
    eng.upload(db)
    flags = eng.default_flags()
    info = eng.db_info()
    if (info["has_literal"] or info["has_glob"]) and "-crypto" not in a.extractors.split(","):
        flags |= 0xE0  # crypto extractors are on by default when the database has strings (match_cmd.rs:290-292)
    fmt = RecordFormatter(db)
    tot = None
    t0 = time.perf_counter()
    nbytes = 0
    out = sys.stdout.buffer
    for path in a.inputs:
        data = np.fromfile(path, dtype=np.uint8)
        nbytes += data.size
        recs, ids = eng.scan(data, flags)
        out.write(fmt.ndjson(recs, ids, data, 0, path))
        c = eng.counters_list()
        tot = c if tot is None else [x + y for x, y in zip(tot, c)]
    out.flush()
    if a.stats and tot is not None:
        dt = time.perf_counter() - t0
        names = ["lines", "bytes", "candidates", "matches"]
        print(json.dumps({"stats": dict(zip(names, tot[:4])), "seconds": round(dt, 3), "MB_per_s": round(nbytes / dt / 1e6, 1)}), file=sys.stderr)
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
