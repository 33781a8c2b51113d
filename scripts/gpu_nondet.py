"""Round-1 open item: config 3 at full size returned 8 065 272 … 8 065 275 matches from run to run with a 512-slot token
reservation.  Reproduce with the token-list audit on: which stage loses the records?  (Run on a B200 via gpurun.)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402
from matchy_b200 import Engine, synth  # noqa: E402

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
nbytes = int(gb * 1e9) // 65536 * 65536
db = synth.build_db(3, 1.0)
log = synth.gen_log(3, nbytes, 1.0)
t0 = time.time()
want = O.Oracle(db).scan_mt(log)
print("oracle: %d matches, counters %r (%.1f s)" % (want[3], want, time.time() - t0), flush=True)
eng = Engine(0, chunk_bytes=2040 << 20)
eng.upload(db)
dev = eng.dev_alloc(nbytes)
eng.dev_upload(dev, log)
variants = [int(x) for x in os.environ.get("VARIANTS", "0").split(",")]
for unit, variant in [(u, v) for v in variants for u in (128, 512, 2048)]:
    eng.set_option("tok_reserve", unit)
    eng.set_option("variant", variant)
    print("---- unit %d variant %d" % (unit, variant), flush=True)
    for verify in (1,):
        eng.set_option("verify_tokens", verify)
        seen = []
        for r in range(runs):
            recs, _ = eng.scan_device(dev, nbytes)
            cnt = eng.counters_list()
            d = eng.debug_counters() if verify else {}
            seen.append(cnt[3])
            ok = cnt == want
            print("unit %4d verify %d run %d: matches %d ip-records %d counters_equal %s %s" % (unit, verify, r, cnt[3], int(np.count_nonzero(recs["kind"] == 1)), ok, d), flush=True)
            if verify:
                for ev in getattr(eng, "debug_events", []):
                    print("      partial warp: block %d base %d (slots %d) warp %d active %08x" % (ev[0], ev[1], d["slots"], ev[2], ev[3]))
                eng.set_option("verify_tokens", 1)  # reset the sums
        print("unit %4d verify %d: distinct match counts %r, oracle %d" % (unit, verify, sorted(set(seen)), want[3]), flush=True)
eng.dev_free(dev)
