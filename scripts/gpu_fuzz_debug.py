"""Find and print the first extractor-fuzz case where the device and the oracle disagree (run on a B200 via gpurun)."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
g.build()
import oracle_lib as O
from matchy_b200 import Engine, synth
from test_gpu_parity import _fuzz_text
db = synth.build_db(1, 0.01)
eng = Engine(0, chunk_bytes=8 << 20); eng.upload(db)
orc = O.Oracle(db)
rng = random.Random(1234)
bad = 0
for it in range(400):
    data = _fuzz_text(rng, rng.randint(0, 60) if it % 8 else rng.randint(300, 3000))
    flags = rng.choice([31, 31, 31, 1, 2, 4, 8, 16, 5, 10, 21])
    want = sorted((s, t, e) for t, s, e in orc.extract(data, flags))
    got = sorted((s, t, e) for t, s, e in eng.extract(data, flags))
    if got != want:
        bad += 1
        miss = sorted(set(want) - set(got)); extra = sorted(set(got) - set(want))
        print("case %d flags %d len %d: missing %d extra %d" % (it, flags, len(data), len(miss), len(extra)))
        for s, t, e in (miss[:6] + extra[:3]):
            print("   %s type %d [%d,%d) tile %d off %d len %d: ...%r[%r]%r..." % ("MISS" if (s, t, e) in miss else "EXTRA", t, s, e, e >> 10, e & 1023, e - s, data[max(0, s - 12):s], data[s:min(e, s + 40)], data[e:e + 12]))
        eng.set_option("fused", 0)
        got2 = sorted((s, t, e) for t, s, e in eng.extract(data, flags))
        eng.set_option("fused", 1)
        print("   unfused path agrees with oracle:", got2 == want)
        if bad >= 4:
            break
print("bad cases:", bad)
