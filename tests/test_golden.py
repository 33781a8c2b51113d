"""Committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py): the reference's own extractor vectors in
machine-readable form, and frozen `matchy match` output (sorted NDJSON + counters) of a 1 MiB slice of every BASELINE
config.  The CPU test pins the oracle to them, the GPU test pins the CUDA path to them."""
import json
import os

import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLD, "extractor_kats.json")))


def _slice(cfg):
    from matchy_b200 import synth
    meta = json.load(open(os.path.join(GOLD, "cfg%d.counters.json" % cfg)))
    db = synth.build_db(cfg, meta["db_scale"])
    assert len(db) == meta["db_bytes"], "the .mxy writer or the database generator changed"
    log = synth.gen_log(cfg, meta["slice_bytes"], meta["db_scale"]).tobytes()
    want_lines = open(os.path.join(GOLD, "cfg%d.ndjson" % cfg), "rb").read().splitlines()
    want_cnt = [meta["counters"][k] for k in ["lines", "bytes", "candidates", "matches"] + ["type%d" % k for k in range(12)]]
    return db, log, want_lines, want_cnt


@pytest.mark.parametrize("kat", KATS, ids=[k["test"] for k in KATS])
def test_oracle_extractor_golden(built, small_dbs, kat):
    o = O.Oracle(small_dbs[1][0])
    got = sorted([t, s.decode()] for t, s in o.extract_strings(kat["input"].encode(), kat["flags"]))
    assert got == sorted(kat["expect"])


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_oracle_scan_golden(built, cfg):
    db, log, want_lines, want_cnt = _slice(cfg)
    o = O.Oracle(db)
    _, cnt = o.scan(log, chunk_size=128 * 1024)
    assert cnt == want_cnt
    assert sorted(o.ndjson(log, "golden.log").splitlines()) == want_lines


@pytest.mark.gpu
@pytest.mark.parametrize("kat", KATS, ids=[k["test"] for k in KATS])
def test_device_extractor_golden(built, kat):
    from matchy_b200 import Engine, ITEM_TYPE_NAMES
    eng = test_device_extractor_golden.__dict__.setdefault("eng", Engine(0, chunk_bytes=1 << 20))
    data = kat["input"].encode()
    got = sorted([ITEM_TYPE_NAMES[t], data[s:e].decode()] for t, s, e in eng.extract(data, kat["flags"]))
    assert got == sorted(kat["expect"])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_device_scan_golden(built, cfg):
    from matchy_b200 import Engine, RecordFormatter
    db, log, want_lines, want_cnt = _slice(cfg)
    eng = Engine(0, chunk_bytes=1 << 20)
    eng.upload(db)
    recs, ids = eng.scan(log)
    assert eng.counters_list() == want_cnt
    assert sorted(RecordFormatter(db).ndjson(recs, ids, log, 0, "golden.log").splitlines()) == want_lines
    eng.close()
