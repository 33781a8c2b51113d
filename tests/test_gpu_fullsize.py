"""Full-size parity gate: the CUDA path against the CPU oracle with the databases at scale 1.0 (the
BASELINE.json indicator counts), 2040 MiB pieces and multi-GiB HBM-resident input — the branches the 1 %-scale suite never
takes (5 M hashes outside the hot filter, 20-bit IPv4 jump table, ~30 GB of scratch, batches of back-to-back pieces).

Counters AND records are compared over the whole input: the oracle side runs on every host core
(Oracle.scan_mt_keep) and both sides come back as arrays in the layout of mgpu_results().
MATCHY_FULLSIZE_GIB (default 4) sets the input size per config.
"""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

GIB = float(os.environ.get("MATCHY_FULLSIZE_GIB", "4"))


def _compare(eng, recs, ids, want_recs, want_ids, want_cnt, what):
    assert eng.counters_list() == want_cnt, what
    assert len(recs) == len(want_recs), what
    assert recs.tobytes() == want_recs.tobytes(), what
    assert ids.tobytes() == want_ids.tobytes(), what


@pytest.mark.parametrize("cfg", [1, 2, 3, 4, 5])
def test_full_scale_parity(built, cfg):
    from matchy_b200 import Engine, synth
    nbytes = (1 << 30) if cfg == 1 else int(GIB * (1 << 30)) // 65536 * 65536  # config 1 is 1 GiB by definition
    db = synth.build_db(cfg, 1.0)
    log = synth.gen_log(cfg, nbytes, 1.0)
    orc = O.Oracle(db)
    want_recs, want_ids, want_cnt = orc.scan_mt_keep(log)
    assert len(want_recs) > 1000
    eng = Engine(0, chunk_bytes=2040 << 20)
    eng.upload(db)
    dev = eng.dev_alloc(nbytes)
    try:
        eng.dev_upload(dev, log)
        recs, ids = eng.scan_device(dev, nbytes)
        _compare(eng, recs, ids, want_recs, want_ids, want_cnt, "resident, first scan")
        recs, ids = eng.scan_device(dev, nbytes)  # same buffers, same answer
        _compare(eng, recs, ids, want_recs, want_ids, want_cnt, "resident, second scan")
        if cfg in (3, 5):  # match-heavy: every piece is sorted on the device (>= 128 K records per piece), the host only checks the order
            eng.debug_counters()
            assert eng.arrived_sorted
    finally:
        eng.dev_free(dev)
    if cfg in (2, 3):  # the host-buffer entry point (double-buffered H2D, memrchr pieces) at the same size
        recs, ids = eng.scan(log)
        _compare(eng, recs, ids, want_recs, want_ids, want_cnt, "host buffer")
    eng.close()


@pytest.mark.parametrize("cfg,key,value", [(3, "cap_rec", 3000), (3, "cap_ip", 40000), (2, "cap_ids", 700), (2, "cap_str", 500), (5, "cap_rec", 2500)])
def test_overflow_redo_is_exact(built, cfg, key, value):
    """Work buffers far too small for a piece: the piece is split at newlines and redone (scan_piece), results unchanged."""
    from matchy_b200 import Engine, synth
    db = synth.build_db(cfg, 0.05)
    log = synth.gen_log(cfg, 96 << 20, 0.05)
    orc = O.Oracle(db)
    want_recs, want_ids, want_cnt = orc.scan_mt_keep(log)
    eng = Engine(0, chunk_bytes=32 << 20)
    eng.upload(db)
    eng.set_option(key, value)
    dev = eng.dev_alloc(len(log))
    try:
        eng.dev_upload(dev, log)
        recs, ids = eng.scan_device(dev, len(log))
        assert eng.timing()["chunks"] > 3 + 8  # more launches than pieces: something was redone
        _compare(eng, recs, ids, want_recs, want_ids, want_cnt, "resident")
    finally:
        eng.dev_free(dev)
    recs, ids = eng.scan(log)
    _compare(eng, recs, ids, want_recs, want_ids, want_cnt, "host buffer")
    eng.close()


def test_token_list_audit(built):
    """The IP token list of every piece, audited on the device: no slot left unwritten, tokens == per-type counters,
    recomputed lookups == IP records; with the shipped reservation unit and with a larger one."""
    from matchy_b200 import Engine, synth
    db = synth.build_db(3, 0.1)
    log = synth.gen_log(3, 512 << 20, 0.1)
    orc = O.Oracle(db)
    want_cnt = orc.scan_mt(log)
    eng = Engine(0, chunk_bytes=1024 << 20)  # (lists sized for 1 GiB pieces: the large reservation units must not overflow them — the redo path has its own test)
    eng.upload(db)
    dev = eng.dev_alloc(len(log))
    try:
        eng.dev_upload(dev, log)
        for unit in (128, 512, 1024):  # (larger units overflow the list on purpose-built inputs only: the redo path has its own test)
            eng.set_option("tok_reserve", unit)
            eng.set_option("verify_tokens", 1)
            recs, _ = eng.scan_device(dev, len(log))
            d = eng.debug_counters()
            cnt = eng.counters_list()
            assert cnt == want_cnt, unit
            assert d["poisoned"] == 0 and d["unknown"] == 0, (unit, d)
            assert d["ipv4"] == cnt[4 + 2] and d["ipv6"] == cnt[4 + 3], (unit, d)
            assert d["lookup_hits"] == int(np.count_nonzero(recs["kind"] == 1)) == cnt[3], (unit, d)
            eng.set_option("verify_tokens", 0)
    finally:
        eng.dev_free(dev)
    eng.close()


# ---- the single-pass scan_kernel (MATCHY_B200_FUSED=1 / Engine(fused=True)): same gates as the default kernel pair ----
@pytest.mark.parametrize("cfg", [2, 3, 4, 5])
def test_fused_scan_kernel_full_scale_parity(built, cfg):
    from matchy_b200 import Engine, synth
    nbytes = int(min(GIB, 2.0) * (1 << 30)) // 65536 * 65536
    db = synth.build_db(cfg, 1.0)
    log = synth.gen_log(cfg, nbytes, 1.0)
    want_recs, want_ids, want_cnt = O.Oracle(db).scan_mt_keep(log)
    eng = Engine(0, chunk_bytes=512 << 20, fused=True)
    eng.upload(db)
    dev = eng.dev_alloc(nbytes)
    try:
        eng.dev_upload(dev, log)
        recs, ids = eng.scan_device(dev, nbytes)
        _compare(eng, recs, ids, want_recs, want_ids, want_cnt, "fused, resident")
    finally:
        eng.dev_free(dev)
    eng.close()


def test_fused_scan_kernel_small_inputs_and_fuzz(built, small_dbs):
    """scan_kernel on the 1 %-scale configs, on adversarial token soup (extraction and scan), and with pieces so small that
    most warps own a single tile."""
    import random
    from matchy_b200 import Engine
    from test_gpu_parity import _fuzz_text, FRAGS
    for cfg, (db, log) in small_dbs.items():
        orc = O.Oracle(db)
        want, wcnt = orc.scan(log, chunk_size=128 * 1024)
        for chunk in (64 << 10, 8 << 20):
            eng = Engine(0, chunk_bytes=chunk, fused=True)
            eng.upload(db)
            eng.scan(log)
            assert eng.counters_list() == wcnt, (cfg, chunk)
            assert eng.records_as_tuples() == want, (cfg, chunk)
            if chunk == 8 << 20 and cfg == 1:
                rng = random.Random(4321)
                for it in range(150):
                    data = _fuzz_text(rng, rng.randint(0, 60) if it % 6 else rng.randint(300, 3000))
                    flags = rng.choice([31, 31, 31, 1, 2, 4, 8, 16, 5, 10, 21, 0xFF])
                    assert sorted(eng.extract(data, flags)) == sorted(orc.extract(data, flags)), (it, flags)
                for data in (b"", b"\n", b"x", b"1.2.3.4", b"evil.com", b"a." * 5000 + b"com\n", b"9" * 5000, b"1.2.3.4 " * 3000, b"ab.cd " * 4000):
                    eng.scan(data)
                    w, c = orc.scan(data)
                    assert eng.counters_list() == c and eng.records_as_tuples() == w, data[:20]
            eng.close()
