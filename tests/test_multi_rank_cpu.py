"""world_size-2 test of the N>1 host path on CPU (gloo): byte-range sharding snapped to newlines + counter all-reduce.
The per-shard scan is stood in for by the oracle (no GPU here); on GPUs bench.py runs the same plumbing over NCCL."""
import os
import subprocess
import sys
import textwrap

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_tile_the_buffer(small_dbs):
    from matchy_b200 import multi
    _, log = small_dbs[2]
    log = log[:200001] + b"last line without newline"
    read_at, n = multi.buffer_reader(log)
    for world in (1, 2, 3, 4, 8, 64):
        rs = [multi.shard_range(read_at, n, r, world) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        for (b0, e0), (b1, e1) in zip(rs, rs[1:]):
            assert e0 == b1
        for b, e in rs:
            assert b == 0 or b == n or log[b - 1:b] == b"\n"


def test_sharded_scan_equals_whole_scan(small_dbs):
    from matchy_b200 import multi
    db, log = small_dbs[5]
    log = log[:400000]
    orc = O.Oracle(db)
    whole_recs, whole_cnt = orc.scan(log, chunk_size=128 * 1024)
    read_at, n = multi.buffer_reader(log)
    for world in (2, 3, 8):
        recs, cnt = [], [0] * 16
        for r in range(world):
            b, e = multi.shard_range(read_at, n, r, world)
            rr, cc = orc.scan(log[b:e], base=b, chunk_size=128 * 1024)
            recs += rr
            cnt = [x + y for x, y in zip(cnt, cc)]
        assert sorted(recs) == whole_recs and cnt == whole_cnt


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import torch.distributed as dist
    import oracle_lib as O
    from matchy_b200 import multi, synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    db = synth.build_db(5, 0.01)
    log = synth.gen_log(5, 6 * 65536, 0.01).tobytes()
    orc = O.Oracle(db)
    read_at, n = multi.buffer_reader(log)
    recs, total = multi.scan_sharded(lambda data, base: orc.scan(data, base=base, chunk_size=128 * 1024), read_at, n, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, recs)
    if rank == 0:
        whole_recs, whole_cnt = orc.scan(log, chunk_size=128 * 1024)
        allr = sorted(r for part in gathered for r in part)
        assert allr == whole_recs, (len(allr), len(whole_recs))
        assert total == whole_cnt, (total, whole_cnt)
        print("OK", world, len(allr), total[:4])
    dist.destroy_process_group()
""")


def test_two_ranks_gloo(tmp_path, built):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "OK 2" in out.stdout
