"""Multi-GPU plumbing: byte-range sharding of a log across ranks (one process per GPU, database replicated) and the
single collective of the path — a sum all-reduce of the summary counters (NCCL over NVLink on GPUs; gloo in CPU tests).

Extraction is line-local, so any newline-aligned partition gives the same match set.  A shard owns the
lines that START inside its nominal byte range: boundaries are snapped forward to the byte after the next '\\n', the
same rule FileReader::next_batch applies to chunk ends (crates/matchy/src/processing/mod.rs:231-245)."""
import os


def nominal_bounds(total_len, world):
    return [total_len * r // world for r in range(world + 1)]


def snap_forward(read_at, pos, total_len, probe=1 << 16):
    """First offset >= pos that starts a line (pos itself if pos == 0 or byte pos-1 is '\\n').
    read_at(offset, n) -> bytes.  Scans forward in `probe`-byte reads."""
    if pos <= 0:
        return 0
    if pos >= total_len:
        return total_len
    if read_at(pos - 1, 1) == b"\n":
        return pos
    at = pos
    while at < total_len:
        buf = read_at(at, min(probe, total_len - at))
        k = buf.find(b"\n")
        if k >= 0:
            return at + k + 1
        at += len(buf)
    return total_len


def shard_range(read_at, total_len, rank, world):
    """[begin, end) of rank's shard; shards tile [0, total_len) exactly and every shard starts at a line start."""
    b = nominal_bounds(total_len, world)
    return snap_forward(read_at, b[rank], total_len), snap_forward(read_at, b[rank + 1], total_len)


def file_reader(path):
    f = open(path, "rb")

    def read_at(off, n):
        f.seek(off)
        return f.read(n)
    return read_at, os.path.getsize(path)


def buffer_reader(buf):
    mv = memoryview(buf)

    def read_at(off, n):
        return bytes(mv[off:off + n])
    return read_at, len(mv)


def allreduce_counters(vec, device=None):
    """Sum a list of non-negative integers over all ranks.  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(vec)
    t = torch.tensor(list(vec), dtype=torch.int64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


def scan_sharded(scan_fn, read_at, total_len, rank, world):
    """Scan this rank's shard with scan_fn(bytes, base) -> (records, counters_list); returns
    (local records, globally summed counters).  Records carry absolute offsets, so ranks' record lists simply concatenate."""
    b, e = shard_range(read_at, total_len, rank, world)
    data = read_at(b, e - b) if e > b else b""
    recs, cnt = scan_fn(data, b)
    return recs, allreduce_counters(cnt)
