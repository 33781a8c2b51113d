"""Mirror of the reference's batch-processing layer (crates/matchy/src/processing/mod.rs): `Worker`,
`WorkerStats`, `MatchResult`, `FileReader`.  `Worker.process_bytes` is the hot path: one call == one
mgpu_scan (tokenize → validate → trie / literal-hash / AC+glob → records) on the database's GPU."""
import os

from . import engine as E
from .database import QueryResult


class WorkerStats:
    """processing/mod.rs:86-128 (the sampled timers are replaced by per-kernel device times)."""
    FIELDS = ["lines_processed", "candidates_tested", "matches_found", "total_bytes", "ipv4_count", "ipv6_count", "domain_count",
              "email_count", "md5_count", "sha1_count", "sha256_count", "sha384_count", "sha512_count", "bitcoin_count",
              "ethereum_count", "monero_count"]

    def __init__(self):
        for f in self.FIELDS:
            setattr(self, f, 0)
        self.kernel_ms = {k: 0.0 for k in E.KERNEL_NAMES}

    def add_counters(self, c, timing=None):
        self.lines_processed += c["lines"]; self.total_bytes += c["bytes"]
        self.candidates_tested += c["candidates"]; self.matches_found += c["matches"]
        t = c["by_type"]
        self.domain_count += t[0]; self.email_count += t[1]; self.ipv4_count += t[2]; self.ipv6_count += t[3]
        self.md5_count += t[4]; self.sha1_count += t[5]; self.sha256_count += t[6]; self.sha384_count += t[7]; self.sha512_count += t[8]
        self.bitcoin_count += t[9]; self.ethereum_count += t[10]; self.monero_count += t[11]
        if timing:
            for k, v in timing["kernel_ms"].items():
                self.kernel_ms[k] += v

    def as_vector(self):
        """[lines, bytes, candidates, matches] — what the multi-GPU path all-reduces."""
        return [getattr(self, f) for f in self.FIELDS]

    def __repr__(self):
        return "WorkerStats(" + ", ".join("%s=%d" % (f, getattr(self, f)) for f in self.FIELDS) + ")"


class MatchResult:
    """processing/mod.rs:131-145."""
    __slots__ = ("matched_text", "match_type", "result", "database_id", "source", "byte_offset")

    def __init__(self, matched_text, match_type, result, database_id, source, byte_offset):
        self.matched_text, self.match_type, self.result = matched_text, match_type, result
        self.database_id, self.source, self.byte_offset = database_id, source, byte_offset

    def __repr__(self):
        return "MatchResult(%r, %s, %r, db=%r, source=%r, offset=%d)" % (self.matched_text, self.match_type, self.result, self.database_id, self.source, self.byte_offset)


class DataBatch:
    def __init__(self, source, data):
        self.source, self.data = source, data


class FileReader:
    """`FileReader::new(path, chunk_size)`; `next_batch` cuts after the last newline of each read and carries the
    tail (processing/mod.rs:152-270).  Plain files and '-' (stdin) only — gzip input is out of scope."""

    def __init__(self, path, chunk_size=128 * 1024):
        self.path = os.fspath(path)
        self.chunk_size = int(chunk_size)
        self._f = os.fdopen(os.dup(0), "rb") if self.path == "-" else open(self.path, "rb")
        self._leftover = b""
        self._eof = False

    def next_batch(self):
        if self._eof:
            return None
        while True:
            buf = self._f.read(self.chunk_size)
            if not buf:
                self._eof = True
                if self._leftover:
                    out, self._leftover = self._leftover, b""
                    return DataBatch(self.path, out)
                return None
            combined = self._leftover + buf
            cut = combined.rfind(b"\n")
            if cut >= 0:
                self._leftover = combined[cut + 1:]
                return DataBatch(self.path, combined[:cut + 1])
            self._leftover = combined

    def batches(self):
        while True:
            b = self.next_batch()
            if b is None:
                return
            yield b

    def close(self):
        self._f.close()


class WorkerBuilder:
    def __init__(self):
        self._extractor = None
        self._dbs = []

    def extractor(self, extractor):
        self._extractor = extractor
        return self

    def add_database(self, database_id, database):
        self._dbs.append((str(database_id), database))
        return self

    def build(self):
        if self._extractor is None:
            raise ValueError("Extractor not set - call .extractor()")
        if not self._dbs:
            raise ValueError("No databases added - call .add_database() at least once")
        return Worker(self._extractor, self._dbs)


class Worker:
    def __init__(self, extractor, databases):
        self._extractor = extractor
        self._dbs = databases
        self._stats = WorkerStats()
        self.last_raw = None  # (recs, ids) of the last process_bytes on the first database, for NDJSON rendering

    @staticmethod
    def builder():
        return WorkerBuilder()

    def stats(self):
        return self._stats

    def reset_stats(self):
        self._stats = WorkerStats()

    def process_bytes(self, data, decode=True):
        """`Worker::process_bytes` (processing/mod.rs:353-448).  Order of the returned matches: by byte offset
        (the reference's order is by item type and is not part of its contract — SURVEY quirk 1)."""
        flags = self._extractor.device_flags()
        out = []
        for n, (db_id, db) in enumerate(self._dbs):
            recs, ids = db.engine.scan(data, flags)
            c = db.engine.counters()
            if n == 0:
                self._stats.add_counters(c, db.engine.timing())
                self.last_raw = (recs.copy(), ids.copy())  # (views of the engine's buffers: keep copies)
            else:
                self._stats.matches_found += c["matches"]
            mv = memoryview(data) if not isinstance(data, memoryview) else data
            for r in recs:
                off, ln = int(r["offset"]), int(r["len"])
                text = bytes(mv[off:off + ln]).decode("utf-8", errors="replace")  # (tokens are validated UTF-8 or ASCII; never raise on a record)
                if r["kind"] == E.KIND_IP:
                    qr = QueryResult("Ip", data=db.decode(int(r["data_offset"])) if decode else int(r["data_offset"]), prefix_len=int(r["prefix_len"]))
                else:
                    a, k = int(r["ids_index"]), int(r["n_ids"])
                    pairs = ids[a:a + k]
                    qr = QueryResult("Pattern", pattern_ids=[int(p["pattern_id"]) for p in pairs],
                                     data=[(None if int(p["data_offset"]) == E.NO_DATA else (db.decode(int(p["data_offset"])) if decode else int(p["data_offset"]))) for p in pairs])
                out.append(MatchResult(text, E.ITEM_TYPE_NAMES[int(r["item_type"])], qr, db_id, "", off))
        return out

    def process_batch(self, batch):
        res = self.process_bytes(batch.data)
        for m in res:
            m.source = batch.source
        return res
