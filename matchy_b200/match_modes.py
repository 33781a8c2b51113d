"""The slower modes of `matchy match` around the same device scan:

* sequential mode (`--threads 1`, bin/match_processor/sequential.rs:205-390): matches in line order — inside a line in
  extract_from_line's order —, real timestamps, canonical address text;
* follow mode (`--follow`, bin/match_processor/follow.rs:22-272): existing content first, then whatever is appended to the
  files, batch by batch through the double-buffered scan; truncation restarts a file at offset 0;
* hot reload (`WatchingDatabase`, crates/matchy/src/watching_database.rs:103-191): the database file is watched and, when it
  changes, uploaded again into the same engine; a generation counter tells callers that it happened.

The hot path is untouched: every batch is one `mgpu_scan`."""
import os
import sys
import threading
import time

import numpy as np

from . import engine as E
from .database import Database, DatabaseError

# extract_from_line emits domains, IPv4, e-mails, IPv6, hashes, Bitcoin, Ethereum, Monero (matchy-extractor/src/lib.rs:1472-1521)
_GROUP_OF_TYPE = np.array([0, 2, 1, 3, 4, 4, 4, 4, 4, 5, 6, 7], dtype=np.int64)


def sequential_order(recs, data, base=0):
    """Indices that put `recs` (any order) into sequential-mode order: by line, then by extractor group, then by position."""
    if len(recs) == 0:
        return np.zeros(0, dtype=np.int64)
    data = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else data, dtype=np.uint8)
    starts = np.concatenate(([0], np.flatnonzero(data == 10) + 1)).astype(np.int64)
    off = recs["offset"].astype(np.int64) - int(base)
    line = np.searchsorted(starts, off, side="right") - 1
    return np.lexsort((off, _GROUP_OF_TYPE[recs["item_type"].astype(np.int64)], line))


def timestamp_text(now=None):
    """`format!("{:.3}", timestamp)` of the wall clock (sequential.rs:47-64)."""
    return "%.3f" % (time.time() if now is None else now)


def render_sequential(fmt, recs, ids, data, source, base=0, now=None):
    """NDJSON of one batch as `matchy match --threads 1` prints it."""
    if len(recs) == 0:
        return b""
    order = sequential_order(recs, data, base)
    return fmt.ndjson(recs[order], ids, data, base, source, timestamp=timestamp_text(now))


class WatchingDatabase:
    """`WatchingDatabase::from(path).open()` (watching_database.rs:103-191): a Database that reloads itself when its file
    changes.  Here "reload" = the new file's sections uploaded into the SAME engine (HBM buffers of the old database are
    freed by the upload), so scans in flight finish on the old tables and the next scan sees the new ones.

        db = WatchingDatabase.from_("threats.mxy").on_reload(lambda gen, path: ...).open()
        db.engine.scan(...)          # always the latest database
        db.generation()              # incremented on every successful reload

    The file is polled (mtime, size, inode) every `interval` seconds by a daemon thread; `check()` does one poll by hand."""

    class Opener:
        def __init__(self, path):
            self._path, self._device, self._chunk, self._interval, self._cb, self._thread = os.fspath(path), 0, 0, 1.0, None, True

        def device(self, d): self._device = int(d); return self
        def chunk_bytes(self, n): self._chunk = int(n); return self
        def poll_interval(self, seconds): self._interval = float(seconds); return self
        def on_reload(self, callback): self._cb = callback; return self
        def no_thread(self): self._thread = False; return self
        def cache_capacity(self, _n): return self
        def no_cache(self): return self

        def open(self):
            return WatchingDatabase(self._path, self._device, self._chunk, self._interval, self._cb, self._thread)

    @staticmethod
    def from_(path):
        return WatchingDatabase.Opener(path)

    def __init__(self, path, device=0, chunk_bytes=0, interval=1.0, callback=None, thread=True):
        self.path = path
        self._db = Database.from_(path).device(device).chunk_bytes(chunk_bytes).open()
        self._sig = self._signature()
        self._gen = 1  # the reference starts at generation 1 after the first load
        self._cb = callback
        self._lock = threading.Lock()
        self._stop = threading.Event()
        self._thread = None
        if thread:
            self._thread = threading.Thread(target=self._run, args=(interval,), daemon=True)
            self._thread.start()

    def _signature(self):
        st = os.stat(self.path)
        return (st.st_mtime_ns, st.st_size, st.st_ino)

    def _run(self, interval):
        while not self._stop.wait(interval):
            try:
                self.check()
            except Exception as e:  # a half-written file: keep the old database, try again on the next tick
                print("[WARN] database reload failed: %s" % e, file=sys.stderr)

    def check(self):
        """One poll: reload when the file changed.  Returns True when a new database is live."""
        try:
            sig = self._signature()
        except OSError:
            return False
        if sig == self._sig:
            return False
        with open(self.path, "rb") as f:
            mxy = f.read()
        with self._lock:
            try:
                self._db.engine.upload(mxy)  # validates first: a bad file leaves the old tables in place
            except E.EngineError as e:
                self._sig = sig  # do not retry the same bad bytes on every tick
                raise DatabaseError(str(e))
            self._db._bytes = mxy
            self._db._fmt = E.RecordFormatter(mxy)
            self._db._info = self._db.engine.db_info()
            self._sig = sig
            self._gen += 1
            gen = self._gen
        if self._cb:
            self._cb(gen, self.path)
        return True

    def generation(self):
        return self._gen

    def snapshot(self):
        return self._db

    @property
    def engine(self):
        return self._db.engine

    def lock(self):
        """Held around a scan + its rendering, so that records and the tables that decode them belong together."""
        return self._lock

    def formatter(self):
        return self._db._fmt

    def lookup(self, query):
        with self._lock:
            return self._db.lookup(query)

    def close(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=5)
        self._db.close()


def follow_files(paths, scan_batch, out, poll=0.2, idle_exit=None, stop=None, warn=None):
    """`matchy match --follow`: process what the files hold, then every batch of appended bytes, until `stop` is set (or nothing
    has arrived for `idle_exit` seconds — tests).  scan_batch(data: np.ndarray, source: str) -> bytes renders one batch.
    A file that shrank was rotated: it starts again at offset 0 (follow.rs:223-226).  Returns total bytes processed."""
    if any(p == "-" for p in paths):
        raise ValueError("--follow mode not supported with stdin")
    warn = warn or sys.stderr
    pos = {p: 0 for p in paths}
    total = 0
    last_activity = time.monotonic()
    while True:
        progressed = False
        for p in paths:
            try:
                size = os.stat(p).st_size
            except OSError:
                if pos[p] != -1:
                    print("[WARN] File deleted/rotated: %s" % p, file=warn)
                    pos[p] = -1
                continue
            if pos[p] == -1 or size < pos[p]:
                pos[p] = 0
            if size == pos[p]:
                continue
            with open(p, "rb") as f:
                f.seek(pos[p])
                data = np.frombuffer(f.read(size - pos[p]), dtype=np.uint8)
            pos[p] += data.size
            total += data.size
            out.write(scan_batch(data, p))
            out.flush()
            progressed = True
        now = time.monotonic()
        if progressed:
            last_activity = now
        elif idle_exit is not None and now - last_activity >= idle_exit:
            return total
        if stop is not None and stop.is_set():
            return total
        if not progressed:
            time.sleep(poll)
