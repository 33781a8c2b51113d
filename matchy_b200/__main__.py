"""`python -m matchy_b200 match DB.mxy LOG [LOG...]` — the `matchy match` output contract on the B200 scan path.
`python -m matchy_b200 extract LOG [LOG...]` — the `matchy extract` output contract (bin/commands/extract_cmd.rs).

Prints one JSON object per match, in the reference's parallel-mode format (bin/match_processor/parallel.rs:297-369: keys
sorted, `timestamp` "0.000", `source` = the file path), and with `--stats` the WorkerStats counters on stderr.  `--threads 1`
gives the sequential-mode output instead (line order, wall-clock timestamps, canonical address text:
bin/match_processor/sequential.rs:205-390), `--follow` keeps reading what is appended to the files (follow.rs) and
`--watch-db` reloads the database when its file changes (watching_database.rs).  `.gz` inputs are inflated on the host, "-"
is stdin.  The CLI proper (progress bars, `--format summary`) is out of scope."""
import argparse
import json
import sys
import time

import numpy as np


def read_input(path):
    """file_reader::open (crates/matchy/src/file_reader.rs:44-110): "-" is stdin, a name ending in .gz (any case) is inflated —
    on the host, one gzip member like flate2's GzDecoder — everything else is read as it is.  Returns a uint8 array."""
    if path == "-":
        return np.frombuffer(sys.stdin.buffer.read(), dtype=np.uint8)
    if path.lower().endswith(".gz"):
        import zlib
        with open(path, "rb") as f:
            return np.frombuffer(zlib.decompressobj(31).decompress(f.read()), dtype=np.uint8)
    return np.fromfile(path, dtype=np.uint8)


def extract_lines(eng, data, flags):
    """`matchy extract` over one input: yields (type code, start, end) in the order LineScanner + extract_from_line produce
    them (bin/cli_utils.rs:9-95, matchy-extractor/src/lib.rs:1472-1521): line by line, and inside a line domains, IPv4,
    e-mails, IPv6, hashes, Bitcoin, Ethereum, Monero, each by position.  Extraction is line-local ('\n' is a boundary,
    SURVEY quirk 5), so one device pass over the whole buffer finds exactly the per-line items.  One difference has to be
    bridged on the host: LineScanner trims ASCII whitespace from both ends of a line, and form feed (0x0C) is whitespace but not
    a token boundary — form feeds inside a line's leading / trailing whitespace run are therefore blanked before the scan.
    Also returns (lines, bytes) as the reference counts them: non-empty trimmed lines, trimmed lengths."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = data.size
    nl = np.flatnonzero(data == 10)
    starts = np.concatenate(([0], nl + 1)).astype(np.int64)
    ends = np.concatenate((nl, [n])).astype(np.int64)
    ws = (data == 32) | ((data >= 9) & (data <= 13) & (data != 11))  # u8::is_ascii_whitespace: SP \t \n \x0C \r
    ff = np.flatnonzero(data == 12)
    if ff.size:
        data = data.copy()
        # first / last non-whitespace byte of each line that holds a form feed
        line_of = np.searchsorted(starts, ff, side="right") - 1
        for li in np.unique(line_of):
            s, e = int(starts[li]), int(ends[li])
            seg = ~ws[s:e]
            nz = np.flatnonzero(seg)
            lo, hi = (s + int(nz[0]), s + int(nz[-1]) + 1) if nz.size else (e, e)
            blank = ff[(line_of == li) & ((ff < lo) | (ff >= hi))]
            data[blank] = 32
    # trimmed length of every line = number of bytes between its first and last non-whitespace byte
    nonws = (~ws).astype(np.int64)
    cum = np.concatenate(([0], np.cumsum(nonws)))
    has = (cum[ends] - cum[starts]) > 0
    idx = np.flatnonzero(~ws)
    first = idx[np.searchsorted(idx, starts[has], side="left")] if idx.size else np.zeros(0, np.int64)
    last = idx[np.searchsorted(idx, ends[has], side="left") - 1] if idx.size else np.zeros(0, np.int64)
    n_lines, n_bytes = int(has.sum()), int((last - first + 1).sum()) if idx.size else 0
    items = eng.extract_array(data, flags)
    if items.shape[0]:
        order = np.array([0, 2, 1, 3, 4, 4, 4, 4, 4, 5, 6, 7], dtype=np.int64)  # by item type code: extract_from_line's group
        line = np.searchsorted(starts, items[:, 1].astype(np.int64), side="right") - 1
        key = np.lexsort((items[:, 1].astype(np.int64), order[items[:, 0].astype(np.int64)], line))
        items = items[key]
    return items, n_lines, n_bytes, data


def cmd_extract(a):
    fmt = a.format.lower()
    if fmt not in ("json", "csv", "text"):
        print("Error: Invalid format '%s', expected: json, csv, or text" % a.format, file=sys.stderr)
        return 1
    want = {"ipv4": True, "ipv6": True, "domains": True, "emails": True}
    if a.types is not None:
        want = dict.fromkeys(want, False)
        for part in [t.strip() for t in a.types.lower().split(",")]:
            if part in ("ipv4", "ip4"): want["ipv4"] = True
            elif part in ("ipv6", "ip6"): want["ipv6"] = True
            elif part in ("domain", "domains"): want["domains"] = True
            elif part in ("email", "emails"): want["emails"] = True
            elif part == "ip": want["ipv4"] = want["ipv6"] = True
            elif part == "all": want = dict.fromkeys(want, True)
            else:
                print("Error: Unknown extraction type '%s', expected: ipv4, ipv6, ip, domain, email, all" % part, file=sys.stderr)
                return 1
        if not any(want.values()):
            print("Error: At least one extraction type must be enabled", file=sys.stderr)
            return 1
    from . import Engine, Extractor
    from .engine import ITEM_TYPE_NAMES
    # Extractor::builder() leaves hashes and the crypto-address extractors at their defaults (on), as extract_cmd.rs does
    try:
        ex = (Extractor.builder().extract_ipv4(want["ipv4"]).extract_ipv6(want["ipv6"]).extract_domains(want["domains"])
              .extract_emails(want["emails"]).min_domain_labels(a.min_labels).require_word_boundaries(not a.no_boundaries).build())
    except Exception as e:
        print("Error: Failed to create pattern extractor: %s" % e, file=sys.stderr)
        return 1
    flags = ex.device_flags()
    eng = Engine(a.device)
    names = [t.lower() for t in ITEM_TYPE_NAMES]
    out = sys.stdout.buffer
    if fmt == "csv":
        out.write(b"type,value\n")
    seen = set() if a.unique else None
    t0 = time.perf_counter()
    lines = nbytes = found = 0
    by = [0] * 12
    for path in a.inputs:
        data = np.frombuffer(sys.stdin.buffer.read(), dtype=np.uint8) if path == "-" else np.fromfile(path, dtype=np.uint8)
        items, nl, nb, data = extract_lines(eng, data, flags)
        lines += nl; nbytes += nb
        raw = data.tobytes()
        buf = []
        for t, s, e in items.tolist():
            text = raw[s:e]
            if a.show_candidates:  # spans are relative to the trimmed line, like the reference's (extract_cmd.rs:229-237)
                ls = raw.rfind(b"\n", 0, s) + 1
                while ls < s and raw[ls] in b" \t\x0c\r":
                    ls += 1
                print("[CANDIDATE] %s at %d-%d: %s" % (ITEM_TYPE_NAMES[t], s - ls, e - ls, text.decode("utf-8", "replace")), file=sys.stderr)
            if seen is not None:
                if text in seen:
                    continue
                seen.add(text)
            if fmt == "json":
                buf.append(b'{"type":"' + names[t].encode() + b'","value":"' + text.replace(b"\\", b"\\\\").replace(b'"', b'\\"') + b'"}\n')
            elif fmt == "csv":
                buf.append(names[t].encode() + b',"' + text.replace(b'"', b'""') + b'"\n')
            else:
                buf.append(text + b"\n")
            found += 1
            by[t] += 1
        out.write(b"".join(buf))
    out.flush()
    if a.stats:
        dt = time.perf_counter() - t0
        print("\n[INFO] === Extraction Complete ===\n[INFO] Lines processed: %s\n[INFO] Patterns found: %s" % (format(lines, ","), format(found, ",")), file=sys.stderr)
        for label, t in (("IPv4", 2), ("IPv6", 3), ("Domains", 0), ("Emails", 1)):
            if by[t]:
                print("[INFO]   %s: %s" % (label, format(by[t], ",")), file=sys.stderr)
        print("[INFO] Throughput: %.2f MB/s\n[INFO] Total time: %.2fs" % ((nbytes / 1e6) / dt if dt > 0 else 0.0, dt), file=sys.stderr)
    eng.close()
    return 0


def cmd_build(a):
    """`matchy build` (bin/commands/build_cmd.rs): text / csv / json inputs -> one .mxy file, written read-only (0444) like the
    reference does.  `-f misp` goes through misp_importer.py; schema validation (the `-t <known schema>` case) is not provided."""
    import os
    from . import builder as B
    if a.format not in ("text", "csv", "json", "misp"):
        print("Error: Unknown format: %s. Use 'text', 'csv', 'json', or 'misp'" % a.format, file=sys.stderr)
        return 1
    b = B.DatabaseBuilder(B.MatchMode.CaseInsensitive if a.case_insensitive else B.MatchMode.CaseSensitive)
    if a.database_type:
        if a.database_type in ("threatdb", "ThreatDB-v1"):  # schemas/mod.rs:137-141: the one built-in schema
            print("Error: schema validation for '%s' is not provided; use a custom --database-type name" % a.database_type, file=sys.stderr)
            return 1
        b.set_database_type(a.database_type)
    if a.description:
        b.set_description(a.desc_lang, a.description)
    total = 0
    try:
        if a.format == "misp":
            from .misp_importer import add_misp_files
            total = add_misp_files(b, a.inputs)
            if a.database_type:
                b.set_database_type(a.database_type)
            if a.description:
                b.set_description(a.desc_lang, a.description)
        else:
            add = {"text": B.add_text_file, "csv": B.add_csv_file, "json": B.add_json_file}[a.format]
            for path in a.inputs:
                total += add(b, path)
        st = b.stats()
        if a.verbose or a.debug:
            print("\nBuilding database:\n  Total entries:   %d\n  IP entries:      %d\n  Literal entries: %d\n  Glob entries:    %d"
                  % (total, st["ip_entries"], st["literal_entries"], st["glob_entries"]))
        data = b.build()
    except (OSError, ValueError) as e:
        print("Error: %s" % e, file=sys.stderr)
        return 1
    if os.path.exists(a.output):
        os.chmod(a.output, 0o644)  # (fs::write on a file the previous build left read-only)
    with open(a.output, "wb") as f:
        f.write(data)
    os.chmod(a.output, 0o444)
    if a.verbose or a.debug:
        print("\n\u2713 Database built successfully!\n  Output:        %s\n  Database size: %.2f MB (%d bytes)" % (a.output, len(data) / (1024.0 * 1024.0), len(data)))
    else:
        print("\u2713 Database built: %s" % a.output)
    return 0


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
    m.add_argument("-j", "--threads", default=None, help="1 = sequential mode (line order, real timestamps); anything else: the parallel-mode output")
    m.add_argument("--follow", action="store_true", help="after the existing content, keep processing what is appended to the files")
    m.add_argument("--follow-idle-exit", type=float, default=None, help="(follow mode) stop after this many seconds without new data; default: run until interrupted")
    m.add_argument("--watch-db", action="store_true", help="reload the database into the engine when its file changes")
    x = sub.add_parser("extract", help="extract IoC tokens from log files, one JSON / CSV / text line per item on stdout")
    x.add_argument("inputs", nargs="+", help='log files, or "-" for stdin')
    x.add_argument("--format", default="json")
    x.add_argument("--types", default=None, help="comma list: ipv4, ipv6, ip, domain, email, all")
    x.add_argument("--min-labels", type=int, default=2)
    x.add_argument("--no-boundaries", action="store_true")
    x.add_argument("-u", "--unique", action="store_true")
    x.add_argument("-s", "--stats", action="store_true")
    x.add_argument("--show-candidates", action="store_true")
    x.add_argument("--device", type=int, default=0)
    bl = sub.add_parser("build", help="write a .mxy database from text / csv / json indicator files (host only: no GPU needed)")
    bl.add_argument("inputs", nargs="+")
    bl.add_argument("-o", "--output", required=True)
    bl.add_argument("-f", "--format", default="text")
    bl.add_argument("-t", "--database-type", default=None)
    bl.add_argument("-d", "--description", default=None)
    bl.add_argument("--desc-lang", default="en")
    bl.add_argument("-v", "--verbose", action="store_true")
    bl.add_argument("--debug", action="store_true")
    bl.add_argument("-i", "--case-insensitive", action="store_true")
    a = ap.parse_args(argv)
    if a.cmd == "extract":
        return cmd_extract(a)
    if a.cmd == "build":
        return cmd_build(a)
    from . import Engine, RecordFormatter
    from . import match_modes as M
    sequential = str(a.threads) == "1" or a.follow  # (follow mode prints sequential-mode lines, follow.rs:209-262)
    watch = None
    if a.watch_db:
        watch = M.WatchingDatabase.from_(a.database).device(a.device).chunk_bytes(a.chunk_mb << 20).on_reload(
            lambda gen, path: print("[INFO] database reloaded (generation %d): %s" % (gen, path), file=sys.stderr)).open()
        eng = watch.engine
    else:
        db = open(a.database, "rb").read()
        eng = Engine(a.device, chunk_bytes=a.chunk_mb << 20)
        eng.upload(db)
        fmt = RecordFormatter(db)
    info = eng.db_info()
    crypto = (info["has_literal"] or info["has_glob"]) and "-crypto" not in a.extractors.split(",")
    tot = [None]
    t0 = time.perf_counter()
    nbytes = [0]
    out = sys.stdout.buffer

    def scan_batch(data, path):
        """One buffer through the device scan, rendered in the selected mode."""
        import contextlib
        with (watch.lock() if watch else contextlib.nullcontext()):
            f = watch.formatter() if watch else fmt
            flags = eng.default_flags() | (0xE0 if crypto else 0)  # crypto extractors are on by default when the database has strings (match_cmd.rs:290-292)
            nbytes[0] += data.size
            recs, ids = eng.scan(data, flags)
            c = eng.counters_list()
            tot[0] = c if tot[0] is None else [x + y for x, y in zip(tot[0], c)]
            return M.render_sequential(f, recs, ids, data, path) if sequential else f.ndjson(recs, ids, data, 0, path)

    try:
        if a.follow:
            M.follow_files(a.inputs, scan_batch, out, idle_exit=a.follow_idle_exit)
        else:
            for path in a.inputs:
                out.write(scan_batch(read_input(path), path))
    except (KeyboardInterrupt, ValueError) as e:
        if isinstance(e, ValueError):
            print("Error: %s" % e, file=sys.stderr)
            return 1
    out.flush()
    if a.stats and tot[0] is not None:
        dt = time.perf_counter() - t0
        names = ["lines", "bytes", "candidates", "matches"]
        print(json.dumps({"stats": dict(zip(names, tot[0][:4])), "seconds": round(dt, 3), "MB_per_s": round(nbytes[0] / dt / 1e6, 1)}), file=sys.stderr)
    if watch:
        watch.close()
    else:
        eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
