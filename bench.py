#!/usr/bin/env python
"""bench.py — `matchy match` log-scan throughput on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one pass of the whole hot path (tokenize -> token validation + string filters -> IP-trie / exact string
lookups -> records) over one batch of synthetic log resident in HBM.

  N = 1   BASELINE.json configs[1]: 100 K globs + 1 M literal domains over 10 GB of DNS/proxy log lines.  The line also
          carries `per_config` (configs 1, 3, 4 measured and parity-checked in the same run) and `alt_path` (the single-pass
          scan_kernel on the same workload).
  N > 1   BASELINE.json configs[4]: the mixed 5 M-indicator database over a 200 GB multi-source log, sharded by byte range
          over the ranks (100 / 50 / 25 GB per GPU, generated in HBM by the device generator), database replicated per GPU,
          no data-path collective; the summary counters are all-reduced over NCCL — strong scaling.  `--config C` with
          several GPUs gives every rank its own `--gb` shard of config C instead (weak scaling).

Every line carries `parity`: the device's counters and records against the CPU oracle on the same bytes (untimed).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C] [--gb G] [--total-gb T]

`--impl reference` times the CPU restatement of the reference matcher (oracle/, kind "port": the Rust reference
cannot be compiled in this environment) on the box's host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    1: "cfg1: 10K-indicator CSV-shaped DB (IPs, CIDRs, literal domains, *.evil globs, hashes) over synthetic nginx access log",
    2: "cfg2: 100K paraglob globs + 1M literal domains over synthetic DNS/proxy log (AC-walk heavy)",
    3: "cfg3: 1M IPv4/IPv6 CIDR prefixes over synthetic firewall/netflow log (ip-trie LPM heavy)",
    4: "cfg4: 5M MD5/SHA1/SHA256 hashes over synthetic EDR/process log (literal-hash probe heavy)",
    5: "cfg5: mixed 5M-indicator threat DB over synthetic multi-source log",
}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons while the timed region runs (B200_PROFILING.md's clocks line): NVML every 10 ms
    (a 10 GB step is ~14 ms, so nvidia-smi's ~100 ms per query would see one sample), nvidia-smi as the fallback."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self._stop_evt = index, [], 0.0, set(), threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x for x in vis.split(",") if x.strip()]
        try:
            return int(ids[index]) if ids else index
        except (ValueError, IndexError):
            return index

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in self.REASONS:
            if bits & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            r = [x.strip() for x in out.split(",")]
            self.sm.append(float(r[0])); self.mx = max(self.mx, float(r[1]))
            for (name, _), v in zip(self.REASONS, r[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.01 if self.nvml else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx or None, "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nvml else "nvidia-smi"}


def ncu_traffic(kernel):
    """DRAM bytes per log byte of `kernel` from the committed ncu --set full capture (profiles/r2_traffic.json, else r1), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)["kernels"][kernel]["dram_bytes_per_log_byte"]), name
        except Exception:
            continue
    return None, None


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: keep this rank's threads — and so the pages of the pinned log buffer it is about to allocate and fill —
    on the NUMA node its GPU hangs off (sysfs local_cpulist of the GPU's PCI function).  With 8 ranks streaming 53 GB/s each from
    host memory, buffers that sit on the other socket make the inter-socket link the bottleneck of the e2e figure.  Best effort:
    returns a description, or None when the topology cannot be read (single node, container without sysfs, ...)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        with open(base + "/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if len(use) < 2 or use == allowed:
            return None
        os.sched_setaffinity(0, use)
        node = None
        try:
            with open(base + "/numa_node") as f:
                node = int(f.read().strip())
        except Exception:
            pass
        return {"pci": bus, "numa_node": node, "cpus": len(use)}
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_port_throughput(db, cfg, scale, threads, target_seconds=12.0, offset_blocks=0):
    """Oracle (CPU port of the reference matcher) on all host cores over a bounded sample; returns (GB/s, sample_bytes, cores, counters)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from matchy_b200 import synth_host as synth
    orc = oracle_lib.Oracle(db)
    cores = threads or os.cpu_count() or 1
    probe = synth.gen_log(cfg, 64 << 20, scale, offset=offset_blocks * 65536)
    rate, cache = 0.0, 10000
    for cap in (10000, 0):  # `matchy match --cache-size`: the default and off; the faster one is what gets timed
        orc.set_cache(cap)
        t0 = time.perf_counter()
        orc.scan_mt(probe, threads=cores)
        r = probe.size / max(time.perf_counter() - t0, 1e-6)
        if r > rate:
            rate, cache = r, cap
    orc.set_cache(cache)
    nbytes = int(min(max(rate * target_seconds, 64 << 20), 4 << 30)) // 65536 * 65536
    sample = probe if nbytes <= probe.size else synth.gen_log(cfg, nbytes, scale, offset=offset_blocks * 65536)
    t0 = time.perf_counter()
    cnt = orc.scan_mt(sample, threads=cores)
    dt = time.perf_counter() - t0
    return sample.size / dt / 1e9, int(sample.size), cores, cnt, cache


def parity_gate(eng, db, host, nbytes, flags, base, record_bytes, cores=0):
    """BASELINE.md §4.4, untimed: the engine's result of its last scan of this resident shard (counters over ALL of it,
    records over its first `record_bytes`) against the CPU oracle on the same bytes.  Returns the `parity` object of the
    JSON line; the caller exits non-zero when something differs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib
    orc = oracle_lib.Oracle(db)
    t0 = time.perf_counter()
    got_cnt = eng.counters_list()
    recs, ids = eng.results()
    want_cnt = orc.scan_mt(host[:nbytes], flags=flags, threads=cores)
    counters_equal = got_cnt == want_cnt
    record_bytes = min(record_bytes, nbytes) // 65536 * 65536  # generator blocks end with a newline
    want_recs, want_ids, _ = orc.scan_mt_keep(host[:record_bytes], flags=flags, threads=cores, base=base)
    k = int(np.searchsorted(recs["offset"], base + record_bytes, side="left")) if len(recs) else 0
    sub = recs[:k]
    pat = sub[sub["kind"] == 2]
    n_ids = int(pat["ids_index"][-1]) + int(pat["n_ids"][-1]) if len(pat) else 0
    records_equal = len(sub) == len(want_recs) and sub.tobytes() == want_recs.tobytes() and ids[:n_ids].tobytes() == want_ids.tobytes()
    return {"bytes_counted": int(nbytes), "bytes_compared": int(record_bytes), "counters_equal": bool(counters_equal),
            "records_equal": bool(records_equal), "records_compared": int(len(want_recs)), "id_pairs_compared": int(len(want_ids)),
            "oracle": "oracle/oracle.cpp (CPU restatement of matchy v1.2.2), all host cores", "seconds": round(time.perf_counter() - t0, 1),
            "counters": {"gpu": got_cnt, "oracle": want_cnt} if not counters_equal else None}


def run_reference(args):
    """The reference arm: the CPU matcher on the host cores, rank 0 only.  Loads oracle/liboracle.so and the host-only generator
    library (libmatchy_synth.so) — no product library, no GPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build(load=False)
    from matchy_b200 import synth_host as synth
    cfg, scale = args.config, args.scale
    db = synth.build_db(cfg, scale)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    orc = oracle_lib.Oracle(db)
    cores = os.cpu_count() or 1
    # size one step at ~6 s of CPU work, with the faster of the two cache settings
    probe = synth.gen_log(cfg, 64 << 20, scale)
    rate, cache = 0.0, 10000
    for cap in (10000, 0):
        orc.set_cache(cap)
        t0 = time.perf_counter(); orc.scan_mt(probe, threads=cores); r = probe.size / max(time.perf_counter() - t0, 1e-6)
        if r > rate:
            rate, cache = r, cap
    orc.set_cache(cache)
    nbytes = int(min(max(rate * 6.0, 64 << 20), 2 << 30)) // 65536 * 65536
    sample = probe if nbytes <= probe.size else synth.gen_log(cfg, nbytes, scale)
    times, cnt = [], None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cnt = orc.scan_mt(sample, threads=cores)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    gbs = sample.size * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": "log_scan_throughput", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg], "config": cfg, "db_scale": scale, "sample_bytes_per_step": int(sample.size)},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "mb_per_s_per_core": 1000 * gbs / cores,
                         "query_cache": "per-thread LRU of %d entries" % cache if cache else "off (--cache-size 0: faster on this stream of mostly distinct tokens)",
                         "reference_published": "200-500 MB/s sequential, 400-2000 MB/s parallel (book/src/commands/matchy-match.md:343-350, hardware unspecified)",
                         "sample": "%d MiB of the cfg%d stream per step, one thread per newline-aligned shard, 128 KiB reads" % (sample.size >> 20, cfg)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lines_per_s": cnt[0] * len(times) / total, "matches_per_s": cnt[3] * len(times) / total,
        "note": "CPU restatement (oracle/oracle.cpp) of matchy v1.2.2's matcher; the Rust reference cannot be built here (no cargo/rustc)",
    }
    with open("/proc/self/maps") as f:
        line["native_libraries"] = sorted({os.path.basename(l.split()[-1]) for l in f if "/matchy_b200/" in l or "/oracle/" in l})
    _emit(line)
    return 0


_REAL_STDOUT = None


def _claim_stdout():
    """Everything libraries print to fd 1 (NCCL's version banner, build chatter) goes to stderr; the JSON line alone reaches stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def measure_resident(eng, dev, nbytes, flags, base, steps, warmup, sync_all, sampler=None):
    """`steps` timed scans of the resident shard after `warmup` untimed ones.  Returns device seconds (sum of the per-scan CUDA
    event spans), wall seconds, per-kernel [ms, launches], kernel launches, counters of the last scan."""
    for _ in range(warmup):
        eng.scan_device(dev, nbytes, flags, base=base)
    sync_all()
    if sampler:
        sampler.start()
    step_ms, kern, launches, counters = [], {}, 0, None
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.scan_device(dev, nbytes, flags, base=base)  # inputs are far larger than the 126 MB L2
        t = eng.timing()
        step_ms.append(t["scan_ms"])
        for k, v in t["kernel_ms"].items():
            kern.setdefault(k, [0.0, 0])
            kern[k][0] += v; kern[k][1] += t["launches"][k]
        launches += sum(t["launches"].values()) + t["aux_launches"]
        counters = eng.counters()
    sync_all()
    return sum(step_ms) / 1000.0, time.perf_counter() - t0, kern, launches, counters


def dominant(kern, nbytes, steps, peak):
    dom = max(kern, key=lambda k: kern[k][0])
    dom_ms, dom_launches = kern[dom]
    bytes_per_launch = nbytes * steps / max(dom_launches, 1)  # 1 algorithmic byte per log byte scanned
    achieved = bytes_per_launch / (dom_ms / max(dom_launches, 1) / 1000.0) / 1e9 if dom_ms > 0 else 0.0
    return dom, achieved, bytes_per_launch


def per_config_block(local_rank, args, peak):
    """Configs 1, 3 and 4 (BASELINE.json configs[0], [2], [3]) in the driver's own run: value, dominant kernel, its roofline
    fraction and the parity flags — 4 GB resident each (generated on the device), records compared on the first 1 GB."""
    import numpy as np
    from matchy_b200 import Engine, synth
    out = {}
    nbytes = int(args.per_config_gb * 1e9) // 65536 * 65536
    pbytes = min(nbytes, (1 << 30))
    for cfg in (1, 3, 4):
        db = synth.build_db(cfg, args.scale)
        eng = Engine(local_rank, chunk_bytes=args.chunk_mb << 20)
        eng.upload(db)
        flags = eng.default_flags()
        dev = eng.dev_alloc(nbytes)
        synth.gen_log_device(local_rank, cfg, dev, nbytes, args.scale)
        dev_s, wall_s, kern, _, counters = measure_resident(eng, dev, nbytes, flags, 0, 3, 3, lambda: None)
        dom, achieved, _ = dominant(kern, nbytes, 3, peak)
        host = synth.gen_log(cfg, pbytes, args.scale)
        eng.scan_device(dev, pbytes, flags)
        par = parity_gate(eng, db, host, pbytes, flags, 0, pbytes)
        out["cfg%d" % cfg] = {"workload": WORKLOADS[cfg], "log_bytes": nbytes, "value": nbytes * 3 / dev_s / 1e9, "value_wall": nbytes * 3 / wall_s / 1e9, "unit": "GB/s", "ms_per_step": 1000 * dev_s / 3,
                              "whole_path_frac": nbytes * 3 / dev_s / 1e9 / peak, "dominant_kernel": dom, "dominant_kernel_frac": achieved / peak,
                              "kernel_ms_per_step": {k: v[0] / 3 for k, v in kern.items()}, "matches": counters["matches"], "lines": counters["lines"],
                              "parity": {k: par[k] for k in ("bytes_compared", "counters_equal", "records_equal", "records_compared")}}
        eng.dev_free(dev)
        eng.close()
    return out


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=0, help="BASELINE.json config 1-5; default: 2 on one GPU, 5 (200 GB sharded by byte range) on several")
    ap.add_argument("--gb", type=float, default=10.0, help="one GPU / explicit --config: log bytes per GPU per step, in GB (1e9)")
    ap.add_argument("--total-gb", type=float, default=200.0, help="several GPUs, config 5: total log bytes, split evenly over the ranks")
    ap.add_argument("--scale", type=float, default=1.0, help="database size scale (1.0 = the BASELINE.json counts)")
    ap.add_argument("--chunk-mb", type=int, default=2040, help="scan piece size (MiB, < 2048); pieces of a resident scan are launched back to back")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="skip the per_config block (configs 1, 3, 4 at 4 GB each)")
    ap.add_argument("--per-config-gb", type=float, default=4.0)
    ap.add_argument("--crypto", action="store_true", help="also run the Bitcoin/Ethereum/Monero extractors (matchy match without --extractors=-crypto)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed full-size parity gate against the CPU oracle")
    ap.add_argument("--parity-gb", type=float, default=2.2, help="bytes of the shard whose RECORDS are compared with the oracle's (counters are compared over all of it)")
    args = ap.parse_args()
    if args.impl == "reference":
        if not args.config:
            args.config = 2 if int(os.environ.get("WORLD_SIZE", "1")) == 1 else 5
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if dist:
        dist.barrier()
    import ctypes as C
    import numpy as np
    from matchy_b200 import Engine, synth
    from matchy_b200 import _native as N

    # One GPU: BASELINE.json configs[1], 10 GB resident (weak: a rank's own shard).  Several GPUs: configs[4], the 200 GB stream
    # split by byte range — strong scaling, every shard generated in its GPU's HBM.
    sharded = world > 1 and args.config in (0, 5)
    cfg = args.config or (5 if world > 1 else 2)
    scale = args.scale
    if sharded:
        # the SAME stream for every N: the total is a whole number of 64 KiB generator blocks divisible by 840 (= lcm(1..8)), so
        # that the all-reduced counters of N = 2, 4, 8 (or any N up to 8) must be identical — blocks end at newlines
        total_blocks = int(args.total_gb * 1e9) // 65536 // 840 * 840
        nbytes = (total_blocks // world if total_blocks % world == 0 else int(args.total_gb * 1e9 / world) // 65536) * 65536
    else:
        nbytes = int(args.gb * 1e9) // 65536 * 65536
    base = rank * nbytes
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    db = synth.build_db(cfg, scale)
    eng = Engine(local_rank, chunk_bytes=args.chunk_mb << 20)
    eng.upload(db)
    info = eng.db_info()
    flags = eng.default_flags() | (0xE0 if args.crypto else 0)  # configs run with --extractors=-crypto semantics unless --crypto

    # host copy (pinned) of what the parity gate and the e2e leg need: the whole shard on one GPU, its first part otherwise
    host_bytes = nbytes if not sharded else min(nbytes, max(int(args.parity_gb * 1e9), 4 << 30) // 65536 * 65536)
    pinned = N.lib().mgpu_host_alloc_pinned(host_bytes)
    if not pinned:
        raise RuntimeError("pinned allocation failed")
    host = np.ctypeslib.as_array(C.cast(pinned, C.POINTER(C.c_uint8)), shape=(host_bytes,))
    synth.gen_log(cfg, host_bytes, scale, offset=base, out=host)
    dev = eng.dev_alloc(nbytes)
    t_gen = time.perf_counter()
    synth.gen_log_device(local_rank, cfg, dev, nbytes, scale, offset=base)  # (same bytes as gen_log: tests/test_gpu_parity.py)
    t_gen = time.perf_counter() - t_gen

    def sync_all():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- HBM-resident timing ("value") ----
    eng.set_keep_results(True)
    sampler = ClockSampler(local_rank)
    dev_s, wall_s, kern, launches, counters = measure_resident(eng, dev, nbytes, flags, base, args.steps, args.warmup, sync_all, sampler)
    eng.debug_counters()
    host_us = dict(eng.host_us)  # host side of the last timed step: the whole mgpu_scan_device call, of which result sort / id re-pack / launches + gather
    clocks = sampler.stop()
    if dist:
        tt = torch.tensor([dev_s, wall_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s, wall_s = float(tt[0]), float(tt[1])
        ct = torch.tensor([counters["lines"], counters["bytes"], counters["candidates"], counters["matches"]] + counters["by_type"],
                          dtype=torch.int64, device="cuda")
        dist.all_reduce(ct, op=dist.ReduceOp.SUM)  # the only collective of the path: summary counters over NVLink
        tot = [int(x) for x in ct.tolist()]
    else:
        tot = [counters["lines"], counters["bytes"], counters["candidates"], counters["matches"]] + counters["by_type"]
    total_bytes = nbytes * world * args.steps
    value = total_bytes / dev_s / 1e9

    # ---- parity gate at the benchmarked size (untimed; rank 0's shard) ----
    parity = None
    if rank == 0 and not args.no_parity:
        if host_bytes == nbytes:  # counters over the whole shard, records over its first part
            parity = parity_gate(eng, db, host, nbytes, flags, base, int(args.parity_gb * 1e9))
        else:                     # the shard exists on the device only: counters and records over its first host_bytes
            eng.scan_device(dev, host_bytes, flags, base=base)
            parity = parity_gate(eng, db, host, host_bytes, flags, base, host_bytes)
            parity["note"] = "shard generated on the device; its first %d bytes regenerated on the host for the oracle" % host_bytes

    # ---- end to end through the C ABI with host buffers ("e2e") ----
    e2e = None
    if not args.no_e2e:
        eng.scan(host, flags, base=base)  # warm-up
        sync_all()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.steps):
            recs, ids = eng.scan(host, flags, base=base)
            d2h += recs.nbytes + ids.nbytes + 192
        sync_all()
        e_s = time.perf_counter() - t0
        if dist:
            tt = torch.tensor([e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_s = float(tt[0])
        e2e = {"value": host_bytes * world * args.steps / e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": host_bytes, "d2h_bytes_per_step": d2h // args.steps,
               "timed": "wall clock around mgpu_scan (pinned host buffer -> double-buffered H2D -> kernels -> records copied per piece to pinned host memory), max over ranks",
               "sample": None if host_bytes == nbytes else "the first %d bytes of every rank's shard per step (the 200 GB stream does not fit host memory)" % host_bytes}

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peak()
    dom, achieved, bytes_per_launch = dominant(kern, nbytes, args.steps, peak)
    per_byte, traffic_file = ncu_traffic(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": per_byte * bytes_per_launch if per_byte is not None else None,
                "traffic_source": None if per_byte is None else "profiles/%s: dram__bytes_read.sum + dram__bytes_write.sum per log byte of this kernel (ncu --set full, one 500 MB piece), scaled to this launch size" % traffic_file,
                "peak_source": peak_src, "bytes_per_launch": bytes_per_launch,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kern.items()},
                "whole_path_frac": (nbytes * args.steps / dev_s / 1e9) / peak if world == 1 else (value / world) / peak}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gbs, sample_bytes, cores, _, cache = cpu_port_throughput(db, cfg, scale, 0)
        cpu = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "mb_per_s_per_core": 1000 * gbs / cores,
               "query_cache": "per-thread LRU of %d entries (matchy match --cache-size, database.rs:725-804)" % cache if cache else "off (--cache-size 0: faster on this stream of mostly distinct tokens)",
               "reference_published": "200-500 MB/s sequential, 400-2000 MB/s parallel (book/src/commands/matchy-match.md:343-350, hardware unspecified)",
               "sample": "%d MiB of the same stream, one thread per newline-aligned shard, 128 KiB reads" % (sample_bytes >> 20)}

    eng.dev_free(dev)
    N.lib().mgpu_host_free_pinned(C.c_void_p(pinned))
    eng.close()

    # ---- the other single-GPU configs and the single-pass kernel, in the same run (one GPU only) ----
    per_config = alt = None
    if rank == 0 and world == 1 and cfg == 2 and not args.no_per_config:
        per_config = per_config_block(local_rank, args, peak)
        e2 = Engine(local_rank, chunk_bytes=args.chunk_mb << 20, fused=True)
        e2.upload(db)
        n2 = min(nbytes, int(8e9) // 65536 * 65536)
        d2 = e2.dev_alloc(n2)
        synth.gen_log_device(local_rank, cfg, d2, n2, scale)
        a_s, _, a_kern, _, a_cnt = measure_resident(e2, d2, n2, flags, 0, 3, 3, lambda: None)
        alt = {"what": "scan_kernel (MATCHY_B200_FUSED=1): one pass over the log, 1 KiB tiles by cp.async.bulk into per-warp shared-memory rings; not the default: see DESIGN.md §4",
               "log_bytes": n2, "value": n2 * 3 / a_s / 1e9, "unit": "GB/s", "kernel_ms_per_step": {k: v[0] / 3 for k, v in a_kern.items()}, "matches": a_cnt["matches"]}
        e2.dev_free(d2)
        e2.close()

    if rank == 0:
        line = {
            "metric": "log_scan_throughput", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000 * dev_s / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[cfg], "config": cfg, "db_scale": scale, "log_bytes_per_gpu": nbytes, "log_bytes_total": nbytes * world,
                       "chunk_bytes": args.chunk_mb << 20, "extractor_flags": flags,
                       "l2": "inputs (%.1f GB per GPU) are larger than L2; no flush needed" % (nbytes / 1e9),
                       "db": {k: info[k] for k in ("node_count", "literal_count", "glob_count", "ac_node_count", "file_bytes")},
                       "parallelism": "byte-range shards x%d, database replicated" % world, "rank0_numa_binding": numa,
                       "generated": "on the device (csrc/synth_device.cu), %.2f s for %.1f GB on rank 0" % (t_gen, nbytes / 1e9)},
            "lines_per_s": tot[0] * args.steps / dev_s, "matches_per_s": tot[3] * args.steps / dev_s,
            "counters": {"lines": tot[0], "bytes": tot[1], "candidates": tot[2], "matches": tot[3], "by_type": tot[4:]},
            "wall_s_timed_region": wall_s, "value_wall": total_bytes / wall_s / 1e9, "host_us_last_step": host_us,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity,
            "per_config": per_config, "alt_path": alt,
        }
        _emit(line)
    if dist:
        dist.destroy_process_group()
    bad = parity is not None and not (parity["counters_equal"] and parity["records_equal"])
    if per_config:
        bad = bad or any(not (v["parity"]["counters_equal"] and v["parity"]["records_equal"]) for v in per_config.values())
    if bad:
        sys.stderr.write("PARITY FAILURE: %r %r\n" % (parity, per_config))
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
