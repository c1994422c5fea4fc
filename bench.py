#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 libcsc hot path.

Metric (BASELINE.json): compress MB/s (10^6 uncompressed bytes per second).
Workload at N=1 (BASELINE.json configs[1]): libcsc -m2 (lazy parser, 6-byte table width 8),
-d64m, 1 GiB of synthetic English-like text on one B200, split csarc -p style into
independent 368 KiB chunks (one libcsc stream per chunk, dict = min(64 MiB, chunk); the chunk
size is part of the parity key and is the same for the reference arm).  At N>1 every rank
compresses its own 1 GiB (weak scaling, contiguous chunk ranges, no collective);
`--scaling strong` splits one fixed job (`--total` bytes) over the ranks instead.

A "step" is one pass of the whole hot path (analyzer -> filters -> match finder -> parser ->
model -> range coder) over the batch.
  value : whole-job MB/s with the chunks already resident in HBM (CSCB200_EncodeChunksDevice),
          timed with CUDA events on the library's launch stream, max over ranks.
  e2e   : the same through the host C ABI (CSCB200_EncodeChunks) from pinned host buffers,
          H2D + D2H + host framing inside the timed region.
  roofline     : the dominant kernel (k_encode) against the measured HBM peak; per-kernel
                 figures for the match-finder / analyzer kernels ride along in "kernels".
  cpu_baseline : the unmodified reference (oracle/_ref) on this box's host cores, one handle
                 per chunk (csarc -p/-t model), bounded sample of the same workload.
  parity_checked : chunks of the timed e2e output byte-compared with the unmodified reference.
  levels / decode / chunk_sweep : the metric's other legs (-m1 / -m3 on text and on the mixed
                 corpus of configs[2], decode of the -m1 / -m3 archives as in configs[4], the
                 chunk-size / ratio trade-off), each measured the same way; `--quick` skips them.
`--impl reference` times only the CPU arm, on the same bytes and with the same config keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LEVEL = 2
DICT = 64 << 20
CHUNK = 368 << 10
TOTAL = 1 << 30
WIDTH = {1: 1, 2: 8, 3: 2, 4: 8, 5: 16}   # 6-byte table width per level (csc_enc.cpp:56-91; -m5: bounded-search rows)


def workload_name(level, chunk, kind="English-like text"):
    return f"libcsc -m{level} -d64m, synthetic {kind}, independent {chunk >> 10} KiB chunks (one stream per chunk), 1 GiB per GPU"


def gen(kind: str, total: int, rank: int) -> np.ndarray:
    from csc_b200 import synth
    tile = 32 << 20
    out = np.empty(total, dtype=np.uint8)
    fn = synth.text if kind == "text" else synth.mixed
    for i, off in enumerate(range(0, total, tile)):
        n = min(tile, total - off)
        out[off:off + n] = fn(n, 12345 + rank * 1000 + i)
    return out


def config_dict(level, chunk, total, kind="text"):
    """identical for both arms (the driver compares them): no run-specific keys in here"""
    return {"workload": workload_name(level, chunk, "English-like text" if kind == "text" else "mixed corpus (text / x86-like / delta tables / random)"),
            "level": level, "dict": DICT, "chunk": chunk, "bytes_per_gpu": total,
            "l2": "inputs (1 GiB/GPU) exceed the 126 MB L2"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._stop_evt = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self._stop_evt.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[2 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(chunks, nthreads: int, level: int, outs=None):
    """unmodified reference, one handle per chunk on nthreads host threads; returns (seconds, streams)"""
    from oracle import refbind
    t = time.perf_counter()
    res = refbind.run_chunks(chunks, DICT, level, nthreads)
    return time.perf_counter() - t, res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refbind
    if not refbind.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built and /root/reference absent"}))
        return 0
    cores = os.cpu_count() or 1
    total = args.total
    data = gen("text", total, 0)
    chunks = [data[o:o + args.chunk] for o in range(0, total, args.chunk)]
    refbind.run_chunks(chunks[:cores], DICT, args.level, cores)   # page in, spin up
    for _ in range(max(0, args.warmup - 1)):
        refbind.run_chunks(chunks[: 4 * cores], DICT, args.level, cores)
    times, comp = [], 0
    for _ in range(args.steps):
        dt, outs = cpu_reference_run(chunks, cores, args.level)
        comp = sum(o.size for o in outs)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = total / (ms / 1e3) / 1e6
    line = {"impl": "reference", "metric": "compress_MBps", "value": round(val, 2), "unit": "MB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args.level, args.chunk, total), "ratio": round(comp / total, 5),
            "cpu_baseline": {"value": round(val, 2), "unit": "MB/s", "cores": cores, "kind": "reference",
                             "sample": f"the whole workload per step ({total >> 20} MiB, {len(chunks)} chunks), unmodified libcsc "
                                       f"(oracle/_ref), one handle per chunk on {cores} host threads"},
            "e2e": {"value": round(val, 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


class Arena:
    """pinned host in / out buffers + a device copy of one corpus, chunked"""

    def __init__(self, torch, api, data, chunk):
        self.total = int(data.size)
        self.chunk = chunk
        self.n = (self.total + chunk - 1) // chunk
        self.sizes = [min(chunk, self.total - i * chunk) for i in range(self.n)]
        self.h_in = torch.empty(self.total, dtype=torch.uint8).pin_memory()
        self.h_in.numpy()[:] = data
        self.bound = api.encode_bound(chunk)
        self.h_out = torch.empty(self.n * self.bound, dtype=torch.uint8).pin_memory()
        self.d_in = self.h_in.cuda(non_blocking=False)
        self.in_ptrs = [self.h_in.data_ptr() + i * chunk for i in range(self.n)]
        self.out_ptrs = [self.h_out.data_ptr() + i * self.bound for i in range(self.n)]
        self.dev_ptrs = [self.d_in.data_ptr() + i * chunk for i in range(self.n)]
        self.caps = [self.bound] * self.n

    def stream(self, i, size):
        return self.h_out.numpy()[i * self.bound: i * self.bound + size]


def timed_device(api, csc, A, level, steps):
    """device-resident path: CUDA-event time of the library's launch stream, summed per kernel"""
    kern, gpu_ms, launches, dsz = {}, 0.0, 0, None
    for _ in range(steps):
        dsz = api.encode_chunks_device(A.dev_ptrs, A.sizes, DICT, level)
        st = csc.get_stats()
        gpu_ms += st["ms_total_gpu"]
        launches += st["kernel_launches"]
        for k, v in st.items():
            if k.startswith("ms_"):
                kern[k] = kern.get(k, 0.0) + v
    return gpu_ms / steps, {k: v / steps for k, v in kern.items()}, launches, dsz


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--total", type=int, default=TOTAL, help="bytes per GPU (weak) or of the whole job (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg and the parity sample")
    ap.add_argument("--quick", action="store_true", help="headline only: skip the levels / decode / chunk-sweep legs")
    ap.add_argument("--chunk", type=int, default=CHUNK, help="independent chunk size (part of the parity key)")
    ap.add_argument("--level", type=int, default=LEVEL)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import csc_b200
    from csc_b200 import api
    csc_b200.set_device(local)
    if not csc_b200.available():
        raise RuntimeError("no CUDA device: csc_b200 has no CPU fallback")
    level, chunk = args.level, args.chunk

    # weak: every rank its own `total`; strong: contiguous chunk range of one fixed job
    if args.scaling == "strong":
        nch = (args.total + chunk - 1) // chunk
        lo, hi = api.chunk_range(nch, rank, world)   # the library's own partition rule
        job = gen("text", args.total, 0)
        data = job[lo * chunk: min(hi * chunk, args.total)]
        job_bytes = args.total
    else:
        data = gen("text", args.total, rank)
        job_bytes = args.total * world
    total = int(data.size)
    A = Arena(torch, api, data, chunk)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (allocations, clocks)
    for _ in range(max(args.warmup, 3)):
        api.encode_chunks_device(A.dev_ptrs, A.sizes, DICT, level)
    out_sizes = api.encode_chunks_raw(A.in_ptrs, A.sizes, A.out_ptrs, A.caps, DICT, level)

    # ---- timed: kernel-only (inputs resident in HBM)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    ms_step, kern, launches, dsz = timed_device(api, csc_b200, A, level, args.steps)
    barrier()
    wall_dev = time.perf_counter() - t0
    # ---- timed: end to end through the host C ABI (pinned host in -> pinned host out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_sizes = api.encode_chunks_raw(A.in_ptrs, A.sizes, A.out_ptrs, A.caps, DICT, level)
    barrier()
    wall_e2e = time.perf_counter() - t0
    sampler.stop()
    assert dsz == out_sizes, "device-resident and host paths disagree on stream sizes"

    ms_e2e = 1e3 * wall_e2e / args.steps
    comp = sum(out_sizes)
    ratio = comp / total
    if world > 1:
        t = torch.tensor([ms_step, ms_e2e, 1e3 * wall_dev / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_e2e, wall_dev_ms = [float(x) for x in t.tolist()]
        c = torch.tensor([comp, total], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        ratio = float(c[0].item()) / float(c[1].item())
    else:
        wall_dev_ms = 1e3 * wall_dev / args.steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = job_bytes / (ms_step / 1e3) / 1e6
    e2e = job_bytes / (ms_e2e / 1e3) / 1e6
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    W = WIDTH[level]

    def kernel_table(kern, total, ratio, W):
        """per-stage time, implementation bytes per input byte (what the stage really moves) and GB/s"""
        impl = {"ms_analyze": 1.0, "ms_filters": 2.0, "ms_mf_hash": 1.0 + 12.0, "ms_mf_sort": 3 * 2 * 12.0, "ms_mf_rows": 12.0 + 4.0 * W,
                "ms_mf_extend": 1.0 + 4.0 * W + 2.0 * W, "ms_mf_chains": 2 * (1.0 + 12.0 + 2 * 2 * 12.0 + 12.0 + 4.0 + 1.0 + 4.0 + 2.0),
                "ms_encode": 1.0 + 6.0 * (W + 2), "ms_rc": 2.0 * 3.0 + ratio}
        out = {}
        for k, bpb in impl.items():
            ms = kern.get(k, 0.0)
            if ms > 0:
                gbs = total * bpb / (ms / 1e3) / 1e9
                out[k[3:]] = {"ms": round(ms, 3), "bytes_per_input_byte": bpb, "GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 5)}
        mf = sum(kern.get(k, 0.0) for k in ("ms_mf_hash", "ms_mf_sort", "ms_mf_rows", "ms_mf_extend", "ms_mf_chains"))
        if mf > 0:   # the match-finder stage as a whole against SURVEY 8d's compulsory 1 + 6 (W + 2) bytes per position
            gbs = total * (1.0 + 6.0 * (W + 2)) / (mf / 1e3) / 1e9
            out["match_finder_total"] = {"ms": round(mf, 3), "bytes_per_input_byte": 1.0 + 6.0 * (W + 2), "GBps": round(gbs, 1),
                                         "frac_of_hbm_peak": round(gbs / peak, 5), "note": "algorithmic bytes (SURVEY 8d), all MF kernels"}
        return out

    per_kernel = kernel_table(kern, total, ratio, W)
    # DRAM bytes of one k_encode launch from the committed ncu capture of this very configuration (else null)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_k_encode_traffic.json")))
        w = tj["workload"]
        if (w["level"], w["chunk"], w["bytes_per_gpu"]) == (level, chunk, total):
            traffic, traffic_src = int(tj["dram_bytes_per_launch"]), "profiles/r02_k_encode_traffic.json (ncu, one launch)"
    except Exception:
        pass
    enc_ms = kern.get("ms_encode", 0.0)
    algo_bytes = total * (1.0 + 6.0 * (W + 2) + 2.0 * 3.0)   # window byte, W+2 candidates of 6 B, ~3 decisions of 2 B written
    achieved = algo_bytes / (enc_ms / 1e3) / 1e9 if enc_ms else 0.0
    line = {
        "metric": "compress_MBps", "value": round(value, 2), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(level, chunk, args.total), "ratio": round(ratio, 5),
        "e2e": {"value": round(e2e, 2), "unit": "MB/s", "h2d_bytes_per_step": int(total), "d2h_bytes_per_step": int(comp)},
        "gpu_launches": int(launches),
        "wall_ms_per_step_device_path": round(wall_dev_ms, 2),
        "roofline": {"bound": "hbm", "kernel": "k_encode (parser + adaptive model, one stream per warp; latency-bound)",
                     "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 6),
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": int(algo_bytes), "peak_source": peak_src,
                     "algorithmic_bytes": "(1 + 6*(W+2) + 6) B per input byte: window byte, W+2 candidates of 6 B (SURVEY 8d), ~3 decisions of 2 B"},
        "kernels": per_kernel,
        "clocks": sampler.summary(),
    }

    from oracle import refbind
    have_ref = refbind.available() and not args.no_cpu
    cores = os.cpu_count() or 1
    if have_ref:
        # ---- parity of the timed output: first / middle / last / random chunks against the unmodified reference
        rng = np.random.default_rng(7)
        pick = sorted(set([0, 1, A.n // 2, A.n - 1] + [int(x) for x in rng.integers(0, A.n, size=44)]))
        ins = [data[i * chunk: i * chunk + A.sizes[i]] for i in pick]
        exp = refbind.run_chunks(ins, DICT, level, cores)
        same = sum(1 for i, e in zip(pick, exp) if out_sizes[i] == e.size and bool((A.stream(i, out_sizes[i]) == e).all()))
        line["parity_checked"] = {"n": len(pick), "identical": same, "against": "unmodified libcsc (oracle/_ref), same level / dict / chunking"}
        # ---- CPU baseline: the reference on this box's host cores, bounded sample of the same chunks
        sample = min(total, max(8, cores) * 32 * chunk)   # 32 chunks per host thread, after one warm-up pass
        chunks = [data[o:o + chunk] for o in range(0, sample, chunk)]
        cpu_reference_run(chunks[:cores], cores, level)
        dt, _ = cpu_reference_run(chunks, cores, level)
        line["cpu_baseline"] = {"value": round(sample / dt / 1e6, 2), "unit": "MB/s", "cores": cores, "kind": "reference",
                                "sample": f"first {sample >> 20} MiB of the workload ({len(chunks)} chunks), unmodified libcsc "
                                          f"(oracle/_ref), one handle per chunk on {cores} threads"}
    elif not args.no_cpu:
        line["cpu_baseline"] = {"value": None, "unit": "MB/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}

    if not args.quick and world == 1:
        legs = {}

        def leg(name, A2, data2, lvl, kind, with_cpu=True):
            api.encode_chunks_device(A2.dev_ptrs, A2.sizes, DICT, lvl)
            ms, kk, _, _ = timed_device(api, csc_b200, A2, lvl, 2)
            api.encode_chunks_raw(A2.in_ptrs, A2.sizes, A2.out_ptrs, A2.caps, DICT, lvl)
            t1 = time.perf_counter()
            osz = api.encode_chunks_raw(A2.in_ptrs, A2.sizes, A2.out_ptrs, A2.caps, DICT, lvl)
            e = time.perf_counter() - t1
            r = sum(osz) / A2.total
            Wl = WIDTH[lvl]
            enc = kk.get("ms_encode", 0.0)
            ach = A2.total * (1.0 + 6.0 * (Wl + 2) + 6.0) / (enc / 1e3) / 1e9 if enc else 0.0
            d = {"value": round(A2.total / ms / 1e3, 2), "unit": "MB/s", "ms_per_step": round(ms, 2),
                 "e2e": {"value": round(A2.total / e / 1e6, 2), "unit": "MB/s", "h2d_bytes_per_step": A2.total, "d2h_bytes_per_step": int(sum(osz))},
                 "config": config_dict(lvl, A2.chunk, A2.total, kind), "ratio": round(r, 5),
                 "roofline": {"bound": "hbm", "kernel": "k_encode", "achieved": round(ach, 2), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 6), "traffic": None},
                 "kernels_ms": {k[3:]: round(v, 2) for k, v in kk.items() if v > 0.05}}
            if have_ref and with_cpu:
                smp = min(A2.total, cores * 8 * A2.chunk)
                cs = [data2[o:o + A2.chunk] for o in range(0, smp, A2.chunk)]
                dt2, ex = cpu_reference_run(cs, cores, lvl)
                d["cpu_baseline"] = {"value": round(smp / dt2 / 1e6, 2), "unit": "MB/s", "cores": cores, "kind": "reference",
                                     "sample": f"first {smp >> 20} MiB ({len(cs)} chunks)"}
                idn = sum(1 for i, x in enumerate(ex) if osz[i] == x.size and bool((A2.stream(i, osz[i]) == x).all()))
                d["parity_checked"] = {"n": len(ex), "identical": idn}
            legs[name] = d
            return osz

        def decode_leg(name, A2, osz):
            ins = [A2.out_ptrs[i] for i in range(A2.n)]
            h_back = torch.empty(A2.total + 64 * A2.n, dtype=torch.uint8).pin_memory()
            outs = [h_back.data_ptr() + i * (A2.chunk + 64) for i in range(A2.n)]
            api.decode_chunks_raw(ins, osz, outs, A2.sizes)
            t1 = time.perf_counter()
            got = api.decode_chunks_raw(ins, osz, outs, A2.sizes)
            e = time.perf_counter() - t1
            st = csc_b200.get_stats()
            ok = got == A2.sizes and all(bool((h_back.numpy()[i * (A2.chunk + 64): i * (A2.chunk + 64) + A2.sizes[i]] ==
                                              A2.h_in.numpy()[i * A2.chunk: i * A2.chunk + A2.sizes[i]]).all()) for i in (0, A2.n // 2, A2.n - 1))
            r = sum(osz) / A2.total
            kms = st["ms_decode"] + st["ms_unfilter"]
            # the same archives four times over in ONE call (configs[4] decodes 16 GB of archives): the decoder is a latency
            # chain per stream, so its throughput is the number of streams a wave keeps resident
            R = 4
            big = torch.empty(R * (A2.total + 64 * A2.n), dtype=torch.uint8).pin_memory()
            outs4 = [big.data_ptr() + (k * A2.n + i) * (A2.chunk + 64) for k in range(R) for i in range(A2.n)]
            api.decode_chunks_raw(ins * R, list(osz) * R, outs4, A2.sizes * R)   # (buffers of this size are made here, outside the timed call)
            t1 = time.perf_counter()
            got4 = api.decode_chunks_raw(ins * R, list(osz) * R, outs4, A2.sizes * R)
            e4 = time.perf_counter() - t1
            st4 = csc_b200.get_stats()
            ok4 = got4 == A2.sizes * R and bool((big.numpy()[(R - 1) * A2.n * (A2.chunk + 64): (R - 1) * A2.n * (A2.chunk + 64) + A2.sizes[0]] ==
                                                 A2.h_in.numpy()[: A2.sizes[0]]).all())
            batch = {"chunks": R * A2.n, "bytes": R * A2.total, "value": round(R * A2.total / st4["ms_total_gpu"] / 1e3, 2),
                     "e2e": round(R * A2.total / e4 / 1e6, 2), "unit": "MB/s", "waves": st4["waves"],
                     "kernels_ms": {"decode": round(st4["ms_decode"], 2), "unfilter": round(st4["ms_unfilter"], 2)}, "roundtrip_ok": ok4}
            del big
            return {"value": round(A2.total / st["ms_total_gpu"] / 1e3, 2), "unit": "MB/s (uncompressed, device time incl. copies)", "batch_x4": batch,
                    "e2e": {"value": round(A2.total / e / 1e6, 2), "unit": "MB/s", "h2d_bytes_per_step": int(sum(osz)), "d2h_bytes_per_step": A2.total},
                    "kernels_ms": {"decode": round(st["ms_decode"], 2), "unfilter": round(st["ms_unfilter"], 2)},
                    "roofline": {"bound": "hbm", "kernel": "k_decode", "achieved": round(A2.total * (1 + r) / (kms / 1e3) / 1e9, 2) if kms else 0.0,
                                 "peak": peak, "unit": "GB/s", "frac": round(A2.total * (1 + r) / (kms / 1e3) / 1e9 / peak, 6) if kms else 0.0, "traffic": None},
                    "roundtrip_ok": bool(ok), "archive": name}

        dec = {}
        o1 = leg("m1_text", A, data, 1, "text")
        dec["m1_text"] = decode_leg("m1_text", A, o1)
        o3 = leg("m3_text", A, data, 3, "text")
        dec["m3_text"] = decode_leg("m3_text", A, o3)
        # the chunk-size / ratio trade-off of the headline level (csarc's own minimum split is 1 MiB, csarc.cpp:533-535)
        sweep = []
        for ck in (1 << 20, 4 << 20):
            A3 = Arena(torch, api, data, ck)
            api.encode_chunks_device(A3.dev_ptrs, A3.sizes, DICT, level)
            ms, _, _, dz = timed_device(api, csc_b200, A3, level, 1)
            sweep.append({"chunk": ck, "MBps": round(total / ms / 1e3, 2), "ratio": round(sum(dz) / total, 5)})
            del A3
        sweep.insert(0, {"chunk": chunk, "MBps": round(value, 2), "ratio": round(ratio, 5)})
        if have_ref:
            t1 = time.perf_counter()
            one = refbind.encode(data[: 64 << 20], DICT, level)
            sweep.append({"chunk": "unchunked (reference, first 64 MiB as ONE stream, 1 host thread)", "MBps": round((64 << 20) / (time.perf_counter() - t1) / 1e6, 2),
                          "ratio": round(one.size / (64 << 20), 5)})
        del A
        torch.cuda.empty_cache()
        mixed = gen("mixed", args.total, rank)
        AM = Arena(torch, api, mixed, chunk)
        om1 = leg("m1_mixed", AM, mixed, 1, "mixed")
        dec["m1_mixed"] = decode_leg("m1_mixed", AM, om1)
        om3 = leg("m3_mixed", AM, mixed, 3, "mixed")
        dec["m3_mixed"] = decode_leg("m3_mixed", AM, om3)
        line["levels"] = legs
        line["decode"] = dec
        line["chunk_sweep"] = sweep
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
