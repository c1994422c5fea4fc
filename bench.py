#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 libcsc hot path.

Metric (BASELINE.json): compress MB/s (10^6 uncompressed bytes per second).
Workload at N=1 (BASELINE.json configs[1]): libcsc -m2 (lazy parser, 6-byte table width 8),
-d64m, 1 GiB of synthetic English-like text on one B200, split csarc -p style into
independent 368 KiB chunks (one libcsc stream per chunk, dict = min(64 MiB, chunk); the chunk
size is part of the parity key and is the same for the reference arm; 368 KiB makes 2850 streams
per GiB, just under the 2960 warps the encoder kernel keeps resident on 148 SMs).  At N>1
every rank compresses its own 1 GiB (weak scaling, contiguous chunk ranges, no collective).

A "step" is one pass of the whole hot path (analyzer -> filters -> match finder -> parser ->
range coder) over the batch.
  value : whole-job MB/s with the chunks already resident in HBM (CSCB200_EncodeChunksDevice),
          timed with CUDA events on the library's launch stream, max over ranks.
  e2e   : the same through the host C ABI (CSCB200_EncodeChunks) from pinned host buffers,
          H2D + D2H + host framing inside the timed region.
  roofline     : the dominant kernel (k_encode) against the measured HBM peak; per-kernel
                 figures for the match-finder / analyzer kernels ride along in "kernels".
  cpu_baseline : the unmodified reference (oracle/_ref) on this box's host cores, one handle
                 per chunk (csarc -p/-t model), bounded sample of the same workload.
`--impl reference` times only that CPU arm.
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
WORKLOAD = "libcsc -m2 -d64m, 1 GiB synthetic English-like text, 2850 independent 368 KiB chunks per GPU"


def gen_text(total: int, rank: int) -> np.ndarray:
    from csc_b200 import synth
    tile = 32 << 20
    out = np.empty(total, dtype=np.uint8)
    for i, off in enumerate(range(0, total, tile)):
        n = min(tile, total - off)
        out[off:off + n] = synth.text(n, 12345 + rank * 1000 + i)
    return out


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


def cpu_reference_run(chunks, nthreads: int):
    """unmodified reference, one handle per chunk on nthreads host threads; returns (seconds, out bytes)"""
    from oracle import refbind
    t = time.perf_counter()
    outs = refbind.run_chunks(chunks, DICT, LEVEL, nthreads)
    dt = time.perf_counter() - t
    return dt, sum(o.size for o in outs)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refbind
    if not refbind.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built and /root/reference absent"}))
        return 0
    cores = os.cpu_count() or 1
    sample = min(TOTAL, max(8, cores) * 32 * CHUNK)   # bounded sample of the same workload: 32 chunks per host thread
    data = gen_text(sample, 0)
    chunks = [data[o:o + CHUNK] for o in range(0, sample, CHUNK)]
    for _ in range(args.warmup):
        cpu_reference_run(chunks[:cores], cores)
    times = []
    for _ in range(args.steps):
        dt, _ = cpu_reference_run(chunks, cores)
        times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = sample / (ms / 1e3) / 1e6
    line = {"impl": "reference", "metric": "compress_MBps", "value": round(val, 2), "unit": "MB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "level": LEVEL, "dict": DICT, "chunk": CHUNK},
            "cpu_baseline": {"value": round(val, 2), "unit": "MB/s", "cores": cores, "kind": "reference",
                             "sample": f"{sample >> 20} MiB of the workload per step ({len(chunks)} chunks), libcsc via oracle/_ref"},
            "e2e": {"value": round(val, 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    global CHUNK, LEVEL, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--total", type=int, default=TOTAL, help="bytes per GPU (default 1 GiB)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--chunk", type=int, default=CHUNK, help="independent chunk size (part of the parity key)")
    ap.add_argument("--level", type=int, default=LEVEL)
    args = ap.parse_args()
    if args.chunk != CHUNK or args.level != LEVEL:
        CHUNK, LEVEL = args.chunk, args.level
        WORKLOAD = f"libcsc -m{LEVEL} -d64m, synthetic English-like text, independent {CHUNK >> 10} KiB chunks"
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

    total = args.total
    data = gen_text(total, rank)
    nchunks = (total + CHUNK - 1) // CHUNK
    sizes = [min(CHUNK, total - i * CHUNK) for i in range(nchunks)]
    # pinned host buffers (e2e) and a device copy (kernel-only)
    h_in = torch.empty(total, dtype=torch.uint8).pin_memory()
    h_in.numpy()[:] = data
    bound = api.encode_bound(CHUNK)
    h_out = torch.empty(nchunks * bound, dtype=torch.uint8).pin_memory()
    d_in = h_in.cuda(non_blocking=False)
    in_ptrs = [h_in.data_ptr() + i * CHUNK for i in range(nchunks)]
    out_ptrs = [h_out.data_ptr() + i * bound for i in range(nchunks)]
    dev_ptrs = [d_in.data_ptr() + i * CHUNK for i in range(nchunks)]
    caps = [bound] * nchunks

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (allocations, clocks)
    for _ in range(max(args.warmup, 3)):
        api.encode_chunks_device(dev_ptrs, sizes, DICT, LEVEL)
    out_sizes = api.encode_chunks_raw(in_ptrs, sizes, out_ptrs, caps, DICT, LEVEL)

    # ---- timed: kernel-only (inputs resident in HBM)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    kern = {}
    gpu_ms = 0.0
    launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dsz = api.encode_chunks_device(dev_ptrs, sizes, DICT, LEVEL)
        st = csc_b200.get_stats()
        gpu_ms += st["ms_total_gpu"]
        launches += st["kernel_launches"]
        for k, v in st.items():
            if k.startswith("ms_"):
                kern[k] = kern.get(k, 0.0) + v
    barrier()
    wall_dev = time.perf_counter() - t0
    # ---- timed: end to end through the host C ABI (pinned host in -> pinned host out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_sizes = api.encode_chunks_raw(in_ptrs, sizes, out_ptrs, caps, DICT, LEVEL)
    barrier()
    wall_e2e = time.perf_counter() - t0
    sampler.stop()
    assert dsz == out_sizes, "device-resident and host paths disagree on stream sizes"

    ms_step = gpu_ms / args.steps
    ms_e2e = 1e3 * wall_e2e / args.steps
    comp = sum(out_sizes)
    ratio = comp / total
    if world > 1:
        t = torch.tensor([ms_step, ms_e2e, 1e3 * wall_dev / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_e2e, wall_dev_ms = [float(x) for x in t.tolist()]
        c = torch.tensor([comp], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        ratio = float(c.item()) / (total * world)
    else:
        wall_dev_ms = 1e3 * wall_dev / args.steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    job_bytes = total * world
    value = job_bytes / (ms_step / 1e3) / 1e6
    e2e = job_bytes / (ms_e2e / 1e3) / 1e6
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    W = 8  # -m2 table width
    per_kernel = {}
    # algorithmic bytes per input byte of each stage
    algo = {"ms_analyze": 1.0, "ms_filters": 2.0, "ms_mf_hash": 1.0 + 12.0, "ms_mf_sort": 3 * 2 * 12.0, "ms_mf_rows": 12.0 + 4.0 * W,
            "ms_mf_extend": 1.0 + 4.0 * W + 2.0 * W, "ms_mf_chains": 2 * (1.0 + 12.0 + 2 * 2 * 12.0 + 12.0 + 4.0 + 1.0 + 4.0 + 2.0),
            "ms_encode": 1.0 + ratio + 6.0 * (W + 2)}
    # DRAM bytes of one k_encode launch from the committed ncu capture of this very configuration (else null)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_k_encode_traffic.json")))
        w = tj["workload"]
        if (w["level"], w["chunk"], w["bytes_per_gpu"]) == (LEVEL, CHUNK, total):
            traffic, traffic_src = int(tj["dram_bytes_per_launch"]), "profiles/r01_k_encode_traffic.json (ncu, one launch)"
    except Exception:
        pass
    for k, bpb in algo.items():
        ms = kern.get(k, 0.0) / args.steps
        if ms > 0:
            gbs = total * bpb / (ms / 1e3) / 1e9
            per_kernel[k[3:]] = {"ms": round(ms, 3), "algo_bytes_per_input_byte": bpb, "GBps": round(gbs, 1),
                                 "frac_of_hbm_peak": round(gbs / peak, 5)}
    enc_ms = kern.get("ms_encode", 0.0) / args.steps
    achieved = total * (1.0 + ratio + 6.0 * (W + 2)) / (enc_ms / 1e3) / 1e9 if enc_ms else 0.0
    line = {
        "metric": "compress_MBps", "value": round(value, 2), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "level": LEVEL, "dict": DICT, "chunk": CHUNK, "bytes_per_gpu": total,
                   "l2": "inputs (1 GiB/GPU) exceed the 126 MB L2", "ratio": round(ratio, 5),
                   "stream_identical_to_libcsc": "checked in tests/test_gpu_parity.py",
                   "wall_ms_per_step_device_path": round(wall_dev_ms, 2)},
        "e2e": {"value": round(e2e, 2), "unit": "MB/s", "h2d_bytes_per_step": int(total), "d2h_bytes_per_step": int(comp)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_encode (parser + range coder, one stream per warp; latency-bound)",
                     "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 6),
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": int(total * (1.0 + ratio + 6.0 * (W + 2))), "peak_source": peak_src,
                     "algorithmic_bytes": "(1 + ratio + 6*(W+2)) B per input byte: window byte, coded bytes, W+2 candidates of 6 B (SURVEY 8d)"},
        "kernels": per_kernel,
        "clocks": sampler.summary(),
    }
    if not args.no_cpu:
        from oracle import refbind
        if refbind.available():
            cores = os.cpu_count() or 1
            sample = min(total, max(8, cores) * 32 * CHUNK)   # 32 chunks per host thread, after one warm-up pass
            chunks = [data[o:o + CHUNK] for o in range(0, sample, CHUNK)]
            cpu_reference_run(chunks[:cores], cores)
            dt, _ = cpu_reference_run(chunks, cores)
            line["cpu_baseline"] = {"value": round(sample / dt / 1e6, 2), "unit": "MB/s", "cores": cores, "kind": "reference",
                                    "sample": f"first {sample >> 20} MiB of the workload ({len(chunks)} chunks), unmodified libcsc "
                                              f"(oracle/_ref), one handle per chunk on {cores} threads"}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "MB/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
