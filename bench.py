#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native bzip2 block compressor (contract: see the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size-mb M] [--parallel P]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Metric (BASELINE.json): input MB/s (10^6 bytes of uncompressed input) compressed into a bit-exact .bz2, 900k blocks.
Workload at N=1 = BASELINE configs[1]: 1 GiB mixed synthetic corpus (text + source-code-like + binary), -9, on one B200.
N > 1 is weak scaling: the stream is N GiB (the 1 GiB corpus repeated N times), its block boundaries are found by one
boundary scan, and every rank compresses a contiguous range of ceil(B/N) blocks -- no data-path collective.
A step = one pass of the whole hot path (boundary scan, RLE1+CRC, BWT, MTF/RLE2, Huffman, bit packing) over the rank's shard.

  value     device-resident: input already in HBM, packed bits left in HBM; CUDA events on the launching stream, max over ranks
  e2e       same work through the C ABI with HOST buffers: pinned host input -> H2D -> kernels -> D2H of the bit stream
  roofline  dominant kernel (by summed CUDA-event time over the timed steps) against the measured HBM copy bandwidth
  cpu_baseline  the reference's own code built for the CPU (oracle/_ref), all host threads, on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "input MB/s compressed, 900k blocks, bit-exact .bz2"
UNIT = "MB/s"
BASE, MULT = 100000, 9
SAMPLE_MB = 128  # bounded CPU sample (prefix of the same corpus)



# stdout carries exactly ONE JSON line: everything else that a library may print there (e.g. NCCL's version banner) goes to stderr
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit_line(line: dict) -> None:
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(n_gpus, size_mb):
    return (f"{size_mb} MiB mixed synthetic corpus (text/code/binary segments of 16 MiB, corpus.mixed_corpus), -9 = 900k blocks, base 100000"
            + (f", repeated x{n_gpus} and sharded by contiguous block ranges over {n_gpus} GPUs" if n_gpus > 1 else ", 1 B200"))


def make_corpus(size_mb):
    from bzip2_opencl_b200 import corpus
    workers = max(1, min(32, (os.cpu_count() or 8)))
    return corpus.mixed_corpus(size_mb << 20, workers=workers)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def single_thread_reference(sample: bytes) -> dict:
    """The same reference code on ONE host thread (north_star: single-thread next to all-cores), on the first 16 MiB of the sample."""
    from oracle import refimpl
    part = sample[:16 << 20]
    t = time.perf_counter()
    refimpl.compress(part, MULT, BASE, threads=1)
    dt = time.perf_counter() - t
    return {"value": round(len(part) / 1e6 / dt, 3), "unit": UNIT, "cores": 1, "sample": f"first {len(part) >> 20} MiB, {dt:.1f} s"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref), all host threads, bounded sample."""
    if rank != 0:
        return
    from oracle import refimpl
    threads = os.cpu_count() or 1
    sample = make_corpus(min(SAMPLE_MB, args.size_mb))
    times = []
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        out = refimpl.compress(sample, MULT, BASE, threads=threads)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = len(sample) / 1e6 / (ms / 1e3)
    sample_desc = f"first {len(sample) >> 20} MiB of the workload per step (compressed to {len(out)} bytes)"
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args.gpus, args.size_mb), "sample": sample_desc},
            "cpu_baseline": {"value": round(val, 3), "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample_desc,
                             "single_thread": single_thread_reference(sample)},
            "e2e": {"value": round(val, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size-mb", type=int, default=1024, help="MiB of corpus per GPU")
    ap.add_argument("--parallel", type=int, default=600, help="blocks per device batch (the reference's -p)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from bzip2_opencl_b200 import api, multi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- synthetic input: the 1 GiB corpus on the host (pinned), once in HBM
    data = make_corpus(args.size_mb)
    n1 = len(data)
    host_in = torch.empty(n1, dtype=torch.uint8, pin_memory=True)
    host_in.numpy()[:] = np.frombuffer(data, dtype=np.uint8)
    d_unit = host_in.cuda(non_blocking=False)
    comp = api.Compressor(MULT, args.parallel, BASE, device=local_rank)

    if world > 1:
        d_full = d_unit.repeat(world)            # the N GiB stream; every rank scans it (metadata only leaves the GPU)
        total_bytes = n1 * world
        starts = comp.scan_blocks(ptr=d_full.data_ptr(), nbytes=total_bytes)
        lo, hi, b_lo, b_hi = multi.shard_byte_ranges(starts, total_bytes, world)[rank]
        d_in = d_full[lo:hi]
        # host copy of this rank's byte range for the e2e leg
        host_shard = torch.empty(hi - lo, dtype=torch.uint8, pin_memory=True)
        pos = 0
        while pos < hi - lo:
            o = (lo + pos) % n1
            take = min(n1 - o, hi - lo - pos)
            host_shard[pos:pos + take] = host_in[o:o + take]
            pos += take
    else:
        d_in, host_shard, total_bytes, lo, hi = d_unit, host_in, n1, 0, n1
    n_in = int(d_in.numel())
    cap = comp.max_output_bytes(n_in) + 64
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    host_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident steps
    res = None
    for _ in range(args.warmup):
        res = comp.compress_blocks_ptr(d_in.data_ptr(), n_in, d_out.data_ptr(), cap)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        res = comp.compress_blocks_ptr(d_in.data_ptr(), n_in, d_out.data_ptr(), cap)
        launches += comp.stats()["kernel_launches"]
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    stats = comp.stats()
    # per-kernel times (CUDA events around every launch) come from separate, untimed passes so that the event traffic stays
    # out of the headline number
    comp.set_profiling(True)
    prof_steps = min(2, args.steps)
    for _ in range(prof_steps):
        comp.compress_blocks_ptr(d_in.data_ptr(), n_in, d_out.data_ptr(), cap)
    torch.cuda.synchronize()
    ktimes = comp.kernel_times()
    comp.set_profiling(False)
    value = total_bytes / 1e6 / (ms_dev / 1e3)

    # ---- end to end through the C ABI with host buffers (pinned): H2D + kernels + D2H inside the timed region
    for _ in range(1):
        comp.compress_blocks_ptr(host_shard.data_ptr(), n_in, host_out.data_ptr(), cap)
    barrier()
    t0 = time.perf_counter()
    e2e_res = None
    for _ in range(args.steps):
        e2e_res = comp.compress_blocks_ptr(host_shard.data_ptr(), n_in, host_out.data_ptr(), cap)
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks(1e3 * (time.perf_counter() - t0)) / args.steps
    e2e_value = total_bytes / 1e6 / (ms_e2e / 1e3)
    e2e_stats = comp.stats()
    out_bytes = (e2e_res["nbits"] + 7) // 8
    assert e2e_res["nbits"] == res["nbits"] and (e2e_res["crcs"] == res["crcs"]).all()

    # ---- roofline of the dominant kernel (summed CUDA-event time over the timed steps)
    peak, peak_src = load_peaks()
    side_stream = ("k_rle_walk", "k_rle_tile_seg", "k_rle_reduce")   # overlapped with the main-stream kernels: not part of the step's critical path
    top = max(((k, v) for k, v in ktimes.items() if k not in side_stream), key=lambda kv: kv[1][0])
    name, (k_ms, k_n, k_bytes) = top
    achieved = (k_bytes / 1e9) / (k_ms / 1e3) if k_ms > 0 and k_bytes else None
    step_kernel_ms = sum(v[0] for v in ktimes.values()) / prof_steps
    # DRAM traffic of that kernel per launch from the committed ncu capture of this same command (profiles/), if there is one
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if name in tj["kernels"]:
            traffic, traffic_src = int(tj["kernels"][name]["dram_bytes_per_launch"]), tj.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": name, "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4) if achieved else None, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "launches_per_step": k_n // prof_steps, "avg_launch_ms": round(k_ms / max(1, k_n), 4),
                "share_of_step": round(k_ms / prof_steps / ms_dev, 4),
                "algorithmic_bytes_per_launch": k_bytes // max(1, k_n),
                "kernels_ms_per_step": {k: round(v[0] / prof_steps, 3) for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0]) if v[1]},
                "sum_kernel_ms_per_step": round(step_kernel_ms, 3),
                "stage_ms_last_step": {k[3:]: round(float(stats[k]), 3) for k in ("ms_total", "ms_rle", "ms_bwt_sort0", "ms_bwt_rounds", "ms_bwt_finish", "ms_mtf",
                                                                                  "ms_huffman", "ms_emit", "ms_walk")}}
    n_rle_total = sum_over_ranks(stats["n_rle"])
    rho = sum_over_ranks(stats["bwt_active_volume"]) / max(1.0, n_rle_total)

    # ---- CPU baseline + parity spot check on rank 0 at N=1 (bounded sample of the same workload)
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import refimpl
        threads = os.cpu_count() or 1
        sample = data[:min(SAMPLE_MB, args.size_mb) << 20]
        t = time.perf_counter()
        ref_out = refimpl.compress(sample, MULT, BASE, threads=threads)
        dt = time.perf_counter() - t
        cpu_baseline = {"value": round(len(sample) / 1e6 / dt, 3), "unit": UNIT, "cores": threads, "kind": "reference",
                        "sample": f"first {len(sample) >> 20} MiB of the workload, oracle/_ref (reference source built for the CPU), {dt:.1f} s",
                        "single_thread": single_thread_reference(sample)}
        ours = comp.compress(sample)
        parity = "bit-identical to oracle/_ref on the sample" if ours == ref_out else "MISMATCH vs oracle/_ref on the sample"

    # ---- the decompress side, reported next to the headline: the stream just produced, back to bytes on the device
    decompress = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            stream = comp.compress(data)
            comp.close()
            dcomp = api.Compressor(MULT, max(args.parallel, int(stats["n_blocks"]) + 8), BASE, device=local_rank)  # every block in one batch
            d_s = torch.frombuffer(bytearray(stream), dtype=torch.uint8).cuda()
            d_o = torch.empty(n1 + 64, dtype=torch.uint8, device="cuda")
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                t = time.perf_counter()
                nout = dcomp.decompress_ptr(d_s.data_ptr(), len(stream), d_o.data_ptr(), n1 + 64)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t
                best = dt if best is None else min(best, dt)
            ok = nout == n1 and bool(torch.equal(d_o[:n1], d_unit))
            decompress = {"value": round(n1 / 1e6 / best, 2), "unit": "output MB/s", "ms": round(1e3 * best, 2), "blocks": int(dcomp.stats()["n_blocks"]),
                          "roundtrip_equals_input": ok, "what": "bz2b200_decompress_stream, device-resident, best of 3"}
            dcomp.close()
        except Exception as e:  # the headline number must not depend on this
            decompress = {"error": str(e)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_dev, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": workload_name(world, args.size_mb), "input_bytes_total": total_bytes, "blocks_per_rank": int(stats["n_blocks"]),
                           "blocks_in_flight": args.parallel, "lanes": int(os.environ.get("BZ2B200_LANES", "2")),
                           "stage_ms_note": "stage_ms_last_step sums the lanes (batches overlap), ms_per_step is the wall-clock of the step", "l2": "inputs (>= 1 GiB per rank) and sort state are larger than the 126 MB L2; no flush needed",
                           "rho_prefix_doubling": round(rho, 3), "bwt_rounds": int(stats["bwt_rounds"]),
                           "compressed_bytes_rank0": int(out_bytes), "parity_check": parity},
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "ms_per_step": round(ms_e2e, 3), "h2d_bytes_per_step": int(n_in),
                        "d2h_bytes_per_step": int(out_bytes),
                        "last_step": {k: round(float(e2e_stats[k]), 3) for k in ("ms_total", "ms_h2d", "ms_d2h")}},
                "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "decompress": decompress, "clocks": clocks}
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
