#!/usr/bin/env python3
"""The multi-device handle (bz2b200_create_multi) on every GPU of the box, in ONE process: host-pinned and device-resident buffers,
block sizes -1/-5/-9, the OutputStream facade over all devices; every stream compared with the one-GPU handle's.
  python tools/multi_probe.py [GiB of mixed corpus, default 4] > profiles/rNN_multi_handle.json"""
import hashlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bzip2_opencl_b200 import api, corpus

gib = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ndev = api.load_library().bz2b200_device_count()
unit = corpus.mixed_corpus(1 << 30, workers=min(32, os.cpu_count() or 8))
n = gib << 30
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for k in range(gib):
    host.numpy()[k << 30:(k + 1) << 30] = np.frombuffer(unit, dtype=np.uint8)
res = {"devices": ndev, "input_bytes": n, "workload": f"{gib} GiB = the 1 GiB mixed corpus x{gib}, -9, base 100000"}


def run(c, src_ptr, nbytes, out, reps=2):
    cap = out.numel()
    best, r = None, None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = c.compress_blocks_ptr(src_ptr, nbytes, out.data_ptr(), cap)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return r, best


def digest(out, r):
    h = hashlib.sha256(out[:(r["nbits"] + 7) // 8].numpy().tobytes() if out.device.type == "cpu" else out[:(r["nbits"] + 7) // 8].cpu().numpy().tobytes())
    h.update(np.ascontiguousarray(r["crcs"]).tobytes())
    return h.hexdigest()


one = api.Compressor(9, 600, 100000, device=0)
cap = one.max_output_bytes(n) + 64
out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
r1, t1 = run(one, host.data_ptr(), n, out)
want = digest(out, r1)
res["one_gpu_host_buffers"] = {"MB/s": round(n / 1e6 / t1, 1), "ms": round(1e3 * t1, 2)}
one.close()
many = api.Compressor(9, 600, 100000, devices=list(range(ndev)))
r2, t2 = run(many, host.data_ptr(), n, out)
res["multi_handle_host_buffers"] = {"MB/s": round(n / 1e6 / t2, 1), "ms": round(1e3 * t2, 2), "equals_one_gpu": digest(out, r2) == want, "blocks": len(r2["crcs"])}
d_in = host[:min(n, 4 << 30)].cuda()
d_out = torch.empty(one.max_output_bytes(d_in.numel()) + 64 if False else cap, dtype=torch.uint8, device="cuda")
r3, t3 = run(many, d_in.data_ptr(), d_in.numel(), d_out)
ok3 = True
if d_in.numel() == n:
    ok3 = digest(d_out, r3) == want
res["multi_handle_buffers_on_gpu0"] = {"MB/s": round(d_in.numel() / 1e6 / t3, 1), "ms": round(1e3 * t3, 2), "equals_one_gpu": ok3,
                                        "note": "input and output in GPU 0's HBM: the other devices' slices and parts travel over NVLink (peer copies)"}
del d_in, d_out
many.close()
torch.cuda.empty_cache()
sizes = {}
for mult in (1, 5, 9):
    a = api.Compressor(mult, 600, 100000, device=0)
    b = api.Compressor(mult, 600, 100000, devices=list(range(ndev)))
    m = 1 << 30
    ra, ta = run(a, host.data_ptr(), m, out, reps=1)
    da = digest(out, ra)
    rb, tb = run(b, host.data_ptr(), m, out, reps=1)
    sizes[f"-{mult}"] = {"one_gpu_MB/s": round(m / 1e6 / ta, 1), "multi_MB/s": round(m / 1e6 / tb, 1), "equal": digest(out, rb) == da, "blocks": len(rb["crcs"])}
    a.close(); b.close()
res["block_sizes_1GiB_host_buffers"] = sizes
# the facade over every device: file -> OutputStream -> file
shm = "/dev/shm" if os.access("/dev/shm", os.W_OK) else "/tmp"
fin, fout = os.path.join(shm, "mp.in"), os.path.join(shm, "mp.bz2")
with open(fin, "wb") as f:
    for k in range(gib):
        f.write(unit)
exe = os.path.join(ROOT, "bzip2_opencl_b200", "facade_bench")
fac = {}
for name, env, sink in (("all_devices", {}, fout), ("all_devices_to_null", {}, "/dev/null"), ("one_device", {"BZ2B200_DEVICES": "0"}, fout)):
    best = None
    for _ in range(2):
        r = subprocess.run([exe, fin, sink, "9", "10", "fd"], capture_output=True, text=True, env=dict(os.environ, **env), timeout=900)
        if r.returncode != 0:
            best = {"error": r.stderr[-200:]}
            break
        j = json.loads(r.stdout.strip().splitlines()[-1])
        if best is None or j["ms"] < best["ms"]:
            best = j
    if "ms" in best:
        best["MB/s"] = round(n / 1e6 / (best["ms"] / 1e3), 1)
        if sink == fout:
            best["sha256"] = hashlib.sha256(open(fout, "rb").read()).hexdigest()
    fac[name] = best
res["facade_file_to_file"] = fac
if "sha256" in fac.get("all_devices", {}) and "sha256" in fac.get("one_device", {}):
    res["facade_file_to_file"]["all_devices_equals_one_device"] = fac["all_devices"]["sha256"] == fac["one_device"]["sha256"]
os.unlink(fin)
if os.path.exists(fout):
    os.unlink(fout)
print(json.dumps(res, indent=1))
