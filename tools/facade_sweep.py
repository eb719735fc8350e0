"""Development aid: file -> OutputStream facade -> .bz2 on the GPU box for a few settings of the facade's host pipeline (reader threads, first
batch, batch cap) on the 1 GiB mixed corpus; one traced run.  Writes gpurun_out/facade_sweep.json."""
import hashlib, json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bzip2_opencl_b200 import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "bzip2_opencl_b200", "facade_bench")
t0 = time.time()
data = corpus.mixed_corpus(1 << 30, workers=min(32, os.cpu_count() or 8))
fin, fout = "/dev/shm/fs_in.bin", "/dev/shm/fs_out.bz2"
open(fin, "wb").write(data)
res = {"cpu_count": os.cpu_count(), "corpus_s": round(time.time() - t0, 1), "runs": []}
want = None


def run(env, out=fout, reps=2, note=""):
    global want
    best = None
    err = ""
    for _ in range(reps):
        r = subprocess.run([exe, fin, out, "9", "10", "fd"], capture_output=True, text=True, timeout=120, env=dict(os.environ, BZ2B200_DEVICES="0", **env))
        if r.returncode != 0:
            res["runs"].append({"env": env, "error": r.stderr[-300:]})
            return
        j = json.loads(r.stdout.strip().splitlines()[-1])
        err = r.stderr
        if best is None or j["ms"] < best["ms"]:
            best = j
    rec = {"env": env, "note": note, "ms": best["ms"], "setup_ms": best["setup_ms"], "MB/s": round(len(data) / 1e3 / best["ms"], 1)}
    if out == fout:
        got = hashlib.sha256(open(fout, "rb").read()).hexdigest()
        want = want or got
        rec["same_stream"] = got == want
    if env.get("BZ2B200_FACADE_TRACE"):
        rec["trace"] = err.splitlines()[-40:]
    res["runs"].append(rec)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "facade_sweep.json"), "w"), indent=1)


run({"BZ2B200_FACADE_READERS": "8", "BZ2B200_FACADE_FIRST_MB": "128"}, note="round-2 settings except the equal batches")
run({}, note="new defaults")
run({"BZ2B200_FACADE_TRACE": "1"}, reps=1, note="new defaults, traced")
for rd in ("12", "16", "24", "32"):
    run({"BZ2B200_FACADE_READERS": rd})
for first in ("32", "128"):
    run({"BZ2B200_FACADE_FIRST_MB": first})
run({"BZ2B200_FACADE_BLOCKS": "300"}, note="batch cap 257 MiB")
run({"BZ2B200_FACADE_BLOCKS": "200"}, note="batch cap 172 MiB")
run({}, out="/dev/null", note="new defaults, /dev/null sink")
os.unlink(fin)
os.unlink(fout)
print(json.dumps([{k: v for k, v in r.items() if k != "trace"} for r in res["runs"]], indent=0))
