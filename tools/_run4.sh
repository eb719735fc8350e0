timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -k "facade or app or cli" 2>&1 | tail -4
python - <<'PY'
import os, subprocess, json, sys, hashlib
sys.path.insert(0, '.')
from bzip2_opencl_b200 import corpus, api
data = corpus.mixed_corpus(1 << 30, workers=16)
open('/dev/shm/fb.in', 'wb').write(data)
c = api.Compressor(9, 600, 100000); want = hashlib.sha256(c.compress(data)).hexdigest(); c.close()
for mode in ('fd', 'fd', 'vec'):
    r = subprocess.run(['bzip2_opencl_b200/facade_bench', '/dev/shm/fb.in', '/dev/shm/fb.bz2', '9', '10', mode], capture_output=True, text=True)
    j = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else r.stderr
    got = hashlib.sha256(open('/dev/shm/fb.bz2', 'rb').read()).hexdigest()
    print(mode, j, 'MB/s', round(len(data) / 1e3 / j['ms'], 1) if isinstance(j, dict) else '', got == want)
os.unlink('/dev/shm/fb.in'); os.unlink('/dev/shm/fb.bz2')
PY
