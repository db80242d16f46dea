#!/bin/bash
# Round 2, the last 80 seconds of GPU budget: the host-loader parity test (incl. the split packed/raw batches) and
# one short bench with the live packer balanced against the link (MFCD_PACK_FRACTION=auto).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 30 python -m pytest tests/test_gpu_epoch.py -m gpu -q -k host_resident -p no:cacheprovider > $O/pytest_hybrid.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_hybrid.log | cut -c1-300
MFCD_PACK_FRACTION=auto timeout 40 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-rooflines --e2e-format wire8_live > $O/bench_hybrid.json 2> $O/bench_hybrid.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_hybrid.json") if l.startswith("{")][-1]); e = d["e2e"]
    print("value %.4g k1 %s kernel %s" % (d["value"], d["roofline"]["k1_ms"], d["roofline"]["kernel"]))
    print("e2e %.4g %s h2d %s packer %s" % (e["value"], e["format"], e["h2d_bytes_per_step"], e.get("host_packer")))
except Exception as ex:
    print("unreadable", ex); print(open("gpurun_out/bench_hybrid.err").read()[-1200:])
PY
