#!/bin/bash
mkdir -p gpurun_out
echo "=== cls tests"; timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_e2e.py tests/test_gpu_host.py -m gpu -q --no-header -p no:cacheprovider -k "cls or config1 or device_resident or golden or predict" > gpurun_out/pytest_cls.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/pytest_cls.log | cut -c1-300
echo "=== c5 cls"; WSI_C5_STRIDES=256,64 timeout 900 python bench.py --config c5 > gpurun_out/bench_c5b.json 2> gpurun_out/bench_c5b.err; echo "exit $?"; tail -n 3 gpurun_out/bench_c5b.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c5b.json"))
for s in d["sweep"]:
    print(s["model"], s["stride"], round(s["ms_per_step"]), round(s["slide_mpx_per_s"],1), {k:(v["ms"], v["share"]) for k,v in s["stages"].items() if k in ("stitch","conv")})
PY
