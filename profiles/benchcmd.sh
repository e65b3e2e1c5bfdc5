python -m pytest tests -m gpu -x -q 2>&1 | tail -3; python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_tmp.json").read().strip().splitlines()[-1])
print(round(d["value"],2), round(d["ms_per_step"],3), {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
print(d["parity"]["values_differing"], d["clocks"])
PY
