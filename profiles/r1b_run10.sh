# cropped masked smoothing: tests + bench variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { tag=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r1b_$tag.json 2> gpurun_out/r1b_$tag.err; python - gpurun_out/r1b_$tag.json "$@" <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print(" ".join(sys.argv[2:]), "|", round(d["value"],2), "Gvox/s", round(d["ms_per_step"],3), "ms", {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()}, "e2e", e.get("value") and round(e["value"],2), (e.get("batch") or {}).get("value"), d["gpu_launches"])
except Exception as ex:
    print("FAILED", sys.argv[2:], ex); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
}
run main --no-e2e
run lung --mask lung --no-e2e
run hist_lung --workload hist --mask lung
run hist_rois --workload hist --mask lung --rois 50
