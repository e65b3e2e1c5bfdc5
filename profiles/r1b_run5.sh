# overlap_scales experiment
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { tag=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r1b_$tag.json 2> gpurun_out/r1b_$tag.err; python - gpurun_out/r1b_$tag.json "$@" <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ".join(sys.argv[2:]), "|", round(d["value"],2), "Gvox/s", round(d["ms_per_step"],3), "ms", {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()}, d["clocks"])
except Exception as e:
    print("FAILED", sys.argv[2:], e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
}
run main
run main_ov --overlap
run lung --mask lung
run lung_ov --mask lung --overlap
run plain_ov --arith plain --overlap
