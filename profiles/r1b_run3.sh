# session 3, pass 3: histogram sink with 32-bit offsets + byte counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1b_tests.log
tail -4 gpurun_out/r1b_tests.log
L=image-feature-extraction_b200/lib
for lib in $L/exp/libife_old.so $L/libife_cuda.so; do
  IFE_CUDA_LIB=$PWD/$lib timeout 300 python profiles/exp_hist.py 2>&1 | tail -1
done | tee gpurun_out/r1b_ab3.log
run() { tag=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r1b_$tag.json 2> gpurun_out/r1b_$tag.err; python - gpurun_out/r1b_$tag.json "$@" <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ".join(sys.argv[2:]), "|", round(d["value"],2), "Gvox/s", round(d["ms_per_step"],3), "ms", {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()}, "e2e", round(d["e2e"]["value"],2) if d.get("e2e") else None, d["clocks"])
except Exception as e:
    print("FAILED", sys.argv[2:], e)
PY
}
run hist_lung --workload hist --mask lung
run hist_ones --workload hist --mask ones
run hist_rois --workload hist --mask lung --rois 50
