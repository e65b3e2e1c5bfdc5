run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/bv.json 2> gpurun_out/bv.err; python - "$@" <<PY
import json,sys,os
d=json.loads(open("gpurun_out/bv.json").read().strip().splitlines()[-1])
print(os.path.basename(os.environ.get("IFE_CUDA_LIB","product")), " ".join(sys.argv[1:]), "|", round(d["value"],2), "Gvox/s", round(d["ms_per_step"],3), "ms", "features", round(d["roofline"]["kernels"]["features_fused"]["ms_per_launch"],3))
PY
}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run --workload hist --mask lung
run --workload hist --mask ones
