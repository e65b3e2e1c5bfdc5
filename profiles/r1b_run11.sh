# box skip in the extract feature kernel: A/B against the build before it
L=$PWD/image-feature-extraction_b200/lib
run() { tag=$1; lib=$2; shift; shift; IFE_CUDA_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r1b_$tag.json 2> gpurun_out/r1b_$tag.err; python - gpurun_out/r1b_$tag.json $tag <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "|", round(d["value"],2), "Gvox/s", round(d["ms_per_step"],3), "ms", {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()})
except Exception as ex:
    print("FAILED", sys.argv[2:], ex); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
}
run main_pre $L/exp/libife_prebox.so
run main_new $L/libife_cuda.so
run lung_pre $L/exp/libife_prebox.so --mask lung
run lung_new $L/libife_cuda.so --mask lung
timeout 600 python -m pytest tests -m gpu -x -q -k "support_box or full_size or multiscale or ragged or overlap or host_tools" 2>&1 | tail -3
