# bench variants beyond the default line (each prints a short summary)
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bv.json 2> gpurun_out/bv.err; python - "$@" <<PY
import json,sys
d=json.loads(open("gpurun_out/bv.json").read().strip().splitlines()[-1])
print(" ".join(sys.argv[1:]), "|", round(d["value"],2), "Gvox/s", round(d["ms_per_step"],3), "ms", {k:round(v["ms_per_launch"],3) for k,v in d["roofline"]["kernels"].items()}, "e2e", round(d["e2e"]["value"],2) if d.get("e2e") else None)
PY
}
run --workload extract --mask lung
run --workload hist --mask lung
run --workload hist --mask ones
run --workload hist --mask lung --rois 50
run --workload extract --mask ones --arith plain
