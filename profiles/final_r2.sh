# Round-2 closing measurements, one B200.  Run by hand through gpurun; results are copied from
# gpurun_out/ into profiles/.  No number printed under ncu is ever used as a bench value.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/final_main.json 2> gpurun_out/final_main.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
for v in "lung:--workload extract --mask lung" "plain:--workload extract --arith plain" "hist_ones:--workload hist --mask ones --rois 0"; do
  name=${v%%:*}; args=${v#*:}
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hist --no-slab $args > gpurun_out/final_$name.json 2> gpurun_out/final_$name.err
done
python profiles/cli_walltime.py 400 > gpurun_out/final_cli.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-hist --no-slab > gpurun_out/plain_launch.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-hist --no-slab > gpurun_out/ncu_launch.log 2>&1
python profiles/prof_run.py 1.2 > gpurun_out/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:iir_tma|features_march4" -s 4 -c 4 -o gpurun_out/prof_r2_final -f python profiles/prof_run.py 1.2 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
python profiles/prof_run.py 1.2 --hist --eq > gpurun_out/plain_prof2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:features_march4 -s 2 -c 1 -o gpurun_out/prof_r2_hist -f python profiles/prof_run.py 1.2 --hist --eq > gpurun_out/ncu_hist.log 2>&1
tail -2 gpurun_out/ncu_hist.log
