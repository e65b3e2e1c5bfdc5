# Round-1 closing measurements, session 3 (support box + new histogram sink), one B200.
# Run by hand through gpurun; results are copied from gpurun_out/ into profiles/.
# No number printed under ncu is ever used as a bench value.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/final_main.json 2> gpurun_out/final_main.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
for v in "lung:--workload extract --mask lung" "hist_lung:--workload hist --mask lung" "hist_ones:--workload hist --mask ones" "hist_rois:--workload hist --mask lung --rois 50" "plain:--workload extract --arith plain"; do
  name=${v%%:*}; args=${v#*:}
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline $args > gpurun_out/final_$name.json 2> gpurun_out/final_$name.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:gauss_pass|features_march" -s 4 -c 4 -o gpurun_out/prof_r1b_final -f python profiles/prof_run.py 1.2 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
ncu --set full --clock-control none --import-source on -k regex:features_march -s 6 -c 1 -o gpurun_out/prof_r1b_hist2 -f python profiles/prof_run.py 1.2 --hist --eq > gpurun_out/ncu_hist2.log 2>&1
tail -2 gpurun_out/ncu_hist2.log
