set -x
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err
python tools/bench_lbs.py --json gpurun_out/lbs_sweep_r1.json | tail -8
python tools/bench_configs.py > gpurun_out/configs_r1.jsonl; cat gpurun_out/configs_r1.jsonl
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python tools/bench_lbs.py --batches 16384 --reps 2 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_lbs_r1.csv python tools/bench_lbs.py --batches 16384 --reps 2 > gpurun_out/ncu_lbs_launch.log 2>&1
python tools/profile_fit.py --batch 4096 --iters 100 --reps 2 > gpurun_out/plain_fit_r1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"smplify_fit|tc_" -s 3 -c 3 -o gpurun_out/fit_r1_final -f python tools/profile_fit.py --batch 4096 --iters 100 --reps 2 > gpurun_out/ncu_fit_r1.log 2>&1
python tools/bench_lbs.py --batches 8192 --reps 1 > gpurun_out/plain_lbs_r1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tc_ -s 7 -c 5 -o gpurun_out/lbs_r1_final -f python tools/bench_lbs.py --batches 8192 --reps 1 > gpurun_out/ncu_lbs_r1.log 2>&1
SMPLB200_FIT_VARIANT=3 python tools/phase_clocks.py run --batch 2368 > gpurun_out/phase_s16.txt; SMPLB200_FIT_VARIANT=4 python tools/phase_clocks.py run --batch 1776 > gpurun_out/phase_s12.txt; python tools/phase_clocks.py run --batch 32 > gpurun_out/phase_s4.txt
for b in 592 1184 1776 2368; do echo -n "B=$b "; python tools/profile_fit.py --batch $b --reps 3 | tail -1; done > gpurun_out/tile_times.txt; cat gpurun_out/tile_times.txt
tail -c 400 gpurun_out/bench_r1.json; tail -c 300 gpurun_out/bench_ref_r1.json
