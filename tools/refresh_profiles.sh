# Round-2 profile refresh: one GPU call.  Every ncu run comes after the same command has exited 0 without ncu.  The .ncu-rep
# files are summarised on the box and deleted (gpurun_out/ is limited to 64 MiB).
set -x
R=r2
O=gpurun_out
python bench.py > $O/bench_$R.json 2> $O/bench_$R.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$R.json 2> $O/bench_ref_$R.err
python tools/bench_lbs.py --json $O/lbs_sweep_$R.json | tail -8
python tools/bench_configs.py > $O/configs_$R.jsonl; cat $O/configs_$R.jsonl
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-lbs --no-config4 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-lbs --no-config4 > $O/ncu_launch.log 2>&1
python tools/bench_lbs.py --batches 16384 --reps 2 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file $O/launches_lbs_$R.csv python tools/bench_lbs.py --batches 16384 --reps 2 > $O/ncu_lbs_launch.log 2>&1
python tools/profile_fit.py --batch 4096 --iters 100 --reps 2 > $O/plain_fit_$R.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"smplify_fit|tc_" -s 3 -c 3 -o /tmp/fit_$R -f python tools/profile_fit.py --batch 4096 --iters 100 --reps 2 > $O/ncu_fit_$R.log 2>&1
python tools/ncu_summary.py /tmp/fit_$R.ncu-rep > $O/fit_pair_kernel_$R.md
python tools/ncu_lines.py /tmp/fit_$R.ncu-rep 30 smplify_fit > $O/fit_pair_kernel_lines_$R.txt
python tools/ncu_json.py fit /tmp/fit_$R.ncu-rep 4096 > $O/fit_kernel_$R.json
python tools/bench_lbs.py --batches 8192 --reps 1 > $O/plain_lbs_$R.log 2>&1 && ncu --set full --clock-control none -k regex:"tc_|pose_" -s 9 -c 7 -o /tmp/lbs_$R -f python tools/bench_lbs.py --batches 8192 --reps 1 > $O/ncu_lbs_$R.log 2>&1
python tools/ncu_summary.py /tmp/lbs_$R.ncu-rep > $O/lbs_tc_kernels_$R.md
python tools/ncu_json.py lbs /tmp/lbs_$R.ncu-rep 8192 > $O/lbs_tc_kernels_$R.json
python tools/phase_clocks.py run --batch 2368 --pair > $O/phase_pair16_$R.txt; python tools/phase_clocks.py run --batch 1776 --pair > $O/phase_pair12_$R.txt
SMPLB200_FIT_VARIANT=11 python tools/phase_clocks.py run --batch 2368 > $O/phase_tile16_$R.txt
for b in 1024 1776 2368 4096 8192; do echo -n "B=$b "; python tools/profile_fit.py --batch $b --reps 3 | tail -1; done > $O/tile_times_$R.txt; cat $O/tile_times_$R.txt
du -sh $O; ls -la $O | head -40
