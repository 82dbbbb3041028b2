# Round-2 profile refresh after the small-batch cluster kernel: one GPU call.  Every ncu run comes after the same command has
# exited 0 without ncu; the .ncu-rep stays on the box (gpurun_out/ is limited to 64 MiB), its summary comes back.
set -x
R=r2
O=gpurun_out
python bench.py > $O/bench_$R.json 2> $O/bench_$R.err
python tools/bench_configs.py > $O/configs_$R.jsonl; cat $O/configs_$R.jsonl
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-lbs --no-config4 > /dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference --no-lbs --no-config4 > $O/ncu_launch.log 2>&1
python tools/profile_fit.py --batch 32 --iters 100 --reps 2 > $O/plain_split_$R.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"smplify_fit_split" -s 1 -c 1 -o /tmp/split_$R -f python tools/profile_fit.py --batch 32 --iters 100 --reps 2 > $O/ncu_split_$R.log 2>&1
python tools/ncu_summary.py /tmp/split_$R.ncu-rep > $O/fit_split_kernel_$R.md
python tools/phase_clocks.py run --batch 32 --split > $O/phase_split8_$R.txt; python tools/phase_clocks.py run --batch 100 --split > $O/phase_split4_$R.txt; python tools/phase_clocks.py run --batch 256 --split > $O/phase_split2_$R.txt
SMPLB200_FIT_VARIANT=11 python tools/phase_clocks.py run --batch 32 > $O/phase_tile4_$R.txt
for b in 8 32 64 100 148 256 296; do echo -n "B=$b "; python tools/profile_fit.py --batch $b --reps 3 | tail -1; done > $O/split_times_$R.txt; cat $O/split_times_$R.txt
for b in 32 256; do echo -n "tile kernel B=$b "; SMPLB200_FIT_VARIANT=11 python tools/profile_fit.py --batch $b --reps 3 | tail -1; done >> $O/split_times_$R.txt
du -sh $O
