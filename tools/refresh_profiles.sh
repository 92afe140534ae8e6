#!/bin/bash
# Regenerates the round's evidence under gpurun_out/ on the GPU box (one GPU).  Order matters: every ncu pass
# follows a plain run of the same command that exited 0.  Post-process here with tools/refresh_profiles_post.py.
set -u
R=${1:-r02}
O=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $O/${R}_bench_under_profile_config.json 2> $O/pre.err || { echo "plain run failed"; tail -5 $O/pre.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv $CMD > $O/ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:corrected_fused -s 3 -c 1 -f -o $O/${R}_corrected $CMD > $O/ncu1.log 2>&1
CMDC="python bench.py --mode compat --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMDC > $O/pre_compat.json 2> $O/pre_compat.err && \
ncu --set full --clock-control none --import-source on -k regex:compat_fused -s 3 -c 1 -f -o $O/${R}_compat $CMDC > $O/ncu2.log 2>&1
# window 4096 (C5's shape): one batch of 592 streams x 430 frames per mode
python tools/run_4096_once.py > $O/${R}_4096_rates.txt 2> $O/pre4096.err && {
ncu --set full --clock-control none --import-source on -k regex:compat_fused -s 2 -c 1 -f -o $O/${R}_compat4096 python tools/run_4096_once.py > $O/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:corrected_fused -s 2 -c 1 -f -o $O/${R}_corrected4096 python tools/run_4096_once.py > $O/ncu4.log 2>&1
}
python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err
python bench.py --impl reference --steps 5 --warmup 2 > $O/${R}_bench_reference.json 2> /dev/null
python tools/run_configs.py > $O/${R}_configs_body.md 2> $O/configs.err
python tools/run_configs.py --sweep > $O/${R}_sweep_body.md 2> $O/sweep.err
python tools/stream_sweep.py > $O/${R}_stream_sweep_body.md 2> $O/ssweep.err
python tools/fft_bench.py > $O/${R}_fft_bench.md 2> $O/fft.err
python tools/voices_probe.py > $O/${R}_voices.md 2> $O/voices.err
[ -x tools/micro/exchange_shfl ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/exchange_shfl tools/micro/exchange_shfl.cu
tools/micro/exchange_shfl > $O/${R}_exchange_micro.md 2> $O/micro.err
for f in ncu1 ncu2 ncu3 ncu4; do tail -n 2 $O/$f.log; done; tail -c 300 $O/${R}_bench.json; tail -3 $O/${R}_configs_body.md
