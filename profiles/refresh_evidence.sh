#!/bin/bash
# Round evidence in one GPU call: tests, bench (default and the driver's 20-step form), reference arm, microbench,
# order (A) / random schedule / training share, ncu launch list and one --set full capture of the top kernels.
#   gpurun --timeout 1500 -- 'bash profiles/refresh_evidence.sh r02'
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/${R}_gputest_tail.log
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${R}_bench_n1_k20.json 2> /dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2> /dev/null
python profiles/microbench.py --reps 40 --json $O/${R}_microbench.json > $O/${R}_microbench.log 2>&1
python profiles/fork_order.py --json $O/${R}_fork_order.json > $O/fork.log 2>&1
python profiles/fork_order.py --order classic --cpu-steps 4 --json $O/${R}_random_schedule.json > $O/rand.log 2>&1
python profiles/train_share.py --json $O/${R}_train_share_n1.json > $O/ts.log 2>&1
python profiles/host_overhead.py --batch 16 --stable --steps 2000 > $O/${R}_host_overhead_b16.json 2>&1
python profiles/e2e_only.py 1500 > $O/${R}_e2e_only.txt 2>&1
python profiles/short_run.py > $O/${R}_short_run.txt 2>&1
# ncu passes last (never a bench value): launch list of the bench command, then one full capture of the top kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-stage-timing > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'filter2d_kernel|resize_rb_kernel|diffjpeg' -s 7 -c 7 -f -o $O/${R}_full \
    python profiles/prof_kernels.py blur1 resize1 jpeg1 g1 > $O/ncu_full.log 2>&1
ls -la $O | tail -30
