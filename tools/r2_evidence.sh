#!/bin/bash
# Round-2 evidence run on one B200: GPU tests, both bench arms, the ncu launch list and one full capture of the sweep.
# usage (under gpurun): bash tools/r2_evidence.sh <tag>
tag=${1:-r2}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q > $o/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?" 
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; echo "ref rc=$?"
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-ops --no-e2e > $o/${tag}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tprf_fused -s 3 -c 1 -o $o/${tag}_fused -f \
  python tools/fused_prof.py 8192 65536 8 2 5 > $o/${tag}_ncu_fused.log 2>&1; echo "ncu full rc=$?"
ncu -i $o/${tag}_fused.ncu-rep --page raw --csv > $o/${tag}_fused_raw.csv 2>/dev/null
tail -3 $o/${tag}_gpu_tests.log
