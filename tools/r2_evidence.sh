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
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tprf_fused|k_tprf_grid' -s 6 -c 2 -o $o/${tag}_sweep -f \
  python bench.py --steps 2 --warmup 1 --skip-ops --no-e2e > $o/${tag}_ncu_sweep.log 2>&1; echo "ncu full rc=$?"
ncu -i $o/${tag}_sweep.ncu-rep --page raw --csv > $o/${tag}_sweep_raw.csv 2>/dev/null
python tools/traffic_json.py $o/${tag}_sweep_raw.csv $o/${tag}_tprf_fused_traffic.json
timeout 600 ncu --set full --clock-control none -k regex:k_softmax_strip -s 1 -c 1 -o $o/${tag}_smstrip -f python tools/softmax_once.py > $o/${tag}_ncu_smstrip.log 2>&1
ncu -i $o/${tag}_smstrip.ncu-rep --page raw --csv > $o/${tag}_smstrip_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:k_dgemm -c 2 -o $o/${tag}_dgemm -f python tools/dgemm_once.py > $o/${tag}_ncu_dgemm.log 2>&1
ncu -i $o/${tag}_dgemm.ncu-rep --page raw --csv > $o/${tag}_dgemm_raw.csv 2>/dev/null
tail -3 $o/${tag}_gpu_tests.log
