#!/bin/bash
# Final evidence of round 2 on one B200: GPU tests, smoke(), both bench arms, the ncu launch list of the bench.
# usage (under gpurun): bash tools/r2_final.sh <tag>
tag=${1:-r2i}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -x -q > $o/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 $o/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_ref.json 2> $o/${tag}_bench_ref.err; echo "ref rc=$?"
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-ops --no-e2e > $o/${tag}_ncu_list.log 2>&1; echo "ncu list rc=$?"
