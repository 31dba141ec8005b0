#!/bin/bash
# One gpurun call's worth of evidence: GPU parity tests, kernel A/B at full size, the bench line,
# the ncu launch list of the bench command and one `--set full` capture of the dominant kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
# Everything lands in gpurun_out/<tag>_*; summaries worth keeping are copied to profiles/ by hand.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1

echo "== pytest -m gpu"
timeout -k 10 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log

echo "== A/B bitstream vs dfa"
timeout -k 10 400 python tools/ab_flat.py 16 > $OUT/${TAG}_ab.jsonl 2> $OUT/${TAG}_ab.err
cat $OUT/${TAG}_ab.jsonl

if [ -n "$EXP" ]; then
  echo "== experiments"
  bash tools/micro/exp.sh > $OUT/${TAG}_exp.txt 2>&1
  cat $OUT/${TAG}_exp.txt
fi

echo "== bench"
timeout -k 10 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench.json
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_ref.json

echo "== ncu launch list (bench command)"
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/${TAG}_ncu_bench.log 2>&1

echo "== ncu --set full (dominant kernel, 16 GiB)"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:'cgx_flat_jit|scan_flat_kernel|scan_dfa_kernel' \
  -s 3 -c 1 -f -o $OUT/${TAG}_scan_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT
