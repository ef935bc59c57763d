#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full ncu capture of the top kernel.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json
for W in "sortpairs_u64_u32_2^28_uniform" "sortpairs_u64_u32_2^28_entropy0.201" "sortkeys_f32_desc_2^28_bits8_24" "sortkeys_f32_desc_2^28" "sortkeys_i64_desc_2^28_bits16_48" "sortkeys_i64_desc_2^28" "sortpairs_u32_u32_2^28_uniform" "sortkeys_u32_2^28_entropy0.201" "sortkeys_u32_2^28_equal" "sortkeys_u32_2^28_few16" "sortkeys_u32_2^28_sorted"; do
  timeout 300 python bench.py --workload "$W" --steps 10 --warmup 3 --no-cpu-baseline >> $OUT/bench_other.jsonl 2>> $OUT/bench.err
done
# the adjacent callers: top-k and segmented sort against the unmodified cub on the same GPU
timeout 600 python tools/topk_bench.py > $OUT/topk_bench.jsonl 2>> $OUT/bench.err
timeout 300 python tools/segmented_bench.py --out $OUT/segmented_bench.jsonl > /dev/null 2>> $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:onesweep_kernel -s 4 -c 1 -f -o $OUT/onesweep_full \
  python tools/prof_one.py "sortkeys_u32_2^28_uniform" 0 2 > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:histogram_kernel -s 1 -c 1 -f -o $OUT/histogram_full \
  python tools/prof_one.py "sortkeys_u32_2^28_uniform" 0 2 >> $OUT/ncu_full.log 2>&1
ls -la $OUT
