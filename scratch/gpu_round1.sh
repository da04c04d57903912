set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest3.log 2>&1; tail -5 gpurun_out/pytest3.log
python bench.py --verbose --ref-gpu > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -25 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json | cut -c1-3000
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"; tail -3 gpurun_out/ncu1.log
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmm_rowsplit -c 5 -o gpurun_out/prof_r1_spmm python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"; tail -3 gpurun_out/ncu2.log
