set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest4.log 2>&1; tail -3 gpurun_out/pytest4.log
python bench.py --verbose --ref-gpu > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench rc=$?"; tail -12 gpurun_out/bench_r1b.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"spmm_rowsplit|transpose_pass|bitmap" -c 13 -o gpurun_out/prof_r1_spmm python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
