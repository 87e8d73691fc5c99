set -x
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/final/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/final/smoke.log 2>&1
for w in c1 c2 c4 c5; do timeout 600 python bench.py --steps 20 --warmup 3 --workload $w > gpurun_out/final/bench_$w.log 2>gpurun_out/final/bench_$w.err; done
timeout 900 python bench.py --steps 10 --warmup 3 --workload c3 --windows 512 --no-cpu-baseline > gpurun_out/final/bench_c3.log 2>gpurun_out/final/bench_c3.err
timeout 600 python bench.py > gpurun_out/final/bench_default.log 2>gpurun_out/final/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref_c4.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --workload c4 --no-cpu-baseline"
$CMD > gpurun_out/final/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches_c4.csv $CMD > gpurun_out/final/ncu_c4.log 2>&1
python scripts/lin_times.py c4 > gpurun_out/final/lin_times.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_lin_tile2 -s 5 -c 1 -o gpurun_out/final/lin_c4 -f python scripts/lin_times.py c4 > gpurun_out/final/ncu_lin.log 2>&1
python scripts/phase_times.py c4 > gpurun_out/final/phase_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_chol_banded_c2 -s 3 -c 1 -o gpurun_out/final/band_c4 -f python scripts/phase_times.py c4 > gpurun_out/final/ncu_band.log 2>&1
python scripts/phase_times.py c1 c2 c4 c5 > gpurun_out/final/phase.log 2>&1
tail -3 gpurun_out/final/pytest_gpu.log; tail -1 gpurun_out/final/smoke.log; grep profiled gpurun_out/final/phase.log; ls -la gpurun_out/final | head -40
