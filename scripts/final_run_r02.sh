# Round 2: the full single-GPU validation (tests, smoke, bench on every workload, reference arm, ncu launch list and
# --set full captures of the dominant kernels).  Outputs under gpurun_out/final2/; scripts/make_profiles_r02.py turns them
# into the tracked summaries under profiles/.
set -x
O=gpurun_out/final2
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.log 2>&1
for w in c1 c2 c4 c5; do timeout 600 python bench.py --steps 20 --warmup 3 --workload $w > $O/bench_$w.log 2>$O/bench_$w.err; done
timeout 900 python bench.py --steps 10 --warmup 3 --workload c3 --windows 512 --no-cpu-baseline > $O/bench_c3.log 2>$O/bench_c3.err
timeout 900 python bench.py --workload c2seq > $O/bench_c2seq.log 2>$O/bench_c2seq.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_c4.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --workload c4 --no-cpu-baseline"
$CMD > $O/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4.csv $CMD > $O/ncu_c4.log 2>&1
python scripts/lin_times.py c4 > $O/lin_times.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_lin_slot -s 5 -c 1 -o $O/lin_c4 -f python scripts/lin_times.py c4 > $O/ncu_lin.log 2>&1
python scripts/lin_times.py c3 >> $O/lin_times.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_lin_slot -s 5 -c 1 -o $O/lin_c3 -f python scripts/lin_times.py c3 > $O/ncu_lin_c3.log 2>&1
python scripts/phase_times.py c4 > $O/phase_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_chol_banded_c2 -s 3 -c 1 -o $O/band_c4 -f python scripts/phase_times.py c4 > $O/ncu_band.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_backsub -s 5 -c 1 -o $O/backsub_c4 -f python scripts/phase_times.py c4 > $O/ncu_backsub.log 2>&1
python scripts/phase_batch.py > $O/phase.log 2>&1
tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log; cat $O/phase.log; ls -la $O | head -50
