# GPU session of round 1 (second half): full GPU tests, benches of configs[1..3], ncu captures of the thermal family
mkdir -p gpurun_out
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_pytest.log
python bench.py > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err
python bench.py --workload cfg3 > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err
python bench.py --workload cfg4 --steps 2 --warmup 1 > gpurun_out/r2_bench_cfg4.json 2> gpurun_out/r2_bench_cfg4.err
PROF_THERMAL=1 PROF_B1=32768 PROF_B4=4096 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_th.csv python profiles/prof_driver.py > gpurun_out/r2_prof_th.log 2>&1
PROF_THERMAL=1 PROF_B1=32768 PROF_B4=4096 ncu --set full --clock-control none --import-source on -k regex:k_resjac -c 1 -s 1 -o gpurun_out/k1_th_r1 -f python profiles/prof_driver.py > gpurun_out/r2_ncu_k1.log 2>&1
PROF_THERMAL=1 PROF_B1=4096 PROF_B4=4096 ncu --set full --clock-control none --import-source on -k regex:k_simulate -c 1 -s 1 -o gpurun_out/k4_th_r1 -f python profiles/prof_driver.py > gpurun_out/r2_ncu_k4.log 2>&1
tail -3 gpurun_out/r2_pytest.log; cut -c1-300 gpurun_out/r2_bench_cfg2.json; cut -c1-300 gpurun_out/r2_bench_cfg3.json; cut -c1-300 gpurun_out/r2_bench_cfg4.json; tail -3 gpurun_out/r2_prof_th.log
