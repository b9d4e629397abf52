# Closing GPU session of round 1 (after the paired vector passes and the K1 write-loop change): GPU tests, smoke,
# bench lines of all configs + the reference arm, launch list of the default bench command, ncu captures of K1 / K4.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/re_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/re_smoke.log 2>&1
python bench.py > gpurun_out/re_bench_cfg2.json 2> gpurun_out/re_bench_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/re_bench_cfg2_reference.json 2>/dev/null
python bench.py --workload cfg3 > gpurun_out/re_bench_cfg3.json 2> gpurun_out/re_bench_cfg3.err
python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/re_bench_cfg4.json 2> gpurun_out/re_bench_cfg4.err
python bench.py --workload cfg5 --steps 3 --warmup 3 > gpurun_out/re_bench_cfg5.json 2> gpurun_out/re_bench_cfg5.err
python bench.py --workload cfg5n10 --steps 3 --warmup 3 > gpurun_out/re_bench_cfg5n10.json 2> gpurun_out/re_bench_cfg5n10.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1e_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/re_ncu_bench.log 2>&1
PROF_B1=65536 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_resjac -c 1 -s 1 -o gpurun_out/k1_r1i -f python profiles/prof_driver.py > gpurun_out/re_ncu_k1.log 2>&1
PROF_B1=8192 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_simulate -c 1 -s 1 -o gpurun_out/k4_r1i -f python profiles/prof_driver.py > gpurun_out/re_ncu_k4.log 2>&1
for k in k1_r1i k4_r1i; do python profiles/ncu_extract.py gpurun_out/$k.ncu-rep > gpurun_out/${k}_ncu_summary.txt 2>/dev/null; done
cat gpurun_out/re_pytest.log; tail -2 gpurun_out/re_smoke.log
for f in cfg2 cfg2_reference cfg3 cfg4 cfg5 cfg5n10; do grep "^{" gpurun_out/re_bench_$f.json | cut -c1-200; done
