#!/bin/bash
# Closing GPU session of the second sitting of round 2 (one B200), final build: GPU tests, smoke, bench lines of every config +
# the reference arm, launch list of the default bench command, K4 ncu capture, family probe, memcheck
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/rz3_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rz3_smoke.log 2>&1
python bench.py > gpurun_out/rz3_bench_cfg2.json 2> gpurun_out/rz3_bench_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/rz3_bench_cfg2_reference.json 2>/dev/null
for w in cfg3 cfg4 cfg5n10; do python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/rz3_bench_$w.json 2> gpurun_out/rz3_bench_$w.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r3_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extra none > gpurun_out/rz3_ncu_bench.log 2>&1
PROF_B1=8192 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_simulate -c 1 -s 1 -o gpurun_out/k4_r3z -f python profiles/prof_driver.py > gpurun_out/rz3_ncu_k4.log 2>&1
python profiles/ncu_extract.py gpurun_out/k4_r3z.ncu-rep > gpurun_out/k4_r3z_ncu_summary.txt 2>/dev/null
SAN_FAMILIES=iso,thermal,sei,wsei,wthsei,thsp,iso12 compute-sanitizer --tool memcheck python profiles/sanitize_driver.py 2>&1 | tail -12 > gpurun_out/rz3_memcheck.txt
for f in iso thermal sei wide wsei wth thsei wthsei mhc lgm lgmth iso_r12 iso_r14 iso_sp thermal_r12 thermal_r14 thermal_sp sei_r12 sei_r14 sei_sp; do python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done > gpurun_out/rz3_families.txt
cat gpurun_out/rz3_pytest.log; cut -c1-150 gpurun_out/rz3_families.txt; tail -1 gpurun_out/rz3_smoke.log
for f in cfg2 cfg2_reference cfg3 cfg4 cfg5n10; do grep "^{" gpurun_out/rz3_bench_$f.json | cut -c1-200; done
grep -E "ERROR SUMMARY" gpurun_out/rz3_memcheck.txt
