#!/bin/bash
# Closing GPU session of round 2 (one B200): GPU tests, smoke, bench lines of every config + the reference arm, launch list
# of the default bench command, ncu captures of K1 (all families) and K4, sanitizer.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/rz_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rz_smoke.log 2>&1
python bench.py > gpurun_out/rz_bench_cfg2.json 2> gpurun_out/rz_bench_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/rz_bench_cfg2_reference.json 2>/dev/null
for w in cfg3 cfg4 cfg5n10; do python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/rz_bench_$w.json 2> gpurun_out/rz_bench_$w.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --extra none > gpurun_out/rz_ncu_bench.log 2>&1
PROF_B1=65536 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_resjac -c 1 -s 1 -o gpurun_out/k1_r2z -f python profiles/prof_driver.py > gpurun_out/rz_ncu_k1.log 2>&1
PROF_B1=8192 PROF_B4=8192 ncu --set full --clock-control none --import-source on -k regex:k_simulate -c 1 -s 1 -o gpurun_out/k4_r2z -f python profiles/prof_driver.py > gpurun_out/rz_ncu_k4.log 2>&1
for fam in thermal sei wsei; do
  PROF_FAMILY=$fam PROF_B1=32768 PROF_B4=1024 ncu --set full --clock-control none --import-source on -k regex:k_resjac -c 1 -s 1 -o gpurun_out/k1_${fam}_r2z -f python profiles/prof_driver.py > gpurun_out/rz_ncu_k1_$fam.log 2>&1
done
for k in k1_r2z k4_r2z k1_thermal_r2z k1_sei_r2z k1_wsei_r2z; do python profiles/ncu_extract.py gpurun_out/$k.ncu-rep > gpurun_out/${k}_ncu_summary.txt 2>/dev/null; done
compute-sanitizer --tool memcheck python profiles/sanitize_driver.py 2>&1 | tail -14 > gpurun_out/rz_memcheck.txt
SAN_FAMILIES=iso compute-sanitizer --tool racecheck python profiles/sanitize_driver.py > gpurun_out/rz_racecheck_full.txt 2>&1
SAN_FAMILIES=thsei compute-sanitizer --tool racecheck python profiles/sanitize_driver.py 2>&1 | tail -3 > gpurun_out/rz_racecheck_thsei.txt
SAN_FAMILIES=wide compute-sanitizer --tool racecheck python profiles/sanitize_driver.py > gpurun_out/rz_racecheck_wide_full.txt 2>&1
SAN_FAMILIES=iso compute-sanitizer --tool synccheck python profiles/sanitize_driver.py 2>&1 | tail -4 > gpurun_out/rz_synccheck.txt
for f in iso thermal sei wide wsei wth thsei wthsei mhc lgm lgmth; do python profiles/k4_probe.py 16384 $f 2>&1 | tail -1; done > gpurun_out/rz_families.txt
cat gpurun_out/rz_pytest.log; cat gpurun_out/rz_families.txt; tail -2 gpurun_out/rz_smoke.log
for f in cfg2 cfg2_reference cfg3 cfg4 cfg5n10; do grep "^{" gpurun_out/rz_bench_$f.json | cut -c1-160; done
grep -E "RACECHECK SUMMARY|ERROR SUMMARY" gpurun_out/rz_*check*.txt
