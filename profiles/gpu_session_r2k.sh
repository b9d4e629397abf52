#!/bin/bash
cd "$(dirname "$0")/.."
for fam in iso sei thermal wsei; do
  B=65536; [ $fam = thermal ] && B=32768
  python profiles/k1_probe.py $B $fam 2>&1 | tail -1
  PLB_LIB=$PWD/profiles/variants/libplb_r1.so python profiles/k1_probe.py $B $fam 2>&1 | tail -1
done
for fam in thermal wsei sei; do
PROF_FAMILY=$fam PROF_B1=32768 PROF_B4=1024 ncu --set full --clock-control none --import-source on -k regex:k_resjac -c 1 -s 1 -o gpurun_out/k1_${fam}_r2a -f python profiles/prof_driver.py > gpurun_out/r2n_ncu_$fam.log 2>&1
python profiles/ncu_extract.py gpurun_out/k1_${fam}_r2a.ncu-rep > gpurun_out/k1_${fam}_r2a_ncu_summary.txt 2>/dev/null
done
grep -h "duration\|registers_per\|shared_mem_per\|warps_active\|issue_active\|inst_executed.sum" gpurun_out/k1_*_r2a_ncu_summary.txt
