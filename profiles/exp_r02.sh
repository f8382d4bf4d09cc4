#!/bin/bash
# round-2 tuning pass: kernel variants + Python-call profile.  Usage: gpurun -- 'bash profiles/exp_r02.sh tag'
tag=${1:-x}; out=gpurun_out/$tag; mkdir -p $out
python profiles/exp_variants.py base > $out/variants.txt 2>&1
for v in w20 w24; do
  CLEDB_B200_LIB=$PWD/solar-coronal-inversion_b200/lib/variants/libcledb_b200_$v.so python profiles/exp_variants.py $v >> $out/variants.txt 2>&1
done
cat $out/variants.txt
python profiles/time_invproc.py 256 > $out/time_invproc.txt 2>&1; head -40 $out/time_invproc.txt
