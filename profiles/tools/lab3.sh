#!/bin/bash
# build p1_lab3 for each RJ_LAB value given, run them on a B200 through gpurun, keep the output as gpurun_out/<tag>.txt
#   profiles/tools/lab3.sh <tag> <RJ_LAB value>...
set -e
tag=$1; shift
cd /root/repo/profiles/microbench
for v in "$@"; do
	nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -DRJ_LAB=$v -I../../include -I../../midoridb_b200/csrc -o p1_lab3_$v p1_lab3.cu 2>&1 | grep -E "error" || true
done
cd /root/repo
gpurun --timeout 600 -- "cd profiles/microbench && (for v in $*; do echo RJ_LAB=\$v; timeout 60 ./p1_lab3_\$v 28 | grep -E 'hints 0|warp|accounted'; done) 2>&1 | tee /root/repo/gpurun_out/$tag.txt" 2>&1 | grep -v "^\[gpurun\] s"
