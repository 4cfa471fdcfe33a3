#!/bin/bash
# ncu --set full capture of a few launches of the named kernels in ONE chain of the D=4, N=3 side program (host-driven,
# plain launches), exported on the box as CSV pages small enough to travel back (the .ncu-rep itself is kept only if < 40 MB).
#   usage: tools/ncu_capture.sh <tag> <kernel-regex> <launch-skip> <count> [extra python args]
set -u
tag=$1; regex=$2; skip=$3; count=$4; shift 4
out=gpurun_out/${tag}
KBP_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${regex}" --launch-skip ${skip} -c ${count} \
  -o ${out} -f python tools/side_timing.py 4 3 --one-plain "$@" > ${out}.log 2>&1
ncu -i ${out}.ncu-rep --page raw --csv > ${out}_raw.csv 2>/dev/null
ncu -i ${out}.ncu-rep --page details --csv > ${out}_details.csv 2>/dev/null
ncu -i ${out}.ncu-rep --page source --csv --print-source sass > ${out}_source_sass.csv 2>/dev/null
sz=$(stat -c %s ${out}.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 40000000 ]; then rm -f ${out}.ncu-rep; fi
ls -la gpurun_out | grep ${tag}
