#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count> <profile_kernels.py args...>
# Captures `ncu --set full` for the matching launches and exports compact CSVs (raw metrics + per-source-line view) into
# gpurun_out/ so the report itself (often > 64 MiB with source import) does not have to travel back.
set -u
tag=$1; regex=$2; skip=$3; count=$4; shift 4
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c "$count" -f -o /tmp/ncu/$tag \
    python tools/profile_kernels.py "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/ncu/$tag.ncu-rep --page source --csv --print-source cuda 2>/dev/null | gzip -9 > gpurun_out/${tag}_source_cuda.csv.gz
ncu -i /tmp/ncu/$tag.ncu-rep --page details 2>/dev/null | gzip -9 > gpurun_out/${tag}_details.txt.gz
sz=$(stat -c %s /tmp/ncu/$tag.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -lt 20000000 ]; then cp /tmp/ncu/$tag.ncu-rep gpurun_out/; fi
echo "$tag: report $sz bytes"; ls -la gpurun_out | tail -8
