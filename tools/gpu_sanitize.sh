#!/bin/bash
# usage (under gpurun): tools/gpu_sanitize.sh -- compute-sanitizer memcheck over the cluster-path, restraint, parity, MD and
# edge tests, racecheck over the small cases; output in gpurun_out/r02_compute_sanitizer.txt
OUT=gpurun_out/r02_compute_sanitizer.txt
: > $OUT
run() { echo "### compute-sanitizer $*" | tee -a $OUT; timeout 1500 compute-sanitizer "$@" 2>&1 | grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|error|hazard" | head -40 | tee -a $OUT; }
run --tool memcheck python -m pytest tests/test_gpu_cluster.py tests/test_gpu_restraints.py -q -x -m gpu
run --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_md.py tests/test_gpu_reference_golden.py -q -x -m gpu -k "not 100k"
run --tool racecheck python -m pytest tests/test_gpu_restraints.py tests/test_gpu_parity.py -q -x -m gpu -k "cfg1 or restraint_energy"
run --tool memcheck python -m pytest tests/test_gpu_gb.py tests/test_gpu_opls.py tests/test_gpu_ewald.py tests/test_gpu_pme.py -q -x -m gpu
run --tool racecheck python -m pytest tests/test_gpu_gb.py -q -x -m gpu -k "cfg1_host_guest or own_charges"
# the displaced-atom rows kernels (block-per-row with 512 threads and the small persistent blocks): memcheck over the cluster
# tests with the small-block launch forced on, racecheck over the tiny cluster-path case in both launches
SDMB200_SIDE_SMALL=1 run --tool memcheck python -m pytest tests/test_gpu_cluster.py -q -x -m gpu
SDMB200_SIDE_SMALL=1 run --tool racecheck python -m pytest tests/test_gpu_edge.py -q -x -m gpu -k "tiny_ragged or three_displacement_groups"
run --tool racecheck python -m pytest tests/test_gpu_edge.py -q -x -m gpu -k "tiny_ragged"
