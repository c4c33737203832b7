#!/bin/bash
mkdir -p gpurun_out
timeout 28 python -m pytest tests/test_ref_maps.py -m gpu -k refmesh -q > gpurun_out/pytest_refmesh.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_refmesh.txt
tail -5 gpurun_out/pytest_refmesh.txt
