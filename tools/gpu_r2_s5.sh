#!/bin/bash
mkdir -p gpurun_out/r2s5; O=gpurun_out/r2s5
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt
tail -25 $O/pytest_gpu.txt
python examples/sky_localisation.py > $O/sky.txt 2>&1; tail -3 $O/sky.txt
