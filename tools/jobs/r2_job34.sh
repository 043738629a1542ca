#!/bin/bash
# round 2, GPU job 34: landslide marks with scrambled unit types (every branch of the "type does not grow" rule), K = 4 and 8
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_shia.py -x -q -p no:cacheprovider -k "scrambled or test_slides" ) > gpurun_out/j34_pytest.log 2>&1
( WMF_SHIA_K=8 timeout 900 python -m pytest tests/test_gpu_shia.py -x -q -p no:cacheprovider -k "scrambled" ) > gpurun_out/j34_pytest_k8.log 2>&1
