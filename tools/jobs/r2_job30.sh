#!/bin/bash
# round 2, GPU job 30: hydrology capped at two CTAs per SM (unused dynamic shared memory) so that k_sed_tb shares the SMs
mkdir -p gpurun_out
for pad in 0 100; do echo "pad $pad KB"; WMF_SED_HPAD_KB=$pad timeout 600 python tools/bench_scenarios.py --member-axis 1,4,8 2> gpurun_out/j30_scen_p$pad.err | cut -c150-300; done > gpurun_out/j30_scen.log
