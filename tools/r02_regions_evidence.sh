#!/bin/bash
# ncu evidence for the accumulator-region build: launch list of a cfg3pp rollout + full capture of one timestep (cfg3pp, cfg 3)
set -u
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_cfg3pp_b64.csv \
    python tools/run_once.py cfg3pp 64 1 > $O/s6_launches.log 2>&1
# second prediction step of cfg3pp: 4 cells x (C, M, O) + GHU + head = 14 conv_halo launches per step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel --launch-skip 28 --launch-count 14 \
    -f -o $O/r02_ncu_cfg3pp_step_b256 python tools/run_once.py cfg3pp 256 1 > $O/s6_ncu.log 2>&1
# cfg 3 (ST-LSTM): 3 cells x (C, M, O, adapter) + head = 13 per step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel --launch-skip 26 --launch-count 13 \
    -f -o $O/r02_ncu_cfg3_step_b256 python tools/run_once.py cfg3 256 1 > $O/s6_ncu3.log 2>&1
timeout 200 python tools/layer_profile.py cfg3 256 > $O/s6_layers_cfg3.txt 2>&1
timeout 200 python tools/layer_profile.py cfg3pp 256 > $O/s6_layers_cfg3pp.txt 2>&1
