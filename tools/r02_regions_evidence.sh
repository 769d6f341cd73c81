#!/bin/bash
# ncu evidence for the accumulator-region build: launch list of a cfg3pp rollout + full capture of the launches of one
# timestep (cell 0: C, M, O; GHU; cell 1: C, M, O), summarised ON the box (the .ncu-rep stays in /tmp: 64 MiB pull limit)
set -u
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_cfg3pp_b64.csv \
    python tools/run_once.py cfg3pp 64 1 > $O/s6_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel --launch-skip 28 --launch-count 7 \
    -f -o /tmp/r02_ncu_cfg3pp_step_b256 python tools/run_once.py cfg3pp 256 1 > $O/s6_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02_ncu_cfg3pp_step_b256.ncu-rep "cfg3pp (PredRNN++: Causal LSTM x4 + GHU, 256 sequences), accumulator-region build: cell 0 (C, M, O), GHU, cell 1 (C, M, O) of one timestep, ncu --set full" > $O/r02_ncu_cfg3pp_step_b256.md 2> $O/s6_ncu_summary.err
timeout 200 python tools/layer_profile.py cfg3 256 > $O/s6_layers_cfg3.txt 2>&1
timeout 200 python tools/layer_profile.py cfg3pp 256 > $O/s6_layers_cfg3pp.txt 2>&1
