#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/step_once.py > gpurun_out/launchlist.log 2>&1
tail -2 gpurun_out/launchlist.log; wc -l gpurun_out/launches.csv
