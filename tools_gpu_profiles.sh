#!/bin/bash
# round-end evidence: (a) launch list of one step with device time + DRAM bytes per launch, (b) ncu --set full on the hot kernels,
# (c) default bench (with CPU baseline), (d) reference arm.  Results land in gpurun_out/ (<= 64 MiB: the big .ncu-rep is exported to CSV on the
# box and dropped); tools/make_profiles.py turns them into profiles/.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/step_once.py > gpurun_out/launchlist.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"attn_fwd_kernel|attn_bwd|gemm_tcgen05|talking_fwd|talking_bwd_rows|th8_fwd|th8_bwd|layernorm" -s 17 -c 19 -o gpurun_out/prof_hot -f python tools/prof_attn.py 2 > gpurun_out/ncu_hot.log 2>&1
ncu -i gpurun_out/prof_hot.ncu-rep --page raw --csv > gpurun_out/prof_hot_raw.csv 2>/dev/null
rm -f gpurun_out/prof_hot.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_kernel" -s 2 -c 1 -o gpurun_out/prof_attn_fused -f python tools/prof_attn.py 2 > gpurun_out/ncu_attn_fused.log 2>&1
# round 2 kernels: fused talking-heads (tcgen05) at cfg2, H=16 mma.sync kernels at cfg4 (make_profiles.py keeps the LAST launch of every kernel instantiation)
timeout 900 ncu --set full --clock-control none -k regex:"tf_fwd_kernel|tf_bwd_kernel|tf_merge|th16_fwd|th16_bwd" -c 48 -o gpurun_out/prof_fused -f python tools/prof_fused.py 2 > gpurun_out/ncu_fused.log 2>&1
ncu -i gpurun_out/prof_fused.ncu-rep --page raw --csv > gpurun_out/prof_fused_raw.csv 2>/dev/null
rm -f gpurun_out/prof_fused.ncu-rep
for nt in 256 384 512; do SPE_TH16_BWD_THREADS=$nt timeout 300 python tools/dev/th16_check.py 2>&1 | grep "N=4150" | sed "s/^/bwd_threads=$nt /" >> gpurun_out/th16_ab.log; done
timeout 900 python bench.py --config cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/th16_ab.log; tail -c 300 gpurun_out/ncu_fused.log; tail -c 600 gpurun_out/bench_default.json; echo; cat gpurun_out/bench_reference.json | cut -c1-300; wc -l gpurun_out/launches.csv gpurun_out/prof_hot_raw.csv; du -sh gpurun_out; tail -2 gpurun_out/ncu_hot.log
