#!/bin/bash
# Remote GPU session with per-step timeouts; stops early when the kernel sanity step fails (a hang must not eat the budget).
# usage: tools/gpu_session.sh <tag> [steps...]   steps ⊂ {dbg micro ops golden ncu bench bench_bf16 launches ncu_hbm}
TAG=$1; shift
STEPS=${@:-dbg micro ops golden}
O=gpurun_out
mkdir -p $O
for S in $STEPS; do
  echo "=== $S"
  case $S in
    dbg)
      timeout 150 python tools/debug_planes.py > $O/${TAG}_dbg.log 2>&1; rc=$?
      grep -c "equal: True" $O/${TAG}_dbg.log; grep "False\|Error\|error" $O/${TAG}_dbg.log | head -5
      if [ $rc -ne 0 ] || grep -q "False" $O/${TAG}_dbg.log; then echo "debug step failed/hung/mismatch (rc=$rc): stopping"; tail -5 $O/${TAG}_dbg.log; exit 1; fi ;;
    trap)
      PASCO_NVCC_FLAGS=-DPASCO_HANG_TRAP python -m pasco_b200.build --force > $O/${TAG}_trapbuild.log 2>&1
      timeout 200 python tools/debug_planes.py wgrad > $O/${TAG}_trap.log 2>&1
      grep "HANG" $O/${TAG}_trap.log | sort | uniq -c | sort -rn | head -40; tail -5 $O/${TAG}_trap.log | cut -c1-300
      python -m pasco_b200.build --force > /dev/null 2>&1 ;;
    profile)
      timeout 300 python tools/profile_step.py --out $O/${TAG}_profile.txt 2>&1 | tail -2; head -45 $O/${TAG}_profile.txt | cut -c1-200 ;;
    micro)
      timeout 200 python tools/conv_microbench.py --occ 0.5 0.1 --channels 64 128 256 --out $O/${TAG}_micro.jsonl 2>&1 | grep conv3 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if "precision" in d: print(d["N"], d["C"], d["precision"], {k: round(v, 3) for k, v in d.items() if k.endswith('_ms')})" ;;
    ops)
      timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -m gpu > $O/${TAG}_ops.log 2>&1; tail -4 $O/${TAG}_ops.log ;;
    golden)
      timeout 400 python -m pytest tests/test_golden_r2.py tests/test_golden_net.py tests/test_dropin_reference.py tests/test_ensemble.py -q -m gpu --tb=short -s > $O/${TAG}_golden.log 2>&1
      grep -v "^Scale\|Overall keep\|dropout_type" $O/${TAG}_golden.log | grep "Error\|passed\|failed" | cut -c1-500 ;;
    ncu)
      timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_conv_pl -c 1 -o $O/${TAG}_conv_pl python tools/conv_microbench.py --occ 0.5 --channels 64 --out /tmp/x.jsonl > /dev/null 2>&1
      timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_pl -c 1 -o $O/${TAG}_wgrad_pl python tools/conv_microbench.py --occ 0.5 --channels 64 --out /tmp/x.jsonl > /dev/null 2>&1
      ls -la $O/${TAG}_*pl.ncu-rep ;;
    bench)
      timeout 500 python bench.py --steps 8 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -3 $O/${TAG}_bench.err; cut -c1-1500 $O/${TAG}_bench.json ;;
    bench_bf16)
      timeout 400 python bench.py --steps 8 --warmup 3 --precision bf16 --no-cpu-baseline > $O/${TAG}_bench_bf16.json 2> $O/${TAG}_bench_bf16.err; cut -c1-600 $O/${TAG}_bench_bf16.json ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 14000 -c 6000 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1; wc -l $O/${TAG}_launches.csv ;;
    ncu_hbm)
      # HBM-bound kernels of one bench step: slim section set, summarised ON THE BOX (the .ncu-rep stays in /tmp: gpurun merges
      # at most 64 MiB back)
      timeout 500 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section WarpStateStats --section Occupancy \
        --clock-control none -k regex:'k_col_sums|k_bn_bwd|k_affine_act|k_split_planes|k_gather_rows|k_scatter_rows|k_kernel_map|k_hash|k_mask|k_xattn|k_maxpool|k_scatter_max|k_dense|k_pack|k_bn_finalize' \
        -s 1300 -c 45 -o /tmp/${TAG}_hbm python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_hbm.log 2>&1
      python tools/ncu_summarise.py /tmp/${TAG}_hbm.ncu-rep --out $O/${TAG}_ncu_hbm.json --note "one bench step (configs[1] shape, fp32 mode), 45 launches of the non-convolution kernels" | head -30 ;;
  esac
done
