#!/bin/bash
# Lean evidence pass (GPU-minute budget): bench (both arms), decode-step launch list, one full ncu capture of the
# decode-step kernels.  Usage (via gpurun): bash tools/gpu_final.sh <tag>
tag=${1:-r02_final}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 python bench.py --profile-out gpurun_out/event_profile_$tag.json > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc $?"; cut -c1-400 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.json 2>/dev/null; echo "ref rc $?"
# launch list of five decode steps (warm L2, every kernel of the chain), summarised per kernel
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 2200 -c 360 --csv \
  --log-file gpurun_out/declist_$tag.csv \
  python bench.py --steps 1 --warmup 0 --tokens 40 --no-e2e --no-cpu-baseline --no-gpu-eager --no-profile --no-graph > gpurun_out/declist_$tag.log 2>&1
echo "declist rc $?"
python tools/ncu_summarize.py gpurun_out/declist_$tag.csv > gpurun_out/declist_$tag.md; cat gpurun_out/declist_$tag.md
# full capture of the decode-step kernels
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'decode_cross_persist|decode_self_persist|gemm_tc_skinny|splitk_ln|sample_step|embed_ln' \
  -s 300 -c 40 -f -o gpurun_out/full_dec_$tag python bench.py --steps 1 --warmup 0 --tokens 12 --no-e2e --no-cpu-baseline --no-gpu-eager --no-profile --no-graph > gpurun_out/full_dec_$tag.log 2>&1
echo "full rc $?"
ncu -i gpurun_out/full_dec_$tag.ncu-rep --page raw --csv > gpurun_out/full_dec_${tag}_raw.csv 2>/dev/null
python tools/ncu_traffic.py gpurun_out/full_dec_${tag}_raw.csv decode_cross_persist 100 bf16 gpurun_out/roofline_traffic.json > gpurun_out/full_dec_$tag.md
head -45 gpurun_out/full_dec_$tag.md
rm -f gpurun_out/*.ncu-rep
# encoder GEMM shapes: ncu duration + tensor-pipe activity per launch (tools/bench_gemm.py, one launch per shape)
bash tools/gpu_gemm_ncu.sh $tag > gpurun_out/gemm_tensor_$tag.txt 2>&1; tail -40 gpurun_out/gemm_tensor_$tag.txt
# launch list of one encoder pass: per-kernel time, DRAM bytes and tensor-pipe activity
bash tools/gpu_enc_ncu.sh $tag > gpurun_out/enclist_$tag.txt 2>&1; tail -30 gpurun_out/enclist_$tag.txt
