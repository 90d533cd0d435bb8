#!/bin/bash
# One GPU-box pass for a round: tests, smoke, bench (both arms), launch list, one full ncu capture.
# Usage (via gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
bash tools/gpu_checks.sh
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/smoke_$tag.log
timeout 600 python bench.py --profile-out gpurun_out/event_profile_$tag.json > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc $?"; cut -c1-1500 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$tag.json 2>/dev/null; echo "ref rc $?"
# launch list of one step with 32 of the 255 decode steps (graph off so every kernel is listed; ncu serialises and
# replays each launch, a full 255-step list takes > 15 min): encode, cross-K/V, prefill, 31 decode steps, reward
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv \
  --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 1 --warmup 0 --tokens 32 --no-e2e --no-cpu-baseline --no-profile --no-graph > gpurun_out/ncu_list_$tag.log 2>&1
echo "launch list rc $?"
python tools/ncu_summarize.py gpurun_out/launches_$tag.csv > gpurun_out/launches_$tag.md
gzip -f gpurun_out/launches_$tag.csv
head -30 gpurun_out/launches_$tag.md
# one full capture: the decode-step kernels (skip the prefill/first steps)
bash tools/gpu_ncu_full.sh dec_$tag 'decode_cross_persist|decode_self_persist|gemm_tc_skinny|splitk_ln|sample_step|embed_ln' 300 72 12
python tools/ncu_traffic.py gpurun_out/full_dec_${tag}_raw.csv decode_cross_persist 89 bf16 gpurun_out/roofline_traffic.json > gpurun_out/full_dec_$tag.md
cat gpurun_out/full_dec_$tag.md | head -50
