#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a small -m gpu subset that covers K1, the K2 sample / main passes
# (1-CTA and CTA-pair variants), the threshold kernels, K3 in its three widths with the fused tail, K4, K5 and the
# peer kernels with fake ranks.  Summaries -> gpurun_out/<tag>_sanitizer_*.log
set -u
mkdir -p gpurun_out
T=${1:-r2s}
SEL="test_l2norm_matches_f_normalize or test_simtopk_accumulator_tiles_match_fp32_matmul or test_tiny_golden or test_session_fused_vote_equals_oracle_vote or test_peer_exchange_with_fake_ranks_on_one_gpu or test_peer_prediction_payload or test_k3_never_certifies or test_vote_kernel_against_oracle or test_merge_with_short_shards or test_certification_fallback"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "$SEL" > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard|Invalid|error" gpurun_out/${T}_sanitizer_$tool.log | sort | uniq -c | sort -rn | head -12
done
