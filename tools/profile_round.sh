#!/bin/bash
# One GPU call that produces every ncu artefact of a round under gpurun_out/ (run through gpurun; summaries are made on the
# build box by tools/make_profiles.py):  tools/profile_round.sh <tag>
tag=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_forward.py > gpurun_out/${tag}_launches.log 2>&1
for k in k_corr_lookup k_tc_linear k_setconv_edge_pairs k_corr_topk_vec k_knn_branch k_corr_gemm k_knn_grid; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 --profile-from-start off -o gpurun_out/${tag}_$k python tools/profile_forward.py --iters 2 --warm 1 > gpurun_out/${tag}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${tag}_*
