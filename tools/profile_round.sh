#!/bin/bash
# One GPU call that produces every ncu artefact of a round under gpurun_out/ (run through gpurun; summaries are made on the
# build box by tools/make_profiles.py):  tools/profile_round.sh <tag>
tag=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_forward.py > gpurun_out/${tag}_launches.log 2>&1
# (kernel, launches to skip inside the profiled forward of 4 iterations): loop kernels skip the first iterations, once-per-forward kernels none
for ks in k_corr_lookup:2 k_tc_linear:20 k_setconv_edge_pairs:6 k_knn_branch:2 k_corr_topk_vec:0 k_corr_gemm:0 k_knn_grid:0; do
  k=${ks%%:*}; s=${ks##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 --profile-from-start off -o gpurun_out/${tag}_$k python tools/profile_forward.py --iters 4 --warm 1 > gpurun_out/${tag}_ncu_$k.log 2>&1
done
ls -la gpurun_out/${tag}_*.ncu-rep
