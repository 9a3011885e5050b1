# round-2 ncu evidence (run under gpurun on one B200): launch list and one --set full capture of a prove step; only CSV comes back
# (the .ncu-rep of 16 kernels with source is > 64 MiB).  Usage: bash tools/ncu_r02.sh [tag]
TAG=${1:-r02}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-tensor-probe --no-extras --sustained-s 0"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_share_ntt2|k_hash_records|k_eval|k_fs1|k_fs2|k_assemble|k_tails|k_expand|k_keygen|k_ntt_f|k_open|k_derive" -s 112 -c 16 -o /tmp/ncu_$TAG $CMD > gpurun_out/ncu_full.log 2>&1
echo full rc=$?
ncu -i /tmp/ncu_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2> gpurun_out/ncu_export.log
ncu -i /tmp/ncu_$TAG.ncu-rep --page source --csv -k regex:k_share_ntt2 > gpurun_out/ncu_${TAG}_share_source.csv 2>> gpurun_out/ncu_export.log
ncu -i /tmp/ncu_$TAG.ncu-rep --page details --csv > gpurun_out/ncu_${TAG}_details.csv 2>> gpurun_out/ncu_export.log
ls -la /tmp/ncu_$TAG.ncu-rep gpurun_out | tail -12
