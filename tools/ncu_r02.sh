set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_s3.log 2>&1; echo rc=$? >> gpurun_out/gputest_s3.log; tail -3 gpurun_out/gputest_s3.log
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-tensor-probe --no-extras --sustained-s 0"
$CMD > gpurun_out/plain_r02.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:"k_share_ntt2|k_hash_records|k_eval|k_fs1|k_fs2|k_assemble|k_tails|k_expand|k_keygen" -s 90 -c 20 -o gpurun_out/ncu_r02 $CMD > gpurun_out/ncu_full.log 2>&1
echo full rc=$?
ls -la gpurun_out | tail
