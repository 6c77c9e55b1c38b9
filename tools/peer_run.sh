# 2-GPU check of the multi-GPU paths: NCCL and peer-memory exchange tests, then nocomm / peer timing (gpurun --gpus 2 -- bash tools/peer_run.sh)
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_dist.py -q > gpurun_out/r2p_dist_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2p_dist_tests.log
for mode in nocomm peer_nccl_sync; do
  case $mode in
    nocomm) export JMAC_DIST_DIAG=nocomm JMAC_PEER_SYNC=flags;;
    peer_nccl_sync) export JMAC_DIST_DIAG= JMAC_PEER_SYNC=nccl;;
  esac
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --sections "" > gpurun_out/r2p_bench_n2_$mode.json 2> gpurun_out/r2p_bench_n2_$mode.err; echo "bench $mode rc=$?"
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2p_bench_n2_$mode.json
done
