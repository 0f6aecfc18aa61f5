for br in ${BRS:-4000000 12500000 25000000 50000000}; do
echo "== batch_reads $br"
MCX_MULTI_BIN_FRAC=${FRAC:-0.08} MCX_MULTI_BATCH_READS=$br MCX_MULTI_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/exp_mb_$br.log 2>&1
grep -E "stage ms|value" gpurun_out/exp_mb_$br.log | cut -c1-330 || true
grep -E "Error|error|assert" gpurun_out/exp_mb_$br.log | head -5
done
