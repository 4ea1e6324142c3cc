set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -8
# DP correctness on 8 ranks (real model over NCCL vs the single-process full batch)
$TR --nproc-per-node 8 --master-port 29601 tools/dp_proof.py --prec fp32 --out gpurun_out/r02_dp_proof_n8_fp32.json > gpurun_out/r02_dp_proof_n8.log 2>&1; echo "dp_proof rc=$?"
$TR --nproc-per-node 8 --master-port 29602 tools/dp_proof.py --prec bf16 --out gpurun_out/r02_dp_proof_n8_bf16.json >> gpurun_out/r02_dp_proof_n8.log 2>&1; echo "dp_proof bf16 rc=$?"
# config 4: encoder-only sweep, batch-sharded (8 clouds) and point-sharded (1 cloud), 1/2/4/8 GPUs
python tools/bench_configs.py --config 4 > gpurun_out/r02_config4_n1.jsonl 2> gpurun_out/r02_config4_n1.err; echo "cfg4 n1 rc=$?"
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2961$n tools/bench_configs.py --config 4 > gpurun_out/r02_config4_n${n}_batch.jsonl 2> gpurun_out/r02_config4_n${n}_batch.err; echo "cfg4 batch n=$n rc=$?"
  $TR --nproc-per-node $n --master-port 2962$n tools/bench_configs.py --config 4 --clouds 1 > gpurun_out/r02_config4_n${n}_points.jsonl 2> gpurun_out/r02_config4_n${n}_points.err; echo "cfg4 points n=$n rc=$?"
done
python tools/bench_configs.py --config 4 --clouds 1 > gpurun_out/r02_config4_n1_1cloud.jsonl 2>> gpurun_out/r02_config4_n1.err
# headline at 1 and 8 GPUs on this box
python bench.py --no-cpu > gpurun_out/r02_bench8box_n1.json 2> gpurun_out/r02_bench8box_n1.err; echo "bench n1 rc=$?"
NCCL_DEBUG=INFO $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench8box_n8.json 2> gpurun_out/r02_bench8box_n8.err; echo "bench n8 rc=$?"
cut -c1-300 gpurun_out/r02_bench8box_n1.json; cut -c1-300 gpurun_out/r02_bench8box_n8.json
cat gpurun_out/r02_config4_n8_batch.jsonl | cut -c1-300
grep -c "NVLS\|nranks 8" gpurun_out/r02_bench8box_n8.err
