set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2961$n tools/bench_configs.py --config 4 > gpurun_out/r02c_config4_n${n}_batch.jsonl 2> gpurun_out/r02c_config4_n${n}_batch.err; echo "cfg4 batch n=$n rc=$?"
  $TR --nproc-per-node $n --master-port 2962$n tools/bench_configs.py --config 4 --clouds 1 > gpurun_out/r02c_config4_n${n}_points.jsonl 2> gpurun_out/r02c_config4_n${n}_points.err; echo "cfg4 points n=$n rc=$?"
done
python bench.py --no-cpu > gpurun_out/r02c_bench8box_n1.json 2> gpurun_out/r02c_bench8box_n1.err; echo "bench n1 rc=$?"
for n in 2 4 8; do
NCCL_DEBUG=VERSION $TR --nproc-per-node $n --master-port 2963$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > gpurun_out/r02c_bench8box_n$n.json 2> gpurun_out/r02c_bench8box_n$n.err; echo "bench n$n rc=$?"
done
for n in 1 2 4 8; do python -c "
import json; d=json.load(open('gpurun_out/r02c_bench8box_n$n.json')); print($n, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"; done
