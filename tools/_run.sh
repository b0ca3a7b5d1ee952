python -m pytest tests/test_dist_nccl.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s3_bench_2gpu.json 2> gpurun_out/s3_bench_2gpu.err; tail -c 600 gpurun_out/s3_bench_2gpu.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/s3_bench_2gpu.json').read().strip().splitlines() if l.startswith('{')][-1]);print(d['value'],d['ms_per_step'],d['e2e']['value'], d['e2e']['h2d_gbs_per_rank']); g=d['global_ba']; print(g['ms_per_iteration'], g['stage_ms_rank0'], g['collective_share_rank0'], g.get('solver'))"
