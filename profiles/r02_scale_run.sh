mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo8.txt 2>&1
for n in 8 4 2; do
  extra="--no-sub"; [ $n = 8 ] && extra=""
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 200 --warmup 20 --e2e-steps 30 $extra > gpurun_out/r02_final_scale_n$n.json 2> gpurun_out/r02_final_scale_n$n.err
  echo "n=$n rc=$?"; tail -c 300 gpurun_out/r02_final_scale_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_final_scale_ref_n8.json 2> gpurun_out/r02_final_scale_ref_n8.err; echo "ref rc=$?"
