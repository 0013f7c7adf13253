set -e
out=gpurun_out/r02_final_sweep.jsonl; : > $out
for b in balance box legacy_box test intrian hat humanb box4 leg2 leg balance2 balance3 insect quad; do
  python bench.py --body $b --steps 300 --warmup 30 --no-e2e --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'what':'K1 $b k_sub=1','us':round(d['roofline']['kernel_us'],1),'frac':round(d['roofline']['frac'],3),'env_steps_per_s':d['value'],'layout':d['config']['state_layout']}))" >> $out
done
python bench.py --config 4 --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-sub 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'what':'config 4 (quad, k_sub=8)','us':round(d['roofline']['kernel_us'],1),'frac':round(d['roofline']['frac'],3),'env_steps_per_s':d['value']}))" >> $out
for p in fused-fp32 fused-tf32 torch; do python bench.py --config 5 --policy $p --steps 640 --warmup 96 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'what':'config 5 policy=$p','us':round(d['ms_per_step']*1e3,1),'env_steps_per_s':d['value']}))" >> $out; done
for b in test balance1 balance2 balance3 leg2 box humanb insect; do python bench.py --config 6 --pkg-body $b --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'what':'package lineage $b, 1 update per launch','us':round(d['roofline']['kernel_us'],1),'frac':round(d['roofline']['frac'],3),'env_updates_per_s':d['value']}))" >> $out; done
python bench.py --probe-stream >> $out
cat $out
