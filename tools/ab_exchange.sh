out=${1:-gpurun_out/ab}; mkdir -p $out
for rep in 1 2; do for ex in auto p2p p2p-early p2p-fused; do
  timeout 90 python bench.py --exchange $ex --steps 1200 --no-cpu-baseline --no-extras --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2 exchange=$ex ms_per_step=%.4f launches/step=%d | ' % (d['ms_per_step'], d['gpu_launches']//d['steps']) + ' '.join('%s=%.2fus' % (k['part'], k['us']) for k in r['step_kernels']))"
done; done | tee $out/ab_exchange_loopback.log
