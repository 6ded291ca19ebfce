# interleaved A/B of the whole training step on ONE box: $1 = environment variable toggled between 0 and 1
VAR=$1
for rep in 1 2; do
  for v in 0 1; do
    env $VAR=$v python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --no-cudnn 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('$VAR=$v rep $rep: %.2f ms/step  e2e %.1f  clock %s  fwd %.2f ms (%.0f TF/s)  wgrad %.2f ms (%.0f TF/s)' % (d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], k['conv3d_umma_fwd']['ms_per_step'], k['conv3d_umma_fwd']['tflops'], k['conv3d_umma_wgrad']['ms_per_step'], k['conv3d_umma_wgrad']['tflops']))"
  done
done
