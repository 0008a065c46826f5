run() { # name, env, extra args
  name=$1; shift; envs=$1; shift
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $NG --steps 20 --warmup 5 "$@" > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1]); print('$name', '%.4g steps/s' % d['value'], '%.1f us/step' % (d['ms_per_step']*1e3), 'gather_ok', d.get('gather_ok'), 'kernel %.1f us' % (d['roofline']['kernel_ms']*1e3))
except Exception as ex: print('$name FAILED', ex); print(open('gpurun_out/${TAG}_$name.err').read()[-1200:])
"
}
