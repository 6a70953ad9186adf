python -m pytest tests/test_gpu_run_batch.py -m gpu -q -x 2>&1 | tail -2
F="--no-cpu-baseline --no-e2e --no-search --no-traj --steps 5"
for pr in f64 f32; do for n in 131072 262144 32768; do
python bench.py $F --precision $pr --filters $n --tsteps 10000 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); s=d['shared_covariance']; print('$pr $n', '%.4g'%s['value'], s.get('frac_of_hbm_peak'))"
done; done
