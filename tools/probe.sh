# diagnostics: component costs of the fused kernel via LZGPU_PROBE bits (results invalid when bits are set)
mkdir -p gpurun_out
T=${TILE:-512}
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --tile-chunks $T --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1),'GiB/s  ms',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    elif 'Error' in l or 'error' in l: print(l.strip()[:200])
"; }
for P in "$@"; do run LZGPU_PROBE=$P; done
