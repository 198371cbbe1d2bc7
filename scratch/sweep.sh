for L in 3 4 5 6 7 8; do
echo "L=$L"; QCSIM_TILE_LOW=$L QCSIM_TILE_VARIANT=1 python bench.py --qubits 28 --steps 10 --no-cpu-baseline --no-kernel-sweep 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' value',d['value'],'ms/step',d['ms_per_step'],'launches',d['gpu_launches'],'passes/step',d['roofline']['passes_per_step'],'GB/s',d['roofline']['achieved'], 'e2e', d['e2e']['value'])
"
done
