for V in 0 1 2; do for K in 12 11 10; do for L in 3 4; do
echo "V=$V K=$K L=$L"; QCSIM_TILE_VARIANT=$V QCSIM_TILE_BITS=$K QCSIM_TILE_LOW=$L python bench.py --qubits 28 --steps 10 --no-cpu-baseline --no-kernel-sweep 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(' value',d['value'],'ms/step',d['ms_per_step'],'launches',d['gpu_launches'],'passes/step',d['roofline']['passes_per_step'],'GB/s',d['roofline']['achieved'])
"
done; done; done
