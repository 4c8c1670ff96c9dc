B="timeout 300 python bench.py --quick --steps 64 --warmup 8 --workload rand50k --replicas 2048 --engine gather"
for sl in 128 256 512; do for st in 1 8; do
ODESAT_GATHER_SLAB=$sl ODESAT_GATHER_SLAB_STEPS=$st $B 2>&1 | tail -1 | cut -c1-140,200-330
done; done
ODESAT_GATHER_SLAB=0 $B --precision f64 2>&1 | tail -1 | cut -c1-140
