for so in odesat_b200/csrc/libodesat_b200.so lib_row64.so lib_row64_nohole.so lib_row128_nohole.so; do
for sched in exact balanced; do
 slack=1; case $so in *nohole*) slack=0;; esac
 echo "== $so slack=$slack $sched"
 ODESAT_SKIP_BUILD=1 ODESAT_B200_SO=$PWD/$so ODESAT_TILE_SLACK=$slack python bench.py --quick --steps 64 --warmup 3 --schedule $sched 2>&1 | tail -1 | cut -c1-140
done; done
