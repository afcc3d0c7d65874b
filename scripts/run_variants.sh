#!/bin/bash
# run the 8-view C2 sweep bench for every prebuilt library variant in _variants/ (experiment helper)
cp hc-mvs_b200/libhcmvs_b200.so /tmp/lib_orig.so
for f in _variants/lib_*.so; do
  n=$(basename $f .so)
  cp $f hc-mvs_b200/libhcmvs_b200.so
  python bench.py --views 8 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  echo "$n: $(python scripts/show_bench.py gpurun_out/var_$n.json | grep -E '^sweep')"
done
cp /tmp/lib_orig.so hc-mvs_b200/libhcmvs_b200.so
