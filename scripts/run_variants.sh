#!/bin/bash
# run the 8-view C2 sweep bench (and the NCC parity tests) for every prebuilt library variant in _variants/ (experiment helper)
cp hc-mvs_b200/libhcmvs_b200.so /tmp/lib_orig.so
echo "default: $(bash scripts/bench_quick.sh | cut -c1-120)"
for f in _variants/lib_*.so; do
  n=$(basename $f .so)
  cp $f hc-mvs_b200/libhcmvs_b200.so
  echo "$n: $(bash scripts/bench_quick.sh | cut -c1-120)"
  python -m pytest tests/test_gpu_parity.py -q -x -m gpu 2>&1 | tail -1
done
cp /tmp/lib_orig.so hc-mvs_b200/libhcmvs_b200.so
