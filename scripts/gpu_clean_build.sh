#!/bin/bash
# Clean rebuild of every native library ON the GPU box (nvcc from the sources of the snapshot, nothing prebuilt is used),
# then smoke() and a short parity run against the fresh binaries: the log is the build-provenance record of the round.
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/r2w_clean_build.log
{
  echo "== removing prebuilt artefacts"; rm -rf jrl-qp_b200/_build oracle/_build; ls jrl-qp_b200/_build 2>&1 | head -2
  echo "== nvcc: $(nvcc --version | tail -2 | head -1)"; nvidia-smi --query-gpu=name,driver_version --format=csv,noheader
  echo "== build()"; time python -c "import __graft_entry__ as g; g.build()" 2>&1 | cut -c1-260
  echo "== freshly built:"; ls -la --time-style=full-iso jrl-qp_b200/_build/*.so oracle/_build/*.so
  sha256sum jrl-qp_b200/_build/libjrlqp_b200.so
  echo "== smoke()"; python -c "import __graft_entry__ as g; g.smoke()"
  echo "== parity tests against the fresh build"; python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -x -q 2>&1 | tail -3
} > $LOG 2>&1
tail -12 $LOG
