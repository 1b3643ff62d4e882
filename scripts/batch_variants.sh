for v in "$@"; do echo "== $v"; GIBBS_B200_LIB=$PWD/gibbssampler_b200/csrc/$v/libgibbs_b200.so python scripts/batch_timing.py 2>&1 | grep "chains/launch 2"; done
