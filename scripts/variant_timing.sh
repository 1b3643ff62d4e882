#!/bin/bash
# stage timing of library variants built with make B=_v_<name> EXTRA=...
for v in "$@"; do
  echo -n "$v: "; GIBBS_B200_LIB=$PWD/gibbssampler_b200/csrc/$v/libgibbs_b200.so python scripts/stage_timing.py
done
