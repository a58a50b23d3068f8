#!/bin/bash
# tools/ab_flags.sh "<flags A>" "<flags B>" ... : rebuild with each EXTRA flag set on the GPU box and print the per-family profile
set -u
for f in "$@"; do
  make -C r3m_b200/csrc clean >/dev/null; make -C r3m_b200/csrc -j16 EXTRA="$f" >/dev/null 2>&1
  echo "EXTRA=$f"; python tools/profile_step.py 2>&1 | tail -1
done
