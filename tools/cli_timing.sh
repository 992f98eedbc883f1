#!/bin/bash
# wall clock of the C++ drop-in (upcgen) on BASELINE config 2 with 10^6 events: HepMC output, ROOT output, and none
set -e
make -s -C upcgen_b200/host
W=$(mktemp -d)
for MODE in "0 0" "1 0" "0 1"; do
  set -- $MODE
  D=$W/h$1r$2; mkdir -p $D
  python - "$D" "$1" "$2" <<'PY'
import sys
sys.path.insert(0, ".")
from upcgen_b200.config import config_text
d, h, r = sys.argv[1:4]
open(d + "/parameters.in", "w").write(config_text("cfg2", f"NEVENTS 1000000\nUSE_HEPMC_OUTPUT {h}\nUSE_ROOT_OUTPUT {r}\n"))
PY
  ( cd $D; s=$(date +%s%N); $OLDPWD/upcgen_b200/host/upcgen -debug 0 > out.log 2> err.log; e=$(date +%s%N);
    echo "HEPMC=$1 ROOT=$2: $(( (e - s) / 1000000 )) ms; $(grep -h 'cross section' out.log | head -1); $(ls -la events.* 2>/dev/null | awk '{print $5, $9}' | tr '\n' ' ')" )
done
