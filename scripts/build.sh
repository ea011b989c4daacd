#!/bin/bash
# Build liblagomorph_b200.so without importing the package (the package import needs the library).
cd $(dirname $0)/.. && python - "$@" <<'PY'
import importlib.util, sys
spec = importlib.util.spec_from_file_location("_b", "lagomorph_b200/build.py")
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
print(m.build(force="--force" in sys.argv, verbose=True))
PY
