#!/usr/bin/env bash
# Compiles the reference's OWN CPU implementation of the RoPE path -- rope_2d_cpu / rope_2d in
# /root/reference/src/model/encoder/backbone/croco/curope/curope.cpp -- from the sources where they lie,
# into oracle/_ref/curope_ref*.so (git-ignored, travels to the GPU box).  Test infrastructure only.
# The reference's CUDA translation unit (kernels.cu) does not compile against torch 2.11
# (kernels.cu:101 uses tokens.type()), so rope_2d_cuda is satisfied by the throwing stub next to this script.
# The rasterizer (diff_gauss_pose) is a third-party package absent from /root/reference: unbuildable here.
set -euo pipefail
cd "$(dirname "$0")"
REF=/root/reference/src/model/encoder/backbone/croco/curope/curope.cpp
[ -f "$REF" ] || { echo "reference sources not present; keeping any prebuilt oracle/_ref"; exit 0; }
mkdir -p _ref
PY=${PYTHON:-python}
read -r TORCH_INC TORCH_LIB PY_INC EXT <<<"$($PY - <<'PYEOF'
import sysconfig, os, torch
from torch.utils import cpp_extension as ce
inc = " ".join("-I" + p for p in ce.include_paths())
print(inc.replace(" ", ";"), os.path.join(os.path.dirname(torch.__file__), "lib"), sysconfig.get_paths()["include"],
      sysconfig.get_config_var("EXT_SUFFIX"))
PYEOF
)"
g++ -O2 -std=c++17 -fPIC -shared -DTORCH_EXTENSION_NAME=curope_ref -DTORCH_API_INCLUDE_EXTENSION_H \
    ${TORCH_INC//;/ } -I"$PY_INC" "$REF" rope_cuda_stub.cpp \
    -L"$TORCH_LIB" -Wl,-rpath,"$TORCH_LIB" -ltorch -ltorch_cpu -ltorch_python -lc10 \
    -o "_ref/curope_ref$EXT"
echo "built oracle/_ref/curope_ref$EXT"
