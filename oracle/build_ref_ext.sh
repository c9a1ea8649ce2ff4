#!/usr/bin/env bash
# TEST INFRASTRUCTURE (oracle). Builds the UNMODIFIED reference pointnet2 CUDA extension
# (/root/reference/model/pointnet2/_ext_src, bindings.cpp:11-24) for sm_100a into oracle/_ref/
# so that GPU parity tests can compare the B200 kernels with the reference's own kernels.
# Sources are compiled where they lie; nothing is copied into the repo. Outputs only in oracle/_ref/.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/model/pointnet2/_ext_src" ] || { echo "reference not present; skipping"; exit 0; }
mkdir -p "$OUT/pointnet2_ref" "$OUT/build"
python - "$REF" "$OUT" <<'PY'
import sys, glob, os, shutil
ref, out = sys.argv[1], sys.argv[2]
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
os.environ.setdefault("MAX_JOBS", "8")
from torch.utils.cpp_extension import load
src = sorted(glob.glob(f"{ref}/model/pointnet2/_ext_src/src/*.cpp") + glob.glob(f"{ref}/model/pointnet2/_ext_src/src/*.cu"))
inc = f"{ref}/model/pointnet2/_ext_src/include"
m = load(name="_ext", sources=src, extra_include_paths=[inc],
         extra_cflags=["-O2", f"-I{inc}"],
         extra_cuda_cflags=["-O2", f"-I{inc}", "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=f"{out}/build", verbose=False, is_python_module=False)
so = glob.glob(f"{out}/build/_ext*.so")[0]
shutil.copy(so, f"{out}/pointnet2_ref/_ext.so")
print("built", f"{out}/pointnet2_ref/_ext.so")
PY
