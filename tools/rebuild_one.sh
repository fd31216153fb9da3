#!/bin/bash
# tools/rebuild_one.sh bicgstab.cu [...]: recompile the given csrc files with the flags of sprsolve_b200/build.py and relink the library
set -e
cd "$(dirname "$0")/.."
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr -Wno-deprecated-gpu-targets -ccbin /usr/bin/g++"
for f in "$@"; do
  X=""; [ "$f" = spmv.cu ] && X="-Xptxas=${SPB_SPMV_PTXAS:--O1}"
  nvcc $X $F $EXTRA -c sprsolve_b200/csrc/$f -o sprsolve_b200/build/${f%.cu}.o &
done
wait
python - <<'P'
import os, subprocess
from sprsolve_b200 import build as b
objs = [os.path.join(b.OBJ, s.replace('.cu', '.o')) for s in b.SOURCES]
cuda_lib = '/usr/local/cuda/lib64'
r = subprocess.run(['/usr/bin/g++', '-shared', '-fPIC', '-o', b.LIB, *objs, f'-L{cuda_lib}', f'-Wl,-rpath,{cuda_lib}', '-lcudart', '-ldl', '-lpthread'], capture_output=True, text=True)
print("link", r.returncode, r.stderr[-500:])
P
