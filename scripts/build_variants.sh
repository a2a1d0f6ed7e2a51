# build experiment variants of libpicca_b200.so into gpurun_variants/ (git-ignored, shipped by gpurun)
# usage: scripts/build_variants.sh tag1 "flags1" tag2 "flags2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_variants
while [ $# -gt 1 ]; do
  tag=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -shared \
       -I include -I picca_b200/csrc $flags -o gpurun_variants/lib_$tag.so picca_b200/csrc/*.cu 2>&1 | grep -v "warning\|Remark\|dlj\|^$\|\^" || true
  echo built gpurun_variants/lib_$tag.so "($flags)"
done
