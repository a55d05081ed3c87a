#!/usr/bin/env bash
# tools/build_variant.sh NAME FILE "-DFLAG=.. ..."  ->  variants/NAME.so : the library with FILE.cu recompiled under extra flags
# (A/B kernel experiments on the GPU box: copy variants/NAME.so over newtonnet_b200/lib/libnewtonnet_b200.so between runs)
set -euo pipefail
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
name="$1"; file="$2"; flags="$3"
src="$root/newtonnet_b200/csrc"
mkdir -p "$root/variants"
objs=()
for f in nbr gemm_simt gemm_tc gemm_ts gemm_chain gemm_tn_tc message_tc pair_ops eval train_ops p2p md_ops; do
  if [ "$f" == "$file" ]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
      --expt-relaxed-constexpr -DNN_BUILD $flags -c "$src/$f.cu" -o "$root/variants/$name.$f.o"
    objs+=("$root/variants/$name.$f.o")
  else
    objs+=("$src/obj/$f.o")
  fi
done
/usr/local/cuda/bin/nvcc -Wno-deprecated-gpu-targets -shared -o "$root/variants/$name.so" "${objs[@]}" -lcudart
echo "built variants/$name.so"
