#!/bin/bash
# ThreadSanitizer pass over the host-threads entropy backend: fb_host_entropy.cpp rebuilt with -fsanitize=thread, linked with the
# (uninstrumented) rest of the library, driven through fb_host_decode with the group index on many threads.  No GPU needed.
# usage: tools/tsan_host_entropy.sh file.fuif [repetitions] [threads]        (clean = prints "ok, N groups" and nothing else)
set -e
C=$(dirname "$0")/../fuif_b200/csrc
T=$(mktemp -d)
make -C "$C" > /dev/null
g++ -O1 -g -std=c++17 -fPIC -fsanitize=thread -pthread -c "$C/fb_host_entropy.cpp" -o "$T/ts.o"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$T/ts.so" "$C/fb_image.o" "$C/fb_transforms.o" "$C/fb_maniac.o" "$C/fb_maniac_enc.o" \
    "$T/ts.o" -Xcompiler -fPIC -Xcompiler -pthread -Xcompiler -fsanitize=thread
g++ -O1 -g -fsanitize=thread "$(dirname "$0")/tsan_host_entropy.cpp" -o "$T/tsan" -ldl
TSAN_OPTIONS="halt_on_error=0" "$T/tsan" "$T/ts.so" "$1" "${2:-3}" "${3:-16}"
rm -rf "$T"
