# ASan + UBSan over the product's contact code built for the host (no GPU needed)
set -e
cd "$(dirname "$0")/.."
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -shared -fPIC -w \
    tests/host_contacts_shim.cpp -o /tmp/libhc_asan.so
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 \
    HC_SO=/tmp/libhc_asan.so python tools/contacts_asan.py
