# every probe case in its own process (a rejected descriptor faults the whole CUDA context)
n=$(./build_probe/probe_umma count)
for i in $(seq 0 $((n-1))); do timeout 60 ./build_probe/probe_umma $i || true; done
timeout 60 ./build_probe/probe_umma split
