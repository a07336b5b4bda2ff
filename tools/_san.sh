export PATH=/usr/local/cuda/bin:$PATH
timeout 280 compute-sanitizer --tool memcheck --print-limit 5 python tools/debug_deep.py 20000 110 > gpurun_out/san_mem.log 2>&1; echo memcheck rc=$?
timeout 280 compute-sanitizer --tool racecheck --print-limit 5 python tools/debug_deep.py 6000 110 > gpurun_out/san_race.log 2>&1; echo racecheck rc=$?
timeout 280 compute-sanitizer --tool initcheck --print-limit 5 python tools/debug_deep.py 20000 110 > gpurun_out/san_init.log 2>&1; echo initcheck rc=$?
timeout 200 compute-sanitizer --tool synccheck --print-limit 5 python tools/debug_deep.py 20000 110 > gpurun_out/san_sync.log 2>&1; echo synccheck rc=$?
