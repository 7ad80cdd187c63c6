import os, sys, time
sys.path.insert(0, '/root/repo')
import xsbench_b200 as xs
inp = xs.read_CLI(["-s","large","-m","event","-G","unionized","-k","6"])
sd = xs.materials_only(inp)
for cfg in ("1","0","1","0"):
    os.environ["XSB200_LAUNCH_CACHE"]=cfg
    gpu = xs.move_simulation_data_to_device(inp, sd)
    for _ in range(5): gpu.run(inp)
    t0=time.perf_counter(); dev=0.0
    for _ in range(40):
        r=gpu.run(inp); dev+=r.device_seconds
    dt=(time.perf_counter()-t0)/40
    print(f"LAUNCH_CACHE={cfg}: wall {1e3*dt:.3f} ms/step, device {1e3*dev/40:.3f} ms, gap {1e3*(dt-dev/40):.3f} ms, host_seconds {1e3*r.host_seconds:.3f}", flush=True)
    gpu.release()
