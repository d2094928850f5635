"""How the CPU arm (oracle port, pthreads) scales on this host: frames/s at 1..N threads + cgroup CPU quota."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
from oracle import pyoracle as po
pkg_synth = g._load_pkg().synth
fr = np.concatenate([pkg_synth.stream_frames(640, 480, s, 16)[0] for s in range(8)], axis=0)
fr = np.concatenate([fr, fr], axis=0)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(p):
        print(p, open(p).read().strip())
pa = np.arange(len(fr), dtype=np.int32); pb = ((pa + 1) % len(fr)).astype(np.int32)
po.extract_match_batch(fr[:8], pa[:8], pb[:8] % 8, 1000, 8)
for t in (1, 8, 16, 32, 64, 128):
    if t > (os.cpu_count() or 1):
        break
    n = min(len(fr), max(8, 4 * t))
    t0 = time.perf_counter(); po.extract_match_batch(fr[:n], pa[:n], (pa[:n] + 1) % n, 1000, t); dt = time.perf_counter() - t0
    print("threads %3d: %d frames in %.2f s -> %.1f fps" % (t, n, dt, n / dt))
